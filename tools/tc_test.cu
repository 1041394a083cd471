// tc_test.cu — standalone validation of the tcgen05 (kind::tf32) conventions used by kernels_tc.cuh:
// no-swizzle canonical shared-memory layouts (K-major and MN-major), descriptors, TMEM alloc/ld, and
// the 3xTF32 split.  nvcc -gencode arch=compute_100a,code=sm_100a -o tc_test tc_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                 // c_format = F32
  d |= 2u << 7;                 // a_format = TF32
  d |= 2u << 10;                // b_format = TF32
  d |= (uint32_t)a_mn_major << 15;
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, int accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}

// element offset (in floats) of (row, k) in a no-swizzle K-major tile: 8-row × 16-byte core matrices,
// k-chunks (4 floats) adjacent (LBO = 128 B), 8-row groups SBO apart
__host__ __device__ inline int kmajor_off(int row, int k, int Ktile) {
  return (row % 8) * 4 + (k % 4) + (k / 4) * 32 + (row / 8) * (Ktile / 4) * 32;
}
// (mn, k) in a no-swizzle MN-major tile: 8(k) × 16-byte (4 mn) core matrices; mn-groups adjacent
// (SBO = 128 B), k-groups of 8 LBO apart
__host__ __device__ inline int mnmajor_off(int mn, int k, int MNtile) {
  return (mn % 4) + (k % 8) * 4 + (mn / 4) * 32 + (k / 8) * (MNtile / 4) * 32;
}

template <int M, int N, int K, int AMN, int BMN, int SPLIT>
__global__ void __launch_bounds__(128) test_kernel(const float* A, const float* B, float* D) {
  // A: M×K row-major (K contiguous) ; B: N×K row-major.  Computes D = A·Bᵀ (M×N row-major).
  extern __shared__ __align__(1024) float smem[];
  float* sAh = smem;
  float* sAl = sAh + M * K;
  float* sBh = sAl + M * K;
  float* sBl = sBh + N * K;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    const float x = A[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int off = AMN ? mnmajor_off(r, k, M) : kmajor_off(r, k, K);
    sAh[off] = SPLIT ? hi : x;
    sAl[off] = x - hi;
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    const float x = B[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int off = BMN ? mnmajor_off(r, k, N) : kmajor_off(r, k, K);
    sBh[off] = SPLIT ? hi : x;
    sBl[off] = x - hi;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes → async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N, AMN, BMN);
    int acc = 0;
    for (int term = 0; term < (SPLIT ? 3 : 1); ++term) {
      const float* a = (term == 2) ? sAl : sAh;
      const float* b = (term == 1) ? sBl : sBh;
      for (int k0 = 0; k0 < K; k0 += 8) {
        uint64_t ad, bd;
        if (AMN) ad = make_desc(smem_u32(a) + (k0 / 8) * (M / 4) * 128, (M / 4) * 128, 128);
        else ad = make_desc(smem_u32(a) + (k0 / 4) * 128, 128, (K / 4) * 128);
        if (BMN) bd = make_desc(smem_u32(b) + (k0 / 8) * (N / 4) * 128, (N / 4) * 128, 128);
        else bd = make_desc(smem_u32(b) + (k0 / 4) * 128, 128, (K / 4) * 128);
        mma_tf32(tmem, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)));
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&s_bar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // epilogue: warp w reads lanes 32w..32w+31 (M=128) — for M=64 rows 16w..16w+15 sit in lanes 32w..32w+15
  uint32_t v[64];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
        "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int lane = tid & 31;
  int row;
  if (M == 128) row = warp * 32 + lane;
  else row = (lane < 16) ? warp * 16 + lane : -1;
  if (row >= 0)
    for (int j = 0; j < N && j < 64; ++j) D[row * N + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

template <int M, int N, int K, int AMN, int BMN, int SPLIT>
void run(const char* name) {
  std::vector<float> A(M * K), B(N * K), D(M * N, -1.f);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto& x : B) x = (float)rand() / RAND_MAX - 0.5f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, D.size() * 4);
  const int smem = (2 * M * K + 2 * N * K) * 4;
  auto kern = test_kernel<M, N, K, AMN, BMN, SPLIT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  kern<<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * (double)B[j * K + k];
      maxerr = fmax(maxerr, fabs(s - (double)D[i * N + j]));
      maxref = fmax(maxref, fabs(s));
    }
  printf("%-40s M=%d N=%d K=%d AMN=%d BMN=%d split=%d : %s  max|err| = %.3e (max|ref| %.3f)\n", name, M, N, K, AMN, BMN,
         SPLIT, cudaGetErrorString(e), maxerr, maxref);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
  run<128, 64, 32, 0, 0, 0>("K-major A, K-major B, 1xTF32");
  run<128, 64, 32, 0, 0, 1>("K-major A, K-major B, 3xTF32");
  run<128, 64, 32, 1, 0, 1>("MN-major A, K-major B, 3xTF32");
  run<128, 64, 32, 1, 1, 1>("MN-major A, MN-major B, 3xTF32");
  run<128, 64, 64, 0, 0, 1>("K-major, K=64, 3xTF32");
  run<64, 64, 32, 0, 0, 1>("M=64 K-major, 3xTF32");
  run<64, 32, 64, 0, 0, 1>("M=64 N=32 K=64, 3xTF32");
  run<128, 32, 32, 1, 0, 1>("MN-major A N=32, 3xTF32");
  return 0;
}
