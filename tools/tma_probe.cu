// tma_probe.cu — stand-alone probes run once on the B200 box (results quoted in DESIGN.md / profiles/):
//  1. shared-memory image of a TMA (cp.async.bulk.tensor) load for the swizzle modes the tcgen05 TF32 operands need
//     (SWIZZLE_128B for K-major tiles, SWIZZLE_128B_ATOM_32B for MN-major tiles), tensor map read from GLOBAL memory;
//  2. whether tcgen05 kind::tf32 truncates or rounds the low 13 mantissa bits of an fp32 operand;
//  3. cuBLAS-free peak micro-benchmarks that the rooflines of this repo use as denominators:
//     tcgen05 kind::tf32 dense MMA, fp32 FFMA, fp64 DFMA and fp64 DMMA (mma.sync.m8n8k4.f64).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) { printf("cuTensorMapEncodeTiled not found\n"); exit(1); }
  return (EncodeTiled)fn;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// ---- 1. TMA image ------------------------------------------------------------------------------------
// 3-D tensor [d2][d1][d0 floats]; loads box (b0,b1,b2) at (c0,c1,c2) and copies the shared-memory image out.
__global__ void tma_dump(const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bytes, float* out) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) smem[i] = -1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = smem[i];
}

// ---- 2. tf32 operand rounding ----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, int acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major no-swizzle [rows][8 floats]: (row%8)*4 + k%4 + (k/4)*32 + (row/8)*64
__device__ __forceinline__ int km8(int row, int k) { return (row % 8) * 4 + (k % 4) + (k / 4) * 32 + (row / 8) * 64; }
// D[m][n] = a[m] * (n == 0): A(m, k=0) = a[m], B(n=0,k=0) = 1.  Output: the value the tensor core used for a[m].
__global__ void __launch_bounds__(128) tf32_round_probe(const float* a, float* out) {
  __shared__ __align__(1024) float sA[128 * 8];
  __shared__ __align__(1024) float sB[16 * 8];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 8; i += 128) sA[i] = 0.f;
  for (int i = tid; i < 16 * 8; i += 128) sB[i] = 0.f;
  __syncthreads();
  sA[km8(tid, 0)] = a[tid];
  if (tid == 0) sB[km8(0, 0)] = 1.f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    mma_tf32(tmem, make_desc(smem_u32(sA), 128, 256, 0), make_desc(smem_u32(sB), 128, 256, 0), make_idesc(128, 16), 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[tid] = __uint_as_float(v[0]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

// ---- 3. peaks ----------------------------------------------------------------------------------------
// tcgen05 kind::tf32: every CTA issues `iters` × 4 MMAs of M=128, N=256, K=8 on resident (zero) shared-memory tiles
__global__ void __launch_bounds__(128) tf32_peak(int iters) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 32; i += 128) smem[i] = 0.f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 256);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 32);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_tf32(tmem, make_desc(a0 + ks * 256, 128, 1024, 0), make_desc(b0 + ks * 256, 128, 1024, 0), idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
__global__ void __launch_bounds__(256) ffma_peak(float* out, int iters) {
  float a[8], b = 1.0001f, c = 0.5f;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;
}
__global__ void __launch_bounds__(256) dfma_peak(double* out, int iters) {
  double a[8], b = 1.0001, c = 0.5;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) dmma_peak(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <class F> static float time_ms(F f, int rep = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < rep; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  EncodeTiled enc = get_encode();
  // ---------------- 1. TMA images ----------------
  const int D0 = 256, D1 = 64, D2 = 4;  // floats, rows, slices
  std::vector<float> h((size_t)D0 * D1 * D2);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *dsrc, *dout;
  CK(cudaMalloc(&dsrc, h.size() * 4));
  CK(cudaMalloc(&dout, 64 * 1024));
  CK(cudaMemcpy(dsrc, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap* dmap;
  CK(cudaMalloc(&dmap, sizeof(CUtensorMap)));
  CK(cudaFuncSetAttribute(tma_dump, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  struct Cfg { CUtensorMapSwizzle sw; const char* name; } cfgs[] = {
      {CU_TENSOR_MAP_SWIZZLE_NONE, "NONE"}, {CU_TENSOR_MAP_SWIZZLE_128B, "128B"}, {CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "128B_ATOM_32B"}};
  for (auto& cfg : cfgs) {
    CUtensorMap tm;
    cuuint64_t gdim[3] = {D0, D1, D2};
    cuuint64_t gstr[2] = {(cuuint64_t)D0 * 4, (cuuint64_t)D0 * D1 * 4};
    cuuint32_t box[3] = {32, 16, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dsrc, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, cfg.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== TMA swizzle %s: encode result %d\n", cfg.name, (int)r);
    if (r != CUDA_SUCCESS) continue;
    CK(cudaMemcpy(dmap, &tm, sizeof(tm), cudaMemcpyHostToDevice));
    const uint32_t bytes = 32 * 16 * 4;
    CK(cudaMemset(dout, 0, 64 * 1024));
    tma_dump<<<1, 128, 16 * 1024>>>(dmap, 32, 8, 1, bytes, dout);  // box origin: float 32, row 8, slice 1
    cudaError_t e = cudaDeviceSynchronize();
    printf("   kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(bytes / 4);
    CK(cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost));
    // decode: every loaded value encodes (slice, row, float); print for each 128-byte smem row the 16-byte chunk order
    int bad_none = 0, bad_128 = 0, bad_a32 = 0;
    for (int row = 0; row < 16; ++row) {
      if (row < 8) printf("   smem row %2d: chunks(src 16B-chunk index within the row):", row);
      for (int ch = 0; ch < 8; ++ch) {
        const float v = o[row * 32 + ch * 4];
        const long idx = (long)v;
        const int f = (int)(idx % D0), rr = (int)((idx / D0) % D1);
        const int src_chunk = (f - 32) / 4, src_row = rr - 8;
        if (row < 8) printf(" r%dc%d", src_row, src_chunk);
        if (!(src_row == row && src_chunk == ch)) ++bad_none;
        if (!(src_row == row && src_chunk == (ch ^ (row & 7)))) ++bad_128;
        // ATOM_32B hypothesis: 32-byte chunk index (ch>>1) XOR (row & 3), 16-byte half kept
        if (!(src_row == row && src_chunk == ((((ch >> 1) ^ (row & 3)) << 1) | (ch & 1)))) ++bad_a32;
      }
      if (row < 8) printf("\n");
    }
    printf("   mismatches vs hypotheses: none=%d  sw128(16B chunk ^ row%%8)=%d  atom32(32B chunk ^ row%%4)=%d\n", bad_none, bad_128, bad_a32);
  }
  // permuted-stride 4-D map (dim1 stride > dim2 stride): accepted by the driver?
  {
    CUtensorMap tm;
    cuuint64_t gdim[4] = {32, D1, D0 / 32, D2};
    cuuint64_t gstr[3] = {(cuuint64_t)D0 * 4, 128, (cuuint64_t)D0 * D1 * 4};
    cuuint32_t box[4] = {32, 16, 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dsrc, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== permuted-stride 4-D map (32f, rows, 128B-blocks, slices): encode result %d\n", (int)r);
  }
  // ---------------- 2. tf32 rounding ----------------
  {
    std::vector<float> a(128), o(128);
    for (int i = 0; i < 128; ++i) {
      uint32_t bits = 0x3F800000u | ((uint32_t)(i + 1) << 10) | 0x1FFFu * (i & 1) | ((i & 2) ? 0x1000u : 0u);
      memcpy(&a[i], &bits, 4);
    }
    float *da, *dro;
    CK(cudaMalloc(&da, 512)); CK(cudaMalloc(&dro, 512));
    CK(cudaMemcpy(da, a.data(), 512, cudaMemcpyHostToDevice));
    tf32_round_probe<<<1, 128>>>(da, dro);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(o.data(), dro, 512, cudaMemcpyDeviceToHost));
    int ntrunc = 0, nrna = 0, nother = 0;
    for (int i = 0; i < 128; ++i) {
      uint32_t in, out; memcpy(&in, &a[i], 4); memcpy(&out, &o[i], 4);
      const uint32_t tr = in & 0xFFFFE000u;
      const uint32_t rn = (in + 0x1000u) & 0xFFFFE000u;  // round-to-nearest, ties away (cvt.rna)
      if (tr == rn) continue;  // indistinguishable
      if (out == tr) ++ntrunc; else if (out == rn) ++nrna; else ++nother;
    }
    printf("== tf32 operand handling of fp32 bits: truncated %d, rounded(rna) %d, other %d (of the distinguishable samples)\n", ntrunc, nrna, nother);
    for (int i = 0; i < 4; ++i) { uint32_t in, out; memcpy(&in, &a[i], 4); memcpy(&out, &o[i], 4); printf("   in %08x -> used %08x\n", in, out); }
  }
  // ---------------- 3. peaks ----------------
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  {
    const size_t sm = (128 + 256) * 32 * 4 + 1024;
    CK(cudaFuncSetAttribute(tf32_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int iters = 20000;
    const float ms = time_ms([&] { tf32_peak<<<sms, 128, sm>>>(iters); });
    const double fl = 2.0 * 128 * 256 * 8 * 4.0 * iters * sms;
    printf("== tcgen05 kind::tf32 dense M128 N256 K8, 1 CTA/SM x %d SMs: %.3f ms -> %.1f TFLOP/s\n", sms, ms, fl / ms / 1e9);
  }
  float* df; double* dd;
  CK(cudaMalloc(&df, 64)); CK(cudaMalloc(&dd, 64));
  {
    const int iters = 4000, blocks = sms * 8;
    const float ms = time_ms([&] { ffma_peak<<<blocks, 256>>>(df, iters); });
    const double fl = 2.0 * 64 * iters * 256.0 * blocks;
    printf("== fp32 FFMA: %.3f ms -> %.1f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  {
    const int iters = 1000, blocks = sms * 8;
    const float ms = time_ms([&] { dfma_peak<<<blocks, 256>>>(dd, iters); });
    const double fl = 2.0 * 64 * iters * 256.0 * blocks;
    printf("== fp64 DFMA: %.3f ms -> %.1f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  {
    const int iters = 2000, blocks = sms * 8;
    const float ms = time_ms([&] { dmma_peak<<<blocks, 256>>>(dd, iters); });
    const double fl = 2.0 * 8 * 8 * 4 * 8.0 * iters * 8.0 * blocks;  // per warp: 8 MMAs of m8n8k4 per iteration, 8 warps per block
    printf("== fp64 DMMA m8n8k4: %.3f ms -> %.1f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  return 0;
}
