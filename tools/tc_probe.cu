// Layout discovery for tcgen05 MN-major / no-swizzle: A smem tile holds its own float index, B = [I8;0]
// (K-major, known-good), so D[m][k] = index of the smem float the hardware used for A(m,k).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__global__ void __launch_bounds__(128) probe(float* D, int amn, uint32_t lbo, uint32_t sbo, uint32_t layout, int nfloats) {
  extern __shared__ __align__(1024) float smem[];
  float* sA = smem;               // nfloats
  float* sB = smem + 4096;        // 16x8 K-major: kmajor: (row%8)*4 + k%4 + (k/4)*32 + (row/8)*64
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4096; i += 128) sA[i] = (i < nfloats) ? (float)i : 0.f;
  for (int i = tid; i < 16 * 8; i += 128) {
    const int r = i / 8, k = i % 8;
    sB[(r % 8) * 4 + (k % 4) + (k / 4) * 32 + (r / 8) * 64] = (r == k) ? 1.f : 0.f;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amn << 15) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    uint64_t ad = make_desc(smem_u32(sA), lbo, sbo, layout);
    uint64_t bd = make_desc(smem_u32(sB), 128, 256, 0);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0));
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)));
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0));
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) D[tid * 16 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}
int main() {
  float* dD; cudaMalloc(&dD, 128 * 16 * 4);
  std::vector<float> D(128 * 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 4 + 1024);
  struct Cfg { int amn; uint32_t lbo, sbo, layout; const char* name; } cfgs[] = {
      {1, 1024, 512, 1, "MN sw128_32B lbo=1024 sbo=512"},
      {1, 512, 2048, 1, "MN sw128_32B lbo=512 sbo=2048"},
  };
  for (auto& c : cfgs) {
    cudaMemset(dD, 0, 128 * 16 * 4);
    probe<<<1, 128, 4096 * 4 + 1024>>>(dD, c.amn, c.lbo, c.sbo, c.layout, 4096);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    printf("== %s : %s\n", c.name, cudaGetErrorString(e));
    int ms[] = {0, 1, 2, 7, 8, 9, 15, 16, 17, 24, 31, 32, 33, 40, 64, 96, 127};
    for (int m : ms) {
      printf("  m=%3d :", m);
      for (int k = 0; k < 8; ++k) printf(" %6.0f", D[m * 16 + k]);
      printf("\n");
    }
  }
  return 0;
}
