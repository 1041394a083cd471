// tma_bw.cu — load-only bandwidth of the TMA unit for the box shapes the mode-product kernel can use.
// One persistent CTA per SM, a ring of 16 KB stages, one producer thread, one consumer thread (no compute).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
// mode 0: cp.async.bulk of 16 KB contiguous.  mode 1: `nbox` tensor boxes per stage; box b of stage t at coordinates
// (c0 = (b * bx0) % dim0 ... supplied by the host through `cstep`): generic: x = ((t*nbox+b) * bx0) % dim0, y = ((t*nbox+b) * bx0 / dim0) * by
__global__ void __launch_bounds__(64) bw_kernel(const CUtensorMap* tm, const float* src, int mode, int nbox, int bx0, int by, int dim0,
                                                uint32_t box_bytes, long long ntiles, int nstage) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[16], empty[16];
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t stage_bytes = 16384;
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(stage_bytes) : "memory");
      const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
      if (mode == 0) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src + t * 4096), "r"(stage_bytes), "r"(smem_u32(&full[s])) : "memory");
      } else {
        for (int b = 0; b < nbox; ++b) {
          const long long lin = (t * nbox + b) * (long long)bx0;
          const int x = (int)(lin % dim0), y = (int)(lin / dim0) * by;
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(dst + b * box_bytes), "l"(tm), "r"(x), "r"(y), "r"(smem_u32(&full[s])) : "memory");
        }
      }
      if (++s == nstage) { s = 0; ph ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(smem_u32(&full[s]), ph);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      if (++s == nstage) { s = 0; ph ^= 1u; }
    }
  }
}
int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiled enc = (EncodeTiled)fn;
  const size_t bytes = 1ull << 30;  // 1 GiB source
  float* src; CK(cudaMalloc(&src, bytes)); CK(cudaMemset(src, 0, bytes));
  CUtensorMap* dmap; CK(cudaMalloc(&dmap, sizeof(CUtensorMap)));
  CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const long long ntiles = bytes / 16384;
  struct V { const char* name; int mode; int dim0 /*floats per row*/; int bx0, by; CUtensorMapSwizzle sw; };
  const V vs[] = {
      {"cp.async.bulk 16 KB contiguous", 0, 0, 0, 0, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"box 32f x 128 rows, rows 128 B apart (contiguous), SW128", 1, 32, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B},
      {"box 32f x 128 rows, rows 256 B apart (LAST chi32), SW128", 1, 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B},
      {"box 32f x 32 rows, rows 8 KB apart (MID inner 1024), SW128_ATOM32", 1, 2048, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B},
      {"box 32f x 32 rows, rows 8 KB apart, SW128", 1, 2048, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B},
      {"box 32f x 32 rows, rows 8 KB apart, no swizzle", 1, 2048, 32, 32, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"box 128f x 32 rows, rows 8 KB apart, no swizzle", 1, 2048, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"box 256f x 16 rows, rows 8 KB apart, no swizzle", 1, 2048, 256, 16, CU_TENSOR_MAP_SWIZZLE_NONE},
  };
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  for (auto& v : vs) {
    int nbox = 1; uint32_t box_bytes = 16384;
    if (v.mode == 1) {
      CUtensorMap tm;
      const cuuint64_t rows = bytes / 4 / v.dim0;
      cuuint64_t gd[2] = {(cuuint64_t)v.dim0, rows};
      cuuint64_t gs[1] = {(cuuint64_t)v.dim0 * 4};
      cuuint32_t box[2] = {(cuuint32_t)v.bx0, (cuuint32_t)v.by};
      cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("%-70s : encode failed %d\n", v.name, (int)r); continue; }
      CK(cudaMemcpy(dmap, &tm, sizeof(tm), cudaMemcpyHostToDevice));
      box_bytes = (uint32_t)v.bx0 * v.by * 4;
      nbox = 16384 / box_bytes;
    }
    for (int nstage : {4, 8, 12}) {
      for (int grid : {prop.multiProcessorCount, 2 * prop.multiProcessorCount}) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(a);
          bw_kernel<<<grid, 64, nstage * 16384>>>(dmap, src, v.mode, nbox, v.bx0, v.by, v.dim0, box_bytes, ntiles, nstage);
          cudaEventRecord(b);
          CK(cudaEventSynchronize(b));
          float ms; cudaEventElapsedTime(&ms, a, b);
          if (ms < best) best = ms;
        }
        printf("%-70s stages %2d grid %3d : %.3f ms  %.0f GB/s\n", v.name, nstage, grid, best, bytes / best / 1e6);
      }
    }
  }
  return 0;
}
