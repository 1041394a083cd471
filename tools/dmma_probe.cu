// dmma_probe.cu — how the fp64 tensor-core rate of one SM depends on the number of resident warps, the number of
// independent accumulators per warp and the MMA shape (the Hermitian Gram of the simple update, kernels_dmma.cuh, runs one
// CTA of 10 warps per SM).  One CTA per SM (forced by a large dynamic shared-memory request), registers only.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_probe tools/dmma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int NA, int NB>
__global__ void __launch_bounds__(1024) k884(double* out, int iters) {
  double c[NA][NB][2];
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  double a[NA], b[NB];
#pragma unroll
  for (int i = 0; i < NA; ++i) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int j = 0; j < NB; ++j) b[j] = 1.0 + threadIdx.x * 1e-4 + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) s += c[i][j][0] + c[i][j][1];
  if (s == 123.456) out[0] = s;
}

// m16n8k8: A 16×8 (4 doubles per lane), B 8×8 (2 per lane), C 16×8 (4 per lane)
template <int NA, int NB>
__global__ void __launch_bounds__(1024) k1688(double* out, int iters) {
  double c[NA][NB][4];
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[i][j][q] = 0.0;
  double a[NA][4], b[NB][2];
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) a[i][q] = threadIdx.x * 1e-3 + i + q;
#pragma unroll
  for (int j = 0; j < NB; ++j) { b[j][0] = 1.0 + threadIdx.x * 1e-4 + j; b[j][1] = 0.5 + j; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[i][j][0]), "+d"(c[i][j][1]), "+d"(c[i][j][2]), "+d"(c[i][j][3])
                     : "d"(a[i][0]), "d"(a[i][1]), "d"(a[i][2]), "d"(a[i][3]), "d"(b[j][0]), "d"(b[j][1]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) s += c[i][j][0] + c[i][j][1] + c[i][j][2] + c[i][j][3];
  if (s == 123.456) out[0] = s;
}

template <class F> static float time_ms(F f, int rep = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < rep; ++r) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms / rep;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* dd;
  cudaMalloc(&dd, 64);
  const size_t smem = 120 * 1024;  // one CTA per SM
  cudaFuncSetAttribute(k884<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k884<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k884<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k1688<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k1688<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 4000;
  const int ws[] = {4, 8, 10, 12, 16, 20, 24, 32};
  printf("fp64 tensor-core rate, one CTA per SM (%d SMs), TFLOP/s\n", sms);
  printf("%-28s", "warps per SM");
  for (int w : ws) printf(" %6d", w);
  printf("\n");
  auto row = [&](const char* name, auto launch, double fl_per_warp_iter) {
    printf("%-28s", name);
    for (int w : ws) {
      const float ms = time_ms([&] { launch(w); });
      printf(" %6.1f", fl_per_warp_iter * iters * w * sms / ms / 1e9);
    }
    printf("\n");
  };
  row("m8n8k4  16 acc (4x4)", [&](int w) { k884<4, 4><<<sms, w * 32, smem>>>(dd, iters); }, 512.0 * 16);
  row("m8n8k4   8 acc (4x2)", [&](int w) { k884<4, 2><<<sms, w * 32, smem>>>(dd, iters); }, 512.0 * 8);
  row("m8n8k4   4 acc (2x2)", [&](int w) { k884<2, 2><<<sms, w * 32, smem>>>(dd, iters); }, 512.0 * 4);
  row("m16n8k8  8 acc tiles (2x4)", [&](int w) { k1688<2, 4><<<sms, w * 32, smem>>>(dd, iters); }, 2048.0 * 8);
  row("m16n8k8  2 acc tiles (1x2)", [&](int w) { k1688<1, 2><<<sms, w * 32, smem>>>(dd, iters); }, 2048.0 * 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
