// tc2g_test.cu — stand-alone correctness + bandwidth test of the TMA-fed tcgen05 Gram contraction (kernels_tc2g.cuh)
// against a double-precision CPU evaluation of  G[i][j] = Σ_col conj(X[i][col])·Y[j][col].
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc2g_test tc2g_test.cu ; ./tc2g_test [bench]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../tensornetworkquantumsimulator.jl_b200/csrc/kernels_tc2g.cuh"

using namespace tnqs;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Case { const char* name; unsigned outer; int chi; unsigned inner; };

// `batch` independent (X, Y) pairs of the same shape (the first pair random, the others copies of it)
static bool run_case(const Case& c, int batch, bool timing) {
  const long long CC = (long long)c.outer * c.inner;
  const long long n = CC * c.chi;
  std::mt19937_64 rng(4321);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float2> hx((size_t)n), hy((size_t)n);
  for (auto& x : hx) { x.x = nd(rng); x.y = nd(rng); }
  // Y correlated with X so that G has a dominant diagonal like a BP message
  for (size_t i = 0; i < hy.size(); ++i) { hy[i].x = 0.7f * hx[i].x + 0.5f * nd(rng); hy[i].y = 0.7f * hx[i].y + 0.5f * nd(rng); }
  float2 *dx, *dy;
  CK(cudaMalloc(&dx, (size_t)n * 8 * batch));
  CK(cudaMalloc(&dy, (size_t)n * 8 * batch));
  for (int b = 0; b < batch; ++b) {
    CK(cudaMemcpy(dx + (size_t)b * n, hx.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dy + (size_t)b * n, hy.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  }
  tc2g::GPlan plan;
  bool ok = true;
  for (int b = 0; b < batch; ++b) {
    tc2g::GramShape t{dx + (size_t)b * n, dy + (size_t)b * n, c.chi, c.outer, c.inner, (unsigned)CC};
    ok = ok && tc2g::plan_add(plan, t, b);
  }
  if (!ok) { printf("%-34s : NOT ELIGIBLE for the TMA path\n", c.name); return false; }
  tc2g::plan_finish(plan);
  tc2g::GLaunch& L = plan.launches[0];
  L.gm.dbg = tc2g::env_int("TNQS_TC2G_DBG", 0);
  L.gm.nstage = std::min(L.gm.nstage, tc2g::env_int("TNQS_TC2G_NSTAGE", 99));
  const size_t per = (size_t)c.chi * c.chi;
  std::vector<double2*> parts(batch);
  for (int b = 0; b < batch; ++b) {
    CK(cudaMalloc(&parts[b], (size_t)L.nslots[b] * per * sizeof(double2)));
    CK(cudaMemset(parts[b], 0xFF, (size_t)L.nslots[b] * per * sizeof(double2)));  // NaN pattern: unwritten partials are detected
    L.tasks[b].partial = parts[b];
  }
  tc2g::GramTask2* dt;
  tc2g::GItem* di;
  CK(cudaMalloc(&dt, L.tasks.size() * sizeof(tc2g::GramTask2)));
  CK(cudaMemcpy(dt, L.tasks.data(), L.tasks.size() * sizeof(tc2g::GramTask2), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&di, L.items.size() * sizeof(tc2g::GItem)));
  CK(cudaMemcpy(di, L.items.data(), L.items.size() * sizeof(tc2g::GItem), cudaMemcpyHostToDevice));
  auto launch = [&] {
    if (L.last) tc2g::tc2_gram_kernel<true><<<L.grid, tc2g::G_THREADS, L.smem>>>(dt, di, (int)L.items.size(), L.gm);
    else tc2g::tc2_gram_kernel<false><<<L.grid, tc2g::G_THREADS, L.smem>>>(dt, di, (int)L.items.size(), L.gm);
  };
  launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-34s : KERNEL ERROR %s\n", c.name, cudaGetErrorString(e)); exit(3); }
  // reduce the partial slots on the host (first and last copy)
  auto reduce = [&](int b) {
    std::vector<double2> hp((size_t)L.nslots[b] * per);
    CK(cudaMemcpy(hp.data(), parts[b], hp.size() * sizeof(double2), cudaMemcpyDeviceToHost));
    std::vector<std::complex<double>> g(per, 0.0);
    for (int s = 0; s < L.nslots[b]; ++s)
      for (size_t k = 0; k < per; ++k) g[k] += std::complex<double>(hp[(size_t)s * per + k].x, hp[(size_t)s * per + k].y);
    return g;
  };
  const auto g0 = reduce(0), gl = reduce(batch - 1);
  long long nan_count = 0, diff_last = 0;
  for (size_t k = 0; k < per; ++k) {
    if (std::isnan(g0[k].real()) || std::isnan(g0[k].imag())) ++nan_count;
    if (g0[k] != gl[k]) ++diff_last;
  }
  // sampled entries against a double-precision evaluation; scale = a typical diagonal entry
  const int nsamp = (double)per * CC > 3e8 ? std::max(8, (int)(3e8 / CC)) : (int)per;
  std::uniform_int_distribution<int> pick(0, (int)per - 1);
  double maxerr = 0, scale = 0;
  for (int sidx = 0; sidx < nsamp; ++sidx) {
    const int k = nsamp == (int)per ? sidx : pick(rng);
    const int i = k / c.chi, j = k % c.chi;
    std::complex<double> acc = 0;
    for (long long o = 0; o < c.outer; ++o)
      for (long long q = 0; q < c.inner; ++q) {
        const float2 x = hx[(size_t)((o * c.chi + i) * c.inner + q)], y = hy[(size_t)((o * c.chi + j) * c.inner + q)];
        acc += std::conj(std::complex<double>(x.x, x.y)) * std::complex<double>(y.x, y.y);
      }
    if (!std::isnan(g0[k].real())) maxerr = std::max(maxerr, std::abs(g0[k] - acc));
    if (i == j) scale = std::max(scale, std::abs(acc));
  }
  if (scale == 0) scale = 1.4 * CC;  // E|x|² · 0.7 · CC
  const bool pass = nan_count == 0 && diff_last == 0 && maxerr / scale < 1e-5;
  const tc2g::GGeom& gm = L.gm;
  printf("%-34s : %s  max|err|/diag = %.2e  unwritten = %lld  copy-mismatch = %lld  [%s chi %d stacked %d nb %d kch %d nstage %d flush %d pfd %d stage %u ncol %d items %zu grid %d slots %d smem %zu]\n",
         c.name, pass ? "ok  " : "FAIL", maxerr / scale, nan_count, diff_last, L.last ? "LAST" : "MID", gm.chi, gm.stacked, gm.nb, gm.kch, gm.nstage, gm.flush, gm.pfd, gm.stage,
         gm.ncol, L.items.size(), L.grid, L.nslots[0], L.smem);
  if (timing) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(a); launch(); cudaEventRecord(b);
      CK(cudaEventSynchronize(b));
      float ms; cudaEventElapsedTime(&ms, a, b);
      best = std::min(best, ms);
    }
    const double bytes = 2.0 * 8.0 * (double)n * batch;
    printf("%-34s   batch %d: %.3f ms  -> %.0f GB/s algorithmic (X+Y %.2f GB), %.1f TFLOP/s\n", "", batch, best, bytes / best / 1e6, bytes / 1e9,
           8.0 * c.chi * c.chi * (double)CC * batch / best / 1e9);
  }
  for (auto p : parts) cudaFree(p);
  cudaFree(dt); cudaFree(di); cudaFree(dx); cudaFree(dy);
  return pass;
}

int main(int argc, char** argv) {
  const bool bench = argc > 1 && !strcmp(argv[1], "bench");
  CK(cudaFuncSetAttribute(tc2g::tc2_gram_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  CK(cudaFuncSetAttribute(tc2g::tc2_gram_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  if (argc > 1 && !strcmp(argv[1], "quick")) {  // bandwidth of the two main shapes only (parameter sweeps through the TNQS_TC2G_* variables)
    run_case({"bench MID chi32 leg1 x48", 2 * 32, 32, 1024}, 48, true);
    run_case({"bench LAST chi32 x48", 2 * 32 * 32 * 32, 32, 1}, 48, true);
    run_case({"bench MID chi64 leg1 x4", 2 * 64, 64, 4096}, 4, true);
    run_case({"bench LAST chi64 x4", 2 * 64 * 64 * 64, 64, 1}, 4, true);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "prof")) {  // one launch each for ncu
    run_case({"prof MID chi32 leg1 x48", 2 * 32, 32, 1024}, 48, false);
    run_case({"prof LAST chi32 x48", 2 * 32 * 32 * 32, 32, 1}, 48, false);
    run_case({"prof MID chi64 leg1 x4", 2 * 64, 64, 4096}, 4, false);
    run_case({"prof LAST chi64 x4", 2 * 64 * 64 * 64, 64, 1}, 4, false);
    return 0;
  }
  const Case cases[] = {
      {"MID chi32 leg2 (inner 32)", 2 * 32 * 32, 32, 32},
      {"MID chi32 leg1 (inner 1024)", 2 * 32, 32, 1024},
      {"MID chi32 leg0 (inner 32768)", 2, 32, 32768},
      {"MID chi16 (inner 16)", 2 * 16 * 16 * 4, 16, 16},
      {"MID chi16 (inner 256)", 2 * 16 * 4, 16, 256},
      {"MID chi64 (inner 64)", 2 * 64, 64, 64},
      {"MID chi64 (inner 4096)", 2, 64, 4096},
      {"MID chi48 (inner 48)", 70, 48, 48},
      {"MID chi24 (inner 48)", 210, 24, 48},
      {"MID ragged units CC=3360", 70, 32, 48},
      {"LAST chi32", 2 * 32 * 32 * 8, 32, 1},
      {"LAST chi16", 2 * 16 * 16 * 16, 16, 1},
      {"LAST chi64", 2 * 64 * 64, 64, 1},
      {"LAST chi48", 5000, 48, 1},
      {"LAST ragged CC=3001", 3001, 32, 1},
  };
  int nfail = 0;
  for (auto& c : cases) nfail += run_case(c, 2, false) ? 0 : 1;
  printf("%d case(s) failed\n", nfail);
  if (bench) {
    // 48 interior chi=32 site tensors (16.8 MB each; X + Y = 1.6 GB, beyond the 126 MB L2)
    run_case({"bench MID chi32 leg1 x48", 2 * 32, 32, 1024}, 48, true);
    run_case({"bench MID chi32 leg2 x48", 2 * 32 * 32, 32, 32}, 48, true);
    run_case({"bench MID chi32 leg0 x48", 2, 32, 32768}, 48, true);
    run_case({"bench LAST chi32 x48", 2 * 32 * 32 * 32, 32, 1}, 48, true);
    run_case({"bench MID chi64 leg1 x4", 2 * 64, 64, 4096}, 4, true);
    run_case({"bench MID chi64 leg0 x4", 2, 64, 262144}, 4, true);
    run_case({"bench LAST chi64 x4", 2 * 64 * 64 * 64, 64, 1}, 4, true);
    run_case({"bench MID chi16 z6 x4", 2 * 16 * 16, 16, 4096}, 16, true);
  }
  return nfail;
}
