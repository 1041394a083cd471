// tc2_test.cu — stand-alone correctness + bandwidth test of the TMA-fed tcgen05 mode product (kernels_tc2.cuh)
// against a double-precision CPU evaluation of  Out[p',o,c,n] = Σ_{p,b} In[p,o,b,n]·Mat[(p,b),(p',c)].
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc2_test tc2_test.cu ; ./tc2_test [bench]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <vector>

#include "../tensornetworkquantumsimulator.jl_b200/csrc/kernels_tc2.cuh"

using namespace tnqs;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Case { const char* name; int P_in, P_out; unsigned outer; int chi_in, chi_out; unsigned inner; };

static int g_stage_override = 0;
static bool g_instr = false;  // run the instrumented instantiation (debug switches / role counters)

// runs `batch` independent copies of the product (distinct tensors, same matrix); returns max error / rms over samples
static bool run_case(const Case& c, int batch, bool timing) {
  const long long CC = (long long)c.outer * c.inner;
  const long long in_plane = CC * c.chi_in, out_plane = CC * c.chi_out;
  const long long in_n = in_plane * c.P_in, out_n = out_plane * c.P_out;
  const int KK = c.P_in * c.chi_in, MM = c.P_out * c.chi_out;
  std::mt19937_64 rng(1234);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float2> hin((size_t)in_n), hmat((size_t)KK * MM);
  for (auto& x : hin) { x.x = nd(rng); x.y = nd(rng); }
  for (auto& x : hmat) { x.x = nd(rng); x.y = nd(rng); }
  float2 *din, *dout, *dmat;
  CK(cudaMalloc(&din, (size_t)in_n * 8 * batch));
  CK(cudaMalloc(&dout, (size_t)out_n * 8 * batch));
  CK(cudaMalloc(&dmat, hmat.size() * 8));
  for (int b = 0; b < batch; ++b) CK(cudaMemcpy(din + (size_t)b * in_n, hin.data(), (size_t)in_n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmat, hmat.data(), hmat.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, (size_t)out_n * 8 * batch));  // NaN pattern: unwritten outputs are detected

  tc2::Plan plan;
  std::map<tc2::ImageKey, float*> images;
  std::vector<void*> imgs;
  bool ok = true;
  for (int b = 0; b < batch; ++b) {
    tc2::ModeShape t{};
    t.in = din + (size_t)b * in_n; t.out = dout + (size_t)b * out_n; t.mat = dmat;
    t.ips = in_plane; t.ops = out_plane; t.chi_in = c.chi_in; t.chi_out = c.chi_out; t.KK = KK; t.MM = MM;
    t.outer = c.outer; t.inner = c.inner; t.CC = (unsigned)CC;
    ok = ok && tc2::plan_add(plan, t, [&](size_t bytes) { void* p; CK(cudaMalloc(&p, bytes)); imgs.push_back(p); return p; }, images);
  }
  if (!ok) { printf("%-34s : NOT ELIGIBLE for the TMA path\n", c.name); return false; }
  if (!tc2::plan_finish(plan)) { printf("%-34s : plan_finish failed\n", c.name); return false; }
  if (g_stage_override > 0)
    for (auto& L : plan.launches) { L.gm.nstage = std::min(L.gm.nstage, g_stage_override); L.gm.nlo = std::min(L.gm.nlo, L.gm.nstage); }
  tc2::PrepTask2* dprep;
  CK(cudaMalloc(&dprep, plan.preps.size() * sizeof(tc2::PrepTask2)));
  CK(cudaMemcpy(dprep, plan.preps.data(), plan.preps.size() * sizeof(tc2::PrepTask2), cudaMemcpyHostToDevice));
  tc2::tc2_prep_kernel<<<(unsigned)plan.preps.size(), 256>>>(dprep);
  CK(cudaGetLastError());
  std::vector<tc2::ModeTask2*> dt(plan.launches.size(), nullptr);
  std::vector<tc2::Item*> dc(plan.launches.size(), nullptr);
  for (size_t g = 0; g < plan.launches.size(); ++g) {
    auto& L = plan.launches[g];
    CK(cudaMalloc(&dt[g], L.tasks.size() * sizeof(tc2::ModeTask2)));
    CK(cudaMemcpy(dt[g], L.tasks.data(), L.tasks.size() * sizeof(tc2::ModeTask2), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dc[g], L.items.size() * sizeof(tc2::Item)));
    CK(cudaMemcpy(dc[g], L.items.data(), L.items.size() * sizeof(tc2::Item), cudaMemcpyHostToDevice));
  }
  auto launch = [&] {
    for (size_t g = 0; g < plan.launches.size(); ++g) {
      auto& L = plan.launches[g];
      if (g_instr) {
        if (L.last) tc2::tc2_mode_kernel<true, true><<<L.grid, tc2::T2_THREADS, L.smem>>>(dt[g], dc[g], (int)L.items.size(), L.gm);
        else tc2::tc2_mode_kernel<false, true><<<L.grid, tc2::T2_THREADS, L.smem>>>(dt[g], dc[g], (int)L.items.size(), L.gm);
      } else {
        if (L.last) tc2::tc2_mode_kernel<true><<<L.grid, tc2::T2_THREADS, L.smem>>>(dt[g], dc[g], (int)L.items.size(), L.gm);
        else tc2::tc2_mode_kernel<false><<<L.grid, tc2::T2_THREADS, L.smem>>>(dt[g], dc[g], (int)L.items.size(), L.gm);
      }
    }
  };
  launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-34s : KERNEL ERROR %s\n", c.name, cudaGetErrorString(e)); exit(3); }
  const tc2::Launch& L0 = plan.launches[0];
  // verification: all outputs of copy 0 when small, a random sample otherwise; the last copy is compared with copy 0
  std::vector<float2> hout((size_t)out_n), hlast((size_t)out_n);
  CK(cudaMemcpy(hout.data(), dout, (size_t)out_n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hlast.data(), dout + (size_t)(batch - 1) * out_n, (size_t)out_n * 8, cudaMemcpyDeviceToHost));
  const long long nsamp = std::min<long long>(out_n, 200000);
  double maxerr = 0, sumsq = 0;
  long long nan_count = 0, diff_last = 0;
  std::uniform_int_distribution<long long> pick(0, out_n - 1);
  for (long long sidx = 0; sidx < nsamp; ++sidx) {
    const long long idx = (nsamp == out_n) ? sidx : pick(rng);
    const int pp = (int)(idx / out_plane);
    long long r = idx - pp * out_plane;
    const long long o = r / ((long long)c.chi_out * c.inner);
    r -= o * (long long)c.chi_out * c.inner;
    const int cc = (int)(r / c.inner);
    const long long n = r - (long long)cc * c.inner;
    std::complex<double> acc = 0;
    for (int p = 0; p < c.P_in; ++p)
      for (int b = 0; b < c.chi_in; ++b) {
        const float2 a = hin[(size_t)(p * in_plane + (o * c.chi_in + b) * c.inner + n)];
        const float2 m = hmat[(size_t)(p * c.chi_in + b) * MM + pp * c.chi_out + cc];
        acc += std::complex<double>(a.x, a.y) * std::complex<double>(m.x, m.y);
      }
    const float2 got = hout[(size_t)idx];
    if (std::isnan(got.x) || std::isnan(got.y)) { ++nan_count; continue; }
    const double err = std::abs(std::complex<double>(got.x, got.y) - acc);
    maxerr = std::max(maxerr, err);
    sumsq += std::norm(acc);
  }
  for (long long i = 0; i < out_n; ++i)
    if (memcmp(&hout[(size_t)i], &hlast[(size_t)i], 8) != 0) ++diff_last;
  for (long long i = 0; i < out_n; ++i)
    if (std::isnan(hout[(size_t)i].x)) ++nan_count;
  const double rms = std::sqrt(sumsq / std::max<long long>(1, nsamp));
  const bool pass = nan_count == 0 && diff_last == 0 && maxerr / rms < 2e-5;
  const tc2::Geom& gm = L0.gm;
  size_t ntask = 0, nitem = 0;
  for (auto& L : plan.launches) { ntask += L.tasks.size(); nitem += L.items.size(); }
  printf("%-34s : %s  max|err|/rms = %.2e  unwritten = %lld  copy-mismatch = %lld  [%zu launch(es) %s tasks %zu items %zu grid %d T %d kch %d nchunk %d NNp %d nstage %d nlo %d nimg %d nbuf %d smem %zu]\n",
         c.name, pass ? "ok  " : "FAIL", maxerr / rms, nan_count, diff_last, plan.launches.size(), L0.last ? "LAST" : "MID", ntask, nitem, L0.grid, gm.T,
         gm.kch, gm.nchunk, gm.NNp, gm.nstage, gm.nlo, gm.nimg, gm.nbuf, L0.smem);
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
    const double bytes = 8.0 * ((double)in_n + (double)out_n) * batch;
    printf("%-34s   batch %d: %.3f ms  -> %.0f GB/s algorithmic (in+out %.2f GB), %.1f TFLOP/s\n", "", batch, best, bytes / best / 1e6, bytes / 1e9,
           8.0 * KK * MM * (double)CC * batch / best / 1e9);
  }
  for (void* p : imgs) cudaFree(p);
  cudaFree(dprep);
  for (size_t g = 0; g < dt.size(); ++g) { cudaFree(dt[g]); cudaFree(dc[g]); }
  cudaFree(din); cudaFree(dout); cudaFree(dmat);
  return pass;
}

int main(int argc, char** argv) {
  const bool bench = argc > 1 && !strcmp(argv[1], "bench");
  if (argc > 2) g_stage_override = atoi(argv[2]);
  CK(cudaFuncSetAttribute(tc2::tc2_mode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  CK(cudaFuncSetAttribute(tc2::tc2_mode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  CK(cudaFuncSetAttribute((tc2::tc2_mode_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  CK(cudaFuncSetAttribute((tc2::tc2_mode_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  g_instr = argc > 1 && (!strcmp(argv[1], "roles") || !strcmp(argv[1], "decomp"));
  const Case cases[] = {
      {"MID chi32 leg2 (inner 32)", 1, 1, 2 * 32 * 32, 32, 32, 32},
      {"MID chi32 leg1 (inner 1024)", 1, 1, 2 * 32, 32, 32, 1024},
      {"MID chi32 leg0 (inner 32768)", 1, 1, 2, 32, 32, 32768},
      {"MID chi16 (inner 16)", 1, 1, 2 * 16 * 16, 16, 16, 16},
      {"MID chi16 (inner 256)", 1, 1, 2 * 16, 16, 16, 256},
      {"MID chi64 (inner 64)", 1, 1, 2 * 64, 64, 64, 64},
      {"MID chi48->48 (inner 48)", 1, 1, 70, 48, 48, 48},
      {"MID ragged tiles CC=1680", 1, 1, 35, 32, 32, 48},
      {"MID final chi32 keep27 (2 planes)", 2, 2, 32, 32, 27, 1024},
      {"MID final chi16 (planes in box)", 2, 2, 16 * 16, 16, 16, 16},
      {"MID final chi64 keep64 (windows)", 2, 2, 64, 64, 64, 64},
      {"MID chi32 -> 80 rows (windows)", 1, 1, 64, 32, 80, 32},
      {"LAST chi32", 1, 1, 2 * 32 * 32 * 8, 32, 32, 1},
      {"LAST chi16", 1, 1, 2 * 16 * 16 * 16, 16, 16, 1},
      {"LAST chi64", 1, 1, 2 * 64 * 64, 64, 64, 1},
      {"LAST ragged CC=3000", 1, 1, 3000, 32, 32, 1},
      {"LAST final chi32 keep26 (2 planes)", 2, 2, 32768, 32, 26, 1},
      {"LAST chi32 -> 80 (windows)", 1, 1, 4096, 32, 80, 1},
  };
  if (argc > 1 && !strcmp(argv[1], "roles")) {  // per-role cycle accounting
    unsigned long long* dprof;
    CK(cudaMalloc(&dprof, 148 * 32 * 8));
    for (int dbg : {0, 15}) {
      CK(cudaMemcpyToSymbol(tc2::g_tc2_dbg, &dbg, sizeof(int)));
      const Case cs[] = {{"roles MID chi32 leg1 x48", 1, 1, 2 * 32, 32, 32, 1024}, {"roles LAST chi32 x48", 1, 1, 2 * 32 * 32 * 32, 32, 32, 1},
                         {"roles MID chi64 leg1 x4", 1, 1, 2 * 64, 64, 64, 4096}, {"roles LAST chi64 x4", 1, 1, 2 * 64 * 64 * 64, 64, 64, 1},
                         {"roles MID final chi64 x4", 2, 2, 64, 64, 64, 4096}};
      for (auto& c : cs) {
        CK(cudaMemset(dprof, 0, 148 * 32 * 8));
        CK(cudaMemcpyToSymbol(tc2::g_tc2_prof, &dprof, sizeof(dprof)));
        run_case(c, c.chi_in >= 64 ? 4 : 48, false);
        std::vector<unsigned long long> h(148 * 32);
        CK(cudaMemcpy(h.data(), dprof, h.size() * 8, cudaMemcpyDeviceToHost));
        double a[32] = {0};
        for (int b = 0; b < 148; ++b) for (int i = 0; i < 32; ++i) a[i] += (double)h[b * 32 + i] / 148.0;
        printf("dbg=%d %s (avg cycles per CTA)\n", dbg, c.name);
        printf("  producer: wait image-free %.0f, tensormap fence %.0f, wait stage-empty %.0f, issue %.0f, total %.0f\n", a[0], a[1], a[2], a[3], a[4]);
        printf("  mma     : wait image %.0f, wait tmem-empty %.0f, wait full %.0f, wait lo %.0f, issue %.0f, total %.0f\n", a[8], a[9], a[10], a[11], a[12], a[13]);
        printf("  splitter: wait full %.0f, work %.0f, total %.0f\n", a[16], a[17], a[18]);
        printf("  epilogue0: wait tmem-full %.0f, wait staging %.0f, work(incl staging wait) %.0f, total %.0f\n", a[20], a[21], a[22], a[23]);
        printf("  epilogue1: wait tmem-full %.0f, wait staging %.0f, work(incl staging wait) %.0f, total %.0f\n", a[24], a[25], a[26], a[27]);
        printf("  store issue (thread 0 of each group): %.0f %.0f\n", a[28], a[29]);
        unsigned long long* nul = nullptr;
        CK(cudaMemcpyToSymbol(tc2::g_tc2_prof, &nul, sizeof(nul)));
      }
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "decomp")) {  // pipeline decomposition with the debug switches (results are garbage by design)
    for (int dbg : {0, 2, 7, 15}) {
      CK(cudaMemcpyToSymbol(tc2::g_tc2_dbg, &dbg, sizeof(int)));
      printf("dbg=%d\n", dbg);
      run_case({"decomp MID chi32 leg1 x48", 1, 1, 2 * 32, 32, 32, 1024}, 48, true);
      run_case({"decomp LAST chi32 x48", 1, 1, 2 * 32 * 32 * 32, 32, 32, 1}, 48, true);
      run_case({"decomp MID chi64 leg1 x4", 1, 1, 2 * 64, 64, 64, 4096}, 4, true);
      run_case({"decomp LAST chi64 x4", 1, 1, 2 * 64 * 64 * 64, 64, 64, 1}, 4, true);
      run_case({"decomp MID final chi64 x4", 2, 2, 64, 64, 64, 4096}, 4, true);
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "prof")) {  // one launch for ncu
    run_case({"prof MID chi32 leg1 x48", 1, 1, 2 * 32, 32, 32, 1024}, 48, false);
    return 0;
  }
  int nfail = 0;
  for (auto& c : cases) nfail += run_case(c, 2, false) ? 0 : 1;
  printf("%d case(s) failed\n", nfail);
  if (bench) {
    // bandwidth: 48 interior chi=32 site tensors (16.8 MB each, 0.8 GB in + 0.8 GB out: beyond the 126 MB L2)
    run_case({"bench MID chi32 leg1 x48", 1, 1, 2 * 32, 32, 32, 1024}, 48, true);
    run_case({"bench MID chi32 leg2 x48", 1, 1, 2 * 32 * 32, 32, 32, 32}, 48, true);
    run_case({"bench MID chi32 leg0 x48", 1, 1, 2, 32, 32, 32768}, 48, true);
    run_case({"bench LAST chi32 x48", 1, 1, 2 * 32 * 32 * 32, 32, 32, 1}, 48, true);
    run_case({"bench MID final chi32 x24", 2, 2, 32, 32, 32, 1024}, 24, true);
    run_case({"bench MID chi64 leg1 x4", 1, 1, 2 * 64, 64, 64, 4096}, 4, true);
    run_case({"bench MID final chi64 x4", 2, 2, 64, 64, 64, 4096}, 4, true);
    run_case({"bench LAST chi64 x4", 1, 1, 2 * 64 * 64 * 64, 64, 64, 1}, 4, true);
  }
  return nfail;
}
