// kernels_dmma.cuh — fp64 tensor-core (DMMA, mma.sync m8n8k4 f64) Hermitian Gram of a site tensor with
// itself: the reduced-factor Gram of the simple update (R†R of the thin QR at simple_update.jl:47-48),
// G[i][j] = Σ_col conj(X[i][col]) · X[j][col],  i = (plane, active index), exact products of the
// stored values accumulated in fp64.
//
// The complex Hermitian rank-K update is run as a REAL symmetric one: with the 2·MM real rows
// W[(i,0)] = Re X[i], W[(i,1)] = Im X[i],  S = W·Wᵀ gives
//     Re G[i][j] = S[(i,0),(j,0)] + S[(i,1),(j,1)],   Im G[i][j] = S[(i,0),(j,1)] − S[(i,1),(j,0)],
// and only the 32×32 blocks of S on or below the diagonal are computed (10 of 16 for MM = 64), one
// block per warp, 16 m8n8k4 accumulator tiles each.  A CTA streams its K-split of the tensor through
// two shared-memory stages of KCH columns (next chunk prefetched into registers while the current
// one is multiplied), converting the stored scalars to fp64 on the way in.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels_tensor.cuh"

namespace tnqs {

constexpr int DG_KCH = 32;            // columns per stage
constexpr int DG_LD = DG_KCH + 4;     // row stride (doubles): rows land 32 B apart modulo 256 B → conflict-free fragment loads
constexpr int DG_WARPS = 10;
constexpr int DG_THREADS = DG_WARPS * 32;
constexpr int DG_MAXPF = 14;          // prefetch registers (complex elements) per thread: ≥ MM·KCH / DG_THREADS for MM ≤ 128

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// grid = (nsplit_max, block groups, tasks); partial[split][i·MM + j] receives the split's contribution
// (both triangles are written so that the generic gram_reduce_kernel can finish the job)
template <typename R, bool INNER1>
__global__ void __launch_bounds__(DG_THREADS) gram_dmma_kernel(const GramTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  extern __shared__ __align__(16) double dg_smem[];
  const GramTask t = tasks[blockIdx.z];
  const int split = blockIdx.x;
  if (split >= t.nsplit) return;
  const int MM = t.MM, rows = 2 * MM;
  const int R32 = (rows + 31) / 32;
  const int nblk = R32 * (R32 + 1) / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int blk = blockIdx.y * DG_WARPS + warp;
  if (blockIdx.y * DG_WARPS >= nblk) return;
  // block (br, bc), br ≥ bc, enumerated row by row
  int br = 0, bc = 0;
  {
    int b = blk < nblk ? blk : 0, r = 0;
    while (b >= r + 1) { b -= r + 1; ++r; }
    br = r; bc = b;
  }
  const bool active = blk < nblk;
  const int prow = R32 * 32;                 // padded rows held per stage
  double* stage0 = dg_smem;
  double* stage1 = dg_smem + (size_t)prow * DG_LD;
  const C* __restrict__ X = (const C*)t.X;
  const unsigned cb = (unsigned)split * t.cols_per_split;
  const unsigned ce = min(t.CC, cb + t.cols_per_split);
  const int nchunk = (int)((ce - cb + DG_KCH - 1) / DG_KCH);
  const int per_thread = (MM * DG_KCH + DG_THREADS - 1) / DG_THREADS;  // ≤ DG_MAXPF

  // zero the padding rows once (rows ≥ 2·MM of both stages)
  for (int idx = tid; idx < (prow - rows) * DG_LD; idx += DG_THREADS) {
    stage0[(size_t)rows * DG_LD + idx] = 0.0;
    stage1[(size_t)rows * DG_LD + idx] = 0.0;
  }

  C pf[DG_MAXPF];
  auto gload = [&](int ch) {
    const unsigned k0 = cb + (unsigned)ch * DG_KCH;
#pragma unroll
    for (int r = 0; r < DG_MAXPF; ++r) {
      pf[r] = c_zero<C>();
      if (r >= per_thread) continue;
      const int idx = tid + r * DG_THREADS;
      if (idx >= MM * DG_KCH) continue;
      int kk, i;
      if (INNER1) { i = idx % MM; kk = idx / MM; } else { kk = idx % DG_KCH; i = idx / DG_KCH; }
      const unsigned col = k0 + kk;
      if (col < ce) {
        const int p = i / t.chi, l = i - p * t.chi;
        long long a;
        if (INNER1) a = p * t.xps + (long long)col * t.chi + l;
        else { const unsigned o = col / t.inner, n = col - o * t.inner; a = p * t.xps + ((long long)o * t.chi + l) * t.inner + n; }
        pf[r] = X[a];
      }
    }
  };
  auto sstore = [&](double* st) {
#pragma unroll
    for (int r = 0; r < DG_MAXPF; ++r) {
      if (r >= per_thread) continue;
      const int idx = tid + r * DG_THREADS;
      if (idx >= MM * DG_KCH) continue;
      int kk, i;
      if (INNER1) { i = idx % MM; kk = idx / MM; } else { kk = idx % DG_KCH; i = idx / DG_KCH; }
      st[(size_t)(2 * i) * DG_LD + kk] = (double)pf[r].x;
      st[(size_t)(2 * i + 1) * DG_LD + kk] = (double)pf[r].y;
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }

  if (nchunk > 0) gload(0);
  const int fr = lane >> 2, fk = lane & 3;  // fragment row / k of this lane
  for (int ch = 0; ch < nchunk; ++ch) {
    double* st = (ch & 1) ? stage1 : stage0;
    sstore(st);
    if (ch + 1 < nchunk) gload(ch + 1);
    __syncthreads();  // stage `st` complete; the other stage was last read two iterations ago
    if (active) {
      const double* arow = st + (size_t)(br * 32 + fr) * DG_LD + fk;
      const double* brow = st + (size_t)(bc * 32 + fr) * DG_LD + fk;
#pragma unroll
      for (int ks = 0; ks < DG_KCH / 4; ++ks) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = arow[(size_t)(8 * a) * DG_LD + 4 * ks];
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = brow[(size_t)(8 * b) * DG_LD + 4 * ks];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
      }
    }
    // the next iteration overwrites the OTHER stage, which every warp finished reading before the
    // barrier above; this stage is rewritten two iterations from now, after another barrier
  }
  // ---- epilogue: S block → complex G entries ----------------------------------------------------------
  if (!active) return;
  double2* __restrict__ P = t.partial + (long long)split * MM * MM;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      // this lane: S[row][col], S[row][col+1] with row = br·32 + 8a + fr = (i, part), col = bc·32 + 8b + 2·fk = (j, 0)
      const double c0 = acc[a][b][0], c1 = acc[a][b][1];
      const double d0 = __shfl_xor_sync(0xffffffffu, c0, 4);  // partner row (i, 1−part)
      const double d1 = __shfl_xor_sync(0xffffffffu, c1, 4);
      const int row = br * 32 + 8 * a + fr;
      if (row & 1) continue;
      const int i = row >> 1, j = (bc * 32 + 8 * b + 2 * fk) >> 1;
      if (i >= MM || j >= MM || j > i) continue;  // on diagonal blocks keep the lower triangle only
      double2 g; g.x = c0 + d1; g.y = c1 - d0;
      P[(long long)i * MM + j] = g;
      if (i != j) { double2 h; h.x = g.x; h.y = -g.y; P[(long long)j * MM + i] = h; }
    }
}

}  // namespace tnqs
