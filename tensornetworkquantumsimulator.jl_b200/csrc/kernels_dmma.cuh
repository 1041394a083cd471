// kernels_dmma.cuh — fp64 tensor-core (DMMA, mma.sync m8n8k4 f64) Hermitian Gram of a site tensor with
// itself: the reduced-factor Gram of the simple update (R†R of the thin QR at simple_update.jl:47-48),
// G[i][j] = Σ_col conj(X[i][col]) · X[j][col],  i = (plane, active index), exact products of the
// stored values accumulated in fp64.
//
// The complex Hermitian rank-K update is run as a REAL symmetric one: with the 2·MM real rows
// W[(i,0)] = Re X[i], W[(i,1)] = Im X[i],  S = W·Wᵀ gives
//     Re G[i][j] = S[(i,0),(j,0)] + S[(i,1),(j,1)],   Im G[i][j] = S[(i,0),(j,1)] − S[(i,1),(j,0)],
// and only the 32×32 blocks of S on or below the diagonal are computed (10 of 16 for MM = 64; of the diagonal
// blocks only the 8×8 tiles on or below THEIR diagonal), 16 m8n8k4 accumulator tiles per off-diagonal block.  A CTA streams its K-split of the tensor through
// two shared-memory stages of KCH columns (chunk ch + 1 is converted to fp64 and stored while chunk ch is
// multiplied, chunk ch + 2 is in flight from global memory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels_tensor.cuh"

namespace tnqs {

constexpr int DG_KCH = 32;            // columns per stage
constexpr int DG_LD = DG_KCH + 4;     // row stride (doubles): rows land 32 B apart modulo 256 B → conflict-free fragment loads

// Warps per CTA and staging registers per thread of the two instantiations.
//   NW = 8, MM ≤ 64 (S has at most 4×4 blocks): one CTA covers all of S.  The 6 off-diagonal blocks take one warp each
//   (16 tiles); the 4 diagonal blocks need only the 10 tiles on or below their own diagonal and their B fragments ARE
//   their A fragments, so two of them share a warp (20 tiles, 4 fragment loads per block and k-step).  Two warps per SM
//   sub-partition with 32 / 32 / 36 / 36 tiles per k-step instead of 3:3:2:2 warps of 16 (tools/dmma_probe.cu: the fp64
//   pipe reaches 30.1 TF/s with 10 resident warps, 35.4 with 8, 36.2 with 12, peak 37.0), and 136 instead of 160 MMAs.
//   NW = 12, MM ≤ 128: 36 lower blocks of a 256-row S = 3 CTAs of 12 warps, one block per warp.
template <int NW> struct DgCfg { static constexpr int MAXPF = NW == 12 ? 11 : 8; };  // ≥ ⌈MM / NW⌉ (+ mapping padding)

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// grid = (nsplit_max, block groups, tasks); partial[split][i·MM + j] receives the split's contribution
// (both triangles are written so that the generic gram_reduce_kernel can finish the job)
//
// Staging without per-element index arithmetic: which element (row i, column kk of the chunk) a thread moves is fixed for
// the whole K range, so its global offset (relative to the chunk) and its shared-memory offset are computed ONCE; a chunk
// then costs one add per element.  MID (active leg not innermost) needs inner % KCH == 0 for that (all columns of a chunk
// then share the outer index — true for every saturated lattice state); other shapes take the generic path, which
// re-derives (outer, inner) per element as before.  Thread ↔ element mapping: MID: a warp moves 32 consecutive columns
// of one row (256 contiguous bytes in, conflict-free 8-byte shared-memory stores); INNER1: a warp moves 4 consecutive
// rows × 8 consecutive columns (eight 32-byte sectors in, 2-way conflicts at worst on the way out — the row stride of
// 72 doubles would make a row-major assignment 16-way conflicted).
template <typename R, bool INNER1, int NW>
__global__ void __launch_bounds__(NW * 32) gram_dmma_kernel(const GramTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  constexpr int NT = NW * 32;
  constexpr int MAXPF = DgCfg<NW>::MAXPF;
  extern __shared__ __align__(16) double dg_smem[];
  const GramTask t = tasks[blockIdx.z];
  const int split = blockIdx.x;
  if (split >= t.nsplit) return;
  const int MM = t.MM, rows = 2 * MM;
  const int R32 = (rows + 31) / 32;
  const int nblk = R32 * (R32 + 1) / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr bool PAIRED = NW == 8;  // diagonal blocks in pairs (see DgCfg)
  const int blk = blockIdx.y * NW + warp;
  if (PAIRED ? blockIdx.y > 0 : blockIdx.y * NW >= nblk) return;
  // block (br, bc), br ≥ bc, enumerated row by row; PAIRED: off-diagonal blocks first, then pairs of diagonal blocks
  int br = 0, bc = 0;
  bool active = blk < nblk, is_diag = false;
  int ndd = 0;  // PAIRED diagonal warp: blocks (br, br) and, if ndd == 2, (br + 1, br + 1)
  if (!PAIRED) {
    int b = blk < nblk ? blk : 0, r = 0;
    while (b >= r + 1) { b -= r + 1; ++r; }
    br = r; bc = b;
  } else {
    const int noff = R32 * (R32 - 1) / 2;
    if (warp < noff) {
      int b = warp, r = 1;
      while (b >= r) { b -= r; ++r; }
      br = r; bc = b; active = true;
    } else {
      is_diag = true;
      br = bc = 2 * (warp - noff);
      ndd = min(2, R32 - br);
      active = ndd > 0;
    }
  }
  const int prow = R32 * 32;                 // padded rows held per stage
  double* stage0 = dg_smem;
  double* stage1 = dg_smem + (size_t)prow * DG_LD;
  const C* __restrict__ X = (const C*)t.X;
  const unsigned cb = (unsigned)split * t.cols_per_split;  // multiple of KCH
  const unsigned ce = min(t.CC, cb + t.cols_per_split);
  const int nchunk = (int)((ce - cb + DG_KCH - 1) / DG_KCH);

  // zero the padding rows once (rows ≥ 2·MM of both stages)
  for (int idx = tid; idx < (prow - rows) * DG_LD; idx += NT) {
    stage0[(size_t)rows * DG_LD + idx] = 0.0;
    stage1[(size_t)rows * DG_LD + idx] = 0.0;
  }

  // ---- per-thread element assignment (fixed for the whole K range) ----
  const bool fast = INNER1 || (t.inner % DG_KCH == 0);
  const int ngrp = INNER1 ? ((MM + 3) >> 2) * 4 : MM;   // warp-sized groups of elements per chunk
  const int per_thread = (ngrp + NW - 1) / NW;          // ≤ MAXPF (host: MM ≤ 64 for 10 warps, ≤ 128 for 12)
  unsigned goff[MAXPF];                                 // element offset inside the chunk; ~0u: nothing to move
  // row / chunk column of entry r: cheap functions of (warp, lane, r) with r a compile-time constant after unrolling
  auto row_of = [&](int r) { const int grp = warp + r * NW; return INNER1 ? (grp >> 2) * 4 + (lane & 3) : grp; };
  auto kk_of = [&](int r) { const int grp = warp + r * NW; return INNER1 ? (grp & 3) * 8 + (lane >> 2) : lane; };
#pragma unroll
  for (int r = 0; r < MAXPF; ++r) {
    goff[r] = ~0u;
    const int grp = warp + r * NW;
    if (r >= per_thread || grp >= ngrp) continue;
    const int i = row_of(r), kk = kk_of(r);
    if (i >= MM) continue;
    const int p = i / t.chi, l = i - p * t.chi;
    if (INNER1) goff[r] = (unsigned)(p * t.xps + (long long)kk * t.chi + l);
    else goff[r] = (unsigned)(p * t.xps + (long long)l * t.inner + kk);  // + (o·χ·inner + n0) of the chunk
  }
  // MID: (outer, inner) position of the chunk's first column, advanced per chunk
  unsigned o_c = 0, n_c = 0;
  if (!INNER1) { o_c = cb / t.inner; n_c = cb - o_c * t.inner; }
  const long long ostride = (long long)t.chi * t.inner;

  C pf[MAXPF];
  auto gload = [&](int ch) {
    const unsigned k0 = cb + (unsigned)ch * DG_KCH;
    const unsigned ncols = min((unsigned)DG_KCH, ce - k0);
    if (fast) {
      const C* __restrict__ Xc = INNER1 ? X + (long long)k0 * t.chi : X + ((long long)o_c * ostride + n_c);
#pragma unroll
      for (int r = 0; r < MAXPF; ++r) {
        pf[r] = c_zero<C>();
        if (r < per_thread && goff[r] != ~0u && (unsigned)kk_of(r) < ncols) pf[r] = Xc[goff[r]];
      }
      if (!INNER1) { n_c += DG_KCH; if (n_c >= t.inner) { n_c = 0; ++o_c; } }
    } else {
      // generic MID path: the columns of a chunk straddle outer slices
#pragma unroll
      for (int r = 0; r < MAXPF; ++r) {
        pf[r] = c_zero<C>();
        if (r >= per_thread || goff[r] == ~0u) continue;
        const unsigned col = k0 + (unsigned)kk_of(r);
        if (col < ce) {
          const unsigned o = col / t.inner, n = col - o * t.inner;
          pf[r] = X[(long long)(goff[r] - (unsigned)kk_of(r)) + (long long)o * ostride + n];
        }
      }
    }
  };
  auto sstore = [&](double* st) {
#pragma unroll
    for (int r = 0; r < MAXPF; ++r) {
      if (r >= per_thread || goff[r] == ~0u) continue;
      double* const d = st + 2 * row_of(r) * DG_LD + kk_of(r);
      d[0] = (double)pf[r].x;
      d[DG_LD] = (double)pf[r].y;
    }
  };

  // accumulator tiles: off-diagonal block: tile (a, b) at 4a + b; diagonal pair: block d, tile (a, b ≤ a) at 10d + a(a+1)/2 + b
  constexpr int NACC = PAIRED ? 20 : 16;
  double acc[NACC][2];
#pragma unroll
  for (int q = 0; q < NACC; ++q) { acc[q][0] = 0.0; acc[q][1] = 0.0; }

  // Software pipeline: while stage (ch & 1) is multiplied, chunk ch + 1 (already in registers) is converted and stored
  // into the other stage and chunk ch + 2 is requested from global memory; one barrier per chunk.  The stores and the
  // DMMAs of a chunk are independent instruction streams, so warps that finish storing start multiplying while others
  // are still converting: the fp64 tensor pipe no longer idles during the staging phase.
  const int fr = lane >> 2, fk = lane & 3;  // fragment row / k of this lane
  if (nchunk > 0) {
    gload(0);
    sstore(stage0);
    if (nchunk > 1) gload(1);
  }
  __syncthreads();
  for (int ch = 0; ch < nchunk; ++ch) {
    double* st = (ch & 1) ? stage1 : stage0;
    if (ch + 1 < nchunk) sstore((ch & 1) ? stage0 : stage1);  // last read by the multiplication of chunk ch − 1, before the barrier
    if (ch + 2 < nchunk) gload(ch + 2);
    if (active && !(PAIRED && is_diag)) {
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
          for (int b = 0; b < 4; ++b) dmma884(acc[4 * a + b][0], acc[4 * a + b][1], af[a], bf[b]);
      }
    } else if (PAIRED && active) {
      // the B fragment of a row group has the same (row, k) ↔ lane mapping as its A fragment
      const double* arow = st + (size_t)(br * 32 + fr) * DG_LD + fk;
#pragma unroll
      for (int ks = 0; ks < DG_KCH / 4; ++ks) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          if (d < ndd) {
            double af[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = arow[(size_t)(32 * d + 8 * a) * DG_LD + 4 * ks];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b <= a; ++b) {
                const int q = (PAIRED ? 10 : 0) * d + a * (a + 1) / 2 + b;
                dmma884(acc[q < NACC ? q : 0][0], acc[q < NACC ? q : 0][1], af[a], af[b]);
              }
          }
        }
      }
    }
    __syncthreads();
  }
  // ---- epilogue: S block → complex G entries ----------------------------------------------------------
  if (!active) return;
  double2* __restrict__ P = t.partial + (long long)split * MM * MM;
  // this lane holds S[row][col], S[row][col+1] of tile (a, b) of block (rb, cb): row = rb·32 + 8a + fr = (i, part),
  // col = cb·32 + 8b + 2·fk = (j, 0)
  auto emit = [&](int rb, int cbk, int a, int b, double c0, double c1) {
    const double d0 = __shfl_xor_sync(0xffffffffu, c0, 4);  // partner row (i, 1−part)
    const double d1 = __shfl_xor_sync(0xffffffffu, c1, 4);
    const int row = rb * 32 + 8 * a + fr;
    if (row & 1) return;
    const int i = row >> 1, j = (cbk * 32 + 8 * b + 2 * fk) >> 1;
    if (i >= MM || j >= MM || j > i) return;  // on diagonal blocks keep the lower triangle only
    double2 g; g.x = c0 + d1; g.y = c1 - d0;
    P[(long long)i * MM + j] = g;
    if (i != j) { double2 h; h.x = g.x; h.y = -g.y; P[(long long)j * MM + i] = h; }
  };
  if (!(PAIRED && is_diag)) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) emit(br, bc, a, b, acc[4 * a + b][0], acc[4 * a + b][1]);
  } else {
#pragma unroll
    for (int d = 0; d < 2; ++d)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          const int q = (PAIRED ? 10 : 0) * d + a * (a + 1) / 2 + b;
          // every lane of the warp takes part in the shuffles of a tile; blocks past the matrix hold zeros and write nothing
          if (d < ndd) emit(br + d, br + d, a, b, acc[q < NACC ? q : 0][0], acc[q < NACC ? q : 0][1]);
        }
  }
}

}  // namespace tnqs
