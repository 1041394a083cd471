// kernels_tc2g.cuh — TMA-fed, warp-specialised tcgen05 Gram contraction for ComplexF32 tensors (sm_100a).
//
//   G[i][j] = Σ_col conj(X[i][col]) · Y[j][col]          (closing contraction of a BP message update, K10 of SURVEY.md §2b:
//                                                          abstractbeliefpropagationcache.jl:162-190; X = site tensor, Y = the
//                                                          tensor with the other messages absorbed)
//
// Same mathematics as tc::tc_gram_kernel (kernels_tc.cuh: real MMAs on the interleaved (re,im) floats, 3-term TF32 split,
// fp32 accumulation in TMEM over a K range, fp64 partial sums reduced in a fixed order), re-built around the memory system
// like the mode product of kernels_tc2.cuh:
//
//   * both tensors arrive by TMA in the canonical tcgen05 operand layout, so no thread computes an address of a streamed
//     tensor:  MID  (active leg not innermost): K-major SWIZZLE_128B boxes (32 floats of (inner,ri) × χ rows);
//              LAST (active leg innermost):     MN-major SWIZZLE_128B_ATOM_32B boxes (32 floats of (j,ri) × kch columns);
//   * the RAW tile is the TF32 `hi` operand (the tensor core truncates); splitter warps only write what cannot be had
//     from memory: lo = rna(x − trunc(x)) of both tensors and, for MID, the rotated copy Ŷ = (Yi, −Yr) (hi and lo) that
//     turns Σ_k X_f[i][k]·Ŷ_f[j][k] into Im G;
//   * hi/lo operands are STACKED along M and N so that one UTCHMMA per k-step produces every needed term
//     (χ ≤ 32: A' = [hi; lo] (M = 128), B' = [hi, lo]; χ = 64: B' = [hi, lo] and a second MMA for lo·hi): 4–8 instead of
//     12–24 MMA instructions per 16 KB of streamed data, issued by one thread;
//   * a ring of stages with full / lo-ready / empty mbarriers decouples producer, splitters, MMA issuer and epilogue;
//     the accumulator is double-buffered in TMEM, the (rare) epilogue of a work item overlaps the next item's MMAs.
//
// Bytes per unit: every element of X and Y is read once (2·8·χ·CC per message); nothing tensor-sized is written.
#pragma once
#include "kernels_tc2.cuh"

namespace tnqs {
namespace tc2g {

using tc::make_desc;
using tc::make_idesc;
using tc::mma_tf32;
using tc::smem_u32;
using tc2::bc;
using tc2::bcu;
using tc2::elect_one;
using tc2::ld_tmem16;
using tc2::lo_part;
using tc2::mbar_arrive;
using tc2::mbar_expect_tx;
using tc2::mbar_init;
using tc2::mbar_wait;
using tc2::tmap_acquire;
using tc2::tmem_ld_wait;
using tc2::umma_commit;

constexpr int G_THREADS = 512;  // warp 0: TMA producer · warp 1: MMA issuer (owns TMEM) · warps 2–3: idle · warps 4–11: splitters · warps 12–15: epilogue
constexpr int G_SPLIT = 256;
constexpr int G_MAX_STAGES = 8;
constexpr size_t G_SMEM_BUDGET = 220 * 1024;

struct alignas(64) GramTask2 {
  CUtensorMap x_map, y_map;  // MID: (32 floats, χ rows, inner/16 blocks, outer) box (32, χ, 1, 1) SWIZZLE_128B
                             // LAST: (32 floats, CC rows, 2χ/32 blocks) box (32, kch, 2χ/32) SWIZZLE_128B_ATOM_32B
  CUtensorMap x_pf, y_pf;    // MID: the same tensor with a box of `pfg` K blocks (pfg·128 contiguous bytes per row): L2 prefetch only
  const void* X;             // LAST: L2 prefetch of contiguous column ranges
  const void* Y;
  double2* partial;          // [nslot][χ·χ]
  unsigned nbi;              // MID: inner / 16 (K blocks per outer slice)
  unsigned units;            // MID: K blocks; LAST: columns
  int nitems;                // work items of this task (the stacked variants write slots nitems … 2·nitems−1 too)
  int pfg;                   // MID: K blocks per prefetch box
  int pad_[6];
};
struct GItem { int task, unit0, nunits, slot; };  // MID unit = one K block of 16 complex columns; LAST unit = kch columns

struct GGeom {
  int chi;
  int stacked;      // hi/lo stacked along M as well (4χ ≤ 128)
  int nb;           // MID: K blocks per stage
  int kch;          // LAST: columns (K) per stage
  int nstage;
  uint32_t stage;   // bytes of one stage
  int ncol;         // TMEM columns of one accumulator (MID 2χ, LAST 4χ)
  int flush;        // stages accumulated in TMEM before the epilogue adds the accumulator to its fp32 running sums (the
                    // tensor core truncates every accumulation: ≈ 3.5e-8 relative per step, measured with tools/tc2g_test.cu)
  int pfd;          // L2 prefetch distance of the producer, in stages (0: none)
  int dbg;          // stand-alone pipeline decomposition (tools/tc2g_test.cu): 1 = splitters idle, 2 = no MMAs.  Zero in the product.
};

template <bool LAST>
__global__ void __launch_bounds__(G_THREADS, 1)
tc2_gram_kernel(const GramTask2* __restrict__ tasks, const GItem* __restrict__ items, int nitems, const GGeom gm) {
  extern __shared__ __align__(1024) uint8_t smemg[];
  __shared__ __align__(8) uint64_t bar_full[G_MAX_STAGES], bar_lo[G_MAX_STAGES], bar_empty[G_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nstage = gm.nstage, chi = gm.chi;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * gm.ncol)) tmem_cols <<= 1;

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_lo[s], G_SPLIT); mbar_init(&bar_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_tfull[b], 1); mbar_init(&bar_tempty[b], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  // bytes of one χ-row K block (MID) / of one 32-float MN block of a stage (LAST)
  const uint32_t R = LAST ? (uint32_t)gm.kch * 128u : (uint32_t)chi * 128u;
  const int nbk = chi >> 4;  // LAST: 32-float blocks per tensor row

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // Loads run `nstage` stages ahead of the MMAs; L2 prefetches (no shared memory needed) run gm.pfd stages ahead of the
    // loads with boxes of 512 contiguous bytes per tensor row, so DRAM sees long bursts and the loads hit L2.
    int s = 0; uint32_t ph = 0;
    for (int ii = blockIdx.x; ii < nitems; ii += gridDim.x) {
      GItem im = items[ii];
      im.task = bc(im.task); im.unit0 = bc(im.unit0); im.nunits = bc(im.nunits);
      const GramTask2* __restrict__ tp = tasks + im.task;
      if (elect_one()) { tmap_acquire(&tp->x_map); tmap_acquire(&tp->y_map); }
      if (!LAST) {
        const unsigned nbi = bcu(tp->nbi);
        const int pfg = bc(tp->pfg);
        if (gm.pfd > 0 && elect_one()) { tmap_acquire(&tp->x_pf); tmap_acquire(&tp->y_pf); }
        unsigned o = (unsigned)im.unit0 / nbi, n = (unsigned)im.unit0 - o * nbi;
        // prefetch cursor (block index, multiple of pfg) and its (o, n)
        int pf = im.unit0 / pfg * pfg;
        unsigned po = (unsigned)pf / nbi, pn = (unsigned)pf - po * nbi;
        const int item_end = im.unit0 + im.nunits;
        const int nst = (im.nunits + gm.nb - 1) / gm.nb;
        for (int st = 0; st < nst; ++st) {
          if (gm.pfd > 0) {
            const int lim = min(item_end, im.unit0 + (st + gm.pfd) * gm.nb);
            while (pf < lim) {
              if (elect_one()) {
                asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
                             ::"l"(&tp->y_pf), "r"(0), "r"(0), "r"((int)pn), "r"((int)po) : "memory");
                asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
                             ::"l"(&tp->x_pf), "r"(0), "r"(0), "r"((int)pn), "r"((int)po) : "memory");
              }
              pf += pfg; pn += (unsigned)pfg;
              if (pn >= nbi) { pn -= nbi; ++po; }
            }
          }
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          if (elect_one()) mbar_expect_tx(&bar_full[s], (uint32_t)gm.nb * 2u * R);
          const uint32_t bar = smem_u32(&bar_full[s]);
          uint32_t dst = smem_u32(smemg + (size_t)s * gm.stage);
          for (int b = 0; b < gm.nb; ++b, dst += 6u * R) {
            // blocks past the item's range still arrive (they belong to the next item or lie outside the tensor, where the
            // TMA unit fills zeros); the MMA issuer skips them
            if (elect_one()) {
              asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                           ::"r"(dst), "l"(&tp->y_map), "r"(0), "r"(0), "r"((int)n), "r"((int)o), "r"(bar) : "memory");
              asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                           ::"r"(dst + 4u * R), "l"(&tp->x_map), "r"(0), "r"(0), "r"((int)n), "r"((int)o), "r"(bar) : "memory");
            }
            if (++n == nbi) { n = 0; ++o; }
          }
          if (++s == nstage) { s = 0; ph ^= 1u; }
        }
      } else {
        const unsigned long long rowb = (unsigned long long)chi * 8ull;  // bytes of one column (a row of 2χ floats)
        const unsigned long long xp = (unsigned long long)tp->X, yp = (unsigned long long)tp->Y;
        const long long cols = (long long)bcu(tp->units);
        long long row = (long long)im.unit0 * gm.kch;
        const long long row_end = min(cols, row + (long long)im.nunits * gm.kch);
        long long pfr = row;
        for (int st = 0; st < im.nunits; ++st, row += gm.kch) {
          if (gm.pfd > 0) {
            const long long lim = min(row_end, row + (long long)gm.pfd * gm.kch);
            while (pfr < lim) {
              const unsigned bytes = (unsigned)(min((long long)gm.kch, row_end - pfr) * (long long)rowb);
              if (elect_one()) {
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(yp + (unsigned long long)pfr * rowb), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xp + (unsigned long long)pfr * rowb), "r"(bytes) : "memory");
              }
              pfr += gm.kch;
            }
          }
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          if (elect_one()) mbar_expect_tx(&bar_full[s], 2u * (uint32_t)nbk * R);
          const uint32_t bar = smem_u32(&bar_full[s]);
          const uint32_t dst = smem_u32(smemg + (size_t)s * gm.stage);
          if (elect_one()) {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(&tp->x_map), "r"(0), "r"((int)row), "r"(0), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst + 2u * (uint32_t)nbk * R), "l"(&tp->y_map), "r"(0), "r"((int)row), "r"(0), "r"(bar) : "memory");
          }
          if (++s == nstage) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // Every operand below derives from kernel parameters and loop counters (uniform registers); lane 0 issues.
    int s = 0; uint32_t ph = 0;
    int buf = 0; uint32_t bph = 0;
    const uint32_t sbase = smem_u32(smemg);
    // MID: D[(Y | Ŷ | Yl | Ŷl rows), (Xh | Xl cols)];  LAST: D[(Xh | Xl rows (i,ri)), (Yh | Yl cols (j,rj))]
    const uint32_t idesc1 = LAST ? make_idesc(128, 4 * chi, 1, 1) : make_idesc(128, 2 * chi, 0, 0);
    const uint32_t idesc2 = LAST ? make_idesc(128, 2 * chi, 1, 1) : make_idesc(128, chi, 0, 0);
    for (int ii = blockIdx.x; ii < nitems; ii += gridDim.x) {
      const int nunits = bc(items[ii].nunits);
      const int nst = LAST ? nunits : (nunits + gm.nb - 1) / gm.nb;
      int acc = 0, left = nunits, since = 0;
      uint32_t d_tmem = tmem;
      for (int st = 0; st < nst; ++st) {
        if (since == 0) {  // a fresh accumulator: the epilogue must have drained its previous content
          mbar_wait(smem_u32(&bar_tempty[buf]), bph ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;");
          d_tmem = tmem + (uint32_t)buf * (uint32_t)gm.ncol;
          acc = 0;
        }
        mbar_wait(smem_u32(&bar_full[s]), ph);
        mbar_wait(smem_u32(&bar_lo[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t st0 = sbase + (uint32_t)s * gm.stage;
        const bool flush_now = (since + 1 == gm.flush) || st == nst - 1;
        if (elect_one()) {
          if (!(gm.dbg & 2)) {
            if (!LAST) {
              const int nbv = min(gm.nb, left);
              for (int b = 0; b < nbv; ++b) {
                const uint32_t blk = st0 + (uint32_t)b * 6u * R;
                const uint64_t a1 = make_desc(blk, 16u, 1024u, 2), b1 = make_desc(blk + 4u * R, 16u, 1024u, 2);
                if (gm.stacked) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) mma_tf32(d_tmem, a1 + 2u * ks, b1 + 2u * ks, idesc1, (acc | b | ks) != 0);
                } else {
                  const uint64_t a2 = make_desc(blk + 2u * R, 16u, 1024u, 2);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) {
                    mma_tf32(d_tmem, a1 + 2u * ks, b1 + 2u * ks, idesc1, (acc | b | ks) != 0);
                    mma_tf32(d_tmem, a2 + 2u * ks, b1 + 2u * ks, idesc2, 1);
                  }
                }
              }
            } else {
              const uint64_t a1 = make_desc(st0, R, 512u, 1), b1 = make_desc(st0 + 2u * (uint32_t)nbk * R, R, 512u, 1);
              const uint64_t a2 = make_desc(st0 + (uint32_t)nbk * R, R, 512u, 1);
              const int nks = gm.kch >> 3;
              if (gm.stacked) {
                for (int ks = 0; ks < nks; ++ks) mma_tf32(d_tmem, a1 + 64u * ks, b1 + 64u * ks, idesc1, (acc | ks) != 0);
              } else {
                for (int ks = 0; ks < nks; ++ks) {
                  mma_tf32(d_tmem, a1 + 64u * ks, b1 + 64u * ks, idesc1, (acc | ks) != 0);
                  mma_tf32(d_tmem, a2 + 64u * ks, b1 + 64u * ks, idesc2, 1);
                }
              }
            }
          }
          umma_commit(&bar_empty[s]);
          if (flush_now) umma_commit(&bar_tfull[buf]);
        }
        __syncwarp();
        acc = 1;
        left -= gm.nb;
        if (++s == nstage) { s = 0; ph ^= 1u; }
        ++since;
        if (flush_now) {
          since = 0;
          if (++buf == 2) { buf = 0; bph ^= 1u; }
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // =============================== splitters ===============================
    // Thread t owns float4 entries t, t + 256, … of a stage; their shared-memory offsets do not depend on the stage, so the
    // index arithmetic is done once.  MID: entry e = (block b, float4 f of the χ×128-byte region): reads Y raw and X raw, writes
    // Ŷh, Yl, Ŷl and Xl;  LAST: entry = float4 of the X raw / Y raw regions: writes Xl and Yl.
    const int t = tid - 128;
    constexpr int NE = 4;  // entries per thread (host guarantees ≤ 4·256 float4 per raw tensor and stage)
    const int per = LAST ? nbk * (int)(R >> 4) : (int)(R >> 4);
    const int n16 = LAST ? per : gm.nb * per;
    uint32_t eoff[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
      const int i = t + k * G_SPLIT;
      if (LAST) eoff[k] = (uint32_t)i * 16u;
      else { const int b = i / per, f = i - b * per; eoff[k] = (uint32_t)b * 6u * R + (uint32_t)f * 16u; }
    }
    int s = 0; uint32_t ph = 0;
    for (int ii = blockIdx.x; ii < nitems; ii += gridDim.x) {
      const int nunits = bc(items[ii].nunits);
      const int nst = LAST ? nunits : (nunits + gm.nb - 1) / gm.nb;
      for (int st = 0; st < nst; ++st) {
        mbar_wait(smem_u32(&bar_full[s]), ph);
        uint8_t* const sb = smemg + (size_t)s * gm.stage;
        if (!(gm.dbg & 1)) {
          if (!LAST) {
            float4 y[NE], x[NE];
#pragma unroll
            for (int k = 0; k < NE; ++k)
              if (t + k * G_SPLIT < n16) {
                y[k] = *reinterpret_cast<const float4*>(sb + eoff[k]);
                x[k] = *reinterpret_cast<const float4*>(sb + eoff[k] + 4u * R);
              }
#pragma unroll
            for (int k = 0; k < NE; ++k)
              if (t + k * G_SPLIT < n16) {
                uint8_t* const r = sb + eoff[k];
                float4 l;
                l.x = lo_part(y[k].x); l.y = lo_part(y[k].y); l.z = lo_part(y[k].z); l.w = lo_part(y[k].w);
                *reinterpret_cast<float4*>(r + R) = make_float4(y[k].y, -y[k].x, y[k].w, -y[k].z);
                *reinterpret_cast<float4*>(r + 2u * R) = l;
                *reinterpret_cast<float4*>(r + 3u * R) = make_float4(l.y, -l.x, l.w, -l.z);
                *reinterpret_cast<float4*>(r + 5u * R) = make_float4(lo_part(x[k].x), lo_part(x[k].y), lo_part(x[k].z), lo_part(x[k].w));
              }
          } else {
            const uint32_t T = (uint32_t)per * 16u;  // bytes of one tensor region: [X raw | Xl | Y raw | Yl]
            float4 y[NE], x[NE];
#pragma unroll
            for (int k = 0; k < NE; ++k)
              if (t + k * G_SPLIT < n16) {
                x[k] = *reinterpret_cast<const float4*>(sb + eoff[k]);
                y[k] = *reinterpret_cast<const float4*>(sb + eoff[k] + 2u * T);
              }
#pragma unroll
            for (int k = 0; k < NE; ++k)
              if (t + k * G_SPLIT < n16) {
                *reinterpret_cast<float4*>(sb + eoff[k] + T) = make_float4(lo_part(x[k].x), lo_part(x[k].y), lo_part(x[k].z), lo_part(x[k].w));
                *reinterpret_cast<float4*>(sb + eoff[k] + 3u * T) = make_float4(lo_part(y[k].x), lo_part(y[k].y), lo_part(y[k].z), lo_part(y[k].w));
              }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&bar_lo[s]);
        if (++s == nstage) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= 12) {
    // =============================== epilogue ===============================
    // Every gm.flush stages the accumulator is added to fp32 running sums held in registers (round-to-nearest adds: the
    // truncation bias of long TMEM accumulations does not build up); one fp64 partial per work item leaves at its end.
    const int quad = warp & 3;
    const int m = quad * 32 + lane;  // accumulator row of this thread
    int buf = 0; uint32_t bph = 0;
    for (int ii = blockIdx.x; ii < nitems; ii += gridDim.x) {
      GItem im = items[ii];
      im.task = bc(im.task); im.slot = bc(im.slot); im.nunits = bc(im.nunits);
      const GramTask2* __restrict__ tp = tasks + im.task;
      double2* const P0 = tp->partial + (size_t)im.slot * chi * chi;
      double2* const P1 = tp->partial + (size_t)(im.slot + bc(tp->nitems)) * chi * chi;
      const int nst = LAST ? im.nunits : (im.nunits + gm.nb - 1) / gm.nb;
      const int nfl = (nst + gm.flush - 1) / gm.flush;
      if (!LAST) {
        // row groups of χ rows: Y → Re, Ŷ → Im (hi terms: columns i and χ + i), then (stacked) Yl → Re, Ŷl → Im (columns i)
        const int grp = m / chi, j = m - grp * chi;
        const bool valid = grp < (gm.stacked ? 4 : 2);
        const bool hi = grp < 2;
        float run[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int q = 0; q < 16; ++q) run[c][q] = 0.f;
        for (int f = 0; f < nfl; ++f) {
          const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * (uint32_t)gm.ncol;
          mbar_wait(smem_u32(&bar_tfull[buf]), bph);
          asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c * 16 < chi) {
              uint32_t v[16];
              ld_tmem16(trow + (uint32_t)(c * 16), v);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 16; ++q) run[c][q] += __uint_as_float(v[q]);
              ld_tmem16(trow + (uint32_t)(chi + c * 16), v);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 16; ++q) run[c][q] += hi ? __uint_as_float(v[q]) : 0.f;
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;");
          mbar_arrive(&bar_tempty[buf]);
          if (++buf == 2) { buf = 0; bph ^= 1u; }
        }
        if (valid) {
          double* const dst = reinterpret_cast<double*>((hi ? P0 : P1) + j) + (grp & 1);
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int i = c * 16 + q;
              if (i < chi) dst[(size_t)i * chi * 2] = (double)run[c][q];
            }
        }
      } else {
        // rows (i,ri) of X | Xl, columns (j,rj) of Y | Yl: re = D[(i,0),(j,0)] + D[(i,1),(j,1)], im = D[(i,0),(j,1)] − D[(i,1),(j,0)]
        const int grp = m / (2 * chi), mm = m - grp * 2 * chi;
        const bool valid = grp < (gm.stacked ? 2 : 1);
        const bool hi = grp == 0;
        const int i = mm >> 1, ri = mm & 1;
        float run[8][8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int q = 0; q < 8; ++q) run[c][q] = 0.f;
        for (int f = 0; f < nfl; ++f) {
          const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * (uint32_t)gm.ncol;
          mbar_wait(smem_u32(&bar_tfull[buf]), bph);
          asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (c * 16 < 2 * chi) {
              uint32_t v[16];
#pragma unroll
              for (int h = 0; h < 2; ++h) {  // columns of Yh, then of Yl (the latter count for the rows of Xh only)
                ld_tmem16(trow + (uint32_t)(h * 2 * chi + c * 16), v);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float own = __uint_as_float(v[2 * q]);
                  const float other = __shfl_xor_sync(0xffffffffu, __uint_as_float(v[2 * q + 1]), 1);
                  const float val = ri ? (other - own) : (own + other);
                  run[c][q] += (h == 0 || hi) ? val : 0.f;
                }
              }
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;");
          mbar_arrive(&bar_tempty[buf]);
          if (++buf == 2) { buf = 0; bph ^= 1u; }
        }
        if (valid) {
          double* const dst = reinterpret_cast<double*>((hi ? P0 : P1) + (size_t)i * chi) + ri;
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int jj = c * 8 + q;
              if (jj < chi) dst[(size_t)jj * 2] = (double)run[c][q];
            }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct GramShape { const void* X; const void* Y; int chi; unsigned outer, inner, CC; };

inline bool eligible(const GramShape& t) {
  if (!tc2::encode_fn()) return false;
  const bool last = t.inner == 1;
  if (t.chi < 16 || t.chi > 64) return false;
  if ((reinterpret_cast<uintptr_t>(t.X) & 15) || (reinterpret_cast<uintptr_t>(t.Y) & 15)) return false;
  if ((double)t.CC * t.chi < 65536.0) return false;  // tiny tensors stay on the simple kernels
  if (last) return t.chi % 16 == 0;
  if (t.chi % 8 != 0 || t.inner % 16 != 0) return false;
  if (4 * t.chi > 128 && t.chi % 16 != 0) return false;  // second MMA has N = χ
  return true;
}
inline int last_kch(int chi) { return std::max(8, (1024 / chi) / 8 * 8); }
inline int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
inline int mid_nb(int chi) { static const int o = env_int("TNQS_TC2G_NB", 0); return std::max(1, std::min(o > 0 ? o : 64 / chi, 128 / chi)); }  // ≤ 1024 float4 per raw tensor and stage

inline bool build_task(const GramShape& t, GramTask2& k) {
  const bool last = t.inner == 1;
  tc2::EncodeTiledFn enc = tc2::encode_fn();
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  for (int w = 0; w < 2; ++w) {
    CUtensorMap* map = w ? &k.y_map : &k.x_map;
    void* ptr = const_cast<void*>(w ? t.Y : t.X);
    if (!last) {
      cuuint64_t gd[4] = {32, (cuuint64_t)t.chi, (cuuint64_t)t.inner / 16, (cuuint64_t)t.outer};
      cuuint64_t gs[3] = {(cuuint64_t)t.inner * 8, 128, (cuuint64_t)t.chi * t.inner * 8};
      cuuint32_t box[4] = {32, (cuuint32_t)t.chi, 1, 1};
      if (enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, gd, gs, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              tc2::l2promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    } else {
      cuuint64_t gd[3] = {32, (cuuint64_t)t.CC, (cuuint64_t)t.chi / 16};
      cuuint64_t gs[2] = {(cuuint64_t)t.chi * 8, 128};
      cuuint32_t box[3] = {32, (cuuint32_t)last_kch(t.chi), (cuuint32_t)t.chi / 16};
      if (enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, gd, gs, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, tc2::l2promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    }
  }
  k.nbi = last ? 1u : t.inner / 16;
  k.units = last ? t.CC : t.outer * (t.inner / 16);
  k.X = t.X; k.Y = t.Y;
  k.pfg = 1;
  if (!last) {
    // prefetch boxes: pfg adjacent K blocks = pfg·128 contiguous bytes of every tensor row (one DRAM burst per row)
    const unsigned nbi = t.inner / 16;
    k.pfg = nbi % 4 == 0 ? 4 : (nbi % 2 == 0 ? 2 : 1);
    for (int w = 0; w < 2; ++w) {
      cuuint64_t gd[4] = {32, (cuuint64_t)t.chi, (cuuint64_t)nbi, (cuuint64_t)t.outer};
      cuuint64_t gs[3] = {(cuuint64_t)t.inner * 8, 128, (cuuint64_t)t.chi * t.inner * 8};
      cuuint32_t box[4] = {32, (cuuint32_t)t.chi, (cuuint32_t)k.pfg, 1};
      if (enc(w ? &k.y_pf : &k.x_pf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(w ? t.Y : t.X), gd, gs, box, ones,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    }
  }
  return true;
}

struct GLaunch {
  bool last = false;
  GGeom gm{};
  std::vector<GramTask2> tasks;
  std::vector<long long> units;  // per task
  std::vector<int> ids;          // caller's task index
  std::vector<GItem> items;
  std::vector<int> nslots;       // per task: partial slots to reduce
  int grid = 0;
  size_t smem = 0;
};
struct GPlan {
  std::vector<GLaunch> launches;
  bool empty() const { return launches.empty(); }
};
inline bool plan_add(GPlan& pl, const GramShape& t, int id) {
  if (!eligible(t)) return false;
  GramTask2 k{};
  if (!build_task(t, k)) return false;
  const bool last = t.inner == 1;
  GLaunch* L = nullptr;
  for (auto& c : pl.launches)
    if (c.last == last && c.gm.chi == t.chi) L = &c;
  if (!L) {
    pl.launches.emplace_back();
    L = &pl.launches.back();
    L->last = last;
    GGeom& gm = L->gm;
    gm.chi = t.chi;
    gm.stacked = 4 * t.chi <= 128 ? 1 : 0;
    gm.nb = mid_nb(t.chi);
    gm.kch = last_kch(t.chi);
    gm.stage = last ? (uint32_t)t.chi * (uint32_t)gm.kch * 32u : (uint32_t)gm.nb * 6u * (uint32_t)t.chi * 128u;
    gm.ncol = last ? 4 * t.chi : 2 * t.chi;
    // the M = 128 operand of a narrow tensor (χ < 32) reads up to 16 KB from its block base: keep that inside the allocation
    const size_t slack = 16 * 1024;
    gm.nstage = (int)std::min<size_t>(G_MAX_STAGES, (G_SMEM_BUDGET - slack) / gm.stage);
    L->smem = (size_t)gm.nstage * gm.stage + slack;
    static const int acc_steps = env_int("TNQS_TC2G_ACC", 64), pf_kb = env_int("TNQS_TC2G_PFKB", 128);
    const int per_stage = last ? gm.kch / 8 : gm.nb * 4;  // accumulation steps (UTCHMMA k-steps) per stage
    gm.flush = std::max(1, acc_steps / per_stage);
    gm.pfd = pf_kb * 1024 / (int)(last ? 2 * t.chi * gm.kch * 8 : gm.nb * 2 * t.chi * 128);  // stages of raw data
    gm.dbg = 0;
  }
  L->tasks.push_back(k);
  L->ids.push_back(id);
  L->units.push_back(last ? ((long long)t.CC + L->gm.kch - 1) / L->gm.kch : (long long)t.outer * (t.inner / 16));
  return true;
}
// Work items: ~6 per SM over the launch, every task cut into equal pieces (a multiple of the stage granularity).
inline void plan_finish(GPlan& pl, int sms = 148) {
  for (auto& L : pl.launches) {
    long long total = 0;
    for (long long u : L.units) total += u;
    const int gran = L.last ? 1 : L.gm.nb;
    const long long min_units = 64ll * gran;  // ≥ 64 stages per item amortise its epilogue
    const long long target = std::max(min_units, total / ((long long)sms * 6));
    L.items.clear();
    L.nslots.assign(L.tasks.size(), 0);
    for (size_t i = 0; i < L.tasks.size(); ++i) {
      const long long u = L.units[i];
      const int nit = (int)std::max<long long>(1, (u + target - 1) / target);
      long long per = (u + nit - 1) / nit;
      per = (per + gran - 1) / gran * gran;
      int slot = 0;
      for (long long u0 = 0; u0 < u; u0 += per, ++slot) L.items.push_back({(int)i, (int)u0, (int)std::min(per, u - u0), slot});
      L.tasks[i].nitems = slot;
      L.nslots[i] = slot * (L.gm.stacked ? 2 : 1);
    }
    L.grid = (int)std::min<size_t>((size_t)sms, L.items.size());
  }
}

}  // namespace tc2g
}  // namespace tnqs
