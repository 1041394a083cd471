// kernels_tc2.cuh — TMA-fed, warp-specialised tcgen05 mode product for ComplexF32 site tensors (sm_100a).
//
//   Out[p', o, c, n] = Σ_{p,b} In[p, o, b, n] · Mat[(p,b), (p',c)]          (K2/K6/K7/K10 of SURVEY.md §2b)
//
// Same mathematics as tc::tc_mode_kernel (kernels_tc.cuh: real MMAs on the interleaved (re,im) floats, fp32
// accuracy through a 3-term TF32 split, fp32 accumulation in TMEM), re-built around the memory system:
//
//   * tensor tiles arrive by TMA (cp.async.bulk.tensor, one tensor map per task kept in global memory) straight in
//     the canonical tcgen05 operand layout — SWIZZLE_128B_ATOM_32B for the MN-major tile of the MID variant,
//     SWIZZLE_128B for the K-major tile of the LAST variant (tools/tma_probe.cu shows both images) — so no thread
//     touches an address of the streamed tensor;
//   * the hardware truncates an fp32 operand to TF32 (tools/tma_probe.cu: "truncated 112, rounded 0"), so the RAW
//     tile is the `hi` operand as it lies in shared memory and only lo = rna_tf32(x − trunc_tf32(x)) is computed,
//     element-wise and layout-agnostic, by four splitter warps (3 ALU instructions per float);
//   * a ring of `nstage` raw-tile stages with full / lo-ready / empty mbarriers decouples the TMA producer, the
//     splitters, the MMA issuers and the epilogue; the lo tiles live in their OWN, shorter ring (`nlo` slots): a raw
//     stage has to wait out the DRAM latency, a lo tile only the few hundred cycles between the splitters and the
//     MMAs, so the shared memory not spent on idle lo slots buys more bytes in flight (χ = 64: B images 64 KB +
//     output staging 64 KB leave 92 KB — 3 raw + 2 lo slots of 16 KB instead of 2 + 2; the stand-alone
//     decomposition, tools/tc2_test.cu `decomp`, showed the bare load pipeline at 4.2 TB/s with two stages in
//     flight); the accumulator is multi-buffered in TMEM, so the MMAs of tile t+1 overlap the epilogue of t;
//   * results leave through a shared-memory staging tile and TMA stores (cp.async.bulk.tensor … bulk_group).
//
// Bytes per unit: every input element is read once and every output element written once (8·(χ_in+χ_out)·CC per
// task); the B images (the small matrix, pre-split into hi/lo) stay resident in shared memory.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <utility>
#include <vector>

#include "kernels_tc.cuh"

namespace tnqs {
namespace tc2 {

using tc::make_desc;
using tc::make_idesc;
using tc::mma_tf32;
using tc::smem_u32;
using tc::tf32_rna;

constexpr int T2_THREADS = 608;  // warp 0: TMA producer · warps 1, 2: MMA issuers (even / odd tiles; warp 1 owns TMEM) · warps 3–10: splitters · warps 11–14 / 15–18: epilogue groups 0 / 1
constexpr int T2_SPLIT = 256;    // splitter threads
constexpr int MAX_STAGES = 8;
constexpr size_t SMEM_BUDGET = 220 * 1024;

struct alignas(64) ModeTask2 {
  CUtensorMap in_map;   // MID: (32 floats, χ_in, P_in, inner/16 blocks, outer) box (32, rows, planes, bpb, 1) SWIZZLE_128B_ATOM_32B
                        // LAST: (2χ_in floats, CC, P_in) box (32, 128, 1) SWIZZLE_128B
  CUtensorMap out_map;  // MID: (32 floats, χ_out, inner/16 blocks, outer, P_out) box (32, nc, bpb, 1, 1) no swizzle
                        // LAST: (2χ_out floats, CC, P_out) box (32, 128, 1) SWIZZLE_128B
  const float* image;   // [nchunk][hi|lo][NNp × kch] K-major, no swizzle
  int kch;              // K extent of a stage: rows of the active leg (MID) or floats (LAST); 16 or 32
  int nchunk;           // stages per tile
  int NNp;              // MMA N (floats)
  int chi_in;           // MID: χ_in; LAST: 2χ_in (floats per plane row)
  int npl_out, pp0;     // output planes written by this task, first of them
  int nc, c0;           // MID: output rows per plane / first row; LAST: output floats (= NNp) / first float
  unsigned inner, CC;
  int ntiles;
  int nsib;             // host: number of sibling windows of the same product starting at this task (0 on the others)
  int bpb;              // MID: 128-byte column blocks per TMA box (4 when inner % 64 == 0, else 2 or 1)
  int pad_;
};

// ---- B images for a window of the matrix columns -------------------------------------------------------------
struct PrepTask2 {
  const float2* mat;  // KKc × MMc complex row-major
  float* image;
  int KKc, MMc;
  int NNp, nchunk, kch, last;
  int npl, pp0, nc, c0, chi_out;  // column window in units of complex output indices
};
__host__ __device__ inline int kmajor_off2(int row, int k, int kch) {
  return (row % 8) * 4 + (k % 4) + (k / 4) * 32 + (row / 8) * (kch / 4) * 32;
}
__global__ void __launch_bounds__(256) tc2_prep_kernel(const PrepTask2* __restrict__ tasks) {
  const PrepTask2 t = tasks[blockIdx.x];
  const int per = t.NNp * t.kch;
  for (int idx = threadIdx.x; idx < t.nchunk * per; idx += blockDim.x) {
    const int ch = idx / per, r = idx - ch * per;
    const int n = r / t.kch, kk = r - n * t.kch;
    const int k = ch * t.kch + kk;
    const int jl = n >> 1;  // local complex output index
    const int pl = jl / t.nc, cc = t.c0 + (jl - pl * t.nc);
    const bool col_ok = pl < t.npl && cc < t.chi_out;
    const int jsrc = (t.pp0 + pl) * t.chi_out + cc;
    float v = 0.f;
    if (!t.last) {
      // B[(j',part)][j]: part 0 → Re Mat[j][j'], part 1 → Im
      if (col_ok && k < t.KKc) { const float2 m = t.mat[(long long)k * t.MMc + jsrc]; v = (n & 1) ? m.y : m.x; }
    } else {
      // B̂[(j',ri')][(j,ri)] = [[Mr, −Mi],[Mi, Mr]]
      const int rip = n & 1, j = k >> 1, ri = k & 1;
      if (col_ok && j < t.KKc) {
        const float2 m = t.mat[(long long)j * t.MMc + jsrc];
        v = (rip == ri) ? m.x : (rip ? m.y : -m.y);
      }
    }
    const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
    float* base = t.image + (long long)ch * 2 * per;
    base[kmajor_off2(n, kk, t.kch)] = hi;
    base[per + kmajor_off2(n, kk, t.kch)] = lo;
  }
}

// debug switches for the stand-alone pipeline decomposition (tools/tc2_test.cu): 1 = splitters idle, 2 = no MMAs,
// 4 = epilogue skips TMEM reads and arithmetic, 8 = no TMA stores.  Zero in the product.
__device__ int g_tc2_dbg = 0;
// optional per-CTA cycle accounting of the four roles (32 counters per CTA); null in the product
__device__ unsigned long long* g_tc2_prof = nullptr;
#define TC2_T0() const long long t0_ = prof ? clock64() : 0
#define TC2_ACC(i) do { if (prof) acc[i] += clock64() - t0_; } while (0)

// mbarrier wait with the spin loop INSIDE the asm block: the compiler then sees straight-line, convergent code and can
// keep the warp-uniform TMA / MMA operands in uniform registers
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

// ---- device helpers ----------------------------------------------------------------------------------------
// one lane of a converged warp (cute::elect_one_sync): ptxas then knows that the guarded region runs on a single lane and
// issues its UTCHMMA / UTCBAR directly instead of wrapping each one into a lane-serialising vote / elect loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmap_acquire(const CUtensorMap* m) {
  // the map was written by a host copy into (re-used) global memory: order the TMA unit's descriptor fetch after it
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(m) : "memory");
}
// warp-uniform value (lets ptxas keep it in a uniform register: TMA / MMA operands then need no per-lane broadcast loop)
__device__ __forceinline__ int bc(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ unsigned bcu(unsigned v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ float lo_part(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);  // what the tensor core uses of x
  return tf32_rna(x - hi);
}
__device__ __forceinline__ void ld_tmem16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------
// Work item: a strided set of tiles of one task.  The grid is persistent (one CTA per SM); CTA c runs items
// c, c + gridDim.x, … and every role walks the same list, so the stage / accumulator / image rings keep flowing
// across item boundaries (the B image of the next item is fetched while the current one is still being multiplied).
struct Item { int task, tile0, stride, pad_; };  // tiles tile0 + i·stride, i < Geom::T (tiles past the task's last one are no-ops)
// launch-wide shared-memory geometry (maxima over the tasks of the launch)
struct Geom {
  uint32_t slot;   // bytes of one raw (= one lo) ring slot
  uint32_t outb;   // bytes of one output staging buffer (one per epilogue group)
  uint32_t imgb;   // bytes of one B-image buffer
  int nstage, nimg, nbuf, ncol;  // ring depths; accumulator buffers of `ncol` TMEM columns each
  int nlo;         // slots of the lo ring (2 ≤ nlo ≤ nstage)
  // shape class of the launch (every task of a launch has the same): the MMA warp then works on kernel parameters and
  // loop counters only, i.e. on values ptxas keeps in uniform registers — its UTCHMMA operands need no per-lane broadcast
  int kch, nchunk, NNp;
  int T;           // tiles per work item
};

template <bool LAST, bool INSTR = false>
__global__ void __launch_bounds__(T2_THREADS, 1)
tc2_mode_kernel(const ModeTask2* __restrict__ tasks, const Item* __restrict__ items, int nitems, const Geom gm) {
  extern __shared__ __align__(1024) uint8_t smem2[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_lo[MAX_STAGES], bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[4], bar_tempty[4], bar_ifull[2], bar_iempty[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31;
  // warp-uniform warp index (as CUTLASS' canonical_warp_idx_sync): the role dispatch below is then a uniform branch and the
  // issue warps' operands can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nstage = gm.nstage, nimg = gm.nimg, nbuf = gm.nbuf;
  const int nlo = gm.nlo;
  uint8_t* const s_stage = smem2;  // [nstage] raw tiles
  uint8_t* const s_lo = s_stage + (size_t)nstage * gm.slot;  // [nlo] lo tiles
  uint8_t* const s_out = s_lo + (size_t)nlo * gm.slot;
  uint8_t* const s_img = s_out + 2 * (size_t)gm.outb;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(nbuf * gm.ncol)) tmem_cols <<= 1;

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_lo[s], T2_SPLIT); mbar_init(&bar_empty[s], 2); }
    for (int b = 0; b < 4; ++b) { mbar_init(&bar_tfull[b], 1); mbar_init(&bar_tempty[b], 128); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_ifull[b], 1); mbar_init(&bar_iempty[b], 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  // instrumentation (debug switches, per-role cycle counters) exists only in the INSTR instantiation: values loaded from
  // memory would otherwise make the issue warps' control flow non-uniform in the compiler's eyes
  const int dbg = INSTR ? g_tc2_dbg : 0;
  unsigned long long* const prof = (INSTR && g_tc2_prof) ? g_tc2_prof + (size_t)blockIdx.x * 32 : nullptr;
  long long acc[6] = {0, 0, 0, 0, 0, 0};
  const long long tstart = prof ? clock64() : 0;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // The whole warp runs the loop on warp-uniform values (so the TMA operands live in uniform registers); lane 0 issues.
    // Everything per tile is additions: no division.
    {
      int s = 0; uint32_t ph = 0;  // stage ring position
      int ib = 0; uint32_t iph = 0;  // image ring position
      for (int ii = blockIdx.x; ii < nitems; ii += gridDim.x) {
        Item im = items[ii];
        im.task = bc(im.task); im.tile0 = bc(im.tile0); im.stride = bc(im.stride);
        const ModeTask2* __restrict__ tp = tasks + im.task;
        const int kch = gm.kch, nchunk = gm.nchunk, chi_in = bc(tp->chi_in), bpb = LAST ? 4 : bc(tp->bpb);
        const unsigned inner = bcu(tp->inner);
        const uint32_t stg = LAST ? 128u * 128u : (uint32_t)kch * 512u;
        {  // B image of this item, once the MMAs that read the buffer's previous content are done
          { TC2_T0(); mbar_wait(smem_u32(&bar_iempty[ib]), iph ^ 1u); TC2_ACC(0); }
          const uint32_t img_bytes = (uint32_t)nchunk * 2u * (uint32_t)gm.NNp * (uint32_t)kch * 4u;
          if (elect_one()) mbar_expect_tx(&bar_ifull[ib], img_bytes);
          if (elect_one())
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(s_img + (size_t)ib * gm.imgb)), "l"(tp->image), "r"(img_bytes), "r"(smem_u32(&bar_ifull[ib]))
                       : "memory");
          if (++ib == nimg) { ib = 0; iph ^= 1u; }
        }
        { TC2_T0(); if (elect_one()) tmap_acquire(&tp->in_map); TC2_ACC(1); }
        // MID: column of the tile's first block as (o, n), advanced by (d_o, d_n) per tile
        unsigned o = 0, n = 0, d_o = 0, d_n = 0;
        if (!LAST) {
          const unsigned col = (unsigned)im.tile0 * 64u, step = (unsigned)im.stride * 64u;
          o = col / inner; n = col - o * inner;
          d_o = step / inner; d_n = step - d_o * inner;
        }
        const int kpl = LAST ? 1 : (chi_in >= kch ? 0 : kch / chi_in);  // MID: planes per stage when a stage spans planes
        int tile = im.tile0;
        for (int ti = 0; ti < gm.T; ++ti, tile += im.stride) {
          int p = 0, b0 = 0;  // MID: plane / first row of the stage; LAST: plane / first float
          for (int ch = 0; ch < nchunk; ++ch) {
            { TC2_T0(); mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u); TC2_ACC(2); }
            const long long tiss_ = prof ? clock64() : 0;
            if (elect_one()) mbar_expect_tx(&bar_full[s], stg);
            const uint32_t dst = smem_u32(s_stage + (size_t)s * gm.slot);
            const uint32_t bar = smem_u32(&bar_full[s]);
            if (!LAST) {
              unsigned oq = o, nq = n;
              for (int q = 0; q < 4; q += bpb) {
                if (elect_one())
                asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                             ::"r"(dst + (uint32_t)q * (uint32_t)kch * 128u), "l"(&tp->in_map), "r"(0), "r"(b0), "r"(p), "r"((int)(nq >> 4)), "r"((int)oq), "r"(bar)
                             : "memory");
                nq += 16u * (unsigned)bpb;
                if (nq >= inner) { nq -= inner; ++oq; }
              }
              if (kpl) p += kpl;
              else { b0 += kch; if (b0 >= chi_in) { b0 = 0; ++p; } }
            } else {
              if (elect_one())
              asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                           ::"r"(dst), "l"(&tp->in_map), "r"(b0), "r"(tile * 128), "r"(p), "r"(bar)
                           : "memory");
              b0 += 32; if (b0 >= chi_in) { b0 = 0; ++p; }
            }
            if (prof) acc[3] += clock64() - tiss_;
            if (++s == nstage) { s = 0; ph ^= 1u; }
          }
          if (!LAST) { n += d_n; o += d_o; if (n >= inner) { n -= inner; ++o; } }
        }
      }
      if (prof && lane == 0) { prof[0] = acc[0]; prof[1] = acc[1]; prof[2] = acc[2]; prof[3] = acc[3]; prof[4] = clock64() - tstart; }
    }
  } else if (warp <= 2) {
    // =============================== MMA issuers (warp 1: even tiles, warp 2: odd tiles) ===============================
    // One UTCHMMA of this size (M 128, N 2χ', K 8) keeps the tensor pipe busy for 32–64 cycles but takes ~80 cycles to
    // dispatch from one warp; two issuing warps on alternate tiles (separate TMEM accumulators) overlap their dispatch.
    // The whole warp runs the loop; lane 0 issues.  Every value below comes from kernel parameters, block / grid indices
    // and loop counters, so the descriptors sit in uniform registers and consecutive UTCHMMAs are a few instructions apart.
    // The shared-memory descriptors of a stage are a base value plus small constants (the start-address field holds
    // addr >> 4 and never carries: shared memory is < 256 KB).
    {
      const int mw = warp - 1;
      int tl = 0;
      int s = 0; uint32_t ph = 0;
      int l = 0;                      // lo ring position (advances with every chunk, like s)
      int ib = 0; uint32_t iph = 0;
      int buf = 0; uint32_t bph = 0;  // accumulator ring position
      const int kch = gm.kch, nchunk = gm.nchunk, NNp = gm.NNp;
      const uint64_t stage_step = (uint64_t)(gm.slot >> 4);
      const uint32_t idesc = make_idesc(128, NNp, LAST ? 0 : 1, 0);
      const uint32_t per_b = (uint32_t)NNp * (uint32_t)kch * 4u;
      const uint64_t a_raw0 = LAST ? make_desc(smem_u32(s_stage), 16u, 1024u, 2)                       // K-major SW128: SBO = 8-row group
                                   : make_desc(smem_u32(s_stage), (uint32_t)kch * 128u, 512u, 1);      // MN-major SW128/32B: LBO = MN-atom stride, SBO = K-atom (4 rows)
      const uint64_t a_lo0 = LAST ? make_desc(smem_u32(s_lo), 16u, 1024u, 2) : make_desc(smem_u32(s_lo), (uint32_t)kch * 128u, 512u, 1);
      const uint64_t a_k = LAST ? 2u : 64u;                                                             // 32 B / 1024 B per k-step of 8
      const uint64_t b_img0 = make_desc(smem_u32(s_img), 128u, (uint32_t)kch * 32u, 0);
      const uint64_t b_lo_off = (uint64_t)(per_b >> 4), chunk_step = (uint64_t)((2u * per_b) >> 4), img_step = (uint64_t)(gm.imgb >> 4);
      const int nks = kch >> 3;
      const int n_my = ((int)nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      for (int k = 0; k < n_my; ++k) {
        { TC2_T0(); mbar_wait(smem_u32(&bar_ifull[ib]), iph); TC2_ACC(0); }
        const uint64_t b_hi0 = b_img0 + (uint64_t)ib * img_step;
        for (int ti = 0; ti < gm.T; ++ti, ++tl) {
          if ((tl & 1) != mw) {
            // The other issuer's tile: observe every phase of its stages (a parity wait must never fall a phase behind) and
            // take part in releasing them (empty[] counts both issuers), so a stage cannot be refilled before this warp has
            // seen it — the protocol does not depend on how far one warp runs ahead of the other.
            for (int ch = 0; ch < nchunk; ++ch) {
              mbar_wait(smem_u32(&bar_full[s]), ph);
              if (lane == 0) mbar_arrive(&bar_empty[s]);
              if (++s == nstage) { s = 0; ph ^= 1u; }
              if (++l == nlo) l = 0;
            }
            if (++buf == nbuf) { buf = 0; bph ^= 1u; }
            continue;
          }
          { TC2_T0(); mbar_wait(smem_u32(&bar_tempty[buf]), bph ^ 1u); TC2_ACC(1); }
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t d_tmem = tmem + (uint32_t)buf * (uint32_t)gm.ncol;
          uint64_t bh = b_hi0;
          for (int ch = 0; ch < nchunk; ++ch, bh += chunk_step) {
            { TC2_T0(); mbar_wait(smem_u32(&bar_full[s]), ph); TC2_ACC(2); }
            const long long tiss_ = prof ? clock64() : 0;
            long long tlo_ = 0;
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint64_t araw = a_raw0 + (uint64_t)s * stage_step;
            const uint64_t alo = a_lo0 + (uint64_t)l * stage_step;
            if (!(dbg & 2) && elect_one()) {
              // terms hi·hi and hi·lo need only the raw tile; lo·hi waits for the splitters
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < nks) mma_tf32(d_tmem, araw + ks * a_k, bh + ks * 16u, idesc, (ch | ks) != 0);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < nks) mma_tf32(d_tmem, araw + ks * a_k, bh + b_lo_off + ks * 16u, idesc, 1);
            }
            __syncwarp();
            { TC2_T0(); mbar_wait(smem_u32(&bar_lo[s]), ph); if (prof) { tlo_ = clock64() - t0_; acc[3] += tlo_; } }
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (elect_one()) {
              if (!(dbg & 2)) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (ks < nks) mma_tf32(d_tmem, alo + ks * a_k, bh + ks * 16u, idesc, 1);
              }
              umma_commit(&bar_empty[s]);
              if (ch == nchunk - 1) umma_commit(&bar_tfull[buf]);
            }
            __syncwarp();
            if (prof) acc[4] += clock64() - tiss_ - tlo_;
            if (++s == nstage) { s = 0; ph ^= 1u; }
            if (++l == nlo) l = 0;
          }
          if (++buf == nbuf) { buf = 0; bph ^= 1u; }
        }
        if (elect_one()) umma_commit(&bar_iempty[ib]);  // the image buffer is free once every MMA of the item has read it
        if (++ib == nimg) { ib = 0; iph ^= 1u; }
      }
      if (prof && lane == 0 && mw == 0) { prof[8] = acc[0]; prof[9] = acc[1]; prof[10] = acc[2]; prof[11] = acc[3]; prof[12] = acc[4]; prof[13] = clock64() - tstart; }
    }
  } else if (warp < 11) {
    // =============================== splitters: lo = rna(x − trunc(x)) ===============================
    const int t = tid - 96;
    int s = 0; uint32_t ph = 0;
    int l = 0;                       // lo slot of the current chunk
    int s2 = 0; uint32_t ph2 = 0;    // raw stage / phase of the chunk that used this lo slot before (nlo chunks earlier)
    {
      const int n_my = ((int)nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int nchunks = n_my * gm.T * gm.nchunk;
      const int n16 = LAST ? 1024 : gm.kch * 32;  // 16-byte pieces of a stage
      for (int c = 0; c < nchunks; ++c) {
        { TC2_T0(); mbar_wait(smem_u32(&bar_full[s]), ph); TC2_ACC(0); }
        const long long tw_ = prof ? clock64() : 0;
        if (c >= nlo) {
          // the MMAs that read this lo slot (chunk c − nlo) must be complete: they commit to that chunk's empty barrier.
          // nlo ≤ nstage, so that barrier cannot be more than one phase ahead of the one waited for.
          mbar_wait(smem_u32(&bar_empty[s2]), ph2);
          if (++s2 == nstage) { s2 = 0; ph2 ^= 1u; }
        }
        const float4* raw = reinterpret_cast<const float4*>(s_stage + (size_t)s * gm.slot);
        float4* lo = reinterpret_cast<float4*>(s_lo + (size_t)l * gm.slot);
        if (!(dbg & 1))
        for (int i0 = t; i0 < n16; i0 += 4 * T2_SPLIT) {
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + u * T2_SPLIT < n16) x[u] = raw[i0 + u * T2_SPLIT];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + u * T2_SPLIT < n16) {
              float4 y;
              y.x = lo_part(x[u].x); y.y = lo_part(x[u].y); y.z = lo_part(x[u].z); y.w = lo_part(x[u].w);
              lo[i0 + u * T2_SPLIT] = y;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&bar_lo[s]);
        if (prof) acc[1] += clock64() - tw_;
        if (++s == nstage) { s = 0; ph ^= 1u; }
        if (++l == nlo) l = 0;
      }
    }
    if (prof && t == 0) { prof[16] = acc[0]; prof[17] = acc[1]; prof[18] = clock64() - tstart; }
  } else {
    // =============================== epilogue (two groups of four warps; group g takes the tiles with tl % 2 == g) ======
    const int eg = (warp - 11) >> 2;
    const int e = tid - 352 - eg * 128;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    uint8_t* const sO = s_out + (size_t)eg * gm.outb;
    const int bar_a = 1 + 2 * eg, bar_b = 2 + 2 * eg;
    int tl = 0;
    int buf = 0; uint32_t bph = 0;  // accumulator ring position (advanced for every tile, both groups)
    for (int ii = blockIdx.x; ii < nitems; ii += gridDim.x) {
      Item im = items[ii];
      im.task = bc(im.task); im.tile0 = bc(im.tile0); im.stride = bc(im.stride);
      const ModeTask2* __restrict__ tp = tasks + im.task;
      const int NNp = gm.NNp;
      const unsigned inner = bcu(tp->inner), CC = bcu(tp->CC);
      const int npl = bc(tp->npl_out), pp0 = bc(tp->pp0), nc = bc(tp->nc), c0 = bc(tp->c0), bpb = LAST ? 4 : bc(tp->bpb);
      if (e == 0) tmap_acquire(&tp->out_map);
      // MID: column of the tile's first block as (o, n), advanced by (d_o, d_n) per tile (no division per tile)
      unsigned o = 0, n = 0, d_o = 0, d_n = 0;
      if (!LAST) {
        const unsigned col = (unsigned)im.tile0 * 64u, step = (unsigned)im.stride * 64u;
        o = col / inner; n = col - o * inner;
        d_o = step / inner; d_n = step - d_o * inner;
      }
      int tile = im.tile0;
      for (int ti = 0; ti < gm.T; ++ti, ++tl, tile += im.stride) {
        const int mybuf = buf;
        const uint32_t myph = bph;
        const unsigned to = o, tn = n;
        if (++buf == nbuf) { buf = 0; bph ^= 1u; }
        if (!LAST) { n += d_n; o += d_o; if (n >= inner) { n -= inner; ++o; } }
        if ((tl & 1) != eg) continue;
        const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)mybuf * (uint32_t)gm.ncol;
        { TC2_T0(); mbar_wait(smem_u32(&bar_tfull[mybuf]), myph); TC2_ACC(0); }
        const long long te_ = prof ? clock64() : 0;
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (dbg & 4) {
          if (e < 32 && elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(bar_b) : "memory");
        } else {
          // Registers only: pairs of 16-column TMEM loads (32 registers) with constant indexing, then singles.
          // Before the first shared-memory write the TMA stores that last read this group's staging buffer must have
          // finished reading it (waited for after the first TMEM loads are back, so the two latencies overlap).
          const int ng = NNp >> 4;
          bool first = true;
          const float sgn = (lane & 1) ? 1.f : -1.f;
          float* const dstm = reinterpret_cast<float*>(sO) + (size_t)quad * (size_t)nc * 32 + lane;  // MID: [plane][blk = quad][row < nc][32 floats]
          const int nrows = npl * nc;
          const int row = quad * 32 + lane;                                                                   // LAST: [box][row][128 B], chunks XOR (row & 7)
          uint8_t* const dstl = sO + (size_t)row * 128;
          auto staging_free = [&]() {
            if (first) {
              TC2_T0();
              if (e < 32 && elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              asm volatile("bar.sync %0, 128;" ::"r"(bar_b) : "memory");
              TC2_ACC(1);
              first = false;
            }
          };
          auto emit = [&](const uint32_t (&v)[16], int g) {
            if (!LAST) {
              // D[(col,ri), (j',part)] → re = D[(c,0),(j',0)] − D[(c,1),(j',1)], im = D[(c,1),(j',0)] + D[(c,0),(j',1)]
              float other[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) other[q] = __shfl_xor_sync(0xffffffffu, __uint_as_float(v[2 * q + 1]), 1);  // all shuffles in flight first
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int jl = g * 8 + q;  // local output index = plane · nc + row
                if (jl < nrows) dstm[(jl < nc ? jl : jl + 3 * nc) * 32] = fmaf(sgn, other[q], __uint_as_float(v[2 * q]));
              }
            } else {
              // D[col, (j',ri')] is the output row as stored
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int cc = g * 4 + q;
                float4 y;
                y.x = __uint_as_float(v[4 * q]); y.y = __uint_as_float(v[4 * q + 1]);
                y.z = __uint_as_float(v[4 * q + 2]); y.w = __uint_as_float(v[4 * q + 3]);
                *reinterpret_cast<float4*>(dstl + (size_t)(cc >> 3) * 16384 + (size_t)(((cc & 7) ^ (row & 7)) << 4)) = y;
              }
            }
          };
          int g0 = 0;
          for (; g0 + 2 <= ng; g0 += 2) {
            uint32_t va[16], vb[16];
            ld_tmem16(trow + (uint32_t)g0 * 16u, va);
            ld_tmem16(trow + (uint32_t)g0 * 16u + 16u, vb);
            tmem_ld_wait();
            if (g0 + 2 >= ng) {  // the accumulator is in registers: hand the TMEM buffer back before the shuffles / stores
              asm volatile("tcgen05.fence::before_thread_sync;");
              mbar_arrive(&bar_tempty[mybuf]);
            }
            staging_free();
            emit(va, g0); emit(vb, g0 + 1);
          }
          for (; g0 < ng; ++g0) {
            uint32_t va[16];
            ld_tmem16(trow + (uint32_t)g0 * 16u, va);
            tmem_ld_wait();
            if (g0 + 1 >= ng) {
              asm volatile("tcgen05.fence::before_thread_sync;");
              mbar_arrive(&bar_tempty[mybuf]);
            }
            staging_free();
            emit(va, g0);
          }
        }
        if (dbg & 4) {
          asm volatile("tcgen05.fence::before_thread_sync;");
          mbar_arrive(&bar_tempty[mybuf]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_a) : "memory");
        const long long tst_ = prof ? clock64() : 0;
        if (e < 32 && !(dbg & 8)) {  // first warp of the group, uniform operands; lane 0 issues
          if (!LAST) {
            unsigned oq = to, nq = tn, col0 = (unsigned)tile * 64u;
            for (int q = 0; q < 4; q += bpb) {
              if (col0 < CC) {
                const uint32_t src = smem_u32(sO) + (uint32_t)q * (uint32_t)nc * 128u;
                for (int pl = 0; pl < npl; ++pl)
                  if (elect_one())
                  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                               ::"l"(&tp->out_map), "r"(0), "r"(c0), "r"((int)(nq >> 4)), "r"((int)oq), "r"(pp0 + pl), "r"(src + (uint32_t)pl * 4u * (uint32_t)nc * 128u)
                               : "memory");
              }
              col0 += 16u * (unsigned)bpb; nq += 16u * (unsigned)bpb;
              if (nq >= inner) { nq -= inner; ++oq; }
            }
          } else {
            for (int bx = 0; bx < NNp / 32; ++bx) {
              if (bx * 32 >= nc) break;
              if (elect_one())
              asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                           ::"l"(&tp->out_map), "r"(c0 + bx * 32), "r"(tile * 128), "r"(pp0), "r"(smem_u32(sO) + (uint32_t)bx * 16384u)
                           : "memory");
            }
          }
          if (elect_one()) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (prof) { acc[2] += clock64() - te_; acc[3] += clock64() - tst_; }
      }
    }
    if (e < 32 && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all output writes complete
    if (prof && e == 0) { prof[20 + 4 * eg] = acc[0]; prof[21 + 4 * eg] = acc[1]; prof[22 + 4 * eg] = acc[2]; prof[23 + 4 * eg] = clock64() - tstart; prof[28 + eg] = acc[3]; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side: which mode products the TMA path takes and how they are cut into tasks
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline CUtensorMapL2promotion l2promo() {
  static int v = [] { const char* e = getenv("TNQS_TC2_PROMO"); return e ? atoi(e) : 3; }();
  return (CUtensorMapL2promotion)v;
}
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

struct ModeShape {  // one mode product, complex element units (mirrors tnqs::ModeTask)
  const void* in; void* out; const void* mat;
  long long ips, ops;
  int chi_in, chi_out, KK, MM;
  unsigned outer, inner, CC;
};

// true when the TMA kernel can run this product; fills the window list {pp0, npl, c0 (rows | floats), nc}
struct Window { int pp0, npl, c0, nc, NNp; };
inline bool plan_windows(const ModeShape& t, int* kch_out, std::vector<Window>& win) {
  win.clear();
  if (!encode_fn()) return false;
  const bool last = t.inner == 1;
  const int P_in = t.KK / t.chi_in, P_out = t.MM / t.chi_out;
  if (P_in * t.chi_in != t.KK || P_out * t.chi_out != t.MM) return false;
  if ((reinterpret_cast<uintptr_t>(t.in) & 15) || (reinterpret_cast<uintptr_t>(t.out) & 15)) return false;
  if ((double)t.CC * t.KK < 32768.0) return false;  // tiny tensors: launch-bound either way, keep them on the simple kernels
  if (P_in > 1 && ((t.ips * 8) & 15)) return false;
  if (P_out > 1 && ((t.ops * 8) & 15)) return false;
  if (!last) {
    if (t.inner % 16 != 0 || t.chi_in % 16 != 0) return false;
    int kch;
    if (t.KK % 32 == 0 && (t.chi_in % 32 == 0 || (t.chi_in == 16 && P_in % 2 == 0))) kch = 32;
    else if (t.KK % 16 == 0) kch = 16;
    else return false;
    *kch_out = kch;
    if (2 * P_out * t.chi_out <= 128 && (size_t)t.KK * 2 * ((2 * P_out * t.chi_out + 15) / 16 * 16) * 4 <= 96 * 1024) {
      win.push_back({0, P_out, 0, t.chi_out, (2 * P_out * t.chi_out + 15) / 16 * 16});
    } else {
      const int wrows = (size_t)t.KK * 2 * 128 * 4 > 96 * 1024 ? 32 : 64;  // keep the resident B images ≤ 96 KB
      for (int pp = 0; pp < P_out; ++pp)
        for (int c0 = 0; c0 < t.chi_out; c0 += wrows) {
          const int nc = std::min(wrows, t.chi_out - c0);
          win.push_back({pp, 1, c0, nc, (2 * nc + 15) / 16 * 16});
        }
    }
  } else {
    if (t.chi_in % 16 != 0 || t.chi_out % 2 != 0) return false;
    *kch_out = 32;
    const int wfl = (size_t)(2 * t.KK) * 2 * 128 * 4 > 96 * 1024 ? 64 : 128;  // keep the resident B images ≤ 96 KB
    for (int pp = 0; pp < P_out; ++pp)
      for (int f0 = 0; f0 < 2 * t.chi_out; f0 += wfl) {
        const int nf = std::min(wfl, 2 * t.chi_out - f0);
        const int NNp = (nf + 31) / 32 * 32;
        win.push_back({pp, 1, f0, nf, NNp});
      }
  }
  // the B images of one window must leave room for at least two stages and one output buffer
  for (auto& w : win) {
    const size_t img = (size_t)(last ? 2 * t.KK : t.KK) * 2 * w.NNp * 4;
    const size_t stg = last ? 16384 : (size_t)*kch_out * 512;
    const size_t outb = last ? (size_t)w.NNp * 512 : (size_t)w.NNp * 256;
    if (img + 2 * 2 * stg + 2 * outb > SMEM_BUDGET) return false;  // two raw + two lo slots at least
  }
  return true;
}

// fills every field of `k` except image / cta bookkeeping; returns false if the driver refuses a map
inline bool build_task(const ModeShape& t, const Window& w, int kch, ModeTask2& k) {
  const bool last = t.inner == 1;
  const int P_in = t.KK / t.chi_in, P_out = t.MM / t.chi_out;
  EncodeTiledFn enc = encode_fn();
  k.kch = kch;
  k.NNp = w.NNp;
  k.npl_out = w.npl; k.pp0 = w.pp0; k.nc = w.nc; k.c0 = w.c0;
  k.inner = t.inner; k.CC = t.CC;
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  if (!last) {
    k.nchunk = t.KK / kch;
    k.chi_in = t.chi_in;
    k.ntiles = (int)((t.CC + 63) / 64);
    const int rows = std::min(kch, t.chi_in), planes = kch / rows;
    const int bpb = t.inner % 64 == 0 ? 4 : (t.inner % 32 == 0 ? 2 : 1);
    k.bpb = bpb;
    const cuuint64_t plane_in = (cuuint64_t)(P_in > 1 ? t.ips * 8 : (cuuint64_t)t.outer * t.chi_in * t.inner * 8);
    const cuuint64_t plane_out = (cuuint64_t)(P_out > 1 ? t.ops * 8 : (cuuint64_t)t.outer * t.chi_out * t.inner * 8);
    {  // (32 floats, χ_in rows, planes, 128-byte blocks of the inner extent, outer): the shared-memory image of a box is
       // [block][plane][row][32 floats] = one K-contiguous MN-atom column per block
      cuuint64_t gd[5] = {32, (cuuint64_t)t.chi_in, (cuuint64_t)P_in, (cuuint64_t)t.inner / 16, (cuuint64_t)t.outer};
      cuuint64_t gs[4] = {(cuuint64_t)t.inner * 8, plane_in, 128, (cuuint64_t)t.chi_in * t.inner * 8};
      cuuint32_t box[5] = {32, (cuuint32_t)rows, (cuuint32_t)planes, (cuuint32_t)bpb, 1};
      if (enc(&k.in_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(t.in), gd, gs, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, l2promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    }
    {  // (32 floats, χ_out rows, blocks, outer, planes): staging tile [block][row][32 floats] per plane
      cuuint64_t gd[5] = {32, (cuuint64_t)t.chi_out, (cuuint64_t)t.inner / 16, (cuuint64_t)t.outer, (cuuint64_t)P_out};
      cuuint64_t gs[4] = {(cuuint64_t)t.inner * 8, 128, (cuuint64_t)t.chi_out * t.inner * 8, plane_out};
      cuuint32_t box[5] = {32, (cuuint32_t)w.nc, (cuuint32_t)bpb, 1, 1};
      if (enc(&k.out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, t.out, gd, gs, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    }
  } else {
    k.nchunk = 2 * t.KK / 32;
    k.chi_in = 2 * t.chi_in;
    k.ntiles = (int)((t.CC + 127) / 128);
    {
      cuuint64_t gd[3] = {(cuuint64_t)t.chi_in * 2, (cuuint64_t)t.CC, (cuuint64_t)P_in};
      cuuint64_t gs[2] = {(cuuint64_t)t.chi_in * 8, (cuuint64_t)(P_in > 1 ? t.ips * 8 : (cuuint64_t)t.CC * t.chi_in * 8)};
      cuuint32_t box[3] = {32, 128, 1};
      if (enc(&k.in_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(t.in), gd, gs, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, l2promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    }
    {
      cuuint64_t gd[3] = {(cuuint64_t)t.chi_out * 2, (cuuint64_t)t.CC, (cuuint64_t)P_out};
      cuuint64_t gs[2] = {(cuuint64_t)t.chi_out * 8, (cuuint64_t)(P_out > 1 ? t.ops * 8 : (cuuint64_t)t.CC * t.chi_out * 8)};
      cuuint32_t box[3] = {32, 128, 1};
      if (enc(&k.out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, t.out, gd, gs, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    }
  }
  k.nsib = 0;
  return true;
}

// A batch of mode products cut into tasks, grouped into launches by shape class (variant, kch, nchunk, NNp), with the B
// images they need (one per distinct (matrix, window)) and the work items of each launch.
struct Launch {
  bool last = false;
  Geom gm{};
  std::vector<ModeTask2> tasks;
  std::vector<Item> items;
  int grid = 0;
  size_t smem = 0;
};
struct Plan {
  std::vector<Launch> launches;
  std::vector<PrepTask2> preps;
  double flops = 0;
  bool empty() const { return launches.empty(); }
};
struct ImageKey {
  const void* mat; int KK, MM, last, pp0, npl, c0, nc, kch;
  bool operator<(const ImageKey& o) const {
    if (mat != o.mat) return mat < o.mat;
    const int a[8] = {KK, MM, last, pp0, npl, c0, nc, kch}, b[8] = {o.KK, o.MM, o.last, o.pp0, o.npl, o.c0, o.nc, o.kch};
    for (int i = 0; i < 8; ++i) if (a[i] != b[i]) return a[i] < b[i];
    return false;
  }
};
// Adds product `t` to the plan if the TMA path can run it (returns false otherwise, plan untouched).
// alloc_image(bytes) returns device memory for a B image; `images` caches them per (matrix, window).
template <class Alloc, class Cache>
inline bool plan_add(Plan& pl, const ModeShape& t, Alloc&& alloc_image, Cache& images) {
  int kch = 0;
  std::vector<Window> win;
  if (!plan_windows(t, &kch, win)) return false;
  const bool last = t.inner == 1;
  std::vector<ModeTask2> mine;
  std::vector<PrepTask2> preps;
  std::vector<std::pair<ImageKey, float*>> fresh;
  for (auto& w : win) {
    ModeTask2 k{};
    if (!build_task(t, w, kch, k)) return false;
    // complex-unit window of the matrix columns
    const int c0c = last ? w.c0 / 2 : w.c0, ncc = last ? (w.nc + 1) / 2 : w.nc;
    ImageKey key{t.mat, t.KK, t.MM, last ? 1 : 0, w.pp0, w.npl, c0c, ncc, kch};
    float* img = nullptr;
    auto f = images.find(key);
    if (f != images.end()) img = f->second;
    for (auto& fr : fresh) if (!(fr.first < key) && !(key < fr.first)) img = fr.second;
    if (!img) {
      img = (float*)alloc_image((size_t)k.nchunk * 2 * k.NNp * kch * 4);
      PrepTask2 p{};
      p.mat = (const float2*)t.mat; p.image = img; p.KKc = t.KK; p.MMc = t.MM; p.NNp = k.NNp; p.nchunk = k.nchunk; p.kch = kch;
      p.last = last ? 1 : 0; p.npl = w.npl; p.pp0 = w.pp0; p.nc = ncc; p.c0 = c0c; p.chi_out = t.chi_out;
      preps.push_back(p);
      fresh.push_back({key, img});
    }
    k.image = img;
    mine.push_back(k);
  }
  for (auto& fr : fresh) images[fr.first] = fr.second;
  for (auto& p : preps) pl.preps.push_back(p);
  // windows of one product may fall into different shape classes (a narrower last window): siblings are the runs of
  // consecutive windows of the same class
  for (size_t i = 0; i < mine.size();) {
    size_t j = i;
    while (j < mine.size() && mine[j].NNp == mine[i].NNp && mine[j].nchunk == mine[i].nchunk) ++j;
    Launch* L = nullptr;
    for (auto& c : pl.launches)
      if (c.last == last && c.gm.kch == kch && c.gm.nchunk == mine[i].nchunk && c.gm.NNp == mine[i].NNp) L = &c;
    if (!L) {
      pl.launches.emplace_back();
      L = &pl.launches.back();
      L->last = last; L->gm.kch = kch; L->gm.nchunk = mine[i].nchunk; L->gm.NNp = mine[i].NNp;
    }
    mine[i].nsib = (int)(j - i);
    for (size_t q = i; q < j; ++q) L->tasks.push_back(mine[q]);
    i = j;
  }
  pl.flops += 8.0 * t.KK * t.MM * (double)t.CC;
  return true;
}
// Work items and the launch geometry.  Tiles of a task are dealt round-robin to its items and the items of one task
// (and of the sibling windows of one product) are adjacent in the list, so that the CTAs running at the same time stream
// neighbouring pieces of the same tensor rows (same DRAM pages; the second read of a shared input tile hits L2).
inline bool plan_finish(Plan& pl, int sms = 148) {
  for (auto& L : pl.launches) {
    Geom& gm = L.gm;
    const bool last = L.last;
    long long total = 0;
    for (auto& k : L.tasks) total += k.ntiles;
    gm.slot = last ? 16384u : (uint32_t)gm.kch * 512u;
    gm.outb = last ? (uint32_t)gm.NNp * 512u : (uint32_t)gm.NNp * 256u;
    gm.imgb = (uint32_t)gm.nchunk * 2u * (uint32_t)gm.NNp * (uint32_t)gm.kch * 4u;
    gm.ncol = gm.NNp;
    gm.outb = (gm.outb + 1023u) & ~1023u;
    gm.imgb = (gm.imgb + 1023u) & ~1023u;
    gm.nbuf = 4 * gm.ncol <= 512 ? 4 : 2;
    gm.nimg = 2;
    // ring slots of `slot` bytes left after the output staging and the B images: two image buffers when that still leaves
    // 4 raw + 2 lo slots, else one; 2 lo slots (3 when there is room for 8), the rest raw stages
    long long nslots = ((long long)SMEM_BUDGET - 2ll * gm.outb - 2ll * gm.imgb) / (long long)gm.slot;
    if (nslots < 6) { gm.nimg = 1; nslots = ((long long)SMEM_BUDGET - 2ll * gm.outb - (long long)gm.imgb) / (long long)gm.slot; }
    if (nslots < 4) return false;
    static const int lo_override = [] { const char* e = getenv("TNQS_TC2_NLO"); return e ? atoi(e) : 0; }();      // tuning
    static const int st_override = [] { const char* e = getenv("TNQS_TC2_NSTAGE"); return e ? atoi(e) : 0; }();
    gm.nlo = nslots >= 8 ? 3 : 2;
    if (lo_override >= 2) gm.nlo = (int)std::min<long long>(lo_override, nslots / 2);
    gm.nstage = (int)std::min<long long>(MAX_STAGES, nslots - gm.nlo);
    if (st_override >= 2) gm.nstage = std::min(gm.nstage, st_override);
    if (gm.nlo > gm.nstage) gm.nlo = gm.nstage;
    L.smem = (size_t)(gm.nstage + gm.nlo) * gm.slot + 2 * (size_t)gm.outb + (size_t)gm.nimg * gm.imgb;
    // every item has exactly T tiles (those past a task's last tile are no-ops): ~24 items per SM, 4 ≤ T ≤ 64
    const int T = (int)std::max<long long>(4, std::min<long long>(64, total / ((long long)sms * 24)));
    gm.T = T;
    L.items.clear();
    for (size_t i = 0; i < L.tasks.size();) {
      const int nsib = std::max(1, L.tasks[i].nsib);
      const int ntiles = L.tasks[i].ntiles;
      const int nit = std::max(1, (ntiles + T - 1) / T);
      for (int j = 0; j < nit; ++j)
        for (int w = 0; w < nsib; ++w) L.items.push_back({(int)i + w, j, nit, 0});
      i += nsib;
    }
    L.grid = (int)std::min<size_t>((size_t)sms, L.items.size());
  }
  return true;
}

}  // namespace tc2
}  // namespace tnqs
