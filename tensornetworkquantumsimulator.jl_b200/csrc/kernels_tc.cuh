// kernels_tc.cuh — tcgen05 (5th-gen tensor core) variants of the tensor-streaming kernels for
// ComplexF32 states.  fp32 accuracy is kept with the 3×TF32 split (a = a_hi + a_lo, both TF32;
// a·b ≈ a_hi·b_hi + a_hi·b_lo + a_lo·b_hi, fp32 accumulation in TMEM).
//
// A complex mode product  Out[j'][col] = Σ_j Mat[j][j'] · A[j][col]  is issued as REAL MMAs on the
// interleaved (re,im) data as it lies in HBM — no de-interleaving pass:
//
//   MID  (active leg not innermost; (col,ri) is the contiguous direction of the tensor):
//        D[(col,ri), (j',part)] = Σ_j  T[(col,ri), j] · [Mr | Mi][j, (j',part)]
//        A operand = tensor tile, MN-major (SWIZZLE_128B_BASE32B, the only MN-major layout TF32 has),
//        M = 128 rows = 64 complex columns;  epilogue: re = D[(c,0),(j',0)] − D[(c,1),(j',1)],
//        im = D[(c,1),(j',0)] + D[(c,0),(j',1)]  (one lane-pair shuffle).
//   LAST (active leg innermost; (j,ri) is contiguous):
//        D[col, (j',ri')] = Σ_{(j,ri)} T[col,(j,ri)] · B̂[(j,ri),(j',ri')],  B̂ = [[Mr, Mi],[−Mi, Mr]]
//        A operand = tensor rows, K-major (no swizzle), M = 128 columns; D is the output as stored.
//
// The small matrix is expanded once into "B images" that already have the canonical shared-memory
// layout (K-major, no swizzle, hi and lo parts), so a CTA fetches it with one cp.async.bulk (TMA
// bulk copy) and keeps it resident while it streams its range of tensor tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tnqs {
namespace tc {

constexpr int KC = 32;  // K floats per staged chunk (4 MMA k-steps of 8)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor bit layout)
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, int accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 x, float4& hi, float4& lo) {
  hi.x = tf32_rna(x.x); hi.y = tf32_rna(x.y); hi.z = tf32_rna(x.z); hi.w = tf32_rna(x.w);
  lo.x = tf32_rna(x.x - hi.x); lo.y = tf32_rna(x.y - hi.y); lo.z = tf32_rna(x.z - hi.z); lo.w = tf32_rna(x.w - hi.w);
}

// ------------------------------------------------------------------------------------------------
// B images
// ------------------------------------------------------------------------------------------------
struct PrepTask {
  const float2* mat;   // KKc × MMc complex row-major
  float* image;        // [nchunk][hi|lo][NNp × KC] canonical K-major no-swizzle
  int KKc, MMc;        // complex dims
  int NNp;             // padded N (floats), multiple of 16
  int nchunk;          // K chunks of KC floats
  int last;            // 0: MID expansion, 1: LAST expansion
};

// K-major / no-swizzle tile [rows × KC floats]: 8-row × 16-byte core matrices, the KC/4 core matrices
// of a row group adjacent (LBO = 128 B), row groups SBO = KC/4·128 B apart
__host__ __device__ inline int kmajor_off(int row, int k) { return (row % 8) * 4 + (k % 4) + (k / 4) * 32 + (row / 8) * (KC / 4) * 32; }

__global__ void __launch_bounds__(256) tc_prep_b_kernel(const PrepTask* __restrict__ tasks) {
  const PrepTask t = tasks[blockIdx.x];
  const int per = t.NNp * KC;
  for (int idx = threadIdx.x; idx < t.nchunk * per; idx += blockDim.x) {
    const int ch = idx / per, r = idx - ch * per;
    const int n = r / KC, kk = r - n * KC;
    const int k = ch * KC + kk;  // real K index
    float v = 0.f;
    if (!t.last) {
      // B[(j',part)][j]: part 0 → Re Mat[j][j'], part 1 → Im
      const int jp = n >> 1, part = n & 1;
      if (k < t.KKc && jp < t.MMc) { const float2 m = t.mat[(long long)k * t.MMc + jp]; v = part ? m.y : m.x; }
    } else {
      // B̂[(j',ri')][(j,ri)] = [[Mr, −Mi],[Mi, Mr]] (rows ri', columns ri)
      const int jp = n >> 1, rip = n & 1, j = k >> 1, ri = k & 1;
      if (j < t.KKc && jp < t.MMc) {
        const float2 m = t.mat[(long long)j * t.MMc + jp];
        v = (rip == ri) ? m.x : (rip ? m.y : -m.y);
      }
    }
    const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
    float* base = t.image + (long long)ch * 2 * per;
    base[kmajor_off(n, kk)] = hi;
    base[per + kmajor_off(n, kk)] = lo;
  }
}

// ------------------------------------------------------------------------------------------------
// mode product
// ------------------------------------------------------------------------------------------------
struct TcModeTask {
  const float2* in;
  float2* out;
  const float* image;     // B images
  long long ips, ops;     // plane strides (complex elements)
  int chi_in, chi_out;    // complex
  int P_in, P_out;
  int KKc, MMc;           // P_in*chi_in, P_out*chi_out
  int NNp;                // padded 2*MMc
  int nchunk;
  unsigned outer, inner, CC;
  int ntiles;             // ceil(CC / cols per tile)
  int tiles_per_cta;
  int cta_begin;          // first CTA (blockIdx.x) of this task
};

// smem: [ B images (nchunk·2·NNp·KC floats) | A_hi (128·KC) | A_lo (128·KC) | output staging ]
// 256 threads: all stage A (4 float4 each, next tile prefetched into registers while the current one
// is multiplied and written back); thread 0 issues the MMAs; warp w drains TMEM lane quadrant w&3,
// column half w>>2, through a shared-memory staging tile so that global stores are 16-byte coalesced.
constexpr int TC_THREADS = 256;

template <bool LAST>
__device__ __forceinline__ void tc_load_tile(const TcModeTask& t, int ch, unsigned c0, int tid, float4 (&x)[4]) {
  if (!LAST) {
    // 32 k-rows × 32 float4; this thread: float4 column l = tid&31 (complex columns c0+2l, c0+2l+1),
    // k-rows kl = (tid>>5) + 8j
    const int l = tid & 31;
    const unsigned col = c0 + 2 * l;
    const bool cv = col < t.CC;
    const unsigned o = cv ? col / t.inner : 0, n = cv ? col - o * t.inner : 0;
    const float2* base = t.in + (long long)o * t.chi_in * t.inner + n;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = ch * KC + (tid >> 5) + 8 * j;
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cv && k < t.KKc) {
        const int p = (t.P_in == 1) ? 0 : k / t.chi_in, b = k - p * t.chi_in;
        x[j] = __ldg(reinterpret_cast<const float4*>(base + p * t.ips + (long long)b * t.inner));
      }
    }
  } else {
    // 128 rows × 8 float4; this thread: row%8 = tid&7, float4 q = (tid>>3)&7, row groups rg = (tid>>6) + 4j
    const int l8 = tid & 7, q = (tid >> 3) & 7;
    const int kf = ch * KC + q * 4;  // first real k of the float4
    const int KKf = 2 * t.KKc;
    const int jc = kf >> 1;
    const int p = (t.P_in == 1) ? 0 : jc / t.chi_in, b = jc - p * t.chi_in;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = ((tid >> 6) + 4 * j) * 8 + l8;
      const unsigned col = c0 + row;
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < t.CC && kf < KKf) {
        const float2* src = t.in + p * t.ips + (long long)col * t.chi_in + b;
        if (b + 1 < t.chi_in && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          x[j] = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          const float2 a0 = src[0];
          x[j].x = a0.x; x[j].y = a0.y;
          const int jc1 = jc + 1;
          if (jc1 < t.KKc) {
            const int p1 = jc1 / t.chi_in, b1 = jc1 - p1 * t.chi_in;
            const float2 a1 = t.in[p1 * t.ips + (long long)col * t.chi_in + b1];
            x[j].z = a1.x; x[j].w = a1.y;
          }
        }
      }
    }
  }
}

template <bool LAST>
__device__ __forceinline__ void tc_store_stage(int tid, const float4 (&x)[4], float* sAh, float* sAl) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 hi, lo;
    split4(x[j], hi, lo);
    int off;
    if (!LAST) {
      const int l = tid & 31, kl = (tid >> 5) + 8 * j;
      const int r = kl & 3, ak = kl >> 2, am = l >> 3, c = (l & 7) >> 1, half = l & 1;
      off = (am * 512 + ak * 2048 + r * 128 + ((c ^ r) * 32) + half * 16) >> 2;
    } else {
      const int l8 = tid & 7, q = (tid >> 3) & 7, rg = (tid >> 6) + 4 * j;
      off = l8 * 4 + q * 32 + rg * 256;
    }
    *reinterpret_cast<float4*>(sAh + off) = hi;
    *reinterpret_cast<float4*>(sAl + off) = lo;
  }
}

template <bool LAST>
__global__ void __launch_bounds__(TC_THREADS) tc_mode_kernel(const TcModeTask* __restrict__ tasks, const int* __restrict__ cta_task) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar_mma, s_bar_b;
  const TcModeTask t = tasks[cta_task[blockIdx.x]];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b_floats = t.nchunk * 2 * t.NNp * KC;
  float* sB = smem;
  float* sAh = smem + ((b_floats + 255) & ~255);  // 1024-byte aligned: the swizzle is address based
  float* sAl = sAh + 128 * KC;
  // The output staging tile aliases the A stage: once the tile's MMAs have completed the stage is dead
  // (the next tile waits in registers).  MID: [NNp/2][128] floats; LAST: [128][NNp] floats with the
  // float4 column XOR-swizzled by (row & 7) to keep the row-per-lane stores conflict free.
  float* sOut = sAh;
  int tmem_cols = 32;
  while (tmem_cols < t.NNp) tmem_cols <<= 1;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar_mma)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar_b)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {  // one TMA bulk copy brings the whole B image set; completion on s_bar_b
    const uint32_t bytes = (uint32_t)b_floats * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar_b)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)),
                 "l"(t.image), "r"(bytes), "r"(smem_u32(&s_bar_b))
                 : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = make_idesc(128, t.NNp, LAST ? 0 : 1, 0);
  const int per_b = t.NNp * KC;
  uint32_t phase = 0;
  bool b_ready = false;

  // tiles are dealt round-robin to the CTAs of a task (tile = cta_local + i·ncta): CTAs resident together
  // stream neighbouring 512-byte pieces of the same tensor rows (same DRAM pages)
  const int cta_local = blockIdx.x - t.cta_begin;
  const int ncta_task = (t.ntiles + t.tiles_per_cta - 1) / t.tiles_per_cta;
  const int tile0 = cta_local;
  const int tile1 = t.ntiles;
  constexpr int COLS = LAST ? 128 : 64;  // complex columns per tile

  float4 xr[4];
  if (tile0 < tile1) tc_load_tile<LAST>(t, 0, (unsigned)tile0 * COLS, tid, xr);

  for (int tile = tile0; tile < tile1; tile += ncta_task) {
    const unsigned c0 = (unsigned)tile * COLS;
    for (int ch = 0; ch < t.nchunk; ++ch) {
      tc_store_stage<LAST>(tid, xr, sAh, sAl);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      // prefetch the next chunk / tile into registers while this one is multiplied
      {
        int nch = ch + 1, ntile = tile;
        if (nch == t.nchunk) { nch = 0; ntile = tile + ncta_task; }
        if (ntile < tile1) tc_load_tile<LAST>(t, nch, (unsigned)ntile * COLS, tid, xr);
      }
      if (!b_ready) { mbar_wait(smem_u32(&s_bar_b), 0); b_ready = true; }
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        const float* bh = sB + (long long)ch * 2 * per_b;
        const float* bl = bh + per_b;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a_base = smem_u32(term == 2 ? sAl : sAh);
          const uint32_t b_base = smem_u32(term == 1 ? bl : bh);
#pragma unroll
          for (int ks = 0; ks < KC / 8; ++ks) {
            uint64_t ad;
            if (!LAST) ad = make_desc(a_base + ks * 2 * 2048, 512, 2048, 1);  // MN-major SW128_32B: LBO = MN-atom, SBO = K-atom stride
            else ad = make_desc(a_base + ks * 256, 128, 1024, 0);             // K-major no swizzle: LBO = k-chunk, SBO = row-group stride
            const uint64_t bd = make_desc(b_base + ks * 256, 128, 1024, 0);
            mma_tf32(tmem, ad, bd, idesc, (ch | term | ks) != 0);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar_mma)) : "memory");
      }
      mbar_wait(smem_u32(&s_bar_mma), phase);
      phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // ---- epilogue: TMEM → registers → staging tile → coalesced global stores ---------------------
    const int quad = warp & 3, half = warp >> 2;
    // XOR swizzle of the staging tile's float4 columns: stay inside an aligned power-of-two group
    const int nf4 = t.NNp >> 2;
    const int swmask = min(7, (nf4 & -nf4) - 1);
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    const int nhalf = (t.NNp / 16 + 1) / 2;  // 16-column groups handled by half 0
    const int g0 = half ? nhalf : 0, g1 = half ? t.NNp / 16 : nhalf;
    if (!LAST) {
      const int f = quad * 32 + lane;  // row of D = float index (col_local, ri) within the tile
      const float sgn = (lane & 1) ? 1.f : -1.f;
      for (int g = g0; g < g1; ++g) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(trow + g * 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float own = __uint_as_float(v[2 * q]);
          const float other = __shfl_xor_sync(0xffffffffu, __uint_as_float(v[2 * q + 1]), 1);
          sOut[(g * 8 + q) * 128 + f] = own + sgn * other;  // [jp][(col,ri)]
        }
      }
    } else {
      const int row = quad * 32 + lane;
      float* dst = sOut + row * t.NNp;
      const int sw = row & swmask;
      for (int g = g0; g < g1; ++g) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(trow + g * 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 y;
          y.x = __uint_as_float(v[4 * q]); y.y = __uint_as_float(v[4 * q + 1]);
          y.z = __uint_as_float(v[4 * q + 2]); y.w = __uint_as_float(v[4 * q + 3]);
          *reinterpret_cast<float4*>(dst + (((g * 4 + q) ^ sw) << 2)) = y;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();  // staging tile complete; TMEM free for the next tile's MMAs
    if (!LAST) {
      // rows jp of the staging tile are 64 complex columns = 512 contiguous bytes of the output
      const int l = tid & 31;
      const unsigned col = c0 + 2 * l;
      if (col < t.CC) {
        const unsigned o = col / t.inner, n = col - o * t.inner;
        float2* obase = t.out + (long long)o * t.chi_out * t.inner + n;
        for (int jp = tid >> 5; jp < t.MMc; jp += TC_THREADS / 32) {
          const int pp = (t.P_out == 1) ? 0 : jp / t.chi_out, c = jp - pp * t.chi_out;
          const float4 y = *reinterpret_cast<const float4*>(sOut + jp * 128 + 4 * l);
          *reinterpret_cast<float4*>(obase + pp * t.ops + (long long)c * t.inner) = y;
        }
      }
    } else {
      // the tile's output rows are contiguous in global memory when there is one output plane
      const int w4 = (2 * t.MMc) >> 2;  // float4 per output row (2·MMc floats); MMc even → exact
      const int total = 128 * w4;
      const bool vec_ok = ((2 * t.MMc) & 3) == 0 && t.P_out == 1 && ((reinterpret_cast<uintptr_t>(t.out) & 15) == 0);
      if (vec_ok) {
        for (int idx = tid; idx < total; idx += TC_THREADS) {
          const int row = idx / w4, q = idx - row * w4;
          const unsigned col = c0 + row;
          if (col < t.CC) {
            const float4 y = *reinterpret_cast<const float4*>(sOut + row * t.NNp + ((q ^ (row & swmask)) << 2));
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(t.out + (long long)col * t.chi_out) + 4 * q) = y;
          }
        }
      } else {
        for (int idx = tid; idx < 128 * t.MMc; idx += TC_THREADS) {
          const int row = idx / t.MMc, jp = idx - row * t.MMc;
          const unsigned col = c0 + row;
          if (col < t.CC) {
            const int pp = jp / t.chi_out, c = jp - pp * t.chi_out;
            float2 y;
            const int f = 2 * jp;
            const float* src = sOut + row * t.NNp + ((((f >> 2) ^ (row & swmask)) << 2) | (f & 3));
            y.x = src[0];
            y.y = src[1];
            t.out[pp * t.ops + (long long)col * t.chi_out + c] = y;
          }
        }
      }
    }
    __syncthreads();  // copy-out done: the staging tile (= A stage) may be overwritten by the next tile
  }
  if (!b_ready) mbar_wait(smem_u32(&s_bar_b), 0);  // never leave a bulk copy in flight
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

// ------------------------------------------------------------------------------------------------
// Gram:  partial[split][i][j] = Σ_{col ∈ split} conj(X[i][col]) · Y[j][col]      (one plane, i,j < χ ≤ 64)
// ------------------------------------------------------------------------------------------------
//   MID  (inner ≥ 16): per outer slice the rows X[i,:], Y[j,:] are K-contiguous real vectors of
//        (n,ri) floats.  D[(j,part), i] = Σ_k Ycat[(j,part),k] · X[i,k] with Ycat[(j,0)] = Y[j],
//        Ycat[(j,1)] = (Yi, −Yr) rotated copy built in registers; re = part 0, im = part 1.
//        Both operands K-major / no swizzle.
//   LAST (inner == 1): each column is a contiguous row of (i,ri) floats → both operands MN-major
//        (SWIZZLE_128B_BASE32B).  D[(i,ri),(j,rj)]: re = D[(i,0),(j,0)] + D[(i,1),(j,1)],
//        im = D[(i,0),(j,1)] − D[(i,1),(j,0)].
// K is streamed through two shared-memory stages; the accumulator stays in TMEM for the whole split
// and is drained once.  M is 64 (χ ≤ 32) or 128.
struct TcGramTask {
  const float2* X;
  const float2* Y;
  double2* partial;        // [nsplit][chi*chi]
  int chi;
  int MMp, NNp;            // padded MMA M and N (floats)
  unsigned outer, inner, CC;
  int nsplit;
  unsigned cols_per_split;
  int cta_begin;
};

// NIT = (X,Y) float4 pairs a thread stages per K step, PD = how many K steps ahead the global loads
// run (register prefetch ring): the DRAM latency of step st+PD overlaps the split / MMA of step st.
template <bool LAST, int NIT, int PD>
__global__ void __launch_bounds__(TC_THREADS) tc_gram_kernel(const TcGramTask* __restrict__ tasks, const int* __restrict__ cta_task) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[2];
  const TcGramTask t = tasks[cta_task[blockIdx.x]];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x - t.cta_begin;
  const int chi = t.chi;
  // stage layout (floats): [A_hi | A_lo | B_hi | B_lo], A = MMp×KC, B = NNp×KC
  const int a_floats = t.MMp * KC, b_floats = t.NNp * KC;
  const int stage_floats = 2 * a_floats + 2 * b_floats;
  int tmem_cols = 32;
  while (tmem_cols < t.NNp) tmem_cols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero both stages once: padded rows / columns must read as zeros
  for (int i = tid; i < 2 * stage_floats; i += TC_THREADS) smem[i] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = make_idesc(t.MMp, t.NNp, LAST ? 1 : 0, LAST ? 1 : 0);

  // K-steps are dealt round-robin to the nsplit CTAs of a task (step s of this CTA = global step
  // s·nsplit + split): CTAs that are resident together then read neighbouring 128-byte pieces of the same
  // tensor rows, i.e. the same DRAM pages, instead of nsplit·2χ far-apart sequential streams per task.
  constexpr int SCOLS = LAST ? 32 : 16;  // complex columns per stage (KC real K either way)
  const unsigned ce = t.CC;
  const int total_steps = (int)((t.CC + SCOLS - 1) / SCOLS);
  const int nstage = split < total_steps ? (total_steps - split + t.nsplit - 1) / t.nsplit : 0;
  uint32_t ph[2] = {0, 0};
  int used[2] = {0, 0};

  // ---- per-thread constants of the staging pattern ----------------------------------------------------
  // MID : item (row, q): row = (idx&7) + 8·(idx>>6), q = (idx>>3)&7 (row%8 fastest → conflict-free stores);
  //       global element offset inside an outer slice: row·inner + 2q
  // LAST: item (kl, l): column kl of the stage, float4 l of its 2χ-float row
  int it_row[NIT], it_q[NIT];       // MID: row, q        LAST: kl, l
  int offA[NIT], offB[NIT];         // shared-memory float offsets
  bool it_ok[NIT];
  const int w4 = chi >> 1;
  const int sboA = (t.MMp >> 5) * 512, sboB = (t.NNp >> 5) * 512;
#pragma unroll
  for (int j = 0; j < NIT; ++j) {
    const int idx = tid + j * TC_THREADS;
    if (!LAST) {
      const int row = (idx & 7) + 8 * (idx >> 6), q = (idx >> 3) & 7;
      it_row[j] = row; it_q[j] = q; it_ok[j] = row < chi;
      const int m0 = 2 * row;  // Ycat rows (j,0) and (j,1)
      offA[j] = (m0 & 7) * 4 + q * 32 + (m0 >> 3) * 256;
      offB[j] = (row & 7) * 4 + q * 32 + (row >> 3) * 256;
    } else {
      const int kl = idx / w4, l = idx - kl * w4;
      it_row[j] = kl; it_q[j] = l; it_ok[j] = kl < SCOLS;
      const int r = kl & 3, ak = kl >> 2, am = l >> 3, c = (l & 7) >> 1, half = l & 1;
      const int inatom = r * 128 + ((c ^ r) * 32) + half * 16;
      offA[j] = (am * 512 + ak * sboA + inatom) >> 2;
      offB[j] = (am * 512 + ak * sboB + inatom) >> 2;
    }
  }

  float4 rx[PD][NIT], ry[PD][NIT];
  auto load_stage = [&](int st, float4 (&x)[NIT], float4 (&y)[NIT]) {
    const unsigned k0 = ((unsigned)st * (unsigned)t.nsplit + (unsigned)split) * SCOLS;
    if (!LAST) {
      // 16 complex columns of one outer slice (inner % 16 == 0 and every step starts on a multiple of 16)
      const unsigned o = k0 / t.inner, n0 = k0 - o * t.inner;
      const int nvalid = (int)min((unsigned)SCOLS, ce - k0);  // multiple of 2 by construction
      const long long base = (long long)o * chi * t.inner + n0;
#pragma unroll
      for (int j = 0; j < NIT; ++j) {
        x[j] = make_float4(0.f, 0.f, 0.f, 0.f); y[j] = x[j];
        if (it_ok[j] && 2 * it_q[j] < nvalid) {
          const long long a = base + (long long)it_row[j] * t.inner + 2 * it_q[j];
          y[j] = __ldg(reinterpret_cast<const float4*>(t.Y + a));
          x[j] = __ldg(reinterpret_cast<const float4*>(t.X + a));
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NIT; ++j) {
        x[j] = make_float4(0.f, 0.f, 0.f, 0.f); y[j] = x[j];
        const unsigned col = k0 + it_row[j];
        if (it_ok[j] && col < ce) {
          const long long a = (long long)col * chi + 2 * it_q[j];
          y[j] = __ldg(reinterpret_cast<const float4*>(t.Y + a));
          x[j] = __ldg(reinterpret_cast<const float4*>(t.X + a));
        }
      }
    }
  };
#pragma unroll
  for (int p = 0; p < PD; ++p)
    if (p < nstage) load_stage(p, rx[p], ry[p]);

  for (int st0 = 0; st0 < nstage; st0 += PD) {
#pragma unroll
    for (int p = 0; p < PD; ++p) {
      const int st = st0 + p;
      if (st >= nstage) break;
      const int bsel = st & 1;
      float* sA = smem + bsel * stage_floats;
      float* sAl = sA + a_floats;
      float* sBh = sAl + a_floats;
      float* sBl = sBh + b_floats;
      if (used[bsel]) { mbar_wait(smem_u32(&s_bar[bsel]), ph[bsel]); ph[bsel] ^= 1; }  // MMAs that read this stage are done
#pragma unroll
      for (int j = 0; j < NIT; ++j) {
        if (!it_ok[j]) continue;
        const float4 x = rx[p][j], y = ry[p][j];
        float4 hi, lo;
        if (!LAST) {
          split4(y, hi, lo);
          *reinterpret_cast<float4*>(sA + offA[j]) = hi;
          *reinterpret_cast<float4*>(sAl + offA[j]) = lo;
          const float4 yr = make_float4(y.y, -y.x, y.w, -y.z);  // (Yi, −Yr)
          split4(yr, hi, lo);
          *reinterpret_cast<float4*>(sA + offA[j] + 4) = hi;  // row m0+1 (m0 even → same 8-row group)
          *reinterpret_cast<float4*>(sAl + offA[j] + 4) = lo;
          split4(x, hi, lo);
          *reinterpret_cast<float4*>(sBh + offB[j]) = hi;
          *reinterpret_cast<float4*>(sBl + offB[j]) = lo;
        } else {
          split4(x, hi, lo);  // A = X (conjugated side: rows (i,ri))
          *reinterpret_cast<float4*>(sA + offA[j]) = hi;
          *reinterpret_cast<float4*>(sAl + offA[j]) = lo;
          split4(y, hi, lo);
          *reinterpret_cast<float4*>(sBh + offB[j]) = hi;
          *reinterpret_cast<float4*>(sBl + offB[j]) = lo;
        }
      }
      if (st + PD < nstage) load_stage(st + PD, rx[p], ry[p]);  // refill this ring slot
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a_base = smem_u32(term == 2 ? sAl : sA);
          const uint32_t b_base = smem_u32(term == 1 ? sBl : sBh);
#pragma unroll
          for (int ks = 0; ks < KC / 8; ++ks) {
            uint64_t ad, bd;
            if (!LAST) {
              ad = make_desc(a_base + ks * 256, 128, 1024, 0);
              bd = make_desc(b_base + ks * 256, 128, 1024, 0);
            } else {
              ad = make_desc(a_base + ks * 2 * sboA, 512, sboA, 1);
              bd = make_desc(b_base + ks * 2 * sboB, 512, sboB, 1);
            }
            mma_tf32(tmem, ad, bd, idesc, (st | term | ks) != 0);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[bsel])) : "memory");
      }
      used[bsel] = 1;
    }
  }
  // drain: wait for the last commit of each stage buffer (commits complete in issue order)
  for (int bsel = 0; bsel < 2; ++bsel)
    if (used[bsel]) { mbar_wait(smem_u32(&s_bar[bsel]), ph[bsel]); ph[bsel] ^= 1; }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // ---- epilogue (once per CTA): TMEM → partial sums -------------------------------------------------
  double2* __restrict__ P = t.partial + (long long)split * chi * chi;
  if (warp < 4) {
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    int m;  // row of D held by this lane
    if (t.MMp == 128) m = warp * 32 + lane;
    else m = (lane < 16) ? warp * 16 + lane : -1;  // M = 64: rows 16w..16w+15 live in lanes 0..15 of quadrant w
    for (int n0 = 0; n0 < t.NNp; n0 += 16) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(trow + n0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!LAST) {
        // D[(j,part), i]: this lane holds row (j,part); columns i = n0..n0+15
        if (m >= 0) {
          const int j = m >> 1, part = m & 1;
          if (j < chi) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int i = n0 + q;
              if (i < chi) reinterpret_cast<double*>(&P[(long long)i * chi + j])[part] = (double)__uint_as_float(v[q]);
            }
          }
        }
      } else {
        // D[(i,ri),(j,rj)]: even lane (i,0) makes re, odd lane (i,1) makes im
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float own = __uint_as_float(v[2 * q]);
          const float other = __shfl_xor_sync(0xffffffffu, __uint_as_float(v[2 * q + 1]), 1);
          if (m >= 0) {
            const int i = m >> 1, ri = m & 1, j = (n0 >> 1) + q;
            if (i < chi && j < chi) {
              const float val = ri ? (other - own) : (own + other);
              reinterpret_cast<double*>(&P[(long long)i * chi + j])[ri] = (double)val;
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

}  // namespace tc
}  // namespace tnqs
