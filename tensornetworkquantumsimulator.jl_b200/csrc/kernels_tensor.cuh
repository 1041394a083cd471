// kernels_tensor.cuh — the kernels that stream site tensors (d·χ^z elements) through the SMs.
//
// Every dense contraction of the BP / simple-update path (SURVEY.md §2b K2,K6,K7,K10,K12,K13) is
// one of two shapes once the site tensor T[s, l_0 … l_{z-1}] is viewed around one "active" leg as
// A[j][col], j = (plane, active index), col = (outer, inner):
//
//   mode product   Out[j'][col] = Σ_j  Mat[j][j'] · A[j][col]            (tensor × small matrix)
//   Gram           G[i][j]      = Σ_col conj(X[i][col]) · Y[j][col]      (tensor × tensor → small)
//
// Both are written as shared-memory tiled SIMT GEMMs (64×64 tile, 4×4 complex micro-tile per
// thread) with gather addressing of the tensor operand, templated on the real type.  This file is
// the exact-precision path (fp32 FFMA for ComplexF32, fp64 DFMA for ComplexF64 and for the
// reduced-factor Gram); kernels_tc.cuh holds the tcgen05 variants for ComplexF32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tnqs {

template <typename R> struct Cx;
template <> struct Cx<float> { using type = float2; };
template <> struct Cx<double> { using type = double2; };

template <typename C> __device__ __forceinline__ C c_zero() { C z; z.x = 0; z.y = 0; return z; }
// acc += a*b
template <typename C> __device__ __forceinline__ void c_fma(C& acc, const C a, const C b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
template <typename C> __device__ __forceinline__ void c_fma_conj(C& acc, const C a, const C b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}

constexpr int TK = 16;

// ------------------------------------------------------------------------------------------------
// mode product
// ------------------------------------------------------------------------------------------------
struct ModeTask {
  const void* in;
  void* out;
  const void* mat;        // [(p,b)][(p',c)] row-major, KK × MM, tensor scalar type
  long long ips, ops;     // plane strides of in / out (elements)
  int chi_in, chi_out;    // active-leg dimension before / after
  int KK, MM;             // P_in*chi_in, P_out*chi_out
  unsigned outer, inner;  // per-plane view [outer][chi][inner]
  unsigned CC;            // outer*inner
  int tiles_m, tiles_c;
};

// Out[p', o, c, n] = Σ_{p,b} In[p, o, b, n] · Mat[(p,b), (p',c)]
// Tile TMv (output active index) × TCv (columns), 256 threads, 4×4 complex micro-tile.
// INNER1: the active leg is the innermost one (inner == 1) — the contiguous direction is then the
// active index itself, so loads run k-fastest and stores m-fastest to stay coalesced.
template <typename R, bool INNER1, int TMv, int TCv>
__global__ void __launch_bounds__(256) mode_product_kernel(const ModeTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  constexpr int NT = 256;
  static_assert((TMv / 4) * (TCv / 4) == NT, "tile/threads mismatch");
  constexpr int FX = INNER1 ? TMv / 4 : TCv / 4;
  const ModeTask t = tasks[blockIdx.y];
  const int tile = blockIdx.x;
  if (tile >= t.tiles_m * t.tiles_c) return;
  const int m0 = (tile % t.tiles_m) * TMv;
  const unsigned c0 = (unsigned)(tile / t.tiles_m) * TCv;
  const C* __restrict__ in = (const C*)t.in;
  const C* __restrict__ mat = (const C*)t.mat;
  C* __restrict__ out = (C*)t.out;

  __shared__ C As[TK][TMv + 1];
  __shared__ C Bs[TK][TCv + 1];
  const int tid = threadIdx.x;
  const int tx = tid % FX, ty = tid / FX;

  C acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = c_zero<C>();

  for (int k0 = 0; k0 < t.KK; k0 += TK) {
    // A tile: Mat[k0+kk][m0+mm]
#pragma unroll
    for (int r = 0; r < (TK * TMv) / NT; ++r) {
      const int idx = tid + r * NT;
      const int mm = idx % TMv, kk = idx / TMv;
      const int k = k0 + kk, m = m0 + mm;
      C v = c_zero<C>();
      if (k < t.KK && m < t.MM) v = mat[(long long)k * t.MM + m];
      As[kk][mm] = v;
    }
    // B tile: In[k0+kk][c0+cc]
#pragma unroll
    for (int r = 0; r < (TK * TCv) / NT; ++r) {
      const int idx = tid + r * NT;
      int kk, cc;
      if (INNER1) { kk = idx % TK; cc = idx / TK; } else { cc = idx % TCv; kk = idx / TCv; }
      const int k = k0 + kk;
      const unsigned col = c0 + cc;
      C v = c_zero<C>();
      if (k < t.KK && col < t.CC) {
        const int p = k / t.chi_in, b = k - p * t.chi_in;
        long long a;
        if (INNER1) {
          a = p * t.ips + (long long)col * t.chi_in + b;
        } else {
          const unsigned o = col / t.inner, n = col - o * t.inner;
          a = p * t.ips + ((long long)o * t.chi_in + b) * t.inner + n;
        }
        v = in[a];
      }
      Bs[kk][cc] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      C a[4], b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = INNER1 ? As[kk][tx + FX * j] : As[kk][4 * ty + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = INNER1 ? Bs[kk][4 * ty + i] : Bs[kk][tx + FX * i];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) c_fma(acc[j][i], a[j], b[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = m0 + (INNER1 ? tx + FX * j : 4 * ty + j);
    if (m >= t.MM) continue;
    const int pp = m / t.chi_out, c = m - pp * t.chi_out;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned col = c0 + (INNER1 ? 4 * ty + i : tx + FX * i);
      if (col >= t.CC) continue;
      long long a;
      if (INNER1) {
        a = pp * t.ops + (long long)col * t.chi_out + c;
      } else {
        const unsigned o = col / t.inner, n = col - o * t.inner;
        a = pp * t.ops + ((long long)o * t.chi_out + c) * t.inner + n;
      }
      out[a] = acc[j][i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Gram
// ------------------------------------------------------------------------------------------------
struct GramTask {
  const void* X;          // conjugated operand
  const void* Y;
  long long xps, yps;     // plane strides (elements)
  int chi, MM;            // active-leg dim; MM = P*chi
  unsigned outer, inner, CC;
  int tiles;              // ceil(MM/TI) per side
  int nsplit;
  unsigned cols_per_split;
  double2* partial;       // [nsplit][MM*MM], row-major [i][j]
};

// partial[split][i][j] = Σ_{col in split} conj(X[i][col]) · Y[j][col]
// Tile TI×TI outputs per CTA, (TI/4)² threads, 4×4 micro-tile; AccR is the accumulation type.
template <typename R, typename AccR, bool INNER1, int TI>
__global__ void __launch_bounds__((TI / 4) * (TI / 4)) gram_kernel(const GramTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  using CA = typename Cx<AccR>::type;
  constexpr int NT = (TI / 4) * (TI / 4);
  constexpr int FX = TI / 4;
  const GramTask t = tasks[blockIdx.z];
  const int split = blockIdx.x;
  const int tile = blockIdx.y;
  if (split >= t.nsplit || tile >= t.tiles * t.tiles) return;
  const int i0 = (tile / t.tiles) * TI, j0 = (tile % t.tiles) * TI;
  const C* __restrict__ X = (const C*)t.X;
  const C* __restrict__ Y = (const C*)t.Y;
  const unsigned cb = (unsigned)split * t.cols_per_split;
  unsigned ce = cb + t.cols_per_split;
  if (ce > t.CC) ce = t.CC;

  __shared__ CA Xs[TK][TI + 1];
  __shared__ CA Ys[TK][TI + 1];
  const int tid = threadIdx.x;
  const int tx = tid % FX, ty = tid / FX;
  CA acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = c_zero<CA>();

  for (unsigned k0 = cb; k0 < ce; k0 += TK) {
#pragma unroll
    for (int r = 0; r < (TK * TI) / NT; ++r) {
      const int idx = tid + r * NT;
      int kk, ii;
      if (INNER1) { ii = idx % TI; kk = idx / TI; } else { kk = idx % TK; ii = idx / TK; }
      const unsigned col = k0 + kk;
      CA xv = c_zero<CA>(), yv = c_zero<CA>();
      if (col < ce) {
        unsigned o = 0, n = 0;
        if (!INNER1) { o = col / t.inner; n = col - o * t.inner; }
        const int i = i0 + ii;
        if (i < t.MM) {
          const int p = i / t.chi, l = i - p * t.chi;
          const long long a = INNER1 ? p * t.xps + (long long)col * t.chi + l
                                     : p * t.xps + ((long long)o * t.chi + l) * t.inner + n;
          const C v = X[a];
          xv.x = (AccR)v.x; xv.y = (AccR)v.y;
        }
        const int j = j0 + ii;
        if (j < t.MM) {
          const int p = j / t.chi, l = j - p * t.chi;
          const long long a = INNER1 ? p * t.yps + (long long)col * t.chi + l
                                     : p * t.yps + ((long long)o * t.chi + l) * t.inner + n;
          const C v = Y[a];
          yv.x = (AccR)v.x; yv.y = (AccR)v.y;
        }
      }
      Xs[kk][ii] = xv;
      Ys[kk][ii] = yv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      CA x[4], y[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) x[a] = Xs[kk][tx + FX * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) y[b] = Ys[kk][4 * ty + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) c_fma_conj(acc[a][b], x[a], y[b]);
    }
    __syncthreads();
  }
  double2* __restrict__ P = t.partial + (long long)split * t.MM * t.MM;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + tx + FX * a;
    if (i >= t.MM) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j0 + 4 * ty + b;
      if (j >= t.MM) continue;
      double2 v; v.x = (double)acc[a][b].x; v.y = (double)acc[a][b].y;
      P[(long long)i * t.MM + j] = v;
    }
  }
}

struct ReduceTask {
  const double2* partial;
  double2* out;
  int nsplit, MM;
  int transpose;  // out[j*MM+i] = Σ partial[i*MM+j]
};

// fixed-order (deterministic) sum over the K-splits
__global__ void gram_reduce_kernel(const ReduceTask* __restrict__ tasks) {
  const ReduceTask t = tasks[blockIdx.y];
  const int n2 = t.MM * t.MM;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n2; idx += gridDim.x * blockDim.x) {
    double sx = 0, sy = 0;
    for (int s = 0; s < t.nsplit; ++s) {
      const double2 v = t.partial[(long long)s * n2 + idx];
      sx += v.x; sy += v.y;
    }
    const int i = idx / t.MM, j = idx - i * t.MM;
    double2 r; r.x = sx; r.y = sy;
    t.out[t.transpose ? (j * t.MM + i) : idx] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// Frobenius norm + scale (simple_update.jl:70-74) and the one-site gate (simple_update.jl:26-28)
// ------------------------------------------------------------------------------------------------
constexpr int NORM_BLOCKS = 64;

struct NormTask {
  void* data;
  long long n;       // complex elements
  double* partial;   // [NORM_BLOCKS]
};

template <typename R>
__global__ void __launch_bounds__(256) sumsq_kernel(const NormTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  const NormTask t = tasks[blockIdx.y];
  const C* __restrict__ d = (const C*)t.data;
  double s = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.n;
       i += (long long)gridDim.x * blockDim.x) {
    const C v = d[i];
    s += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) t.partial[blockIdx.x] = red[0];
}

template <typename R>
__global__ void __launch_bounds__(256) scale_kernel(const NormTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  const NormTask t = tasks[blockIdx.y];
  double s = 0;
  for (int i = 0; i < NORM_BLOCKS; ++i) s += t.partial[i];
  if (!(s > 0)) return;
  const R f = (R)(1.0 / sqrt(s));
  C* __restrict__ d = (C*)t.data;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.n;
       i += (long long)gridDim.x * blockDim.x) {
    C v = d[i];
    v.x *= f; v.y *= f;
    d[i] = v;
  }
}

struct OneSiteTask {
  void* data;        // in place: T[s'][n] = Σ_s U[s'][s] T[s][n]
  long long plane;   // elements per physical plane
  int d;
  double2 U[16];     // d×d row-major, d ≤ 4
};

template <typename R>
__global__ void __launch_bounds__(256) onesite_kernel(const OneSiteTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  const OneSiteTask& t = tasks[blockIdx.y];
  C* __restrict__ data = (C*)t.data;
  const int d = t.d;
  C U[16];
  for (int i = 0; i < d * d; ++i) { U[i].x = (R)t.U[i].x; U[i].y = (R)t.U[i].y; }
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < t.plane;
       n += (long long)gridDim.x * blockDim.x) {
    C in[4], o[4];
    for (int s = 0; s < d; ++s) in[s] = data[s * t.plane + n];
    for (int sp = 0; sp < d; ++sp) {
      o[sp] = c_zero<C>();
      for (int s = 0; s < d; ++s) c_fma(o[sp], U[sp * d + s], in[s]);
    }
    for (int sp = 0; sp < d; ++sp) data[sp * t.plane + n] = o[sp];
  }
}

// fill a χ×χ message with the identity / a diagonal (default_message, apply_gates.jl:126-136)
struct DiagTask {
  void* out;
  int chi;
  const double* diag;  // nullptr → identity
  const double* scale_sumsq;  // optional: divide diag by sqrt(*scale_sumsq)
};

template <typename R>
__global__ void diag_fill_kernel(const DiagTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  const DiagTask t = tasks[blockIdx.y];
  C* __restrict__ o = (C*)t.out;
  double f = 1.0;
  if (t.scale_sumsq) { const double s = *t.scale_sumsq; f = s > 0 ? 1.0 / sqrt(s) : 1.0; }
  const int n2 = t.chi * t.chi;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n2; idx += gridDim.x * blockDim.x) {
    const int i = idx / t.chi, j = idx - i * t.chi;
    C v = c_zero<C>();
    if (i == j) v.x = (R)(t.diag ? t.diag[i] * f : 1.0);
    o[idx] = v;
  }
}

}  // namespace tnqs
