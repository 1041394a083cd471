// kernels_small.cuh — the O(χ³) dense algebra of the path, batched one CTA per matrix, all in
// fp64 regardless of the state's scalar type (the reference itself promotes the message
// eigendecomposition to Float64: src/utils.jl:94-108).
//
//   jacobi_kernel        one-sided (Hestenes) Jacobi with round-robin pair ordering: complex SVD of
//                        θ (simple_update.jl:53-59), Hermitian eigendecomposition of BP messages
//                        (utils.jl:18-35) and of the reduced-factor Gram matrix (replaces the QR at
//                        simple_update.jl:47-48)
//   msg_* / su_* / bp_*  glue between the tensor-streaming kernels and the factorizations
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels_tensor.cuh"
#include "kernels_jacobi.cuh"

namespace tnqs {

__device__ __forceinline__ double2 z_mul(const double2 a, const double2 b) {
  double2 r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum of one double; every thread gets the result.  `red` needs ≥ 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double s = lane < nw ? red[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) red[32] = s;
  }
  __syncthreads();
  return red[32];
}

// ------------------------------------------------------------------------------------------------
// one-sided Jacobi, global-memory (L2) resident: the fallback for matrices too large for the
// shared-memory cluster kernel of kernels_jacobi.cuh (stacked rows > 256 or more than 256 columns)
// ------------------------------------------------------------------------------------------------
// RPL = rows held per lane: each lane keeps its rows of both columns of a pair in registers, so a
// pair costs one L2 round trip (all loads issued back to back) instead of one per row chunk.
template <int RPL>
__global__ void __launch_bounds__(RPL <= 4 ? 1024 : (RPL == 8 ? 512 : 256)) jacobi_kernel(const JacobiTask* __restrict__ tasks,
                                                      int max_sweeps, double tol, double* __restrict__ nonconv) {
  const JacobiTask t = tasks[blockIdx.x];
  const int n = t.n, m = t.m;
  const int ne = n + (n & 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double2* __restrict__ A = t.A;
  double2* __restrict__ V = t.V;
  __shared__ int s_rot;
  __shared__ double s_part[32];
  __shared__ double s_floor;
  __shared__ unsigned char s_dead[512];  // columns already known to be numerically null
  for (int j = threadIdx.x; j < 512; j += blockDim.x) s_dead[j] = 0;
  // orthogonality threshold ∝ √m·eps (round-off floor of an m-term dot product, as in LAPACK xGESVJ)
  const double tol2 = tol * tol * (double)(m > 1 ? m : 1);
  // Columns whose norm falls below 1e-20·‖A‖_F are numerically null (rank-deficient θ / Gram
  // matrices): rotating them only chases round-off and never terminates, so they are left alone.
  {
    double part = 0;
    for (long long idx = threadIdx.x; idx < (long long)m * n; idx += blockDim.x) {
      const double2 x = A[idx];
      part += x.x * x.x + x.y * x.y;
    }
    part = warp_sum(part);
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      double f = 0;
      for (int w = 0; w < nwarps; ++w) f += s_part[w];
      s_floor = 1e-40 * f;
    }
    __syncthreads();
  }
  const double floor2 = s_floor;
  if (n >= 2) {
    bool conv = false;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
      if (threadIdx.x == 0) s_rot = 0;
      __syncthreads();
      for (int step = 0; step < ne - 1; ++step) {
        for (int pair = warp; pair < ne / 2; pair += nwarps) {
          int p, q;
          if (pair == 0) { p = step; q = ne - 1; }
          else { p = (step + pair) % (ne - 1); q = (step - pair + (ne - 1)) % (ne - 1); }
          if (p >= n || q >= n) continue;
          if (s_dead[p] | s_dead[q]) continue;
          if (p > q) { const int tmp = p; p = q; q = tmp; }
          double2* __restrict__ ap = A + (long long)p * m;
          double2* __restrict__ aq = A + (long long)q * m;
          double2 xr[RPL], yr[RPL];
#pragma unroll
          for (int k = 0; k < RPL; ++k) {
            const int i = lane + 32 * k;
            if (i < m) { xr[k] = ap[i]; yr[k] = aq[i]; }
            else { xr[k].x = xr[k].y = 0; yr[k].x = yr[k].y = 0; }
          }
          double a = 0, b = 0, gx = 0, gy = 0;
#pragma unroll
          for (int k = 0; k < RPL; ++k) {
            const double2 x = xr[k], y = yr[k];
            a += x.x * x.x + x.y * x.y;
            b += y.x * y.x + y.y * y.y;
            gx += x.x * y.x + x.y * y.y;   // conj(x)*y
            gy += x.x * y.y - x.y * y.x;
          }
          a = warp_sum(a); b = warp_sum(b); gx = warp_sum(gx); gy = warp_sum(gy);
          const double g2 = gx * gx + gy * gy;
          if (lane == 0) {
            if (!(a > floor2)) s_dead[p] = 1;
            if (!(b > floor2)) s_dead[q] = 1;
          }
          if (g2 > tol2 * a * b && a > floor2 && b > floor2) {
            if (lane == 0) s_rot = 1;
            const double ig = rsqrt(g2);
            const double zeta = 0.5 * (b - a) * ig;
            const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = rsqrt(1.0 + tt * tt), s = c * tt;
            double2 ph; ph.x = gx * ig; ph.y = -gy * ig;  // e^{-iφ}
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
              const int i = lane + 32 * k;
              if (i < m) {
                const double2 x = xr[k];
                const double2 y = z_mul(yr[k], ph);
                double2 xn, yn;
                xn.x = c * x.x - s * y.x; xn.y = c * x.y - s * y.y;
                yn.x = s * x.x + c * y.x; yn.y = s * x.y + c * y.y;
                ap[i] = xn; aq[i] = yn;
              }
            }
            if (V) {
              double2* __restrict__ vp = V + (long long)p * n;
              double2* __restrict__ vq = V + (long long)q * n;
#pragma unroll
              for (int k = 0; k < RPL; ++k) {
                const int i = lane + 32 * k;
                if (i < n) { xr[k] = vp[i]; yr[k] = vq[i]; }
              }
#pragma unroll
              for (int k = 0; k < RPL; ++k) {
                const int i = lane + 32 * k;
                if (i < n) {
                  const double2 x = xr[k];
                  const double2 y = z_mul(yr[k], ph);
                  double2 xn, yn;
                  xn.x = c * x.x - s * y.x; xn.y = c * x.y - s * y.y;
                  yn.x = s * x.x + c * y.x; yn.y = s * x.y + c * y.y;
                  vp[i] = xn; vq[i] = yn;
                }
              }
            }
          }
        }
        __syncthreads();
      }
      const int rot = s_rot;
      __syncthreads();
      if (!rot) { conv = true; break; }
    }
    if (!conv && threadIdx.x == 0 && nonconv) nonconv[0] = 1.0;  // still rotating after max_sweeps: reported to the host
  }
  // column norms, then rank sort (descending, ties by index)
  for (int j = warp; j < n; j += nwarps) {
    double a = 0;
    const double2* aj = A + (long long)j * m;
    for (int i = lane; i < m; i += 32) { const double2 x = aj[i]; a += x.x * x.x + x.y * x.y; }
    a = warp_sum(a);
    if (lane == 0) t.sval[j] = sqrt(a);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double sj = t.sval[j];
    int rank = 0;
    for (int i = 0; i < n; ++i) {
      const double si = t.sval[i];
      rank += (si > sj) || (si == sj && i < j);
    }
    t.perm[rank] = j;
  }
}

// ------------------------------------------------------------------------------------------------
// BP message → (√M, projector) for the simple-update gauge (simple_update.jl:38-41, utils.jl:18-26)
// ------------------------------------------------------------------------------------------------
struct MsgEigTask {
  const void* M;   // χ×χ row-major message, tensor scalar type
  double2* A;      // χ×χ column-major work (Hermitian part of M, then M·V)
  double2* V;      // χ×χ column-major
  void* sqrtM;     // out: Q √D Q†, row-major, tensor scalar type
  void* proj;      // out: Q 1[kept] Q†, row-major, tensor scalar type
  int chi;
  int* flags;      // out: [0] = projector is the identity, [1] = DomainError (negative eigenvalue ≥ cutoff)
  double* errflags;  // engine-wide error words ([1] is raised on a DomainError), all-reduced across ranks before the host reads them
  double* lam;     // work [χ]
};

template <typename R>
__global__ void msg_prepare_kernel(const MsgEigTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  const MsgEigTask t = tasks[blockIdx.x];
  const C* __restrict__ M = (const C*)t.M;
  const int n = t.chi;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx % n, j = idx / n;  // column-major target
    const C a = M[i * n + j], b = M[j * n + i];
    double2 h; h.x = 0.5 * ((double)a.x + (double)b.x); h.y = 0.5 * ((double)a.y - (double)b.y);
    t.A[idx] = h;
    double2 v; v.x = (i == j) ? 1.0 : 0.0; v.y = 0.0;
    t.V[idx] = v;
  }
}

template <typename R>
__global__ void msg_finish_kernel(const MsgEigTask* __restrict__ tasks, double cutoff) {
  using C = typename Cx<R>::type;
  const MsgEigTask t = tasks[blockIdx.x];
  const int n = t.chi;
  __shared__ int s_allkept, s_domain;
  if (threadIdx.x == 0) { s_allkept = 1; s_domain = 0; }
  __syncthreads();
  // λ_j = Re(v_j† (H v_j))
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double l = 0;
    for (int i = 0; i < n; ++i) {
      const double2 v = t.V[i + (long long)j * n], a = t.A[i + (long long)j * n];
      l += v.x * a.x + v.y * a.y;
    }
    const bool kept = !(l == 0.0 || fabs(l) < cutoff);
    if (!kept) s_allkept = 0;
    if (kept && l < 0) { s_domain = 1; l = 0; }
    t.lam[j] = kept ? l : 0.0;
  }
  __syncthreads();
  C* __restrict__ S = (C*)t.sqrtM;
  C* __restrict__ P = (C*)t.proj;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int a = idx / n, b = idx - a * n;  // row-major [a][b]
    double sx = 0, sy = 0, px = 0, py = 0;
    for (int j = 0; j < n; ++j) {
      const double l = t.lam[j];
      if (l <= 0) continue;
      const double2 va = t.V[a + (long long)j * n], vb = t.V[b + (long long)j * n];
      const double rx = va.x * vb.x + va.y * vb.y;  // va * conj(vb)
      const double ry = va.y * vb.x - va.x * vb.y;
      const double f = sqrt(l);
      sx += f * rx; sy += f * ry; px += rx; py += ry;
    }
    C s; s.x = (R)sx; s.y = (R)sy;
    C p; p.x = (R)px; p.y = (R)py;
    S[idx] = s; P[idx] = p;
  }
  __syncthreads();
  if (threadIdx.x == 0) { t.flags[0] = s_allkept; t.flags[1] = s_domain; if (s_domain && t.errflags) t.errflags[1] = 1.0; }
}

// ------------------------------------------------------------------------------------------------
// simple-update reduced factors
// ------------------------------------------------------------------------------------------------
struct HermTask {
  const double2* G;  // n×n row-major
  double2* A;        // n×n column-major Hermitian part
  double2* V;        // identity
  int n;
};
__global__ void herm_prepare_kernel(const HermTask* __restrict__ tasks) {
  const HermTask t = tasks[blockIdx.x];
  const int n = t.n;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx % n, j = idx / n;
    const double2 a = t.G[(long long)i * n + j], b = t.G[(long long)j * n + i];
    double2 h; h.x = 0.5 * (a.x + b.x); h.y = 0.5 * (a.y - b.y);
    t.A[idx] = h;
    double2 v; v.x = (i == j) ? 1.0 : 0.0; v.y = 0.0;
    t.V[idx] = v;
  }
}

// Pivoted Cholesky of the (PSD) reduced-factor Gram matrix, the preconditioner of its Jacobi
// eigendecomposition: with P·G·Pᵀ = L·L†, one-sided Jacobi on the columns of L gives L·W = U·Σ, hence
// G = Pᵀ·U·Σ²·U†·P — eigenvalues σ², eigenvectors Pᵀ·U.  Jacobi on L converges in about half the sweeps it
// needs on G itself (whose columns are graded by the eigenvalues), needs no accumulated V, and returns the
// small eigenvalues to high relative accuracy (Veselić–Hari).  One CTA per matrix, matrix in shared memory.
struct CholTask {
  const double2* G;  // n×n row-major
  double2* A;        // out: L, n×n column-major (columns ≥ rank are zero); after chol_finish: G·V = V·Σ²
  double2* V;        // out of chol_finish: eigenvectors, n×n column-major
  int* piv;          // [n] row i of L is row piv[i] of G
  const double* sval;  // [n] column norms after the Jacobi (chol_finish)
  int n;
  double2* scratch;  // n×n global work matrix when the matrix does not fit shared memory, else nullptr
};

// Block size: 256 threads when the matrix lives in shared memory, 1024 when it lives in the L2-resident scratch (n > 96:
// the trailing update of a step is then a stream of L2 round trips, and four times the threads keep four times as many
// of them in flight).
__global__ void __launch_bounds__(1024) chol_prepare_kernel(const CholTask* __restrict__ tasks, double tol) {
  extern __shared__ __align__(16) unsigned char chol_raw[];
  const CholTask t = tasks[blockIdx.x];
  const int n = t.n, tid = threadIdx.x, nt = blockDim.x;
  // [n][n] trailing matrix / L (row-major): shared memory, or the task's global scratch (L2-resident) for large n
  double2* S = t.scratch ? t.scratch : reinterpret_cast<double2*>(chol_raw);
  double* d = reinterpret_cast<double*>(chol_raw + (t.scratch ? 0 : (size_t)n * n * sizeof(double2)));  // [n] running diagonal
  int* piv = reinterpret_cast<int*>(d + n);                      // [n]
  __shared__ double s_best[32];
  __shared__ int s_besti[32];
  __shared__ int s_p, s_stop;
  __shared__ double s_dmax0;
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx - i * n;
    const double2 a = t.G[idx], b = t.G[(long long)j * n + i];
    double2 h; h.x = 0.5 * (a.x + b.x); h.y = 0.5 * (a.y - b.y);
    S[idx] = h;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) { d[i] = S[i * n + i].x; piv[i] = i; }
  __syncthreads();
  int rank = n;
  for (int k = 0; k < n; ++k) {
    // ---- pivot: largest remaining diagonal entry, ties to the smallest index (deterministic) ----------
    double best = -1.0; int besti = n;
    for (int i = k + tid; i < n; i += nt) { const double v = d[i]; if (v > best) { best = v; besti = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_besti[tid >> 5] = besti; }
    __syncthreads();
    if (tid == 0) {
      double b = -1.0; int bi = n;
      for (int w = 0; w < (nt + 31) / 32; ++w)
        if (s_best[w] > b || (s_best[w] == b && s_besti[w] < bi)) { b = s_best[w]; bi = s_besti[w]; }
      if (k == 0) s_dmax0 = b;
      s_p = bi;
      s_stop = !(b > tol * s_dmax0) || !(b > 0.0);
    }
    __syncthreads();
    if (s_stop) { rank = k; break; }
    const int p = s_p;
    if (p != k) {  // symmetric permutation k <-> p of the trailing matrix and of the finished columns of L
      for (int j = tid; j < n; j += nt) { const double2 a = S[k * n + j]; S[k * n + j] = S[p * n + j]; S[p * n + j] = a; }
      __syncthreads();
      for (int i = tid; i < n; i += nt) {
        if (i >= k) {  // columns k, p only matter inside the trailing block (the finished part of L has columns < k)
          const double2 a = S[i * n + k]; S[i * n + k] = S[i * n + p]; S[i * n + p] = a;
        }
      }
      if (tid == 0) { const double x = d[k]; d[k] = d[p]; d[p] = x; const int q = piv[k]; piv[k] = piv[p]; piv[p] = q; }
      __syncthreads();
    }
    const double lkk = sqrt(d[k]);
    const double inv = 1.0 / lkk;
    __syncthreads();
    for (int i = k + tid; i < n; i += nt) {
      double2 v = S[i * n + k];
      if (i == k) { v.x = lkk; v.y = 0.0; } else { v.x *= inv; v.y *= inv; }
      S[i * n + k] = v;
    }
    __syncthreads();
    // trailing update S[i][j] −= L[i][k]·conj(L[j][k]) for i, j > k (both triangles: keeps the swaps trivial)
    const int m = n - k - 1;
    for (int idx = tid; idx < m * m; idx += nt) {
      const int i = k + 1 + idx / m, j = k + 1 + idx % m;
      const double2 li = S[i * n + k], lj = S[j * n + k];
      double2 v = S[i * n + j];
      v.x -= li.x * lj.x + li.y * lj.y;
      v.y -= li.y * lj.x - li.x * lj.y;
      S[i * n + j] = v;
      if (i == j) d[i] = v.x;
    }
    __syncthreads();
  }
  // ---- write L column-major, zero above the diagonal and beyond the rank ---------------------------------
  for (int idx = tid; idx < n * n; idx += nt) {
    const int j = idx / n, i = idx - j * n;  // column j, row i
    double2 v; v.x = 0; v.y = 0;
    if (j < rank && i >= j) v = S[i * n + j];
    t.A[idx] = v;
  }
  for (int i = tid; i < n; i += nt) t.piv[i] = piv[i];
}

// after the Jacobi on L: V = Pᵀ·U (U = normalised columns), A ← V·Σ² (= G·V, what su_theta expects)
__global__ void __launch_bounds__(256) chol_finish_kernel(const CholTask* __restrict__ tasks) {
  const CholTask t = tasks[blockIdx.x];
  const int n = t.n;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int j = idx / n, i = idx - j * n;
    const double sg = t.sval[j];
    double2 v = t.A[idx];
    const double f = sg > 0 ? 1.0 / sg : 0.0;
    v.x *= f; v.y *= f;
    t.V[(long long)j * n + t.piv[i]] = v;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int j = idx / n;
    const double sg = t.sval[j];
    double2 v = t.V[idx];
    v.x *= sg * sg; v.y *= sg * sg;
    t.A[idx] = v;
  }
}

// K = A†·A for an m×n column-major A (the θ of a gate), row-major n×n out — Gram of the θ columns
struct SmallGemmTask {
  const double2* A;   // m×n column-major
  const double2* B;   // n×n column-major (apply) or unused
  double2* out;
  int m, n;
};
__global__ void __launch_bounds__(256) colgram_kernel(const SmallGemmTask* __restrict__ tasks) {
  const SmallGemmTask t = tasks[blockIdx.x];
  const int m = t.m, n = t.n;
  // 4×4 register tiles over the lower triangle of the n×n result
  const int nt4 = (n + 3) / 4;
  // grid (matrices, slices): the 4×4 output tiles are dealt round-robin to the threads of all slices
  for (int tile = blockIdx.y * blockDim.x + threadIdx.x; tile < nt4 * nt4; tile += blockDim.x * gridDim.y) {
    const int ti = tile / nt4, tj = tile - ti * nt4;
    if (tj > ti) continue;
    double2 acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) { acc[a][b].x = 0; acc[a][b].y = 0; }
    for (int k = 0; k < m; ++k) {
      double2 x[4], y[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = 4 * ti + a, j = 4 * tj + a;
        x[a] = i < n ? t.A[(long long)i * m + k] : make_double2(0, 0);
        y[a] = j < n ? t.A[(long long)j * m + k] : make_double2(0, 0);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {  // conj(x_a)·y_b
          acc[a][b].x += x[a].x * y[b].x + x[a].y * y[b].y;
          acc[a][b].y += x[a].x * y[b].y - x[a].y * y[b].x;
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = 4 * ti + a, j = 4 * tj + b;
        if (i < n && j < n) {
          t.out[(long long)i * n + j] = acc[a][b];
          if (ti != tj) { double2 c = acc[a][b]; c.y = -c.y; t.out[(long long)j * n + i] = c; }
        }
      }
  }
}
// out = A·B : (m×n)·(n×n), all column-major — θ times the approximate right singular vectors
__global__ void __launch_bounds__(256) colapply_kernel(const SmallGemmTask* __restrict__ tasks) {
  const SmallGemmTask t = tasks[blockIdx.x];
  const int m = t.m, n = t.n;
  const int mt4 = (m + 3) / 4, nt4 = (n + 3) / 4;
  for (int tile = blockIdx.y * blockDim.x + threadIdx.x; tile < mt4 * nt4; tile += blockDim.x * gridDim.y) {
    const int tj = tile / mt4, ti = tile - tj * mt4;  // consecutive threads → consecutive row tiles
    double2 acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) { acc[a][b].x = 0; acc[a][b].y = 0; }
    for (int k = 0; k < n; ++k) {
      double2 x[4], y[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = 4 * ti + a, j = 4 * tj + a;
        x[a] = i < m ? t.A[(long long)k * m + i] : make_double2(0, 0);
        y[a] = j < n ? t.B[(long long)j * n + k] : make_double2(0, 0);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          acc[a][b].x += x[a].x * y[b].x - x[a].y * y[b].y;
          acc[a][b].y += x[a].x * y[b].y + x[a].y * y[b].x;
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = 4 * ti + a, j = 4 * tj + b;
        if (i < m && j < n) t.out[(long long)j * m + i] = acc[a][b];
      }
  }
}

struct SuGateTask {
  // per site: eigen-decomposition of the Gram matrix of the gauged tensor (n = d·χ_b)
  const double2* GA[2];  // G·V (column-major) after Jacobi
  const double2* GV[2];  // V
  double* sq[2];         // out [n]: √λ_r (0 for dropped directions)
  double* isq[2];        // out [n]: 1/√λ_r (0 for dropped)
  int d[2];
  int chi_b;
  int full;              // number of singular values the reference's thin QR leaves: min_s(r_s·d_s)
  double2 gate[256];     // (d0·d1)² row-major [(s0',s1')][(s0,s1)]
  double2* theta;        // (n0·d0)×(n1·d1) column-major, Jacobi work
  double2* theta0;       // copy kept for the right factor
  // after the θ Jacobi
  const double* sval;    // [n1·d1]
  const int* perm;
  double* sigma;         // out [n1·d1] singular values, descending
  int* keep;             // out
  double* err;           // out
  double* sumsq_kept;    // out Σ kept σ²
  double2* Rp;           // work (n1·d1) × keep_max
  void* X[2];            // out: [(s,b)][(s',c)] row-major, tensor scalar type
};

// many small device-to-device copies in one launch (a clone's messages): 16-byte aligned buffers, sizes multiple of 8
struct CopyTask { const void* src; void* dst; unsigned long long bytes; };
__global__ void __launch_bounds__(256) copy_many_kernel(const CopyTask* __restrict__ tasks) {
  const CopyTask t = tasks[blockIdx.x];
  const unsigned long long n16 = t.bytes / 16;
  const uint4* __restrict__ s = reinterpret_cast<const uint4*>(t.src);
  uint4* __restrict__ d = reinterpret_cast<uint4*>(t.dst);
  for (unsigned long long i = threadIdx.x; i < n16; i += blockDim.x) d[i] = s[i];
  const unsigned char* sb = reinterpret_cast<const unsigned char*>(t.src);
  unsigned char* db = reinterpret_cast<unsigned char*>(t.dst);
  for (unsigned long long i = n16 * 16 + threadIdx.x; i < t.bytes; i += blockDim.x) db[i] = sb[i];
}

// λ, R = Λ^{1/2}V† and θ = gate·(R_0 ⊗_b R_1)   (simple_update.jl:47-51 with R†R = Gram)
// grid (gates, slices): every slice recomputes the 2·n eigenvalues (n dot products of length n, cheap) into shared
// memory and fills its share of the θ entries, so a batch of few gates (a sharded run leaves 16 per rank) still spreads
// over the SMs; slice 0 also publishes √λ and 1/√λ for su_factors.
constexpr int SU_MAXN = 256;  // d·χ_b of a site (host: launch_jacobi takes matrices of up to 512 rows = d·n)
__global__ void su_theta_kernel(const SuGateTask* __restrict__ tasks, double tolG) {
  const SuGateTask& t = tasks[blockIdx.x];
  __shared__ double red[34];
  __shared__ double s_sq[2][SU_MAXN];
  const int chi = t.chi_b;
  for (int site = 0; site < 2; ++site) {
    const int n = t.d[site] * chi;
    const double2* __restrict__ GA = t.GA[site];
    const double2* __restrict__ GV = t.GV[site];
    double lmax = 0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
      double l = 0;
      for (int i = 0; i < n; ++i) {
        const double2 v = GV[i + (long long)r * n], a = GA[i + (long long)r * n];
        l += v.x * a.x + v.y * a.y;
      }
      s_sq[site][r] = l;  // stash λ
      lmax = fmax(lmax, l);
    }
    __syncthreads();  // block-wide max of λ
    double mx = lmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m2 = 0;
      for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m2 = fmax(m2, red[w]);
      red[33] = m2;
    }
    __syncthreads();
    const double thr = tolG * red[33];
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
      const double l = s_sq[site][r];
      const bool kept = l > thr && l > 0;
      s_sq[site][r] = kept ? sqrt(l) : 0.0;
      if (blockIdx.y == 0) {
        t.sq[site][r] = kept ? sqrt(l) : 0.0;
        t.isq[site][r] = kept ? 1.0 / sqrt(l) : 0.0;
      }
    }
    __syncthreads();
  }
  const int d0 = t.d[0], d1 = t.d[1];
  const int n0 = d0 * chi, n1 = d1 * chi;
  const int rows = n0 * d0, cols = n1 * d1;
  const double2* __restrict__ V0 = t.GV[0];
  const double2* __restrict__ V1 = t.GV[1];
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < rows * cols; idx += blockDim.x * gridDim.y) {
    const int rho = idx % rows, kap = idx / rows;
    const int r0 = rho / d0, s0p = rho - r0 * d0;
    const int r1 = kap / d1, s1p = kap - r1 * d1;
    const double w = s_sq[0][r0] * s_sq[1][r1];
    double2 acc; acc.x = 0; acc.y = 0;
    if (w != 0.0) {
      for (int s0 = 0; s0 < d0; ++s0)
        for (int s1 = 0; s1 < d1; ++s1) {
          const double2 g = t.gate[(s0p * d1 + s1p) * (d0 * d1) + (s0 * d1 + s1)];
          if (g.x == 0.0 && g.y == 0.0) continue;
          // Σ_b conj(V0[(s0,b), r0]) conj(V1[(s1,b), r1])
          double wx = 0, wy = 0;
          const double2* v0 = V0 + (long long)r0 * n0 + s0 * chi;
          const double2* v1 = V1 + (long long)r1 * n1 + s1 * chi;
          for (int b = 0; b < chi; ++b) {
            const double2 x = v0[b], y = v1[b];
            wx += x.x * y.x - x.y * y.y;
            wy -= x.x * y.y + x.y * y.x;
          }
          acc.x += g.x * wx - g.y * wy;
          acc.y += g.x * wy + g.y * wx;
        }
      acc.x *= w; acc.y *= w;
    }
    t.theta[idx] = acc;
    t.theta0[idx] = acc;
  }
}

// NDTensors truncate! on P = σ² (simple_update.jl:53-59): maxdim first, then the relative cutoff
// NDTensors `truncate!` on P = σ² (called through factorize_svd, simple_update.jl:53-59): drop from the tail while
// n > maxdim (whatever mindim says); then either the absolute test (P[n] ≤ cutoff, truncerr left unscaled) or the
// summed test (discarded + P[n] ≤ cutoff·scale, scale = ΣP with use_relative_cutoff, else 1; truncerr /= scale),
// both only while n > mindim.
__global__ void su_truncate_kernel(const SuGateTask* __restrict__ tasks, int ntasks, int maxdim,
                                   int mindim, double cutoff, int use_abs, int use_rel) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ntasks) return;
  const SuGateTask& t = tasks[g];
  const int n = t.d[1] * t.chi_b * t.d[1];
  const int full = t.full;  // ≤ min(rows, cols); the columns beyond are structurally zero
  double total = 0;
  for (int k = 0; k < n; ++k) {
    const double s = t.sval[t.perm[k]];
    t.sigma[k] = s;
    if (k < full) total += s * s;
  }
  int keep = full;
  double disc = 0;
  if (mindim < 1) mindim = 1;
  if (maxdim > 0)
    while (keep > maxdim && keep > 1) { const double s = t.sigma[keep - 1]; disc += s * s; --keep; }
  double err;
  if (use_abs) {
    if (cutoff >= 0)
      while (keep > mindim) {
        const double s = t.sigma[keep - 1];
        if (s * s <= cutoff) { disc += s * s; --keep; } else break;
      }
    err = disc;
  } else {
    const double scale = use_rel ? (total > 0 ? total : 1.0) : 1.0;
    if (cutoff >= 0)
      while (keep > mindim) {
        const double s = t.sigma[keep - 1];
        if (disc + s * s <= cutoff * scale) { disc += s * s; --keep; } else break;
      }
    err = disc / scale;
  }
  if (full <= 1) err = 0.0;  // a single candidate is never truncated
  double kept = 0;
  for (int k = 0; k < keep; ++k) kept += t.sigma[k] * t.sigma[k];
  *t.keep = keep;
  *t.err = err;
  *t.sumsq_kept = kept;
}

// X_0 = V_0 Λ_0^{-1/2} · U√σ,  X_1 = V_1 Λ_1^{-1/2} · conj(V_θ)√σ  (simple_update.jl:53-64 folded:
// T' = (T ×_ext P) · R⁺ · new factor)
// Two launches, grid (gates, slices) each: PART 0 writes the right factor Rp and X_0, PART 1 (which reads all of Rp)
// writes X_1.  Every output entry is an independent dot product, so the slices of a gate split them evenly.
template <typename R, int PART>
__global__ void su_factors_kernel(const SuGateTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  const SuGateTask& t = tasks[blockIdx.x];
  const int chi = t.chi_b, d0 = t.d[0], d1 = t.d[1];
  const int n0 = d0 * chi, n1 = d1 * chi;
  const int rows = n0 * d0, cols = n1 * d1;
  const int keep = *t.keep;
  const int first = blockIdx.y * blockDim.x + threadIdx.x, step = blockDim.x * gridDim.y;
  if (PART == 0) {
    // Numerically null directions (σ_c ≤ 2e-13·‖θ‖_F: the Jacobi stops rotating such columns, so they are
    // not orthogonal to the others in the relative sense) get zero factor columns instead of noise/σ^{3/2};
    // their weight σ_c·u_c v_c† in the two-site tensor is below 2e-13 either way.
    __shared__ double s_thr;
    if (threadIdx.x == 0) {
      double tot = 0;
      for (int k = 0; k < cols; ++k) tot += t.sigma[k] * t.sigma[k];
      s_thr = 2e-13 * sqrt(tot);
    }
    __syncthreads();
    const double sthr = s_thr;
    // right factor: Rp[κ, c] = Σ_ρ θ0[ρ,κ] conj(Uσ[ρ, perm c]) / σ_c^{3/2}
    for (int idx = first; idx < cols * keep; idx += step) {
      const int kap = idx % cols, c = idx / cols;
      const double sg = t.sigma[c];
      double2 acc; acc.x = 0; acc.y = 0;
      if (sg > sthr) {
        const double2* th = t.theta0 + (long long)kap * rows;
        const double2* u = t.theta + (long long)t.perm[c] * rows;
        for (int rho = 0; rho < rows; ++rho) {
          const double2 a = th[rho], b = u[rho];
          acc.x += a.x * b.x + a.y * b.y;
          acc.y += a.y * b.x - a.x * b.y;
        }
        const double f = 1.0 / (sg * sqrt(sg));
        acc.x *= f; acc.y *= f;
      }
      t.Rp[idx] = acc;
    }
    // X_0[(s,b),(s0',c)] = Σ_r V0[(s,b),r] isq0[r] · Uσ[(r,s0'), perm c]/√σ_c
    C* __restrict__ X = (C*)t.X[0];
    const int mm = d0 * keep;
    for (int idx = first; idx < n0 * mm; idx += step) {
      const int row = idx / mm, col = idx - row * mm;
      const int sp = col / keep, c = col - sp * keep;
      const double sg = t.sigma[c];
      double2 acc; acc.x = 0; acc.y = 0;
      if (sg > sthr) {
        const double2* u = t.theta + (long long)t.perm[c] * rows;
        for (int r = 0; r < n0; ++r) {
          const double w = t.isq[0][r];
          if (w == 0.0) continue;
          const double2 v = t.GV[0][row + (long long)r * n0];
          const double2 l = u[r * d0 + sp];
          acc.x += w * (v.x * l.x - v.y * l.y);
          acc.y += w * (v.x * l.y + v.y * l.x);
        }
        const double f = 1.0 / sqrt(sg);
        acc.x *= f; acc.y *= f;
      }
      C o; o.x = (R)acc.x; o.y = (R)acc.y;
      X[idx] = o;
    }
  } else {
    C* __restrict__ X = (C*)t.X[1];
    const int mm = d1 * keep;
    for (int idx = first; idx < n1 * mm; idx += step) {
      const int row = idx / mm, col = idx - row * mm;
      const int sp = col / keep, c = col - sp * keep;
      double2 acc; acc.x = 0; acc.y = 0;
      for (int r = 0; r < n1; ++r) {
        const double w = t.isq[1][r];
        if (w == 0.0) continue;
        const double2 v = t.GV[1][row + (long long)r * n1];
        const double2 l = t.Rp[(r * d1 + sp) + (long long)c * cols];
        acc.x += w * (v.x * l.x - v.y * l.y);
        acc.y += w * (v.x * l.y + v.y * l.x);
      }
      C o; o.x = (R)acc.x; o.y = (R)acc.y;
      X[idx] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BP message finalisation: sum-normalise, convergence measure, write
// (abstractbeliefpropagationcache.jl:182-187, beliefpropagationcache.jl:17-21)
// ------------------------------------------------------------------------------------------------
struct BpFinTask {
  const double2* g;  // χ×χ row-major un-normalised new message m[l][l']
  const void* old_msg;
  void* new_msg;
  double* diff;      // out: 1 − |⟨new,old⟩|²/(‖new‖²‖old‖²)
  int chi;
};

template <typename R>
__global__ void __launch_bounds__(256) bp_finalize_kernel(const BpFinTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  const BpFinTask t = tasks[blockIdx.x];
  __shared__ double red[34];
  const int n2 = t.chi * t.chi;
  double sx = 0, sy = 0;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) { sx += t.g[i].x; sy += t.g[i].y; }
  sx = block_sum(sx, red);
  sy = block_sum(sy, red);
  // m / Σm unless the sum is exactly zero
  double ix = 1.0, iy = 0.0;
  if (sx != 0.0 || sy != 0.0) { const double nn = sx * sx + sy * sy; ix = sx / nn; iy = -sy / nn; }
  const C* __restrict__ old = (const C*)t.old_msg;
  C* __restrict__ nw = (C*)t.new_msg;
  double dx = 0, dy = 0, na = 0, nb = 0;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const double2 g = t.g[i];
    C v; v.x = (R)(g.x * ix - g.y * iy); v.y = (R)(g.x * iy + g.y * ix);
    nw[i] = v;
    const C o = old[i];
    const double ax = v.x, ay = v.y, bx = o.x, by = o.y;
    dx += ax * bx + ay * by;  // conj(a)*b
    dy += ax * by - ay * bx;
    na += ax * ax + ay * ay;
    nb += bx * bx + by * by;
  }
  dx = block_sum(dx, red); dy = block_sum(dy, red);
  na = block_sum(na, red); nb = block_sum(nb, red);
  if (threadIdx.x == 0) *t.diff = 1.0 - (dx * dx + dy * dy) / (na * nb);
}

// ------------------------------------------------------------------------------------------------
// random_tensornetworkstate on the device (tensornetworkstate.jl:93-103: iid normal entries): a counter-based
// generator (SplitMix64 of (seed, vertex, element) → Box–Muller), so the state depends only on the seed
// ------------------------------------------------------------------------------------------------
struct RandTask { void* data; long long n; unsigned long long key; };
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
template <typename R>
__global__ void __launch_bounds__(256) randn_kernel(const RandTask* __restrict__ tasks) {
  using C = typename Cx<R>::type;
  const RandTask t = tasks[blockIdx.y];
  C* __restrict__ out = (C*)t.data;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long h = splitmix64(t.key ^ splitmix64((unsigned long long)i));
    const double u1 = ((double)(h >> 40) + 0.5) * (1.0 / 16777216.0);          // (0,1), 24 bits
    const double u2 = ((double)((h >> 8) & 0xFFFFFFull)) * (1.0 / 16777216.0);  // [0,1)
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    C z;
    z.x = (R)(r * cs); z.y = (R)(r * sn);
    out[i] = z;
  }
}

}  // namespace tnqs
