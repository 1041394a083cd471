// kernels_jacobi.cuh — shared-memory-resident one-sided (Hestenes) Jacobi, one thread-block CLUSTER
// per matrix.
//
// The columns of the stacked matrix [A; V] (A = m×n work matrix, V = n×n accumulated rotations or
// absent) are cut into 2·C blocks of BC columns.  The C CTAs of a cluster play a round-robin
// tournament over the blocks: in block-round t every CTA holds one block pair in shared memory and
// orthogonalises its column pairs there (first block-round of a sweep: all pairs among its 2·BC
// columns; later block-rounds: the BC² cross pairs), then the blocks go back to global memory (L2),
// the cluster synchronises and the next block pairing is loaded.  One sweep therefore visits every
// column pair exactly once — a cyclic Jacobi ordering — with all rotations running on shared memory
// instead of one L2 round trip per pair (the per-SM L2 path was the limit of the previous kernel).
// C = 1 (everything fits one CTA) never leaves shared memory between sweeps.
//
// A column pair is owned by LPP lanes (16: two pairs per warp, 32: one), each lane keeping RPL rows
// of both columns in registers between the dot products and the rotation.
//
// Determinism: the schedule is fixed and every reduction has a fixed order, so all ranks of a
// multi-GPU run that repeat the same factorisation get bit-identical results (engine.cu relies on
// that to keep bond dimensions and messages replicated).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tnqs {

struct JacobiTask {
  double2* A;    // m×n column-major, overwritten by A·V (columns = σ_j u_j)
  double2* V;    // n×n column-major accumulated right rotations, or nullptr
  int m, n;
  double* sval;  // [n] column norms of the result
  int* perm;     // [n] column indices by descending sval
};

struct JacobiAux {       // per task scratch in global memory (zeroed before the launch)
  int rot[64];           // rot[s] != 0: sweep s rotated something
  double part[8];        // per-CTA partial ‖A‖_F²
  unsigned char dead[512];  // numerically null columns
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// tournament pairing (circle method) of `ne` players (even), round `step`, pair index `pair`
__device__ __forceinline__ void rr_pair(int ne, int step, int pair, int& p, int& q) {
  if (pair == 0) { p = step; q = ne - 1; }
  else { p = (step + pair) % (ne - 1); q = (step - pair + (ne - 1)) % (ne - 1); }
}

template <int LPP, int RPL>
__global__ void __launch_bounds__(512) jacobi_cluster_kernel(const JacobiTask* __restrict__ tasks, JacobiAux* __restrict__ aux,
                                                             int BC, int C, int ld, int max_sweeps, double tol,
                                                             double dead_rel2) {
  extern __shared__ __align__(16) unsigned char jsm_raw[];
  double2* cols = reinterpret_cast<double2*>(jsm_raw);  // [2·BC][ld]
  __shared__ unsigned char s_dead[64];
  __shared__ int s_gcol[64];
  __shared__ int s_rot;
  __shared__ double s_red[16];

  const int task = blockIdx.x / C, crank = blockIdx.x - task * C;
  const JacobiTask t = tasks[task];
  JacobiAux* __restrict__ ax = aux + task;
  const int n = t.n, m = t.m;
  const int mt = m + (t.V ? n : 0);
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int nslots = 2 * BC, nb = 2 * C;
  const int pi = tid / LPP, l = tid % LPP;          // pair slot of this lane group, lane inside it
  const unsigned gmask = (LPP == 32) ? 0xffffffffu : (0xffffu << (16 * ((tid & 31) >> 4)));

  auto col_ptr = [&](int gcol, int row) -> double2* {  // global address of stacked row `row` of column gcol
    return row < m ? t.A + (long long)gcol * m + row : t.V + (long long)gcol * n + (row - m);
  };
  auto load_blocks = [&](int bp, int bq) {
    for (int s = tid; s < nslots; s += nthreads) {
      const int g = (s < BC ? bp : bq) * BC + (s % BC);
      s_gcol[s] = g < n ? g : -1;
      s_dead[s] = g < n ? __ldcg(&ax->dead[g]) : 1;
    }
    for (int idx = tid; idx < nslots * mt; idx += nthreads) {
      const int s = idx / mt, row = idx - s * mt;
      const int g = (s < BC ? bp : bq) * BC + (s % BC);
      double2 v; v.x = 0; v.y = 0;
      if (g < n) v = __ldcg(col_ptr(g, row));
      cols[s * ld + row] = v;
    }
    __syncthreads();
  };
  auto store_blocks = [&]() {
    for (int idx = tid; idx < nslots * mt; idx += nthreads) {
      const int s = idx / mt, row = idx - s * mt;
      const int g = s_gcol[s];
      if (g >= 0) __stcg(col_ptr(g, row), cols[s * ld + row]);
    }
    for (int s = tid; s < nslots; s += nthreads)
      if (s_gcol[s] >= 0) ax->dead[s_gcol[s]] = s_dead[s];
  };

  // ---- ‖A‖_F² (fixed summation order) → null-column floor ---------------------------------------------
  int bp, bq;
  rr_pair(nb, 0, crank, bp, bq);
  if (bp > bq) { const int x = bp; bp = bq; bq = x; }
  load_blocks(bp, bq);
  {
    double part = 0;
    for (int idx = tid; idx < nslots * m; idx += nthreads) {
      const int s = idx / m, row = idx - s * m;
      const double2 x = cols[s * ld + row];
      part += x.x * x.x + x.y * x.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      double f = 0;
      for (int w = 0; w < (nthreads + 31) / 32; ++w) f += s_red[w];
      ax->part[crank] = f;
      __threadfence();
    }
    if (C > 1) cluster_sync_all(); else __syncthreads();
  }
  double fro = 0;
  for (int c = 0; c < C; ++c) fro += __ldcg(&ax->part[c]);
  const double floor2 = dead_rel2 * fro;
  const double tol2 = tol * tol * (double)(m > 1 ? m : 1);

  // ---- sweeps ------------------------------------------------------------------------------------------
  int sweep = 0;
  if (n >= 2) {
    for (; sweep < max_sweeps; ++sweep) {
      if (tid == 0) s_rot = 0;
      for (int bt = 0; bt < nb - 1; ++bt) {
        if (C > 1 && !(sweep == 0 && bt == 0)) {
          rr_pair(nb, bt, crank, bp, bq);
          if (bp > bq) { const int x = bp; bp = bq; bq = x; }
          load_blocks(bp, bq);
        } else {
          __syncthreads();
        }
        const int nrounds = (bt == 0) ? nslots - 1 : BC;
        for (int r = 0; r < nrounds; ++r) {
          int s1 = 0, s2 = 1;
          if (pi >= BC) {}  // spare lanes of a block smaller than one warp only keep the barriers company
          else if (bt == 0) { rr_pair(nslots, r, pi, s1, s2); if (s1 > s2) { const int x = s1; s1 = s2; s2 = x; } }
          else { s1 = pi; s2 = BC + (pi + r) % BC; }
          if (pi < BC && !(s_dead[s1] | s_dead[s2])) {
            double2* __restrict__ cp = cols + s1 * ld;
            double2* __restrict__ cq = cols + s2 * ld;
            double2 xr[RPL], yr[RPL];
            double a = 0, b = 0, gx = 0, gy = 0;
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
              const int i = l + LPP * k;
              if (i < mt) { xr[k] = cp[i]; yr[k] = cq[i]; }
              else { xr[k].x = xr[k].y = 0; yr[k].x = yr[k].y = 0; }
              if (i < m) {
                const double2 x = xr[k], y = yr[k];
                a += x.x * x.x + x.y * x.y;
                b += y.x * y.x + y.y * y.y;
                gx += x.x * y.x + x.y * y.y;   // conj(x)·y
                gy += x.x * y.y - x.y * y.x;
              }
            }
#pragma unroll
            for (int o = LPP / 2; o > 0; o >>= 1) {
              a += __shfl_xor_sync(gmask, a, o);
              b += __shfl_xor_sync(gmask, b, o);
              gx += __shfl_xor_sync(gmask, gx, o);
              gy += __shfl_xor_sync(gmask, gy, o);
            }
            const double g2 = gx * gx + gy * gy;
            const bool alive_p = a > floor2, alive_q = b > floor2;
            if (l == 0) {
              if (!alive_p) s_dead[s1] = 1;
              if (!alive_q) s_dead[s2] = 1;
            }
            if (alive_p && alive_q && g2 > tol2 * a * b) {
              if (l == 0) s_rot = 1;
              const double ig = rsqrt(g2);
              const double zeta = 0.5 * (b - a) * ig;
              const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
              const double c = rsqrt(1.0 + tt * tt), s = c * tt;
              const double phx = gx * ig, phy = -gy * ig;  // e^{-iφ}
#pragma unroll
              for (int k = 0; k < RPL; ++k) {
                const int i = l + LPP * k;
                if (i < mt) {
                  const double2 x = xr[k];
                  double2 y;
                  y.x = yr[k].x * phx - yr[k].y * phy;
                  y.y = yr[k].x * phy + yr[k].y * phx;
                  double2 xn, yn;
                  xn.x = c * x.x - s * y.x; xn.y = c * x.y - s * y.y;
                  yn.x = s * x.x + c * y.x; yn.y = s * x.y + c * y.y;
                  cp[i] = xn; cq[i] = yn;
                }
              }
            }
          }
          __syncthreads();
        }
        if (C > 1) {
          store_blocks();
          if (bt == nb - 2 && tid == 0 && s_rot) atomicOr(&ax->rot[sweep], 1);
          __threadfence();
          cluster_sync_all();
        }
      }
      int rot;
      if (C > 1) rot = __ldcg(&ax->rot[sweep]);
      else { rot = s_rot; if (tid == 0 && rot) ax->rot[sweep] = 1; __syncthreads(); }
      if (!rot) { ++sweep; break; }
    }
  }
  // ---- singular values (column norms of the A part) and their descending order --------------------------
  // after the last block-round (or, for C == 1, always) this CTA's columns are in shared memory
  for (int s = tid / 32; s < nslots; s += nthreads / 32) {
    const int g = s_gcol[s];
    if (g < 0) continue;
    double a = 0;
    for (int i = tid & 31; i < m; i += 32) { const double2 x = cols[s * ld + i]; a += x.x * x.x + x.y * x.y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0) t.sval[g] = sqrt(a);
  }
  if (C == 1) { __syncthreads(); store_blocks(); }
  __threadfence();
  if (C > 1) cluster_sync_all(); else __syncthreads();
  if (crank == 0) {
    for (int j = tid; j < n; j += nthreads) {
      const double sj = __ldcg(&t.sval[j]);
      int rank = 0;
      for (int i = 0; i < n; ++i) {
        const double si = __ldcg(&t.sval[i]);
        rank += (si > sj) || (si == sj && i < j);
      }
      t.perm[rank] = j;
    }
  }
}

}  // namespace tnqs
