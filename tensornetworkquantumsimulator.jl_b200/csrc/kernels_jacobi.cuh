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

// LOCALP: every lane group computes the rotation of its own pair right after the dot products (all its lanes hold the
// sums) and applies it at once: one barrier per round instead of three, and nobody waits for the one warp that otherwise
// computes all rotation parameters (ncu, 18 matrices on the whole GPU: 60 % of the warp time at barriers, 22 % of all
// samples behind that warp).  It repeats the fp64 rsqrt chains in every lane, which costs throughput when the SMs are
// full (batching them was worth 1.25× there), so the host picks it only for launches that leave SMs idle — the sharded
// run, small lattices.  Same operations per pair in the same order: the results are bit-identical.
template <int LPP, int RPL, int MAXT, int MINB, bool LOCALP = false>
__global__ void __launch_bounds__(MAXT, MINB) jacobi_cluster_kernel(const JacobiTask* __restrict__ tasks, JacobiAux* __restrict__ aux,
                                                             int BC, int C, int ld, int max_sweeps, double tol,
                                                             double dead_rel2, double* __restrict__ nonconv) {
  extern __shared__ __align__(16) unsigned char jsm_raw[];
  double2* cols = reinterpret_cast<double2*>(jsm_raw);  // [2·BC][ld]
  __shared__ unsigned char s_dead[64];
  __shared__ int s_gcol[64];
  __shared__ int s_rot, s_anyrot;
  __shared__ double s_red[16];
  __shared__ double s_dot[16][4];    // per pair of the round: a, b, Re g, Im g
  __shared__ double s_rotp[16][4];   // c, s, e^{-iφ}
  __shared__ int s_pair[16][2], s_do[16];
  __shared__ unsigned char s_sched[31][16][2];  // pair schedule of the first block-round (all pairs among 2·BC slots)

  const int task = blockIdx.x / C, crank = blockIdx.x - task * C;
  const JacobiTask t = tasks[task];
  JacobiAux* __restrict__ ax = aux + task;
  const int n = t.n, m = t.m;
  const int mt = m + (t.V ? n : 0);
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int nslots = 2 * BC, nb = 2 * C;
  const int pi = tid / LPP, l = tid % LPP;          // pair slot of this lane group, lane inside it
  const unsigned gmask = (LPP == 32) ? 0xffffffffu : (0xffffu << (16 * ((tid & 31) >> 4)));

  // rows ≥ mt of every shared-memory column stay zero (ld = LPP·RPL ≥ mt), so the row loops need no bound
  // checks; the dot products run over the A part only: per-lane bit k set ⇔ row l + LPP·k < m
  unsigned dotmask = 0;
#pragma unroll
  for (int k = 0; k < RPL; ++k) dotmask |= (unsigned)(l + LPP * k < m) << k;
  for (int idx = tid; idx < (nslots - 1) * BC; idx += nthreads) {
    const int r = idx / BC, p = idx - r * BC;
    int a, b;
    rr_pair(nslots, r, p, a, b);
    if (a > b) { const int x = a; a = b; b = x; }
    s_sched[r][p][0] = (unsigned char)a; s_sched[r][p][1] = (unsigned char)b;
  }
  auto col_ptr = [&](int gcol, int row) -> double2* {  // global address of stacked row `row` of column gcol
    return row < m ? t.A + (long long)gcol * m + row : t.V + (long long)gcol * n + (row - m);
  };
  auto load_blocks = [&](int bp, int bq) {
    for (int s = tid; s < nslots; s += nthreads) {
      const int g = (s < BC ? bp : bq) * BC + (s % BC);
      s_gcol[s] = g < n ? g : -1;
      s_dead[s] = g < n ? __ldcg(&ax->dead[g]) : 1;
    }
    // ld is a power of two; 8 independent L2 loads in flight per thread before the first store
    const int total = nslots * ld, ldsh = 31 - __clz(ld);
    for (int base = tid; base < total; base += 8 * nthreads) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * nthreads;
        const int s = idx >> ldsh, row = idx & (ld - 1);
        const int g = (s < BC ? bp : bq) * BC + (s & (BC - 1));
        v[u].x = 0; v[u].y = 0;
        if (idx < total && g < n && row < mt) v[u] = __ldcg(col_ptr(g, row));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * nthreads;
        if (idx < total) cols[idx] = v[u];
      }
    }
    __syncthreads();
  };
  auto store_blocks = [&]() {
    const int total = nslots * ld, ldsh = 31 - __clz(ld);
    for (int idx = tid; idx < total; idx += nthreads) {
      const int s = idx >> ldsh, row = idx & (ld - 1);
      const int g = s_gcol[s];
      if (g >= 0 && row < mt) __stcg(col_ptr(g, row), cols[idx]);
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
    bool conv = false;
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
        double2 xr[RPL], yr[RPL];
        for (int r = 0; r < nrounds; ++r) {
          // ---- phase A: every lane group loads its rows of the two columns and forms a, b, g ------------
          int s1 = 0, s2 = 1;
          if (pi >= BC) {}  // spare lanes of a block smaller than one warp only keep the barriers company
          else if (bt == 0) { s1 = s_sched[r][pi][0]; s2 = s_sched[r][pi][1]; }
          else { s1 = pi; s2 = pi + r; s2 = BC + (s2 >= BC ? s2 - BC : s2); }
          const bool live = pi < BC && !(s_dead[s1] | s_dead[s2]);
          double2* __restrict__ cp = cols + s1 * ld;
          double2* __restrict__ cq = cols + s2 * ld;
          // cross block-rounds: lane group pi keeps ITS column (slot pi) in registers for the whole visit and
          // only the rotating partner goes through shared memory (halves the shared-memory traffic)
          const bool xres = bt != 0;
          if (xres && r == 0 && pi < BC) {
#pragma unroll
            for (int k = 0; k < RPL; ++k) xr[k] = cp[l + LPP * k];
          }
          if (live) {
            double a = 0, b = 0, gx = 0, gy = 0;
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
              if (!xres) xr[k] = cp[l + LPP * k];
              yr[k] = cq[l + LPP * k];
              if ((dotmask >> k) & 1u) {
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
            if (LOCALP) {
              // phases B and C in place (same formulas as below)
              const double g2 = gx * gx + gy * gy;
              const bool alive_p = a > floor2, alive_q = b > floor2;
              if (l == 0) { if (!alive_p) s_dead[s1] = 1; if (!alive_q) s_dead[s2] = 1; }
              if (alive_p && alive_q && g2 > tol2 * a * b) {
                if (l == 0) s_rot = 1;
                const double dl = b - a;
                const double ig = rsqrt(g2);
                const double rh = rsqrt(dl * dl + 4.0 * g2);
                const double c2 = 0.5 + 0.5 * fabs(dl) * rh;
                const double rc = rsqrt(c2);
                const double c = c2 * rc;
                const double s = (dl >= 0 ? 1.0 : -1.0) * (g2 * ig) * rh * rc;
                const double phx = gx * ig, phy = -gy * ig;
#pragma unroll
                for (int k = 0; k < RPL; ++k) {
                  const double2 x = xr[k];
                  double2 y;
                  y.x = yr[k].x * phx - yr[k].y * phy;
                  y.y = yr[k].x * phy + yr[k].y * phx;
                  double2 xn, yn;
                  xn.x = c * x.x - s * y.x; xn.y = c * x.y - s * y.y;
                  yn.x = s * x.x + c * y.x; yn.y = s * x.y + c * y.y;
                  if (xres) xr[k] = xn; else cp[l + LPP * k] = xn;
                  cq[l + LPP * k] = yn;
                }
              }
            } else if (l == 0) { s_dot[pi][0] = a; s_dot[pi][1] = b; s_dot[pi][2] = gx; s_dot[pi][3] = gy; s_pair[pi][0] = s1; s_pair[pi][1] = s2; }
          } else if (!LOCALP && l == 0 && pi < BC) {
            s_pair[pi][0] = -1;
          }
          if (LOCALP) { __syncthreads(); continue; }  // the rotated columns are visible to the next round's pairing
          if (!__syncthreads_or(live ? 1 : 0)) continue;  // nothing alive in this round
          // ---- phase B: one warp turns the BC dot products into rotations (the fp64 rsqrt chains are
          //      issued once per round instead of once per lane group) ---------------------------------------
          if (tid < 32) {
            bool rotate = false;
            if (tid < BC && s_pair[tid][0] >= 0) {
              const double a = s_dot[tid][0], b = s_dot[tid][1], gx = s_dot[tid][2], gy = s_dot[tid][3];
              const double g2 = gx * gx + gy * gy;
              const bool alive_p = a > floor2, alive_q = b > floor2;
              if (!alive_p) s_dead[s_pair[tid][0]] = 1;
              if (!alive_q) s_dead[s_pair[tid][1]] = 1;
              if (alive_p && alive_q && g2 > tol2 * a * b) {
                rotate = true;
                // ζ = δ/2|g|, t = sgn(δ)/(|ζ|+√(1+ζ²))  ⇒  c² = ½ + ½|δ|/h,  s = sgn(δ)·|g|/(h·c),  h = √(δ²+4|g|²)
                const double dl = b - a;
                const double ig = rsqrt(g2);
                const double rh = rsqrt(dl * dl + 4.0 * g2);
                const double c2 = 0.5 + 0.5 * fabs(dl) * rh;
                const double rc = rsqrt(c2);
                s_rotp[tid][0] = c2 * rc;                                        // c
                s_rotp[tid][1] = (dl >= 0 ? 1.0 : -1.0) * (g2 * ig) * rh * rc;   // s
                s_rotp[tid][2] = gx * ig;                                        // e^{-iφ}
                s_rotp[tid][3] = -gy * ig;
              }
            }
            if (tid < BC) s_do[tid] = rotate ? 1 : 0;
            const unsigned any = __ballot_sync(0xffffffffu, rotate);
            if (tid == 0) { s_anyrot = any != 0; if (any) s_rot = 1; }
          }
          __syncthreads();
          if (!s_anyrot) continue;  // converged pairs only: the columns are untouched
          // ---- phase C: apply the rotations to the rows still held in registers ---------------------------
          if (live && s_do[pi]) {
            const double c = s_rotp[pi][0], s = s_rotp[pi][1], phx = s_rotp[pi][2], phy = s_rotp[pi][3];
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
              const double2 x = xr[k];
              double2 y;
              y.x = yr[k].x * phx - yr[k].y * phy;
              y.y = yr[k].x * phy + yr[k].y * phx;
              double2 xn, yn;
              xn.x = c * x.x - s * y.x; xn.y = c * x.y - s * y.y;
              yn.x = s * x.x + c * y.x; yn.y = s * x.y + c * y.y;
              if (xres) xr[k] = xn; else cp[l + LPP * k] = xn;
              cq[l + LPP * k] = yn;
            }
          }
          __syncthreads();
        }
        if (bt != 0 && pi < BC) {  // the resident columns go back to shared memory before the blocks are stored
#pragma unroll
          for (int k = 0; k < RPL; ++k) cols[pi * ld + l + LPP * k] = xr[k];
        }
        if (bt != 0) __syncthreads();
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
      if (!rot) { ++sweep; conv = true; break; }
    }
    if (!conv && crank == 0 && tid == 0 && nonconv) nonconv[0] = 1.0;  // still rotating after max_sweeps: reported to the host
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
