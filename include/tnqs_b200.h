/* tnqs_b200.h — C-ABI of the B200-native belief-propagation simple-update engine.
 *
 * The reference (TensorNetworkQuantumSimulator.jl) has no FFI/plugin seam: its only backend
 * switch is Adapt-based storage swapping (src/Apply/apply_gates.jl:41-44,
 * src/MessagePassing/abstractbeliefpropagationcache.jl:262-287).  The boundary is therefore drawn
 * at the Julia functions examples/2dIsing_dynamics.jl uses; each entry point below names the
 * reference function (file:line under /root/reference) whose work it replaces.  INTEGRATION.md
 * shows the Julia `ccall` methods a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every host buffer belongs to the caller and is read/written
 *     only during the call; every entry point is synchronous on return.
 *   - complex numbers are interleaved (re, im).  Site tensors and messages cross the boundary in
 *     the state's own scalar type (`TNQS_C64` = 2×float, `TNQS_C128` = 2×double); gate and
 *     observable matrices always cross as complex128.
 *   - vertices and edges are 0-based integers; the host language keeps the name ↔ integer map.
 *   - a site tensor is a dense row-major array T[s, l_0, l_1, …] : the physical index first, then
 *     one bond leg per incident edge in increasing edge id.
 *   - a message on the directed edge src→dst is a row-major χ×χ matrix m[ket, bra]
 *     (src/TensorNetworks/tensornetworkstate.jl:72-75).
 *   - return value: 0 on success, otherwise a TNQS_E* code; `tnqs_last_error()` holds the text
 *     (thread-local).  The Julia shim turns codes into `error()` / `ArgumentError`.
 */
#ifndef TNQS_B200_H
#define TNQS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tnqs_state* tnqs_handle;

enum { TNQS_C64 = 0, TNQS_C128 = 1 };

enum {
  TNQS_OK = 0,
  TNQS_EINVAL = 1,      /* bad argument (maps to ArgumentError)                                  */
  TNQS_ENOTADJ = 2,     /* two-site gate on non-adjacent vertices  (apply_gates.jl:114-120)      */
  TNQS_ENSITES = 3,     /* gate on <1 or >2 vertices               (apply_gates.jl:109-112)      */
  TNQS_ECUDA = 4,       /* CUDA runtime failure                                                  */
  TNQS_EDOMAIN = 5,     /* sqrt of a negative message eigenvalue ≥ cutoff (Julia DomainError in
                           utils.jl:21-22)                                                       */
  TNQS_ECAPACITY = 6,   /* caller buffer too small                                               */
  TNQS_ENOGPU = 7       /* no CUDA device: there is no CPU fallback                              */
};

/* keyword arguments of simple_update / factorize_svd  (src/Apply/simple_update.jl:21-24,53-59) */
typedef struct {
  int32_t maxdim;            /* <=0: unlimited                                                   */
  int32_t mindim;            /* default 1                                                        */
  double  cutoff;            /* <0: none; relative cutoff on σ² (NDTensors truncate!)            */
  int32_t normalize_tensors; /* default 1 (simple_update.jl:23)                                  */
  double  sqrt_cutoff;       /* <0: default 10*eps(real(T)) (simple_update.jl:32-33)             */
  /* the remaining factorize_svd keywords that simple_update forwards (simple_update.jl:53-59) */
  int32_t use_absolute_cutoff; /* default 0: 1 = drop while σ² ≤ cutoff, truncerr unscaled (NDTensors truncate!) */
  int32_t use_relative_cutoff; /* default 1: cutoff is relative to Σσ²; 0 = relative to 1              */
  int32_t svd_alg;           /* 0 "divide_and_conquer" (default), 1 "qr_iteration", 2 "recursive": LAPACK driver
                                names of the reference; this library always runs its Jacobi SVD, whose singular
                                values and truncation errors agree with either driver to round-off          */
  int32_t reserved;
} tnqs_apply_opts;

/* keyword arguments of update(bpc; …)  (src/MessagePassing/beliefpropagationcache.jl:56-72,103-119) */
typedef struct {
  int32_t maxiter;           /* <=0: default (25 loopy / 1 tree)                                 */
  double  tolerance;         /* <0: default (1e-5 c64 / 1e-8 c128, none on trees); NaN → none    */
  int32_t use_tolerance;     /* 0: no convergence test (tolerance=nothing)                       */
  const int32_t* edge_sequence; /* 2*n_seq ints (src,dst) or NULL → the cache's own sequence     */
  int32_t n_seq;
} tnqs_bp_opts;

typedef struct {
  int32_t niter;             /* sweeps performed                                                 */
  int32_t converged;         /* 1 if avg diff ≤ tolerance (or no tolerance)                      */
  double  diff;              /* final average message_diff (beliefpropagationcache.jl:17-21)     */
} tnqs_bp_report;

/* --- lifetime ------------------------------------------------------------------------------- */

/* BeliefPropagationCache(ψ) (beliefpropagationcache.jl:27-31) for a network on the given graph.
 * Site tensors start as |0…0> product tensors with every bond = bond_dim[e]; upload real data with
 * tnqs_set_site.  No BP is run; all messages are the identity default.  `device` is the CUDA
 * ordinal.  The default BP schedule (forest_cover_edge_sequence) is supplied by the host through
 * tnqs_set_edge_sequence. */
int  tnqs_create(int dtype, int nv, int ne, const int32_t* edge_uv /*2*ne*/,
                 const int32_t* phys_dim /*nv*/, const int32_t* bond_dim /*ne*/, int device,
                 tnqs_handle* out);
/* Base.copy(bpc) (beliefpropagationcache.jl:35-37): functional-update semantics of
 * apply_gates (apply_gates.jl:55) and update (abstractbeliefpropagationcache.jl:228). */
int  tnqs_clone(tnqs_handle in, tnqs_handle* out);
void tnqs_destroy(tnqs_handle h);

/* --- state import / export: network(ψ_bpc), tn[v], setindex_preserve!, messages ------------ */

/* setindex_preserve! (abstracttensornetwork.jl:40-43); shape[0]=d, shape[1+k] = dim of leg k.
 * Changing a bond dimension resets the two messages on that edge to the identity default. */
int  tnqs_set_site(tnqs_handle h, int v, const void* data, int ndim, const int64_t* shape);
int  tnqs_site_shape(tnqs_handle h, int v, int* ndim /*in: capacity, out: ndim*/, int64_t* shape);
/* network(ψ_bpc)[v] (beliefpropagationcache.jl:24) */
int  tnqs_get_site(tnqs_handle h, int v, void* data, int64_t capacity_elems);
/* setmessage! / message / deletemessage! (abstractbeliefpropagationcache.jl:86-102) */
int  tnqs_set_message(tnqs_handle h, int src, int dst, const void* chi_x_chi, int chi);
int  tnqs_get_message(tnqs_handle h, int src, int dst, void* out, int64_t capacity_elems,
                      int* chi, int* is_set /*0: identity default*/);
int  tnqs_delete_messages(tnqs_handle h);
/* maxvirtualdim / virtualinds (abstracttensornetwork.jl:24-29) */
int  tnqs_get_bond_dims(tnqs_handle h, int32_t* out /*ne*/);
/* edge_sequence field of the cache (beliefpropagationcache.jl:14,39) */
int  tnqs_set_edge_sequence(tnqs_handle h, const int32_t* seq /*2*n*/, int n);

/* --- the hot path ---------------------------------------------------------------------------- */

/* apply_gates(circuit::Vector{<:ITensor}, ψ_bpc; apply_kwargs, bp_update_kwargs, update_cache)
 * (apply_gates.jl:46-98) including apply_gate! (:101-143) and simple_update
 * (simple_update.jl:21-77): walks the gate list with the reference's BP-refresh rule
 * (:60-90), runs every stretch between refreshes as batched device launches, mutates `h` in place
 * (callers wanting the reference's copy semantics clone first).
 *   nverts[i] ∈ {1,2}; verts[2*i], verts[2*i+1]; gate i is a d^n×d^n row-major complex128 matrix,
 *   kron(first, second) basis order, gates packed back to back in `gate_mats`.
 *   trunc_err[i] = spec.truncerr of gate i (0 for one-site gates).
 *   reports: one entry per BP refresh performed, up to max_reports; *n_reports = how many ran.
 *   Size limit of this build: the two-site factorisation (θ of simple_update.jl:51) may have at most 512 rows,
 *   d·min(∏ external dims, d·χ) ≤ 512 with χ ≤ max(current bond, maxdim) — χ ≤ 128 for qubits.  A call that could
 *   exceed it returns TNQS_EINVAL before the state is touched (the reference has no such limit). */
int  tnqs_apply_gates(tnqs_handle h, int ngates, const int32_t* nverts, const int32_t* verts,
                      const double* gate_mats, const tnqs_apply_opts* aopts,
                      const tnqs_bp_opts* bopts, int update_cache, double* trunc_err,
                      tnqs_bp_report* reports, int max_reports, int* n_reports);

/* update(bpc; maxiter, tolerance, edge_sequence) (abstractbeliefpropagationcache.jl:223-259) with
 * updated_message (:162-190), message_diff (beliefpropagationcache.jl:17-21).  Sequential
 * (Gauss–Seidel) semantics over the edge sequence are preserved exactly; independent updates of
 * the sequence are grouped into dependency levels and each level is one batch of launches. */
int  tnqs_bp_update(tnqs_handle h, const tnqs_bp_opts* opts, tnqs_bp_report* report);

/* expect(alg"bp", cache, (op, [v])) (expect.jl:59-82): un-normalised numerator/denominator ratio
 * for single-site operators; ops are d×d complex128 row-major; out is complex128 per observable. */
int  tnqs_expect_local(tnqs_handle h, int nobs, const int32_t* verts, const double* op_mats,
                       double* out /*2*nobs*/);
/* adjacent two-site observable (Steiner tree = the edge, expect.jl:67); ops are d×d each. */
int  tnqs_expect_two_site(tnqs_handle h, int nobs, const int32_t* verts /*2*nobs*/,
                          const double* op_mats /*2 d×d per obs*/, double* out /*2*nobs*/);

/* vertex_scalar(bpc, v) (abstractbeliefpropagationcache.jl:22-28): the contraction of T_v, conj(T_v) and the
 * messages into v — the numerator terms of partitionfunction / freenergy (:289-303); out is complex128 per
 * vertex.  edge_scalar (beliefpropagationcache.jl:47-49) only needs the two χ×χ messages: host side. */
int  tnqs_vertex_scalars(tnqs_handle h, int n, const int32_t* verts, double* out /*2*n*/);
/* tn[v] <- factor_v * tn[v] in place (rescale_vertices!, beliefpropagationcache.jl:82-101);
 * factors are complex128, each vertex at most once. */
int  tnqs_scale_sites(tnqs_handle h, int n, const int32_t* verts, const double* factors /*2*n*/);

/* tn[v] <- tn[v] x_{bond to nbr} M for every listed (v, nbr): M is chi x chi row-major [in][out] complex128,
 * matrices packed back to back; bond dimensions unchanged.  The device half of symmetric_gauge!
 * (symmetric_gauge.jl:1-56: psi_src * inv_rootX * U * sqrtS etc.); the chi x chi algebra stays on the host. */
int  tnqs_apply_leg_matrices(tnqs_handle h, int n, const int32_t* verts, const int32_t* nbrs,
                             const double* mats);

/* One site of a multi-site region contraction — expect(alg"bp") over a Steiner path (src/expect.jl:59-82) and
 * reduced_density_matrix(alg"bp") (src/rdm.jl:52-73).  The ket tensor of vertex v absorbs, on every bond leg except the one
 * towards `open_nbr`, either custom_mats[i] (χ×χ complex128 row-major [ket][bra], for the leg towards custom_nbrs[i]: the
 * partial contraction arriving from the previous site of the path) or that leg's BP message; `op` (d×d complex128 row-major
 * O[s'][s], or NULL) acts on the physical index; the result is contracted with conj(T_v) over all closed indices.
 * out[ket][bra], n×n complex128: n = d·χ (open_phys=1, open_nbr>=0: E[(s,b),(s',b')]), χ (open_phys=0) or d (open_nbr<0).
 * The host walks the path with it (api.py: expect / reduced_density_matrix). */
int  tnqs_site_contract(tnqs_handle h, int v, int n_custom, const int32_t* custom_nbrs, const double* custom_mats,
                        int open_nbr, int open_phys, const double* op, double* out, int64_t capacity, int* n_out);

/* random_tensornetworkstate(eltype, g; bond_dimension) (src/TensorNetworks/tensornetworkstate.jl:93-103) generated in place
 * on the device: every site tensor of the handle (created with the wanted bond dimensions) is filled with iid
 * N(0,1) + i N(0,1) entries from a counter-based generator keyed by (seed, vertex); normalize != 0 scales each tensor to
 * unit Frobenius norm.  All messages return to their identity default.  The RNG stream is this library's own (Julia's
 * is not reproducible outside Julia). */
int  tnqs_randomize_sites(tnqs_handle h, uint64_t seed, int normalize);

/* --- multi-GPU: vertex ownership + NCCL exchange --------------------------- */

/* Join an NCCL communicator: every rank holds the full graph, owns the site tensors of the
 * vertices with owner[v]==rank, and exchanges cut-edge messages / reduced factors inside
 * tnqs_apply_gates / tnqs_bp_update.  unique_id is the 128-byte ncclUniqueId produced by
 * tnqs_comm_unique_id on rank 0 and broadcast by the host (torch.distributed). */
int  tnqs_comm_unique_id(void* out128);
int  tnqs_comm_init(tnqs_handle h, int rank, int nranks, const void* unique_id128,
                    const int32_t* owner /*nv*/);

/* --- diagnostics ----------------------------------------------------------------------------- */

/* per-handle counters since creation / last reset: kernels launched by this library, and device
 * time (ms, CUDA events) spent in BP updates and in gate application. */
typedef struct {
  int64_t kernel_launches;
  double  bp_ms, su_ms;
  int64_t bp_messages;       /* message updates performed                                        */
  int64_t two_site_gates;
  int64_t bp_sweeps;
  double  mode_ms, gram_ms, small_ms; /* per-kernel-family device time when profiling is on      */
  double  mode_flops, gram_flops;     /* algorithmic real flops issued: 8·KK·MM·CC / 8·MM²·CC    */
  int64_t mode_launches, gram_launches;
  int64_t tc_launches;                /* launches of the tcgen05 (tensor-core) kernels             */
  double  mode_bytes, gram_bytes;     /* algorithmic HBM bytes: tensors read + written once        */
  double  wall_ms;                    /* host wall time spent inside tnqs_apply_gates / tnqs_bp_update
                                         (device time bp_ms + su_ms below it means the host is the limit) */
  double  sync_ms;                    /* part of wall_ms the host spent blocked waiting for the device;
                                         wall_ms - sync_ms = host preparation / enqueue time               */
  int64_t tma_launches;               /* of tc_launches: the TMA-fed warp-specialised kernel (kernels_tc2.cuh) */
} tnqs_stats;
int  tnqs_get_stats(tnqs_handle h, tnqs_stats* out, int reset);
int  tnqs_set_profiling(tnqs_handle h, int on);

const char* tnqs_last_error(void);
const char* tnqs_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TNQS_B200_H */
