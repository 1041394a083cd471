// engine.cuh — host-side engine: flat device-resident block layout indexed by graph vertex / edge,
// the BP level scheduler, the batched simple-update pipeline and local expectation values.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/tnqs_b200.h"
#include "kernels_small.cuh"
#include "kernels_tensor.cuh"
#include "kernels_tc.cuh"
#include "kernels_tc2.cuh"
#include "kernels_tc2g.cuh"
#include "kernels_dmma.cuh"

namespace tnqs {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define TNQS_CUDA(x)                                                                           \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess)                                                                     \
      throw ::tnqs::Error(TNQS_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e_));        \
  } while (0)

}  // namespace tnqs
#include "comm.cuh"
namespace tnqs {

struct Leg {
  int edge, nbr;
};

// one chain of mode products on a site tensor: result = T_v ×_{pos_0} M_0 ×_{pos_1} M_1 …
struct Chain {
  int v = -1;
  std::vector<std::pair<int, const void*>> steps;  // (bond-leg position, χ×χ matrix)
  const void* result = nullptr;                    // filled by run_chains
};

class Engine {
 public:
  Engine(int dtype, int nv, int ne, const int32_t* edge_uv, const int32_t* phys, const int32_t* bond,
         int device);
  Engine(const Engine& o);  // deep copy (tnqs_clone)
  ~Engine();

  // import / export
  void set_site(int v, const void* data, int ndim, const int64_t* shape);
  void site_shape(int v, int* ndim, int64_t* shape) const;
  void get_site(int v, void* data, int64_t cap);
  void set_message(int src, int dst, const void* data, int chi);
  void get_message(int src, int dst, void* out, int64_t cap, int* chi, int* is_set);
  void delete_messages();
  void get_bond_dims(int32_t* out) const;
  void set_edge_sequence(const int32_t* seq, int n);

  // hot path
  void apply_gates(int ngates, const int32_t* nverts, const int32_t* verts, const double* mats,
                   const tnqs_apply_opts* ao, const tnqs_bp_opts* bo, int update_cache, double* errs,
                   tnqs_bp_report* reports, int max_reports, int* n_reports);
  tnqs_bp_report bp_update(const tnqs_bp_opts* opts);
  void expect_local(int nobs, const int32_t* verts, const double* ops, double* out);
  void expect_two_site(int nobs, const int32_t* verts, const double* ops, double* out);
  void vertex_scalars(int n, const int32_t* verts, double* out);
  void site_contract(int v, int n_custom, const int32_t* custom_nbrs, const double* custom_mats, int open_nbr, int open_phys,
                     const double* op, double* out, int64_t cap, int* out_n);
  void scale_sites(int n, const int32_t* verts, const double* factors);
  void randomize_sites(unsigned long long seed, int normalize);
  void apply_leg_matrices(int n, const int32_t* verts, const int32_t* nbrs, const double* mats);

  // multi-GPU: vertex ownership + NCCL exchange of the replicated small data (messages, Gram matrices)
  void comm_init(int rank, int nranks, const void* unique_id128, const int32_t* owner);
  bool owns(int v) const { return owner_.empty() || owner_[v] == rank_; }

  void get_stats(tnqs_stats* out, int reset);
  void set_profiling(int on) { profiling_ = on != 0; }

  int dtype() const { return dtype_; }

 private:
  // ---- static description -------------------------------------------------------------------
  int dtype_, esz_, device_, nv_, ne_;
  std::vector<int> eu_, ev_, phys_, bond_;
  std::vector<std::vector<Leg>> inc_;
  std::vector<int> seq_;  // default BP edge sequence (src,dst pairs)
  bool is_tree_ = false;

  // ---- device-resident state ----------------------------------------------------------------
  std::vector<void*> site_;     // [nv]  T_v[s, l_0 …] row-major
  std::vector<void*> msg_;      // [2*ne] directed edge 2e (eu→ev), 2e+1 (ev→eu): χ×χ [ket][bra]
  std::vector<void*> msg_next_; // staging for a BP level
  std::vector<int> msg_dim_;    // χ the buffer currently holds (0: not materialised)
  std::vector<int> msg_next_dim_;
  // A clone places every message and its staging twin in ONE allocation (≈2·10³ stream-ordered allocations and as many
  // copies per functional copy otherwise); buffers inside it are never freed individually (dfree skips them), messages
  // whose dimension changes later move to their own allocation.
  char* msg_arena_ = nullptr;
  size_t msg_arena_bytes_ = 0;
  bool in_msg_arena(const void* p) const {
    return msg_arena_ && (const char*)p >= msg_arena_ && (const char*)p < msg_arena_ + msg_arena_bytes_;
  }
  std::vector<std::vector<int>> sshape_;  // bond-leg dims each site buffer was written with
  std::vector<char> msg_set_;   // 0: identity default (messages(bpc) is empty for it)
  double* d_errflags_ = nullptr;  // [0] Jacobi non-convergence, [1] DomainError; summed over ranks before the host reads them
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  std::vector<void*> temps_;    // stream-ordered temporaries freed by free_temps()
  // Site-tensor buffers released by a gate batch, keyed by size: with saturated bond dimensions the next batch needs exactly
  // these sizes again, so in steady state no site tensor goes through cudaMallocAsync / cudaFreeAsync (the stream-ordered pool
  // fragments under 268 MB blocks and then maps fresh memory: 80-420 ms stalls at chi = 64).  Reuse is ordered by the one stream.
  std::map<size_t, std::vector<void*>> site_pool_;
  size_t site_pool_bytes_ = 0;
  void* site_alloc(size_t bytes);
  void site_release(void* p, size_t bytes);
  void trim_site_pool(size_t keep_bytes);
 public:
  // Out of device memory: every live engine of the device drops its recycled site buffers, the cached scratch slabs go
  // back to the driver, the device is drained; the failed allocation is then retried once (engine.cu: dalloc).
  static void emergency_trim(int device);
 private:
  std::vector<std::pair<char*, size_t>> slabs_;  // GiB-sized scratch slabs for tensor-sized temporaries (process-wide cache)
  size_t slab_cur_ = 0, slab_off_ = 0;
  std::vector<char*> arena_;    // cached chunks for small temporaries
  size_t arena_cur_ = 0, arena_off_ = 0;
  // Task tables go host→device through a ring of (device, pinned-host mirror) chunks so that the
  // copies are truly asynchronous (a pageable cudaMemcpyAsync above 64 KB drains the stream first and
  // would serialise host preparation with device execution).  One set per free_temps() epoch; a set is
  // reused only after the event recorded at the end of its previous epoch has completed.
  struct UpChunk { char* dev; char* host; };
  struct UpSet { std::vector<UpChunk> chunks; size_t cur = 0, off = 0; cudaEvent_t done = nullptr; bool pending = false; };
  static constexpr int kUpSets = 4;
  static constexpr size_t kUpChunk = 4ull << 20;
  UpSet up_[kUpSets];
  int up_cur_ = 0;
  void* up_alloc(size_t bytes, void** host);
  tnqs_stats stats_{};
  bool profiling_ = false;
  int wall_depth_ = 0;
  bool use_tc_ = true;          // tcgen05 path for ComplexF32 (env TNQS_TC=0 disables it)
  bool use_tc2g_ = true;        // TMA-fed warp-specialised Gram contraction (env TNQS_TC2G=0: the LDG-fed tcgen05 kernel)
  bool use_tc2_ = true;         // TMA-fed warp-specialised mode product (env TNQS_TC2=0: the LDG-fed tcgen05 kernel)
  bool use_fast_svd_ = true;        // Cholesky-preconditioned θ-SVD with Jacobi polish (env TNQS_FAST_SVD=0: plain Jacobi on θ)
  bool use_chol_ = true;            // Cholesky-preconditioned eigendecomposition of the reduced-factor Gram (env TNQS_CHOL=0: Jacobi on G)
  bool use_dmma_ = true;            // fp64 tensor-core Hermitian Gram (env TNQS_DMMA=0: SIMT fp64 kernel)
  bool use_cluster_jacobi_ = true;  // shared-memory cluster Jacobi (env TNQS_CLUSTER_JACOBI=0: L2-resident kernel)
  std::shared_ptr<CommHandle> comm_;  // null: single GPU
  std::vector<int> owner_;            // owner rank per vertex (empty: everything local)
  int rank_ = 0, nranks_ = 1;
  struct Bcast { void* ptr; size_t bytes; int root; };
  void exchange(const std::vector<Bcast>& items);       // grouped broadcasts on the engine stream
  void allreduce_sum(double* dptr, size_t count);

  // ---- helpers -------------------------------------------------------------------------------
  void* dalloc(size_t bytes);
  void* talloc(size_t bytes);  // temporary, freed by free_temps()
  void dfree(void* p);
  void release_slabs();
  void free_temps();
  template <class T> T* upload(const std::vector<T>& v);
  int dedge(int src, int dst) const;
  int leg_pos(int v, int e) const;
  long long site_elems(int v) const;
  void leg_view(int v, int pos, unsigned* outer, int* chi, unsigned* inner) const;
  void materialize_message(int de);
  void check_shapes() const;
  size_t scratch_budget(size_t need) const;
  bool c64() const { return dtype_ == TNQS_C64; }

  void launch_mode(std::vector<ModeTask>& tasks);
  std::vector<ModeTask> launch_mode_tc(std::vector<ModeTask>& tasks);
  void launch_gram(std::vector<GramTask>& tasks, bool acc_double, std::vector<double2*>& outs,
                   bool transpose);
  void launch_jacobi(std::vector<JacobiTask>& tasks, double dead_rel2);
  void run_chains(std::vector<Chain>& chains);
  ModeTask mode_task(int v, int pos, const void* in, void* out, const void* mat) const;
  GramTask gram_task(int v, int pos, int planes, const void* X, const void* Y) const;

  void local_rdms(int n, const int32_t* verts, std::vector<std::complex<double>>& rho, std::vector<size_t>& offs);
  void normalize_sites(const std::vector<int>& vs);
  void apply_one_site_batch(const std::vector<std::pair<int, std::vector<std::complex<double>>>>& g,
                            bool normalize);
  void apply_two_site_batch(const std::vector<int>& gate_ids, const int32_t* verts,
                            const double* mats, const std::vector<size_t>& mat_off,
                            const tnqs_apply_opts& ao, double* errs);
  std::vector<std::vector<int>> bp_levels(const std::vector<int>& seq) const;
  void bp_level(const std::vector<int>& seq, const std::vector<int>& items, double* d_diff);
  void count_launch(int n = 1) { stats_.kernel_launches += n; }
  friend struct ProfScope;
};

}  // namespace tnqs
