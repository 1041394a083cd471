// capi.cu — the extern "C" boundary declared in include/tnqs_b200.h.  Everything behind it is
// C++/CUDA; exceptions are translated into status codes + a thread-local message here.
#include "engine.cuh"

using tnqs::Engine;
using tnqs::Error;

struct tnqs_state {
  Engine* eng;
};

static thread_local std::string g_last_error;

template <class F>
static int guarded(F&& f) {
  try {
    f();
    return TNQS_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return TNQS_EINVAL;
  } catch (...) {
    g_last_error = "unknown error";
    return TNQS_EINVAL;
  }
}

static Engine& E(tnqs_handle h) {
  if (!h || !h->eng) throw Error(TNQS_EINVAL, "null handle");
  return *h->eng;
}

extern "C" {

int tnqs_create(int dtype, int nv, int ne, const int32_t* edge_uv, const int32_t* phys_dim,
                const int32_t* bond_dim, int device, tnqs_handle* out) {
  return guarded([&] {
    if (!out || !phys_dim || (ne > 0 && (!edge_uv || !bond_dim))) throw Error(TNQS_EINVAL, "null argument");
    *out = nullptr;
    Engine* e = new Engine(dtype, nv, ne, edge_uv, phys_dim, bond_dim, device);
    *out = new tnqs_state{e};
  });
}

int tnqs_clone(tnqs_handle in, tnqs_handle* out) {
  return guarded([&] {
    if (!out) throw Error(TNQS_EINVAL, "null argument");
    *out = nullptr;
    Engine* e = new Engine(E(in));
    *out = new tnqs_state{e};
  });
}

void tnqs_destroy(tnqs_handle h) {
  if (!h) return;
  try { delete h->eng; } catch (...) {}
  delete h;
}

int tnqs_set_site(tnqs_handle h, int v, const void* data, int ndim, const int64_t* shape) {
  return guarded([&] {
    if (!data || !shape) throw Error(TNQS_EINVAL, "null argument");
    E(h).set_site(v, data, ndim, shape);
  });
}
int tnqs_site_shape(tnqs_handle h, int v, int* ndim, int64_t* shape) {
  return guarded([&] {
    if (!ndim || !shape) throw Error(TNQS_EINVAL, "null argument");
    E(h).site_shape(v, ndim, shape);
  });
}
int tnqs_get_site(tnqs_handle h, int v, void* data, int64_t capacity_elems) {
  return guarded([&] {
    if (!data) throw Error(TNQS_EINVAL, "null argument");
    E(h).get_site(v, data, capacity_elems);
  });
}
int tnqs_set_message(tnqs_handle h, int src, int dst, const void* m, int chi) {
  return guarded([&] {
    if (!m) throw Error(TNQS_EINVAL, "null argument");
    E(h).set_message(src, dst, m, chi);
  });
}
int tnqs_get_message(tnqs_handle h, int src, int dst, void* out, int64_t cap, int* chi, int* is_set) {
  return guarded([&] {
    if (!out || !chi || !is_set) throw Error(TNQS_EINVAL, "null argument");
    E(h).get_message(src, dst, out, cap, chi, is_set);
  });
}
int tnqs_delete_messages(tnqs_handle h) {
  return guarded([&] { E(h).delete_messages(); });
}
int tnqs_get_bond_dims(tnqs_handle h, int32_t* out) {
  return guarded([&] {
    if (!out) throw Error(TNQS_EINVAL, "null argument");
    E(h).get_bond_dims(out);
  });
}
int tnqs_set_edge_sequence(tnqs_handle h, const int32_t* seq, int n) {
  return guarded([&] {
    if (n < 0 || (n > 0 && !seq)) throw Error(TNQS_EINVAL, "bad edge sequence");
    E(h).set_edge_sequence(seq, n);
  });
}

int tnqs_apply_gates(tnqs_handle h, int ngates, const int32_t* nverts, const int32_t* verts,
                     const double* gate_mats, const tnqs_apply_opts* aopts, const tnqs_bp_opts* bopts,
                     int update_cache, double* trunc_err, tnqs_bp_report* reports, int max_reports,
                     int* n_reports) {
  return guarded([&] {
    if (ngates < 0 || (ngates > 0 && (!nverts || !verts || !gate_mats || !trunc_err)))
      throw Error(TNQS_EINVAL, "null argument");
    E(h).apply_gates(ngates, nverts, verts, gate_mats, aopts, bopts, update_cache, trunc_err, reports,
                     max_reports, n_reports);
  });
}

int tnqs_bp_update(tnqs_handle h, const tnqs_bp_opts* opts, tnqs_bp_report* report) {
  return guarded([&] {
    const tnqs_bp_report r = E(h).bp_update(opts);
    if (report) *report = r;
  });
}

int tnqs_expect_local(tnqs_handle h, int nobs, const int32_t* verts, const double* op_mats, double* out) {
  return guarded([&] {
    if (nobs > 0 && (!verts || !op_mats || !out)) throw Error(TNQS_EINVAL, "null argument");
    E(h).expect_local(nobs, verts, op_mats, out);
  });
}
int tnqs_vertex_scalars(tnqs_handle h, int n, const int32_t* verts, double* out) {
  return guarded([&] {
    if (n > 0 && (!verts || !out)) throw Error(TNQS_EINVAL, "null argument");
    E(h).vertex_scalars(n, verts, out);
  });
}
int tnqs_scale_sites(tnqs_handle h, int n, const int32_t* verts, const double* factors) {
  return guarded([&] {
    if (n > 0 && (!verts || !factors)) throw Error(TNQS_EINVAL, "null argument");
    E(h).scale_sites(n, verts, factors);
  });
}
int tnqs_site_contract(tnqs_handle h, int v, int n_custom, const int32_t* custom_nbrs, const double* custom_mats, int open_nbr,
                       int open_phys, const double* op, double* out, int64_t capacity, int* n_out) {
  return guarded([&] {
    if (!out || !n_out || (n_custom > 0 && (!custom_nbrs || !custom_mats))) throw Error(TNQS_EINVAL, "null argument");
    E(h).site_contract(v, n_custom, custom_nbrs, custom_mats, open_nbr, open_phys, op, out, capacity, n_out);
  });
}
int tnqs_randomize_sites(tnqs_handle h, uint64_t seed, int normalize) {
  return guarded([&] { E(h).randomize_sites((unsigned long long)seed, normalize); });
}
int tnqs_apply_leg_matrices(tnqs_handle h, int n, const int32_t* verts, const int32_t* nbrs, const double* mats) {
  return guarded([&] {
    if (n > 0 && (!verts || !nbrs || !mats)) throw Error(TNQS_EINVAL, "null argument");
    E(h).apply_leg_matrices(n, verts, nbrs, mats);
  });
}
int tnqs_expect_two_site(tnqs_handle h, int nobs, const int32_t* verts, const double* op_mats, double* out) {
  return guarded([&] {
    if (nobs > 0 && (!verts || !op_mats || !out)) throw Error(TNQS_EINVAL, "null argument");
    E(h).expect_two_site(nobs, verts, op_mats, out);
  });
}

int tnqs_comm_unique_id(void* out128) {
  return guarded([&] {
    if (!out128) throw Error(TNQS_EINVAL, "null argument");
    tnqs::NcclApi& api = tnqs::NcclApi::get();
    tnqs::NcclUniqueId id;
    api.check(api.GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out128, &id, sizeof(id));
  });
}
int tnqs_comm_init(tnqs_handle h, int rank, int nranks, const void* unique_id128, const int32_t* owner) {
  return guarded([&] {
    if (!owner || (nranks > 1 && !unique_id128)) throw Error(TNQS_EINVAL, "null argument");
    E(h).comm_init(rank, nranks, unique_id128, owner);
  });
}

int tnqs_get_stats(tnqs_handle h, tnqs_stats* out, int reset) {
  return guarded([&] {
    if (!out) throw Error(TNQS_EINVAL, "null argument");
    E(h).get_stats(out, reset);
  });
}
int tnqs_set_profiling(tnqs_handle h, int on) {
  return guarded([&] { E(h).set_profiling(on); });
}

const char* tnqs_last_error(void) { return g_last_error.c_str(); }
const char* tnqs_version(void) { return "tnqs_b200 0.1.0 (sm_100a)"; }

}  // extern "C"
