// engine.cu — implementation of the host-side engine (see engine.cuh).
#include "engine.cuh"

#include <chrono>
#include <functional>
#include <cstdio>
#include <numeric>
#include <set>
#include <tuple>
#include <cstdlib>

namespace tnqs {

using cplx = std::complex<double>;

// TNQS_SLOWLOG=1: report host-side CUDA calls that take longer than 3 ms (allocator / driver stalls)
struct SlowLog {
  const char* what; std::chrono::steady_clock::time_point t0;
  static bool on() { static const bool v = std::getenv("TNQS_SLOWLOG") != nullptr; return v; }
  explicit SlowLog(const char* w) : what(w), t0(std::chrono::steady_clock::now()) {}
  ~SlowLog() {
    if (!on()) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms > 3.0) std::fprintf(stderr, "[tnqs slow] %s took %.1f ms\n", what, ms);
  }
};

// host time spent blocked on the device (stats: wall_ms − sync_ms = host time spent preparing / enqueueing)
struct WaitScope {
  double* acc; std::chrono::steady_clock::time_point t0;
  explicit WaitScope(double* a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~WaitScope() { *acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// Process-wide free list of upload chunks per device: apply_gates / update clone the engine for the
// reference's functional-copy semantics, and pinning host memory (cudaHostAlloc) costs milliseconds.
namespace {
struct UpPool {
  std::mutex mu;
  std::map<int, std::vector<std::pair<char*, char*>>> free_;
  static UpPool& get() { static UpPool p; return p; }
  std::pair<char*, char*> take(int device, size_t bytes) {
    {
      std::lock_guard<std::mutex> lk(mu);
      auto& v = free_[device];
      if (!v.empty()) { auto c = v.back(); v.pop_back(); return c; }
    }
    char *d = nullptr, *h = nullptr;
    SlowLog sl("cudaMalloc+cudaHostAlloc(upload chunk)");
    if (cudaMalloc(&d, bytes) != cudaSuccess) throw Error(TNQS_ECUDA, "cudaMalloc(upload chunk) failed");
    if (cudaHostAlloc(&h, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaFree(d); throw Error(TNQS_ECUDA, "cudaHostAlloc(upload chunk) failed"); }
    return {d, h};
  }
  void give(int device, char* d, char* h) {
    std::lock_guard<std::mutex> lk(mu);
    free_[device].push_back({d, h});
  }
};
struct SlabPool {
  std::mutex mu;
  std::map<int, std::vector<std::pair<char*, size_t>>> free_;
  static SlabPool& get() { static SlabPool p; return p; }
  std::pair<char*, size_t> take(int device, size_t bytes) {
    {
      std::lock_guard<std::mutex> lk(mu);
      auto& v = free_[device];
      int best = -1;
      for (int i = 0; i < (int)v.size(); ++i)
        if (v[i].second >= bytes && (best < 0 || v[i].second < v[best].second)) best = i;
      if (best >= 0) { auto c = v[best]; v.erase(v.begin() + best); return c; }
    }
    char* d = nullptr;
    SlowLog sl("cudaMalloc(slab)");
    if (cudaMalloc(&d, bytes) != cudaSuccess) {
      cudaGetLastError();
      trim(device);  // give cached slabs and the idle part of the stream-ordered pool back to the driver, retry once
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { cudaDeviceSynchronize(); cudaMemPoolTrimTo(pool, 0); }
      if (cudaMalloc(&d, bytes) != cudaSuccess) throw Error(TNQS_ECUDA, "cudaMalloc(scratch slab) failed: out of device memory");
    }
    return {d, bytes};
  }
  void give(int device, char* d, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu);
    free_[device].push_back({d, bytes});
  }
  size_t free_bytes(int device) {
    std::lock_guard<std::mutex> lk(mu);
    size_t t = 0;
    for (auto& c : free_[device]) t += c.second;
    return t;
  }
  void trim(int device) {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& c : free_[device]) cudaFree(c.first);
    free_[device].clear();
  }
};
}  // namespace

namespace {
struct EngineRegistry {
  std::mutex mu;
  std::set<Engine*> live;
  static EngineRegistry& get() { static EngineRegistry r; return r; }
};
}  // namespace
void Engine::emergency_trim(int device) {
  {
    EngineRegistry& reg = EngineRegistry::get();
    std::lock_guard<std::mutex> lk(reg.mu);
    for (Engine* e : reg.live) {
      if (e->device_ != device) continue;
      for (auto& kv : e->site_pool_)
        for (void* p : kv.second) cudaFreeAsync(p, e->stream_);
      e->site_pool_.clear();
      e->site_pool_bytes_ = 0;
    }
  }
  cudaDeviceSynchronize();
  SlabPool::get().trim(device);
  cudaGetLastError();
}

// Opt-in to > 48 KB dynamic shared memory for every kernel that needs it.  The attribute is per device, so it is
// tracked per device ordinal (a second cache on another GPU of the same process must opt in again).
static void ensure_kernel_attributes(int device) {
  static std::mutex mu;
  static bool done[64] = {false};
  std::lock_guard<std::mutex> lk(mu);
  if (device >= 0 && device < 64 && done[device]) return;
  auto set = [](const void* f, int bytes) { TNQS_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); };
  set((const void*)tc::tc_mode_kernel<false>, 200 * 1024);
  set((const void*)tc::tc_mode_kernel<true>, 200 * 1024);
  set((const void*)tc2::tc2_mode_kernel<false>, 225 * 1024);
  set((const void*)tc2::tc2_mode_kernel<true>, 225 * 1024);
  set((const void*)tc2g::tc2_gram_kernel<false>, 225 * 1024);
  set((const void*)tc2g::tc2_gram_kernel<true>, 225 * 1024);
  set((const void*)tc::tc_gram_kernel<false, 1, 2>, 200 * 1024);
  set((const void*)tc::tc_gram_kernel<false, 2, 2>, 200 * 1024);
  set((const void*)tc::tc_gram_kernel<true, 1, 2>, 200 * 1024);
  set((const void*)tc::tc_gram_kernel<true, 2, 2>, 200 * 1024);
  set((const void*)tc::tc_gram_kernel<true, 4, 1>, 200 * 1024);
  set((const void*)gram_dmma_kernel<float, false, 8>, 160 * 1024);
  set((const void*)gram_dmma_kernel<float, true, 8>, 160 * 1024);
  set((const void*)gram_dmma_kernel<double, false, 8>, 160 * 1024);
  set((const void*)gram_dmma_kernel<double, true, 8>, 160 * 1024);
  set((const void*)gram_dmma_kernel<float, false, 12>, 160 * 1024);
  set((const void*)gram_dmma_kernel<float, true, 12>, 160 * 1024);
  set((const void*)gram_dmma_kernel<double, false, 12>, 160 * 1024);
  set((const void*)gram_dmma_kernel<double, true, 12>, 160 * 1024);
  set((const void*)chol_prepare_kernel, 160 * 1024);
  if (device >= 0 && device < 64) done[device] = true;
}

// ------------------------------------------------------------------------------------------------
// construction / destruction
// ------------------------------------------------------------------------------------------------
Engine::Engine(int dtype, int nv, int ne, const int32_t* edge_uv, const int32_t* phys,
               const int32_t* bond, int device)
    : dtype_(dtype), esz_(dtype == TNQS_C64 ? 8 : 16), device_(device), nv_(nv), ne_(ne) {
  if (dtype != TNQS_C64 && dtype != TNQS_C128) throw Error(TNQS_EINVAL, "dtype must be TNQS_C64 or TNQS_C128");
  if (nv <= 0 || ne < 0) throw Error(TNQS_EINVAL, "need nv > 0, ne >= 0");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw Error(TNQS_ENOGPU, "no CUDA device visible: tnqs_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) throw Error(TNQS_EINVAL, "bad CUDA device ordinal");
  TNQS_CUDA(cudaSetDevice(device));
  ensure_kernel_attributes(device);
  eu_.resize(ne); ev_.resize(ne); bond_.resize(ne); phys_.assign(phys, phys + nv);
  inc_.assign(nv, {});
  std::set<std::pair<int, int>> seen;
  for (int e = 0; e < ne; ++e) {
    const int u = edge_uv[2 * e], v = edge_uv[2 * e + 1];
    if (u < 0 || v < 0 || u >= nv || v >= nv || u == v) throw Error(TNQS_EINVAL, "bad edge endpoints");
    if (!seen.insert({std::min(u, v), std::max(u, v)}).second) throw Error(TNQS_EINVAL, "duplicate edge");
    if (bond[e] < 1) throw Error(TNQS_EINVAL, "bond dimension must be >= 1");
    eu_[e] = u; ev_[e] = v; bond_[e] = bond[e];
    inc_[u].push_back({e, v});
    inc_[v].push_back({e, u});
  }
  for (int v = 0; v < nv; ++v)
    if (phys_[v] < 1 || phys_[v] > 4) throw Error(TNQS_EINVAL, "physical dimension must be in 1..4");
  // tree test (default_bp_maxiter, beliefpropagationcache.jl:39)
  {
    std::vector<char> vis(nv, 0);
    std::vector<int> st{0};
    vis[0] = 1;
    int cnt = 1;
    while (!st.empty()) {
      const int x = st.back(); st.pop_back();
      for (auto& l : inc_[x]) if (!vis[l.nbr]) { vis[l.nbr] = 1; ++cnt; st.push_back(l.nbr); }
    }
    is_tree_ = (cnt == nv) && (ne == nv - 1);
  }
  TNQS_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  TNQS_CUDA(cudaEventCreate(&ev0_));
  TNQS_CUDA(cudaEventCreate(&ev1_));
  { const char* e = std::getenv("TNQS_TC"); use_tc_ = !(e && e[0] == '0'); }
  { const char* e = std::getenv("TNQS_TC2"); use_tc2_ = !(e && e[0] == '0'); }
  { const char* e = std::getenv("TNQS_TC2G"); use_tc2g_ = !(e && e[0] == '0'); }
  { const char* e = std::getenv("TNQS_FAST_SVD"); use_fast_svd_ = !(e && e[0] == '0'); }
  { const char* e = std::getenv("TNQS_CHOL"); use_chol_ = !(e && e[0] == '0'); }
  { const char* e = std::getenv("TNQS_DMMA"); use_dmma_ = !(e && e[0] == '0'); }
  { const char* e = std::getenv("TNQS_CLUSTER_JACOBI"); use_cluster_jacobi_ = !(e && e[0] == '0'); }
  cudaMemPool_t pool;
  TNQS_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX;
  TNQS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  site_.assign(nv, nullptr);
  sshape_.assign(nv, {});
  for (int v = 0; v < nv; ++v) {
    sshape_[v].clear();
    for (auto& l : inc_[v]) sshape_[v].push_back(bond_[l.edge]);
    const long long n = site_elems(v);
    site_[v] = dalloc((size_t)n * esz_);
    TNQS_CUDA(cudaMemsetAsync(site_[v], 0, (size_t)n * esz_, stream_));
    if (c64()) { const float one[2] = {1.f, 0.f}; TNQS_CUDA(cudaMemcpyAsync(site_[v], one, 8, cudaMemcpyHostToDevice, stream_)); }
    else { const double one[2] = {1.0, 0.0}; TNQS_CUDA(cudaMemcpyAsync(site_[v], one, 16, cudaMemcpyHostToDevice, stream_)); }
  }
  msg_.assign(2 * ne, nullptr);
  msg_next_.assign(2 * ne, nullptr);
  msg_dim_.assign(2 * ne, 0);
  msg_next_dim_.assign(2 * ne, 0);
  msg_set_.assign(2 * ne, 0);
  d_errflags_ = (double*)dalloc(2 * sizeof(double));
  TNQS_CUDA(cudaMemsetAsync(d_errflags_, 0, 2 * sizeof(double), stream_));
  TNQS_CUDA(cudaStreamSynchronize(stream_));
  { EngineRegistry& reg = EngineRegistry::get(); std::lock_guard<std::mutex> lk(reg.mu); reg.live.insert(this); }
}

Engine::Engine(const Engine& o)
    : dtype_(o.dtype_), esz_(o.esz_), device_(o.device_), nv_(o.nv_), ne_(o.ne_), eu_(o.eu_), ev_(o.ev_),
      phys_(o.phys_), bond_(o.bond_), inc_(o.inc_), seq_(o.seq_), is_tree_(o.is_tree_), sshape_(o.sshape_) {
  use_tc_ = o.use_tc_;
  use_tc2_ = o.use_tc2_;
  use_tc2g_ = o.use_tc2g_;
  use_cluster_jacobi_ = o.use_cluster_jacobi_;
  use_dmma_ = o.use_dmma_;
  use_chol_ = o.use_chol_;
  use_fast_svd_ = o.use_fast_svd_;
  profiling_ = o.profiling_;
  comm_ = o.comm_; owner_ = o.owner_; rank_ = o.rank_; nranks_ = o.nranks_;
  TNQS_CUDA(cudaSetDevice(device_));
  ensure_kernel_attributes(device_);
  TNQS_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  TNQS_CUDA(cudaEventCreate(&ev0_));
  TNQS_CUDA(cudaEventCreate(&ev1_));
  TNQS_CUDA(cudaStreamSynchronize(o.stream_));
  site_.assign(nv_, nullptr);
  for (int v = 0; v < nv_; ++v) {
    if (!o.site_[v]) continue;  // not owned by this rank
    const size_t b = (size_t)site_elems(v) * esz_;
    site_[v] = dalloc(b);
    TNQS_CUDA(cudaMemcpyAsync(site_[v], o.site_[v], b, cudaMemcpyDeviceToDevice, stream_));
  }
  msg_.assign(2 * ne_, nullptr);
  msg_next_.assign(2 * ne_, nullptr);
  msg_dim_ = o.msg_dim_;
  msg_next_dim_.assign(2 * ne_, 0);
  msg_set_ = o.msg_set_;
  {
    // every message and its BP staging twin in one allocation, copied by one kernel
    size_t tot = 0;
    std::vector<size_t> off(2 * ne_, 0);
    for (int de = 0; de < 2 * ne_; ++de) {
      if (!o.msg_[de] || !o.msg_dim_[de]) { msg_dim_[de] = 0; continue; }
      const size_t b = ((size_t)msg_dim_[de] * msg_dim_[de] * esz_ + 255) & ~size_t(255);
      off[de] = tot;
      tot += 2 * b;
    }
    if (tot > 0) {
      msg_arena_ = (char*)dalloc(tot);
      msg_arena_bytes_ = tot;
      std::vector<CopyTask> ct;
      for (int de = 0; de < 2 * ne_; ++de) {
        if (!msg_dim_[de]) continue;
        const size_t raw = (size_t)msg_dim_[de] * msg_dim_[de] * esz_;
        const size_t b = (raw + 255) & ~size_t(255);
        msg_[de] = msg_arena_ + off[de];
        msg_next_[de] = msg_arena_ + off[de] + b;
        msg_next_dim_[de] = msg_dim_[de];
        ct.push_back({o.msg_[de], msg_[de], (unsigned long long)raw});
      }
      CopyTask* dct = upload(ct);
      copy_many_kernel<<<(unsigned)ct.size(), 256, 0, stream_>>>(dct);
      TNQS_CUDA(cudaGetLastError());
    }
  }
  d_errflags_ = (double*)dalloc(2 * sizeof(double));
  TNQS_CUDA(cudaMemsetAsync(d_errflags_, 0, 2 * sizeof(double), stream_));
  TNQS_CUDA(cudaStreamSynchronize(stream_));
  { EngineRegistry& reg = EngineRegistry::get(); std::lock_guard<std::mutex> lk(reg.mu); reg.live.insert(this); }
}

Engine::~Engine() {
  { EngineRegistry& reg = EngineRegistry::get(); std::lock_guard<std::mutex> lk(reg.mu); reg.live.erase(this); }
  cudaSetDevice(device_);
  if (stream_) cudaStreamSynchronize(stream_);
  for (void* p : temps_) cudaFreeAsync(p, stream_);
  for (char* p : arena_) cudaFreeAsync(p, stream_);
  for (auto& sl : slabs_) SlabPool::get().give(device_, sl.first, sl.second);
  slabs_.clear();
  for (auto& s : up_) {  // the stream is idle: the chunks can serve another engine
    for (auto& c : s.chunks) UpPool::get().give(device_, c.dev, c.host);
    s.chunks.clear();
    if (s.done) cudaEventDestroy(s.done);
  }
  for (void* p : site_) if (p) cudaFreeAsync(p, stream_);
  for (auto& kv : site_pool_) for (void* p : kv.second) cudaFreeAsync(p, stream_);
  site_pool_.clear();
  for (void* p : msg_) if (p && !in_msg_arena(p)) cudaFreeAsync(p, stream_);
  for (void* p : msg_next_) if (p && !in_msg_arena(p)) cudaFreeAsync(p, stream_);
  if (msg_arena_) cudaFreeAsync(msg_arena_, stream_);
  if (d_errflags_) cudaFreeAsync(d_errflags_, stream_);
  if (stream_) { cudaStreamSynchronize(stream_); cudaStreamDestroy(stream_); }
  if (ev0_) cudaEventDestroy(ev0_);
  if (ev1_) cudaEventDestroy(ev1_);
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
void* Engine::dalloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  SlowLog sl("cudaMallocAsync");
  cudaError_t err = cudaMallocAsync(&p, bytes, stream_);
  if (err == cudaErrorMemoryAllocation) {
    // e.g. a functional copy of a 58 GB state next to the buffers the source engine keeps for recycling: release every
    // cache of the process on this device and try once more
    cudaGetLastError();
    emergency_trim(device_);
    err = cudaMallocAsync(&p, bytes, stream_);
  }
  TNQS_CUDA(err);
  return p;
}
// Temporaries: small ones are bump-allocated from cached 32 MiB chunks (thousands per gate batch —
// one cudaMallocAsync each would dominate the host time), large ones go to the stream-ordered pool.
void* Engine::talloc(size_t bytes) {
  constexpr size_t kChunk = 32ull << 20, kSmall = 1ull << 20, kSlab = 1ull << 30;
  if (bytes >= kSmall) {
    // tensor-sized temporaries are carved from 1 GiB slabs: a BP level needs ~10³ of them and one
    // cudaMallocAsync + cudaFreeAsync per buffer made the host the bottleneck of the sweep
    // The slabs come from a process-wide cache (cudaMalloc once, reused by every engine and clone): the
    // stream-ordered pool cannot coalesce its free blocks into GiB-sized ones and would keep mapping new memory.
    bytes = (bytes + 255) & ~size_t(255);
    while (true) {
      if (slab_cur_ < slabs_.size() && slab_off_ + bytes <= slabs_[slab_cur_].second) break;
      if (slab_cur_ < slabs_.size()) { ++slab_cur_; slab_off_ = 0; }
      if (slab_cur_ >= slabs_.size()) {
        try {
          slabs_.push_back(SlabPool::get().take(device_, std::max(kSlab, bytes)));
        } catch (const Error&) {  // out of memory even after the slab cache was trimmed: drop the recycled site buffers too
          emergency_trim(device_);
          slabs_.push_back(SlabPool::get().take(device_, std::max(kSlab, bytes)));
        }
        slab_off_ = 0;
      }
    }
    void* p = slabs_[slab_cur_].first + slab_off_;
    slab_off_ += bytes;
    return p;
  }
  bytes = (bytes + 255) & ~size_t(255);
  if (bytes == 0) bytes = 256;
  if (arena_cur_ >= arena_.size() || arena_off_ + bytes > kChunk) {
    if (arena_cur_ < arena_.size() && arena_off_ > 0) ++arena_cur_;
    if (arena_cur_ >= arena_.size()) arena_.push_back((char*)dalloc(kChunk));
    arena_off_ = 0;
  }
  void* p = arena_[arena_cur_] + arena_off_;
  arena_off_ += bytes;
  return p;
}
// hand the scratch slabs back to the process-wide cache; only when the stream has drained
void Engine::release_slabs() {
  for (auto& sl : slabs_) SlabPool::get().give(device_, sl.first, sl.second);
  slabs_.clear();
  slab_cur_ = 0; slab_off_ = 0;
}
void Engine::dfree(void* p) {
  if (!p || in_msg_arena(p)) return;  // arena messages live as long as the engine
  SlowLog sl("cudaFreeAsync");
  TNQS_CUDA(cudaFreeAsync(p, stream_));
}
void* Engine::site_alloc(size_t bytes) {
  auto f = site_pool_.find(bytes);
  if (f != site_pool_.end() && !f->second.empty()) {
    void* p = f->second.back();
    f->second.pop_back();
    site_pool_bytes_ -= bytes;
    return p;
  }
  return dalloc(bytes);
}
void Engine::site_release(void* p, size_t bytes) {
  if (!p) return;
  if (bytes < (1ull << 20)) { dfree(p); return; }  // small tensors: the pool handles them well
  site_pool_[bytes].push_back(p);
  site_pool_bytes_ += bytes;
}
void Engine::trim_site_pool(size_t keep_bytes) {
  for (auto it = site_pool_.begin(); it != site_pool_.end() && site_pool_bytes_ > keep_bytes;) {
    while (!it->second.empty() && site_pool_bytes_ > keep_bytes) {
      dfree(it->second.back());
      it->second.pop_back();
      site_pool_bytes_ -= it->first;
    }
    if (it->second.empty()) it = site_pool_.erase(it); else ++it;
  }
}
void Engine::free_temps() {
  for (void* p : temps_) TNQS_CUDA(cudaFreeAsync(p, stream_));
  temps_.clear();
  slab_cur_ = 0; slab_off_ = 0;  // slabs stay with the engine until it is destroyed
  arena_cur_ = 0;  // chunks stay cached; reuse is ordered by the single engine stream
  arena_off_ = 0;
  // close the upload epoch: its pinned mirrors may be rewritten only after the copies queued so far ran
  UpSet& s = up_[up_cur_];
  if (s.cur != 0 || s.off != 0) {
    if (!s.done) TNQS_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    TNQS_CUDA(cudaEventRecord(s.done, stream_));
    s.pending = true;
    up_cur_ = (up_cur_ + 1) % kUpSets;
    UpSet& n = up_[up_cur_];
    if (n.pending) { { WaitScope wsc(&stats_.sync_ms); TNQS_CUDA(cudaEventSynchronize(n.done)); } n.pending = false; }
    n.cur = 0; n.off = 0;
  }
}
// device block + its pinned host mirror from the current upload set
void* Engine::up_alloc(size_t bytes, void** host) {
  bytes = (bytes + 255) & ~size_t(255);
  if (bytes == 0) bytes = 256;
  if (bytes > kUpChunk) { *host = nullptr; return talloc(bytes); }  // oversized table: plain (synchronising) copy
  UpSet& s = up_[up_cur_];
  if (s.cur >= s.chunks.size() || s.off + bytes > kUpChunk) {
    if (s.cur < s.chunks.size() && s.off > 0) ++s.cur;
    if (s.cur >= s.chunks.size()) {
      auto c = UpPool::get().take(device_, kUpChunk);
      s.chunks.push_back({c.first, c.second});
    }
    s.off = 0;
  }
  *host = s.chunks[s.cur].host + s.off;
  void* d = s.chunks[s.cur].dev + s.off;
  s.off += bytes;
  return d;
}
template <class T> T* Engine::upload(const std::vector<T>& v) {
  void* h = nullptr;
  T* d = (T*)up_alloc(v.size() * sizeof(T), &h);
  if (v.empty()) return d;
  const void* src = v.data();
  if (h) { std::memcpy(h, v.data(), v.size() * sizeof(T)); src = h; }
  SlowLog sl("cudaMemcpyAsync(upload)");
  TNQS_CUDA(cudaMemcpyAsync(d, src, v.size() * sizeof(T), cudaMemcpyHostToDevice, stream_));
  return d;
}
int Engine::dedge(int src, int dst) const {
  if (src < 0 || src >= nv_ || dst < 0 || dst >= nv_) throw Error(TNQS_EINVAL, "vertex out of range");
  for (auto& l : inc_[src])
    if (l.nbr == dst) return 2 * l.edge + (eu_[l.edge] == src ? 0 : 1);
  throw Error(TNQS_ENOTADJ, "vertices " + std::to_string(src) + " and " + std::to_string(dst) + " do not share an edge");
}
int Engine::leg_pos(int v, int e) const {
  for (size_t k = 0; k < inc_[v].size(); ++k) if (inc_[v][k].edge == e) return (int)k;
  throw Error(TNQS_EINVAL, "edge not incident to vertex");
}
long long Engine::site_elems(int v) const {
  long long n = phys_[v];
  for (auto& l : inc_[v]) n *= bond_[l.edge];
  return n;
}
// view of T_v around bond leg `pos` with the physical index folded into `outer`;
// pos == -1 selects the physical index itself.
void Engine::leg_view(int v, int pos, unsigned* outer, int* chi, unsigned* inner) const {
  long long o = 1, in = 1;
  const int z = (int)inc_[v].size();
  if (pos < 0) {
    for (int k = 0; k < z; ++k) in *= bond_[inc_[v][k].edge];
    *chi = phys_[v];
  } else {
    o = phys_[v];
    for (int k = 0; k < pos; ++k) o *= bond_[inc_[v][k].edge];
    for (int k = pos + 1; k < z; ++k) in *= bond_[inc_[v][k].edge];
    *chi = bond_[inc_[v][pos].edge];
  }
  if (o * in >= (1ll << 31) || o >= (1ll << 31) || in >= (1ll << 31))
    throw Error(TNQS_EINVAL, "site tensor too large for 32-bit column indexing");
  *outer = (unsigned)o; *inner = (unsigned)in;
}
void Engine::check_shapes() const {
  for (int v = 0; v < nv_; ++v)
    for (size_t k = 0; k < inc_[v].size(); ++k)
      if (sshape_[v][k] != bond_[inc_[v][k].edge])
        throw Error(TNQS_EINVAL, "site " + std::to_string(v) + " leg " + std::to_string(k) +
                                     " disagrees with the bond dimension of its edge (set both endpoint tensors)");
}
void Engine::materialize_message(int de) {
  const int chi = bond_[de / 2];
  if (msg_[de] && msg_dim_[de] == chi) return;
  if (msg_[de]) dfree(msg_[de]);
  msg_[de] = dalloc((size_t)chi * chi * esz_);
  msg_dim_[de] = chi;
  msg_set_[de] = 0;
  std::vector<DiagTask> t(1);
  t[0].out = msg_[de]; t[0].chi = chi; t[0].diag = nullptr; t[0].scale_sumsq = nullptr;
  DiagTask* d = upload(t);
  const int nb = std::min(64, (chi * chi + 255) / 256);
  if (c64()) diag_fill_kernel<float><<<dim3(nb, 1), 256, 0, stream_>>>(d);
  else diag_fill_kernel<double><<<dim3(nb, 1), 256, 0, stream_>>>(d);
  count_launch();
}
// `need` = bytes of tensor-sized temporaries the caller is about to carve.  When they fit the scratch slabs this
// process already holds, no driver query is made: cudaMemGetInfo was measured to stall 10–70 ms sporadically
// (tools/jitter_probe.py), several times per Trotter layer.
size_t Engine::scratch_budget(size_t need) const {
  {
    size_t have = SlabPool::get().free_bytes(device_);
    for (auto& sl : slabs_) have += sl.second;
    if (need > 0 && need + need / 32 + (32ull << 20) <= have) return have;  // ≤ 1/60 of a slab is lost to packing
  }
  size_t fr = 0, tot = 0;
  { SlowLog sl("cudaMemGetInfo"); cudaMemGetInfo(&fr, &tot); }
  cudaMemPool_t pool;
  uint64_t reserved = 0, used = 0;
  if (cudaDeviceGetDefaultMemPool(&pool, device_) == cudaSuccess) {
    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
  }
  size_t mine = 0;
  for (auto& sl : slabs_) mine += sl.second;
  const size_t avail = fr + (size_t)(reserved > used ? reserved - used : 0) + SlabPool::get().free_bytes(device_) + mine + site_pool_bytes_;
  return (size_t)(0.7 * (double)avail);
}

// ------------------------------------------------------------------------------------------------
// import / export
// ------------------------------------------------------------------------------------------------
void Engine::set_site(int v, const void* data, int ndim, const int64_t* shape) {
  if (v < 0 || v >= nv_) throw Error(TNQS_EINVAL, "vertex out of range");
  const int z = (int)inc_[v].size();
  if (ndim != z + 1) throw Error(TNQS_EINVAL, "site tensor must have 1 + degree indices");
  if (shape[0] != phys_[v]) throw Error(TNQS_EINVAL, "physical dimension mismatch");
  TNQS_CUDA(cudaSetDevice(device_));
  for (int k = 0; k < z; ++k) {
    if (shape[1 + k] < 1) throw Error(TNQS_EINVAL, "bond dimension must be >= 1");
    const int e = inc_[v][k].edge;
    if (bond_[e] != (int)shape[1 + k]) {
      bond_[e] = (int)shape[1 + k];
      for (int de = 2 * e; de < 2 * e + 2; ++de) {  // messages on a resized bond fall back to the default
        if (msg_[de]) { dfree(msg_[de]); msg_[de] = nullptr; }
        msg_dim_[de] = 0; msg_set_[de] = 0;
      }
    }
    sshape_[v][k] = (int)shape[1 + k];
  }
  dfree(site_[v]);
  site_[v] = nullptr;
  if (!owns(v)) return;  // shape bookkeeping only: the tensor lives on its owner
  const size_t b = (size_t)site_elems(v) * esz_;
  site_[v] = dalloc(b);
  TNQS_CUDA(cudaMemcpyAsync(site_[v], data, b, cudaMemcpyHostToDevice, stream_));
  TNQS_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::site_shape(int v, int* ndim, int64_t* shape) const {
  if (v < 0 || v >= nv_) throw Error(TNQS_EINVAL, "vertex out of range");
  const int z = (int)inc_[v].size();
  if (*ndim < z + 1) throw Error(TNQS_ECAPACITY, "shape buffer too small");
  *ndim = z + 1;
  shape[0] = phys_[v];
  for (int k = 0; k < z; ++k) shape[1 + k] = sshape_[v][k];
}
void Engine::get_site(int v, void* data, int64_t cap) {
  if (v < 0 || v >= nv_) throw Error(TNQS_EINVAL, "vertex out of range");
  long long n = phys_[v];
  for (int x : sshape_[v]) n *= x;
  if (cap < n) throw Error(TNQS_ECAPACITY, "site buffer too small");
  if (!owns(v) || !site_[v]) throw Error(TNQS_EINVAL, "site " + std::to_string(v) + " is owned by rank " + std::to_string(owner_.empty() ? 0 : owner_[v]));
  TNQS_CUDA(cudaSetDevice(device_));
  TNQS_CUDA(cudaMemcpyAsync(data, site_[v], (size_t)n * esz_, cudaMemcpyDeviceToHost, stream_));
  TNQS_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::set_message(int src, int dst, const void* data, int chi) {
  const int de = dedge(src, dst);
  if (chi != bond_[de / 2]) throw Error(TNQS_EINVAL, "message dimension must equal the bond dimension");
  TNQS_CUDA(cudaSetDevice(device_));
  if (msg_[de] && msg_dim_[de] != chi) { dfree(msg_[de]); msg_[de] = nullptr; }
  if (!msg_[de]) msg_[de] = dalloc((size_t)chi * chi * esz_);
  msg_dim_[de] = chi;
  msg_set_[de] = 1;
  TNQS_CUDA(cudaMemcpyAsync(msg_[de], data, (size_t)chi * chi * esz_, cudaMemcpyHostToDevice, stream_));
  TNQS_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::get_message(int src, int dst, void* out, int64_t cap, int* chi, int* is_set) {
  const int de = dedge(src, dst);
  TNQS_CUDA(cudaSetDevice(device_));
  materialize_message(de);
  const int n = msg_dim_[de];
  *chi = n;
  *is_set = msg_set_[de];
  if (cap < (int64_t)n * n) throw Error(TNQS_ECAPACITY, "message buffer too small");
  TNQS_CUDA(cudaMemcpyAsync(out, msg_[de], (size_t)n * n * esz_, cudaMemcpyDeviceToHost, stream_));
  TNQS_CUDA(cudaStreamSynchronize(stream_));
  free_temps();
}
void Engine::delete_messages() {
  TNQS_CUDA(cudaSetDevice(device_));
  for (int de = 0; de < 2 * ne_; ++de) {
    if (msg_[de]) { dfree(msg_[de]); msg_[de] = nullptr; }
    msg_dim_[de] = 0; msg_set_[de] = 0;
  }
}
void Engine::get_bond_dims(int32_t* out) const {
  for (int e = 0; e < ne_; ++e) out[e] = bond_[e];
}
void Engine::set_edge_sequence(const int32_t* seq, int n) {
  std::vector<int> s(seq, seq + 2 * n);
  for (int i = 0; i < n; ++i) (void)dedge(s[2 * i], s[2 * i + 1]);
  seq_ = s;
}
void Engine::get_stats(tnqs_stats* out, int reset) {
  *out = stats_;
  if (reset) stats_ = tnqs_stats{};
}

// ------------------------------------------------------------------------------------------------
// multi-GPU plumbing: every rank holds the graph, the bond dimensions and ALL
// messages; site tensors live only on their owner.  What crosses NVLink is O(χ²) per edge:
// the messages of a BP level and the reduced-factor Gram matrices of a gate batch.
// ------------------------------------------------------------------------------------------------
void Engine::comm_init(int rank, int nranks, const void* unique_id128, const int32_t* owner) {
  if (nranks < 1 || rank < 0 || rank >= nranks) throw Error(TNQS_EINVAL, "bad rank / nranks");
  TNQS_CUDA(cudaSetDevice(device_));
  std::vector<int> own(owner, owner + nv_);
  for (int v = 0; v < nv_; ++v)
    if (own[v] < 0 || own[v] >= nranks) throw Error(TNQS_EINVAL, "owner rank out of range");
  if (nranks > 1) {
    NcclApi& api = NcclApi::get();
    NcclUniqueId id;
    std::memcpy(&id, unique_id128, sizeof(id));
    auto h = std::make_shared<CommHandle>();
    api.check(api.CommInitRank(&h->comm, nranks, id, rank), "ncclCommInitRank");
    h->rank = rank; h->nranks = nranks;
    comm_ = h;
  }
  rank_ = rank; nranks_ = nranks; owner_ = own;
  for (int v = 0; v < nv_; ++v)
    if (!owns(v) && site_[v]) { dfree(site_[v]); site_[v] = nullptr; }
  TNQS_CUDA(cudaStreamSynchronize(stream_));
}

// Every item travels from its root to all ranks.  The list is replicated information, so every rank derives the same
// packing: the items of root r go, in list order, into slots r·per … of one staging buffer; each rank packs what it owns
// (one copy kernel), ONE all-gather moves everything, one copy kernel unpacks what the rank does not own.  A BP level of
// the 16×16 lattice is 480 messages: as a group of 480 ncclBroadcast calls the exchange cost more than the level's
// tensor kernels on 8 GPUs.  (TNQS_EXCHANGE=bcast keeps the grouped broadcasts.)
void Engine::exchange(const std::vector<Bcast>& items) {
  if (nranks_ <= 1 || items.empty()) return;
  NcclApi& api = NcclApi::get();
  static const bool use_bcast = [] { const char* e = std::getenv("TNQS_EXCHANGE"); return e && std::string(e) == "bcast"; }();
  if (use_bcast || items.size() < 4) {
    api.check(api.GroupStart(), "ncclGroupStart");
    for (auto& b : items)
      api.check(api.Broadcast(b.ptr, b.ptr, b.bytes, kNcclChar, b.root, comm_->comm, stream_), "ncclBroadcast");
    api.check(api.GroupEnd(), "ncclGroupEnd");
    stats_.kernel_launches += 1;
    return;
  }
  const int R = nranks_;
  size_t slot = 0;
  std::vector<int> cnt(R, 0);
  for (auto& b : items) {
    if (b.root < 0 || b.root >= R) throw Error(TNQS_EINVAL, "exchange: bad root rank");
    slot = std::max(slot, b.bytes);
    ++cnt[b.root];
  }
  slot = (slot + 255) & ~size_t(255);
  const int per = *std::max_element(cnt.begin(), cnt.end());
  char* stage = (char*)talloc(slot * (size_t)per * R);
  std::vector<CopyTask> pack, unpack;
  std::vector<int> next(R, 0);
  for (auto& b : items) {
    char* s = stage + ((size_t)b.root * per + next[b.root]++) * slot;
    if (b.root == rank_) pack.push_back({b.ptr, s, (unsigned long long)b.bytes});
    else unpack.push_back({s, b.ptr, (unsigned long long)b.bytes});
  }
  if (!pack.empty()) {
    CopyTask* d = upload(pack);
    copy_many_kernel<<<(unsigned)pack.size(), 256, 0, stream_>>>(d);
  }
  api.check(api.AllGather(stage + (size_t)rank_ * per * slot, stage, slot * (size_t)per, kNcclChar, comm_->comm, stream_), "ncclAllGather(exchange)");
  if (!unpack.empty()) {
    CopyTask* d = upload(unpack);
    copy_many_kernel<<<(unsigned)unpack.size(), 256, 0, stream_>>>(d);
  }
  TNQS_CUDA(cudaGetLastError());
  stats_.kernel_launches += 3;
}

void Engine::allreduce_sum(double* dptr, size_t count) {
  if (nranks_ <= 1 || count == 0) return;
  NcclApi& api = NcclApi::get();
  api.check(api.AllReduce(dptr, dptr, count, kNcclFloat64, kNcclSum, comm_->comm, stream_), "ncclAllReduce");
}

// ------------------------------------------------------------------------------------------------
// kernel launch wrappers
// ------------------------------------------------------------------------------------------------
struct ProfScope {
  Engine* e; double* acc; bool on; cudaStream_t s; cudaEvent_t a, b;
  ProfScope(Engine* e_, bool on_, cudaStream_t s_, cudaEvent_t a_, cudaEvent_t b_, double* acc_)
      : e(e_), acc(acc_), on(on_), s(s_), a(a_), b(b_) { if (on) cudaEventRecord(a, s); }
  ~ProfScope() {
    if (on) { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); *acc += ms; }
  }
};

ModeTask Engine::mode_task(int v, int pos, const void* in, void* out, const void* mat) const {
  ModeTask t{};
  t.in = in; t.out = out; t.mat = mat;
  int chi;
  leg_view(v, pos, &t.outer, &chi, &t.inner);
  t.ips = t.ops = 0;
  t.chi_in = t.chi_out = chi;
  t.KK = t.MM = chi;
  t.CC = t.outer * t.inner;
  return t;
}

GramTask Engine::gram_task(int v, int pos, int planes, const void* X, const void* Y) const {
  GramTask t{};
  t.X = X; t.Y = Y;
  leg_view(v, pos, &t.outer, &t.chi, &t.inner);
  if (planes > 1) {  // physical index kept open: planes are not folded into `outer`
    t.outer /= (unsigned)phys_[v];
    t.xps = t.yps = (long long)t.outer * t.chi * t.inner;
  }
  t.MM = planes * t.chi;
  t.CC = t.outer * t.inner;
  return t;
}

template <typename R, bool I1>
static void launch_mode_variant(const ModeTask* d, int ntasks, int maxtiles, bool small, cudaStream_t s) {
  for (int off = 0; off < ntasks; off += 65535) {
    const int nb = std::min(65535, ntasks - off);
    dim3 grid(maxtiles, nb);
    if (small) mode_product_kernel<R, I1, 32, 128><<<grid, 256, 0, s>>>(d + off);
    else mode_product_kernel<R, I1, 64, 64><<<grid, 256, 0, s>>>(d + off);
  }
}

// tcgen05 path for ComplexF32 mode products (kernels_tc.cuh).  Returns the tasks it did not take.
std::vector<ModeTask> Engine::launch_mode_tc(std::vector<ModeTask>& tasks_all) {
  std::vector<ModeTask> rest;
  if (!c64() || !use_tc_) return tasks_all;
  // ---- TMA-fed warp-specialised kernel (kernels_tc2.cuh) first; what it does not take goes to the older tcgen05 kernel ----
  std::vector<ModeTask> tasks;
  if (use_tc2_) {
    tc2::Plan plan;
    std::map<tc2::ImageKey, float*> images2;
    for (auto& t : tasks_all) {
      tc2::ModeShape ms{t.in, t.out, t.mat, t.ips, t.ops, t.chi_in, t.chi_out, t.KK, t.MM, t.outer, t.inner, t.CC};
      if (!tc2::plan_add(plan, ms, [&](size_t b) { return talloc(b); }, images2)) tasks.push_back(t);
    }
    if (!plan.empty()) {
      if (!tc2::plan_finish(plan)) throw Error(TNQS_ECUDA, "tc2 mode product: shared-memory plan failed");
      tc2::PrepTask2* dp = upload(plan.preps);
      if (!plan.preps.empty()) {
        tc2::tc2_prep_kernel<<<(unsigned)plan.preps.size(), 256, 0, stream_>>>(dp);
        count_launch();
      }
      for (auto& L : plan.launches) {  // one launch per shape class (variant, K chunking, N)
        tc2::ModeTask2* dt = upload(L.tasks);
        tc2::Item* di = upload(L.items);
        if (L.last) tc2::tc2_mode_kernel<true><<<L.grid, tc2::T2_THREADS, L.smem, stream_>>>(dt, di, (int)L.items.size(), L.gm);
        else tc2::tc2_mode_kernel<false><<<L.grid, tc2::T2_THREADS, L.smem, stream_>>>(dt, di, (int)L.items.size(), L.gm);
        count_launch();
        stats_.mode_launches += 1;
        stats_.tc_launches += 1;
        stats_.tma_launches += 1;
      }
      stats_.mode_flops += plan.flops;
      TNQS_CUDA(cudaGetLastError());
    }
    if (tasks.empty()) return rest;
  } else {
    tasks = tasks_all;
  }
  struct ImgKey { const void* mat; int KK, MM, last; bool operator<(const ImgKey& o) const {
    return std::tie(mat, KK, MM, last) < std::tie(o.mat, o.KK, o.MM, o.last); } };
  std::map<ImgKey, float*> images;
  std::vector<tc::PrepTask> preps;
  std::vector<tc::TcModeTask> tt[2];
  std::vector<int> cta_task[2];
  size_t smem_max[2] = {0, 0};
  double flops = 0;
  for (auto& t : tasks) {
    const bool last = t.inner == 1;
    const int NNp = (2 * t.MM + 15) / 16 * 16;
    const int KKf = last ? 2 * t.KK : t.KK;
    const int nchunk = (KKf + tc::KC - 1) / tc::KC;
    const size_t b_bytes = (size_t)nchunk * 2 * NNp * tc::KC * 4;
    bool ok = NNp <= 256 && b_bytes <= 96 * 1024 && (double)t.CC * t.KK >= 8192.0;
    if (!last) ok = ok && (t.inner % 2 == 0) && (t.ips % 2 == 0) && (t.ops % 2 == 0);
    ok = ok && ((reinterpret_cast<uintptr_t>(t.in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(t.out) & 15) == 0);
    if (!ok) { rest.push_back(t); continue; }
    const int P_in = t.KK / t.chi_in, P_out = t.MM / t.chi_out;
    ImgKey key{t.mat, t.KK, t.MM, last ? 1 : 0};
    auto it = images.find(key);
    if (it == images.end()) {
      float* img = (float*)talloc(b_bytes);
      tc::PrepTask pt{};
      pt.mat = (const float2*)t.mat; pt.image = img; pt.KKc = t.KK; pt.MMc = t.MM; pt.NNp = NNp; pt.nchunk = nchunk; pt.last = last ? 1 : 0;
      preps.push_back(pt);
      it = images.emplace(key, img).first;
    }
    tc::TcModeTask k{};
    k.in = (const float2*)t.in; k.out = (float2*)t.out; k.image = it->second;
    k.ips = t.ips; k.ops = t.ops; k.chi_in = t.chi_in; k.chi_out = t.chi_out; k.P_in = P_in; k.P_out = P_out;
    k.KKc = t.KK; k.MMc = t.MM; k.NNp = NNp; k.nchunk = nchunk; k.outer = t.outer; k.inner = t.inner; k.CC = t.CC;
    const int cols = last ? 128 : 64;
    k.ntiles = (int)((t.CC + cols - 1) / cols);
    const int g = last ? 1 : 0;
    tt[g].push_back(k);
    // output staging aliases the A stage (2·128·KC floats = 32 KB ≥ 128·NNp·4 for NNp ≤ 64; larger N adds the rest)
    const size_t out_bytes = last ? (size_t)128 * NNp * 4 : (size_t)(NNp / 2) * 128 * 4;
    const size_t a_bytes = std::max<size_t>((size_t)2 * 128 * tc::KC * 4, out_bytes);
    smem_max[g] = std::max(smem_max[g], ((b_bytes / 4 + 255) & ~size_t(255)) * 4 + a_bytes);
    flops += 8.0 * t.KK * t.MM * (double)t.CC;
  }
  if (tt[0].empty() && tt[1].empty()) return rest;
  {
    tc::PrepTask* dp = upload(preps);
    tc::tc_prep_b_kernel<<<(unsigned)preps.size(), 256, 0, stream_>>>(dp);
    count_launch();
  }
  for (int g = 0; g < 2; ++g) {
    if (tt[g].empty()) continue;
    long long total_tiles = 0;
    for (auto& k : tt[g]) total_tiles += k.ntiles;
    const int tpc = (int)std::max<long long>(4, std::min<long long>(32, total_tiles / (148 * 8)));
    int ncta = 0;
    for (size_t i = 0; i < tt[g].size(); ++i) {
      auto& k = tt[g][i];
      k.tiles_per_cta = tpc;
      k.cta_begin = ncta;
      const int n = (k.ntiles + tpc - 1) / tpc;
      for (int c = 0; c < n; ++c) cta_task[g].push_back((int)i);
      ncta += n;
    }
    tc::TcModeTask* dt = upload(tt[g]);
    int* dc = upload(cta_task[g]);
    if (g == 0) tc::tc_mode_kernel<false><<<ncta, tc::TC_THREADS, smem_max[g], stream_>>>(dt, dc);
    else tc::tc_mode_kernel<true><<<ncta, tc::TC_THREADS, smem_max[g], stream_>>>(dt, dc);
    count_launch();
    stats_.mode_launches += 1;
    stats_.tc_launches += 1;
  }
  stats_.mode_flops += flops;
  TNQS_CUDA(cudaGetLastError());
  return rest;
}

void Engine::launch_mode(std::vector<ModeTask>& tasks_in) {
  if (tasks_in.empty()) return;
  for (auto& t : tasks_in) stats_.mode_bytes += (double)esz_ * ((double)t.KK + t.MM) * (double)t.CC;
  ProfScope ps(this, profiling_, stream_, ev0_, ev1_, &stats_.mode_ms);
  std::vector<ModeTask> tasks = launch_mode_tc(tasks_in);
  if (tasks.empty()) return;
  for (int pass = 0; pass < 2; ++pass) {
    const bool inner1 = pass == 1;
    std::vector<ModeTask> grp;
    int maxMM = 0;
    for (auto& t : tasks) if ((t.inner == 1) == inner1) { grp.push_back(t); maxMM = std::max(maxMM, t.MM); }
    if (grp.empty()) continue;
    const bool small = maxMM <= 32;
    const int tm = small ? 32 : 64, tc = small ? 128 : 64;
    int maxtiles = 0;
    for (auto& t : grp) {
      t.tiles_m = (t.MM + tm - 1) / tm;
      t.tiles_c = (int)((t.CC + tc - 1) / tc);
      maxtiles = std::max(maxtiles, t.tiles_m * t.tiles_c);
    }
    ModeTask* d = upload(grp);
    if (c64()) {
      if (inner1) launch_mode_variant<float, true>(d, (int)grp.size(), maxtiles, small, stream_);
      else launch_mode_variant<float, false>(d, (int)grp.size(), maxtiles, small, stream_);
    } else {
      if (inner1) launch_mode_variant<double, true>(d, (int)grp.size(), maxtiles, small, stream_);
      else launch_mode_variant<double, false>(d, (int)grp.size(), maxtiles, small, stream_);
    }
    count_launch(((int)grp.size() + 65534) / 65535);
    stats_.mode_launches += ((int)grp.size() + 65534) / 65535;
    for (auto& t : grp) stats_.mode_flops += 8.0 * t.KK * t.MM * (double)t.CC;
    TNQS_CUDA(cudaGetLastError());
  }
}

template <typename R, typename A, bool I1>
static void launch_gram_variant(const GramTask* d, int ntasks, int maxsplit, int maxtiles2, bool small,
                                cudaStream_t s) {
  for (int off = 0; off < ntasks; off += 65535) {
    const int nb = std::min(65535, ntasks - off);
    dim3 grid(maxsplit, maxtiles2, nb);
    if (small) gram_kernel<R, A, I1, 32><<<grid, 64, 0, s>>>(d + off);
    else gram_kernel<R, A, I1, 64><<<grid, 256, 0, s>>>(d + off);
  }
}

// outs[i] receives the MM×MM result of task i (row-major [i][j], or transposed)
void Engine::launch_gram(std::vector<GramTask>& tasks, bool acc_double, std::vector<double2*>& outs,
                         bool transpose) {
  if (tasks.empty()) return;
  ProfScope ps(this, profiling_, stream_, ev0_, ev1_, &stats_.gram_ms);
  std::vector<ReduceTask> red(tasks.size());
  std::vector<char> done(tasks.size(), 0);
  for (auto& t : tasks) stats_.gram_bytes += (double)esz_ * 2.0 * t.MM * (double)t.CC;
  // ---- TMA-fed warp-specialised tcgen05 path (kernels_tc2g.cuh): ComplexF32, fp32 accumulation (BP messages), one plane ----
  if (c64() && use_tc_ && use_tc2g_ && !acc_double) {
    tc2g::GPlan plan;
    for (size_t i = 0; i < tasks.size(); ++i) {
      const GramTask& t = tasks[i];
      if (t.MM != t.chi) continue;
      tc2g::GramShape gs{t.X, t.Y, t.chi, t.outer, t.inner, t.CC};
      if (tc2g::plan_add(plan, gs, (int)i)) done[i] = 1;
    }
    if (!plan.empty()) {
      tc2g::plan_finish(plan);
      for (auto& L : plan.launches) {
        for (size_t k = 0; k < L.tasks.size(); ++k) {
          const int i = L.ids[k];
          const int chi = tasks[i].chi;
          L.tasks[k].partial = (double2*)talloc((size_t)L.nslots[k] * chi * chi * sizeof(double2));
          ReduceTask& r = red[i];
          r.partial = L.tasks[k].partial; r.out = outs[i]; r.nsplit = L.nslots[k]; r.MM = chi; r.transpose = transpose ? 1 : 0;
          stats_.gram_flops += 8.0 * chi * chi * (double)tasks[i].CC;
        }
        tc2g::GramTask2* dt = upload(L.tasks);
        tc2g::GItem* di = upload(L.items);
        if (L.last) tc2g::tc2_gram_kernel<true><<<L.grid, tc2g::G_THREADS, L.smem, stream_>>>(dt, di, (int)L.items.size(), L.gm);
        else tc2g::tc2_gram_kernel<false><<<L.grid, tc2g::G_THREADS, L.smem, stream_>>>(dt, di, (int)L.items.size(), L.gm);
        count_launch();
        stats_.gram_launches += 1;
        stats_.tc_launches += 1;
        stats_.tma_launches += 1;
      }
      TNQS_CUDA(cudaGetLastError());
    }
  }
  // ---- tcgen05 path: ComplexF32, fp32 accumulation (BP messages), one plane, χ ≤ 64 ------------------
  if (c64() && use_tc_ && !acc_double) {
    std::vector<tc::TcGramTask> tt[2];
    std::vector<int> ids[2];
    size_t smem_max[2] = {0, 0};
    long long work[2] = {0, 0};
    for (size_t i = 0; i < tasks.size(); ++i) {
      const GramTask& t = tasks[i];
      if (done[i]) continue;
      const bool last = t.inner == 1;
      bool ok = t.MM == t.chi && t.chi <= 64 && t.chi % 2 == 0 && (double)t.CC * t.chi >= 4096.0;
      if (!last) ok = ok && t.inner % 16 == 0;
      ok = ok && ((reinterpret_cast<uintptr_t>(t.X) & 15) == 0) && ((reinterpret_cast<uintptr_t>(t.Y) & 15) == 0);
      if (!ok) continue;
      tc::TcGramTask k{};
      k.X = (const float2*)t.X; k.Y = (const float2*)t.Y; k.chi = t.chi;
      k.MMp = 2 * t.chi <= 64 ? 64 : 128;
      k.NNp = last ? (2 * t.chi + 31) / 32 * 32 : (t.chi + 15) / 16 * 16;
      k.outer = t.outer; k.inner = t.inner; k.CC = t.CC;
      const int g = last ? 1 : 0;
      tt[g].push_back(k);
      ids[g].push_back((int)i);
      smem_max[g] = std::max(smem_max[g], (size_t)2 * (2 * k.MMp + 2 * k.NNp) * tc::KC * 4);
      work[g] += t.CC;
      done[i] = 1;
    }
    for (int g = 0; g < 2; ++g) {
      if (tt[g].empty()) continue;
      const unsigned scols = g ? 32 : 16;
      // K-steps are interleaved over the CTAs of a task (kernels_tc.cuh): many splits per task keep the CTAs
      // that run together on neighbouring DRAM pages; ≥ 2048 columns (≥ 64 steps) per CTA amortise the epilogue
      const long long target_cols = std::max<long long>(2048, work[g] / (148 * 64));
      std::vector<int> cta_task;
      int ncta = 0;
      for (size_t k = 0; k < tt[g].size(); ++k) {
        tc::TcGramTask& t = tt[g][k];
        long long ns = std::max<long long>(1, (t.CC + target_cols - 1) / target_cols);
        ns = std::min<long long>(ns, 4096);
        unsigned cps = (unsigned)((t.CC + ns - 1) / ns);
        cps = (cps + scols - 1) / scols * scols;
        t.cols_per_split = cps;
        t.nsplit = (int)((t.CC + cps - 1) / cps);
        t.partial = (double2*)talloc((size_t)t.nsplit * t.chi * t.chi * sizeof(double2));
        t.cta_begin = ncta;
        for (int c = 0; c < t.nsplit; ++c) cta_task.push_back((int)k);
        ncta += t.nsplit;
        const int i = ids[g][k];
        ReduceTask& r = red[i];
        r.partial = t.partial; r.out = outs[i]; r.nsplit = t.nsplit; r.MM = t.chi; r.transpose = transpose ? 1 : 0;
        stats_.gram_flops += 8.0 * t.chi * t.chi * (double)t.CC;
      }
      tc::TcGramTask* dt = upload(tt[g]);
      int* dc = upload(cta_task);
      int maxchi = 0;
      for (auto& t : tt[g]) maxchi = std::max(maxchi, t.chi);
      if (g == 0) {
        if (maxchi <= 32) tc::tc_gram_kernel<false, 1, 2><<<ncta, tc::TC_THREADS, smem_max[g], stream_>>>(dt, dc);
        else tc::tc_gram_kernel<false, 2, 2><<<ncta, tc::TC_THREADS, smem_max[g], stream_>>>(dt, dc);
      } else {
        if (maxchi <= 16) tc::tc_gram_kernel<true, 1, 2><<<ncta, tc::TC_THREADS, smem_max[g], stream_>>>(dt, dc);
        else if (maxchi <= 32) tc::tc_gram_kernel<true, 2, 2><<<ncta, tc::TC_THREADS, smem_max[g], stream_>>>(dt, dc);
        else tc::tc_gram_kernel<true, 4, 1><<<ncta, tc::TC_THREADS, smem_max[g], stream_>>>(dt, dc);
      }
      count_launch();
      stats_.gram_launches += 1;
      stats_.tc_launches += 1;
      TNQS_CUDA(cudaGetLastError());
    }
  }
  // ---- fp64 tensor-core path (kernels_dmma.cuh): Hermitian Gram of a tensor with itself, fp64 accumulation ----
  if (acc_double && use_dmma_) {
    for (int pass = 0; pass < 2; ++pass) {
      const bool inner1 = pass == 1;
      std::vector<GramTask> grp;
      std::vector<int> ids;
      int maxMM = 0;
      long long work = 0;
      for (size_t i = 0; i < tasks.size(); ++i) {
        const GramTask& t = tasks[i];
        if (done[i] || (t.inner == 1) != inner1) continue;
        if (t.X != t.Y || t.xps != t.yps || t.MM < 16 || t.MM > 128 || t.CC < 512) continue;
        if ((double)t.CC * t.MM >= 4.0e9) continue;  // 32-bit element offsets inside the kernel
        grp.push_back(t); ids.push_back((int)i); maxMM = std::max(maxMM, t.MM); work += t.CC;
        done[i] = 1;
      }
      if (grp.empty()) continue;
      // ~4 CTAs per SM overall, at least 1024 columns per split.  One CTA is resident per SM (registers), so the launch runs
      // in waves of 148 CTAs of equal length: among a few nearby split sizes take the one that wastes least of its last wave
      // (924 CTAs = 6.24 waves ran as 7; the K-splits are reduced in a fixed order, so the result stays deterministic).
      const int NW = maxMM > 64 ? 12 : 8;  // warps per CTA (kernels_dmma.cuh); 8: one CTA covers all blocks, diagonal blocks paired
      long long target_cols = std::max<long long>(1024, work / (148 * 4));
      {
        auto ctas_for = [&](long long tc) {
          long long total = 0;
          for (auto& t : grp) {
            const long long ns = std::max<long long>(1, std::min<long long>(4096, (t.CC + tc - 1) / tc));
            long long cps = (t.CC + ns - 1) / ns;
            cps = (cps + DG_KCH - 1) / DG_KCH * DG_KCH;
            const int R32 = (2 * t.MM + 31) / 32, nblk = R32 * (R32 + 1) / 2;
            total += ((t.CC + cps - 1) / cps) * (NW == 8 ? 1 : (nblk + NW - 1) / NW);
          }
          return total;
        };
        double best_eff = -1.0;
        long long best_tc = target_cols;
        for (int pct : {100, 90, 80, 72, 64, 56, 50}) {
          const long long tc = std::max<long long>(1024, target_cols * pct / 100);
          const long long total = ctas_for(tc);
          const double eff = (double)total / (double)((total + 147) / 148 * 148);
          if (eff > best_eff + 0.02) { best_eff = eff; best_tc = tc; }  // prefer the larger split unless a smaller one gains ≥ 2 %
        }
        target_cols = best_tc;
      }
      int maxsplit = 1, maxgroups = 1;
      for (size_t k = 0; k < grp.size(); ++k) {
        GramTask& t = grp[k];
        long long ns = std::max<long long>(1, std::min<long long>(4096, (t.CC + target_cols - 1) / target_cols));
        unsigned cps = (unsigned)((t.CC + ns - 1) / ns);
        cps = (cps + DG_KCH - 1) / DG_KCH * DG_KCH;
        t.cols_per_split = cps;
        t.nsplit = (int)((t.CC + cps - 1) / cps);
        t.partial = (double2*)talloc((size_t)t.nsplit * t.MM * t.MM * sizeof(double2));
        maxsplit = std::max(maxsplit, t.nsplit);
        const int R32 = (2 * t.MM + 31) / 32, nblk = R32 * (R32 + 1) / 2;
        maxgroups = std::max(maxgroups, NW == 8 ? 1 : (nblk + NW - 1) / NW);
        ReduceTask& r = red[ids[k]];
        r.partial = t.partial; r.out = outs[ids[k]]; r.nsplit = t.nsplit; r.MM = t.MM; r.transpose = transpose ? 1 : 0;
        stats_.gram_flops += 8.0 * t.MM * t.MM * (double)t.CC;
      }
      const int prow = (2 * maxMM + 31) / 32 * 32;
      const size_t smem = (size_t)2 * prow * DG_LD * sizeof(double);
      GramTask* d = upload(grp);
      for (int off = 0; off < (int)grp.size(); off += 65535) {
        const int nbz = std::min(65535, (int)grp.size() - off);
        dim3 grid(maxsplit, maxgroups, nbz);
#define TNQS_DG(RT, I1) do { if (NW == 12) gram_dmma_kernel<RT, I1, 12><<<grid, 12 * 32, smem, stream_>>>(d + off); \
                             else gram_dmma_kernel<RT, I1, 8><<<grid, 8 * 32, smem, stream_>>>(d + off); } while (0)
        if (c64()) { if (inner1) TNQS_DG(float, true); else TNQS_DG(float, false); }
        else { if (inner1) TNQS_DG(double, true); else TNQS_DG(double, false); }
#undef TNQS_DG
        count_launch();
        stats_.gram_launches += 1;
      }
      TNQS_CUDA(cudaGetLastError());
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    const bool inner1 = pass == 1;
    std::vector<GramTask> grp;
    std::vector<int> ids;
    int maxMM = 0;
    for (size_t i = 0; i < tasks.size(); ++i)
      if (!done[i] && (tasks[i].inner == 1) == inner1) { grp.push_back(tasks[i]); ids.push_back((int)i); maxMM = std::max(maxMM, tasks[i].MM); }
    if (grp.empty()) continue;
    const bool small = maxMM <= 32;
    const int ti = small ? 32 : 64;
    int maxsplit = 1, maxtiles2 = 1;
    long long ctas_unsplit = 0;
    for (auto& t : grp) { t.tiles = (t.MM + ti - 1) / ti; ctas_unsplit += (long long)t.tiles * t.tiles; }
    const long long target = 148ll * 8;
    for (size_t k = 0; k < grp.size(); ++k) {
      GramTask& t = grp[k];
      long long want = (target + ctas_unsplit - 1) / ctas_unsplit;        // splits wanted per task
      const long long cap = std::max<long long>(1, t.CC / 256);           // ≥256 columns per split
      long long ns = std::max<long long>(1, std::min(want, cap));
      ns = std::min<long long>(ns, 1024);
      unsigned cps = (unsigned)((t.CC + ns - 1) / ns);
      cps = (cps + TK - 1) / TK * TK;
      t.cols_per_split = cps;
      t.nsplit = (int)((t.CC + cps - 1) / cps);
      t.partial = (double2*)talloc((size_t)t.nsplit * t.MM * t.MM * sizeof(double2));
      maxsplit = std::max(maxsplit, t.nsplit);
      maxtiles2 = std::max(maxtiles2, t.tiles * t.tiles);
      ReduceTask& r = red[ids[k]];
      r.partial = t.partial; r.out = outs[ids[k]]; r.nsplit = t.nsplit; r.MM = t.MM; r.transpose = transpose ? 1 : 0;
    }
    GramTask* d = upload(grp);
    const int n = (int)grp.size();
    if (c64()) {
      if (acc_double) {
        if (inner1) launch_gram_variant<float, double, true>(d, n, maxsplit, maxtiles2, small, stream_);
        else launch_gram_variant<float, double, false>(d, n, maxsplit, maxtiles2, small, stream_);
      } else {
        if (inner1) launch_gram_variant<float, float, true>(d, n, maxsplit, maxtiles2, small, stream_);
        else launch_gram_variant<float, float, false>(d, n, maxsplit, maxtiles2, small, stream_);
      }
    } else {
      if (inner1) launch_gram_variant<double, double, true>(d, n, maxsplit, maxtiles2, small, stream_);
      else launch_gram_variant<double, double, false>(d, n, maxsplit, maxtiles2, small, stream_);
    }
    count_launch((n + 65534) / 65535);
    stats_.gram_launches += (n + 65534) / 65535;
    for (auto& t : grp) stats_.gram_flops += 8.0 * t.MM * t.MM * (double)t.CC;
    TNQS_CUDA(cudaGetLastError());
  }
  ReduceTask* dr = upload(red);
  int maxMM = 0;
  for (auto& r : red) maxMM = std::max(maxMM, r.MM);
  const int nb = std::max(1, std::min(64, (maxMM * maxMM + 255) / 256));
  for (int off = 0; off < (int)red.size(); off += 65535) {
    const int cnt = std::min(65535, (int)red.size() - off);
    gram_reduce_kernel<<<dim3(nb, cnt), 256, 0, stream_>>>(dr + off);
    count_launch();
  }
  TNQS_CUDA(cudaGetLastError());
}

template <int LPP, int RPL, int MAXT, int MINB, bool LOCALP>
static void launch_jacobi_cluster(const JacobiTask* d, JacobiAux* aux, int ntasks, int BC, int C, int ld, size_t smem,
                                  double dead_rel2, double* nonconv, cudaStream_t s) {
  // the shared-memory opt-in is a per-device attribute of the function
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    TNQS_CUDA(cudaFuncSetAttribute(jacobi_cluster_kernel<LPP, RPL, MAXT, MINB, LOCALP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(ntasks * C));
  cfg.blockDim = dim3((unsigned)std::max(32, BC * LPP));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  SlowLog sl("cudaLaunchKernelEx(jacobi)");
  TNQS_CUDA(cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel<LPP, RPL, MAXT, MINB, LOCALP>, d, aux, BC, C, ld, 40, 8.9e-16, dead_rel2, nonconv));
}

void Engine::launch_jacobi(std::vector<JacobiTask>& tasks, double dead_rel2) {
  if (tasks.empty()) return;
  ProfScope ps(this, profiling_, stream_, ev0_, ev1_, &stats_.small_ms);
  int maxn = 0, maxm = 0, maxmt = 0;
  for (auto& t : tasks) {
    maxn = std::max(maxn, t.n);
    maxm = std::max(maxm, std::max(t.m, t.n));
    maxmt = std::max(maxmt, t.m + (t.V ? t.n : 0));
  }
  JacobiTask* d = upload(tasks);
  const unsigned nb = (unsigned)tasks.size();
  if (maxmt <= 256 && maxn <= 256 && use_cluster_jacobi_) {
    // shared-memory cluster kernel (kernels_jacobi.cuh)
    int LPP = maxmt <= 128 ? 16 : 32;
    if (const char* e = std::getenv("TNQS_JACOBI_LPP")) { if (std::atoi(e) == 32 && maxmt > 32) LPP = 32; }  // tuning override
    int RPL = 1;
    while (RPL * LPP < maxmt) RPL <<= 1;
    int npad = 2;
    while (npad < maxn) npad <<= 1;
    const int ld = RPL * LPP;  // ≥ maxmt: the pad rows of every shared-memory column are kept at zero
    int BC, C;
    if (npad <= 32) { BC = npad / 2; C = 1; }
    else {
      // both splits run n−1 rounds per sweep; take the one that needs fewer waves of CTAs, then the larger block
      auto waves = [&](int bc) {
        const int c = npad / (2 * bc);
        const size_t sm = (size_t)2 * bc * ld * sizeof(double2) + 2048;
        const int thr = bc * LPP;
        const int per_sm = std::max(1, (int)std::min<size_t>(std::min<size_t>(227 * 1024 / sm, 2048 / thr), 65536 / (thr * (RPL >= 8 ? 128 : 64))));
        return (int)(((long long)nb * c + 148ll * per_sm - 1) / (148ll * per_sm));
      };
      const bool can16 = npad / 32 <= 8, can8 = npad / 16 <= 8;
      if (can16 && (!can8 || waves(16) <= waves(8))) { BC = 16; C = npad / 32; }
      else { BC = 8; C = npad / 16; }
      if (const char* e = std::getenv("TNQS_JACOBI_BC")) {  // tuning override
        const int want = std::atoi(e);
        if (want == 8 && can8) { BC = 8; C = npad / 16; }
        if (want == 16 && can16) { BC = 16; C = npad / 32; }
      }
    }
    const size_t smem = (size_t)2 * BC * ld * sizeof(double2);
    JacobiAux* aux = (JacobiAux*)talloc(sizeof(JacobiAux) * nb);
    TNQS_CUDA(cudaMemsetAsync(aux, 0, sizeof(JacobiAux) * nb, stream_));
    // register budget: MINB CTAs of MAXT threads per SM
    const int thr = std::max(32, BC * LPP);
    // launches that leave SMs idle are latency-bound: every lane group then computes its own rotation (one barrier per
    // round, kernels_jacobi.cuh); TNQS_JACOBI_LOCALP=0/1 forces the choice
    bool localp = (long long)nb * C <= 148;
    if (const char* e = std::getenv("TNQS_JACOBI_LOCALP")) localp = std::atoi(e) != 0;
#define TNQS_JAC2(L, R, P) do { if (thr <= 128) launch_jacobi_cluster<L, R, 128, (R >= 8 ? 4 : 6), P>(d, aux, (int)nb, BC, C, ld, smem, dead_rel2, d_errflags_, stream_); \
    else if (thr <= 256) launch_jacobi_cluster<L, R, 256, (R >= 8 ? 2 : 3), P>(d, aux, (int)nb, BC, C, ld, smem, dead_rel2, d_errflags_, stream_); \
    else launch_jacobi_cluster<L, R, 512, 1, P>(d, aux, (int)nb, BC, C, ld, smem, dead_rel2, d_errflags_, stream_); } while (0)
#define TNQS_JAC(L, R) do { if (localp) TNQS_JAC2(L, R, true); else TNQS_JAC2(L, R, false); } while (0)
    bool launched = true;
    try {
      if (LPP == 16) {
        if (RPL <= 1) TNQS_JAC(16, 1);
        else if (RPL == 2) TNQS_JAC(16, 2);
        else if (RPL == 4) TNQS_JAC(16, 4);
        else TNQS_JAC(16, 8);
      } else {
        if (RPL <= 2) TNQS_JAC(32, 2);
        else if (RPL == 4) TNQS_JAC(32, 4);
        else TNQS_JAC(32, 8);
      }
    } catch (const Error&) {
      // a cluster shape the device refuses (launch errors are synchronous): the global-memory kernel below
      // factorises the same task table
      cudaGetLastError();
      launched = false;
    }
#undef TNQS_JAC
#undef TNQS_JAC2
    if (launched) {
    count_launch();
    TNQS_CUDA(cudaGetLastError());
    const bool dbg = std::getenv("TNQS_JACOBI_DEBUG") != nullptr;
    if (dbg) {  // sweeps actually run per matrix (rot[s] is set by every sweep that rotated something)
      std::vector<JacobiAux> h(nb);
      TNQS_CUDA(cudaMemcpyAsync(h.data(), aux, sizeof(JacobiAux) * nb, cudaMemcpyDeviceToHost, stream_));
      TNQS_CUDA(cudaStreamSynchronize(stream_));
      int mn = 1000, mx = 0; double avg = 0;
      for (auto& a : h) { int c = 0; for (int i = 0; i < 64; ++i) c += a.rot[i] != 0; mn = std::min(mn, c); mx = std::max(mx, c); avg += c; }
      std::fprintf(stderr, "[jacobi] tasks=%u maxn=%d maxmt=%d LPP=%d RPL=%d BC=%d C=%d smem=%zu rotating sweeps min/avg/max=%d/%.1f/%d\n",
                   nb, maxn, maxmt, LPP, RPL, BC, C, smem, mn, avg / nb, mx);
    }
    return;
    }
  }
  const int pairs = (maxn + 1) / 2;
  const int maxwarps = maxm <= 128 ? 32 : (maxm <= 256 ? 16 : 8);  // register budget per lane grows with m
  const int warps = std::max(1, std::min(maxwarps, pairs));
  if (maxm <= 32) jacobi_kernel<1><<<nb, warps * 32, 0, stream_>>>(d, 40, 8.9e-16, d_errflags_);
  else if (maxm <= 64) jacobi_kernel<2><<<nb, warps * 32, 0, stream_>>>(d, 40, 8.9e-16, d_errflags_);
  else if (maxm <= 128) jacobi_kernel<4><<<nb, warps * 32, 0, stream_>>>(d, 40, 8.9e-16, d_errflags_);
  else if (maxm <= 256) jacobi_kernel<8><<<nb, warps * 32, 0, stream_>>>(d, 40, 8.9e-16, d_errflags_);
  else if (maxm <= 512) jacobi_kernel<16><<<nb, warps * 32, 0, stream_>>>(d, 40, 8.9e-16, d_errflags_);
  else throw Error(TNQS_EINVAL, "matrix too large for the batched Jacobi kernel (max 512 rows)");
  count_launch();
  TNQS_CUDA(cudaGetLastError());
}

// Run every chain: round r applies step r of every chain that has one; ping-pong scratch buffers.
void Engine::run_chains(std::vector<Chain>& chains) {
  size_t maxsteps = 0;
  std::vector<void*> bufA(chains.size(), nullptr), bufB(chains.size(), nullptr);
  for (size_t i = 0; i < chains.size(); ++i) {
    Chain& c = chains[i];
    if (!owns(c.v)) { c.steps.clear(); c.result = nullptr; continue; }
    maxsteps = std::max(maxsteps, c.steps.size());
    c.result = site_[c.v];
    if (c.steps.empty()) continue;
    const size_t b = (size_t)site_elems(c.v) * esz_;
    bufA[i] = talloc(b);
    if (c.steps.size() > 1) bufB[i] = talloc(b);
  }
  for (size_t r = 0; r < maxsteps; ++r) {
    std::vector<ModeTask> tasks;
    for (size_t i = 0; i < chains.size(); ++i) {
      Chain& c = chains[i];
      if (c.steps.size() <= r) continue;
      const void* in = r == 0 ? site_[c.v] : ((r - 1) % 2 == 0 ? bufA[i] : bufB[i]);
      void* out = (r % 2 == 0) ? bufA[i] : bufB[i];
      tasks.push_back(mode_task(c.v, c.steps[r].first, in, out, c.steps[r].second));
      c.result = out;
    }
    launch_mode(tasks);
  }
}

// ------------------------------------------------------------------------------------------------
// belief propagation
// ------------------------------------------------------------------------------------------------

// Group the sequential edge schedule into dependency levels that reproduce Gauss–Seidel semantics:
// an update must come after earlier updates of the messages it reads (RAW), not before earlier
// readers of the message it writes (WAR; same level is fine because a level commits at its end).
std::vector<std::vector<int>> Engine::bp_levels(const std::vector<int>& seq) const {
  const int n = (int)seq.size() / 2;
  std::vector<int> lw(2 * ne_, -1), lr(2 * ne_, -1), lev(n, 0);
  int nlev = 0;
  for (int i = 0; i < n; ++i) {
    const int u = seq[2 * i], v = seq[2 * i + 1];
    const int w = dedge(u, v);
    int L = 0;
    for (auto& l : inc_[u]) {
      if (l.nbr == v) continue;
      const int r = dedge(l.nbr, u);
      if (lw[r] >= 0) L = std::max(L, lw[r] + 1);
    }
    L = std::max(L, lr[w]);
    if (lw[w] >= 0) L = std::max(L, lw[w] + 1);
    lev[i] = L;
    lw[w] = L;
    for (auto& l : inc_[u]) {
      if (l.nbr == v) continue;
      const int r = dedge(l.nbr, u);
      lr[r] = std::max(lr[r], L);
    }
    nlev = std::max(nlev, L + 1);
  }
  std::vector<std::vector<int>> out(nlev);
  for (int i = 0; i < n; ++i) out[lev[i]].push_back(i);
  return out;
}

// one level: every item i is the update of message seq[i] (abstractbeliefpropagationcache.jl:162-190)
void Engine::bp_level(const std::vector<int>& seq, const std::vector<int>& all_items, double* d_diff) {
  // staging buffers for every message of the level (all ranks), compute only what this rank owns
  std::vector<int> items;
  for (int it : all_items) {
    const int de = dedge(seq[2 * it], seq[2 * it + 1]);
    const int chi = bond_[de / 2];
    if (!msg_next_[de] || msg_next_dim_[de] != chi) {
      if (msg_next_[de]) dfree(msg_next_[de]);
      msg_next_[de] = dalloc((size_t)chi * chi * esz_);
      msg_next_dim_[de] = chi;
    }
    if (owns(seq[2 * it])) items.push_back(it);
  }
  // Messages leaving the same vertex share partial products: with the legs split recursively in
  // halves, the branch towards one half first absorbs every message of the other half, so a degree-4
  // vertex sending on all legs costs 8 mode products instead of 4·3 (degree 6: 16 instead of 30).
  struct Group { int u; std::vector<int> its; };
  std::vector<Group> groups;
  {
    std::map<int, int> gidx;
    for (int it : items) {
      const int u = seq[2 * it];
      auto f = gidx.find(u);
      if (f == gidx.end()) { f = gidx.emplace(u, (int)groups.size()).first; groups.push_back({u, {}}); }
      groups[f->second].its.push_back(it);
    }
  }
  struct Node { int v, pos; const void* mat; int in; void* out; int level; };
  std::vector<Node> nodes;
  std::vector<int> leaf;  // per leg position of the current vertex: node id holding the result (−1: the site tensor)
  std::vector<char> want, isset;
  std::vector<const void*> lmat;
  // cur = node that has absorbed every set leg outside L; returns nothing, fills leaf[]
  std::function<void(const std::vector<int>&, int, int, int)> build = [&](const std::vector<int>& L, int cur, int level, int u) {
    bool any = false;
    for (int p : L) any |= want[p] != 0;
    if (!any) return;
    if (L.size() == 1) { leaf[L[0]] = cur; return; }
    const size_t h = L.size() / 2;
    const std::vector<int> L1(L.begin(), L.begin() + h), L2(L.begin() + h, L.end());
    for (int side = 0; side < 2; ++side) {
      const std::vector<int>& keep = side == 0 ? L1 : L2;
      const std::vector<int>& absorb = side == 0 ? L2 : L1;
      bool w = false;
      for (int p : keep) w |= want[p] != 0;
      if (!w) continue;
      int x = cur, lv = level;
      for (int p : absorb) {
        if (!isset[p]) continue;  // identity default: nothing to absorb
        nodes.push_back({u, p, lmat[p], x, nullptr, lv});
        x = (int)nodes.size() - 1;
        ++lv;
      }
      build(keep, x, lv, u);
    }
  };
  auto plan_group = [&](const Group& g, std::vector<std::pair<int, int>>& results /* (item, node) */) {
    const int u = g.u, z = (int)inc_[u].size();
    want.assign(z, 0); isset.assign(z, 0); lmat.assign(z, nullptr); leaf.assign(z, -1);
    for (int p = 0; p < z; ++p) {
      const int de = dedge(inc_[u][p].nbr, u);
      isset[p] = msg_set_[de] ? 1 : 0;
      lmat[p] = msg_[de];
    }
    std::vector<int> legs(z);
    for (int p = 0; p < z; ++p) legs[p] = p;
    for (int it : g.its) want[leg_pos(u, dedge(u, seq[2 * it + 1]) / 2)] = 1;
    build(legs, -1, 0, u);
    for (int it : g.its) results.push_back({it, leaf[leg_pos(u, dedge(u, seq[2 * it + 1]) / 2)]});
  };
  size_t need_total = 0;
  {
    std::vector<std::pair<int, int>> dry;
    for (auto& gr : groups) {
      const size_t n0 = nodes.size();
      plan_group(gr, dry);
      need_total += (nodes.size() - n0) * (((size_t)site_elems(gr.u) * esz_ + 255) & ~size_t(255));
    }
    nodes.clear();
  }
  const size_t budget = scratch_budget(need_total);
  size_t pos = 0;
  while (pos < groups.size()) {
    // chunk by scratch: one buffer per product node
    nodes.clear();
    std::vector<std::pair<int, int>> results;
    size_t end = pos, bytes = 0;
    while (end < groups.size()) {
      const size_t n0 = nodes.size(), r0 = results.size();
      plan_group(groups[end], results);
      const size_t need = (nodes.size() - n0) * (size_t)site_elems(groups[end].u) * esz_;
      if (end > pos && bytes + need > budget) { nodes.resize(n0); results.resize(r0); break; }
      bytes += need;
      ++end;
    }
    int maxlevel = -1;
    for (auto& nd : nodes) {
      nd.out = talloc((size_t)site_elems(nd.v) * esz_);
      maxlevel = std::max(maxlevel, nd.level);
    }
    for (int lv = 0; lv <= maxlevel; ++lv) {
      std::vector<ModeTask> tasks;
      for (auto& nd : nodes)
        if (nd.level == lv) tasks.push_back(mode_task(nd.v, nd.pos, nd.in < 0 ? site_[nd.v] : nodes[nd.in].out, nd.out, nd.mat));
      launch_mode(tasks);
    }
    const int n = (int)results.size();
    std::vector<GramTask> gt(n);
    std::vector<double2*> outs(n);
    std::vector<BpFinTask> fin(n);
    for (int k = 0; k < n; ++k) {
      const int it = results[k].first;
      const int u = seq[2 * it], v = seq[2 * it + 1];
      const int de = dedge(u, v);
      const int e = de / 2;
      gt[k] = gram_task(u, leg_pos(u, e), 1, site_[u], results[k].second < 0 ? site_[u] : nodes[results[k].second].out);
      const int chi = bond_[e];
      outs[k] = (double2*)talloc((size_t)chi * chi * sizeof(double2));
      fin[k].g = outs[k]; fin[k].old_msg = msg_[de]; fin[k].new_msg = msg_next_[de];
      fin[k].diff = d_diff + it; fin[k].chi = chi;
    }
    launch_gram(gt, /*acc_double=*/false, outs, /*transpose=*/true);
    {
      BpFinTask* df = upload(fin);
      if (c64()) bp_finalize_kernel<float><<<n, 256, 0, stream_>>>(df);
      else bp_finalize_kernel<double><<<n, 256, 0, stream_>>>(df);
      count_launch();
      TNQS_CUDA(cudaGetLastError());
    }
    free_temps();
    pos = end;
  }
  if (nranks_ > 1) {  // the level's new messages travel from the owner of their source vertex
    std::vector<Bcast> bc;
    for (int it : all_items) {
      const int de = dedge(seq[2 * it], seq[2 * it + 1]);
      bc.push_back({msg_next_[de], (size_t)msg_next_dim_[de] * msg_next_dim_[de] * esz_, owner_[seq[2 * it]]});
    }
    exchange(bc);
  }
  for (int it : all_items) {  // commit the level
    const int de = dedge(seq[2 * it], seq[2 * it + 1]);
    std::swap(msg_[de], msg_next_[de]);
    std::swap(msg_dim_[de], msg_next_dim_[de]);
    msg_set_[de] = 1;
  }
  stats_.bp_messages += (int64_t)items.size();
}

namespace {
struct WallScope {  // host wall time of the outermost hot-path call
  double* acc; int* depth; std::chrono::steady_clock::time_point t0;
  WallScope(double* a, int* d) : acc(a), depth(d), t0(std::chrono::steady_clock::now()) { ++*depth; }
  ~WallScope() {
    if (--*depth == 0) *acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};
// timing events of one hot-path call; destroyed on every exit path (an Error thrown mid-call must not leak them)
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() { TNQS_CUDA(cudaEventCreate(&a)); if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); a = nullptr; throw Error(TNQS_ECUDA, "cudaEventCreate failed"); } }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  EventPair(const EventPair&) = delete;
  EventPair& operator=(const EventPair&) = delete;
};
}  // namespace

tnqs_bp_report Engine::bp_update(const tnqs_bp_opts* o) {
  WallScope ws(&stats_.wall_ms, &wall_depth_);
  TNQS_CUDA(cudaSetDevice(device_));
  check_shapes();
  std::vector<int> seq;
  if (o && o->edge_sequence && o->n_seq > 0) seq.assign(o->edge_sequence, o->edge_sequence + 2 * o->n_seq);
  else seq = seq_;
  if (seq.empty() && ne_ > 0)
    throw Error(TNQS_EINVAL, "no BP edge sequence: call tnqs_set_edge_sequence or pass one in tnqs_bp_opts");
  const int nseq = (int)seq.size() / 2;
  // defaults (beliefpropagationcache.jl:103-119)
  int maxiter = is_tree_ ? 1 : 25;
  bool use_tol = !is_tree_;
  double tol = c64() ? 1e-5 : 1e-8;
  if (o) {
    if (o->maxiter > 0) maxiter = o->maxiter;
    use_tol = o->use_tolerance != 0;
    if (o->tolerance >= 0) tol = o->tolerance;
  }
  tnqs_bp_report rep{maxiter, use_tol ? 0 : 1, 0.0};
  if (nseq == 0) { rep.niter = 0; rep.converged = 1; return rep; }
  EventPair ev;
  const cudaEvent_t t0 = ev.a, t1 = ev.b;
  TNQS_CUDA(cudaEventRecord(t0, stream_));
  for (int de = 0; de < 2 * ne_; ++de) materialize_message(de);
  free_temps();
  const auto levels = bp_levels(seq);
  struct DiffBuf {  // released on every exit path
    Engine* e; double* p;
    ~DiffBuf() { if (p) e->dfree(p); }
  } diff_buf{this, (double*)dalloc(sizeof(double) * nseq)};
  double* const d_diff = diff_buf.p;
  std::vector<double> h_diff(nseq);
  for (int it = 1; it <= maxiter; ++it) {
    if (nranks_ > 1) TNQS_CUDA(cudaMemsetAsync(d_diff, 0, sizeof(double) * nseq, stream_));
    for (auto& lv : levels) bp_level(seq, lv, d_diff);
    stats_.bp_sweeps += 1;
    if (use_tol) {
      allreduce_sum(d_diff, nseq);
      TNQS_CUDA(cudaMemcpyAsync(h_diff.data(), d_diff, sizeof(double) * nseq, cudaMemcpyDeviceToHost, stream_));
      { WaitScope wsc(&stats_.sync_ms); TNQS_CUDA(cudaStreamSynchronize(stream_)); }
      double s = 0;
      for (double x : h_diff) s += x;
      rep.diff = s / nseq;
      if (rep.diff <= tol) { rep.converged = 1; rep.niter = it; break; }
    }
  }
  TNQS_CUDA(cudaEventRecord(t1, stream_));
  { WaitScope wsc(&stats_.sync_ms); TNQS_CUDA(cudaEventSynchronize(t1)); }
  float ms = 0;
  cudaEventElapsedTime(&ms, t0, t1);
  stats_.bp_ms += ms;
  if (wall_depth_ == 1) release_slabs();  // outermost call, stream drained by the event wait above
  return rep;
}

// ------------------------------------------------------------------------------------------------
// gate application
// ------------------------------------------------------------------------------------------------
void Engine::normalize_sites(const std::vector<int>& vs) {
  if (vs.empty()) return;
  std::vector<NormTask> t(vs.size());
  for (size_t i = 0; i < vs.size(); ++i) {
    t[i].data = site_[vs[i]];
    t[i].n = site_elems(vs[i]);
    t[i].partial = (double*)talloc(sizeof(double) * NORM_BLOCKS);
  }
  NormTask* d = upload(t);
  long long maxn = 0;
  for (auto& x : t) maxn = std::max(maxn, x.n);
  const int nb = (int)std::max<long long>(1, std::min<long long>(512, (maxn + 1023) / 1024));
  for (int off = 0; off < (int)t.size(); off += 65535) {
    const int cnt = std::min(65535, (int)t.size() - off);
    if (c64()) {
      sumsq_kernel<float><<<dim3(NORM_BLOCKS, cnt), 256, 0, stream_>>>(d + off);
      scale_kernel<float><<<dim3(nb, cnt), 256, 0, stream_>>>(d + off);
    } else {
      sumsq_kernel<double><<<dim3(NORM_BLOCKS, cnt), 256, 0, stream_>>>(d + off);
      scale_kernel<double><<<dim3(nb, cnt), 256, 0, stream_>>>(d + off);
    }
    count_launch(2);
  }
  TNQS_CUDA(cudaGetLastError());
}

void Engine::apply_one_site_batch(const std::vector<std::pair<int, std::vector<cplx>>>& g, bool normalize) {
  if (g.empty()) return;
  std::vector<OneSiteTask> t(g.size());
  std::vector<int> vs;
  long long maxplane = 0;
  for (size_t i = 0; i < g.size(); ++i) {
    const int v = g[i].first;
    const int d = phys_[v];
    if (!owns(v)) { t[i].data = nullptr; t[i].plane = 0; t[i].d = d; continue; }
    t[i].data = site_[v];
    t[i].plane = site_elems(v) / d;
    t[i].d = d;
    for (int k = 0; k < d * d; ++k) { t[i].U[k].x = g[i].second[k].real(); t[i].U[k].y = g[i].second[k].imag(); }
    maxplane = std::max(maxplane, t[i].plane);
    vs.push_back(v);
  }
  OneSiteTask* dt = upload(t);
  const int nb = (int)std::max<long long>(1, std::min<long long>(512, (maxplane + 255) / 256));
  for (int off = 0; off < (int)t.size(); off += 65535) {
    const int cnt = std::min(65535, (int)t.size() - off);
    if (c64()) onesite_kernel<float><<<dim3(nb, cnt), 256, 0, stream_>>>(dt + off);
    else onesite_kernel<double><<<dim3(nb, cnt), 256, 0, stream_>>>(dt + off);
    count_launch();
  }
  TNQS_CUDA(cudaGetLastError());
  if (normalize) normalize_sites(vs);
  free_temps();
}

// One batch of vertex-disjoint two-site gates: the whole of simple_update's two-site branch
// (simple_update.jl:29-68) plus apply_gate!'s write-back (apply_gates.jl:126-140), batched.
void Engine::apply_two_site_batch(const std::vector<int>& gate_ids, const int32_t* verts,
                                  const double* mats, const std::vector<size_t>& mat_off,
                                  const tnqs_apply_opts& ao, double* errs) {
  if (gate_ids.empty()) return;
  const double eps = c64() ? 1.1920928955078125e-07 : 2.220446049250313e-16;
  const double sqrt_cutoff = ao.sqrt_cutoff >= 0 ? ao.sqrt_cutoff : 10 * eps;  // simple_update.jl:32-33
  const bool normalize = ao.normalize_tensors != 0;
  // Scratch of a gate: three tensor-sized buffers per site (gauged, projected, new), on the rank that owns the site.  A batch
  // is cut into chunks that fit the budget of the most loaded rank; every rank derives the same cuts from replicated
  // information (ownership, shapes), because chunk boundaries carry collectives.  Fewer chunks matter: the O(χ³) chain of a
  // chunk (Cholesky / Jacobi launches) is latency-bound and costs the same for 30 or 120 gates.
  const int RR = nranks_ > 1 ? nranks_ : 1;
  auto site_need = [&](int v) { return 3 * (size_t)site_elems(v) * esz_; };
  auto rank_of = [&](int v) { return owner_.empty() ? 0 : owner_[v]; };
  size_t budget;
  if (RR == 1) {
    size_t need_total = 0;
    for (int g : gate_ids) need_total += site_need(verts[2 * g]) + site_need(verts[2 * g + 1]);
    budget = scratch_budget(need_total);
  } else {
    std::vector<size_t> state(RR, 0);
    for (int v = 0; v < nv_; ++v) state[rank_of(v)] += (size_t)site_elems(v) * esz_;
    const size_t worst = *std::max_element(state.begin(), state.end());
    const size_t hbm = (size_t)170 << 30;  // usable HBM of a B200 (183 GB) minus context, pools and fragmentation slack
    budget = worst + ((size_t)8 << 30) < hbm ? (size_t)(0.6 * (double)(hbm - worst)) : ((size_t)8 << 30);
  }
  size_t gpos = 0;
  while (gpos < gate_ids.size()) {
    size_t gend = gpos;
    std::vector<size_t> used(RR, 0);
    while (gend < gate_ids.size()) {
      const int g = gate_ids[gend];
      const int a = verts[2 * g], b = verts[2 * g + 1];
      std::vector<size_t> next = used;
      next[rank_of(a)] += site_need(a);
      next[rank_of(b)] += site_need(b);
      if (gend > gpos && *std::max_element(next.begin(), next.end()) > budget) break;
      used.swap(next);
      ++gend;
    }
    const int ng = (int)(gend - gpos);
    // TNQS_SLOWLOG=1: host time of every phase of a batch that took longer than 60 ms in total
    std::vector<std::pair<const char*, double>> phase_log;
    auto phase_t0 = std::chrono::steady_clock::now();
    auto phase_mark = [&](const char* name) {
      if (!SlowLog::on()) return;
      const auto now = std::chrono::steady_clock::now();
      phase_log.push_back({name, std::chrono::duration<double, std::milli>(now - phase_t0).count()});
      phase_t0 = now;
    };

    // ---- 1. environments: eigendecompose every non-default incoming message --------------------
    struct EnvRef { int gate, site, pos, de, task; };
    std::vector<EnvRef> envs;
    std::vector<MsgEigTask> mt;
    for (int k = 0; k < ng; ++k) {
      const int g = gate_ids[gpos + k];
      for (int s = 0; s < 2; ++s) {
        const int v = verts[2 * g + s], o = verts[2 * g + 1 - s];
        if (!owns(v)) continue;  // the owner of the site gauges it
        for (size_t p = 0; p < inc_[v].size(); ++p) {
          const int w = inc_[v][p].nbr;
          if (w == o) continue;
          const int de = dedge(w, v);
          if (!msg_set_[de]) continue;  // identity default ⇒ √M = M^{-1/2} = 1
          const int chi = bond_[de / 2];
          MsgEigTask t{};
          t.M = msg_[de];
          t.A = (double2*)talloc((size_t)chi * chi * sizeof(double2));
          t.V = (double2*)talloc((size_t)chi * chi * sizeof(double2));
          t.sqrtM = talloc((size_t)chi * chi * esz_);
          t.proj = talloc((size_t)chi * chi * esz_);
          t.chi = chi;
          t.flags = nullptr;
          t.errflags = d_errflags_;
          t.lam = (double*)talloc(sizeof(double) * chi);
          envs.push_back({k, s, (int)p, de, (int)mt.size()});
          mt.push_back(t);
        }
      }
    }
    int* d_flags = (int*)talloc(sizeof(int) * 2 * std::max<size_t>(1, mt.size()));
    for (size_t i = 0; i < mt.size(); ++i) mt[i].flags = d_flags + 2 * i;
    if (!mt.empty()) {
      MsgEigTask* dm = upload(mt);
      if (c64()) msg_prepare_kernel<float><<<(unsigned)mt.size(), 256, 0, stream_>>>(dm);
      else msg_prepare_kernel<double><<<(unsigned)mt.size(), 256, 0, stream_>>>(dm);
      count_launch();
      std::vector<JacobiTask> jt(mt.size());
      std::vector<double*> sv(mt.size());
      for (size_t i = 0; i < mt.size(); ++i) {
        jt[i].A = mt[i].A; jt[i].V = mt[i].V; jt[i].m = jt[i].n = mt[i].chi;
        jt[i].sval = (double*)talloc(sizeof(double) * mt[i].chi);
        jt[i].perm = (int*)talloc(sizeof(int) * mt[i].chi);
      }
      launch_jacobi(jt, 1e-40);  // message eigenvalues are kept down to the absolute sqrt_cutoff: no early null columns
      if (c64()) msg_finish_kernel<float><<<(unsigned)mt.size(), 256, 0, stream_>>>(dm, sqrt_cutoff);
      else msg_finish_kernel<double><<<(unsigned)mt.size(), 256, 0, stream_>>>(dm, sqrt_cutoff);
      count_launch();
      TNQS_CUDA(cudaGetLastError());
    }

    phase_mark("msg");
    // ---- 2. gauge: T̃ = T ×_ext √M (simple_update.jl:43-44) -------------------------------------
    std::vector<Chain> gauge(2 * ng);
    for (int k = 0; k < ng; ++k)
      for (int s = 0; s < 2; ++s) gauge[2 * k + s].v = verts[2 * gate_ids[gpos + k] + s];
    for (auto& en : envs) gauge[2 * en.gate + en.site].steps.push_back({en.pos, mt[en.task].sqrtM});
    run_chains(gauge);

    phase_mark("gauge");
    // ---- 3. Gram of the gauged tensor over its external legs (R†R of the QR at :47-48) ---------
    std::vector<GramTask> gt(2 * ng);
    std::vector<double2*> G(2 * ng);
    std::vector<int> nn(2 * ng), epos(2 * ng);
    std::vector<int> ebond(ng);
    for (int k = 0; k < ng; ++k) {
      const int g = gate_ids[gpos + k];
      const int a = verts[2 * g], b = verts[2 * g + 1];
      const int e = dedge(a, b) / 2;
      ebond[k] = e;
      for (int s = 0; s < 2; ++s) {
        const int v = s == 0 ? a : b;
        epos[2 * k + s] = leg_pos(v, e);
        gt[2 * k + s] = gram_task(v, epos[2 * k + s], phys_[v], gauge[2 * k + s].result, gauge[2 * k + s].result);
        nn[2 * k + s] = gt[2 * k + s].MM;
        G[2 * k + s] = (double2*)talloc((size_t)nn[2 * k + s] * nn[2 * k + s] * sizeof(double2));
      }
    }
    {
      // each rank forms the Gram matrices of the sites it owns; every rank then receives all of them
      // and repeats the (deterministic) O(χ³) algebra, so bond dimensions, singular values and
      // messages stay replicated without a second exchange
      std::vector<GramTask> gto;
      std::vector<double2*> Go;
      std::vector<Bcast> bc;
      for (int i = 0; i < 2 * ng; ++i) {
        const int v = verts[2 * gate_ids[gpos + i / 2] + (i & 1)];
        if (owns(v)) { gto.push_back(gt[i]); Go.push_back(G[i]); }
        if (nranks_ > 1) bc.push_back({G[i], (size_t)nn[i] * nn[i] * sizeof(double2), owner_[v]});
      }
      launch_gram(gto, /*acc_double=*/true, Go, /*transpose=*/false);
      exchange(bc);
    }

    phase_mark("gram");
    // ---- 4. eig(G), θ, SVD(θ), truncation ------------------------------------------------------
    // Multi-GPU: the O(χ³) factorisations of gate k run on rank k mod R only (its "solver"); the kept
    // rank, truncation error, singular values and the two new factors then reach every rank with two
    // all-gathers over packed, rank-major records.
    const int R = nranks_ > 1 ? nranks_ : 1;
    const int per = (ng + R - 1) / R;
    auto solver = [&](int k) { return R > 1 ? k % R : rank_; };
    auto slot = [&](int k) { return R > 1 ? (k % R) * per + k / R : k; };
    std::vector<SuGateTask> st(ng);
    std::vector<int> keep_cap(ng);
    int maxcols = 1;
    size_t xstride = 256;
    for (int k = 0; k < ng; ++k) {
      const int g = gate_ids[gpos + k];
      const int d0 = phys_[verts[2 * g]], d1 = phys_[verts[2 * g + 1]];
      const int chi = bond_[ebond[k]];
      // thin QR of the reference: r_s = min(∏ external dims, d_s·χ_b)  (simple_update.jl:47-48)
      long long full = 1ll << 40;
      for (int s = 0; s < 2; ++s) {
        const int v = verts[2 * g + s];
        const long long ext = site_elems(v) / ((long long)phys_[v] * chi);
        full = std::min(full, std::min<long long>(ext, nn[2 * k + s]) * phys_[v]);
      }
      int cap = (int)full;
      if (ao.maxdim > 0) cap = std::min(cap, std::max(ao.maxdim, 1));
      keep_cap[k] = cap;
      SuGateTask& t = st[k];
      std::memset(&t, 0, sizeof(t));
      t.full = (int)full;
      maxcols = std::max(maxcols, nn[2 * k + 1] * d1);
      const size_t x0 = ((size_t)nn[2 * k] * d0 * cap * esz_ + 255) & ~size_t(255);
      const size_t x1 = ((size_t)nn[2 * k + 1] * d1 * cap * esz_ + 255) & ~size_t(255);
      xstride = std::max(xstride, x0 + x1);
    }
    const size_t rec_bytes = 32 + sizeof(double) * (size_t)maxcols;  // {err, Σkept σ², keep, pad, σ[maxcols]}
    char* d_rec = (char*)talloc(rec_bytes * (size_t)R * per);
    char* d_x = (char*)talloc(xstride * (size_t)R * per);
    std::vector<int> mine;
    for (int k = 0; k < ng; ++k) {
      const int g = gate_ids[gpos + k];
      SuGateTask& t = st[k];
      const int d0 = phys_[verts[2 * g]], d1 = phys_[verts[2 * g + 1]];
      char* rec = d_rec + rec_bytes * (size_t)slot(k);
      t.err = (double*)rec; t.sumsq_kept = (double*)(rec + 8); t.keep = (int*)(rec + 16); t.sigma = (double*)(rec + 32);
      char* xb = d_x + xstride * (size_t)slot(k);
      t.X[0] = xb;
      t.X[1] = xb + (((size_t)nn[2 * k] * d0 * keep_cap[k] * esz_ + 255) & ~size_t(255));
      t.d[0] = d0; t.d[1] = d1; t.chi_b = bond_[ebond[k]];
      if (solver(k) == rank_) mine.push_back(k);
    }
    {
      const int nm = (int)mine.size();
      std::vector<HermTask> ht(2 * nm);
      std::vector<JacobiTask> jg(2 * nm);
      for (int q = 0; q < nm; ++q)
        for (int s = 0; s < 2; ++s) {
          const int i = 2 * mine[q] + s, j = 2 * q + s;
          const int n = nn[i];
          ht[j].G = G[i]; ht[j].n = n;
          ht[j].A = (double2*)talloc((size_t)n * n * sizeof(double2));
          ht[j].V = (double2*)talloc((size_t)n * n * sizeof(double2));
          jg[j].A = ht[j].A; jg[j].V = ht[j].V; jg[j].m = jg[j].n = n;
          jg[j].sval = (double*)talloc(sizeof(double) * n);
          jg[j].perm = (int*)talloc(sizeof(int) * n);
        }
      int maxn_g = 0;
      for (auto& h : ht) maxn_g = std::max(maxn_g, h.n);
      if (nm > 0 && use_chol_ && maxn_g <= 256) {
        // Cholesky-preconditioned eigendecomposition (kernels_small.cuh): L in shared memory (n ≤ 96) or in an
        // L2-resident global scratch matrix (χ = 64 gives n = d·χ = 128), Jacobi on L without V
        const bool globg = maxn_g > 96;
        std::vector<CholTask> ct(2 * nm);
        for (int j = 0; j < 2 * nm; ++j) {
          ct[j].G = ht[j].G; ct[j].A = ht[j].A; ct[j].V = ht[j].V; ct[j].n = ht[j].n;
          ct[j].piv = (int*)talloc(sizeof(int) * ht[j].n);
          ct[j].sval = jg[j].sval;
          ct[j].scratch = globg ? (double2*)talloc((size_t)ht[j].n * ht[j].n * sizeof(double2)) : nullptr;
          jg[j].V = nullptr;
        }
        CholTask* dc = upload(ct);
        const size_t sm = (globg ? 0 : (size_t)maxn_g * maxn_g * sizeof(double2)) + (size_t)maxn_g * (sizeof(double) + sizeof(int));
        chol_prepare_kernel<<<2 * nm, globg ? 1024 : 256, sm, stream_>>>(dc, 1e-15);
        count_launch();
        launch_jacobi(jg, 1e-40);  // the null columns of L are exactly zero
        chol_finish_kernel<<<2 * nm, 256, 0, stream_>>>(dc);
        count_launch();
        TNQS_CUDA(cudaGetLastError());
      } else if (nm > 0) {
        HermTask* dh = upload(ht);
        herm_prepare_kernel<<<2 * nm, 256, 0, stream_>>>(dh);
        count_launch();
        // directions with λ ≤ 64·eps·λmax are dropped by su_theta: columns below 1e-15·‖G‖_F may stop rotating early
        launch_jacobi(jg, 1e-30);
      }
      std::vector<JacobiTask> jt(nm);
      std::vector<SuGateTask> stm(nm);
      for (int q = 0; q < nm; ++q) {
        const int k = mine[q];
        const int g = gate_ids[gpos + k];
        SuGateTask& t = st[k];
        const int d0 = t.d[0], d1 = t.d[1];
        for (int s = 0; s < 2; ++s) {
          t.GA[s] = ht[2 * q + s].A; t.GV[s] = ht[2 * q + s].V;
          t.sq[s] = (double*)talloc(sizeof(double) * nn[2 * k + s]);
          t.isq[s] = (double*)talloc(sizeof(double) * nn[2 * k + s]);
        }
        const int D = d0 * d1;
        const double* gm = mats + mat_off[g];
        for (int i = 0; i < D * D; ++i) { t.gate[i].x = gm[2 * i]; t.gate[i].y = gm[2 * i + 1]; }
        const int rows = nn[2 * k] * d0, cols = nn[2 * k + 1] * d1;
        t.theta = (double2*)talloc((size_t)rows * cols * sizeof(double2));
        t.theta0 = (double2*)talloc((size_t)rows * cols * sizeof(double2));
        double* sval = (double*)talloc(sizeof(double) * cols);
        int* perm = (int*)talloc(sizeof(int) * cols);
        t.sval = sval; t.perm = perm;
        t.Rp = (double2*)talloc((size_t)cols * keep_cap[k] * sizeof(double2));
        jt[q].A = t.theta; jt[q].V = nullptr; jt[q].m = rows; jt[q].n = cols; jt[q].sval = sval; jt[q].perm = perm;
        stm[q] = t;
      }
      SuGateTask* ds = nullptr;
      if (nm > 0) {
        ds = upload(stm);
        // slices per gate of the small dense kernels: enough CTAs for ~4 per SM whatever the batch size
        const int su_slices = std::max(1, std::min(16, (148 * 4 + nm - 1) / nm));
        for (int q = 0; q < nm; ++q)
          if (nn[2 * mine[q]] > SU_MAXN || nn[2 * mine[q] + 1] > SU_MAXN) throw Error(TNQS_EINVAL, "simple update: reduced factor larger than 256 rows");
        su_theta_kernel<<<dim3(nm, su_slices), 256, 0, stream_>>>(ds, 64 * 2.220446049250313e-16);
        count_launch();
        // Preconditioned SVD: K = θ†θ → pivoted Cholesky P·K·Pᵀ = L·L† → Jacobi on L (fast: L is the
        // preconditioned form, and only rank(θ) columns are non-zero) → V_K = Pᵀ·U_L ≈ right singular vectors →
        // B = θ·V_K has nearly orthogonal columns → a short Jacobi polish on B restores full relative accuracy.
        // Directions with σ < ~3e-8·σmax (σ² below the fp64 noise of K) come back as σ = 0, so the path is taken
        // only when they cannot matter: ComplexF32 states (fp32 noise 6e-8) or a relative cutoff ≥ 1e-12 on σ² (a hundred
        // such directions together stay below the cutoff, so neither the kept rank nor truncerr can depend on them).
        int maxcols_t = 0, maxrows_t = 0;
        for (auto& j : jt) { maxcols_t = std::max(maxcols_t, j.n); maxrows_t = std::max(maxrows_t, j.m); }
        const bool tail_irrelevant = c64() || (ao.cutoff >= 1e-12 && ao.use_relative_cutoff && !ao.use_absolute_cutoff);
        if (use_fast_svd_ && tail_irrelevant && maxcols_t <= 256 && maxrows_t <= 256 && maxcols_t >= 8) {
          std::vector<CholTask> ct(nm);
          std::vector<SmallGemmTask> g1(nm), g2(nm);
          std::vector<JacobiTask> jl(nm);
          const bool glob = maxcols_t > 96;
          for (int q = 0; q < nm; ++q) {
            const int k = mine[q];
            const int rows = jt[q].m, cols = jt[q].n;
            double2* K = (double2*)talloc((size_t)cols * cols * sizeof(double2));
            double2* L = (double2*)talloc((size_t)cols * cols * sizeof(double2));
            double2* Vk = (double2*)talloc((size_t)cols * cols * sizeof(double2));
            ct[q].G = K; ct[q].A = L; ct[q].V = Vk; ct[q].n = cols;
            ct[q].piv = (int*)talloc(sizeof(int) * cols);
            double* sv1 = (double*)talloc(sizeof(double) * cols);
            ct[q].sval = sv1;
            ct[q].scratch = glob ? (double2*)talloc((size_t)cols * cols * sizeof(double2)) : nullptr;
            g1[q].A = st[k].theta; g1[q].B = nullptr; g1[q].out = K; g1[q].m = rows; g1[q].n = cols;
            g2[q].A = st[k].theta0; g2[q].B = Vk; g2[q].out = st[k].theta; g2[q].m = rows; g2[q].n = cols;
            jl[q].A = L; jl[q].V = nullptr; jl[q].m = cols; jl[q].n = cols; jl[q].sval = sv1;
            jl[q].perm = (int*)talloc(sizeof(int) * cols);
          }
          SmallGemmTask* d1 = upload(g1);
          SmallGemmTask* d2 = upload(g2);
          CholTask* dc = upload(ct);
          // 4×4 output tiles per thread: (n/4)² tiles of a matrix keep at most (n/4)²/256 CTAs busy
          const int gemm_slices = std::max(1, std::min(su_slices, ((maxcols_t + 3) / 4) * ((maxrows_t + 3) / 4) / 256));
          colgram_kernel<<<dim3(nm, gemm_slices), 256, 0, stream_>>>(d1);
          const size_t sm = (glob ? 0 : (size_t)maxcols_t * maxcols_t * sizeof(double2)) + (size_t)maxcols_t * (sizeof(double) + sizeof(int));
          chol_prepare_kernel<<<nm, glob ? 1024 : 256, sm, stream_>>>(dc, 1e-15);
          count_launch(2);
          launch_jacobi(jl, 1e-40);
          chol_finish_kernel<<<nm, 256, 0, stream_>>>(dc);
          colapply_kernel<<<dim3(nm, gemm_slices), 256, 0, stream_>>>(d2);
          count_launch(2);
          TNQS_CUDA(cudaGetLastError());
        }
        // singular values below 1e-13·‖θ‖_F never survive the truncation (σ² < 1e-26 of the total weight)
        launch_jacobi(jt, 1e-26);
        su_truncate_kernel<<<(nm + 63) / 64, 64, 0, stream_>>>(ds, nm, ao.maxdim, ao.mindim, ao.cutoff, ao.use_absolute_cutoff, ao.use_relative_cutoff);
        count_launch();
        TNQS_CUDA(cudaGetLastError());
      }
      if (R > 1) {
        NcclApi& api = NcclApi::get();
        api.check(api.AllGather(d_rec + rec_bytes * (size_t)per * rank_, d_rec, rec_bytes * (size_t)per, kNcclChar, comm_->comm, stream_),
                  "ncclAllGather(records)");
        stats_.kernel_launches += 1;
      }
    phase_mark("factor-enqueue");
      // the host needs the kept ranks to size the new tensors: the one sync of the batch
      std::vector<int> keep(ng), flags(2 * std::max<size_t>(1, mt.size()));
      std::vector<double> err(ng);
      std::vector<char> h_rec(rec_bytes * (size_t)R * per);
      // error words (Jacobi non-convergence, DomainError of a message square root): summed over the ranks BEFORE anyone
      // reads them, so every rank takes the same decision at the same point and nobody is left inside a collective
      double h_errflags[2] = {0.0, 0.0};
      allreduce_sum(d_errflags_, 2);
      TNQS_CUDA(cudaMemcpyAsync(h_errflags, d_errflags_, sizeof(h_errflags), cudaMemcpyDeviceToHost, stream_));
      TNQS_CUDA(cudaMemcpyAsync(h_rec.data(), d_rec, h_rec.size(), cudaMemcpyDeviceToHost, stream_));
      if (!mt.empty())
        TNQS_CUDA(cudaMemcpyAsync(flags.data(), d_flags, sizeof(int) * 2 * mt.size(), cudaMemcpyDeviceToHost, stream_));
      { WaitScope wsc(&stats_.sync_ms); TNQS_CUDA(cudaStreamSynchronize(stream_)); }
      if (h_errflags[0] != 0.0 || h_errflags[1] != 0.0) {
        TNQS_CUDA(cudaMemsetAsync(d_errflags_, 0, 2 * sizeof(double), stream_));
        free_temps();
        if (h_errflags[1] != 0.0) {
          std::string where;
          for (size_t i = 0; i < mt.size(); ++i)
            if (flags[2 * i + 1]) { where = " (message into vertex " + std::to_string(verts[2 * gate_ids[gpos + envs[i].gate] + envs[i].site]) + ")"; break; }
          throw Error(TNQS_EDOMAIN, "DomainError: sqrt of a negative message eigenvalue" + where);
        }
        throw Error(TNQS_ECUDA, "simple update: a one-sided Jacobi factorisation did not converge within 40 sweeps");
      }
      for (int k = 0; k < ng; ++k) {
        const char* rec = h_rec.data() + rec_bytes * (size_t)slot(k);
        std::memcpy(&err[k], rec, sizeof(double));
        std::memcpy(&keep[k], rec + 16, sizeof(int));
        if (keep[k] < 1 || keep[k] > keep_cap[k]) {
          free_temps();
          throw Error(TNQS_ECUDA, "simple update: invalid kept rank returned by the truncation kernel");
        }
      }
      if (nm > 0) {
        const int su_slices = std::max(1, std::min(16, (148 * 4 + nm - 1) / nm));
        const dim3 fg(nm, su_slices);
        if (c64()) { su_factors_kernel<float, 0><<<fg, 256, 0, stream_>>>(ds); su_factors_kernel<float, 1><<<fg, 256, 0, stream_>>>(ds); }
        else { su_factors_kernel<double, 0><<<fg, 256, 0, stream_>>>(ds); su_factors_kernel<double, 1><<<fg, 256, 0, stream_>>>(ds); }
        count_launch(2);
        TNQS_CUDA(cudaGetLastError());
      }
      if (R > 1) {
        NcclApi& api = NcclApi::get();
        api.check(api.AllGather(d_x + xstride * (size_t)per * rank_, d_x, xstride * (size_t)per, kNcclChar, comm_->comm, stream_),
                  "ncclAllGather(factors)");
        stats_.kernel_launches += 1;
      }

    phase_mark("sync+factors");
      // ---- 5. un-gauge with the projector and contract with the new factor (:62-64) -----------
      std::vector<Chain> proj(2 * ng);
      for (int k = 0; k < ng; ++k)
        for (int s = 0; s < 2; ++s) proj[2 * k + s].v = verts[2 * gate_ids[gpos + k] + s];
      for (auto& en : envs)
        if (!flags[2 * en.task]) proj[2 * en.gate + en.site].steps.push_back({en.pos, mt[en.task].proj});
      run_chains(proj);
      std::vector<ModeTask> fin;
      std::vector<void*> newbuf(2 * ng, nullptr);
      std::vector<size_t> oldbytes(2 * ng, 0);
      std::set<size_t> newsizes;
      const size_t site_pool_cap = (size_t)64 << 30;  // a third of a B200's HBM
      for (int k = 0; k < ng; ++k) {
        const int g = gate_ids[gpos + k];
        for (int s = 0; s < 2; ++s) {
          const int v = verts[2 * g + s];
          if (!owns(v)) continue;
          const int d = phys_[v];
          const int pos = epos[2 * k + s];
          fin.emplace_back();
          ModeTask& t = fin.back();
          t = ModeTask{};
          int chi;
          leg_view(v, pos, &t.outer, &chi, &t.inner);
          t.outer /= (unsigned)d;
          t.chi_in = chi; t.chi_out = keep[k];
          t.KK = d * chi; t.MM = d * keep[k];
          t.CC = t.outer * t.inner;
          t.ips = (long long)t.outer * chi * t.inner;
          t.ops = (long long)t.outer * keep[k] * t.inner;
          t.in = proj[2 * k + s].result;
          oldbytes[2 * k + s] = (size_t)site_elems(v) * esz_;  // bond_ still holds the old dimensions
          newbuf[2 * k + s] = site_alloc((size_t)t.ops * d * esz_);
          newsizes.insert((size_t)t.ops * d * esz_);
          t.out = newbuf[2 * k + s];
          t.mat = st[k].X[s];
        }
      }
      launch_mode(fin);
    phase_mark("proj+final");
      // ---- 6. commit: tensors, bond dimension, messages (apply_gates.jl:126-140) ----------------
      std::vector<int> touched;
      std::vector<DiagTask> dt;
      for (int k = 0; k < ng; ++k) {
        const int g = gate_ids[gpos + k];
        const int e = ebond[k];
        bond_[e] = keep[k];
        for (int s = 0; s < 2; ++s) {
          const int v = verts[2 * g + s];
          sshape_[v][epos[2 * k + s]] = keep[k];
          if (!owns(v)) continue;
          // keep the old buffer for the next batch only if this batch asked for the same size (saturated bonds)
          if (newsizes.count(oldbytes[2 * k + s]) && site_pool_bytes_ + oldbytes[2 * k + s] <= site_pool_cap) site_release(site_[v], oldbytes[2 * k + s]);
          else dfree(site_[v]);
          site_[v] = newbuf[2 * k + s];
          touched.push_back(v);
        }
        for (int de = 2 * e; de < 2 * e + 2; ++de) {
          if (msg_[de]) dfree(msg_[de]);
          msg_[de] = dalloc((size_t)keep[k] * keep[k] * esz_);
          msg_dim_[de] = keep[k];
          msg_set_[de] = 1;
          DiagTask d{};
          d.out = msg_[de]; d.chi = keep[k]; d.diag = st[k].sigma;
          d.scale_sumsq = normalize ? st[k].sumsq_kept : nullptr;
          dt.push_back(d);
        }
        errs[g] = err[k];
      }
      if (normalize) normalize_sites(touched);
      DiagTask* dd = upload(dt);
      int maxchi = 1;
      for (auto& d : dt) maxchi = std::max(maxchi, d.chi);
      const int nb = std::max(1, std::min(64, (maxchi * maxchi + 255) / 256));
      if (c64()) diag_fill_kernel<float><<<dim3(nb, (unsigned)dt.size()), 256, 0, stream_>>>(dd);
      else diag_fill_kernel<double><<<dim3(nb, (unsigned)dt.size()), 256, 0, stream_>>>(dd);
      count_launch();
      TNQS_CUDA(cudaGetLastError());
    }
    stats_.two_site_gates += ng;
    free_temps();
    phase_mark("commit");
    if (SlowLog::on()) {
      double tot = 0;
      for (auto& ph : phase_log) tot += ph.second;
      static const double thr = std::getenv("TNQS_PHASELOG") ? 0.0 : 60.0;  // TNQS_PHASELOG=1: every batch
      if (tot > thr) {
        std::fprintf(stderr, "[tnqs slow] two-site batch of %d gates: host", ng);
        for (auto& ph : phase_log) std::fprintf(stderr, " %s %.1f", ph.first, ph.second);
        std::fprintf(stderr, " ms\n");
      }
    }
    gpos = gend;
  }
}

void Engine::apply_gates(int ngates, const int32_t* nverts, const int32_t* verts, const double* mats,
                         const tnqs_apply_opts* aop, const tnqs_bp_opts* bo, int update_cache,
                         double* errs, tnqs_bp_report* reports, int max_reports, int* n_reports) {
  WallScope ws(&stats_.wall_ms, &wall_depth_);
  TNQS_CUDA(cudaSetDevice(device_));
  check_shapes();
  tnqs_apply_opts ao{0, 1, -1.0, 1, -1.0, 0, 1, 0, 0};
  if (aop) ao = *aop;
  if (ao.svd_alg < 0 || ao.svd_alg > 2) throw Error(TNQS_EINVAL, "unknown SVD algorithm");
  if (ao.mindim < 1) ao.mindim = 1;
  // validate everything before touching the state (apply_gates.jl:109-120)
  std::vector<size_t> mat_off(ngates);
  size_t off = 0;
  for (int i = 0; i < ngates; ++i) {
    if (nverts[i] < 1 || nverts[i] > 2)
      throw Error(TNQS_ENSITES, "apply_gate!: only one- and two-site gates are supported; received a gate acting on " +
                                    std::to_string(nverts[i]) + " vertices.");
    const int a = verts[2 * i];
    if (a < 0 || a >= nv_) throw Error(TNQS_EINVAL, "gate vertex out of range");
    int D = phys_[a];
    if (nverts[i] == 2) {
      const int b = verts[2 * i + 1];
      if (b < 0 || b >= nv_) throw Error(TNQS_EINVAL, "gate vertex out of range");
      bool adj = false;
      for (auto& l : inc_[a]) adj |= (l.nbr == b);
      if (!adj)
        throw Error(TNQS_ENOTADJ, "apply_gate!: cannot apply a two-site gate on the non-adjacent vertices " +
                                      std::to_string(a) + " and " + std::to_string(b) + ".");
      D *= phys_[b];
    }
    mat_off[i] = off;
    off += 2 * (size_t)D * D;
  }
  // The batched one-sided Jacobi takes matrices of up to 512 rows; θ of a two-site gate has d·min(∏ external dims, d·χ) rows
  // (simple_update.jl:47-52), χ = the shared bond before the gate — at most max(current bond, maxdim) for every gate of this
  // call when maxdim is given.  Refuse the call before anything is touched rather than in the middle of it.
  for (int i = 0; i < ngates; ++i) {
    if (nverts[i] != 2) continue;
    const int a = verts[2 * i], b = verts[2 * i + 1];
    const int e = dedge(a, b) / 2;
    const long long chi = ao.maxdim > 0 ? std::max(bond_[e], ao.maxdim) : bond_[e];
    for (int s = 0; s < 2; ++s) {
      const int v = s ? b : a;
      const long long ext = site_elems(v) / ((long long)phys_[v] * bond_[e]);
      const long long rows = std::min<long long>(ext, (long long)phys_[v] * chi) * phys_[v];
      if (rows > 512)
        throw Error(TNQS_EINVAL, "apply_gates: the two-site factorisation of the gate on vertices " + std::to_string(a) + " and " +
                                     std::to_string(b) + " would have " + std::to_string(rows) +
                                     " rows; the batched Jacobi SVD of this build takes at most 512 (d^2·chi <= 512, i.e. chi <= 128 for qubits)");
    }
  }
  int nrep = 0;
  EventPair ev;
  const cudaEvent_t t0 = ev.a, t1 = ev.b;
  TNQS_CUDA(cudaEventRecord(t0, stream_));
  const double bp_before = stats_.bp_ms;

  // Between two BP refreshes, order gates by per-vertex dependency depth; every depth is one batch
  // of vertex-disjoint two-site gates plus one batch of (fused) one-site gates.
  std::vector<int> seg;
  auto flush = [&]() {
    if (seg.empty()) return;
    struct Stage { std::vector<int> two; std::vector<std::pair<int, std::vector<cplx>>> one; };
    std::vector<Stage> stages;
    std::vector<int> last_stage(nv_, -1), open_one(nv_, -1);  // open_one: index into stage.one
    for (int g : seg) {
      if (nverts[g] == 1) {
        const int v = verts[2 * g], d = phys_[v];
        std::vector<cplx> U(d * d);
        for (int k = 0; k < d * d; ++k) U[k] = cplx(mats[mat_off[g] + 2 * k], mats[mat_off[g] + 2 * k + 1]);
        if (open_one[v] >= 0) {  // fuse with the preceding one-site gate on this vertex
          auto& prev = stages[last_stage[v]].one[open_one[v]].second;
          std::vector<cplx> W(d * d);
          for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
              cplx s = 0;
              for (int k = 0; k < d; ++k) s += U[i * d + k] * prev[k * d + j];
              W[i * d + j] = s;
            }
          prev = W;
        } else {
          const int sidx = last_stage[v] + 1;
          if ((int)stages.size() <= sidx) stages.resize(sidx + 1);
          stages[sidx].one.push_back({v, U});
          open_one[v] = (int)stages[sidx].one.size() - 1;
          last_stage[v] = sidx;
        }
      } else {
        const int a = verts[2 * g], b = verts[2 * g + 1];
        const int sidx = std::max(last_stage[a], last_stage[b]) + 1;
        if ((int)stages.size() <= sidx) stages.resize(sidx + 1);
        stages[sidx].two.push_back(g);
        last_stage[a] = last_stage[b] = sidx;
        open_one[a] = open_one[b] = -1;
      }
    }
    for (auto& st : stages) {
      apply_two_site_batch(st.two, verts, mats, mat_off, ao, errs);
      apply_one_site_batch(st.one, ao.normalize_tensors != 0);
    }
    seg.clear();
  };

  std::vector<char> affected(nv_, 0);
  for (int i = 0; i < ngates; ++i) {
    errs[i] = 0.0;
    bool need = false;
    if (nverts[i] >= 2) need = affected[verts[2 * i]] || affected[verts[2 * i + 1]];
    if (update_cache && need) {  // apply_gates.jl:68-83
      flush();
      const tnqs_bp_report r = bp_update(bo);
      if (reports && nrep < max_reports) reports[nrep] = r;
      ++nrep;
      std::fill(affected.begin(), affected.end(), 0);
    }
    seg.push_back(i);
    for (int k = 0; k < nverts[i]; ++k) affected[verts[2 * i + k]] = 1;
  }
  flush();
  if (update_cache) {  // apply_gates.jl:93-95
    const tnqs_bp_report r = bp_update(bo);
    if (reports && nrep < max_reports) reports[nrep] = r;
    ++nrep;
  }
  if (n_reports) *n_reports = nrep;
  TNQS_CUDA(cudaEventRecord(t1, stream_));
  { WaitScope wsc(&stats_.sync_ms); TNQS_CUDA(cudaEventSynchronize(t1)); }
  float ms = 0;
  cudaEventElapsedTime(&ms, t0, t1);
  stats_.su_ms += ms - (stats_.bp_ms - bp_before);
  release_slabs();  // stream drained by the event wait above
}

// ------------------------------------------------------------------------------------------------
// expectation values (expect.jl:59-82)
// ------------------------------------------------------------------------------------------------
// un-normalised single-site density matrices ρ_v[s][s'] = ⟨T_v| msgs |T_v⟩ with the physical legs open
// (the common part of expect.jl:59-82 and vertex_scalar, abstractbeliefpropagationcache.jl:22-28)
void Engine::local_rdms(int n, const int32_t* verts, std::vector<cplx>& rho, std::vector<size_t>& offs) {
  TNQS_CUDA(cudaSetDevice(device_));
  check_shapes();
  std::vector<Chain> chains(n);
  for (int i = 0; i < n; ++i) {
    const int v = verts[i];
    if (v < 0 || v >= nv_) throw Error(TNQS_EINVAL, "vertex out of range");
    chains[i].v = v;
    for (size_t p = 0; p < inc_[v].size(); ++p) {
      const int de = dedge(inc_[v][p].nbr, v);
      if (!msg_set_[de] || !msg_[de]) continue;
      chains[i].steps.push_back({(int)p, msg_[de]});
    }
  }
  run_chains(chains);
  std::vector<double2*> outs(n);
  size_t tot = 0;
  offs.assign(n, 0);
  for (int i = 0; i < n; ++i) { offs[i] = tot; tot += (size_t)phys_[verts[i]] * phys_[verts[i]]; }
  double2* d_rho = (double2*)talloc(tot * sizeof(double2));
  TNQS_CUDA(cudaMemsetAsync(d_rho, 0, tot * sizeof(double2), stream_));
  {
    std::vector<GramTask> gto;
    std::vector<double2*> oo;
    for (int i = 0; i < n; ++i) {
      outs[i] = d_rho + offs[i];
      if (!owns(verts[i])) continue;  // the owner contracts; the others contribute zeros to the sum
      gto.push_back(gram_task(verts[i], -1, 1, site_[verts[i]], chains[i].result));
      oo.push_back(outs[i]);
    }
    launch_gram(gto, true, oo, /*transpose=*/true);  // buffer[s*d+s'] = ρ[s][s']
    allreduce_sum(reinterpret_cast<double*>(d_rho), 2 * tot);
  }
  rho.assign(tot, cplx(0, 0));
  TNQS_CUDA(cudaMemcpyAsync(rho.data(), d_rho, tot * sizeof(double2), cudaMemcpyDeviceToHost, stream_));
  TNQS_CUDA(cudaStreamSynchronize(stream_));
  free_temps();
  release_slabs();
}

void Engine::expect_local(int nobs, const int32_t* verts, const double* ops, double* out) {
  if (nobs <= 0) return;
  std::vector<cplx> rho;
  std::vector<size_t> offs;
  local_rdms(nobs, verts, rho, offs);
  size_t opoff = 0;
  for (int i = 0; i < nobs; ++i) {
    const int d = phys_[verts[i]];
    const cplx* r = rho.data() + offs[i];
    cplx num = 0, den = 0;
    for (int s = 0; s < d; ++s) {
      den += r[s * d + s];
      for (int sp = 0; sp < d; ++sp) {
        const cplx O(ops[opoff + 2 * (sp * d + s)], ops[opoff + 2 * (sp * d + s) + 1]);
        num += O * r[s * d + sp];
      }
    }
    opoff += 2 * (size_t)d * d;
    const cplx val = num / den;
    out[2 * i] = val.real(); out[2 * i + 1] = val.imag();
  }
}

// One site of a region contraction (expect.jl:59-82 / rdm.jl:52-73 for a region that is a path): the ket tensor of vertex v
// absorbs, on every bond leg except `open_nbr`, either a caller-supplied matrix (the partial contraction arriving from the
// previous site of the path) or the BP message of that leg; an optional operator acts on its physical index; the result is
// contracted with conj(T_v) over everything except the open indices.  out[ket][bra] row-major, complex128:
//   open_phys = 1, open_nbr = w : E[(s,b),(s',b')], n = d·χ_b     open_phys = 0, open_nbr = w : M[b,b'], n = χ_b
//   open_phys = 1, open_nbr < 0 : ρ[s,s'], n = d
void Engine::site_contract(int v, int n_custom, const int32_t* custom_nbrs, const double* custom_mats, int open_nbr, int open_phys,
                           const double* op, double* out, int64_t cap, int* out_n) {
  TNQS_CUDA(cudaSetDevice(device_));
  check_shapes();
  if (v < 0 || v >= nv_) throw Error(TNQS_EINVAL, "vertex out of range");
  if (open_nbr < 0 && !open_phys) throw Error(TNQS_EINVAL, "tnqs_site_contract: nothing left open");
  const int d = phys_[v];
  int open_pos = -1;
  if (open_nbr >= 0) open_pos = leg_pos(v, dedge(open_nbr, v) / 2);
  const int chi_open = open_pos >= 0 ? bond_[inc_[v][open_pos].edge] : 1;
  const int n = (open_phys ? d : 1) * chi_open;
  *out_n = n;
  if (cap < (int64_t)n * n) throw Error(TNQS_ECAPACITY, "tnqs_site_contract: output buffer too small");
  // device copies of the caller's matrices in the tensor's scalar type
  auto to_device = [&](const double* src, size_t cnt) -> void* {
    void* dm = talloc(cnt * esz_);
    if (c64()) {
      std::vector<float2> h(cnt);
      for (size_t k = 0; k < cnt; ++k) { h[k].x = (float)src[2 * k]; h[k].y = (float)src[2 * k + 1]; }
      TNQS_CUDA(cudaMemcpyAsync(dm, h.data(), cnt * sizeof(float2), cudaMemcpyHostToDevice, stream_));
      TNQS_CUDA(cudaStreamSynchronize(stream_));  // h is a pageable temporary
    } else {
      TNQS_CUDA(cudaMemcpyAsync(dm, src, cnt * sizeof(double2), cudaMemcpyHostToDevice, stream_));
      TNQS_CUDA(cudaStreamSynchronize(stream_));
    }
    return dm;
  };
  std::vector<Chain> chains(1);
  chains[0].v = v;
  std::vector<char> custom_seen(inc_[v].size(), 0);
  size_t moff = 0;
  for (int i = 0; i < n_custom; ++i) {
    const int de = dedge(custom_nbrs[i], v);
    const int p = leg_pos(v, de / 2);
    if (p == open_pos) throw Error(TNQS_EINVAL, "tnqs_site_contract: a matrix was supplied for the open leg");
    if (custom_seen[p]) throw Error(TNQS_EINVAL, "tnqs_site_contract: two matrices for the same leg");
    custom_seen[p] = 1;
    const size_t cnt = (size_t)bond_[de / 2] * bond_[de / 2];
    chains[0].steps.push_back({p, to_device(custom_mats + moff, cnt)});
    moff += 2 * cnt;
  }
  for (size_t p = 0; p < inc_[v].size(); ++p) {
    if ((int)p == open_pos || custom_seen[p]) continue;
    const int de = dedge(inc_[v][p].nbr, v);
    if (!msg_set_[de] || !msg_[de]) continue;  // identity default
    chains[0].steps.push_back({(int)p, msg_[de]});
  }
  if (op) {  // ket ← O·ket: the mode product convention is Out[j'] = Σ_j Mat[j][j']·A[j], so Mat = Oᵀ
    std::vector<double> ot(2 * (size_t)d * d);
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) { ot[2 * (a * d + b)] = op[2 * (b * d + a)]; ot[2 * (a * d + b) + 1] = op[2 * (b * d + a) + 1]; }
    chains[0].steps.push_back({-1, to_device(ot.data(), (size_t)d * d)});
  }
  run_chains(chains);
  double2* d_out = (double2*)talloc((size_t)n * n * sizeof(double2));
  TNQS_CUDA(cudaMemsetAsync(d_out, 0, (size_t)n * n * sizeof(double2), stream_));
  if (owns(v)) {
    std::vector<GramTask> gt(1);
    std::vector<double2*> oo(1, d_out);
    gt[0] = gram_task(v, open_pos, (open_phys && open_pos >= 0) ? d : 1, site_[v], chains[0].result);
    launch_gram(gt, /*acc_double=*/true, oo, /*transpose=*/true);
  }
  allreduce_sum(reinterpret_cast<double*>(d_out), 2 * (size_t)n * n);
  TNQS_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * n * sizeof(double2), cudaMemcpyDeviceToHost, stream_));
  TNQS_CUDA(cudaStreamSynchronize(stream_));
  free_temps();
  release_slabs();
}

// vertex_scalar(bpc, v) (abstractbeliefpropagationcache.jl:22-28): ⟨T_v| incoming messages |T_v⟩ = tr ρ_v
void Engine::vertex_scalars(int n, const int32_t* verts, double* out) {
  if (n <= 0) return;
  std::vector<cplx> rho;
  std::vector<size_t> offs;
  local_rdms(n, verts, rho, offs);
  for (int i = 0; i < n; ++i) {
    const int d = phys_[verts[i]];
    cplx tr = 0;
    for (int s = 0; s < d; ++s) tr += rho[offs[i] + (size_t)s * d + s];
    out[2 * i] = tr.real(); out[2 * i + 1] = tr.imag();
  }
}

// tn[v] ← f_v · tn[v] (rescale_vertices!, beliefpropagationcache.jl:82-101): the one-site kernel with U = f·1
void Engine::scale_sites(int n, const int32_t* verts, const double* factors) {
  if (n <= 0) return;
  TNQS_CUDA(cudaSetDevice(device_));
  check_shapes();
  std::vector<std::pair<int, std::vector<cplx>>> g;
  std::vector<char> seen(nv_, 0);
  for (int i = 0; i < n; ++i) {
    const int v = verts[i];
    if (v < 0 || v >= nv_) throw Error(TNQS_EINVAL, "vertex out of range");
    if (seen[v]) throw Error(TNQS_EINVAL, "tnqs_scale_sites: a vertex may appear only once");
    seen[v] = 1;
    const int d = phys_[v];
    std::vector<cplx> U((size_t)d * d, cplx(0, 0));
    for (int s = 0; s < d; ++s) U[(size_t)s * d + s] = cplx(factors[2 * i], factors[2 * i + 1]);
    g.push_back({v, U});
  }
  apply_one_site_batch(g, /*normalize=*/false);
  TNQS_CUDA(cudaStreamSynchronize(stream_));
}

// random_tensornetworkstate (tensornetworkstate.jl:93-103) generated in place on the device: iid 𝒩(0,1) + i𝒩(0,1) entries
// from a counter-based generator keyed by (seed, vertex), optionally scaled to unit Frobenius norm per tensor.  Messages
// fall back to their identity default.  Sharded caches fill only the tensors they own (same values on any rank count).
void Engine::randomize_sites(unsigned long long seed, int normalize) {
  TNQS_CUDA(cudaSetDevice(device_));
  check_shapes();
  delete_messages();
  std::vector<RandTask> t;
  std::vector<int> vs;
  long long maxn = 0;
  for (int v = 0; v < nv_; ++v) {
    if (!owns(v) || !site_[v]) continue;
    RandTask r{site_[v], site_elems(v), seed * 0x2545F4914F6CDD1Dull + (unsigned long long)v * 0x9E3779B97F4A7C15ull};
    t.push_back(r); vs.push_back(v);
    maxn = std::max(maxn, r.n);
  }
  if (!t.empty()) {
    RandTask* d = upload(t);
    const int nb = (int)std::max<long long>(1, std::min<long long>(1024, (maxn + 1023) / 1024));
    for (int off = 0; off < (int)t.size(); off += 65535) {
      const int cnt = std::min(65535, (int)t.size() - off);
      if (c64()) randn_kernel<float><<<dim3(nb, cnt), 256, 0, stream_>>>(d + off);
      else randn_kernel<double><<<dim3(nb, cnt), 256, 0, stream_>>>(d + off);
      count_launch();
    }
    TNQS_CUDA(cudaGetLastError());
    if (normalize) normalize_sites(vs);
  }
  TNQS_CUDA(cudaStreamSynchronize(stream_));
  free_temps();
}

// tn[v] ← tn[v] ×_{leg(v,nbr)} M for a list of (v, nbr, M): M is χ×χ row-major [in][out] complex128.  Bond
// dimensions do not change.  Used by symmetric_gauge (symmetric_gauge.jl:1-56) and other host-driven regauging.
void Engine::apply_leg_matrices(int n, const int32_t* verts, const int32_t* nbrs, const double* mats) {
  if (n <= 0) return;
  TNQS_CUDA(cudaSetDevice(device_));
  check_shapes();
  std::map<int, int> chain_of;
  std::vector<Chain> chains;
  std::set<std::pair<int, int>> seen;
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const int v = verts[i], w = nbrs[i];
    const int de = dedge(w, v);  // validates adjacency
    if (!seen.insert({v, w}).second) throw Error(TNQS_EINVAL, "tnqs_apply_leg_matrices: a (vertex, neighbour) pair may appear only once");
    const int chi = bond_[de / 2];
    const size_t cnt = (size_t)chi * chi;
    void* dm = talloc(cnt * esz_);
    if (c64()) {
      std::vector<float2> h(cnt);
      for (size_t k = 0; k < cnt; ++k) { h[k].x = (float)mats[off + 2 * k]; h[k].y = (float)mats[off + 2 * k + 1]; }
      TNQS_CUDA(cudaMemcpyAsync(dm, h.data(), cnt * sizeof(float2), cudaMemcpyHostToDevice, stream_));
      TNQS_CUDA(cudaStreamSynchronize(stream_));  // h is a pageable temporary
    } else {
      TNQS_CUDA(cudaMemcpyAsync(dm, mats + off, cnt * sizeof(double2), cudaMemcpyHostToDevice, stream_));
      TNQS_CUDA(cudaStreamSynchronize(stream_));
    }
    off += 2 * cnt;
    auto f = chain_of.find(v);
    if (f == chain_of.end()) { f = chain_of.emplace(v, (int)chains.size()).first; chains.emplace_back(); chains.back().v = v; }
    chains[f->second].steps.push_back({leg_pos(v, de / 2), dm});
  }
  run_chains(chains);
  for (auto& c : chains) {
    if (!owns(c.v) || c.steps.empty() || !c.result) continue;
    TNQS_CUDA(cudaMemcpyAsync(site_[c.v], c.result, (size_t)site_elems(c.v) * esz_, cudaMemcpyDeviceToDevice, stream_));
  }
  TNQS_CUDA(cudaStreamSynchronize(stream_));
  free_temps();
  release_slabs();
}

void Engine::expect_two_site(int nobs, const int32_t* verts, const double* ops, double* out) {
  TNQS_CUDA(cudaSetDevice(device_));
  check_shapes();
  if (nobs <= 0) return;
  std::vector<Chain> chains(2 * nobs);
  std::vector<int> pos(2 * nobs);
  for (int i = 0; i < nobs; ++i) {
    const int a = verts[2 * i], b = verts[2 * i + 1];
    const int e = dedge(a, b) / 2;
    for (int s = 0; s < 2; ++s) {
      const int v = s == 0 ? a : b, o = s == 0 ? b : a;
      chains[2 * i + s].v = v;
      pos[2 * i + s] = leg_pos(v, e);
      for (size_t p = 0; p < inc_[v].size(); ++p) {
        if (inc_[v][p].nbr == o) continue;
        const int de = dedge(inc_[v][p].nbr, v);
        if (!msg_set_[de] || !msg_[de]) continue;
        chains[2 * i + s].steps.push_back({(int)p, msg_[de]});
      }
    }
  }
  run_chains(chains);
  std::vector<GramTask> gt(2 * nobs);
  std::vector<double2*> outs(2 * nobs);
  std::vector<size_t> offs(2 * nobs);
  size_t tot = 0;
  for (int i = 0; i < 2 * nobs; ++i) {
    const int v = chains[i].v;
    gt[i] = gram_task(v, pos[i], phys_[v], site_[v], chains[i].result);
    offs[i] = tot;
    tot += (size_t)gt[i].MM * gt[i].MM;
  }
  double2* d_e = (double2*)talloc(tot * sizeof(double2));
  TNQS_CUDA(cudaMemsetAsync(d_e, 0, tot * sizeof(double2), stream_));
  {
    std::vector<GramTask> gto;
    std::vector<double2*> oo;
    for (int i = 0; i < 2 * nobs; ++i) {
      outs[i] = d_e + offs[i];
      if (!owns(chains[i].v)) continue;
      gto.push_back(gt[i]);
      oo.push_back(outs[i]);
    }
    launch_gram(gto, true, oo, /*transpose=*/false);  // buffer[(s'b')*n + (s b)] = E[s,b,s',b']
    allreduce_sum(reinterpret_cast<double*>(d_e), 2 * tot);
  }
  std::vector<cplx> E(tot);
  TNQS_CUDA(cudaMemcpyAsync(E.data(), d_e, tot * sizeof(double2), cudaMemcpyDeviceToHost, stream_));
  TNQS_CUDA(cudaStreamSynchronize(stream_));
  free_temps();
  release_slabs();
  size_t opoff = 0;
  for (int i = 0; i < nobs; ++i) {
    const int a = verts[2 * i], b = verts[2 * i + 1];
    const int d0 = phys_[a], d1 = phys_[b];
    const int chi = bond_[dedge(a, b) / 2];
    const int n0 = d0 * chi, n1 = d1 * chi;
    const cplx* E0 = E.data() + offs[2 * i];
    const cplx* E1 = E.data() + offs[2 * i + 1];
    const double* O0 = ops + opoff;
    const double* O1 = ops + opoff + 2 * (size_t)d0 * d0;
    opoff += 2 * ((size_t)d0 * d0 + (size_t)d1 * d1);
    cplx num = 0, den = 0;
    for (int s0 = 0; s0 < d0; ++s0) for (int t0 = 0; t0 < d0; ++t0)
      for (int s1 = 0; s1 < d1; ++s1) for (int t1 = 0; t1 < d1; ++t1) {
        cplx rho = 0;  // ρ[s0,s1,t0,t1] = Σ_{b,c} E0[s0,b,t0,c] E1[s1,b,t1,c]
        for (int bb = 0; bb < chi; ++bb) for (int cc = 0; cc < chi; ++cc)
          rho += E0[(size_t)(t0 * chi + cc) * n0 + (s0 * chi + bb)] * E1[(size_t)(t1 * chi + cc) * n1 + (s1 * chi + bb)];
        const cplx o0(O0[2 * (t0 * d0 + s0)], O0[2 * (t0 * d0 + s0) + 1]);
        const cplx o1(O1[2 * (t1 * d1 + s1)], O1[2 * (t1 * d1 + s1) + 1]);
        num += o0 * o1 * rho;
        if (s0 == t0 && s1 == t1) den += rho;
      }
    const cplx val = num / den;
    out[2 * i] = val.real(); out[2 * i + 1] = val.imag();
  }
}

}  // namespace tnqs
