// pawpyseed_b200 engine: device-resident wavefunction state + the C ABI of include/pawpyseed_b200.h.
// There is deliberately no CPU compute path in this file: every overlap / projection / FFT is a
// kernel launch, and entry points fail with an error message when no sm_100 device is usable.
#include <cuda_runtime.h>
#include <cufft.h>
#include <omp.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pawpyseed_b200.h"
#include "host_paw.h"
#include "kernels.cuh"
#include "zgemm.cuh"
#include "fft_launch.h"

using namespace pawb200;

// ---------------------------------------------------------------------------------------
// errors, CUDA helpers, timers
// ---------------------------------------------------------------------------------------
namespace {

thread_local std::string g_error;
thread_local bool g_has_error = false;

void set_error(const std::string& m) {
  g_error = m;
  g_has_error = true;
  if (getenv("PAWB200_VERBOSE")) fprintf(stderr, "[pawb200] error: %s\n", m.c_str());
}

#define CUDA_OK(expr)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e_) + " at " +    \
                               __FILE__ + ":" + std::to_string(__LINE__));                  \
  } while (0)
#define CUFFT_OK(expr)                                                                      \
  do {                                                                                      \
    cufftResult r_ = (expr);                                                                \
    if (r_ != CUFFT_SUCCESS)                                                                \
      throw std::runtime_error("cuFFT error " + std::to_string((int)r_) + " at " + __FILE__ + \
                               ":" + std::to_string(__LINE__));                             \
  } while (0)
// The engine is not re-entrant (one main stream, shared scratch, stream-ordered pools): every C-ABI entry point
// that touches device state is serialised on one process-wide lock, so Python threads that call in with the
// GIL released queue up instead of corrupting each other's buffers.
std::recursive_mutex g_api_mutex;
#define API_BEGIN \
  std::lock_guard<std::recursive_mutex> api_lock_(g_api_mutex); \
  g_has_error = false; \
  try {
#define API_END(ret)                      \
  }                                       \
  catch (const std::exception& e) {       \
    set_error(e.what());                  \
    return ret;                           \
  }
#define API_END_VOID                      \
  }                                       \
  catch (const std::exception& e) {       \
    set_error(e.what());                  \
    return;                               \
  }

// Exceptions must not leave an OpenMP region (std::terminate): loop bodies run through OmpErr::run, the first
// exception is kept and rethrown after the loop.
struct OmpErr {
  std::exception_ptr e;
  std::atomic<bool> has{false};
  template <class F> void run(F&& f) noexcept {
    try { f(); } catch (...) { if (!has.exchange(true)) e = std::current_exception(); }
  }
  void rethrow() { if (has) std::rethrow_exception(e); }
};

cudaStream_t g_stream = 0;   // legacy default stream: ordered with torch's default stream
std::atomic<long long> g_launches{0};
long long g_boxes_scattered = 0, g_boxes_fft = 0, g_slots_projected = 0, g_sphere_samples = 0;
int g_num_sms = 0;

void require_device() {
  static int state = 0;   // 0 unknown, 1 ok, -1 failed
  static std::string why;
  if (state == 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      state = -1;
      why = std::string("no CUDA device available (") +
            (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
            "); pawpyseed_b200 has no CPU fallback";
      cudaGetLastError();
    } else {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceProp p;
      cudaGetDeviceProperties(&p, dev);
      if (p.major < 10) {
        state = -1;
        why = "GPU '" + std::string(p.name) + "' is sm_" + std::to_string(p.major) +
              std::to_string(p.minor) + "; this library is built for sm_100a (B200) only";
      } else {
        state = 1;
        g_num_sms = p.multiProcessorCount;
      }
    }
  }
  if (state < 0) throw std::runtime_error(why);
}

void trim_all_pools();

// Stream-ordered caching allocator: all work runs on one stream, so a block returned to the pool
// can be handed out again without a device synchronisation (later kernels are ordered after the
// earlier users).  Avoids cudaMalloc/cudaFree (and their implicit syncs) on the per-step path.
struct DevPool {
  std::map<size_t, std::vector<void*>> free_;
  size_t cached_bytes = 0;
  static size_t bucket(size_t n) {
    if (n <= (1u << 20)) {
      size_t b = 256;
      while (b < n) b <<= 1;
      return b;
    }
    const size_t g = (size_t)2 << 20;
    return (n + g - 1) / g * g;
  }
  void* get(size_t& n) {
    n = bucket(n);
    auto it = free_.find(n);
    if (it != free_.end() && !it->second.empty()) {
      void* p = it->second.back();
      it->second.pop_back();
      cached_bytes -= n;
      return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
      cudaGetLastError();
      trim_all_pools();
      e = cudaMalloc(&p, n);
      if (e != cudaSuccess)
        throw std::runtime_error(std::string("CUDA: out of device memory allocating ") + std::to_string(n >> 20) + " MiB");
    }
    return p;
  }
  void put(void* p, size_t n) {
    free_[n].push_back(p);
    cached_bytes += n;
  }
  void trim() {
    cudaDeviceSynchronize();
    for (auto& kv : free_)
      for (void* p : kv.second) cudaFree(p);
    free_.clear();
    cached_bytes = 0;
  }
};
DevPool g_pool;
// Second pool for buffers that are written on the ingest streams (coefficients, their interleaved copy, the
// box-order permutation).  Blocks only come back through pawb200_free_pswf, which waits for every stream first,
// so a block taken from it is quiescent and may be used on any stream without ordering it behind the main one.
DevPool g_xpool;
void trim_all_pools() { g_pool.trim(); g_xpool.trim(); }

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool x = false;   // block belongs to g_xpool
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), x(o.x) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; x = o.x; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n) {
    release();
    if (n == 0) return;
    x = false;
    p = g_pool.get(n);
    bytes = n;
  }
  void alloc_x(size_t n) {   // cross-stream buffer, see g_xpool
    release();
    if (n == 0) return;
    x = true;
    p = g_xpool.get(n);
    bytes = n;
  }
  void ensure(size_t n) { if (n > bytes) alloc(n); }
  void release() { if (p) (x ? g_xpool : g_pool).put(p, bytes); p = nullptr; bytes = 0; }
  void zero(cudaStream_t st = g_stream) { if (p) CUDA_OK(cudaMemsetAsync(p, 0, bytes, st)); }
  void zero(size_t n, cudaStream_t st = g_stream) { if (p) CUDA_OK(cudaMemsetAsync(p, 0, std::min(n, bytes), st)); }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Pinned staging arena for host->device uploads of setup data: bump allocation, recycled at the
// synchronisation points the API already has (or when full).
cudaStream_t g_unpack_stream = nullptr;   // ingest unpack stream (also reads the arena), set by ingest_ring()
void sync_arena_users() {
  CUDA_OK(cudaStreamSynchronize(g_stream));
  if (g_unpack_stream) CUDA_OK(cudaStreamSynchronize(g_unpack_stream));
}
// Two halves used alternately: when one fills up the arena moves to the other and only waits for the device work that
// was queued when THAT half was left (an event per consuming stream) - normally long finished.  r02: the single
// region of r01 synchronised the main stream on every wrap, which stalls the host for as long as queued transforms
// wait for coefficients that are still crossing the host link (traced at 4 GPUs: 37 ms per build_site_tables call).
struct PinnedArena {
  unsigned char* base = nullptr;
  size_t cap = 0, used = 0;          // cap: bytes per half
  int cur = 0;
  cudaEvent_t left_main[2] = {nullptr, nullptr}, left_unpack[2] = {nullptr, nullptr};
  bool left_valid[2] = {false, false};
  void* take(size_t n) {
    n = (n + 255) / 256 * 256;
    if (n > cap) {   // grow: everything in flight must land first
      sync_arena_users();
      if (base) cudaFreeHost(base);
      cap = std::max<size_t>(n * 2, (size_t)64 << 20);
      CUDA_OK(cudaMallocHost((void**)&base, 2 * cap));
      used = 0;
      cur = 0;
      left_valid[0] = left_valid[1] = false;
    }
    if (used + n > cap) {
      if (!left_main[0])
        for (int i = 0; i < 2; i++) {
          CUDA_OK(cudaEventCreateWithFlags(&left_main[i], cudaEventDisableTiming));
          CUDA_OK(cudaEventCreateWithFlags(&left_unpack[i], cudaEventDisableTiming));
        }
      CUDA_OK(cudaEventRecord(left_main[cur], g_stream));
      if (g_unpack_stream) CUDA_OK(cudaEventRecord(left_unpack[cur], g_unpack_stream));
      left_valid[cur] = true;
      cur ^= 1;
      if (left_valid[cur]) {          // readers of the half we are about to overwrite
        CUDA_OK(cudaEventSynchronize(left_main[cur]));
        if (g_unpack_stream) CUDA_OK(cudaEventSynchronize(left_unpack[cur]));
        left_valid[cur] = false;
      }
      used = 0;
    }
    void* p = base + (size_t)cur * cap + used;
    used += n;
    return p;
  }
  void reset_after_sync() { used = 0; left_valid[0] = left_valid[1] = false; }
};
PinnedArena g_arena;

void stream_sync() {
  sync_arena_users();
  g_arena.reset_after_sync();
}

// pinned arena block -> device, by kernel (see pinned_fetch_kernel)
void fetch_pinned(void* dst, const void* pinned_src, size_t bytes, cudaStream_t st) {
  if (!bytes) return;
  const unsigned blocks = (unsigned)std::min<size_t>(64, (bytes / 16 + 255) / 256 + 1);
  pinned_fetch_kernel<<<blocks, 256, 0, st>>>((const unsigned char*)pinned_src, (unsigned char*)dst, bytes);
  CUDA_OK(cudaGetLastError());
}

template <typename T>
DevBuf upload(const T* v, size_t n, cudaStream_t st = g_stream, bool xpool = false) {
  DevBuf b;
  if (xpool) b.alloc_x(std::max<size_t>(n, 1) * sizeof(T)); else b.alloc(std::max<size_t>(n, 1) * sizeof(T));
  if (n) {
    void* stg = g_arena.take(n * sizeof(T));
    memcpy(stg, v, n * sizeof(T));
    fetch_pinned(b.p, stg, n * sizeof(T), st);
  }
  return b;
}
template <typename T>
DevBuf upload(const std::vector<T>& v, cudaStream_t st = g_stream, bool xpool = false) {
  return upload(v.data(), v.size(), st, xpool);
}

// ---- stage timers (CUDA events on the launch stream) ---------------------------------------
enum Stage { ST_H2D, ST_SCATTER, ST_FFT, ST_PROJECT, ST_TABLE, ST_GEMM_PS, ST_GEMM_AUG, ST_AUGMENT,
             ST_D2H, ST_COUNT };
struct EventPair { cudaEvent_t a, b; int stage; };
std::vector<EventPair> g_pending;
std::vector<cudaEvent_t> g_event_pool;
double g_stage_ms[ST_COUNT] = {0};
bool g_timing = true;

cudaEvent_t get_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  CUDA_OK(cudaEventCreate(&e));
  return e;
}
struct ScopedStage {
  EventPair ep;
  bool on;
  cudaStream_t st;
  explicit ScopedStage(int stage, cudaStream_t stream = nullptr) : on(g_timing), st(stream ? stream : g_stream) {
    if (!on) return;
    ep.a = get_event();
    ep.b = get_event();
    ep.stage = stage;
    cudaEventRecord(ep.a, st);
  }
  ~ScopedStage() {
    if (!on) return;
    cudaEventRecord(ep.b, st);
    g_pending.push_back(ep);
  }
};
void drain_timers() {
  for (auto& ep : g_pending) {
    cudaEventSynchronize(ep.b);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) g_stage_ms[ep.stage] += ms;
    g_event_pool.push_back(ep.a);
    g_event_pool.push_back(ep.b);
  }
  g_pending.clear();
}

// ---- host wall-clock sections (PAWB200_PROFILE=1 prints them from pawb200_get_timers) ---------
struct HostProf {
  std::map<std::string, double> ms;
  std::map<std::string, long> calls;
};
HostProf g_hostprof;
struct HostSection {
  const char* name;
  std::chrono::steady_clock::time_point t0;
  explicit HostSection(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
  ~HostSection() {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    g_hostprof.ms[name] += ms;
    g_hostprof.calls[name] += 1;
  }
};

// Device-side trace for debugging overlap (PAWB200_TRACE=1): trace_mark records an event on a stream; the marks
// are printed, relative to the first one, when pawb200_get_timers is called.
struct TraceMark { std::string name; cudaEvent_t ev; };
std::vector<TraceMark> g_trace;
bool trace_on() {
  static const bool on = getenv("PAWB200_TRACE") != nullptr;
  return on;
}
void trace_mark(const std::string& name, cudaStream_t st) {
  if (!trace_on()) return;
  TraceMark m{name, nullptr};
  cudaEventCreate(&m.ev);
  cudaEventRecord(m.ev, st);
  g_trace.push_back(m);
}
void trace_dump() {
  if (g_trace.empty()) return;
  cudaDeviceSynchronize();
  for (auto& m : g_trace) {
    float ms = 0;
    cudaEventElapsedTime(&ms, g_trace[0].ev, m.ev);
    fprintf(stderr, "[trace] %8.3f ms  %s\n", ms, m.name.c_str());
  }
  for (auto& m : g_trace) cudaEventDestroy(m.ev);
  g_trace.clear();
}

inline void count_launch(int n = 1) { g_launches += n; }
inline void check_launch() { CUDA_OK(cudaGetLastError()); }

// ---- cuFFT plan cache -------------------------------------------------------------------------
struct PlanKey {
  int n0, n1, n2, batch;
  bool operator<(const PlanKey& o) const {
    return std::tie(n0, n1, n2, batch) < std::tie(o.n0, o.n1, o.n2, o.batch);
  }
};
std::map<PlanKey, cufftHandle> g_plans;
DevBuf g_fft_work;

cufftHandle get_plan(const int* fftg, int batch) {
  PlanKey k{fftg[0], fftg[1], fftg[2], batch};
  auto it = g_plans.find(k);
  if (it != g_plans.end()) return it->second;
  cufftHandle h;
  CUFFT_OK(cufftCreate(&h));
  CUFFT_OK(cufftSetAutoAllocation(h, 0));
  int n[3] = {fftg[0], fftg[1], fftg[2]};
  size_t ws = 0;
  long long nn[3] = {n[0], n[1], n[2]};
  long long dist = (long long)n[0] * n[1] * n[2];
  CUFFT_OK(cufftMakePlanMany64(h, 3, nn, nullptr, 1, dist, nullptr, 1, dist, CUFFT_Z2Z, batch, &ws));
  CUFFT_OK(cufftSetStream(h, g_stream));
  if (ws > g_fft_work.bytes) {
    CUDA_OK(cudaStreamSynchronize(g_stream));
    g_fft_work.alloc(ws);
    for (auto& kv : g_plans) CUFFT_OK(cufftSetWorkArea(kv.second, g_fft_work.p));
  }
  CUFFT_OK(cufftSetWorkArea(h, g_fft_work.p));
  g_plans[k] = h;
  return h;
}

size_t fft_budget_bytes() {
  const char* e = getenv("PAWB200_FFT_BYTES");
  if (e) return (size_t)atoll(e);
  return (size_t)6 << 30;
}

// ---------------------------------------------------------------------------------------
// site table sets
// ---------------------------------------------------------------------------------------
struct ElementList {   // what pawb200_ppot_t points to
  std::vector<Element> el;
};

struct ElemDevStore {   // device copies of the radial data for one table mode
  std::vector<DevBuf> bufs;
  DevBuf dev;           // ElemDev[nelem]
};

// mode 0: projectors, 1: filtered phi-phit on the linear grid, 2: phi-phit on the log grid
ElemDevStore upload_elements(const std::vector<Element>& els, int mode) {
  ElemDevStore st;
  std::vector<ElemDev> host(els.size());
  for (size_t e = 0; e < els.size(); e++) {
    const Element& el = els[e];
    const std::vector<double>& grid = mode == 0 ? el.proj_grid : (mode == 1 ? el.smooth_grid : el.wave_grid);
    const int n = (int)grid.size();
    std::vector<double> f((size_t)el.num_projs * n), spl((size_t)el.num_projs * 3 * n);
    for (int k = 0; k < el.num_projs; k++) {
      const RadialFunc& rf = el.funcs[k];
      const std::vector<double>& v = mode == 0 ? rf.proj : (mode == 1 ? rf.smooth_diffwave : rf.diffwave);
      const Spline& s = mode == 0 ? rf.proj_s : (mode == 1 ? rf.smooth_s : rf.diffwave_s);
      std::copy(v.begin(), v.end(), f.begin() + (size_t)k * n);
      for (int r = 0; r < 3; r++)
        std::copy(s.c[r].begin(), s.c[r].end(), spl.begin() + ((size_t)k * 3 + r) * n);
    }
    std::vector<int> cn, cl, cm;
    for (auto& c : el.chan) { cn.push_back(c.n); cl.push_back(c.l); cm.push_back(c.m); }
    st.bufs.push_back(upload(grid));  host[e].grid = st.bufs.back().as<double>();
    st.bufs.push_back(upload(f));     host[e].f = st.bufs.back().as<double>();
    st.bufs.push_back(upload(spl));   host[e].spl = st.bufs.back().as<double>();
    st.bufs.push_back(upload(cn));    host[e].chan_n = st.bufs.back().as<int>();
    st.bufs.push_back(upload(cl));    host[e].chan_l = st.bufs.back().as<int>();
    st.bufs.push_back(upload(cm));    host[e].chan_m = st.bufs.back().as<int>();
    host[e].n = n;
    host[e].nfunc = el.num_projs;
    host[e].nchan = el.total_projs;
    host[e].rmax = mode == 0 ? el.rmax : el.wave_rmax;
  }
  st.dev = upload(host);
  return st;
}

struct SiteTables {
  int nsites = 0;
  int mode = 0;
  int fftg[3] = {0, 0, 0};
  std::vector<SiteDev> host;          // per site
  std::vector<int> site_id;           // index of each site in the structure
  long total_pts = 0;                 // padded
  long total_tab = 0;                 // double2 elements
  int nproj = 0;                      // concatenated channel count
  DevBuf sites, idx, path, wrap, table, tablek;
  DevBuf ureal, phk, chan_m;          // real-table form of the projection (modes 0/1): see sphere_project_real_kernel
  std::vector<std::vector<int>> by_mt;   // site indices grouped by m-tile count (1..3)
  std::vector<DevBuf> by_mt_dev;
  std::vector<std::vector<int32_t>> host_idx;   // kept for the index-parity accessor
  size_t coord_hash = 0;
};

size_t hash_bytes(const void* p, size_t n, size_t seed = 1469598103934665603ull) {
  const unsigned char* b = (const unsigned char*)p;
  size_t h = seed;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

// Geometry on the host (bit-exact membership), values on the GPU.
std::unique_ptr<SiteTables> build_site_tables(const std::vector<Element>& els, const int* site_list,
                                              int nlist, const int* labels, const double* coords,
                                              const double* lattice, const int* fftg, int mode,
                                              bool keep_host_idx) {
  HostSection hs_("build_site_tables");
  auto T = std::make_unique<SiteTables>();
  T->nsites = nlist;
  T->mode = mode;
  for (int d = 0; d < 3; d++) T->fftg[d] = fftg[d];
  T->host.resize(nlist);
  T->site_id.assign(site_list, site_list + nlist);
  std::vector<SphereGeom> geom(nlist);
  auto* hs_geom = new HostSection("  sphere_geometry");
  OmpErr oe;
#pragma omp parallel for schedule(dynamic)
  for (int s = 0; s < nlist; s++) oe.run([&] {
    const int p = site_list[s];
    const Element& el = els[labels[p]];
    const double rmax = mode == 0 ? el.rmax : el.wave_rmax;
    double radius = rmax;
    if (mode != 2) {
      // utils.c:651-652: (n-1)*rmax/n with the divisor taken from pps[labels[s]] (loop index)
      const int ndiv = els[labels[s]].proj_gridsize;
      radius = (el.proj_gridsize - 1) * rmax / ndiv;
    }
    geom[s] = sphere_geometry(coords + 3 * p, lattice, fftg, rmax, radius);
  });
  delete hs_geom;
  oe.rethrow();
  long pt = 0, tab = 0;
  int lm = 0;
  for (int s = 0; s < nlist; s++) {
    const int p = site_list[s];
    const Element& el = els[labels[p]];
    SiteDev& sd = T->host[s];
    sd.elem = labels[p];
    sd.npts = (int)geom[s].index.size();
    sd.npts_pad = (sd.npts + 31) / 32 * 32;
    sd.pt_off = pt;
    sd.tab_off = tab;
    sd.nlm = el.total_projs;
    sd.lm_off = lm;
    for (int d = 0; d < 3; d++) sd.coord[d] = coords[3 * p + d];
    pt += sd.npts_pad;
    tab += (long)sd.nlm * sd.npts_pad;
    lm += sd.nlm;
    if (sd.nlm > 24) throw std::runtime_error("more than 24 channels per site is not supported");
  }
  T->total_pts = pt;
  T->total_tab = tab;
  T->nproj = lm;
  // The host only ships the unwrapped grid coordinates of the selected points (8 B per point, packed straight into
  // pinned staging memory in parallel over sites); expand_geometry_kernel rebuilds index, offsets and cell shifts.
  const size_t npt = (size_t)std::max<long>(pt, 1);
  // NOTE: a host pointer returned by g_arena.take() is only valid until the next take() (which may wrap or grow
  // the arena), so every other upload of this function happens before the ijk block is taken and the block is
  // fetched to the device before anything else touches the arena.
  T->sites = upload(T->host);
  T->by_mt.assign(4, {});
  for (int s = 0; s < nlist; s++) T->by_mt[(T->host[s].nlm + 7) / 8].push_back(s);
  T->by_mt_dev.resize(4);
  for (int m = 1; m <= 3; m++)
    if (!T->by_mt[m].empty()) T->by_mt_dev[m] = upload(T->by_mt[m]);
  int16_t* ijk = (int16_t*)g_arena.take(npt * 4 * sizeof(int16_t));
#pragma omp parallel for schedule(dynamic)
  for (int s = 0; s < nlist; s++) {
    const SiteDev& sd = T->host[s];
    if (sd.npts) memcpy(ijk + 4 * sd.pt_off, geom[s].ijk.data(), (size_t)sd.npts * 4 * sizeof(int16_t));
    memset(ijk + 4 * (sd.pt_off + sd.npts), 0, (size_t)(sd.npts_pad - sd.npts) * 4 * sizeof(int16_t));
  }
  if (keep_host_idx) {
    T->host_idx.resize(nlist);
    for (int s = 0; s < nlist; s++) T->host_idx[s] = std::move(geom[s].index);
  }
  T->idx.alloc(npt * sizeof(int32_t));
  T->path.alloc(3 * npt * sizeof(double));
  if (mode == 2) T->wrap.alloc(3 * npt * sizeof(int32_t));
  {
    DevBuf dijk(npt * 4 * sizeof(int16_t));
    fetch_pinned(dijk.p, ijk, npt * 4 * sizeof(int16_t), g_stream);
    int maxpad = 0;
    for (auto& sd : T->host) maxpad = std::max(maxpad, sd.npts_pad);
    if (nlist > 0 && maxpad > 0) {
      Lattice9 L9;
      for (int i = 0; i < 9; i++) L9.a[i] = lattice[i];
      dim3 grid((maxpad + 255) / 256, nlist);
      expand_geometry_kernel<<<grid, 256, 0, g_stream>>>(T->sites.as<SiteDev>(), dijk.as<short4>(), L9, fftg[0], fftg[1],
                                                         fftg[2], T->idx.as<int>(), T->path.as<double>(),
                                                         mode == 2 ? T->wrap.as<int>() : nullptr, (long)pt);
      count_launch();
      check_launch();
    }
  }
  T->table.alloc(std::max<size_t>(1, tab) * sizeof(double2));

  if (nlist > 0 && pt > 0) {
    ScopedStage tm(ST_TABLE);
    ElemDevStore ed = upload_elements(els, mode);
    std::vector<double> lat(lattice, lattice + 9);
    DevBuf dlat = upload(lat);
    int maxpts = 0;
    for (auto& sd : T->host) maxpts = std::max(maxpts, sd.npts_pad);
    dim3 grid((maxpts + 127) / 128, nlist);
    site_table_kernel<<<grid, 128, 0, g_stream>>>(
        T->sites.as<SiteDev>(), ed.dev.as<ElemDev>(), T->idx.as<int>(), T->path.as<double>(), pt,
        T->table.as<double2>(), dlat.as<double>(), fftg[0], fftg[1], fftg[2], mode == 2 ? 1 : 0);
    count_launch();
    check_launch();   // ed / dlat return to the stream-ordered pool
    if (mode != 2) {
      // real rows of the table + the m of every channel (k-independent)
      std::vector<int> cm((size_t)std::max(lm, 1), 0);
      for (int s = 0; s < nlist; s++) {
        const Element& el = els[labels[site_list[s]]];
        for (int c = 0; c < T->host[s].nlm; c++) cm[T->host[s].lm_off + c] = el.chan[c].m;
      }
      T->chan_m = upload(cm);
      T->ureal.alloc(std::max<size_t>(1, tab) * sizeof(double));
      dim3 grid2((maxpts + 255) / 256, nlist);
      real_table_kernel<<<grid2, 256, 0, g_stream>>>(T->sites.as<SiteDev>(), T->chan_m.as<int>(),
                                                     T->table.as<double2>(), T->ureal.as<double>());
      count_launch();
      check_launch();
    }
  }
  return T;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// wavefunction state
// ---------------------------------------------------------------------------------------
struct pawb200_ppot {
  ElementList list;
};

namespace { struct PrunedPlan; }

struct HostMatrixCache {
  uint64_t other_id = 0, other_gen = 0, self_gen = 0;
  int flip = -1;
  int recip = 0;
  size_t list_hash = 0;
  std::vector<cdouble> data;   // [NK][nbS][nbR]
  bool valid = false;
};

static std::atomic<uint64_t> g_next_id{1};

struct pawb200_pswf {
  uint64_t id = g_next_id++;
  uint64_t gen = 1;
  int nspin = 0, nwk = 0, nband = 0, ncl = 0;
  double encut = 0;
  double lattice[9], reclattice[9];
  int G_bounds[6] = {0, 0, 0, 0, 0, 0};
  std::vector<KPointInfo> kp;          // per kappa
  std::vector<double> weight;          // per kappa
  std::vector<DevBuf> C;               // per kappa: float2 [nband][ldc]
  std::vector<long> ldc;
  std::vector<DevBuf> Cil;             // per kappa: float2 [ceil(nslot/16)][ldil][16], 16-slot interleaved copy (pass Z)
  std::vector<long> ldil;
  std::vector<char> resident;          // per kappa: this process holds the block
  std::vector<DevBuf> perm_dev;        // per kappa: device copy of the box-order permutation
  struct Chunk { int kap, band_lo, band_hi; cudaEvent_t ready; };
  std::vector<Chunk> chunks;           // ingest chunks: coefficients of [band_lo, band_hi) are valid after `ready`
  ~pawb200_pswf() {
    for (auto& c : chunks) cudaEventDestroy(c.ready);
  }
  // projector state
  std::unique_ptr<pawb200_ppot> pps;
  int num_sites = 0;
  int fftg[3] = {0, 0, 0};
  std::vector<int> labels;
  std::vector<double> coords;
  std::unique_ptr<SiteTables> proj_sites;
  std::vector<DevBuf> P;               // per kappa: double2 [nslot][ldp]
  long ldp = 0;
  bool has_projections = false;
  // setup_projections could not keep the FFT boxes resident: the transform + projection of the bands is deferred
  // to the first consumer, so that overlap_setup_real can project its partial-wave tables in the SAME pass over
  // the boxes instead of transforming every band a second time (see require_projections)
  bool lazy_proj = false;
  // overlap_setup state (this wf's bands projected on the other structure's filtered partial waves)
  std::vector<DevBuf> W;               // per kappa: double2 [nslot][ldw]
  long ldw = 0;
  int wp_num = 0;
  std::vector<int> wp_nlm;             // channels per wave-projection site
  std::vector<DevBuf> CA;              // (f3) per kappa: float2 [nband][ldc] augmentation part in the PW basis
  bool recip_setup = false;            // off-site data came from overlap_setup_recip
  // as wf_S: off-site data
  std::vector<std::vector<cdouble>> omega;
  std::vector<double> dcoords;
  std::vector<int> omega_n1, omega_n2;
  uint64_t overlap_partner = 0;
  // FFT boxes psi~(r) of every slot kept in HBM after setup_projections (if they fit the budget), so
  // overlap_setup_real does not scatter + transform the bands a second time
  std::vector<DevBuf> boxes;
  int boxes_fftg[3] = {0, 0, 0};
  bool boxes_interleaved = false;
  // pruned-FFT plans (column / plane tables) per kappa for the grid in plan_fftg
  std::vector<std::shared_ptr<PrunedPlan>> fft_plans;
  int plan_fftg[3] = {0, 0, 0};   // layout of `boxes`: [group][grid][16] (pruned FFT) or [slot][grid] (cuFFT)
  // real-space table cache (mode 2), keyed by grid + coords
  std::unique_ptr<SiteTables> ae_sites;
  // per-band call caches
  HostMatrixCache pseudo_cache, aug_cache;

  // band-block sharding of one (k,spin) block over ranks (SURVEY 8e level 2): this process transforms / projects /
  // multiplies only the bands [band_lo, band_hi); row buffers (C, P, W) hold band_rows = per * world rows so that the
  // blocks of all ranks can be all-gathered in place (pawb200_get_device_buffer + NCCL)
  int band_lo = 0, band_hi = 0, band_rows = 0;
  int nkappa() const { return nwk * nspin; }
  int halves() const { return ncl ? 2 : 1; }
  int nslot() const { return nband * halves(); }
  int slot_lo() const { return band_lo * halves(); }
  int slot_own() const { return (band_hi - band_lo) * halves(); }
  bool band_sharded() const { return band_lo != 0 || band_hi != nband; }
  int npw_half(int kap) const { return kp[kap].nplane / halves(); }
};

namespace {

int g_shard_rank = 0, g_shard_world = 1;
int g_band_rank = 0, g_band_world = 1;

// ---- WAVECAR ingest -----------------------------------------------------------------------
struct ByteSource {
  const unsigned char* mem = nullptr;
  FILE* fp = nullptr;
  void read(void* dst, long off, size_t n) {
    if (mem) {
      memcpy(dst, mem + off, n);
    } else {
      if (fseek(fp, off, SEEK_SET) != 0 || fread(dst, 1, n, fp) != n)
        throw std::runtime_error("short read from WAVECAR");
    }
  }
};

// Staging ring for raw WAVECAR records.  Copies and the column permutation run on their own streams in band
// chunks; each chunk records an event that the consumers of those bands (FFT / scatter / GEMM launches on the
// main stream) wait on.  The H2D of a wavefunction therefore overlaps both the kernels of the previous
// wavefunction and - with pawb200_set_async_ingest(1) - its own transform pipeline.
struct IngestRing {
  struct Slot {
    void* p = nullptr;
    size_t bytes = 0;
    cudaEvent_t filled = nullptr, free = nullptr;   // copy landed / unpack kernels have consumed it
  };
  // `copy` carries only the H2D transfers, `unpack` (high priority) the permutation / interleave kernels, so a
  // transfer never queues behind a kernel that is waiting for SMs held by the persistent FFT / GEMM kernels.
  cudaStream_t copy = nullptr, unpack = nullptr;
  Slot slots[3];
  unsigned next = 0;
};
IngestRing& ingest_ring() {
  static IngestRing r;
  if (!r.copy) {
    int lo = 0, hi = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_OK(cudaStreamCreateWithFlags(&r.copy, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithPriority(&r.unpack, cudaStreamNonBlocking, hi));
    g_unpack_stream = r.unpack;
    for (auto& sl : r.slots) {
      CUDA_OK(cudaEventCreateWithFlags(&sl.filled, cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&sl.free, cudaEventDisableTiming));
    }
  }
  return r;
}
bool g_async_ingest = false;

// Storage order of the plane waves = FFT-box index order (wrap(g1), wrap(g2), wrap(g3)) with
// wrap(g) = g for g >= 0 and negatives after all non-negatives (independent of the grid size).
// The file order is already sorted by (wrap(g3), wrap(g2), wrap(g1)) (reader.c:230-246 loop nest),
// so two stable counting sorts (by g2, then g1) produce the target order in O(n).
void box_order(KPointInfo& kp) {
  const int ng = (int)(kp.G.size() / 3);
  std::vector<int32_t> cur(ng), nxt(ng);
  for (int w = 0; w < ng; w++) cur[w] = w;
  bool sorted3 = true;   // verify the assumption; fall back to a comparison sort otherwise
  auto key = [](int g, int lo, int hi) { return g >= 0 ? g : (hi + 1) + (g - lo); };
  int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  for (int w = 0; w < ng; w++)
    for (int d = 0; d < 3; d++) {
      lo[d] = std::min(lo[d], (int)kp.G[3 * w + d]);
      hi[d] = std::max(hi[d], (int)kp.G[3 * w + d]);
    }
  for (int w = 1; w < ng && sorted3; w++) {
    long a = 0, b = 0;
    for (int d = 2; d >= 0; d--) {
      a = a * 4096 + key(kp.G[3 * (w - 1) + d], lo[d], hi[d]);
      b = b * 4096 + key(kp.G[3 * w + d], lo[d], hi[d]);
    }
    if (a > b) sorted3 = false;
  }
  if (sorted3 && hi[0] - lo[0] < 4000 && hi[1] - lo[1] < 4000 && hi[2] - lo[2] < 4000) {
    for (int d = 1; d >= 0; d--) {
      const int nk = hi[d] - lo[d] + 2;
      std::vector<int> cnt(nk + 1, 0);
      for (int w = 0; w < ng; w++) cnt[key(kp.G[3 * cur[w] + d], lo[d], hi[d]) + 1]++;
      for (int q = 0; q < nk; q++) cnt[q + 1] += cnt[q];
      for (int w = 0; w < ng; w++) nxt[cnt[key(kp.G[3 * cur[w] + d], lo[d], hi[d])]++] = cur[w];
      cur.swap(nxt);
    }
  } else {
    std::stable_sort(cur.begin(), cur.end(), [&](int a, int b) {
      for (int d = 0; d < 3; d++) {
        const int ka = key(kp.G[3 * a + d], lo[d], hi[d]), kb = key(kp.G[3 * b + d], lo[d], hi[d]);
        if (ka != kb) return ka < kb;
      }
      return false;
    });
  }
  kp.perm = cur;
  kp.pos.resize(ng);
  for (int j = 0; j < ng; j++) kp.pos[kp.perm[j]] = j;
}

void interleave_rows(pawb200_pswf* wf, int kap, int band_lo, int band_hi, cudaStream_t st);
void alloc_interleaved(pawb200_pswf* wf, int kap, cudaStream_t st);

// One band chunk of one (k,spin) block on its way to the device: raw records -> staging slot (copy stream) ->
// box-ordered rows + interleaved copy (unpack stream) -> `ready` event in wf->chunks.
struct ChunkJob {
  pawb200_pswf* wf;
  int kap, b0, nb, per;
  long ld, nrecl;
  int nplane, half_len;
  const unsigned char* from;        // first record of the block (host memory that outlives the copy)
};
// Asynchronous ingest from caller-owned memory issues only the FIRST block of a wavefunction at read time and defers
// the others until some consumer asks for coefficients (wait_coeffs): by then the second wavefunction of a pair has
// usually been read as well, and the deferred chunks of both are issued in (k,spin)-block order - basis block 1, wf
// block 1, basis block 2, ... - so the pseudo GEMM of block k can start when 2(k+1) blocks have crossed the link
// instead of waiting for every block of the basis first (e2e of the 8-block job on 2 GPUs: 968 -> see DESIGN 5).
std::vector<ChunkJob> g_pending_chunks;

void issue_chunk(const ChunkJob& j) {
  IngestRing& ring = ingest_ring();
  pawb200_pswf* wf = j.wf;
  IngestRing::Slot& sl = ring.slots[ring.next++ % 3];
  const size_t raw_bytes = (size_t)j.per * j.ld * sizeof(float2);
  if (raw_bytes > sl.bytes) {
    CUDA_OK(cudaStreamSynchronize(ring.copy));
    CUDA_OK(cudaStreamSynchronize(ring.unpack));
    if (sl.p) cudaFree(sl.p);
    CUDA_OK(cudaMalloc(&sl.p, raw_bytes));
    sl.bytes = raw_bytes;
  }
  CUDA_OK(cudaStreamWaitEvent(ring.copy, sl.free, 0));     // the slot's previous contents were unpacked
  trace_mark("h2d chunk start k=" + std::to_string(j.kap) + " b0=" + std::to_string(j.b0), ring.copy);
  {
    ScopedStage tm(ST_H2D, ring.copy);
    CUDA_OK(cudaMemcpy2DAsync(sl.p, j.ld * sizeof(float2), j.from + (size_t)j.b0 * j.nrecl, j.nrecl,
                              (size_t)j.nplane * sizeof(float2), j.nb, cudaMemcpyHostToDevice, ring.copy));
  }
  CUDA_OK(cudaEventRecord(sl.filled, ring.copy));
  trace_mark("h2d chunk done k=" + std::to_string(j.kap) + " b0=" + std::to_string(j.b0), ring.copy);
  CUDA_OK(cudaStreamWaitEvent(ring.unpack, sl.filled, 0));
  dim3 grid((j.half_len + 255) / 256, std::min(j.nb, 64));
  permute_coeff_kernel<<<grid, 256, 0, ring.unpack>>>((const float2*)sl.p, wf->C[j.kap].as<float2>() + (long)j.b0 * j.ld,
                                                   j.ld, j.nb, wf->ncl ? 2 : 1, j.half_len,
                                                   wf->perm_dev[j.kap].as<int>());
  count_launch();
  check_launch();
  interleave_rows(wf, j.kap, j.b0, j.b0 + j.nb, ring.unpack);
  CUDA_OK(cudaEventRecord(sl.free, ring.unpack));
  trace_mark("unpack done k=" + std::to_string(j.kap) + " b0=" + std::to_string(j.b0), ring.unpack);
  pawb200_pswf::Chunk ck{j.kap, j.b0, j.b0 + j.nb, nullptr};
  CUDA_OK(cudaEventCreateWithFlags(&ck.ready, cudaEventDisableTiming));
  CUDA_OK(cudaEventRecord(ck.ready, ring.unpack));
  wf->chunks.push_back(ck);
}

// Issue every deferred chunk, lowest (k,spin) block first, wavefunctions in the order they were read.
void flush_pending_ingest() {
  if (g_pending_chunks.empty()) return;
  std::vector<ChunkJob> jobs;
  jobs.swap(g_pending_chunks);
  std::stable_sort(jobs.begin(), jobs.end(), [](const ChunkJob& a, const ChunkJob& b) { return a.kap < b.kap; });
  for (auto& j : jobs) issue_chunk(j);
}

void drop_pending_ingest(const pawb200_pswf* wf) {
  g_pending_chunks.erase(std::remove_if(g_pending_chunks.begin(), g_pending_chunks.end(),
                                        [wf](const ChunkJob& j) { return j.wf == wf; }),
                         g_pending_chunks.end());
}

pawb200_pswf* ingest(ByteSource src, const double* kws) {
  HostSection hs_("ingest");
  require_device();
  auto wf = std::make_unique<pawb200_pswf>();
  double h0[3];
  src.read(h0, 0, 24);
  WavecarHeader hd;
  hd.nrecl = (long)std::round(h0[0]);
  hd.nspin = (int)std::round(h0[1]);
  if (hd.nrecl < 96 || hd.nspin < 1 || hd.nspin > 2) throw std::runtime_error("not a WAVECAR header");
  // VASP writes 45200 for complex64 and 45210 for complex128 coefficients; reader.c:143 reads the tag and ignores it,
  // so a double-precision file is silently
  // reinterpreted as complex64 garbage there.  Here it is an error.
  if ((long)std::round(h0[2]) != 45200)
    throw std::runtime_error("unsupported WAVECAR precision tag " + std::to_string((long)std::round(h0[2])) +
                             " (only 45200, complex64 coefficients, is supported - as in the reference)");
  std::vector<double> rec(hd.nrecl / 8);
  src.read(rec.data(), hd.nrecl, 12 * 8);
  hd.nwk = (int)std::round(rec[0]);
  hd.nband = (int)std::round(rec[1]);
  if (hd.nwk <= 0 || hd.nband <= 0) throw std::runtime_error("WAVECAR header: k-point / band count must be positive");
  if ((size_t)hd.nrecl / 8 < 4 + 3 * (size_t)hd.nband)
    throw std::runtime_error("WAVECAR record length " + std::to_string(hd.nrecl) + " is too short for the " +
                             std::to_string(hd.nband) + "-band k-point header (truncated or malformed file)");
  hd.encut = rec[2];
  for (int i = 0; i < 9; i++) hd.lattice[i] = rec[3 + i];
  wavecar_bounds(hd);
  wf->nspin = hd.nspin; wf->nwk = hd.nwk; wf->nband = hd.nband; wf->encut = hd.encut;
  {
    const int per = (hd.nband + g_band_world - 1) / g_band_world;
    wf->band_lo = std::min(hd.nband, g_band_rank * per);
    wf->band_hi = std::min(hd.nband, (g_band_rank + 1) * per);
    wf->band_rows = per * g_band_world;
  }
  memcpy(wf->lattice, hd.lattice, sizeof(hd.lattice));
  memcpy(wf->reclattice, hd.reclattice, sizeof(hd.reclattice));
  const int NK = hd.nwk * hd.nspin;
  wf->kp.resize(NK); wf->weight.resize(NK); wf->C.resize(NK); wf->ldc.assign(NK, 0);
  wf->resident.assign(NK, 0);
  wf->perm_dev.resize(NK);
  unsigned char* stage = nullptr;
  size_t stage_bytes = 0;
  IngestRing& ring = ingest_ring();
  bool first_block = true;
  for (int kap = 0; kap < NK; kap++) {
    const long base = 2 + (long)kap * (1 + hd.nband);
    const size_t hdr_doubles = std::min<size_t>(hd.nrecl / 8, 4 + 3 * (size_t)hd.nband);
    src.read(rec.data(), base * hd.nrecl, hdr_doubles * 8);
    KPointInfo& kp = wf->kp[kap];
    kp.nplane = (int)std::round(rec[0]);
    if (kp.nplane <= 0) throw std::runtime_error("k-point " + std::to_string(kap) + " has no plane waves");
    kp.k[0] = rec[1]; kp.k[1] = rec[2]; kp.k[2] = rec[3];
    kp.energy.resize(hd.nband); kp.occ.resize(hd.nband);
    for (int b = 0; b < hd.nband; b++) { kp.energy[b] = rec[4 + 3 * b]; kp.occ[b] = rec[6 + 3 * b]; }
    wf->weight[kap] = kws ? kws[kap % hd.nwk] : 1.0;
    const bool same_k = kap >= hd.nwk && kp.k[0] == wf->kp[kap - hd.nwk].k[0] &&
                        kp.k[1] == wf->kp[kap - hd.nwk].k[1] && kp.k[2] == wf->kp[kap - hd.nwk].k[2];
    const bool mine = kap % g_shard_world == g_shard_rank;
    if (same_k && !wf->kp[kap - hd.nwk].G.empty()) {   // second spin channel: same k, same list
      kp.G = wf->kp[kap - hd.nwk].G;
      kp.perm = wf->kp[kap - hd.nwk].perm;
      kp.pos = wf->kp[kap - hd.nwk].pos;
    } else if (mine) {
      kp.G = enumerate_g(hd, kp.k, wf->G_bounds);
      box_order(kp);
    } else {
      // sharded read, another rank's block: its plane-wave list is never used here (5-7 ms of host work per k-point
      // that delayed the first H2D chunk of ranks whose own block comes late in the file)
      continue;
    }
    const int ng = (int)(kp.G.size() / 3);
    if (2 * ng == kp.nplane) {
      wf->ncl = 1;
    } else if (ng != kp.nplane) {
      // reader.c:282-284 only prints; a mismatched basis cannot be mapped, so this is an error here
      throw std::runtime_error("plane-wave count mismatch at k-point " + std::to_string(kap) + ": " +
                               std::to_string(ng) + " enumerated vs " + std::to_string(kp.nplane) +
                               " stored (gamma-only WAVECARs are unsupported, as in the reference)");
    }
    if ((long)kp.nplane * 8 > hd.nrecl) throw std::runtime_error("record shorter than nplane");
    if (!mine) continue;   // sharded ingest: not this rank's block
    wf->resident[kap] = 1;
    const long ld = ((long)kp.nplane + 31) / 32 * 32;
    wf->ldc[kap] = ld;
    // cross-stream buffers from the quiescent pool: the unpack stream fills them without being ordered behind
    // whatever the main stream still has queued (e.g. the previous structure's transforms)
    wf->C[kap].alloc_x((size_t)wf->band_rows * ld * sizeof(float2));
    wf->C[kap].zero((size_t)wf->band_rows * ld * sizeof(float2), ring.unpack);
    const unsigned char* from;
    if (src.mem) {
      from = src.mem + (base + 1) * hd.nrecl;
    } else {
      const size_t need = (size_t)hd.nband * hd.nrecl;
      if (need > stage_bytes) {
        if (stage) cudaFreeHost(stage);
        CUDA_OK(cudaMallocHost((void**)&stage, need));
        stage_bytes = need;
      }
      CUDA_OK(cudaStreamSynchronize(ring.copy));   // previous block has left the staging buffer
      src.read(stage, (base + 1) * hd.nrecl, need);
      from = stage;
    }
    // raw records -> HBM in band chunks on the copy stream; the unpack stream permutes each chunk into box order
    // and writes its interleaved copy; consumers wait on the per-chunk events (wait_coeffs)
    {
      wf->perm_dev[kap] = upload(kp.perm, ring.unpack, true);
      alloc_interleaved(wf.get(), kap, ring.unpack);
      const int half_len = kp.nplane / (wf->ncl ? 2 : 1);
      const int nown = wf->band_hi - wf->band_lo;      // band-sharded ranks read only their band block
      const int nchunk = std::max(1, std::min(8, nown / 16));
      // whole GEMM row tiles (64) per chunk when there are enough bands, else whole interleave groups (16)
      const int gran = nown >= 256 ? 64 : 16;
      const int per = ((nown + nchunk - 1) / nchunk + gran - 1) / gran * gran;
      // asynchronous ingest from memory: the first resident block goes out now, the others when a consumer asks
      const bool defer = g_async_ingest && src.mem && !first_block;
      first_block = false;
      for (int b0 = wf->band_lo; b0 < wf->band_hi; b0 += per) {
        const int nb = std::min(per, wf->band_hi - b0);
        ChunkJob job{wf.get(), kap, b0, nb, per, ld, hd.nrecl, kp.nplane, half_len, from};
        if (defer) g_pending_chunks.push_back(job);
        else issue_chunk(job);
      }
    }
  }
  // A second wavefunction has been read while chunks of an earlier one are still deferred: this is the pair case,
  // and from here on the copy engine should never idle - issue everything that is pending, in block order.
  {
    bool others = false;
    for (auto& j : g_pending_chunks) others = others || j.wf != wf.get();
    if (others) flush_pending_ingest();
  }
  // Unless the caller promised to keep its buffer alive (async ingest), wait for the copies - but not for
  // any compute - before returning
  if (!(g_async_ingest && src.mem)) CUDA_OK(cudaStreamSynchronize(ring.copy));
  if (stage) cudaFreeHost(stage);
  return wf.release();
}

// Make the main stream wait until the coefficient rows of bands [band_lo, band_hi) of `kap` have landed.
void wait_coeffs(const pawb200_pswf* wf, int kap, int band_lo, int band_hi, cudaStream_t st = nullptr) {
  flush_pending_ingest();             // deferred chunks (of every wavefunction) are issued before anybody waits
  for (auto& c : wf->chunks)
    if (c.kap == kap && c.band_lo < band_hi && c.band_hi > band_lo)
      CUDA_OK(cudaStreamWaitEvent(st ? st : g_stream, c.ready, 0));
}

// Build / refresh the 16-slot interleaved coefficient copy for bands [band_lo, band_hi) of kappa on `st`.
void interleave_rows(pawb200_pswf* wf, int kap, int band_lo, int band_hi, cudaStream_t st) {
  const int h = wf->halves(), half_len = wf->npw_half(kap);
  if (half_len == 0 || band_hi <= band_lo) return;
  const int slot_lo = band_lo * h, slot_hi = band_hi * h;
  dim3 grid((half_len + 127) / 128, ((slot_hi + 15) >> 4) - (slot_lo >> 4));
  interleave_coeff_kernel<<<grid, 256, 0, st>>>(wf->C[kap].as<float2>(), wf->ldc[kap], h, half_len, slot_lo, slot_hi,
                                                wf->Cil[kap].as<float2>(), wf->ldil[kap]);
  count_launch();
  check_launch();
}

// allocate (zeroed, on the main stream) the interleaved copy of kappa
void alloc_interleaved(pawb200_pswf* wf, int kap, cudaStream_t st) {
  const int NK = wf->nkappa();
  if ((int)wf->Cil.size() != NK) { wf->Cil.resize(NK); wf->ldil.assign(NK, 0); }
  const long ldil = ((long)wf->npw_half(kap) + 1) / 2 * 2;
  const size_t bytes = (size_t)((wf->nslot() + 15) / 16) * ldil * 16 * sizeof(float2);
  wf->ldil[kap] = ldil;
  wf->Cil[kap].alloc_x(std::max<size_t>(bytes, 16));
  wf->Cil[kap].zero(std::max<size_t>(bytes, 16), st);
}

// ---- inverse scatter map -------------------------------------------------------------------
DevBuf build_inverse_map(const pawb200_pswf* wf, int kap, const int* fftg, std::vector<int>* fwd = nullptr) {
  HostSection hs_("build_inverse_map");
  const KPointInfo& kp = wf->kp[kap];
  const int npw = wf->npw_half(kap);
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  // host work is O(npw): grid position of every plane wave; the O(N_grid) map itself is filled on the device
  std::vector<int> lin(npw), val(npw);
  int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  for (int w = 0; w < npw; w++) {
    const int g1 = (kp.G[3 * w] + fftg[0]) % fftg[0];
    const int g2 = (kp.G[3 * w + 1] + fftg[1]) % fftg[1];
    const int g3 = (kp.G[3 * w + 2] + fftg[2]) % fftg[2];
    if (g1 < 0 || g2 < 0 || g3 < 0) throw std::runtime_error("FFT grid smaller than the G range");
    lin[w] = (int)(((long)g1 * fftg[1] + g2) * fftg[2] + g3);
    val[w] = kp.pos[w];
    for (int d = 0; d < 3; d++) {
      lo[d] = std::min(lo[d], (int)kp.G[3 * w + d]);
      hi[d] = std::max(hi[d], (int)kp.G[3 * w + d]);
    }
  }
  if (fwd) *fwd = lin;
  // On an aliased grid several plane waves share a position; the reference's loop lets the last one win
  // (linalg.c:31).  Keep only those winners so the device scatter has no write conflicts.
  if (hi[0] - lo[0] >= fftg[0] || hi[1] - lo[1] >= fftg[1] || hi[2] - lo[2] >= fftg[2]) {
    std::unordered_map<int, int> last;
    for (int w = 0; w < npw; w++) last[lin[w]] = w;
    std::vector<int> l2, v2;
    l2.reserve(last.size()); v2.reserve(last.size());
    for (int w = 0; w < npw; w++)
      if (last[lin[w]] == w) { l2.push_back(lin[w]); v2.push_back(val[w]); }
    lin.swap(l2);
    val.swap(v2);
  }
  DevBuf inv((size_t)ngrid * sizeof(int));
  CUDA_OK(cudaMemsetAsync(inv.p, 0xFF, (size_t)ngrid * sizeof(int), g_stream));    // -1 everywhere
  const int n = (int)lin.size();
  if (n) {
    DevBuf dl = upload(lin), dv = upload(val);
    fill_inverse_map_kernel<<<(n + 255) / 256, 256, 0, g_stream>>>(dl.as<int>(), dv.as<int>(), n, inv.as<int>());
    count_launch();
    check_launch();
  }
  return inv;
}

void launch_scatter(const pawb200_pswf* wf, int kap, int slot0, int nslot, const DevBuf& inv,
                    double2* x, const int* fftg) {
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  const double scale = std::pow(determinant3(wf->lattice), -0.5);   // linalg.c:34
  const int h = wf->halves();
  if (slot0 % h) throw std::runtime_error("slot batches must start on a band boundary");
  const int threads = 256;
  long blocks = std::min<long>((ngrid + threads - 1) / threads, (long)g_num_sms * 16);
  wait_coeffs(wf, kap, slot0 / h, (slot0 + nslot + h - 1) / h);
  ScopedStage tm(ST_SCATTER);
  g_boxes_scattered += nslot;
  scatter_pw_kernel<<<(unsigned)blocks, threads, 0, g_stream>>>(
      wf->C[kap].as<float2>(), wf->ldc[kap], slot0 / h, h, wf->npw_half(kap), inv.as<int>(), x, ngrid,
      nslot, scale);
  count_launch();
  check_launch();
}

void launch_fft(double2* x, const int* fftg, int batch, int direction) {
  cufftHandle plan = get_plan(fftg, batch);
  ScopedStage tm(ST_FFT);
  g_boxes_fft += batch;
  CUFFT_OK(cufftExecZ2Z(plan, (cufftDoubleComplex*)x, (cufftDoubleComplex*)x, direction));
}

template <int MT>
void launch_project_mt(const SiteTables& T, const double2* x, long ngrid, int nslot, double2* P,
                       long ldp, int slot0) {
  if (T.by_mt[MT].empty()) return;
  static bool configured = false;
  const size_t smem = sphere_project_smem(MT);
  if (!configured) {
    CUDA_OK(cudaFuncSetAttribute(sphere_project_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    configured = true;
  }
  dim3 grid((nslot + PROJ_NB - 1) / PROJ_NB, (unsigned)T.by_mt[MT].size());
  sphere_project_kernel<MT><<<grid, 128, smem, g_stream>>>(
      T.sites.as<SiteDev>(), T.by_mt_dev[MT].as<int>(), T.idx.as<int>(), T.tablek.as<double2>(), x, ngrid,
      nslot, P, ldp, slot0);
  count_launch();
  check_launch();
}

void launch_project(const SiteTables& T, const double2* x, long ngrid, int nslot, double2* P, long ldp,
                    int slot0) {
  if (T.nsites == 0 || T.total_pts == 0) return;
  ScopedStage tm(ST_PROJECT);
  g_slots_projected += nslot;
  for (auto& sd : T.host) g_sphere_samples += (long long)nslot * sd.npts;
  launch_project_mt<1>(T, x, ngrid, nslot, P, ldp, slot0);
  launch_project_mt<2>(T, x, ngrid, nslot, P, ldp, slot0);
  launch_project_mt<3>(T, x, ngrid, nslot, P, ldp, slot0);
}

// Per-k phase data of a table set.  interleaved: the real-table kernel needs only dv e^{i k.path} per sphere point;
// otherwise the planar kernel's complex per-k table Tk = conj(T) dv e^{i k.path} is built.
void make_phase_table(SiteTables& T, const pawb200_pswf* wf, int kap, const int* fftg, bool interleaved) {
  if (T.nsites == 0 || T.total_pts == 0) return;
  double kc[3] = {wf->kp[kap].k[0], wf->kp[kap].k[1], wf->kp[kap].k[2]};
  frac_to_cart(kc, wf->reclattice);                                           // projector.c:233-237
  const double dv = determinant3(wf->lattice) / fftg[0] / fftg[1] / fftg[2];   // projector.c:229
  ScopedStage tm(ST_TABLE);
  if (interleaved && T.ureal.p) {
    T.phk.ensure((size_t)T.total_pts * sizeof(double2));
    const unsigned blocks = (unsigned)std::min<long>((T.total_pts + 255) / 256, (long)g_num_sms * 8);
    phase_points_kernel<<<blocks, 256, 0, g_stream>>>(T.path.as<double>(), T.total_pts, T.total_pts,
                                                      T.phk.as<double2>(), kc[0], kc[1], kc[2], dv);
  } else {
    T.tablek.ensure(std::max<size_t>(1, T.total_tab) * sizeof(double2));
    int maxpts = 0;
    for (auto& sd : T.host) maxpts = std::max(maxpts, sd.npts_pad);
    dim3 grid((maxpts + 255) / 256, T.nsites);
    phase_table_kernel<<<grid, 256, 0, g_stream>>>(T.sites.as<SiteDev>(), T.path.as<double>(), T.total_pts,
                                                   T.table.as<double2>(), T.tablek.as<double2>(), kc[0],
                                                   kc[1], kc[2], dv);
  }
  count_launch();
  check_launch();
}

DevBuf g_grid;   // FFT box batch, reused across calls

// ---- pruned band-interleaved FFT (fft3d.cuh) -----------------------------------------------------------
struct PrunedPlan {
  bool ok = false;
  FftGeom g;
  int max_plane_cols = 0;
  DevBuf col_start, col_cnt, zpos, col_run, plane_run, ysrc, xsrc, tw[3];
};

bool factor_pair(int n, int& r1, int& r2) {
  static const int ok[] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 15, 16, 18, 20};
  int best = 1 << 30;
  bool found = false;
  for (int a : ok)
    for (int b : ok)
      if (a * b == n && std::max(a, b) < best) {
        best = std::max(a, b);
        r1 = a;
        r2 = b;
        found = true;
      }
  return found;
}

// n = a * b * c with radices <= 10, as balanced as possible (smallest largest factor); the largest factor goes first
// (phase 1 has the most tasks per thread otherwise)
bool factor_triple(int n, int& r1, int& r2, int& r3) {
  if (n > 420) return false;            // two exchange buffers of n x 256 B must fit the shared memory of an SM
  int best = 1 << 30;
  bool found = false;
  for (int a = 2; a <= 10; a++)
    for (int b = 2; b <= a; b++)
      for (int c = 2; c <= b; c++)
        if (a * b * c == n && a < best) {
          best = a;
          r1 = a; r2 = b; r3 = c;
          found = true;
        }
  return found;
}

std::shared_ptr<PrunedPlan> build_pruned_plan_kp(const KPointInfo& kp, int npw, const int* fftg) {
  auto P = std::make_shared<PrunedPlan>();
  if (getenv("PAWB200_FFT") && std::string(getenv("PAWB200_FFT")) == "cufft") return P;
  FftGeom& g = P->g;
  g.n1 = fftg[0]; g.n2 = fftg[1]; g.n3 = fftg[2];
  g.pf = getenv("PAWB200_FFT_PF") ? atoi(getenv("PAWB200_FFT_PF")) : 1;
  for (int d = 0; d < 3; d++) {
    g.r3[d] = 1;
    if (factor_pair(fftg[d], g.r1[d], g.r2[d])) continue;
    if (getenv("PAWB200_FFT3") && atoi(getenv("PAWB200_FFT3")) == 0) return P;
    if (!factor_triple(fftg[d], g.r1[d], g.r2[d], g.r3[d])) return P;
    g.three = 1;
  }
  if (npw == 0) return P;
  std::vector<int> col_start, col_cnt, col_ypos, zpos(npw), plane_col0, plane_ncol, plane_xpos;
  int last1 = -1, last2 = -1, lastz = -1;
  for (int j = 0; j < npw; j++) {
    const int w = kp.perm[j];
    const int a = ((kp.G[3 * w] % fftg[0]) + fftg[0]) % fftg[0];
    const int b = ((kp.G[3 * w + 1] % fftg[1]) + fftg[1]) % fftg[1];
    const int c = ((kp.G[3 * w + 2] % fftg[2]) + fftg[2]) % fftg[2];
    if (std::abs(kp.G[3 * w]) >= fftg[0] || std::abs(kp.G[3 * w + 1]) >= fftg[1] || std::abs(kp.G[3 * w + 2]) >= fftg[2])
      return P;
    if (a != last1) {
      if (a < last1) return P;                 // not in box order (aliased grid): use the generic path
      plane_col0.push_back((int)col_start.size());
      plane_ncol.push_back(0);
      plane_xpos.push_back(a);
      last1 = a;
      last2 = -1;
    }
    if (b != last2) {
      if (b < last2) return P;
      col_start.push_back(j);
      col_cnt.push_back(0);
      col_ypos.push_back(b);
      plane_ncol.back()++;
      last2 = b;
      lastz = -1;
    }
    if (c <= lastz) return P;                  // duplicate / unordered positions
    lastz = c;
    zpos[j] = c;
    col_cnt.back()++;
  }
  g.ncol = (int)col_start.size();
  g.nplane = (int)plane_col0.size();
  std::vector<int> ysrc((size_t)g.nplane * fftg[1], -1), xsrc(fftg[0], -1);
  for (int p = 0; p < g.nplane; p++) {
    xsrc[plane_xpos[p]] = p;
    for (int c = plane_col0[p]; c < plane_col0[p] + plane_ncol[p]; c++) ysrc[(size_t)p * fftg[1] + col_ypos[c]] = c;
  }
  // cyclic z-run of each column (always one run for a cutoff sphere; checked, with a staged fallback)
  std::vector<int4> col_run(g.ncol);
  bool runs_ok = getenv("PAWB200_FFT_STAGED_Z") == nullptr;
  for (int c = 0; c < g.ncol && runs_ok; c++) {
    const int s0 = col_start[c], cnt = col_cnt[c];
    int gap = -1;
    for (int j = 0; j + 1 < cnt; j++)
      if (zpos[s0 + j + 1] != zpos[s0 + j] + 1) {
        if (gap >= 0) runs_ok = false;
        gap = j;
      }
    if (gap >= 0 && !(zpos[s0] == 0 && zpos[s0 + cnt - 1] == fftg[2] - 1)) runs_ok = false;
    // sorted list = [wrapped-around tail of the run (z = 0..), head of the run (z = zlo..n3-1)]
    col_run[c] = gap < 0 ? make_int4(s0, cnt, zpos[s0], cnt) : make_int4(s0, cnt, zpos[s0 + gap + 1], cnt - (gap + 1));
  }
  if (runs_ok) P->col_run = upload(col_run);
  g.col_run = runs_ok ? P->col_run.as<int4>() : nullptr;
  // cyclic y-run of the active columns of each x-plane (same encoding; consumed by the fused pass Y+X)
  std::vector<int4> plane_run(g.nplane);
  bool planes_ok = true;
  for (int p = 0; p < g.nplane && planes_ok; p++) {
    const int c0 = plane_col0[p], cnt = plane_ncol[p];
    int gap = -1;
    for (int j = 0; j + 1 < cnt; j++)
      if (col_ypos[c0 + j + 1] != col_ypos[c0 + j] + 1) {
        if (gap >= 0) planes_ok = false;
        gap = j;
      }
    if (gap >= 0 && !(col_ypos[c0] == 0 && col_ypos[c0 + cnt - 1] == fftg[1] - 1)) planes_ok = false;
    plane_run[p] = gap < 0 ? make_int4(c0, cnt, col_ypos[c0], cnt)
                           : make_int4(c0, cnt, col_ypos[c0 + gap + 1], cnt - (gap + 1));
  }
  for (int p = 0; p < g.nplane; p++) P->max_plane_cols = std::max(P->max_plane_cols, plane_ncol[p]);
  if (planes_ok) P->plane_run = upload(plane_run);
  g.plane_run = planes_ok ? P->plane_run.as<int4>() : nullptr;
  P->col_start = upload(col_start); P->col_cnt = upload(col_cnt); P->zpos = upload(zpos);
  P->ysrc = upload(ysrc); P->xsrc = upload(xsrc);
  g.col_start = P->col_start.as<int>(); g.col_cnt = P->col_cnt.as<int>(); g.zpos = P->zpos.as<int>();
  g.ysrc = P->ysrc.as<int>(); g.xsrc = P->xsrc.as<int>();
  for (int d = 0; d < 3; d++) {
    std::vector<double2> t(fftg[d]);
    for (int m = 0; m < fftg[d]; m++) {
      const long double a = 2.0L * 3.141592653589793238462643383279502884L * m / fftg[d];
      t[m] = make_double2((double)cosl(a), (double)sinl(a));
    }
    P->tw[d] = upload(t);
    g.tw[d] = P->tw[d].as<double2>();
  }
  init_small_twiddles();
  if (g.three && !g.col_run) return P;      // the three-factor pass Z has no staged variant
  P->ok = true;
  return P;
}

std::shared_ptr<PrunedPlan> build_pruned_plan(const pawb200_pswf* wf, int kap, const int* fftg) {
  return build_pruned_plan_kp(wf->kp[kap], wf->npw_half(kap), fftg);
}

std::shared_ptr<PrunedPlan> get_pruned_plan(pawb200_pswf* wf, int kap, const int* fftg) {
  const int NK = wf->nkappa();
  if ((int)wf->fft_plans.size() != NK || wf->plan_fftg[0] != fftg[0] || wf->plan_fftg[1] != fftg[1] ||
      wf->plan_fftg[2] != fftg[2]) {
    wf->fft_plans.assign(NK, nullptr);
    for (int d = 0; d < 3; d++) wf->plan_fftg[d] = fftg[d];
  }
  if (!wf->fft_plans[kap]) wf->fft_plans[kap] = build_pruned_plan(wf, kap, fftg);
  return wf->fft_plans[kap];
}

DevBuf g_fft_t1, g_fft_t2, g_fft_flags;

size_t fused_l2_budget() {
  // ring of the fused pass Y+X: must stay L2-resident next to the streams that pass through (126 MB L2)
  if (const char* e = getenv("PAWB200_FFT_RING_BYTES")) return (size_t)atoll(e);
  const char* f = getenv("PAWB200_FFT_FUSED");
  return (size_t)((f && atoi(f) == 2) ? 60 : 40) << 20;      // phased variant: 3 slots, two of them live
}

// Inverse transform of slots [slot0, slot0 + nslot) of kappa into X (interleaved groups of FFT_B slots).
void pruned_fft(const pawb200_pswf* wf, int kap, const PrunedPlan& P, int slot0, int nslot, double2* X) {
  const FftGeom& g = P.g;
  const int ngroups = (nslot + FFT_B - 1) / FFT_B;
  const long ngrid = (long)g.n1 * g.n2 * g.n3;
  const size_t t1_grp = (size_t)g.ncol * g.n3 * FFT_B * sizeof(double2);
  const size_t t2_grp = (size_t)g.nplane * g.n2 * g.n3 * FFT_B * sizeof(double2);
  // groups per launch: measured on config 2 (fft ms per step, stand-alone passes) 2: 11.8, 4: 10.8, 8: 10.3, 16: 10.1,
  // all 38: 10.0.  8 groups = 128 bands is also one ingest chunk, so a launch never waits for more than the chunk
  // it needs.
  int gc = 8;
  if (const char* e = getenv("PAWB200_FFT_GROUPS")) gc = std::max(1, atoi(e));
  gc = std::min(gc, ngroups);
  const YxConfig yx = plan_fused_yx(g, g_num_sms, fused_l2_budget());
  g_fft_t1.ensure(t1_grp * gc);
  if (yx.ok) {
    g_fft_t2.ensure(yx.ring_bytes);
    g_fft_flags.ensure(yx.flag_words(gc) * sizeof(unsigned));
  } else {
    g_fft_t2.ensure(t2_grp * gc);
  }
  const double scale = std::pow(determinant3(wf->lattice), -0.5);
  const int h = wf->halves();
  FftInput in;
  in.Cil = (!wf->Cil.empty() && wf->Cil[kap].p) ? wf->Cil[kap].as<float2>() : nullptr;
  in.ldil = in.Cil ? wf->ldil[kap] : 0;
  in.C = wf->C[kap].as<float2>();
  in.ldc = wf->ldc[kap];
  in.halves = h;
  in.half_len = wf->npw_half(kap);
  FftWork w;
  w.T1 = g_fft_t1.as<double2>();
  w.T2 = g_fft_t2.as<double2>();
  w.flags = yx.ok ? g_fft_flags.as<unsigned>() : nullptr;
  ScopedStage tm(ST_FFT);
  g_boxes_fft += nslot;
  for (int g0 = 0; g0 < ngroups; g0 += gc) {
    const int ng = std::min(gc, ngroups - g0);
    const int s0 = slot0 + g0 * FFT_B;
    const int ns = std::min(nslot - g0 * FFT_B, ng * FFT_B);
    wait_coeffs(wf, kap, s0 / h, (s0 + ns + h - 1) / h);
    count_launch(launch_pruned_passes(g, in, s0, ns, ng, scale, w, &yx, X + (long)g0 * ngrid * FFT_B, g_num_sms,
                                      g_stream, P.max_plane_cols));
    trace_mark("fft done slots " + std::to_string(s0) + "+" + std::to_string(ns), g_stream);
  }
  check_launch();
}

template <int MT>
void launch_project_il_mt(const SiteTables& T, const double2* X, long ngrid, int nslot, double2* P, long ldp,
                          int slot0) {
  if (T.by_mt[MT].empty()) return;
  int maxpts = 0;
  for (int s : T.by_mt[MT]) maxpts = std::max(maxpts, T.host[s].npts_pad);
  const int idx_cap = std::min(maxpts, 16384);
  dim3 grid((nslot + PROJ_NB - 1) / PROJ_NB, (unsigned)T.by_mt[MT].size());
  if (T.ureal.p && T.phk.p) {
    // real tables + phase on the samples: 2 DMMA per k-step and channel tile instead of 4
    const size_t smem = sphere_project_real_smem<MT>();
    static size_t configured_r = 0;
    if (smem > configured_r) {
      CUDA_OK(cudaFuncSetAttribute(sphere_project_real_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured_r = smem;
    }
    sphere_project_real_kernel<MT><<<grid, PROJ_THREADS, smem, g_stream>>>(
        T.sites.as<SiteDev>(), T.by_mt_dev[MT].as<int>(), T.idx.as<int>(), T.ureal.as<double>(),
        T.phk.as<double2>(), T.chan_m.as<int>(), X, ngrid, nslot, (nslot + FFT_B - 1) / FFT_B, P, ldp, slot0, idx_cap);
    count_launch();
    check_launch();
    return;
  }
  const size_t smem = sphere_project_smem(MT) + (size_t)idx_cap * sizeof(int);
  static size_t configured = 0;
  if (smem > configured) {
    CUDA_OK(cudaFuncSetAttribute(sphere_project_il_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  sphere_project_il_kernel<MT><<<grid, 128, smem, g_stream>>>(
      T.sites.as<SiteDev>(), T.by_mt_dev[MT].as<int>(), T.idx.as<int>(), T.tablek.as<double2>(), X, ngrid, nslot,
      (nslot + FFT_B - 1) / FFT_B, P, ldp, slot0, idx_cap);
  count_launch();
  check_launch();
}

void launch_project_il(const SiteTables& T, const double2* X, long ngrid, int nslot, double2* P, long ldp,
                       int slot0) {
  if (T.nsites == 0 || T.total_pts == 0) return;
  ScopedStage tm(ST_PROJECT);
  g_slots_projected += nslot;
  for (auto& sd : T.host) g_sphere_samples += (long long)nslot * sd.npts;
  launch_project_il_mt<1>(T, X, ngrid, nslot, P, ldp, slot0);
  launch_project_il_mt<2>(T, X, ngrid, nslot, P, ldp, slot0);
  launch_project_il_mt<3>(T, X, ngrid, nslot, P, ldp, slot0);
  trace_mark("project_il done slots " + std::to_string(slot0) + "+" + std::to_string(nslot), g_stream);
}

size_t keep_boxes_budget() {
  const char* e = getenv("PAWB200_KEEP_BOXES_BYTES");
  if (e) return (size_t)atoll(e);
  // per wavefunction: a quarter of the device (45 GB on a 180 GB B200), so a basis/wf pair leaves half of HBM
  // for coefficients, tables and scratch
  static size_t quarter = 0;
  if (!quarter) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = (size_t)128 << 30;
    quarter = total_b / 4;
  }
  return quarter;
}

// setup_projections, step 1: transform every band of every resident (k,spin) block into resident interleaved
// boxes.  Nothing here depends on the atomic sites, so the kernels are queued before the host builds the sphere
// geometry and tables (which then overlaps the transforms).  Returns false when the boxes do not fit the
// budget or the grid needs the generic path; project_all_bands then transforms batch by batch.
bool prefft_all_bands(pawb200_pswf* wf, const int* fftg) {
  const int NK = wf->nkappa(), nslot = wf->slot_own();      // boxes[kap] holds the slots of the own band block
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  const int ngroups_all = (nslot + FFT_B - 1) / FFT_B;
  const size_t il_bytes = (size_t)ngroups_all * FFT_B * ngrid * sizeof(double2);
  int nres = 0;
  for (int kap = 0; kap < NK; kap++) nres += wf->resident[kap] ? 1 : 0;
  wf->boxes.clear();
  if (il_bytes * (size_t)std::max(nres, 1) > keep_boxes_budget()) return false;
  for (int kap = 0; kap < NK; kap++)
    if (wf->resident[kap] && !get_pruned_plan(wf, kap, fftg)->ok) return false;
  wf->boxes.resize(NK);
  for (int d = 0; d < 3; d++) wf->boxes_fftg[d] = fftg[d];
  wf->boxes_interleaved = true;
  for (int kap = 0; kap < NK; kap++) {
    if (!wf->resident[kap]) continue;
    wf->boxes[kap].alloc(il_bytes);
    if (nslot > 0)
      pruned_fft(wf, kap, *get_pruned_plan(wf, kap, fftg), wf->slot_lo(), nslot, wf->boxes[kap].as<double2>());
  }
  return true;
}

// All bands of all resident (k,spin) blocks of `wf` -> <table|psi~>, written to out[kap] [nslot][ld].
// main_pass: this is setup_projections (boxes may be kept); otherwise kept boxes are reused when present.
// T2 / out2 / ld2 (optional): a second table set projected from the same boxes in the same pass.
void project_all_bands(pawb200_pswf* wf, SiteTables& T, const int* fftg, std::vector<DevBuf>& out,
                       long& ld, bool main_pass, SiteTables* T2 = nullptr, std::vector<DevBuf>* out2 = nullptr,
                       long* ld2 = nullptr) {
  HostSection hs_("project_all_bands");
  const int NK = wf->nkappa();
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  ld = ((long)std::max(T.nproj, 1) + 7) / 8 * 8;
  out.clear();
  out.resize(NK);
  const bool two = T2 && out2 && ld2 && T2->nsites > 0;
  if (T2 && out2 && ld2) {
    *ld2 = ((long)std::max(T2->nproj, 1) + 7) / 8 * 8;
    out2->clear();
    out2->resize(NK);
  }
  const int nslot = wf->slot_own(), slo = wf->slot_lo();   // own band block (all bands unless band-sharded)
  const size_t out_rows = (size_t)wf->band_rows * wf->halves();
  long batch = (long)(fft_budget_bytes() / (sizeof(double2) * ngrid));
  batch = std::max<long>(batch, wf->halves());
  if (batch >= 32) batch = batch / 32 * 32;
  batch -= batch % wf->halves();
  batch = std::max<long>(std::min<long>(batch, nslot), 1);
  int nres = 0;
  for (int kap = 0; kap < NK; kap++) nres += wf->resident[kap] ? 1 : 0;
  const size_t box_bytes = (size_t)nslot * ngrid * sizeof(double2);
  const bool same_grid = wf->boxes_fftg[0] == fftg[0] && wf->boxes_fftg[1] == fftg[1] && wf->boxes_fftg[2] == fftg[2];
  bool keep = false;
  if (main_pass) {
    wf->boxes.clear();
    keep = box_bytes * (size_t)std::max(nres, 1) <= keep_boxes_budget();
    if (keep) {
      wf->boxes.resize(NK);
      for (int d = 0; d < 3; d++) wf->boxes_fftg[d] = fftg[d];
    }
  }
  const int ngroups_all = (nslot + FFT_B - 1) / FFT_B;
  const size_t il_bytes = (size_t)ngroups_all * FFT_B * ngrid * sizeof(double2);
  for (int kap = 0; kap < NK; kap++) {
    if (!wf->resident[kap]) continue;
    out[kap].alloc(out_rows * ld * sizeof(double2));
    out[kap].zero();
    if (T2 && out2 && ld2) {
      (*out2)[kap].alloc(out_rows * *ld2 * sizeof(double2));
      (*out2)[kap].zero();
    }
    if ((T.nsites == 0 && !two) || nslot == 0) continue;
    const bool reuse = !main_pass && same_grid && (int)wf->boxes.size() == NK && wf->boxes[kap].p;
    std::shared_ptr<PrunedPlan> plan = reuse ? nullptr : get_pruned_plan(wf, kap, fftg);
    make_phase_table(T, wf, kap, fftg, reuse ? wf->boxes_interleaved : plan->ok);
    if (two) make_phase_table(*T2, wf, kap, fftg, reuse ? wf->boxes_interleaved : plan->ok);
    auto second_il = [&](const double2* x, int nb, int slot) {
      if (two) launch_project_il(*T2, x, ngrid, nb, (*out2)[kap].as<double2>(), *ld2, slot);
    };
    auto second_planar = [&](const double2* x, int nb, int slot) {
      if (two) launch_project(*T2, x, ngrid, nb, (*out2)[kap].as<double2>(), *ld2, slot);
    };
    if (reuse) {
      for (int s0 = 0; s0 < nslot; s0 += 2048) {
        const int nb = std::min(2048, nslot - s0);
        const double2* xb = wf->boxes[kap].as<double2>() + (long)s0 * ngrid;
        if (wf->boxes_interleaved) {
          launch_project_il(T, xb, ngrid, nb, out[kap].as<double2>(), ld, slo + s0);
          second_il(xb, nb, slo + s0);
        } else {
          launch_project(T, xb, ngrid, nb, out[kap].as<double2>(), ld, slo + s0);
          second_planar(xb, nb, slo + s0);
        }
      }
      continue;
    }
    if (plan->ok) {
      // pruned band-interleaved FFT: chunks of whole 32-slot CTAs
      long chunk = std::max<long>(32, batch / 32 * 32);
      double2* base;
      if (keep) {
        wf->boxes[kap].alloc(il_bytes);
        wf->boxes_interleaved = true;
        base = wf->boxes[kap].as<double2>();
      } else {
        g_grid.ensure((size_t)chunk * ngrid * sizeof(double2));
        base = g_grid.as<double2>();
      }
      for (int s0 = 0; s0 < nslot; s0 += (int)chunk) {
        const int nb = (int)std::min<long>(chunk, nslot - s0);
        double2* x = keep ? base + (long)s0 * ngrid : base;
        pruned_fft(wf, kap, *plan, slo + s0, nb, x);
        launch_project_il(T, x, ngrid, nb, out[kap].as<double2>(), ld, slo + s0);
        second_il(x, nb, slo + s0);
      }
      continue;
    }
    DevBuf inv = build_inverse_map(wf, kap, fftg);
    double2* base;
    if (keep) {
      wf->boxes[kap].alloc(box_bytes);
      wf->boxes_interleaved = false;
      base = wf->boxes[kap].as<double2>();
    } else {
      g_grid.ensure((size_t)batch * ngrid * sizeof(double2));
      base = g_grid.as<double2>();
    }
    for (int s0 = 0; s0 < nslot; s0 += (int)batch) {
      const int nb = (int)std::min<long>(batch, nslot - s0);
      double2* x = keep ? base + (long)s0 * ngrid : base;
      launch_scatter(wf, kap, slo + s0, nb, inv, x, fftg);
      launch_fft(x, fftg, nb, CUFFT_INVERSE);
      launch_project(T, x, ngrid, nb, out[kap].as<double2>(), ld, slo + s0);
      second_planar(x, nb, slo + s0);
    }
  }
}

// Projections are valid on return: a wavefunction whose setup_projections deferred the work (lazy_proj) is
// transformed and projected now - together with a second table set when the caller has one for the same bands.
void require_projections(pawb200_pswf* wf, SiteTables* T2 = nullptr, std::vector<DevBuf>* out2 = nullptr,
                         long* ld2 = nullptr) {
  if (!wf || !wf->has_projections) throw std::runtime_error("setup_projections has not been run");
  if (wf->lazy_proj) {
    wf->lazy_proj = false;
    project_all_bands(wf, *wf->proj_sites, wf->fftg, wf->P, wf->ldp, true, T2, out2, ld2);
  } else if (T2 && out2 && ld2) {
    // boxes (if kept) are reused; otherwise the bands are transformed again for the second table set alone
    project_all_bands(wf, *T2, wf->fftg, *out2, *ld2, false);
  }
}


// ---- GEMM driver ------------------------------------------------------------------------------
DevBuf g_zg_ws;
// The pseudo-overlap GEMM only needs the plane-wave coefficients, and it is tensor-pipe bound while the FFT /
// projection pipeline is HBM bound: it runs on its own high-priority stream with one persistent CTA per SM and
// can overlap the transforms queued on the main stream (PAWB200_GEMM_OVERLAP=1; off by default).
cudaStream_t g_stream2 = nullptr;
struct RawBuf {            // plain cudaMalloc buffer (not from the stream-ordered pool: used across streams)
  void* p = nullptr;
  size_t bytes = 0;
  void ensure(size_t n) {
    if (n <= bytes) return;
    if (p) {   // growing: nothing may still be using the old block
      CUDA_OK(cudaDeviceSynchronize());
      cudaFree(p);
    }
    CUDA_OK(cudaMalloc(&p, n));
    bytes = n;
  }
};
RawBuf g_zg_ws2, g_pblk[2];
cudaEvent_t g_pblk_free[2] = {nullptr, nullptr}, g_pblk_done[2] = {nullptr, nullptr};
bool gemm_overlap_enabled() {
  // measured on B200 (config 2): concurrent execution stretches both sides and nets nothing, so it is opt-in
  static const bool on = getenv("PAWB200_GEMM_OVERLAP") && atoi(getenv("PAWB200_GEMM_OVERLAP")) != 0;
  return on;
}
cudaStream_t gemm_stream() {
  if (!g_stream2) {
    int lo = 0, hi = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_OK(cudaStreamCreateWithPriority(&g_stream2, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 2; i++) {
      CUDA_OK(cudaEventCreateWithFlags(&g_pblk_free[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&g_pblk_done[i], cudaEventDisableTiming));
    }
  }
  return g_stream2;
}

template <typename T, bool K3M>
void run_zgemm_variant(const T* A, long lda, const T* B, long ldb, int M, int N, long Kpad, double2* out,
                       long ldo, bool accumulate, int stage, cudaStream_t st, bool side_stream) {
  static bool configured = false;
  constexpr size_t smem = zgemm_smem_bytes<T, K3M>();
  constexpr int BN = ZgShape<K3M>::BN;
  if (!configured) {
    CUDA_OK(cudaFuncSetAttribute(zgemm_abh_kernel<T, K3M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (M == 0 || N == 0) return;
  ZgPlan plan;
  plan.M = M; plan.N = N; plan.bn = BN;
  plan.tiles_m = (M + ZG_BM - 1) / ZG_BM;
  plan.tiles_n = (N + BN - 1) / BN;
  if (Kpad % ZgTraits<T>::KT) throw std::runtime_error("GEMM K dimension is not padded");
  plan.kiters = Kpad / ZgTraits<T>::KT;
  if (plan.kiters == 0) {
    if (!accumulate) CUDA_OK(cudaMemset2DAsync(out, ldo * sizeof(double2), 0, N * sizeof(double2), M, st));
    return;
  }
  plan.total = (long)plan.tiles_m * plan.tiles_n * plan.kiters;
  // side stream: one CTA per SM so the transforms on the main stream keep most of the register file / smem
  plan.G = (int)std::min<long>(2L * g_num_sms, plan.total);
  const size_t ws_bytes = (size_t)plan.G * 2 * ZG_BM * BN * sizeof(double2);
  double2* ws;
  if (side_stream) {
    g_zg_ws2.ensure(ws_bytes);
    ws = (double2*)g_zg_ws2.p;
  } else {
    g_zg_ws.ensure(ws_bytes);
    ws = g_zg_ws.as<double2>();
  }
  trace_mark(std::string(side_stream ? "side " : "main ") + "zgemm start M=" + std::to_string(M), st);
  ScopedStage tm(stage, st);
  zgemm_abh_kernel<T, K3M><<<plan.G, ZG_THREADS, smem, st>>>(A, lda, B, ldb, plan, out, ldo, accumulate ? 1 : 0, ws);
  count_launch();
  check_launch();
  zgemm_fixup_kernel<<<plan.tiles_m * plan.tiles_n, 256, 0, st>>>(plan, ws, out, ldo, accumulate ? 1 : 0);
  count_launch();
  check_launch();
  trace_mark(std::string(side_stream ? "side " : "main ") + "zgemm done M=" + std::to_string(M) + " K=" + std::to_string(Kpad), st);
}

// PAWB200_GEMM_4M=1 selects the 4-real-product variant (default: 3M, 25 % fewer DMMA instructions)
template <typename T>
void run_zgemm(const T* A, long lda, const T* B, long ldb, int M, int N, long Kpad, double2* out,
               long ldo, bool accumulate, int stage, cudaStream_t st = nullptr, bool side_stream = false) {
  static const bool use4m = getenv("PAWB200_GEMM_4M") != nullptr;
  if (!st) st = g_stream;
  if (use4m)
    run_zgemm_variant<T, false>(A, lda, B, ldb, M, N, Kpad, out, ldo, accumulate, stage, st, side_stream);
  else
    run_zgemm_variant<T, true>(A, lda, B, ldb, M, N, Kpad, out, ldo, accumulate, stage, st, side_stream);
}

int flipped(const pawb200_pswf* wf, int kap, int flip) {
  if (wf->nspin == 2 && flip) return kap < wf->nwk ? kap + wf->nwk : kap - wf->nwk;
  return kap;
}

void check_pair(const pawb200_pswf* S, const pawb200_pswf* R) {
  if (!S || !R) throw std::runtime_error("NULL wavefunction pointer");
  if (S->nwk != R->nwk || S->nspin != R->nspin)
    throw std::runtime_error("wavefunctions have different k-point / spin counts (the reference reads out "
                             "of bounds here, SURVEY 8b); refusing");
  if (S->ncl || R->ncl) throw std::runtime_error("band-pair overlaps are not defined for noncollinear input "
                                                 "(projector.py:74-75)");
}

// true while some ingest chunk of (wf, kap) is still being copied to the device
bool coeffs_in_flight(const pawb200_pswf* wf, int kap) {
  static const bool force = getenv("PAWB200_GEMM_CHUNKED") != nullptr;   // tests: take the chunked path always
  if (force) return true;
  flush_pending_ingest();
  for (auto& c : wf->chunks)
    if (c.kap == kap && cudaEventQuery(c.ready) == cudaErrorNotReady) return true;
  return false;
}

// pseudo overlap block for one kappa into dev [nbS][nbR]
void pseudo_block(pawb200_pswf* S, pawb200_pswf* R, int kap, int flip, double2* out, long ldo,
                  cudaStream_t st = nullptr, bool side_stream = false) {
  const int kr = flipped(R, kap, flip);
  if (!S->resident[kap] || !R->resident[kr]) throw std::runtime_error("(k,spin) block not resident on this rank");
  // the reference takes num_waves from wf_ref->kpts[kpt_num] (pseudoprojector.c:84) for both vectors
  if (S->kp[kap].nplane != R->kp[kr].nplane)
    throw std::runtime_error("plane-wave bases differ between the two wavefunctions at kappa " + std::to_string(kap));
  wait_coeffs(R, kr, 0, R->nband, st);
  if (side_stream && coeffs_in_flight(S, kap)) {
    // rows of S are multiplied ingest chunk by ingest chunk as their H2D copies land, so the GEMM runs under
    // the rest of the transfer instead of after it
    cudaStream_t s2 = st ? st : g_stream;
    for (auto& c : S->chunks) {
      if (c.kap != kap) continue;
      CUDA_OK(cudaStreamWaitEvent(s2, c.ready, 0));
      run_zgemm<float2>(S->C[kap].as<float2>() + (long)c.band_lo * S->ldc[kap], S->ldc[kap], R->C[kr].as<float2>(),
                        R->ldc[kr], c.band_hi - c.band_lo, R->nband, S->ldc[kap], out + (long)c.band_lo * ldo, ldo,
                        false, ST_GEMM_PS, st, side_stream);
    }
    return;
  }
  wait_coeffs(S, kap, 0, S->nband, st);
  // rows of the result = the wf bands this process owns (all of them unless band-sharded)
  run_zgemm<float2>(S->C[kap].as<float2>() + (long)S->band_lo * S->ldc[kap], S->ldc[kap], R->C[kr].as<float2>(),
                    R->ldc[kr], S->band_hi - S->band_lo, R->nband, S->ldc[kap], out + (long)S->band_lo * ldo, ldo, false,
                    ST_GEMM_PS, st, side_stream);
}

struct SiteLists {
  std::vector<int> M_R, M_S, N_R, N_S, N_RS_R, N_RS_S;
  size_t hash() const {
    size_t h = 1469598103934665603ull;
    for (auto* v : {&M_R, &M_S, &N_R, &N_S, &N_RS_R, &N_RS_S}) {
      int n = (int)v->size();
      h = hash_bytes(&n, sizeof(n), h);
      if (n) h = hash_bytes(v->data(), n * sizeof(int), h);
    }
    return h;
  }
};

SiteLists make_lists(int num_M, int num_N_R, int num_N_S, int num_N_RS, const int* M_R, const int* M_S,
                     const int* N_R, const int* N_S, const int* N_RS_R, const int* N_RS_S) {
  SiteLists L;
  auto cp = [](std::vector<int>& d, const int* s, int n) { if (n > 0 && s) d.assign(s, s + n); };
  cp(L.M_R, M_R, num_M); cp(L.M_S, M_S, num_M);
  cp(L.N_R, N_R, num_N_R); cp(L.N_S, N_S, num_N_S);
  cp(L.N_RS_R, N_RS_R, num_N_RS); cp(L.N_RS_S, N_RS_S, num_N_RS);
  return L;
}

// Augmentation operands (see kernels.cuh block_apply_kernel) and the c128 GEMM, for one kappa.
struct AugPlan {
  std::vector<BlockOp> opsS, opsR_P, opsR_W, opsS_W;   // ops reading P_S / P_R / W_R / W_S
  std::vector<cdouble> mats;                           // k-independent matrices (D blocks)
  std::vector<std::pair<int, long>> rs_mats;           // (pair index, offset) of k-dependent blocks
  long K = 0, Kpad = 0;
};

AugPlan plan_aug(const pawb200_pswf* S, const pawb200_pswf* R, const SiteLists& L, bool recip = false) {
  HostSection hs_("plan_aug");
  if (!S->has_projections || !R->has_projections)
    throw std::runtime_error("setup_projections has not been run on both wavefunctions");
  require_projections(const_cast<pawb200_pswf*>(S));
  require_projections(const_cast<pawb200_pswf*>(R));
  const auto& elsR = R->pps->list.el;
  const auto& elsS = S->pps->list.el;
  AugPlan A;
  long col = 0;
  auto siteR = [&](int s) -> const SiteDev& {
    if (s < 0 || s >= R->num_sites) throw std::runtime_error("site index out of range (basis)");
    return R->proj_sites->host[s];
  };
  auto siteS = [&](int s) -> const SiteDev& {
    if (s < 0 || s >= S->num_sites) throw std::runtime_error("site index out of range (wf)");
    return S->proj_sites->host[s];
  };
  // O_M  (projector.c:890-910)
  for (size_t q = 0; q < L.M_R.size(); q++) {
    const SiteDev &r = siteR(L.M_R[q]), &s = siteS(L.M_S[q]);
    const Element& pp = elsR[R->labels[L.M_R[q]]];
    const Element& ps = elsS[S->labels[L.M_S[q]]];
    const long off = (long)A.mats.size();
    for (int i = 0; i < r.nlm; i++)
      for (int j = 0; j < s.nlm; j++) {
        const Channel &ci = pp.chan[i], &cj = ps.chan[j];
        double v = 0;
        if (ci.l == cj.l && ci.m == cj.m) {
          if (cj.n >= pp.num_projs) throw std::runtime_error("matched sites carry different PAW datasets");
          v = pp.aeov[ci.n * pp.num_projs + cj.n] - pp.psov[ci.n * pp.num_projs + cj.n];
        }
        A.mats.push_back(cdouble(v, 0));
      }
    A.opsR_P.push_back({r.lm_off, (int)col, r.nlm, 0, -1});
    A.opsS.push_back({s.lm_off, (int)col, r.nlm, s.nlm, off});
    col += r.nlm;
  }
  // O_R  (:915-924): W_S (S bands on R's N_R sites) against P_R
  if (!recip && !L.N_R.empty()) {
    if (S->wp_num != (int)L.N_R.size()) throw std::runtime_error("overlap_setup_real was not run for these site lists");
    int woff = 0;
    for (size_t q = 0; q < L.N_R.size(); q++) {
      const SiteDev& r = siteR(L.N_R[q]);
      if (S->wp_nlm[q] != r.nlm) throw std::runtime_error("wave-projection channel mismatch (N_R)");
      A.opsR_P.push_back({r.lm_off, (int)col, r.nlm, 0, -1});
      A.opsS_W.push_back({woff, (int)col, r.nlm, 0, -1});
      woff += r.nlm;
      col += r.nlm;
    }
  }
  // O_S  (:929-938): W_R (R bands on S's N_S sites) against P_S
  if (!recip && !L.N_S.empty()) {
    if (R->wp_num != (int)L.N_S.size()) throw std::runtime_error("overlap_setup_real was not run for these site lists");
    int woff = 0;
    for (size_t q = 0; q < L.N_S.size(); q++) {
      const SiteDev& s = siteS(L.N_S[q]);
      if (R->wp_nlm[q] != s.nlm) throw std::runtime_error("wave-projection channel mismatch (N_S)");
      A.opsR_W.push_back({woff, (int)col, s.nlm, 0, -1});
      A.opsS.push_back({s.lm_off, (int)col, s.nlm, 0, -1});
      woff += s.nlm;
      col += s.nlm;
    }
  }
  // O_N  (:944-959)
  if (!L.N_RS_R.empty()) {
    if (S->omega.size() != L.N_RS_R.size()) throw std::runtime_error("off-site overlaps missing: run overlap_setup_real");
    for (size_t q = 0; q < L.N_RS_R.size(); q++) {
      const SiteDev &r = siteR(L.N_RS_R[q]), &s = siteS(L.N_RS_S[q]);
      if (S->omega_n1[q] != r.nlm || S->omega_n2[q] != s.nlm) throw std::runtime_error("off-site block shape mismatch");
      const long off = (long)A.mats.size();
      A.mats.resize(A.mats.size() + (size_t)r.nlm * s.nlm);
      A.rs_mats.push_back({(int)q, off});
      A.opsR_P.push_back({r.lm_off, (int)col, r.nlm, 0, -1});
      A.opsS.push_back({s.lm_off, (int)col, r.nlm, s.nlm, off});
      col += r.nlm;
    }
  }
  A.K = col;
  A.Kpad = (col + 7) / 8 * 8;
  return A;
}

void apply_ops(const std::vector<BlockOp>& ops, const DevBuf& mats, const double2* src, long lds, double2* dst,
               long ldd, int nrows) {
  if (ops.empty()) return;
  DevBuf d = upload(ops);
  dim3 grid(std::min(nrows, 4096), (unsigned)ops.size());
  block_apply_kernel<<<grid, 128, 0, g_stream>>>(d.as<BlockOp>(), mats.as<double2>(), src, lds, dst, ldd, nrows);
  count_launch();
  check_launch();
}

void aug_block(pawb200_pswf* S, pawb200_pswf* R, AugPlan& A, int kap, int flip, double2* out, long ldo,
               bool accumulate) {
  HostSection hs_("aug_block");
  if (A.K == 0) {
    if (!accumulate)
      CUDA_OK(cudaMemset2DAsync(out, ldo * sizeof(double2), 0, R->nband * sizeof(double2), S->nband, g_stream));
    return;
  }
  const int kr = flipped(R, kap, flip);
  if (!S->resident[kap] || !R->resident[kr]) throw std::runtime_error("(k,spin) block not resident on this rank");
  // k-dependent off-site blocks: Omega * exp(2 pi i k_R . d)   (projector.c:953-955)
  for (auto& pr : A.rs_mats) {
    const int q = pr.first;
    const double* d = S->dcoords.data() + 3 * q;
    const double* k = R->kp[kr].k;
    const double ang = 2 * kPi * (k[0] * d[0] + k[1] * d[1] + k[2] * d[2]);
    const cdouble ph = std::exp(cdouble(0, 1) * ang);
    for (size_t e = 0; e < S->omega[q].size(); e++) A.mats[pr.second + e] = S->omega[q][e] * ph;
  }
  DevBuf mats = upload(A.mats.empty() ? std::vector<cdouble>(1) : A.mats);
  DevBuf opS((size_t)S->nband * A.Kpad * sizeof(double2)), opR((size_t)R->nband * A.Kpad * sizeof(double2));
  {
    ScopedStage tm(ST_GEMM_AUG);
    opS.zero();
    opR.zero();
    apply_ops(A.opsS, mats, S->P[kap].as<double2>(), S->ldp, opS.as<double2>(), A.Kpad, S->nband);
    apply_ops(A.opsS_W, mats, S->W.empty() ? nullptr : S->W[kap].as<double2>(), S->ldw, opS.as<double2>(), A.Kpad, S->nband);
    apply_ops(A.opsR_P, mats, R->P[kr].as<double2>(), R->ldp, opR.as<double2>(), A.Kpad, R->nband);
    apply_ops(A.opsR_W, mats, R->W.empty() ? nullptr : R->W[kr].as<double2>(), R->ldw, opR.as<double2>(), A.Kpad, R->nband);
  }
  run_zgemm<double2>(opS.as<double2>() + (long)S->band_lo * A.Kpad, A.Kpad, opR.as<double2>(), A.Kpad,
                     S->band_hi - S->band_lo, R->nband, A.Kpad, out + (long)S->band_lo * ldo, ldo, accumulate, ST_GEMM_AUG);
}

// Full block(s) to host: out[kap - lo][bS][bR]
// (f3) plane-wave part of the aug_recip correction  [projector.c:1016-1027]: out += <CA_R|C_S> + <C_R|CA_S>
void recip_block(pawb200_pswf* S, pawb200_pswf* R, const SiteLists& L, int kap, int flip, double2* out, long ldo) {
  const int kr = flipped(R, kap, flip);
  if (S->kp[kap].nplane != R->kp[kr].nplane)
    throw std::runtime_error("plane-wave bases differ between the two wavefunctions at kappa " + std::to_string(kap));
  if (!L.N_R.empty()) {
    if (R->CA.empty() || !R->CA[kr].p) throw std::runtime_error("overlap_setup_recip was not run for these site lists (N_R)");
    wait_coeffs(S, kap, 0, S->nband);
    run_zgemm<float2>(S->C[kap].as<float2>() + (long)S->band_lo * S->ldc[kap], S->ldc[kap], R->CA[kr].as<float2>(),
                      R->ldc[kr], S->band_hi - S->band_lo, R->nband, S->ldc[kap], out + (long)S->band_lo * ldo, ldo, true,
                      ST_GEMM_AUG);
  }
  if (!L.N_S.empty()) {
    if (S->CA.empty() || !S->CA[kap].p) throw std::runtime_error("overlap_setup_recip was not run for these site lists (N_S)");
    wait_coeffs(R, kr, 0, R->nband);
    run_zgemm<float2>(S->CA[kap].as<float2>() + (long)S->band_lo * S->ldc[kap], S->ldc[kap], R->C[kr].as<float2>(),
                      R->ldc[kr], S->band_hi - S->band_lo, R->nband, S->ldc[kap], out + (long)S->band_lo * ldo, ldo, true,
                      ST_GEMM_AUG);
  }
}

// ---- pseudo GEMMs launched ahead of the projection work (asynchronous ingest) ---------------------------------
// While coefficients are still crossing the host link the main stream sits behind transforms that wait for data,
// and anything queued after them - the pseudo GEMMs of overlap_matrix - cannot start before the last block has been
// projected (r02 trace, 4 blocks per rank: first GEMM at 455 ms of a 993 ms step).  The pseudo overlap needs nothing
// but coefficients, so overlap_setup_real - the first call that sees both wavefunctions - queues the GEMM of every
// resident block on the GEMM stream right away, per ingest chunk, into library-owned blocks; overlap_matrix later
// joins, copies the block (64 MB device to device) and adds the augmentation onto it.
struct PrelaunchedPseudo {
  uint64_t s_id = 0, r_id = 0;
  int nS = 0, nR = 0;
  std::vector<int> slot_of_kap;               // -1: not prelaunched
};
PrelaunchedPseudo g_pre;
std::vector<RawBuf> g_pre_blk;                // block buffers, kept across wavefunction pairs
std::vector<cudaEvent_t> g_pre_done, g_pre_used;

bool any_coeffs_in_flight(const pawb200_pswf* wf) {
  static const bool force = getenv("PAWB200_GEMM_CHUNKED") != nullptr;   // tests: take the prelaunch path always
  if (force) return true;
  for (auto& j : g_pending_chunks)
    if (j.wf == wf) return true;
  for (auto& c : wf->chunks)
    if (cudaEventQuery(c.ready) == cudaErrorNotReady) return true;
  return false;
}

void prelaunch_pseudo(pawb200_pswf* S, pawb200_pswf* R) {
  g_pre = PrelaunchedPseudo();
  static const bool off = getenv("PAWB200_PRELAUNCH") && atoi(getenv("PAWB200_PRELAUNCH")) == 0;
  if (off || S->ncl || R->ncl || S->band_sharded() || R->band_sharded()) return;
  if (!any_coeffs_in_flight(S) && !any_coeffs_in_flight(R)) return;      // resident inputs: nothing to gain
  const int NK = S->nkappa();
  const size_t blk_bytes = (size_t)S->nband * R->nband * sizeof(double2);
  g_pre.s_id = S->id; g_pre.r_id = R->id; g_pre.nS = S->nband; g_pre.nR = R->nband;
  g_pre.slot_of_kap.assign(NK, -1);
  cudaStream_t st2 = gemm_stream();
  int slot = 0;
  for (int kap = 0; kap < NK; kap++) {
    if (!S->resident[kap] || !R->resident[kap] || S->kp[kap].nplane != R->kp[kap].nplane) continue;
    if ((int)g_pre_blk.size() <= slot) {
      g_pre_blk.emplace_back();
      cudaEvent_t a, b;
      CUDA_OK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
      g_pre_done.push_back(a);
      g_pre_used.push_back(b);
    }
    g_pre_blk[slot].ensure(blk_bytes);
    CUDA_OK(cudaStreamWaitEvent(st2, g_pre_used[slot], 0));       // the previous pair's consumer of this buffer
    pseudo_block(S, R, kap, 0, (double2*)g_pre_blk[slot].p, R->nband, st2, true);
    CUDA_OK(cudaEventRecord(g_pre_done[slot], st2));
    g_pre.slot_of_kap[kap] = slot++;
  }
}

// the prelaunched pseudo block of (S, R, kap) copied into b on the main stream; false if there is none
bool take_prelaunched(const pawb200_pswf* S, const pawb200_pswf* R, int kap, int flip, double2* b) {
  if (flip || g_pre.s_id != S->id || g_pre.r_id != R->id || g_pre.nS != S->nband || g_pre.nR != R->nband) return false;
  if (kap >= (int)g_pre.slot_of_kap.size() || g_pre.slot_of_kap[kap] < 0) return false;
  const int slot = g_pre.slot_of_kap[kap];
  const size_t blk_bytes = (size_t)S->nband * R->nband * sizeof(double2);
  CUDA_OK(cudaStreamWaitEvent(g_stream, g_pre_done[slot], 0));
  CUDA_OK(cudaMemcpyAsync(b, g_pre_blk[slot].p, blk_bytes, cudaMemcpyDeviceToDevice, g_stream));
  CUDA_OK(cudaEventRecord(g_pre_used[slot], g_stream));
  g_pre.slot_of_kap[kap] = -1;                                   // one use: the GEMM stream may refill the buffer
  return true;
}

// dev_out: `out` is DEVICE memory [hi - lo][nbS][nbR]; the blocks are computed in place on the main stream and
// nothing is copied or synchronised (the caller orders its own work after the main stream).
void overlap_matrix(pawb200_pswf* S, pawb200_pswf* R, const SiteLists* L, int flip, int lo, int hi,
                    bool pseudo, bool aug, cdouble* out, bool recip = false, bool dev_out = false) {
  HostSection hs_("overlap_matrix");
  check_pair(S, R);
  const int nS = S->nband, nR = R->nband;
  const size_t blk_bytes = (size_t)nS * nR * sizeof(double2);
  // the pseudo GEMM goes to its own stream when asked to (PAWB200_GEMM_OVERLAP) or while wf coefficients are still
  // arriving (async ingest): it then starts per ingest chunk, ahead of the transforms queued on the main stream
  bool side = pseudo && gemm_overlap_enabled();
  for (int kap = lo; kap < hi && pseudo && !side; kap++)
    if (S->resident[kap] && coeffs_in_flight(S, kap)) side = true;
  DevBuf blk;
  if (!side && !dev_out) blk.alloc(blk_bytes);
  AugPlan A;
  if (aug) A = plan_aug(S, R, *L, recip);
  int use = 0;
  for (int kap = lo; kap < hi; kap++) {
    cdouble* dst = out + (size_t)(kap - lo) * nS * nR;
    const int kr = flipped(R, kap, flip);
    if (!S->resident[kap] || !R->resident[kr]) {              // another rank's block
      if (dev_out) CUDA_OK(cudaMemsetAsync(dst, 0, blk_bytes, g_stream));
      else std::fill(dst, dst + (size_t)nS * nR, cdouble(0, 0));
      continue;
    }
    double2* b;
    if (pseudo && g_pre.s_id == S->id && kap < (int)g_pre.slot_of_kap.size() && g_pre.slot_of_kap[kap] >= 0) {
      // the pseudo GEMM of this block was queued by overlap_setup_real (prelaunch_pseudo): join, copy, augment
      DevBuf own;
      if (dev_out) b = (double2*)dst;
      else { own.alloc(blk_bytes); b = own.as<double2>(); }
      if (take_prelaunched(S, R, kap, flip, b)) {
        if (aug) aug_block(S, R, A, kap, flip, b, nR, true);
        if (aug && recip) recip_block(S, R, *L, kap, flip, b, nR);
        if (dev_out) continue;
        ScopedStage tm(ST_D2H);
        CUDA_OK(cudaMemcpyAsync(dst, b, blk_bytes, cudaMemcpyDeviceToHost, g_stream));
        trace_mark("d2h block done", g_stream);
        continue;
      }
    }
    if (side) {
      // pseudo block on the GEMM stream into a dedicated buffer (or the caller's device block); the main stream
      // joins before it adds the augmentation GEMM and copies the block out
      cudaStream_t st2 = gemm_stream();
      const int slot = use++ & 1;
      if (dev_out) {
        b = (double2*)dst;
        CUDA_OK(cudaEventRecord(g_pblk_free[slot], g_stream));     // earlier main-stream users of the block are done
      } else {
        g_pblk[slot].ensure(blk_bytes);
        b = (double2*)g_pblk[slot].p;
      }
      CUDA_OK(cudaStreamWaitEvent(st2, g_pblk_free[slot], 0));
      if (S->band_sharded()) CUDA_OK(cudaMemsetAsync(b, 0, blk_bytes, st2));        // rows of other ranks' bands
      pseudo_block(S, R, kap, flip, b, nR, st2, true);
      CUDA_OK(cudaEventRecord(g_pblk_done[slot], st2));
      CUDA_OK(cudaStreamWaitEvent(g_stream, g_pblk_done[slot], 0));
      if (aug) aug_block(S, R, A, kap, flip, b, nR, true);
      if (aug && recip) recip_block(S, R, *L, kap, flip, b, nR);
      if (dev_out) continue;
      ScopedStage tm(ST_D2H);
      CUDA_OK(cudaMemcpyAsync(dst, b, blk_bytes, cudaMemcpyDeviceToHost, g_stream));
      trace_mark("d2h block done", g_stream);
      CUDA_OK(cudaEventRecord(g_pblk_free[slot], g_stream));
      continue;
    }
    b = dev_out ? (double2*)dst : blk.as<double2>();
    if (S->band_sharded()) CUDA_OK(cudaMemsetAsync(b, 0, blk_bytes, g_stream));    // rows of other ranks' bands
    if (pseudo) pseudo_block(S, R, kap, flip, b, nR);
    if (aug) aug_block(S, R, A, kap, flip, b, nR, pseudo);
    if (aug && recip) recip_block(S, R, *L, kap, flip, b, nR);
    if (dev_out) continue;
    ScopedStage tm(ST_D2H);
    CUDA_OK(cudaMemcpyAsync(dst, b, blk_bytes, cudaMemcpyDeviceToHost, g_stream));
  }
  if (!dev_out) stream_sync();
}

// ---- real-space states -----------------------------------------------------------------------
SiteTables& ae_tables(pawb200_pswf* wf, const int* fftg, const int* labels, const double* coords) {
  require_projections(wf);
  const size_t h = hash_bytes(coords, sizeof(double) * 3 * wf->num_sites, hash_bytes(labels, sizeof(int) * wf->num_sites));
  if (!wf->ae_sites || wf->ae_sites->fftg[0] != fftg[0] || wf->ae_sites->fftg[1] != fftg[1] ||
      wf->ae_sites->fftg[2] != fftg[2] || wf->ae_sites->coord_hash != h) {
    std::vector<int> all(wf->num_sites);
    for (int i = 0; i < wf->num_sites; i++) all[i] = i;
    wf->ae_sites = build_site_tables(wf->pps->list.el, all.data(), wf->num_sites, labels, coords, wf->lattice,
                                     fftg, 2, false);
    wf->ae_sites->coord_hash = h;
  }
  return *wf->ae_sites;
}

// boxes for slots [slot0, slot0+nslot) of kappa, AE-augmented, left in g_grid
void realspace_boxes(pawb200_pswf* wf, int kap, int slot0, int nslot, const int* fftg, SiteTables& T,
                     const DevBuf& inv) {
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  g_grid.ensure((size_t)nslot * ngrid * sizeof(double2));
  double2* x = g_grid.as<double2>();
  launch_scatter(wf, kap, slot0, nslot, inv, x, fftg);
  launch_fft(x, fftg, nslot, CUFFT_INVERSE);
  const double* k = wf->kp[kap].k;
  ScopedStage tm(ST_AUGMENT);
  const long blocks = std::min<long>((ngrid + 255) / 256, (long)g_num_sms * 16);
  if (k[0] != 0 || k[1] != 0 || k[2] != 0) {   // exp(0) = 1 exactly: nothing to do at Gamma
    bloch_phase_kernel<<<(unsigned)blocks, 256, 0, g_stream>>>(x, fftg[0], fftg[1], fftg[2], k[0], k[1], k[2], 1.0, nslot);
    count_launch();
    check_launch();
  }
  if (T.nsites && T.total_pts) {
    int maxpts = 0, maxlm = 0;
    for (auto& sd : T.host) { maxpts = std::max(maxpts, sd.npts); maxlm = std::max(maxlm, sd.nlm); }
    // P rows of these slots: sites are in structure order so lm offsets coincide with proj_sites
    constexpr int NBMAX = 32;
    for (int b0 = 0; b0 < nslot; b0 += NBMAX) {
      const int nb = std::min(NBMAX, nslot - b0);
      dim3 grid(std::max(1, (maxpts + 255) / 256), T.nsites);
      augment_add_kernel<<<grid, 256, sizeof(double2) * nb * maxlm, g_stream>>>(
          T.sites.as<SiteDev>(), T.idx.as<int>(), T.wrap.as<int>(), T.total_pts, T.table.as<double2>(),
          wf->P[kap].as<double2>() + (long)(slot0 + b0) * wf->ldp, wf->ldp, nb, x + (long)b0 * ngrid, ngrid,
          k[0], k[1], k[2]);
      count_launch();
      check_launch();
    }
  }
}

// Same, through the pruned transform: boxes in its band-interleaved layout X[group][g][16] (left in g_grid).
// Used by the density accumulation, which never needs a box in the caller's layout.
void realspace_boxes_il(pawb200_pswf* wf, int kap, const PrunedPlan& plan, int slot0, int nslot, const int* fftg,
                        SiteTables& T) {
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  const int ng = (nslot + FFT_B - 1) / FFT_B;
  g_grid.ensure((size_t)ng * FFT_B * ngrid * sizeof(double2));
  double2* x = g_grid.as<double2>();
  pruned_fft(wf, kap, plan, slot0, nslot, x);
  const double* k = wf->kp[kap].k;
  ScopedStage tm(ST_AUGMENT);
  if (k[0] != 0 || k[1] != 0 || k[2] != 0) {
    const long blocks = std::min<long>((ngrid + 255) / 256, (long)g_num_sms * 16);
    bloch_phase_il_kernel<<<(unsigned)blocks, 256, 0, g_stream>>>(x, fftg[0], fftg[1], fftg[2], k[0], k[1], k[2], 1.0, ng);
    count_launch();
    check_launch();
  }
  if (T.nsites && T.total_pts) {
    int maxpts = 0, maxlm = 0;
    for (auto& sd : T.host) { maxpts = std::max(maxpts, sd.npts); maxlm = std::max(maxlm, sd.nlm); }
    constexpr int NBMAX = 32;
    for (int b0 = 0; b0 < nslot; b0 += NBMAX) {
      const int nb = std::min(NBMAX, nslot - b0);
      dim3 grid(std::max(1, (maxpts + 255) / 256), T.nsites);
      augment_add_il_kernel<<<grid, 256, sizeof(double2) * nb * maxlm, g_stream>>>(
          T.sites.as<SiteDev>(), T.idx.as<int>(), T.wrap.as<int>(), T.total_pts, T.table.as<double2>(),
          wf->P[kap].as<double2>() + (long)slot0 * wf->ldp, wf->ldp, b0, nb, x, ngrid, k[0], k[1], k[2]);
      count_launch();
      check_launch();
    }
  }
}

void check_kpoint(const pawb200_pswf* wf, int band, int kap) {
  if (!wf) throw std::runtime_error("NULL wavefunction pointer");
  if (band < 0 || band >= wf->nband) throw std::runtime_error("band index out of range");
  if (kap < 0 || kap >= wf->nkappa()) throw std::runtime_error("k-point/spin index out of range");
  if (!wf->resident[kap]) throw std::runtime_error("(k,spin) block not resident on this rank");
  if (band < wf->band_lo || band >= wf->band_hi) throw std::runtime_error("band belongs to another rank's band block");
}

// Device -> arbitrary (pageable or pinned) host memory through two page-locked staging buffers: the copy of
// chunk i+1 overlaps the host-side store (or += for accumulate) of chunk i, so large grids leave at PCIe rate instead
// of the pageable-memcpy rate.  Synchronises the main stream.
void d2h_pipelined(void* dst, const void* src_dev, size_t bytes, bool accumulate_f64) {
  static unsigned char* stage[2] = {nullptr, nullptr};
  static cudaEvent_t done[2];
  const size_t chunk = (size_t)32 << 20;
  if (!stage[0]) {
    for (int i = 0; i < 2; i++) {
      CUDA_OK(cudaMallocHost((void**)&stage[i], chunk));
      CUDA_OK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
  }
  const size_t nchunk = (bytes + chunk - 1) / chunk;
  auto issue = [&](size_t c) {
    const size_t off = c * chunk, n = std::min(chunk, bytes - off);
    CUDA_OK(cudaMemcpyAsync(stage[c & 1], (const unsigned char*)src_dev + off, n, cudaMemcpyDeviceToHost, g_stream));
    CUDA_OK(cudaEventRecord(done[c & 1], g_stream));
  };
  if (nchunk) issue(0);
  for (size_t c = 0; c < nchunk; c++) {
    CUDA_OK(cudaEventSynchronize(done[c & 1]));
    const size_t off = c * chunk, n = std::min(chunk, bytes - off);
    if (accumulate_f64) {
      double* d = (double*)((unsigned char*)dst + off);
      const double* s8 = (const double*)stage[c & 1];
      const long cnt = (long)(n / sizeof(double));
      // chunk c+1 may not be issued into the other buffer before chunk c-1's host pass is over: it is (sequential)
      if (c + 1 < nchunk) issue(c + 1);
#pragma omp parallel for schedule(static)
      for (long i = 0; i < cnt; i++) d[i] += s8[i];
    } else {
      if (c + 1 < nchunk) issue(c + 1);
      unsigned char* d = (unsigned char*)dst + off;
      const unsigned char* s1 = stage[c & 1];
      const long parts = 16, per = (long)((n + parts - 1) / parts);
#pragma omp parallel for schedule(static)
      for (long q = 0; q < parts; q++) {
        const long o = q * per;
        if (o < (long)n) memcpy(d + o, s1 + o, (size_t)std::min<long>(per, (long)n - o));
      }
    }
  }
  CUDA_OK(cudaStreamSynchronize(g_stream));
}

void state_to_host(pawb200_c128* out, int band, int kap, pawb200_pswf* wf, const int* fftg, const int* labels,
                   const double* coords) {
  require_device();
  check_kpoint(wf, band, kap);
  SiteTables& T = ae_tables(wf, fftg, labels, coords);
  const int h = wf->halves();
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  std::shared_ptr<PrunedPlan> plan = get_pruned_plan(wf, kap, fftg);
  if (plan->ok && !getenv("PAWB200_DENSITY_GENERIC")) {
    // hand-written pruned transform (one interleave group; the other 15 slots are zero columns), Bloch phase and
    // augmentation on the interleaved box, then the state's slot(s) are copied out in the caller's layout
    realspace_boxes_il(wf, kap, *plan, band * h, h, fftg, T);
    DevBuf planar((size_t)h * ngrid * sizeof(double2));
    const long blocks = std::min<long>((ngrid * h + 255) / 256, (long)g_num_sms * 16);
    extract_il_kernel<<<(unsigned)blocks, 256, 0, g_stream>>>(g_grid.as<double2>(), ngrid, 0, h, planar.as<double2>());
    count_launch();
    check_launch();
    d2h_pipelined(out, planar.p, (size_t)h * ngrid * sizeof(double2), false);
    return;
  }
  DevBuf inv = build_inverse_map(wf, kap, fftg);
  realspace_boxes(wf, kap, band * h, h, fftg, T, inv);
  d2h_pipelined(out, g_grid.p, (size_t)h * ngrid * sizeof(double2), false);
}

void density_to_host(double* Pout, pawb200_pswf* wf, const int* fftg, const int* labels, const double* coords,
                     int only_band, int only_kap, int band_lo = 0, int band_hi = 1 << 30) {
  require_device();
  if (!wf) throw std::runtime_error("NULL wavefunction pointer");
  SiteTables& T = ae_tables(wf, fftg, labels, coords);
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  const int h = wf->halves();
  DevBuf rho(ngrid * sizeof(double));
  rho.zero();
  long batch = std::max<long>(1, (long)(fft_budget_bytes() / (sizeof(double2) * ngrid)) / h);
  batch = std::min<long>(batch, 64);
  const double spin_mult = wf->ncl ? 1.0 : 2 / wf->nspin;   // density.c:164, 187
  for (int kap = 0; kap < wf->nkappa(); kap++) {
    if (only_kap >= 0 && kap != only_kap) continue;
    if (!wf->resident[kap]) continue;
    std::vector<int> bands;
    std::vector<double> wts;
    if (only_band >= 0) {
      bands.push_back(only_band);
      wts.push_back(1.0);                                    // ae_state_density, density.c:35-37
    } else {
      for (int b = std::max(wf->band_lo, band_lo); b < std::min(wf->band_hi, band_hi); b++)   // own band block only
        if (wf->kp[kap].occ[b] > 0) {                        // density.c:168
          bands.push_back(b);
          wts.push_back(wf->weight[kap] * wf->kp[kap].occ[b] * spin_mult);
        }
    }
    if (bands.empty()) continue;
    // pruned transform + interleaved boxes when the grid factors and one 16-slot group fits the box budget
    std::shared_ptr<PrunedPlan> plan = get_pruned_plan(wf, kap, fftg);
    const size_t grp_bytes = (size_t)FFT_B * ngrid * sizeof(double2);
    const long il_groups = std::min<long>(4, (long)(keep_boxes_budget() / grp_bytes));
    if (plan->ok && il_groups >= 1 && !getenv("PAWB200_DENSITY_GENERIC")) {
      const long il_batch = il_groups * FFT_B / h;          // bands per batch
      size_t i = 0;
      while (i < bands.size()) {
        size_t j = i + 1;
        while (j < bands.size() && bands[j] == bands[j - 1] + 1 && (long)(j - i) < il_batch) j++;
        const int nb = (int)(j - i), nslot = nb * h, ng = (nslot + FFT_B - 1) / FFT_B;
        realspace_boxes_il(wf, kap, *plan, bands[i] * h, nslot, fftg, T);
        std::vector<double> w((size_t)ng * FFT_B, 0.0);     // pad slots of the last group weigh nothing
        for (int q = 0; q < nb; q++)
          for (int hh = 0; hh < h; hh++) w[q * h + hh] = wts[i + q];
        DevBuf dw = upload(w);
        const long blocks = std::min<long>((ngrid * 16 + 255) / 256, (long)g_num_sms * 16);
        ScopedStage tm(ST_AUGMENT);
        density_accum_il_kernel<<<(unsigned)blocks, 256, 0, g_stream>>>(g_grid.as<double2>(), ngrid, ng,
                                                                        dw.as<double>(), rho.as<double>());
        count_launch();
        check_launch();
        i = j;
      }
      continue;
    }
    DevBuf inv = build_inverse_map(wf, kap, fftg);
    // occupied bands are contiguous in practice; process maximal runs in batches
    size_t i = 0;
    while (i < bands.size()) {
      size_t j = i + 1;
      while (j < bands.size() && bands[j] == bands[j - 1] + 1 && (long)(j - i) < batch) j++;
      const int nb = (int)(j - i);
      realspace_boxes(wf, kap, bands[i] * h, nb * h, fftg, T, inv);
      std::vector<double> w(nb * h);
      for (int q = 0; q < nb; q++)
        for (int hh = 0; hh < h; hh++) w[q * h + hh] = wts[i + q];
      DevBuf dw = upload(w);
      const long blocks = std::min<long>((ngrid + 255) / 256, (long)g_num_sms * 16);
      ScopedStage tm(ST_AUGMENT);
      density_accum_kernel<<<(unsigned)blocks, 256, 0, g_stream>>>(g_grid.as<double2>(), ngrid, nb * h,
                                                                   dw.as<double>(), rho.as<double>());
      count_launch();
      check_launch();
      i = j;
    }
  }
  d2h_pipelined(Pout, rho.p, (size_t)ngrid * sizeof(double), true);     // the reference accumulates into P
}

}  // namespace

// =======================================================================================
// C ABI
// =======================================================================================
extern "C" {

const char* pawb200_last_error(void) { return g_has_error ? g_error.c_str() : nullptr; }
void pawb200_clear_error(void) { g_has_error = false; }
const char* pawb200_version(void) { return "pawpyseed_b200 0.1 (sm_100a)"; }

int pawb200_device_check(void) {
  API_BEGIN
  require_device();
  return 0;
  API_END(-1)
}

void pawb200_set_async_ingest(int on) { g_async_ingest = on != 0; }

void pawb200_set_host_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

void pawb200_set_band_shard(int rank, int world) {
  if (world < 1) world = 1;
  g_band_rank = ((rank % world) + world) % world;
  g_band_world = world;
}

void* pawb200_get_device_buffer(pawb200_pswf_t* wf, int which, int kappa, long* ld, int* rows, int* own_lo,
                                int* own_hi) {
  API_BEGIN
  if (!wf) throw std::runtime_error("NULL wavefunction pointer");
  if (kappa < 0 || kappa >= wf->nkappa() || !wf->resident[kappa]) throw std::runtime_error("(k,spin) block not resident");
  const int h = wf->halves();
  void* p = nullptr;
  long l = 0;
  int r = 0, lo = 0, hi = 0;
  if (which == 0) {                      // plane-wave coefficients: one row per band
    // everything queued after this call on the main (legacy default) stream sees the landed coefficients
    wait_coeffs(wf, kappa, 0, wf->nband);
    p = wf->C[kappa].p; l = wf->ldc[kappa]; r = wf->band_rows; lo = wf->band_lo; hi = wf->band_hi;
  } else if (which == 1 || which == 2) {   // projections / wave projections: one row per slot
    require_projections(wf);
    std::vector<DevBuf>& v = which == 1 ? wf->P : wf->W;
    if (kappa >= (int)v.size() || !v[kappa].p) throw std::runtime_error("projections have not been set up");
    p = v[kappa].p; l = which == 1 ? wf->ldp : wf->ldw; r = wf->band_rows * h; lo = wf->band_lo * h; hi = wf->band_hi * h;
  } else {
    throw std::runtime_error("unknown buffer id");
  }
  if (ld) *ld = l;
  if (rows) *rows = r;
  if (own_lo) *own_lo = lo;
  if (own_hi) *own_hi = hi;
  return p;
  API_END(nullptr)
}

void pawb200_set_read_shard(int rank, int world) {
  if (world < 1) world = 1;
  g_shard_rank = ((rank % world) + world) % world;
  g_shard_world = world;
}

pawb200_pswf_t* pawb200_read_wavefunctions(const char* filename, const double* kws) {
  API_BEGIN
  ByteSource s;
  s.fp = fopen(filename, "rb");
  if (!s.fp) throw std::runtime_error(std::string("cannot open ") + filename);
  pawb200_pswf* wf = nullptr;
  try {
    wf = ingest(s, kws);
  } catch (...) {
    fclose(s.fp);
    throw;
  }
  fclose(s.fp);
  return wf;
  API_END(nullptr)
}

pawb200_pswf_t* pawb200_read_wavefunctions_from_str(const char* start, const double* kws) {
  API_BEGIN
  if (!start) throw std::runtime_error("NULL WAVECAR buffer");
  ByteSource s;
  s.mem = (const unsigned char*)start;
  return ingest(s, kws);
  API_END(nullptr)
}

void pawb200_free_pswf(pawb200_pswf_t* wf) {
  if (!wf) return;
  drop_pending_ingest(wf);            // deferred chunks that nobody asked for
  cudaStreamSynchronize(ingest_ring().copy);
  cudaStreamSynchronize(ingest_ring().unpack);
  if (g_stream2) cudaStreamSynchronize(g_stream2);
  cudaStreamSynchronize(g_stream);
  delete wf;
}

int pawb200_get_nband(pawb200_pswf_t* wf) { return wf ? wf->nband : 0; }
int pawb200_get_nwk(pawb200_pswf_t* wf) { return wf ? wf->nwk : 0; }
int pawb200_get_nspin(pawb200_pswf_t* wf) { return wf ? wf->nspin : 0; }
int pawb200_is_ncl(pawb200_pswf_t* wf) { return wf ? wf->ncl : 0; }
double pawb200_get_encut(pawb200_pswf_t* wf) { return wf ? wf->encut : 0; }

double pawb200_get_energy(pawb200_pswf_t* wf, int band, int kpt, int spin) {
  API_BEGIN
  if (!wf) throw std::runtime_error("NULL wavefunction pointer");
  const int kap = kpt + spin * wf->nwk;
  if (band < 0 || band >= wf->nband || kpt < 0 || kpt >= wf->nwk || spin < 0 || spin >= wf->nspin)
    throw std::runtime_error("index out of range");
  return wf->kp[kap].energy[band];
  API_END(NAN)
}
double pawb200_get_occ(pawb200_pswf_t* wf, int band, int kpt, int spin) {
  API_BEGIN
  if (!wf) throw std::runtime_error("NULL wavefunction pointer");
  const int kap = kpt + spin * wf->nwk;
  if (band < 0 || band >= wf->nband || kpt < 0 || kpt >= wf->nwk || spin < 0 || spin >= wf->nspin)
    throw std::runtime_error("index out of range");
  return wf->kp[kap].occ[band];
  API_END(NAN)
}
double* pawb200_get_occs(pawb200_pswf_t* wf) {
  if (!wf) return nullptr;
  const int NK = wf->nkappa();
  double* o = (double*)malloc(sizeof(double) * NK * wf->nband);
  for (int k = 0; k < NK; k++)
    for (int b = 0; b < wf->nband; b++) o[b * NK + k] = wf->kp[k].occ[b];
  return o;
}
void pawb200_set_num_sites(pawb200_pswf_t* wf, int nsites) { if (wf) wf->num_sites = nsites; }
void pawb200_free_ptr(void* p) { free(p); }

pawb200_ppot_t* pawb200_get_projector_list(int num_els, const int* labels, const int* ls, const double* wave_grids,
                                           const double* projectors, const double* aewaves,
                                           const double* pswaves, const double* rmaxs, double grid_encut) {
  API_BEGIN
  HostSection hs_("get_projector_list");
  auto p = std::make_unique<pawb200_ppot>();
  p->list.el = build_elements(num_els, labels, ls, wave_grids, projectors, aewaves, pswaves, rmaxs, grid_encut);
  return p.release();
  API_END(nullptr)
}
void pawb200_free_ppot_list(pawb200_ppot_t* pps, int) { delete pps; }

void pawb200_setup_projections(pawb200_pswf_t* wf, pawb200_ppot_t* pps, int num_elems, int num_sites,
                               const int* fftg, const int* labels, const double* coords) {
  API_BEGIN
  HostSection hs_("setup_projections");
  require_device();
  if (!wf || !pps) throw std::runtime_error("NULL argument");
  if ((int)pps->list.el.size() != num_elems) throw std::runtime_error("num_elems does not match the projector list");
  for (int s = 0; s < num_sites; s++)
    if (labels[s] < 0 || labels[s] >= num_elems) throw std::runtime_error("site label out of range");
  wf->pps.reset(pps);
  wf->num_sites = num_sites;
  for (int d = 0; d < 3; d++) wf->fftg[d] = fftg[d];
  wf->labels.assign(labels, labels + num_sites);
  wf->coords.assign(coords, coords + 3 * num_sites);
  wf->gen++;
  wf->ae_sites.reset();
  wf->W.clear();
  wf->wp_num = 0;
  std::vector<int> all(num_sites);
  for (int i = 0; i < num_sites; i++) all[i] = i;
  const bool pre = prefft_all_bands(wf, fftg);     // queue the transforms first; the host work below overlaps them
  wf->proj_sites = build_site_tables(pps->list.el, all.data(), num_sites, labels, coords, wf->lattice, fftg, 0, true);
  wf->has_projections = true;
  wf->lazy_proj = false;
  int nres = 0;
  for (int kap = 0; kap < wf->nkappa(); kap++) nres += wf->resident[kap] ? 1 : 0;
  const size_t box_bytes = (size_t)wf->slot_own() * fftg[0] * fftg[1] * fftg[2] * sizeof(double2);
  static const bool lazy_ok = !(getenv("PAWB200_LAZY") && atoi(getenv("PAWB200_LAZY")) == 0);
  if (!pre && lazy_ok && box_bytes * (size_t)std::max(nres, 1) > keep_boxes_budget()) {
    // The boxes cannot stay resident, so a later overlap_setup_real would have to transform every band again.
    // Defer: the first consumer of the projections runs the pass (require_projections), and overlap_setup_real
    // adds its own table set to that same pass.
    wf->P.clear();
    wf->boxes.clear();
    wf->lazy_proj = true;
  } else {
    project_all_bands(wf, *wf->proj_sites, fftg, wf->P, wf->ldp, !pre);
  }
  API_END_VOID
}

void pawb200_projection_matrix(pawb200_c128* out, pawb200_pswf_t* wf_S, pawb200_pswf_t* wf_R, int num_M,
                               int num_N_R, int num_N_S, int num_N_RS, const int* M_R, const int* M_S,
                               const int* N_R, const int* N_S, const int* N_RS_R, const int* N_RS_S,
                               int flip_spin, int kappa_lo, int kappa_hi, int pseudo_only) {
  API_BEGIN
  require_device();
  check_pair(wf_S, wf_R);
  if (kappa_lo < 0 || kappa_hi > wf_S->nkappa() || kappa_lo > kappa_hi) throw std::runtime_error("bad kappa range");
  SiteLists L = make_lists(num_M, num_N_R, num_N_S, num_N_RS, M_R, M_S, N_R, N_S, N_RS_R, N_RS_S);
  // pseudo_only: 0 = pseudo + augmentation (aug_real), 1 = pseudo only, 2 = pseudo + aug_recip augmentation
  overlap_matrix(wf_S, wf_R, &L, flip_spin, kappa_lo, kappa_hi, true, pseudo_only != 1, (cdouble*)out,
                 pseudo_only == 2);
  API_END_VOID
}

// Same blocks, left in DEVICE memory the caller owns (e.g. a torch tensor): out_dev[kappa - lo][b_S][b_R].  The
// kernels are queued on the library's main stream (the legacy default stream, i.e. torch's default stream) and
// the call returns without synchronising, so a collective launched by the caller on that stream - the NCCL
// all-gather of the per-k matrices - follows without a host round trip.
void pawb200_projection_matrix_dev(void* out_dev, pawb200_pswf_t* wf_S, pawb200_pswf_t* wf_R, int num_M,
                                   int num_N_R, int num_N_S, int num_N_RS, const int* M_R, const int* M_S,
                                   const int* N_R, const int* N_S, const int* N_RS_R, const int* N_RS_S,
                                   int flip_spin, int kappa_lo, int kappa_hi, int pseudo_only) {
  API_BEGIN
  require_device();
  check_pair(wf_S, wf_R);
  if (!out_dev) throw std::runtime_error("NULL device output pointer");
  if (kappa_lo < 0 || kappa_hi > wf_S->nkappa() || kappa_lo > kappa_hi) throw std::runtime_error("bad kappa range");
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, out_dev) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
    cudaGetLastError();
    throw std::runtime_error("pawb200_projection_matrix_dev needs a device pointer");
  }
  SiteLists L = make_lists(num_M, num_N_R, num_N_S, num_N_RS, M_R, M_S, N_R, N_S, N_RS_R, N_RS_S);
  overlap_matrix(wf_S, wf_R, &L, flip_spin, kappa_lo, kappa_hi, true, pseudo_only != 1, (cdouble*)out_dev,
                 pseudo_only == 2, true);
  API_END_VOID
}

void pawb200_pseudoprojection(pawb200_c128* projections, pawb200_pswf_t* wf_ref, pawb200_pswf_t* wf_proj,
                              int BAND_NUM, int flip_spin) {
  API_BEGIN
  require_device();
  check_pair(wf_proj, wf_ref);
  if (BAND_NUM < 0 || BAND_NUM >= wf_proj->nband) throw std::runtime_error("band index out of range");
  HostMatrixCache& c = wf_proj->pseudo_cache;
  const int NK = wf_ref->nkappa(), nS = wf_proj->nband, nR = wf_ref->nband;
  if (!c.valid || c.other_id != wf_ref->id || c.other_gen != wf_ref->gen || c.self_gen != wf_proj->gen ||
      c.flip != (flip_spin ? 1 : 0)) {
    c.data.assign((size_t)NK * nS * nR, cdouble(0, 0));
    overlap_matrix(wf_proj, wf_ref, nullptr, flip_spin, 0, NK, true, false, c.data.data());
    c.other_id = wf_ref->id; c.other_gen = wf_ref->gen; c.self_gen = wf_proj->gen;
    c.flip = flip_spin ? 1 : 0;
    c.valid = true;
  }
  cdouble* out = (cdouble*)projections;
  for (int k = 0; k < NK; k++)
    for (int b = 0; b < nR; b++) out[(size_t)b * NK + k] = c.data[((size_t)k * nS + BAND_NUM) * nR + b];
  API_END_VOID
}

// Argument checks shared by overlap_setup_real / _recip: every site index against the structure it indexes, every
// label against the element list (the reference indexes pps[labels[s]] and coords + 3*s unchecked,
// projector.c:625-719).  The arrays hold num_sites entries of the structure they describe, as in the reference.
void check_overlap_setup_args(const pawb200_pswf* wf_R, const pawb200_pswf* wf_S, const int* labels_R,
                              const int* labels_S, const double* coords_R, const double* coords_S, const int* N_R,
                              const int* N_S, const int* N_RS_R, const int* N_RS_S, int num_N_R, int num_N_S,
                              int num_N_RS) {
  if (num_N_R < 0 || num_N_S < 0 || num_N_RS < 0) throw std::runtime_error("negative site-list length");
  if (!labels_R || !labels_S || !coords_R || !coords_S) throw std::runtime_error("NULL label / coordinate array");
  if ((num_N_R && !N_R) || (num_N_S && !N_S) || (num_N_RS && (!N_RS_R || !N_RS_S)))
    throw std::runtime_error("NULL site list with a non-zero length");
  auto check_struct = [](const pawb200_pswf* wf, const int* labels, const double* coords, const char* who) {
    const int nel = (int)wf->pps->list.el.size();
    for (int s = 0; s < wf->num_sites; s++) {
      if (labels[s] < 0 || labels[s] >= nel) throw std::runtime_error(std::string("site label out of range (") + who + ")");
      for (int d = 0; d < 3; d++)
        if (!std::isfinite(coords[3 * s + d])) throw std::runtime_error(std::string("non-finite site coordinate (") + who + ")");
    }
  };
  check_struct(wf_R, labels_R, coords_R, "basis");
  check_struct(wf_S, labels_S, coords_S, "wf");
  auto check_list = [](const int* v, int n, int nsites, const char* who) {
    for (int i = 0; i < n; i++)
      if (v[i] < 0 || v[i] >= nsites) throw std::runtime_error(std::string("site index out of range in ") + who);
  };
  check_list(N_R, num_N_R, wf_R->num_sites, "N_R");
  check_list(N_S, num_N_S, wf_S->num_sites, "N_S");
  check_list(N_RS_R, num_N_RS, wf_R->num_sites, "N_RS_R");
  check_list(N_RS_S, num_N_RS, wf_S->num_sites, "N_RS_S");
}

// part 3 of overlap_setup_real / _recip (projector.c:682-719, 799-841): off-site partial-wave overlaps, host
// side (O(pairs * channels^2) radial integrals)
void setup_offsite(pawb200_pswf* wf_R, pawb200_pswf* wf_S, const int* labels_R, const int* labels_S,
                   const double* coords_R, const double* coords_S, const int* N_RS_R, const int* N_RS_S,
                   int num_N_RS) {
  const auto& elsR = wf_R->pps->list.el;
  const auto& elsS = wf_S->pps->list.el;
  wf_S->omega.assign(num_N_RS, {});
  wf_S->dcoords.assign(3 * (size_t)num_N_RS, 0.0);
  wf_S->omega_n1.assign(num_N_RS, 0);
  wf_S->omega_n2.assign(num_N_RS, 0);
  OmpErr oe;
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < num_N_RS; i++) oe.run([&] {
    const int s1 = N_RS_R[i], s2 = N_RS_S[i];
    const Element& p1 = elsR[labels_R[s1]];
    const Element& p2 = elsS[labels_S[s2]];
    double R = 0;
    double* d = wf_S->dcoords.data() + 3 * i;
    min_image_path(coords_S + 3 * s2, coords_R + 3 * s1, wf_R->lattice, d, &R);
    auto& om = wf_S->omega[i];
    om.resize((size_t)p1.total_projs * p2.total_projs);
    for (int a = 0; a < p1.total_projs; a++)
      for (int b = 0; b < p2.total_projs; b++) {
        const Channel &ca = p1.chan[a], &cb = p2.chan[b];
        const RadialFunc &fa = p1.funcs[ca.n], &fb = p2.funcs[cb.n];
        om[(size_t)a * p2.total_projs + b] = std::conj(offsite_overlap_recip(
            d, p1.kwave_grid.data(), fa.kwave.data(), fa.kwave_s, p1.wave_gridsize, p2.kwave_grid.data(),
            fb.kwave.data(), fb.kwave_s, p2.wave_gridsize, ca.l, ca.m, cb.l, cb.m));
      }
    wf_S->omega_n1[i] = p1.total_projs;
    wf_S->omega_n2[i] = p2.total_projs;
  });
  oe.rethrow();
}

// (f3) get_aug_freqs for every band of `wf` (projector.c:420-453): the (phi - phit) augmentation of the listed
// sites is accumulated on the FFT grid, transformed forward and gathered to complex64 plane-wave coefficients
// CA[kappa][band][npw] stored like C (box order).  Unlike the reference, which keeps a band's CAs from an
// earlier site list (the `CAs != NULL` early return at :424), the result is always recomputed.
void compute_aug_freqs(pawb200_pswf* wf, const int* site_list, int nlist, const int* labels, const double* coords) {
  HostSection hs_("compute_aug_freqs");
  const int* fftg = wf->fftg;
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  auto T = build_site_tables(wf->pps->list.el, site_list, nlist, labels, coords, wf->lattice, fftg, 1, false);
  std::vector<int> full_off(nlist);
  int maxpts = 0, maxlm = 0;
  for (int s = 0; s < nlist; s++) {
    if (site_list[s] < 0 || site_list[s] >= wf->num_sites) throw std::runtime_error("site index out of range");
    const SiteDev& full = wf->proj_sites->host[site_list[s]];
    if (full.nlm != T->host[s].nlm) throw std::runtime_error("labels differ from those given to setup_projections");
    full_off[s] = full.lm_off;
    maxpts = std::max(maxpts, T->host[s].npts);
    maxlm = std::max(maxlm, T->host[s].nlm);
  }
  DevBuf doff = upload(full_off);
  wf->CA.clear();
  wf->CA.resize(wf->nkappa());
  const double scale = std::pow(determinant3(wf->lattice), 0.5) / fftg[0] / fftg[1] / fftg[2];   // linalg.c:64-65
  constexpr int NBMAX = 32;   // bands per augment launch (shared-memory P tile)
  const long budget = (long)(fft_budget_bytes() / (ngrid * sizeof(double2)));
  const int batch = (int)std::max<long>(1, std::min<long>(wf->nband, budget));
  for (int kap = 0; kap < wf->nkappa(); kap++) {
    if (!wf->resident[kap]) continue;
    const int npw = wf->kp[kap].nplane;
    std::vector<int> fwd;
    build_inverse_map(wf, kap, fftg, &fwd);
    std::vector<int> gsorted(npw);
    for (int j = 0; j < npw; j++) gsorted[j] = fwd[wf->kp[kap].perm[j]];
    DevBuf dg = upload(gsorted);
    wf->CA[kap].alloc((size_t)wf->nband * wf->ldc[kap] * sizeof(float2));
    wf->CA[kap].zero((size_t)wf->nband * wf->ldc[kap] * sizeof(float2));
    double kc[3] = {wf->kp[kap].k[0], wf->kp[kap].k[1], wf->kp[kap].k[2]};
    frac_to_cart(kc, wf->reclattice);                                           // projector.c:288-292
    std::shared_ptr<PrunedPlan> plan = get_pruned_plan(wf, kap, fftg);
    const bool pruned = plan->ok && plan->g.col_run && !plan->g.three && !getenv("PAWB200_DENSITY_GENERIC");
    // one site per launch, no atomics: sums over overlapping spheres are formed in site order (deterministic)
    auto add_sites = [&](double2* x, int b_first, int nbx, int interleaved) {
      for (int c0 = 0; c0 < nbx && T->total_pts; c0 += NBMAX) {
        const int nc = std::min(NBMAX, nbx - c0);
        for (int s = 0; s < nlist; s++) {
          const int npts = T->host[s].npts;
          if (!npts) continue;
          aug_freq_add_site_kernel<<<(npts + 255) / 256, 256, sizeof(double2) * nc * maxlm, g_stream>>>(
              T->sites.as<SiteDev>(), s, full_off[s], T->idx.as<int>(), T->path.as<double>(), T->total_pts,
              T->table.as<double2>(), wf->P[kap].as<double2>() + (long)(b_first + c0) * wf->ldp, wf->ldp, nc,
              interleaved ? x + (long)(c0 / FFT_B) * ngrid * FFT_B : x + (long)c0 * ngrid, ngrid, kc[0], kc[1], kc[2],
              interleaved);
          count_launch();
        }
        check_launch();
      }
    };
    if (pruned) {
      // hand-written forward transform, pruned on the output side (fft_fwd_pass_*): augmentation built straight in
      // the interleaved box layout, coefficients come back interleaved and are transposed into band rows
      const FftGeom& g = plan->g;
      const size_t grp_box = (size_t)FFT_B * ngrid * sizeof(double2);
      const int gmax = (int)std::max<size_t>(1, std::min<size_t>(8, fft_budget_bytes() / grp_box));
      const size_t t1_grp = (size_t)g.ncol * g.n3 * FFT_B * sizeof(double2);
      const size_t t2_grp = (size_t)g.nplane * g.n2 * g.n3 * FFT_B * sizeof(double2);
      const long ldil = ((long)npw + 1) / 2 * 2;
      DevBuf cil((size_t)gmax * ldil * FFT_B * sizeof(float2));
      const int ngroups = (wf->nband + FFT_B - 1) / FFT_B;
      for (int g0 = 0; g0 < ngroups; g0 += gmax) {
        const int ng = std::min(gmax, ngroups - g0);
        const int b0 = g0 * FFT_B, nb = std::min(wf->nband - b0, ng * FFT_B);
        g_grid.ensure((size_t)ng * grp_box);
        g_fft_t1.ensure(t1_grp * ng);
        g_fft_t2.ensure(t2_grp * ng);
        double2* x = g_grid.as<double2>();
        {
          ScopedStage tm(ST_AUGMENT);
          CUDA_OK(cudaMemsetAsync(x, 0, (size_t)ng * grp_box, g_stream));
          add_sites(x, b0, nb, 1);
        }
        FftWork w;
        w.T1 = g_fft_t1.as<double2>();
        w.T2 = g_fft_t2.as<double2>();
        {
          ScopedStage tm(ST_FFT);
          g_boxes_fft += nb;
          count_launch(launch_pruned_forward(g, ng, x, w, cil.as<float2>(), ldil, scale, g_num_sms, g_stream));
        }
        ScopedStage tm(ST_SCATTER);
        dim3 grid((npw + 127) / 128, ng);
        deinterleave_coeff_kernel<<<grid, 256, 0, g_stream>>>(cil.as<float2>(), ldil, npw, b0, nb,
                                                             wf->CA[kap].as<float2>(), wf->ldc[kap]);
        count_launch();
        check_launch();
      }
      continue;
    }
    for (int b0 = 0; b0 < wf->nband; b0 += batch) {
      const int nb = std::min(batch, wf->nband - b0);
      g_grid.ensure((size_t)nb * ngrid * sizeof(double2));
      double2* x = g_grid.as<double2>();
      {
        ScopedStage tm(ST_AUGMENT);
        CUDA_OK(cudaMemsetAsync(x, 0, (size_t)nb * ngrid * sizeof(double2), g_stream));
        add_sites(x, b0, nb, 0);
      }
      launch_fft(x, fftg, nb, CUFFT_FORWARD);
      ScopedStage tm(ST_SCATTER);
      dim3 grid((npw + 255) / 256, nb);
      gather_pw_batch_kernel<<<grid, 256, 0, g_stream>>>(x, ngrid, dg.as<int>(),
                                                         wf->CA[kap].as<float2>() + (long)b0 * wf->ldc[kap],
                                                         wf->ldc[kap], npw, scale);
      count_launch();
      check_launch();
    }
  }
}

void pawb200_overlap_setup_real(pawb200_pswf_t* wf_R, pawb200_pswf_t* wf_S, const int* labels_R,
                                const int* labels_S, const double* coords_R, const double* coords_S,
                                const int* N_R, const int* N_S, const int* N_RS_R, const int* N_RS_S,
                                int num_N_R, int num_N_S, int num_N_RS) {
  API_BEGIN
  HostSection hs_("overlap_setup_real");
  require_device();
  check_pair(wf_S, wf_R);
  if (!wf_R->has_projections || !wf_S->has_projections) throw std::runtime_error("setup_projections has not been run");
  check_overlap_setup_args(wf_R, wf_S, labels_R, labels_S, coords_R, coords_S, N_R, N_S, N_RS_R, N_RS_S, num_N_R,
                           num_N_S, num_N_RS);
  wf_R->gen++;
  wf_S->gen++;
  wf_R->aug_cache.valid = wf_S->aug_cache.valid = false;
  prelaunch_pseudo(wf_S, wf_R);        // asynchronous ingest: the pseudo GEMMs go out before the projection work
  wf_R->W.clear(); wf_S->W.clear();
  wf_R->wp_num = num_N_S; wf_S->wp_num = num_N_R;
  wf_R->wp_nlm.clear(); wf_S->wp_nlm.clear();
  const auto& elsR = wf_R->pps->list.el;
  const auto& elsS = wf_S->pps->list.el;
  // part 1 (projector.c:625-646): filtered partial waves of R's N_R sites on S's grid, all S bands
  std::unique_ptr<SiteTables> T_NR, T_NS;
  if (num_N_R > 0) {
    T_NR = build_site_tables(elsR, N_R, num_N_R, labels_R, coords_R, wf_S->lattice, wf_S->fftg, 1, false);
    for (auto& sd : T_NR->host) wf_S->wp_nlm.push_back(sd.nlm);
  }
  if (num_N_S > 0) {     // part 2 (:649-671)
    T_NS = build_site_tables(elsS, N_S, num_N_S, labels_S, coords_S, wf_R->lattice, wf_R->fftg, 1, false);
    for (auto& sd : T_NS->host) wf_R->wp_nlm.push_back(sd.nlm);
  }
  // One pass over the boxes of each wavefunction for P (if deferred) and W.  The wavefunction that was read first
  // goes first: with asynchronous ingest its coefficients are the ones already arriving, the other one's copies
  // are queued behind them.
  auto pass_S = [&] { require_projections(wf_S, T_NR.get(), T_NR ? &wf_S->W : nullptr, T_NR ? &wf_S->ldw : nullptr); };
  auto pass_R = [&] { require_projections(wf_R, T_NS.get(), T_NS ? &wf_R->W : nullptr, T_NS ? &wf_R->ldw : nullptr); };
  if (wf_R->id <= wf_S->id) { pass_R(); pass_S(); } else { pass_S(); pass_R(); }
  setup_offsite(wf_R, wf_S, labels_R, labels_S, coords_R, coords_S, N_RS_R, N_RS_S, num_N_RS);
  wf_R->recip_setup = wf_S->recip_setup = false;
  wf_R->CA.clear(); wf_S->CA.clear();
  wf_S->overlap_partner = wf_R->id;
  API_END_VOID
}

void pawb200_compensation_terms(pawb200_c128* overlap, int BAND_NUM, pawb200_pswf_t* wf_S, pawb200_pswf_t* wf_R,
                                int num_M, int num_N_R, int num_N_S, int num_N_RS, const int* M_R, const int* M_S,
                                const int* N_R, const int* N_S, const int* N_RS_R, const int* N_RS_S,
                                const int*, const double*, const int*, const double*, const int*, int spin_flip) {
  API_BEGIN
  require_device();
  check_pair(wf_S, wf_R);
  if (BAND_NUM < 0 || BAND_NUM >= wf_S->nband) throw std::runtime_error("band index out of range");
  SiteLists L = make_lists(num_M, num_N_R, num_N_S, num_N_RS, M_R, M_S, N_R, N_S, N_RS_R, N_RS_S);
  HostMatrixCache& c = wf_S->aug_cache;
  const int NK = wf_R->nkappa(), nS = wf_S->nband, nR = wf_R->nband;
  const size_t lh = L.hash();
  if (!c.valid || c.other_id != wf_R->id || c.other_gen != wf_R->gen || c.self_gen != wf_S->gen ||
      c.flip != (spin_flip ? 1 : 0) || c.list_hash != lh || c.recip != 0) {
    c.data.assign((size_t)NK * nS * nR, cdouble(0, 0));
    overlap_matrix(wf_S, wf_R, &L, spin_flip, 0, NK, false, true, c.data.data());
    c.other_id = wf_R->id; c.other_gen = wf_R->gen; c.self_gen = wf_S->gen;
    c.flip = spin_flip ? 1 : 0; c.list_hash = lh; c.recip = 0; c.valid = true;
  }
  cdouble* out = (cdouble*)overlap;
  for (int k = 0; k < NK; k++)
    for (int b = 0; b < nR; b++) out[(size_t)b * NK + k] += c.data[((size_t)k * nS + BAND_NUM) * nR + b];
  API_END_VOID
}

// ---- (f3) method "aug_recip": projector.h:114-121, 130-137; projector.c:727-848, 965-1077 -------------------
void pawb200_overlap_setup_recip(pawb200_pswf_t* wf_R, pawb200_pswf_t* wf_S, const int* labels_R,
                                 const int* labels_S, const double* coords_R, const double* coords_S,
                                 const int* N_R, const int* N_S, const int* N_RS_R, const int* N_RS_S,
                                 int num_N_R, int num_N_S, int num_N_RS) {
  API_BEGIN
  HostSection hs_("overlap_setup_recip");
  require_device();
  check_pair(wf_S, wf_R);
  if (!wf_R->has_projections || !wf_S->has_projections) throw std::runtime_error("setup_projections has not been run");
  check_overlap_setup_args(wf_R, wf_S, labels_R, labels_S, coords_R, coords_S, N_R, N_S, N_RS_R, N_RS_S, num_N_R,
                           num_N_S, num_N_RS);
  wf_R->gen++;
  wf_S->gen++;
  wf_R->aug_cache.valid = wf_S->aug_cache.valid = false;
  wf_R->W.clear(); wf_S->W.clear();
  wf_R->wp_nlm.clear(); wf_S->wp_nlm.clear();
  wf_R->wp_num = num_N_S; wf_S->wp_num = num_N_R;    // projector.c:735-736
  wf_R->CA.clear(); wf_S->CA.clear();
  if (wf_R->band_sharded() || wf_S->band_sharded())
    throw std::runtime_error("method aug_recip is not available on band-sharded wavefunctions (use aug_real)");
  require_projections(wf_R);
  require_projections(wf_S);
  if (num_N_R > 0) compute_aug_freqs(wf_R, N_R, num_N_R, labels_R, coords_R);   // part 1 (:748-767)
  if (num_N_S > 0) compute_aug_freqs(wf_S, N_S, num_N_S, labels_S, coords_S);   // part 2 (:770-792)
  setup_offsite(wf_R, wf_S, labels_R, labels_S, coords_R, coords_S, N_RS_R, N_RS_S, num_N_RS);
  wf_R->recip_setup = wf_S->recip_setup = true;
  wf_S->overlap_partner = wf_R->id;
  API_END_VOID
}

void pawb200_compensation_terms_recip(pawb200_c128* overlap, int BAND_NUM, pawb200_pswf_t* wf_S,
                                      pawb200_pswf_t* wf_R, int num_M, int num_N_R, int num_N_S, int num_N_RS,
                                      const int* M_R, const int* M_S, const int* N_R, const int* N_S,
                                      const int* N_RS_R, const int* N_RS_S, const int*, const double*, const int*,
                                      const double*, const int*, int spin_flip) {
  API_BEGIN
  require_device();
  check_pair(wf_S, wf_R);
  if (BAND_NUM < 0 || BAND_NUM >= wf_S->nband) throw std::runtime_error("band index out of range");
  SiteLists L = make_lists(num_M, num_N_R, num_N_S, num_N_RS, M_R, M_S, N_R, N_S, N_RS_R, N_RS_S);
  HostMatrixCache& c = wf_S->aug_cache;
  const int NK = wf_R->nkappa(), nS = wf_S->nband, nR = wf_R->nband;
  const size_t lh = L.hash();
  if (!c.valid || c.other_id != wf_R->id || c.other_gen != wf_R->gen || c.self_gen != wf_S->gen ||
      c.flip != (spin_flip ? 1 : 0) || c.list_hash != lh || c.recip != 1) {
    c.data.assign((size_t)NK * nS * nR, cdouble(0, 0));
    overlap_matrix(wf_S, wf_R, &L, spin_flip, 0, NK, false, true, c.data.data(), true);
    c.other_id = wf_R->id; c.other_gen = wf_R->gen; c.self_gen = wf_S->gen;
    c.flip = spin_flip ? 1 : 0; c.list_hash = lh; c.recip = 1; c.valid = true;
  }
  cdouble* out = (cdouble*)overlap;
  for (int k = 0; k < NK; k++)
    for (int b = 0; b < nR; b++) out[(size_t)b * NK + k] += c.data[((size_t)k * nS + BAND_NUM) * nR + b];
  API_END_VOID
}

// ---- (f2) k-point desymmetrisation: utils.h:392-393, utils.c:829-1098 -------------------------------------
pawb200_pswf_t* pawb200_expand_symm_wf(pawb200_pswf_t* rwf, int num_kpts, const int* maps, const double* ops,
                                       const double* drs, const double* kws, const int* trs) {
  API_BEGIN
  require_device();
  if (!rwf || num_kpts <= 0) throw std::runtime_error("bad arguments to expand_symm_wf");
  if (rwf->ncl) throw std::runtime_error("desymmetrisation of noncollinear wavefunctions is not defined "
                                         "(NCLWavefunction.desymmetrized_copy raises in the reference too)");
  auto wf = std::make_unique<pawb200_pswf>();
  if (rwf->band_sharded()) throw std::runtime_error("desymmetrisation of a band-sharded wavefunction is not supported");
  wf->nspin = rwf->nspin; wf->nband = rwf->nband; wf->nwk = num_kpts; wf->encut = rwf->encut; wf->ncl = 0;
  wf->band_lo = 0; wf->band_hi = wf->nband; wf->band_rows = wf->nband;
  memcpy(wf->lattice, rwf->lattice, sizeof(wf->lattice));
  memcpy(wf->reclattice, rwf->reclattice, sizeof(wf->reclattice));
  const int NK = num_kpts * wf->nspin;
  wf->kp.resize(NK); wf->weight.resize(NK); wf->C.resize(NK); wf->ldc.assign(NK, 0);
  wf->resident.assign(NK, 0); wf->perm_dev.resize(NK);
  WavecarHeader hd;
  hd.encut = rwf->encut;
  memcpy(hd.lattice, rwf->lattice, sizeof(hd.lattice));
  memcpy(hd.reclattice, rwf->reclattice, sizeof(hd.reclattice));
  for (int d = 0; d < 3; d++) hd.nbmax[d] = rwf->G_bounds[2 * d + 1] - rwf->G_bounds[2 * d] + 2;   // utils.c:925-927
  for (int knum = 0; knum < NK; knum++) {
    const int kq = knum % num_kpts;
    if (maps[kq] < 0 || maps[kq] >= rwf->nwk) throw std::runtime_error("k-point map out of range");
    int rnum = maps[kq];
    const int tr = trs[kq];
    if (knum >= num_kpts && rwf->nspin == 2) rnum += rwf->nwk;
    const KPointInfo& rk = rwf->kp[rnum];
    KPointInfo& nk = wf->kp[knum];
    const double* op = ops + 9 * kq;
    const double* dr = drs + 3 * kq;
    for (int i = 0; i < 3; i++) nk.k[i] = op[3 * i] * rk.k[0] + op[3 * i + 1] * rk.k[1] + op[3 * i + 2] * rk.k[2];
    if (tr == 1) for (int i = 0; i < 3; i++) nk.k[i] *= -1;
    double kdiff[3];
    for (int i = 0; i < 3; i++) {
      kdiff[i] = std::round(nk.k[i]);
      nk.k[i] -= kdiff[i];
      if (std::fabs(nk.k[i] + 0.5) < 0.0001) {
        kdiff[i] -= 1;
        nk.k[i] += 1;
      }
    }
    wf->weight[knum] = kws[kq];
    nk.nplane = rk.nplane;
    nk.energy = rk.energy;
    nk.occ = rk.occ;
    nk.G = enumerate_g(hd, nk.k, wf->G_bounds);
    const int npw = (int)(nk.G.size() / 3);
    if (npw != rk.nplane)
      throw std::runtime_error("desymmetrised k-point has " + std::to_string(npw) + " plane waves, source has " +
                               std::to_string(rk.nplane) + " (operation is not a symmetry of the reciprocal lattice?)");
    box_order(nk);
    // G -> index map of the new list, then source index + phase per new plane wave  (utils.c:994-1050)
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int w = 0; w < npw; w++)
      for (int d = 0; d < 3; d++) {
        lo[d] = std::min(lo[d], (int)nk.G[3 * w + d]);
        hi[d] = std::max(hi[d], (int)nk.G[3 * w + d]);
      }
    const int ng[3] = {hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1};
    std::vector<int> kptinds((size_t)ng[0] * ng[1] * ng[2], -1);
    for (int w = 0; w < npw; w++)
      kptinds[((size_t)(nk.G[3 * w] - lo[0]) * ng[1] + (nk.G[3 * w + 1] - lo[1])) * ng[2] + (nk.G[3 * w + 2] - lo[2])] = w;
    std::vector<int> gmaps(npw, -1);
    std::vector<float2> factors(npw);
    for (int g = 0; g < npw; g++) {
      double pw[3] = {(double)rk.G[3 * g], (double)rk.G[3 * g + 1], (double)rk.G[3 * g + 2]};
      double r[3];
      for (int i = 0; i < 3; i++) r[i] = op[3 * i] * pw[0] + op[3 * i + 1] * pw[1] + op[3 * i + 2] * pw[2];
      for (int i = 0; i < 3; i++) pw[i] = (tr == 1 ? -r[i] : r[i]) + kdiff[i];
      const int gx = (int)std::round(pw[0]), gy = (int)std::round(pw[1]), gz = (int)std::round(pw[2]);
      int w = -1;
      if (gx >= lo[0] && gx <= hi[0] && gy >= lo[1] && gy <= hi[1] && gz >= lo[2] && gz <= hi[2])
        w = kptinds[((size_t)(gx - lo[0]) * ng[1] + (gy - lo[1])) * ng[2] + (gz - lo[2])];
      if (w < 0) throw std::runtime_error("bad plane-wave mapping in expand_symm_wf (operation does not map the basis onto itself)");
      gmaps[w] = g;
      const double ph = nk.k[0] * dr[0] + nk.k[1] * dr[1] + nk.k[2] * dr[2] + pw[0] * dr[0] + pw[1] * dr[1] + pw[2] * dr[2];
      // cexpf of the double argument converted to float complex, like utils.c:1040-1044
      const std::complex<float> f = std::exp(std::complex<float>(0.0f, (float)((tr == 0 ? -1.0 : 1.0) * 2 * kPi * ph)));
      factors[w] = make_float2(f.real(), f.imag());
    }
    for (int w = 0; w < npw; w++)
      if (gmaps[w] < 0) throw std::runtime_error("incomplete plane-wave mapping in expand_symm_wf");
    if (!rwf->resident[rnum]) continue;
    wf->resident[knum] = 1;
    // storage (box) order on both sides: new position j holds file index perm[j]; its source sits at pos_old[...]
    std::vector<int> src(npw);
    std::vector<float2> fac(npw);
    for (int j = 0; j < npw; j++) {
      const int w = nk.perm[j];
      src[j] = rk.pos[gmaps[w]];
      fac[j] = factors[w];
    }
    const long ld = ((long)npw + 31) / 32 * 32;
    wf->ldc[knum] = ld;
    wf->C[knum].alloc((size_t)wf->nband * ld * sizeof(float2));
    wf->C[knum].zero((size_t)wf->nband * ld * sizeof(float2));
    wf->perm_dev[knum] = upload(nk.perm);
    DevBuf dsrc = upload(src), dfac = upload(fac);
    wait_coeffs(rwf, rnum, 0, rwf->nband);
    dim3 grid((npw + 255) / 256, std::min(wf->nband, 64));
    symm_map_kernel<<<grid, 256, 0, g_stream>>>(rwf->C[rnum].as<float2>(), rwf->ldc[rnum], wf->C[knum].as<float2>(), ld,
                                                wf->nband, npw, dsrc.as<int>(), dfac.as<float2>(), tr == 1 ? 1 : 0);
    count_launch();
    check_launch();
    alloc_interleaved(wf.get(), knum, g_stream);
    interleave_rows(wf.get(), knum, 0, wf->nband, g_stream);
  }
  return wf.release();
  API_END(nullptr)
}

// Plane-wave coefficients of (band, kappa) in WAVECAR file order (complex64, nplane entries) - test accessor.
int pawb200_get_coefficients(pawb200_pswf_t* wf, int band, int kappa, pawb200_c64* out) {
  API_BEGIN
  check_kpoint(wf, band, kappa);
  const KPointInfo& kp = wf->kp[kappa];
  const int hl = wf->npw_half(kappa);
  std::vector<float2> row(kp.nplane);
  wait_coeffs(wf, kappa, band, band + 1);
  CUDA_OK(cudaMemcpyAsync(row.data(), wf->C[kappa].as<float2>() + (long)band * wf->ldc[kappa],
                          (size_t)kp.nplane * sizeof(float2), cudaMemcpyDeviceToHost, g_stream));
  stream_sync();
  float2* o = (float2*)out;
  for (int h = 0; h < wf->halves(); h++)
    for (int j = 0; j < hl; j++) o[h * hl + kp.perm[j]] = row[h * hl + j];
  return kp.nplane;
  API_END(-1)
}

int pawb200_get_kpoint(pawb200_pswf_t* wf, int kappa, double* k3, double* weight) {
  if (!wf || kappa < 0 || kappa >= wf->nkappa()) return -1;
  for (int i = 0; i < 3; i++) k3[i] = wf->kp[kappa].k[i];
  if (weight) *weight = wf->weight[kappa];
  return wf->kp[kappa].nplane;
}

// ---- real space ---------------------------------------------------------------------------------
void pawb200_realspace_state(pawb200_c128* x, int BAND_NUM, int KPOINT_NUM, pawb200_pswf_t* wf, const int* fftg,
                             const int* labels, const double* coords) {
  API_BEGIN
  if (wf && wf->ncl) throw std::runtime_error("noncollinear wavefunction: use pawb200_ncl_realspace_state");
  state_to_host(x, BAND_NUM, KPOINT_NUM, wf, fftg, labels, coords);
  API_END_VOID
}
void pawb200_ncl_realspace_state(pawb200_c128* x, int BAND_NUM, int KPOINT_NUM, pawb200_pswf_t* wf,
                                 const int* fftg, const int* labels, const double* coords) {
  API_BEGIN
  if (wf && !wf->ncl) throw std::runtime_error("collinear wavefunction: use pawb200_realspace_state");
  state_to_host(x, BAND_NUM, KPOINT_NUM, wf, fftg, labels, coords);
  API_END_VOID
}
void pawb200_remove_phase(pawb200_c128* x, int KPOINT_NUM, pawb200_pswf_t* wf, const int* fftg) {
  API_BEGIN
  require_device();
  if (!wf || KPOINT_NUM < 0 || KPOINT_NUM >= wf->nkappa()) throw std::runtime_error("k-point index out of range");
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  g_grid.ensure(ngrid * sizeof(double2));
  CUDA_OK(cudaMemcpyAsync(g_grid.p, x, ngrid * sizeof(double2), cudaMemcpyHostToDevice, g_stream));
  const double* k = wf->kp[KPOINT_NUM].k;
  const long blocks = std::min<long>((ngrid + 255) / 256, (long)g_num_sms * 16);
  bloch_phase_kernel<<<(unsigned)blocks, 256, 0, g_stream>>>(g_grid.as<double2>(), fftg[0], fftg[1], fftg[2], k[0],
                                                             k[1], k[2], -1.0, 1);
  count_launch();
  check_launch();
  CUDA_OK(cudaMemcpyAsync(x, g_grid.p, ngrid * sizeof(double2), cudaMemcpyDeviceToHost, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
  API_END_VOID
}
void pawb200_ae_state_density(double* P, int BAND_NUM, int KPOINT_NUM, pawb200_pswf_t* wf, const int* fftg,
                              const int* labels, const double* coords) {
  API_BEGIN
  check_kpoint(wf, BAND_NUM, KPOINT_NUM);
  density_to_host(P, wf, fftg, labels, coords, BAND_NUM, KPOINT_NUM);
  API_END_VOID
}
void pawb200_ae_chg_density(double* P, pawb200_pswf_t* wf, const int* fftg, const int* labels, const double* coords) {
  API_BEGIN
  density_to_host(P, wf, fftg, labels, coords, -1, -1);
  API_END_VOID
}

// Band shard of ae_chg_density: only the occupied bands in [band_lo, band_hi) contribute (same weights), so that
// ranks holding the same wavefunction can split the bands and sum their grids (one all-reduce).
void pawb200_ae_chg_density_bands(double* P, pawb200_pswf_t* wf, const int* fftg, const int* labels,
                                  const double* coords, int band_lo, int band_hi) {
  API_BEGIN
  if (wf && wf->ncl) throw std::runtime_error("band-sharded density is implemented for collinear wavefunctions");
  density_to_host(P, wf, fftg, labels, coords, -1, -1, band_lo, band_hi);
  API_END_VOID
}
void pawb200_ncl_ae_chg_density(double* P, pawb200_pswf_t* wf, const int* fftg, const int* labels,
                                const double* coords) {
  API_BEGIN
  density_to_host(P, wf, fftg, labels, coords, -1, -1);
  API_END_VOID
}

// density.c:205-230 (Projector method 'realspace'): AE states on the grid, brute-force overlap
void pawb200_project_realspace_state(pawb200_c128* projs, int BAND_NUM, pawb200_pswf_t* wf, pawb200_pswf_t* wf_R,
                                     const int* fftg, const int* labels, const double* coords, const int* labels_R,
                                     const double* coords_R) {
  API_BEGIN
  require_device();
  check_pair(wf, wf_R);
  if (BAND_NUM < 0 || BAND_NUM >= wf->nband) throw std::runtime_error("band index out of range");
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  const int NK = wf->nkappa(), nR = wf_R->nband;
  SiteTables& TS = ae_tables(wf, fftg, labels, coords);
  SiteTables& TR = ae_tables(wf_R, fftg, labels_R, coords_R);
  DevBuf state(ngrid * sizeof(double2));
  const double scale = determinant3(wf->lattice) / (double)ngrid;   // density.c:224
  long batch = std::max<long>(1, (long)(fft_budget_bytes() / (sizeof(double2) * ngrid)));
  batch = std::min<long>(batch, 64);
  const int nchunk = 64;
  DevBuf partial((size_t)batch * nchunk * sizeof(double2)), dres((size_t)nR * sizeof(double2));
  std::vector<cdouble> host(nR);
  cdouble* out = (cdouble*)projs;
  for (int kap = 0; kap < NK; kap++) {
    if (!wf->resident[kap] || !wf_R->resident[kap]) {
      for (int b = 0; b < nR; b++) out[(size_t)b * NK + kap] = cdouble(0, 0);
      continue;
    }
    std::shared_ptr<PrunedPlan> planS = get_pruned_plan(wf, kap, fftg), planR = get_pruned_plan(wf_R, kap, fftg);
    if (planS->ok && planR->ok && !getenv("PAWB200_DENSITY_GENERIC")) {
      // both sides through the hand-written pruned transform; the basis states stay in the interleaved layout
      realspace_boxes_il(wf, kap, *planS, BAND_NUM, 1, fftg, TS);
      extract_il_kernel<<<(unsigned)std::min<long>((ngrid + 255) / 256, (long)g_num_sms * 16), 256, 0, g_stream>>>(
          g_grid.as<double2>(), ngrid, 0, 1, state.as<double2>());
      count_launch();
      check_launch();
      const int gbatch = (int)std::max<long>(1, std::min<long>(4, batch / FFT_B));
      DevBuf partial_il((size_t)gbatch * FFT_B * nchunk * sizeof(double2));
      for (int b0 = 0; b0 < nR; b0 += gbatch * FFT_B) {
        const int nb = std::min(gbatch * FFT_B, nR - b0), ng = (nb + FFT_B - 1) / FFT_B;
        realspace_boxes_il(wf_R, kap, *planR, b0, nb, fftg, TR);
        ScopedStage tm(ST_AUGMENT);
        grid_dot_partial_il_kernel<<<dim3(nchunk, ng), 256, 0, g_stream>>>(g_grid.as<double2>(), state.as<double2>(),
                                                                           ngrid, nchunk, partial_il.as<double2>());
        grid_dot_final_kernel<<<(nb + 127) / 128, 128, 0, g_stream>>>(partial_il.as<double2>(), nchunk, nb, scale,
                                                                      dres.as<double2>() + b0);
        count_launch(2);
        check_launch();
      }
      CUDA_OK(cudaMemcpyAsync(host.data(), dres.p, (size_t)nR * sizeof(double2), cudaMemcpyDeviceToHost, g_stream));
      stream_sync();
      for (int b = 0; b < nR; b++) out[(size_t)b * NK + kap] = host[b];
      continue;
    }
    {
      DevBuf inv = build_inverse_map(wf, kap, fftg);
      realspace_boxes(wf, kap, BAND_NUM, 1, fftg, TS, inv);
      CUDA_OK(cudaMemcpyAsync(state.p, g_grid.p, ngrid * sizeof(double2), cudaMemcpyDeviceToDevice, g_stream));
    }
    DevBuf invR = build_inverse_map(wf_R, kap, fftg);
    for (int b0 = 0; b0 < nR; b0 += (int)batch) {
      const int nb = (int)std::min<long>(batch, nR - b0);
      realspace_boxes(wf_R, kap, b0, nb, fftg, TR, invR);
      ScopedStage tm(ST_AUGMENT);
      grid_dot_partial_kernel<<<dim3(nchunk, nb), 256, 0, g_stream>>>(g_grid.as<double2>(), state.as<double2>(), ngrid,
                                                                      nchunk, partial.as<double2>());
      grid_dot_final_kernel<<<(nb + 127) / 128, 128, 0, g_stream>>>(partial.as<double2>(), nchunk, nb, scale,
                                                                    dres.as<double2>() + b0);
      count_launch(2);
      check_launch();
    }
    CUDA_OK(cudaMemcpyAsync(host.data(), dres.p, (size_t)nR * sizeof(double2), cudaMemcpyDeviceToHost, g_stream));
    stream_sync();
    for (int b = 0; b < nR; b++) out[(size_t)b * NK + kap] = host[b];
  }
  API_END_VOID
}

void pawb200_write_volumetric(const char* filename, const double* x, const int* fftg, double scale) {
  API_BEGIN
  // density.c:461-477: x fastest, z slowest, "%E   " five per line.  Host I/O by design.
  FILE* fp = fopen(filename, "w");
  if (!fp) throw std::runtime_error(std::string("cannot open ") + filename);
  std::vector<char> buf(1 << 20);
  setvbuf(fp, buf.data(), _IOFBF, buf.size());
  long t = 1;
  for (int k = 0; k < fftg[2]; k++)
    for (int j = 0; j < fftg[1]; j++)
      for (int i = 0; i < fftg[0]; i++) {
        fprintf(fp, "%E   ", x[((long)i * fftg[1] + j) * fftg[2] + k] * scale);
        if (t % 5 == 0) fputc('\n', fp);
        t++;
      }
  fclose(fp);
  API_END_VOID
}

// ---- single-band FFT entry points ---------------------------------------------------------------
namespace {
// Plan of a caller-supplied G list (the single-band fft3d / fwd_fft3d entry points): box-order permutation + pruned
// geometry.  ok == false -> the caller falls back to scatter + cuFFT (unsupported radix, aliased grid, or a grid too
// large for a whole 16-slot interleave group).
std::shared_ptr<PrunedPlan> plan_for_g_list(const int* Gs, int num_waves, const int* fftg, KPointInfo& kp) {
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  if (num_waves <= 0 || (size_t)ngrid * FFT_B * sizeof(double2) > ((size_t)2 << 30)) return std::make_shared<PrunedPlan>();
  kp.nplane = num_waves;
  kp.G.assign(Gs, Gs + 3 * (size_t)num_waves);
  box_order(kp);
  return build_pruned_plan_kp(kp, num_waves, fftg);
}
void ensure_group_scratch(const FftGeom& g) {
  g_fft_t1.ensure((size_t)g.ncol * g.n3 * FFT_B * sizeof(double2));
  g_fft_t2.ensure((size_t)g.nplane * g.n2 * g.n3 * FFT_B * sizeof(double2));
}
}  // namespace

void pawb200_fft3d(pawb200_c128* x, const int*, const double* lattice, const double*, const int* Gs,
                   const pawb200_c64* Cs, int num_waves, const int* fftg) {
  API_BEGIN
  require_device();
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  {
    KPointInfo kp;
    std::shared_ptr<PrunedPlan> plan = plan_for_g_list(Gs, num_waves, fftg, kp);
    if (plan->ok && plan->g.col_run) {
      // slot 0 of one interleave group carries the band; the hand-written pruned transform does the rest
      const long ldil = ((long)num_waves + 1) / 2 * 2;
      std::vector<float2> il((size_t)ldil * FFT_B, make_float2(0.f, 0.f));
      const float2* C = reinterpret_cast<const float2*>(Cs);
      for (int j = 0; j < num_waves; j++) il[(size_t)j * FFT_B] = C[kp.perm[j]];
      DevBuf dil = upload(il);
      const FftGeom& g = plan->g;
      ensure_group_scratch(g);
      g_grid.ensure((size_t)ngrid * FFT_B * sizeof(double2));
      FftInput in;
      in.Cil = dil.as<float2>(); in.ldil = ldil; in.C = nullptr; in.ldc = 0; in.halves = 1; in.half_len = num_waves;
      FftWork w;
      w.T1 = g_fft_t1.as<double2>(); w.T2 = g_fft_t2.as<double2>();
      const double scale = std::pow(determinant3(lattice), -0.5);
      {
        ScopedStage tm(ST_FFT);
        g_boxes_fft += 1;
        count_launch(launch_pruned_passes(g, in, 0, 1, 1, scale, w, nullptr, g_grid.as<double2>(), g_num_sms, g_stream,
                                          plan->max_plane_cols));
      }
      DevBuf planar((size_t)ngrid * sizeof(double2));
      extract_il_kernel<<<(unsigned)std::min<long>((ngrid + 255) / 256, (long)g_num_sms * 16), 256, 0, g_stream>>>(
          g_grid.as<double2>(), ngrid, 0, 1, planar.as<double2>());
      count_launch();
      check_launch();
      CUDA_OK(cudaMemcpyAsync(x, planar.p, ngrid * sizeof(double2), cudaMemcpyDeviceToHost, g_stream));
      stream_sync();
      return;
    }
  }
  std::vector<int> inv(ngrid, -1);
  for (int w = 0; w < num_waves; w++) {
    const long g1 = (Gs[3 * w] + fftg[0]) % fftg[0], g2 = (Gs[3 * w + 1] + fftg[1]) % fftg[1],
               g3 = (Gs[3 * w + 2] + fftg[2]) % fftg[2];
    if (g1 < 0 || g2 < 0 || g3 < 0) throw std::runtime_error("FFT grid smaller than the G range");
    inv[(g1 * fftg[1] + g2) * fftg[2] + g3] = w;
  }
  DevBuf dinv = upload(inv);
  DevBuf dC((size_t)std::max(num_waves, 1) * sizeof(float2));
  CUDA_OK(cudaMemcpyAsync(dC.p, Cs, (size_t)num_waves * sizeof(float2), cudaMemcpyHostToDevice, g_stream));
  g_grid.ensure(ngrid * sizeof(double2));
  const double scale = std::pow(determinant3(lattice), -0.5);
  const long blocks = std::min<long>((ngrid + 255) / 256, (long)g_num_sms * 16);
  scatter_pw_kernel<<<(unsigned)blocks, 256, 0, g_stream>>>(dC.as<float2>(), 0, 0, 1, num_waves, dinv.as<int>(),
                                                            g_grid.as<double2>(), ngrid, 1, scale);
  count_launch();
  check_launch();
  launch_fft(g_grid.as<double2>(), fftg, 1, CUFFT_INVERSE);
  CUDA_OK(cudaMemcpyAsync(x, g_grid.p, ngrid * sizeof(double2), cudaMemcpyDeviceToHost, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
  API_END_VOID
}

void pawb200_fwd_fft3d(pawb200_c128* x, const int*, const double* lattice, const double*, const int* Gs,
                       pawb200_c64* Cs, int num_waves, const int* fftg) {
  API_BEGIN
  require_device();
  const long ngrid = (long)fftg[0] * fftg[1] * fftg[2];
  {
    KPointInfo kp;
    std::shared_ptr<PrunedPlan> plan = plan_for_g_list(Gs, num_waves, fftg, kp);
    if (plan->ok && plan->g.col_run && !plan->g.three) {
      // forward transform pruned on the output side (fft_fwd_pass_*): the box goes into slot 0 of an interleave group
      const FftGeom& g = plan->g;
      ensure_group_scratch(g);
      g_grid.ensure((size_t)ngrid * FFT_B * sizeof(double2));
      DevBuf planar((size_t)ngrid * sizeof(double2));
      CUDA_OK(cudaMemcpyAsync(planar.p, x, ngrid * sizeof(double2), cudaMemcpyHostToDevice, g_stream));
      CUDA_OK(cudaMemsetAsync(g_grid.p, 0, (size_t)ngrid * FFT_B * sizeof(double2), g_stream));
      insert_il_kernel<<<(unsigned)std::min<long>((ngrid + 255) / 256, (long)g_num_sms * 16), 256, 0, g_stream>>>(
          planar.as<double2>(), ngrid, 0, 1, g_grid.as<double2>());
      count_launch();
      check_launch();
      const long ldil = ((long)num_waves + 1) / 2 * 2;
      DevBuf dil((size_t)ldil * FFT_B * sizeof(float2));
      FftWork w;
      w.T1 = g_fft_t1.as<double2>(); w.T2 = g_fft_t2.as<double2>();
      const double scale = std::pow(determinant3(lattice), 0.5) / fftg[0] / fftg[1] / fftg[2];   // linalg.c:64-65
      {
        ScopedStage tm(ST_FFT);
        g_boxes_fft += 1;
        count_launch(launch_pruned_forward(g, 1, g_grid.as<double2>(), w, dil.as<float2>(), ldil, scale, g_num_sms,
                                           g_stream));
      }
      std::vector<float2> il((size_t)ldil * FFT_B);
      CUDA_OK(cudaMemcpyAsync(il.data(), dil.p, il.size() * sizeof(float2), cudaMemcpyDeviceToHost, g_stream));
      stream_sync();
      float2* C = reinterpret_cast<float2*>(Cs);
      for (int j = 0; j < num_waves; j++) C[kp.perm[j]] = il[(size_t)j * FFT_B];
      return;
    }
  }
  std::vector<int> gi(num_waves);
  for (int w = 0; w < num_waves; w++) {
    const long g1 = (Gs[3 * w] + fftg[0]) % fftg[0], g2 = (Gs[3 * w + 1] + fftg[1]) % fftg[1],
               g3 = (Gs[3 * w + 2] + fftg[2]) % fftg[2];
    gi[w] = (int)((g1 * fftg[1] + g2) * fftg[2] + g3);
  }
  DevBuf dgi = upload(gi);
  DevBuf dC((size_t)std::max(num_waves, 1) * sizeof(float2));
  g_grid.ensure(ngrid * sizeof(double2));
  CUDA_OK(cudaMemcpyAsync(g_grid.p, x, ngrid * sizeof(double2), cudaMemcpyHostToDevice, g_stream));
  launch_fft(g_grid.as<double2>(), fftg, 1, CUFFT_FORWARD);
  const double scale = std::pow(determinant3(lattice), 0.5) / fftg[0] / fftg[1] / fftg[2];   // linalg.c:64-65
  gather_pw_kernel<<<(num_waves + 255) / 256, 256, 0, g_stream>>>(g_grid.as<double2>(), dgi.as<int>(),
                                                                  dC.as<float2>(), num_waves, scale);
  count_launch();
  check_launch();
  // like the in-place DFTI transform, x holds the (scaled) spectrum afterwards
  CUDA_OK(cudaMemcpyAsync(Cs, dC.p, (size_t)num_waves * sizeof(float2), cudaMemcpyDeviceToHost, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
  API_END_VOID
}

// ---- utilities -----------------------------------------------------------------------------------
double pawb200_legendre(int l, int m, double x) { return assoc_legendre(l, m, x); }
void pawb200_Ylm(int l, int m, double theta, double phi, double* o) {
  const cdouble v = sph_harm(l, m, theta, phi);
  o[0] = v.real(); o[1] = v.imag();
}
void pawb200_Ylm2(int l, int m, double ct, double phi, double* o) {
  const cdouble v = sph_harm_cos(l, m, ct, phi);
  o[0] = v.real(); o[1] = v.imag();
}
void pawb200_frac_to_cartesian(double* c, const double* lattice) { frac_to_cart(c, lattice); }
void pawb200_cartesian_to_frac(double* c, const double* rec) { cart_to_frac(c, rec); }
double* pawb200_spline_coeff(const double* x, const double* y, int N) {
  Spline s = make_spline(x, y, N);
  double* o = (double*)malloc(sizeof(double) * 3 * N);
  for (int r = 0; r < 3; r++) std::copy(s.c[r].begin(), s.c[r].end(), o + (size_t)r * N);
  return o;
}
static Spline wrap_spline(const double* s3n, int n) {
  Spline s;
  for (int r = 0; r < 3; r++) s.c[r].assign(s3n + (size_t)r * n, s3n + (size_t)(r + 1) * n);
  return s;
}
double pawb200_proj_interpolate(double r, double rmax, int size, const double* x, const double* f, const double* s) {
  return eval_linear_grid(r, rmax, size, x, f, wrap_spline(s, size));
}
double pawb200_wave_interpolate(double r, int size, const double* x, const double* f, const double* s) {
  return eval_log_grid(r, size, x, f, wrap_spline(s, size));
}
double pawb200_spline_integral(const double* x, const double* a, const double* s, int size) {
  return spline_integrate(x, a, wrap_spline(s, size), size);
}
void pawb200_spherical_bessel_transform(double encut, int l, int N, const double* r, const double* f, double* k_out,
                                        double* fk_out) {
  API_BEGIN
  BesselTransform t(encut, 0, l, N, r);   // pawpyc.pyx:143-145
  std::vector<double> g = t.forward(f, l);
  std::copy(t.kgrid().begin(), t.kgrid().end(), k_out);
  std::copy(g.begin(), g.end(), fk_out);
  API_END_VOID
}
void pawb200_reciprocal_offsite_wave_overlap(const double* dcoord, const double* k1, const double* f1,
                                             const double* s1, int size1, const double* k2, const double* f2,
                                             const double* s2, int size2, int l1, int m1, int l2, int m2,
                                             double* o) {
  API_BEGIN
  const cdouble v = offsite_overlap_recip(dcoord, k1, f1, wrap_spline(s1, size1), size1, k2, f2,
                                          wrap_spline(s2, size2), size2, l1, m1, l2, m2);
  o[0] = v.real(); o[1] = v.imag();
  API_END_VOID
}

// ---- extensions ------------------------------------------------------------------------------------
void pawb200_set_kappa_range(pawb200_pswf_t* wf, int lo, int hi) {
  if (!wf) return;
  API_BEGIN
  // blocks outside the range are dropped: their HBM goes back to the pools (after every stream that may still
  // touch them has drained) and the per-band host caches, which hold full-range matrices, are invalidated
  bool any = false;
  for (int k = 0; k < wf->nkappa(); k++) any = any || ((k < lo || k >= hi) && wf->resident[k]);
  if (!any) return;
  // deferred ingest chunks of the dropped blocks must never be issued (their buffers are about to be released)
  g_pending_chunks.erase(std::remove_if(g_pending_chunks.begin(), g_pending_chunks.end(),
                                        [&](const ChunkJob& j) { return j.wf == wf && (j.kap < lo || j.kap >= hi); }),
                         g_pending_chunks.end());
  cudaStreamSynchronize(ingest_ring().copy);
  cudaStreamSynchronize(ingest_ring().unpack);
  if (g_stream2) cudaStreamSynchronize(g_stream2);
  cudaStreamSynchronize(g_stream);
  for (int k = 0; k < wf->nkappa(); k++) {
    if (!(k < lo || k >= hi) || !wf->resident[k]) continue;
    wf->resident[k] = 0;
    auto drop = [k](std::vector<DevBuf>& v) { if (k < (int)v.size()) v[k].release(); };
    drop(wf->C); drop(wf->Cil); drop(wf->perm_dev); drop(wf->P); drop(wf->W); drop(wf->CA); drop(wf->boxes);
  }
  wf->gen++;
  wf->pseudo_cache.valid = wf->aug_cache.valid = false;
  wf->pseudo_cache.data = std::vector<cdouble>();
  wf->aug_cache.data = std::vector<cdouble>();
  API_END_VOID
}
int pawb200_num_projections(pawb200_pswf_t* wf, int which) {
  if (!wf) return 0;
  if (which == 3) { int n = 0; for (int v : wf->wp_nlm) n += v; return n; }
  return wf->proj_sites ? wf->proj_sites->nproj : 0;
}
int pawb200_get_projections(pawb200_pswf_t* wf, int band, int kappa, int which, pawb200_c128* out) {
  API_BEGIN
  check_kpoint(wf, band, kappa);
  require_projections(wf);
  const int n = pawb200_num_projections(wf, which);
  const std::vector<DevBuf>& src = which == 3 ? wf->W : wf->P;
  const long ld = which == 3 ? wf->ldw : wf->ldp;
  if (src.empty() || !src[kappa].p) throw std::runtime_error("projections not available");
  int slot = band;
  if (wf->ncl) {
    if (which == 0) throw std::runtime_error("noncollinear: ask for up (1) or down (2) projections");
    slot = 2 * band + (which == 2 ? 1 : 0);
  }
  CUDA_OK(cudaMemcpy(out, src[kappa].as<double2>() + (long)slot * ld, (size_t)n * sizeof(double2), cudaMemcpyDeviceToHost));
  return n;
  API_END(-1)
}
int pawb200_get_channel_index(pawb200_pswf_t* wf, int* out) {
  if (!wf || !wf->proj_sites) return 0;
  int n = 0;
  for (int s = 0; s < wf->num_sites; s++) {
    const Element& el = wf->pps->list.el[wf->labels[s]];
    for (auto& c : el.chan) {
      if (out) { out[4 * n] = s; out[4 * n + 1] = c.n; out[4 * n + 2] = c.l; out[4 * n + 3] = c.m; }
      n++;
    }
  }
  return n;
}
int pawb200_get_site_indices(pawb200_pswf_t* wf, int site, int* out, int capacity) {
  if (!wf || !wf->proj_sites || site < 0 || site >= wf->num_sites) return -1;
  const auto& v = wf->proj_sites->host_idx[site];
  if (out) std::copy(v.begin(), v.begin() + std::min<size_t>(capacity, v.size()), out);
  return (int)v.size();
}
// Page-locked host buffers for results: device->host copies into them run at full PCIe rate and asynchronously
void* pawb200_alloc_pinned(size_t bytes) {
  API_BEGIN
  require_device();
  void* p = nullptr;
  CUDA_OK(cudaMallocHost(&p, std::max<size_t>(bytes, 1)));
  return p;
  API_END(nullptr)
}
void pawb200_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

void pawb200_get_timers(pawb200_timers* t) {
  drain_timers();
  trace_dump();
  if (getenv("PAWB200_PROFILE")) {
    for (auto& kv : g_hostprof.ms)
      fprintf(stderr, "[pawb200 host] %-22s %9.2f ms  (%ld calls)\n", kv.first.c_str(), kv.second,
              g_hostprof.calls[kv.first]);
  }
  t->h2d_ms = g_stage_ms[ST_H2D]; t->scatter_ms = g_stage_ms[ST_SCATTER]; t->fft_ms = g_stage_ms[ST_FFT];
  t->project_ms = g_stage_ms[ST_PROJECT]; t->table_ms = g_stage_ms[ST_TABLE];
  t->gemm_pseudo_ms = g_stage_ms[ST_GEMM_PS]; t->gemm_aug_ms = g_stage_ms[ST_GEMM_AUG];
  t->augment_ms = g_stage_ms[ST_AUGMENT]; t->d2h_ms = g_stage_ms[ST_D2H];
  t->launches = g_launches.load();
  t->boxes_scattered = g_boxes_scattered; t->boxes_fft = g_boxes_fft; t->slots_projected = g_slots_projected; t->sphere_samples = g_sphere_samples;
}
void pawb200_reset_timers(void) {
  drain_timers();
  g_hostprof.ms.clear();
  g_hostprof.calls.clear();
  for (auto& v : g_stage_ms) v = 0;
  g_launches = 0;
  g_boxes_scattered = g_boxes_fft = g_slots_projected = g_sphere_samples = 0;
}

}  // extern "C"

// =======================================================================================
// (f4) MomentumMatrix: momentum.h:56-92, momentum.c; pawpyc.pyx:738-807
// =======================================================================================
struct pawb200_density_ft {
  // per element: SBT of (phi_n1 phi_n2 - phit_n1 phit_n2)/r for L = |l1-l2| .. l1+l2 step 2 (momentum.c:116-140,
  // 225-276), as spline tables on the transform's k grid
  struct Elem {
    int N = 0;
    std::vector<double> ks;
    std::map<std::tuple<int, int, int>, int> slot;     // (n1, n2, L) -> slot
    std::vector<std::vector<double>> tab;              // per slot: f[N] + 3N spline coefficients
  };
  std::vector<Elem> el;
};

namespace {

Spline make_spline_v(const std::vector<double>& x, const std::vector<double>& f) {
  return make_spline(x.data(), f.data(), (int)x.size());
}

std::vector<double> pack_spline(const std::vector<double>& f, const Spline& sp) {
  const size_t n = f.size();
  std::vector<double> t(4 * n);
  for (size_t i = 0; i < n; i++) {
    t[i] = f[i];
    t[n + i] = sp.c[0][i];
    t[2 * n + i] = sp.c[1][i];
    t[3 * n + i] = sp.c[2][i];
  }
  return t;
}

// dense map over the wavefunction's G_bounds box: grid position -> storage (box-order) index of kappa's plane waves
struct BoxMap {
  int lo[3], dim[3];
  DevBuf map;
};
BoxMap build_box_map(const pawb200_pswf* wf, int kap) {
  BoxMap B;
  for (int d = 0; d < 3; d++) {
    B.lo[d] = wf->G_bounds[2 * d];
    B.dim[d] = wf->G_bounds[2 * d + 1] - wf->G_bounds[2 * d] + 1;
  }
  const KPointInfo& kp = wf->kp[kap];
  std::vector<int> m((size_t)B.dim[0] * B.dim[1] * B.dim[2], -1);
  for (int w = 0; w < kp.nplane; w++) {
    const int a = kp.G[3 * w] - B.lo[0], b = kp.G[3 * w + 1] - B.lo[1], c = kp.G[3 * w + 2] - B.lo[2];
    m[((size_t)a * B.dim[1] + b) * B.dim[2] + c] = kp.pos[w];
  }
  B.map = upload(m);
  return B;
}

std::vector<cdouble> fetch_projection_row(pawb200_pswf* wf, int kap, int band) {
  require_projections(wf);
  const int np = wf->proj_sites ? wf->proj_sites->nproj : 0;
  std::vector<cdouble> row(std::max(np, 1));
  CUDA_OK(cudaMemcpyAsync(row.data(), wf->P[kap].as<double2>() + (long)band * wf->ldp,
                          (size_t)np * sizeof(double2), cudaMemcpyDeviceToHost, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
  return row;
}

// Runs momentum_site_kernel for one set of per-element terms / per-site weights.
struct SiteTermSet {
  std::vector<MomElem> elems;
  std::vector<MomTerm> terms;
  std::vector<double> radial;
  std::vector<long> slot_off;
  std::vector<long> w_off;
  std::vector<cdouble> W;
};
void run_site_terms(const SiteTermSet& S, pawb200_pswf* wf, int numg, const DevBuf& dig, const double dk[3],
                    double gsign, const int* labels, const double* coords, double phase_sign, double pref,
                    double2* out) {
  const int ns = wf->num_sites;
  std::vector<int> se(ns);
  for (int s = 0; s < ns; s++) se[s] = labels[s];
  std::vector<double> recl(wf->reclattice, wf->reclattice + 9), crd(coords, coords + 3 * ns);
  DevBuf de = upload(S.elems), dt = upload(S.terms), dr = upload(S.radial), dso = upload(S.slot_off);
  DevBuf dse = upload(se), dc = upload(crd), dwo = upload(S.w_off), dW = upload(S.W), drecl = upload(recl);
  const size_t smem = std::max<size_t>(S.terms.size(), 1) * sizeof(double2);
  if (smem > 48 * 1024)
    CUDA_OK(cudaFuncSetAttribute(momentum_site_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  momentum_site_kernel<<<numg, 128, smem, g_stream>>>(
      numg, dig.as<int>(), dk[0], dk[1], dk[2], gsign, drecl.as<double>(), (int)S.elems.size(), de.as<MomElem>(),
      dt.as<MomTerm>(), dr.as<double>(), dso.as<long>(), ns, dse.as<int>(), dc.as<double>(), dwo.as<long>(),
      dW.as<double2>(), phase_sign, pref, out);
  count_launch();
  check_launch();
}

}  // namespace

extern "C" {

void pawb200_momentum_grid_size(pawb200_pswf_t* wf, double* nb1max, double* nb2max, double* nb3max, int* npmax,
                                double encut) {
  API_BEGIN
  if (!wf) throw std::runtime_error("NULL wavefunction pointer");
  WavecarHeader hd;
  hd.encut = encut;
  memcpy(hd.lattice, wf->lattice, sizeof(hd.lattice));
  wavecar_bounds(hd);                                   // reader.c:55-127 `setup`
  *nb1max = hd.nbmax[0]; *nb2max = hd.nbmax[1]; *nb3max = hd.nbmax[2];
  *npmax = hd.npmax;
  API_END_VOID
}

int pawb200_get_momentum_grid(int* igall, pawb200_pswf_t* wf, double nb1max, double nb2max, double nb3max,
                              double encut) {
  API_BEGIN
  if (!wf) throw std::runtime_error("NULL wavefunction pointer");
  WavecarHeader hd;
  hd.encut = encut;
  memcpy(hd.lattice, wf->lattice, sizeof(hd.lattice));
  memcpy(hd.reclattice, wf->reclattice, sizeof(hd.reclattice));
  hd.nbmax[0] = nb1max; hd.nbmax[1] = nb2max; hd.nbmax[2] = nb3max;
  const double k0[3] = {0, 0, 0};
  std::vector<int32_t> g = enumerate_g(hd, k0, nullptr);  // momentum.c:402-445: the reader's loop at k = 0
  memcpy(igall, g.data(), g.size() * sizeof(int32_t));
  return (int)(g.size() / 3);
  API_END(-1)
}

void pawb200_grid_bounds(int* G_bounds, int* gdim, const int* igall, int num_waves) {   // momentum.c:365-388
  for (int w = 0; w < num_waves; w++)
    for (int d = 0; d < 3; d++) {
      const int g = igall[3 * w + d];
      if (g < G_bounds[2 * d]) G_bounds[2 * d] = g;
      else if (g > G_bounds[2 * d + 1]) G_bounds[2 * d + 1] = g;
    }
  for (int d = 0; d < 3; d++) gdim[d] = G_bounds[2 * d + 1] - G_bounds[2 * d] + 1;
}

void pawb200_list_to_grid_map(int* grid, const int* G_bounds, const int* gdim, const int* igall,
                              int num_waves) {                                          // momentum.c:390-400
  (void)G_bounds;
  for (int w = 0; w < num_waves; w++) {
    int G[3];
    for (int d = 0; d < 3; d++) G[d] = (igall[3 * w + d] % gdim[d] + gdim[d]) % gdim[d];
    grid[((long)G[0] * gdim[1] + G[1]) * gdim[2] + G[2]] = w;
  }
}

pawb200_density_ft_t* pawb200_get_all_transforms(pawb200_pswf_t* wf, double encut) {
  API_BEGIN
  (void)encut;                                           // the reference ignores it too (fixed 1e7, momentum.c:129)
  if (!wf || !wf->has_projections) throw std::runtime_error("setup_projections has not been run");
  auto D = std::make_unique<pawb200_density_ft>();
  for (const Element& pp : wf->pps->list.el) {
    pawb200_density_ft::Elem E;
    E.N = pp.wave_gridsize;
    int lmax = 0;
    for (auto& f : pp.funcs) lmax = std::max(lmax, f.l);
    BesselTransform bt(1e7, 0.0, 2 * lmax, E.N, pp.wave_grid.data());
    E.ks.assign(bt.kgrid().begin(), bt.kgrid().begin() + E.N);
    std::vector<double> rho(E.N);
    for (int n1 = 0; n1 < pp.num_projs; n1++)
      for (int n2 = 0; n2 < pp.num_projs; n2++) {
        const RadialFunc &f1 = pp.funcs[n1], &f2 = pp.funcs[n2];
        for (int i = 0; i < E.N; i++)                    // make_rho, momentum.c:109-114
          rho[i] = (f1.aewave[i] * f2.aewave[i] - f1.pswave[i] * f2.pswave[i]) / pp.wave_grid[i];
        for (int L = std::abs(f1.l - f2.l); L <= f1.l + f2.l; L += 2) {
          std::vector<double> tr = bt.forward(rho.data(), L);
          tr.resize(E.N);
          E.slot[{n1, n2, L}] = (int)E.tab.size();
          E.tab.push_back(pack_spline(tr, make_spline_v(E.ks, tr)));
        }
      }
    D->el.push_back(std::move(E));
  }
  return D.release();
  API_END(nullptr)
}

void pawb200_free_density_ft_elem_list(pawb200_density_ft_t* elems, int num_elems) {
  (void)num_elems;
  delete elems;
}

// matrix[g] = < b1,k1,s1 | exp(i (G_g + k1 - k2).r) | b2,k2,s2 >  (momentum.c:278-357).  The plane-wave part is
// accumulated in FP64 (reference: float complex); the one-centre part is FP64 in both.
void pawb200_get_momentum_matrix(pawb200_c128* matrix, int numg, const int* igall, pawb200_pswf_t* wf,
                                 const int* labels, const double* coords, int band1, int kpt1, int spin1, int band2,
                                 int kpt2, int spin2, pawb200_density_ft_t* T, double encut) {
  API_BEGIN
  (void)encut;
  require_device();
  if (!wf || !T) throw std::runtime_error("NULL pointer");
  if (wf->ncl) throw std::runtime_error("momentum matrix elements are implemented for collinear wavefunctions");
  const int kap1 = kpt1 + spin1 * wf->nwk, kap2 = kpt2 + spin2 * wf->nwk;
  check_kpoint(wf, band1, kap1);
  check_kpoint(wf, band2, kap2);
  if (!wf->has_projections) throw std::runtime_error("setup_projections has not been run");
  if (numg <= 0) return;
  std::vector<int> ig(igall, igall + 3 * (size_t)numg);
  DevBuf dig = upload(ig);
  DevBuf out((size_t)numg * sizeof(double2));
  // ---- plane-wave part ----
  {
    BoxMap B = build_box_map(wf, kap1);
    const KPointInfo& k2 = wf->kp[kap2];
    std::vector<int> g2(3 * (size_t)k2.nplane);
    for (int j = 0; j < k2.nplane; j++)
      for (int d = 0; d < 3; d++) g2[3 * j + d] = k2.G[3 * k2.perm[j] + d];       // storage order
    DevBuf dg2 = upload(g2);
    wait_coeffs(wf, kap1, band1, band1 + 1);
    wait_coeffs(wf, kap2, band2, band2 + 1);
    momentum_pseudo_kernel<<<(numg * 32 + 255) / 256, 256, 0, g_stream>>>(
        numg, dig.as<int>(), wf->C[kap1].as<float2>() + (long)band1 * wf->ldc[kap1],
        wf->C[kap2].as<float2>() + (long)band2 * wf->ldc[kap2], dg2.as<int>(), k2.nplane, B.map.as<int>(), B.lo[0],
        B.lo[1], B.lo[2], B.dim[0], B.dim[1], B.dim[2], out.as<double2>());
    count_launch();
    check_launch();
  }
  // ---- one-centre part: per site W[t] = sum over channel pairs feeding term t (momentum.c:318-350, 156-223) ----
  const std::vector<cdouble> P1 = fetch_projection_row(wf, kap1, band1), P2 = fetch_projection_row(wf, kap2, band2);
  SiteTermSet S;
  const auto& els = wf->pps->list.el;
  std::vector<std::map<std::tuple<int, int, int, int>, int>> term_index(els.size());
  for (size_t e = 0; e < els.size(); e++) {
    const auto& E = T->el[e];
    MomElem me;
    me.term_off = (int)S.terms.size();
    me.N = E.N;
    me.ks_off = (long)S.radial.size();
    S.radial.insert(S.radial.end(), E.ks.begin(), E.ks.end());
    std::vector<long> local_off(E.tab.size());
    for (size_t q = 0; q < E.tab.size(); q++) {
      local_off[q] = (long)S.radial.size();
      S.radial.insert(S.radial.end(), E.tab[q].begin(), E.tab[q].end());
    }
    for (auto& kv : E.slot) {
      const int L = std::get<2>(kv.first);
      for (int M = -L; M <= L; M++) {
        term_index[e][{std::get<0>(kv.first), std::get<1>(kv.first), L, M}] = (int)S.terms.size() - me.term_off;
        S.slot_off.push_back(local_off[kv.second]);
        S.terms.push_back(MomTerm{(int)S.slot_off.size() - 1, L, M, (L == 0 && M == 0) ? 1 : 0});
      }
    }
    me.nterm = (int)S.terms.size() - me.term_off;
    S.elems.push_back(me);
  }
  const cdouble I(0, 1);
  for (int s = 0; s < wf->num_sites; s++) {
    const int e = labels[s];
    if (e < 0 || e >= (int)els.size()) throw std::runtime_error("element label out of range");
    const Element& pp = els[e];
    const SiteDev& sd = wf->proj_sites->host[s];
    S.w_off.push_back((long)S.W.size());
    const size_t base = S.W.size();
    S.W.resize(base + S.elems[e].nterm, cdouble(0, 0));
    for (int i = 0; i < pp.total_projs; i++)
      for (int j = 0; j < pp.total_projs; j++) {
        const Channel &ci = pp.chan[i], &cj = pp.chan[j];
        const cdouble pq = std::conj(P1[sd.lm_off + i]) * P2[sd.lm_off + j];
        int lx, ly, mx, my;
        if (ci.l < cj.l) { lx = cj.l; ly = ci.l; mx = cj.m; my = ci.m; }
        else { lx = ci.l; ly = cj.l; mx = ci.m; my = cj.m; }
        if (my < 0) { mx = -mx; my = -my; }
        const int M = cj.m - ci.m;
        for (int L = std::abs(ci.l - cj.l); L <= ci.l + cj.l; L += 2) {
          if (std::abs(M) > L) continue;                                   // Y_L^M vanishes identically
          const cdouble fac = sbt_factor(lx, ly, L, mx, my) * 4 * kPi * std::pow(I, L) * std::pow(-1.0, cj.m);
          S.W[base + term_index[e][{ci.n, cj.n, L, M}]] += pq * fac;
        }
      }
  }
  const double dk[3] = {wf->kp[kap1].k[0] - wf->kp[kap2].k[0], wf->kp[kap1].k[1] - wf->kp[kap2].k[1],
                        wf->kp[kap1].k[2] - wf->kp[kap2].k[2]};
  run_site_terms(S, wf, numg, dig, dk, 1.0, labels, coords, 1.0, 1.0, out.as<double2>());
  CUDA_OK(cudaMemcpyAsync(matrix, out.p, (size_t)numg * sizeof(double2), cudaMemcpyDeviceToHost, g_stream));
  stream_sync();
  API_END_VOID
}

// Cs[g] += C(b,k,s,G_g) with | b,k,s > = V^-1/2 sum_G C exp(i (k+G).r): pseudo coefficient + partial-wave part
// (momentum.c:465-543).  kpt_num is the kappa index k + s*nwk like the reference's call (pawpyc.pyx:790-791).
void pawb200_fullwf_reciprocal(pawb200_c128* Cs, const int* igall, pawb200_pswf_t* wf, int numg, int band_num,
                               int kpt_num, const int* labels, const double* coords) {
  API_BEGIN
  require_device();
  if (wf && wf->ncl) throw std::runtime_error("fullwf_reciprocal is implemented for collinear wavefunctions");
  check_kpoint(wf, band_num, kpt_num);
  if (!wf->has_projections) throw std::runtime_error("setup_projections has not been run");
  if (numg <= 0) return;
  std::vector<int> ig(igall, igall + 3 * (size_t)numg);
  DevBuf dig = upload(ig);
  DevBuf out((size_t)numg * sizeof(double2));
  {
    void* st = g_arena.take((size_t)numg * sizeof(double2));
    memcpy(st, Cs, (size_t)numg * sizeof(double2));                         // the reference accumulates into Cs
    fetch_pinned(out.p, st, (size_t)numg * sizeof(double2), g_stream);
  }
  BoxMap B = build_box_map(wf, kpt_num);
  wait_coeffs(wf, kpt_num, band_num, band_num + 1);
  momentum_pick_kernel<<<(numg + 255) / 256, 256, 0, g_stream>>>(
      numg, dig.as<int>(), wf->C[kpt_num].as<float2>() + (long)band_num * wf->ldc[kpt_num], B.map.as<int>(), B.lo[0],
      B.lo[1], B.lo[2], B.dim[0], B.dim[1], B.dim[2], out.as<double2>());
  count_launch();
  check_launch();
  const std::vector<cdouble> P = fetch_projection_row(wf, kpt_num, band_num);
  SiteTermSet S;
  const auto& els = wf->pps->list.el;
  const cdouble I(0, 1);
  for (size_t e = 0; e < els.size(); e++) {
    const Element& pp = els[e];
    MomElem me;
    me.term_off = (int)S.terms.size();
    me.N = pp.wave_gridsize;
    me.ks_off = (long)S.radial.size();
    S.radial.insert(S.radial.end(), pp.kwave_grid.begin(), pp.kwave_grid.begin() + me.N);
    std::vector<long> foff(pp.num_projs);
    for (int n = 0; n < pp.num_projs; n++) {
      foff[n] = (long)S.radial.size();
      std::vector<double> kw(pp.funcs[n].kwave.begin(), pp.funcs[n].kwave.begin() + me.N);
      const std::vector<double> t = pack_spline(kw, pp.funcs[n].kwave_s);
      S.radial.insert(S.radial.end(), t.begin(), t.end());
    }
    for (int p = 0; p < pp.total_projs; p++) {
      S.slot_off.push_back(foff[pp.chan[p].n]);
      S.terms.push_back(MomTerm{(int)S.slot_off.size() - 1, pp.chan[p].l, pp.chan[p].m, 1});
    }
    me.nterm = pp.total_projs;
    S.elems.push_back(me);
  }
  for (int s = 0; s < wf->num_sites; s++) {
    const Element& pp = els[labels[s]];
    const SiteDev& sd = wf->proj_sites->host[s];
    S.w_off.push_back((long)S.W.size());
    for (int p = 0; p < pp.total_projs; p++) S.W.push_back(P[sd.lm_off + p] * std::pow(I, pp.chan[p].l));
  }
  const double* k = wf->kp[kpt_num].k;
  const double dk[3] = {k[0], k[1], k[2]};
  const double pref = 4 * kPi * std::pow(determinant3(wf->lattice), -0.5);
  run_site_terms(S, wf, numg, dig, dk, -1.0, labels, coords, -1.0, pref, out.as<double2>());
  CUDA_OK(cudaMemcpyAsync(Cs, out.p, (size_t)numg * sizeof(double2), cudaMemcpyDeviceToHost, g_stream));
  stream_sync();
  API_END_VOID
}

// < C1 | shift by dG | C2 > over the momentum grid (momentum.c:547-574); O(numg) host loop like the reference
void pawb200_quick_overlap(const int* dG, const pawb200_c128* C1s, const pawb200_c128* C2s, int numg, const int* Gs,
                           const int* gmap, const int* G_bounds, const int* gdim, double* re_im) {
  const cdouble* c1 = (const cdouble*)C1s;
  const cdouble* c2 = (const cdouble*)C2s;
  cdouble total(0, 0);
  for (int w = 0; w < numg; w++) {
    int GP[3] = {Gs[3 * w] + dG[0], Gs[3 * w + 1] + dG[1], Gs[3 * w + 2] + dG[2]};
    if (GP[0] >= G_bounds[0] && GP[0] <= G_bounds[1] && GP[1] >= G_bounds[2] && GP[1] <= G_bounds[3] &&
        GP[2] >= G_bounds[4] && GP[2] <= G_bounds[5]) {
      for (int d = 0; d < 3; d++) GP[d] = (GP[d] % gdim[d] + gdim[d]) % gdim[d];
      const int wp = gmap[((long)GP[0] * gdim[1] + GP[1]) * gdim[2] + GP[2]];
      if (wp >= 0) total += std::conj(c1[wp]) * c2[w];
    }
  }
  re_im[0] = total.real();
  re_im[1] = total.imag();
}

}  // extern "C"
