// Launchers of the pruned band-interleaved inverse 3-D FFT (kernels: fft3d.cuh).  One launcher per pass,
// templated on the largest radix of ITS axis (10 / 12 / 14 / 16 / 20): thread count, register budget and resident
// CTAs follow the axis, not the worst axis of the grid.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>

#include "fft3d.cuh"

namespace pawb200 {

namespace {

#define FFT_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e_) + " at " +    \
                               __FILE__ + ":" + std::to_string(__LINE__));                  \
  } while (0)

constexpr int kFftSmemOptIn = 227 * 1024;

inline unsigned fft_grid_dim(long lines, int occ, int num_sms) {
  return (unsigned)std::min<long>(lines, (long)num_sms * std::max(occ, 1));
}

// Resident CTAs per SM of a kernel at a given block size / dynamic shared memory, cached per (kernel, threads,
// smem): grids differ between wavefunctions of one process (config 2 and config 3 in one bench run).
struct OccKey {
  const void* fn; int threads; size_t smem;
  bool operator<(const OccKey& o) const {
    return fn != o.fn ? fn < o.fn : threads != o.threads ? threads < o.threads : smem < o.smem;
  }
};
std::map<OccKey, int> g_occ;
std::map<const void*, bool> g_attr_set;

template <class K>
int cached_occupancy(K kernel, int threads, size_t smem) {
  const void* fn = (const void*)kernel;
  if (!g_attr_set[fn]) {
    FFT_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFftSmemOptIn));
    g_attr_set[fn] = true;
  }
  const OccKey key{fn, threads, smem};
  auto it = g_occ.find(key);
  if (it != g_occ.end()) return it->second;
  int occ = 0;
  FFT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
  occ = std::max(occ, 1);
  g_occ[key] = occ;
  return occ;
}

template <int RMAX>
void launch_pass_z(const FftGeom& g, const FftInput& in, int s0, int ns, int ng, double scale, double2* T1,
                   int num_sms, cudaStream_t st) {
  const int threads = std::max(g.r1[2], g.r2[2]) * FFT_B;
  const size_t smem = (size_t)(2 * g.n3 * FFT_B + g.n3) * sizeof(double2);
  const bool runs = g.col_run != nullptr && in.Cil != nullptr;
  if (runs) {
    const int occ = cached_occupancy(fft_pass_z_kernel<RMAX>, threads, smem);
    fft_pass_z_kernel<RMAX><<<fft_grid_dim((long)ng * g.ncol, occ, num_sms), threads, smem, st>>>(
        g, in.Cil, in.ldil, s0, ns, scale, T1, ng);
  } else {
    const int occ = cached_occupancy(fft_pass_z_staged_kernel<RMAX>, threads, smem);
    fft_pass_z_staged_kernel<RMAX><<<fft_grid_dim((long)ng * g.ncol, occ, num_sms), threads, smem, st>>>(
        g, in.C, in.ldc, in.halves, in.half_len, s0, ns, scale, T1, ng);
  }
}

template <int RMAX>
void launch_pass_y(const FftGeom& g, int ng, const double2* T1, double2* T2, int num_sms, cudaStream_t st) {
  const int threads = std::max(g.r1[1], g.r2[1]) * FFT_B;
  const size_t smem = (size_t)(2 * g.n2 * FFT_B + g.n2) * sizeof(double2) + g.n2 * sizeof(int);
  const int occ = cached_occupancy(fft_pass_y_kernel<RMAX>, threads, smem);
  fft_pass_y_kernel<RMAX><<<fft_grid_dim((long)ng * g.nplane * ((g.n3 + FFT_ZC - 1) / FFT_ZC), occ, num_sms), threads,
                            smem, st>>>(g, T1, T2, ng);
}

template <int RMAX>
void launch_pass_x(const FftGeom& g, int ng, const double2* T2, double2* X, int num_sms, cudaStream_t st) {
  const int threads = std::max(g.r1[0], g.r2[0]) * FFT_B;
  const size_t smem = (size_t)(2 * g.n1 * FFT_B + g.n1) * sizeof(double2) + g.n1 * sizeof(int);
  const int occ = cached_occupancy(fft_pass_x_kernel<RMAX>, threads, smem);
  fft_pass_x_kernel<RMAX><<<fft_grid_dim((long)ng * g.n2 * g.n3, occ, num_sms), threads, smem, st>>>(g, T2, X, ng);
}

// three-factor axis (fft3_pass_kernel<PASS>): Q = 16 or 32 task slots of 16 bands
template <int PASS>
void launch_pass3(const FftGeom& g, const FftInput& in, int s0, int ns, int ng, double scale, const double2* src,
                  double2* dst, int num_sms, cudaStream_t st) {
  const int n = PASS == 0 ? g.n1 : PASS == 1 ? g.n2 : g.n3;
  const int threads = (n <= 220 ? 16 : 32) * FFT_B;
  const size_t smem = (size_t)(2 * n * FFT_B + n) * sizeof(double2) + (size_t)std::max(g.n1, g.n2) * sizeof(int);
  const int occ = cached_occupancy(fft3_pass_kernel<PASS>, threads, smem);
  const long lines = PASS == 0 ? (long)ng * g.n2 * g.n3 : PASS == 1 ? (long)ng * g.nplane * g.n3 : (long)ng * g.ncol;
  fft3_pass_kernel<PASS><<<fft_grid_dim(lines, occ, num_sms), threads, smem, st>>>(g, in.Cil, in.ldil, s0, ns, scale,
                                                                                   src, dst, ng);
}

inline int yx_threads(const FftGeom& g) {
  return std::max(std::max(g.r1[0], g.r2[0]), std::max(g.r1[1], g.r2[1])) * FFT_B;
}
inline size_t yx_smem(const FftGeom& g) {
  const int nmax = std::max(g.n1, g.n2);
  return (size_t)(2 * nmax * FFT_B + g.n1 + g.n2) * sizeof(double2) + (size_t)g.nplane * sizeof(int4) +
         (size_t)g.n1 * sizeof(int) + 2 * sizeof(YxItem) + 16;
}

template <int RMAX>
int yx_occupancy(const FftGeom& g) {
  return cached_occupancy(fft_pass_yx_kernel<RMAX>, yx_threads(g), yx_smem(g));
}

template <int RMAX>
void launch_pass_yx(const FftGeom& g, const YxConfig& yx, int ng, const double2* T1, double2* ring, unsigned* flags,
                    double2* X, cudaStream_t st) {
  YxArgs a;
  a.zch = yx.zch; a.nzc = yx.nzc; a.ring = yx.ring; a.lead = yx.lead;
  a.nchunks = ng * yx.nzc;
  a.ticket = flags;
  a.ydone = flags + 8;
  a.xdone = flags + 8 + a.nchunks;
  FFT_CUDA_OK(cudaMemsetAsync(flags, 0, yx.flag_words(ng) * sizeof(unsigned), st));
  fft_pass_yx_kernel<RMAX><<<yx.grid, yx_threads(g), yx_smem(g), st>>>(g, a, T1, ring, X);
}

inline size_t yx2_smem(const FftGeom& g) {
  const int nmax = std::max(g.n1, g.n2);
  return (size_t)(2 * nmax * FFT_B + g.n1 + g.n2) * sizeof(double2) + (size_t)g.nplane * sizeof(int4) +
         (size_t)g.n1 * sizeof(int) + 16;
}

template <int RMAX>
int yx2_occupancy(const FftGeom& g) {
  return cached_occupancy(fft_pass_yx2_kernel<RMAX>, yx_threads(g), yx2_smem(g));
}

// phased fused pass: cooperative launch (every CTA must be resident, see fft_pass_yx2_kernel)
template <int RMAX>
void launch_pass_yx2(const FftGeom& g, const YxConfig& yx, int ng, const double2* T1, double2* ring, unsigned* flags,
                     double2* X, cudaStream_t st) {
  FftGeom gg = g;
  Yx2Args a;
  a.zch = yx.zch; a.nzc = yx.nzc; a.ring = yx.ring;
  a.nchunks = ng * yx.nzc;
  a.ydone = flags;
  a.xdone = flags + a.nchunks;
  FFT_CUDA_OK(cudaMemsetAsync(flags, 0, yx.flag_words(ng) * sizeof(unsigned), st));
  const long items = (long)(a.nchunks + 1) * (long)(g.nplane + g.n2) * yx.zch;
  const unsigned grid = (unsigned)std::min<long>(yx.grid, items);
  void* args[] = {&gg, &a, &T1, &ring, &X};
  FFT_CUDA_OK(cudaLaunchCooperativeKernel((const void*)fft_pass_yx2_kernel<RMAX>, dim3(grid), dim3(yx_threads(g)), args,
                                          yx2_smem(g), st));
}

// ---- TMA-fed passes (fft_pass_tma_kernel) ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// rows x inner rank-3 view of an interleaved scratch array: [outer][mid][16 bands] of complex128, i.e.
// dims (fastest first) {32 doubles, mid, outer}, box {32, zl, box_rows}
bool make_tensor_map(CUtensorMap* m, const void* base, long mid, long outer, int zl, int box_rows) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {32, (cuuint64_t)mid, (cuuint64_t)outer};
  const cuuint64_t strides[2] = {256, (cuuint64_t)mid * 256};
  const cuuint32_t box[3] = {32, (cuuint32_t)zl, (cuuint32_t)box_rows};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

struct TmaPlan {
  bool ok = false;
  TmaPassArgs a;
  int box_rows = 0, threads = 0;
  size_t smem = 0;
};

// Stage / exchange-buffer configuration of a TMA-fed pass (pass 0 = X, 1 = Y): the deepest ring that still leaves
// two CTAs per SM; PAWB200_TMA_ZL / _STAGES / _DBL override (tuning).
TmaPlan plan_tma_pass(const FftGeom& g, int pass, int max_plane_cols) {
  TmaPlan p;
  // Opt-in (PAWB200_FFT_TMA=1).  Measured on config 2 (fft ms/step): register-direct passes 10.4; TMA-fed, two
  // CTAs/SM with a 3-stage ring 12.9, ZL=2 12.0, single exchange buffer (three CTAs/SM) 10.9 - the passes are bound by
  // resident warps (issue / FP64 latency between the two phases of a line), not by bytes in flight, so shared memory
  // spent on stages costs more than the prefetch returns.
  if (env_int("PAWB200_FFT_TMA", 0) == 0 || !tensor_map_encoder() || g.three) return p;
  if (pass == 1 && !g.plane_run) return p;
  const int n = pass == 0 ? g.n1 : g.n2;
  const int zl = std::max(1, std::min(8, env_int("PAWB200_TMA_ZL", 1)));
  int rows;
  if (pass == 0) {
    p.box_rows = std::min(g.nplane, 256);
    rows = (g.nplane + p.box_rows - 1) / p.box_rows * p.box_rows;
  } else {
    p.box_rows = FFT_TMA_COLBOX;
    rows = (max_plane_cols + FFT_TMA_COLBOX - 1) / FFT_TMA_COLBOX * FFT_TMA_COLBOX;
  }
  const size_t stage = (size_t)rows * zl * FFT_B * sizeof(double2);
  auto smem_of = [&](int S, int dbl) {
    return (size_t)S * stage + (size_t)(dbl ? 2 : 1) * n * FFT_B * sizeof(double2) + (size_t)n * sizeof(double2) +
           (pass == 1 ? (size_t)g.nplane * sizeof(int4) : (size_t)g.n1 * sizeof(int)) + (size_t)S * 8 + 32;
  };
  const int prefs[][2] = {{3, 1}, {2, 1}, {3, 0}, {2, 0}};
  int S = 2, dbl = 0;
  bool found = false;
  for (auto& pr : prefs)
    if (!found && smem_of(pr[0], pr[1]) <= (size_t)kFftSmemOptIn / 2) { S = pr[0]; dbl = pr[1]; found = true; }
  S = std::max(2, env_int("PAWB200_TMA_STAGES", S));
  dbl = env_int("PAWB200_TMA_DBL", dbl) ? 1 : 0;
  p.smem = smem_of(S, dbl);
  if (p.smem > (size_t)kFftSmemOptIn) return p;
  p.a.zl = zl; p.a.stages = S; p.a.stage_rows = rows; p.a.dbl = dbl;
  p.threads = std::max(g.r1[pass == 0 ? 0 : 1], g.r2[pass == 0 ? 0 : 1]) * FFT_B;
  p.ok = true;
  return p;
}

template <int RMAX, int PASS>
bool launch_pass_tma(const FftGeom& g, const TmaPlan& p, int ng, const double2* in, double2* out, int num_sms,
                     cudaStream_t st) {
  CUtensorMap tmap;
  const long plane = (long)g.n2 * g.n3;
  const bool ok = PASS == 0 ? make_tensor_map(&tmap, in, plane, (long)g.nplane * ng, p.a.zl, p.box_rows)
                            : make_tensor_map(&tmap, in, g.n3, (long)g.ncol * ng, p.a.zl, p.box_rows);
  if (!ok) return false;
  const int occ = cached_occupancy(fft_pass_tma_kernel<RMAX, PASS>, p.threads, p.smem);
  const long upg = PASS == 0 ? (plane + p.a.zl - 1) / p.a.zl : (long)g.nplane * ((g.n3 + p.a.zl - 1) / p.a.zl);
  fft_pass_tma_kernel<RMAX, PASS><<<fft_grid_dim((long)ng * upg, occ, num_sms), p.threads, p.smem, st>>>(
      tmap, g, p.a, out, ng);
  return true;
}

}  // namespace

#define PAWB200_AXIS_SWITCH(r, CALL) \
  do {                               \
    if ((r) <= 10) { CALL(10); }     \
    else if ((r) <= 12) { CALL(12); } \
    else if ((r) <= 14) { CALL(14); } \
    else if ((r) <= 16) { CALL(16); } \
    else { CALL(20); }               \
  } while (0)

void init_small_twiddles() {
  static bool done = false;
  if (done) return;
  double2 h[FFT_MAXR + 1][FFT_MAXR];
  memset(h, 0, sizeof(h));
  for (int R = 1; R <= FFT_MAXR; R++)
    for (int m = 0; m < R; m++) {
      const long double a = 2.0L * 3.141592653589793238462643383279502884L * m / R;
      h[R][m] = make_double2((double)cosl(a), (double)sinl(a));
    }
  FFT_CUDA_OK(cudaMemcpyToSymbol(c_small_tw, h, sizeof(h)));
  done = true;
}

YxConfig plan_fused_yx(const FftGeom& g, int num_sms, size_t l2_budget_bytes) {
  YxConfig c;
  // Opt-in (PAWB200_FFT_FUSED=1): measured on config 2 the fused pass moves 43 % fewer DRAM bytes (ncu: 2.27 GB
  // instead of 4.0 GB per 128 bands) but the per-line inter-CTA synchronisation makes it slower than the two
  // stand-alone passes (1.29 ms against 0.88 ms), which are latency- rather than bandwidth-bound.
  const char* e = getenv("PAWB200_FFT_FUSED");
  if (!e || atoi(e) == 0) return c;
  if (!g.plane_run || g.three) return c;
  if (yx_smem(g) > (size_t)kFftSmemOptIn) return c;
  const int ryx = std::max(std::max(g.r1[0], g.r2[0]), std::max(g.r1[1], g.r2[1]));
  int occ = 1;
  if (atoi(e) == 2) {
    // phased variant: ring of 3 chunk slots, chunk = zch z values of one band group
    int coop = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return c;
#define OCC2(R) occ = yx2_occupancy<R>(g)
    PAWB200_AXIS_SWITCH(ryx, OCC2);
#undef OCC2
    const size_t per_z = (size_t)g.nplane * g.n2 * FFT_B * sizeof(double2);
    int zmax = (int)std::min<size_t>((size_t)g.n3, l2_budget_bytes / (3 * per_z));
    if (zmax < 1) return c;
    int zch = zmax;
    for (int z = zmax; z >= std::max(1, (2 * zmax + 2) / 3); z--)      // prefer a divisor of n3 close to the budget
      if (g.n3 % z == 0) { zch = z; break; }
    c.mode = 2;
    c.zch = zch;
    c.nzc = (g.n3 + zch - 1) / zch;
    c.ring = 3;
    c.lead = 1;
    c.grid = num_sms * occ;
    c.ring_bytes = (size_t)c.ring * per_z * zch;
    c.ok = true;
    return c;
  }
#define OCC(R) occ = yx_occupancy<R>(g)
  PAWB200_AXIS_SWITCH(ryx, OCC);
#undef OCC
  const long grid = (long)num_sms * occ;
  // z values per chunk: the largest of 4..1 whose ring fits the L2 budget, preferring divisors of n3 (no padding)
  int best = 0;
  for (int pass = 0; pass < 2 && !best; pass++)
    for (int zch = 4; zch >= 1 && !best; zch--) {
      if (pass == 0 && g.n3 % zch) continue;
      const long per_step = (long)(g.n2 + g.nplane) * zch;
      const int lead = (int)((2 * grid + per_step - 1) / per_step) + 1;   // a CTA holds an item and the next ticket
      const size_t bytes = (size_t)(2 * lead) * g.nplane * g.n2 * zch * FFT_B * sizeof(double2);
      if (bytes <= l2_budget_bytes) best = zch;
    }
  if (!best) return c;
  const long per_step = (long)(g.n2 + g.nplane) * best;
  c.zch = best;
  c.nzc = (g.n3 + best - 1) / best;
  c.lead = (int)((2 * grid + per_step - 1) / per_step) + 1;
  c.ring = 2 * c.lead;
  c.grid = (int)grid;
  c.ring_bytes = (size_t)c.ring * g.nplane * g.n2 * best * FFT_B * sizeof(double2);
  c.ok = true;
  return c;
}

int launch_pruned_forward(const FftGeom& g, int ng, const double2* X, const FftWork& w, float2* out_il, long ldil,
                          double scale, int num_sms, cudaStream_t st) {
  if (!g.col_run || g.three) throw std::runtime_error("forward pruned transform needs single-run columns and two-factor axes");
  const int rz = std::max(g.r1[2], g.r2[2]), ry = std::max(g.r1[1], g.r2[1]), rx = std::max(g.r1[0], g.r2[0]);
#define FWD_X(R)                                                                                                    \
  {                                                                                                                 \
    const int threads = rx * FFT_B;                                                                                 \
    const size_t smem = (size_t)(2 * g.n1 * FFT_B + g.n1) * sizeof(double2) + g.n1 * sizeof(int);                  \
    const int occ = cached_occupancy(fft_fwd_pass_x_kernel<R>, threads, smem);                                      \
    fft_fwd_pass_x_kernel<R><<<fft_grid_dim((long)ng * g.n2 * g.n3, occ, num_sms), threads, smem, st>>>(g, X, w.T2, ng); \
  }
#define FWD_Y(R)                                                                                                    \
  {                                                                                                                 \
    const int threads = ry * FFT_B;                                                                                 \
    const size_t smem = (size_t)(2 * g.n2 * FFT_B + g.n2) * sizeof(double2) + g.n2 * sizeof(int);                  \
    const int occ = cached_occupancy(fft_fwd_pass_y_kernel<R>, threads, smem);                                      \
    fft_fwd_pass_y_kernel<R><<<fft_grid_dim((long)ng * g.nplane * ((g.n3 + FFT_ZC - 1) / FFT_ZC), occ, num_sms),    \
                               threads, smem, st>>>(g, w.T2, w.T1, ng);                                             \
  }
#define FWD_Z(R)                                                                                                    \
  {                                                                                                                 \
    const int threads = rz * FFT_B;                                                                                 \
    const size_t smem = (size_t)(2 * g.n3 * FFT_B + g.n3) * sizeof(double2);                                        \
    const int occ = cached_occupancy(fft_fwd_pass_z_kernel<R>, threads, smem);                                      \
    fft_fwd_pass_z_kernel<R><<<fft_grid_dim((long)ng * g.ncol, occ, num_sms), threads, smem, st>>>(                 \
        g, w.T1, out_il, ldil, scale, ng);                                                                          \
  }
  PAWB200_AXIS_SWITCH(rx, FWD_X);
  PAWB200_AXIS_SWITCH(ry, FWD_Y);
  PAWB200_AXIS_SWITCH(rz, FWD_Z);
#undef FWD_X
#undef FWD_Y
#undef FWD_Z
  FFT_CUDA_OK(cudaGetLastError());
  return 3;
}

int launch_pruned_passes(const FftGeom& g, const FftInput& in, int s0, int ns, int ng, double scale,
                         const FftWork& w, const YxConfig* yx, double2* X, int num_sms, cudaStream_t st,
                         int max_plane_cols) {
  const int rz = std::max(g.r1[2], g.r2[2]), ry = std::max(g.r1[1], g.r2[1]), rx = std::max(g.r1[0], g.r2[0]);
  if (g.three) {
    // some axis has three factors: per axis either the three-factor kernel or the two-factor one (no fused / TMA
    // variants, and pass Z needs the run encoding + interleaved coefficients, which the plan guarantees here)
#define PASS_Z(R) launch_pass_z<R>(g, in, s0, ns, ng, scale, w.T1, num_sms, st)
#define PASS_Y(R) launch_pass_y<R>(g, ng, w.T1, w.T2, num_sms, st)
#define PASS_X(R) launch_pass_x<R>(g, ng, w.T2, X, num_sms, st)
    if (g.r3[2] > 1 && (!in.Cil || !g.col_run))
      throw std::runtime_error("three-factor pass Z needs the interleaved coefficient copy and single-run columns");
    if (g.r3[2] > 1) launch_pass3<2>(g, in, s0, ns, ng, scale, nullptr, w.T1, num_sms, st);
    else PAWB200_AXIS_SWITCH(rz, PASS_Z);
    if (g.r3[1] > 1) launch_pass3<1>(g, in, s0, ns, ng, scale, w.T1, w.T2, num_sms, st);
    else PAWB200_AXIS_SWITCH(ry, PASS_Y);
    if (g.r3[0] > 1) launch_pass3<0>(g, in, s0, ns, ng, scale, w.T2, X, num_sms, st);
    else PAWB200_AXIS_SWITCH(rx, PASS_X);
#undef PASS_Z
#undef PASS_Y
#undef PASS_X
    FFT_CUDA_OK(cudaGetLastError());
    return 3;
  }
#define PASS_Z(R) launch_pass_z<R>(g, in, s0, ns, ng, scale, w.T1, num_sms, st)
  PAWB200_AXIS_SWITCH(rz, PASS_Z);
#undef PASS_Z
  if (yx && yx->ok) {
    const int ryx = std::max(ry, rx);
#define PASS_YX(R) launch_pass_yx<R>(g, *yx, ng, w.T1, w.T2, w.flags, X, st)
#define PASS_YX2(R) launch_pass_yx2<R>(g, *yx, ng, w.T1, w.T2, w.flags, X, st)
    if (yx->mode == 2) PAWB200_AXIS_SWITCH(ryx, PASS_YX2);
    else PAWB200_AXIS_SWITCH(ryx, PASS_YX);
#undef PASS_YX
#undef PASS_YX2
    FFT_CUDA_OK(cudaGetLastError());
    return 2;
  }
  const TmaPlan ty = plan_tma_pass(g, 1, max_plane_cols), tx = plan_tma_pass(g, 0, max_plane_cols);
  bool done_y = false, done_x = false;
#define PASS_Y_TMA(R) done_y = launch_pass_tma<R, 1>(g, ty, ng, w.T1, w.T2, num_sms, st)
#define PASS_X_TMA(R) done_x = launch_pass_tma<R, 0>(g, tx, ng, w.T2, X, num_sms, st)
#define PASS_Y(R) launch_pass_y<R>(g, ng, w.T1, w.T2, num_sms, st)
#define PASS_X(R) launch_pass_x<R>(g, ng, w.T2, X, num_sms, st)
  if (ty.ok) PAWB200_AXIS_SWITCH(ry, PASS_Y_TMA);
  if (!done_y) PAWB200_AXIS_SWITCH(ry, PASS_Y);
  if (tx.ok) PAWB200_AXIS_SWITCH(rx, PASS_X_TMA);
  if (!done_x) PAWB200_AXIS_SWITCH(rx, PASS_X);
#undef PASS_Y
#undef PASS_X
#undef PASS_Y_TMA
#undef PASS_X_TMA
  FFT_CUDA_OK(cudaGetLastError());
  return 3;
}

}  // namespace pawb200
