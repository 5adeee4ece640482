// Host-side interface of the pruned, band-interleaved inverse 3-D FFT (kernels in fft3d.cuh, launchers in
// fft_launch.cu - a translation unit of its own so that the engine and the transform compile in parallel).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace pawb200 {

constexpr int FFT_B = 16;          // interleaved bands per group (256 B per grid point)
constexpr int FFT_MAXR = 20;       // largest radix evaluated in registers
constexpr int FFT_ZC = 9;          // z-lines per work unit of the stand-alone pass Y

struct FftGeom {            // device-side description of one (k-point, grid) pruned transform
  int n1, n2, n3;           // grid
  int r1[3], r2[3];         // n_d = r1[d] * r2[d] (index 0: x, 1: y, 2: z)
  int r3[3];                // third factor (1 for two-factor axes): n_d = r1 * r2 * r3, all radices <= 10 then
  int three;                // some axis has three factors (inverse passes only; forward transforms take the generic path)
  int ncol, nplane;         // active (g1,g2) columns / active g1 planes
  const int* col_start;     // [ncol] first sorted plane-wave index of the column
  const int* col_cnt;       // [ncol]
  const int* zpos;          // [npw]  wrapped g3 of each sorted plane wave
  const int4* col_run;      // [ncol] {start, cnt, zlo, nfirst}: the column's plane waves occupy the cyclic z-run
                            //        zlo .. zlo+cnt-1 (mod n3); the first nfirst of the run sit at the END of the
                            //        sorted list (they wrap), see fft_pass_z_kernel.  Null if some column is not a run.
  const int4* plane_run;    // [nplane] {col0, cnt, ylo, nfirst}: same encoding for the active columns of an x-plane
                            //        along y (fused pass Y+X).  Null if some plane is not one cyclic run.
  const int* ysrc;          // [nplane][n2] column index holding (plane, y) or -1
  const int* xsrc;          // [n1] plane index holding x or -1
  const double2* tw[3];     // exp(+2 pi i m / n_d), m < n_d
  int pf;                   // passes prefetch the inputs of their next work item into L2 (PAWB200_FFT_PF, default on)
};

struct FftInput {           // coefficient source of pass Z
  const float2* Cil;        // 16-slot interleaved copy [ceil(nslot/16)][ldil][16] (null: staged pass Z reads C)
  long ldil;
  const float2* C;          // [nband][ldc] box-ordered rows
  long ldc;
  int halves, half_len;     // spinor halves per band, plane waves per half
};

// Fused pass Y+X: the y-transformed planes of a few z values (a "chunk") live in a small ring that stays in
// L2, so the intermediate T2 never travels to HBM (see fft_pass_yx_kernel).
struct YxConfig {
  bool ok = false;
  int mode = 1;             // 1: in-order ticket queue with per-line dependencies; 2: static phases (cooperative launch)
  int zch = 0;              // z values per chunk
  int nzc = 0;              // chunks per band group
  int ring = 0;             // chunk slots in the ring
  int lead = 0;             // chunk steps between the Y items of a chunk and its X items
  int grid = 0;             // persistent CTAs
  size_t ring_bytes = 0;    // ring * nplane * n2 * zch * 256 B
  size_t flag_words(int ngroups) const { return 8 + 2 * (size_t)ngroups * nzc; }
};

struct FftWork {            // scratch owned by the caller
  double2* T1 = nullptr;    // [ng][ncol][n3][16]
  double2* T2 = nullptr;    // stand-alone passes: [ng][nplane][n2][n3][16]; fused: the ring
  unsigned* flags = nullptr;   // fused path: YxConfig::flag_words(ng) words (ticket + per-chunk counters)
};

void init_small_twiddles();
// Work decomposition of the fused pass for this geometry; ok == false when the planes are not y-runs, the ring
// would not fit `l2_budget_bytes`, or PAWB200_FFT_FUSED=0.
YxConfig plan_fused_yx(const FftGeom& g, int num_sms, size_t l2_budget_bytes);
// Transforms `ng` groups (slots s0 .. s0+ns) into X[ng][n1][n2][n3][16]; returns the number of kernels launched.
// max_plane_cols: largest number of active columns in one x-plane (stage size of the TMA-fed pass Y).
int launch_pruned_passes(const FftGeom& g, const FftInput& in, int s0, int ns, int ng, double scale,
                         const FftWork& w, const YxConfig* yx, double2* X, int num_sms, cudaStream_t st,
                         int max_plane_cols);

// Forward transform + gather of `ng` interleaved boxes X[ng][n1][n2][n3][16] (fwd_fft3d, linalg.c:47-79): the result
// is written as complex64 coefficients in the interleaved layout out_il[ng][ldil][16] (box order), multiplied by
// `scale`.  Uses w.T1 / w.T2 as full-size scratch ([ng][ncol][n3][16], [ng][nplane][n2][n3][16]).
int launch_pruned_forward(const FftGeom& g, int ng, const double2* X, const FftWork& w, float2* out_il, long ldil,
                          double scale, int num_sms, cudaStream_t st);

}  // namespace pawb200
