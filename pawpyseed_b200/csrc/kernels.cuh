// sm_100a kernels of the PAW band-projection path (everything except the complex GEMM,
// which lives in zgemm.cuh).  Layouts and roofline notes are in DESIGN.md; each kernel names
// the reference loop it replaces (file:line relative to pawpyseed/core/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pawb200 {

// ---------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
// D(8x8) += A(8x4,row) * B(4x8,col), FP64 tensor core (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// ---------------------------------------------------------------------------------------
// (a2) plane-wave scatter into the FFT box, fused with the zero fill  [linalg.c:22-33]
//   x[slot][g] = inv[g] >= 0 ? scale * widen(C[slot_base(slot) + inv[g]]) : 0
// One thread per grid point, all `nslot` boxes of the batch written from one read of inv[].
// Writes are fully coalesced 16-B stores; algorithmic bytes = 16*N_grid + 8*npw per slot.
//   slot -> coefficient row: band = slot / halves, half = slot % halves (noncollinear: 2)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_pw_kernel(const float2* __restrict__ C, long ldc, int band0, int halves, int half_len,
                  const int* __restrict__ inv, double2* __restrict__ x, long ngrid, int nslot,
                  double scale) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < ngrid; g += stride) {
    const int j = __ldg(inv + g);
    if (j < 0) {
#pragma unroll 4
      for (int s = 0; s < nslot; s++) x[(long)s * ngrid + g] = make_double2(0.0, 0.0);
    } else {
#pragma unroll 4
      for (int s = 0; s < nslot; s++) {
        const int band = band0 + s / halves, half = s % halves;
        const float2 c = __ldg(C + (long)band * ldc + (long)half * half_len + j);
        x[(long)s * ngrid + g] = make_double2(scale * (double)c.x, scale * (double)c.y);
      }
    }
  }
}

// One-time reorder of the coefficient columns into FFT-box index order (z runs contiguous), so the
// scatter reads them coalesced.  Both wavefunctions of a pair get the same permutation (same G list),
// so the band-band GEMM over the plane-wave axis is unaffected.
__global__ void __launch_bounds__(256)
permute_coeff_kernel(const float2* __restrict__ raw, float2* __restrict__ C, long ld, int nband,
                     int halves, int half_len, const int* __restrict__ perm) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= half_len) return;
  const int src = perm[j];
  for (int b = blockIdx.y; b < nband; b += gridDim.y)
    for (int h = 0; h < halves; h++)
      C[(long)b * ld + (long)h * half_len + j] = raw[(long)b * ld + (long)h * half_len + src];
}

// Second coefficient layout for the pruned FFT: slots (band, spinor half) interleaved in groups of 16,
//   Cil[((slot >> 4) * ldil + j) * 16 + (slot & 15)] = C[slot / halves][(slot % halves) * half_len + j]
// so that pass Z reads 128-byte rows (16 slots of one plane wave) straight into registers.
// Tiled transpose: CTA = one 16-slot group x 128 plane waves; only slots in [slot_lo, slot_hi) are written.
__global__ void __launch_bounds__(256)
interleave_coeff_kernel(const float2* __restrict__ C, long ldc, int halves, int half_len, int slot_lo,
                        int slot_hi, float2* __restrict__ Cil, long ldil) {
  __shared__ float2 tile[16][129];
  const int grp = (slot_lo >> 4) + blockIdx.y;
  const int j0 = blockIdx.x * 128;
  for (int e = threadIdx.x; e < 16 * 128; e += 256) {
    const int r = e >> 7, jj = e & 127;
    const int slot = grp * 16 + r, j = j0 + jj;
    float2 v = make_float2(0.f, 0.f);
    if (slot >= slot_lo && slot < slot_hi && j < half_len)
      v = C[(long)(slot / halves) * ldc + (long)(slot % halves) * half_len + j];
    tile[r][jj] = v;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 16 * 128; e += 256) {
    const int jj = e >> 4, r = e & 15;
    const int slot = grp * 16 + r, j = j0 + jj;
    if (slot >= slot_lo && slot < slot_hi && j < half_len) Cil[((long)grp * ldil + j) * 16 + r] = tile[r][jj];
  }
}

// inverse of interleave_coeff_kernel for whole groups: rows[band0 + grp*16 + r][j] = il[grp][j][r]
// (the forward pruned transform leaves its coefficients in the interleaved layout; the GEMM wants band rows)
__global__ void __launch_bounds__(256)
deinterleave_coeff_kernel(const float2* __restrict__ il, long ldil, int npw, int band0, int nband,
                          float2* __restrict__ rows, long ldc) {
  __shared__ float2 tile[16][129];
  const int grp = blockIdx.y;
  const int j0 = blockIdx.x * 128;
  for (int e = threadIdx.x; e < 16 * 128; e += 256) {
    const int jj = e >> 4, r = e & 15;
    const int j = j0 + jj;
    tile[r][jj] = j < npw ? il[((long)grp * ldil + j) * 16 + r] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 16 * 128; e += 256) {
    const int r = e >> 7, jj = e & 127;
    const int band = grp * 16 + r, j = j0 + jj;
    if (band < nband && j < npw) rows[(long)(band0 + band) * ldc + j] = tile[r][jj];
  }
}

// (f2) k-point desymmetrisation  [utils.c:1070-1081]: Cnew[b][j] = fac[j] * Cold[b][src[j]] (conjugated under
// time reversal), in single precision without FMA contraction so it rounds like the reference's complex float
// multiply.
__global__ void __launch_bounds__(256)
symm_map_kernel(const float2* __restrict__ Cold, long ldo, float2* __restrict__ Cnew, long ldn, int nband,
                int npw, const int* __restrict__ src, const float2* __restrict__ fac, int tr) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npw) return;
  const int sj = src[j];
  const float2 f = fac[j];
  for (int b = blockIdx.y; b < nband; b += gridDim.y) {
    const float2 c = Cold[(long)b * ldo + sj];
    float re = __fsub_rn(__fmul_rn(f.x, c.x), __fmul_rn(f.y, c.y));
    float im = __fadd_rn(__fmul_rn(f.x, c.y), __fmul_rn(f.y, c.x));
    if (tr) im = -im;
    Cnew[(long)b * ldn + j] = make_float2(re, im);
  }
}

// Small host->device uploads (tables, index lists, plans) read pinned host memory from a kernel instead of using
// the copy engine, so they never queue behind the bulk coefficient transfers the ingest stream has in flight.
__global__ void __launch_bounds__(256)
pinned_fetch_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, size_t n) {
  const size_t n16 = n / 16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
    reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  for (size_t i = n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

// inverse scatter map: inv[lin[w]] = val[w]  (lin holds distinct grid positions)
__global__ void fill_inverse_map_kernel(const int* __restrict__ lin, const int* __restrict__ val, int n,
                                        int* __restrict__ inv) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < n) inv[lin[w]] = val[w];
}

// (a3) gather back after a forward FFT  [linalg.c:72-77]; narrow to complex64
__global__ void gather_pw_kernel(const double2* __restrict__ x, const int* __restrict__ gidx,
                                 float2* __restrict__ Cout, int npw, double scale) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < npw) {
    const double2 v = x[gidx[w]];
    Cout[w] = make_float2((float)(v.x * scale), (float)(v.y * scale));
  }
}

// ---------------------------------------------------------------------------------------
// (a4) projector / partial-wave tables on sphere points  [utils.c:547-588, 672-686]
// One thread per (site point); evaluates the radial cubic spline once per radial channel and
// the complex Y_lm for each m.  Geometry (index list, membership) was decided on the host.
// ---------------------------------------------------------------------------------------
struct ElemDev {            // per element, device pointers
  const double* grid;       // linear radial grid (proj_grid or smooth_grid), n points
  const double* f;          // [nfunc][n] function values
  const double* spl;        // [nfunc][3][n] spline rows
  const int* chan_n;        // per channel: radial index
  const int* chan_l;
  const int* chan_m;
  int n, nfunc, nchan;
  double rmax;
};

__device__ __forceinline__ double dpow_half(double v) { return sqrt(v); }

__device__ double d_legendre(int l, int m, double x);
__device__ inline double d_ifac(int n) {
  double t = 1;
  for (int q = 2; q <= n; q++) t *= q;
  return t;
}
__device__ inline double d_legendre(int l, int m, double x) {
  // closed sum of utils.c:377-386 for m >= 0; negative m through the symmetry relation
  int am = m < 0 ? -m : m;
  double total = 0;
  for (int n = l; n >= 0 && 2 * n - l - am >= 0; n--) {
    double term = pow(x, (double)(2 * n - l - am)) * d_ifac(2 * n) / d_ifac(2 * n - l - am) /
                  d_ifac(n) / d_ifac(l - n);
    total += ((l - n) & 1) ? -term : term;
  }
  double v = total * ((am & 1) ? -1.0 : 1.0) * pow(1 - x * x, am / 2.0) / (double)(1 << l);
  if (m < 0) v *= ((am & 1) ? -1.0 : 1.0) * d_ifac(l - am) / d_ifac(l + am);
  return v;
}
__device__ inline double2 d_ylm(int l, int m, double theta, double phi) {
  const double PI = 3.14159265358979323846;
  const double norm = sqrt((2 * l + 1) / (4 * PI) * d_ifac(l - m) / d_ifac(l + m));
  const double p = norm * d_legendre(l, m, cos(theta));
  double s, c;
  sincos(m * phi, &s, &c);
  return make_double2(p * c, p * s);
}
__device__ inline void d_angles(const double* v, double r, double* theta, double* phi) {
  const double PI = 3.14159265358979323846;
  if (r == 0) {
    *theta = 0;
    *phi = 0;
    return;
  }
  *theta = acos(v[2] / r);
  if (r - fabs(v[2]) == 0)
    *phi = 0;
  else
    *phi = acos(v[0] / sqrt(v[0] * v[0] + v[1] * v[1]));
  if (v[1] < 0) *phi = 2 * PI - *phi;
}
__device__ inline double d_eval_linear(double r, double rmax, int n, const double* x,
                                       const double* f, const double* spl) {
  if (r > x[n - 1]) return 0;
  if (r < x[0]) return f[0];
  int i = (int)(r / rmax * n);
  if (i > n - 2) i = n - 2;
  const double t = r - x[i];
  return f[i] + t * (spl[i] + t * (spl[n + i] + t * spl[2 * n + i]));
}
__device__ inline double d_eval_log(double r, int n, const double* x, const double* f,
                                    const double* spl) {
  if (r > x[n - 1]) return 0;
  if (r < x[0]) return f[0];
  int i = (int)(log(r / x[0]) / log(x[1] / x[0]));
  if (i > n - 2) i = n - 2;
  const double t = r - x[i];
  return f[i] + t * (spl[i] + t * (spl[n + i] + t * spl[2 * n + i]));
}

struct SiteDev {            // per site in a table set
  int elem;                 // element label
  int npts, npts_pad;       // real / padded point count (pad = multiple of 32)
  long pt_off;              // offset into idx/path arrays (padded points)
  long tab_off;             // offset (double2 elements) of table block [nlm][npts_pad]
  int nlm, lm_off;          // channels of this site, offset into the concatenated channel axis
  double coord[3];          // fractional position of the atom
};

// Sphere geometry on the device from the unwrapped grid coordinates the host selected (bit-exact membership):
//   path = frac_to_cartesian((i/N0, j/N1, k/N2) - coord)  with the operation order of utils.c:656-660 and no FMA
//          contraction, so it equals the host / reference value bit for bit;
//   idx  = wrapped linear index;  wrap = (wrapped - unwrapped) / N  integer cell shifts (density.c:293-295).
// Uploading 8 bytes per point instead of 28-40 keeps the setup off the PCIe critical path.
struct Lattice9 { double a[9]; };
__global__ void __launch_bounds__(256)
expand_geometry_kernel(const SiteDev* __restrict__ sites, const short4* __restrict__ ijk, Lattice9 L, int n0, int n1,
                       int n2, int* __restrict__ idx, double* __restrict__ path, int* __restrict__ wrap, long ld) {
  const SiteDev sd = sites[blockIdx.y];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < sd.npts_pad; p += gridDim.x * blockDim.x) {
    const long q = sd.pt_off + p;
    if (p >= sd.npts) {
      idx[q] = 0;
      path[q] = path[ld + q] = path[2 * ld + q] = 0.0;
      if (wrap) wrap[q] = wrap[ld + q] = wrap[2 * ld + q] = 0;
      continue;
    }
    const short4 v = ijk[q];
    const int i = v.x, j = v.y, k = v.z;
    const double t0 = __dsub_rn(__ddiv_rn((double)i, (double)n0), sd.coord[0]);
    const double t1 = __dsub_rn(__ddiv_rn((double)j, (double)n1), sd.coord[1]);
    const double t2 = __dsub_rn(__ddiv_rn((double)k, (double)n2), sd.coord[2]);
    for (int d = 0; d < 3; d++)
      path[d * ld + q] = __dadd_rn(__dadd_rn(__dmul_rn(t0, L.a[d]), __dmul_rn(t1, L.a[3 + d])), __dmul_rn(t2, L.a[6 + d]));
    const int ii = (i % n0 + n0) % n0, jj = (j % n1 + n1) % n1, kk = (k % n2 + n2) % n2;
    idx[q] = (ii * n1 + jj) * n2 + kk;
    if (wrap) {
      wrap[q] = (ii - i) / n0;
      wrap[ld + q] = (jj - j) / n1;
      wrap[2 * ld + q] = (kk - k) / n2;
    }
  }
}

// mode 0: linear-grid radial function (projector or filtered phi-phit), value = R(r) Y_lm,
//         r from the minimum-image path of the wrapped grid point (utils.c:566-588)
// mode 1: log-grid partial wave difference, value = R(r)/r Y_lm from the direct offset
//         (wave_value2, utils.c:522-545)
__global__ void __launch_bounds__(128)
site_table_kernel(const SiteDev* __restrict__ sites, const ElemDev* __restrict__ elems,
                  const int* __restrict__ idx, const double* __restrict__ path, long path_ld,
                  double2* __restrict__ table, const double* __restrict__ lattice, int n0, int n1,
                  int n2, int mode) {
  const SiteDev sd = sites[blockIdx.y];
  const ElemDev ed = elems[sd.elem];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < sd.npts_pad; p += gridDim.x * blockDim.x) {
    double2* out = table + sd.tab_off + p;
    if (p >= sd.npts) {
      for (int c = 0; c < sd.nlm; c++) out[(long)c * sd.npts_pad] = make_double2(0, 0);
      continue;
    }
    double v[3], r;
    if (mode == 0) {
      const int g = idx[sd.pt_off + p];
      const int i = g / (n1 * n2), rem = g % (n1 * n2);
      const double fr[3] = {(double)i / n0, (double)(rem / n2) / n1, (double)(rem % n2) / n2};
      r = INFINITY;
      for (int a = -1; a <= 1; a++)
        for (int b = -1; b <= 1; b++)
          for (int c = -1; c <= 1; c++) {
            const double t0 = fr[0] + a - sd.coord[0], t1 = fr[1] + b - sd.coord[1],
                         t2 = fr[2] + c - sd.coord[2];
            const double x = t0 * lattice[0] + t1 * lattice[3] + t2 * lattice[6];
            const double y = t0 * lattice[1] + t1 * lattice[4] + t2 * lattice[7];
            const double z = t0 * lattice[2] + t1 * lattice[5] + t2 * lattice[8];
            const double d = sqrt(x * x + y * y + z * z);
            if (d < r) {
              r = d;
              v[0] = x; v[1] = y; v[2] = z;
            }
          }
    } else {
      v[0] = path[sd.pt_off + p];
      v[1] = path[path_ld + sd.pt_off + p];
      v[2] = path[2 * path_ld + sd.pt_off + p];
      r = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    }
    double theta, phi;
    d_angles(v, r, &theta, &phi);
    int last_n = -1;
    double rad = 0;
    for (int c = 0; c < sd.nlm; c++) {
      const int fn = ed.chan_n[c];
      if (fn != last_n) {
        const double* f = ed.f + (long)fn * ed.n;
        const double* sp = ed.spl + (long)fn * 3 * ed.n;
        if (mode == 0) {
          rad = d_eval_linear(r, ed.rmax, ed.n, ed.grid, f, sp);
        } else {
          rad = d_eval_log(r, ed.n, ed.grid, f, sp);
          rad = (r < ed.grid[0]) ? rad / ed.grid[0] : rad / r;
        }
        last_n = fn;
      }
      const double2 y = d_ylm(ed.chan_l[c], ed.chan_m[c], theta, phi);
      out[(long)c * sd.npts_pad] = make_double2(rad * y.x, rad * y.y);
    }
  }
}

// Per-k projector table: Tk = conj(T) * dv * exp(i k_cart . path)   [projector.c:259-263, 269]
// (the phase does not depend on the band, so it is folded into the table once per k-point)
__global__ void __launch_bounds__(256)
phase_table_kernel(const SiteDev* __restrict__ sites, const double* __restrict__ path, long path_ld,
                   const double2* __restrict__ table, double2* __restrict__ tablek, double kx,
                   double ky, double kz, double dv) {
  const SiteDev sd = sites[blockIdx.y];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < sd.npts_pad; p += gridDim.x * blockDim.x) {
    const long q = sd.pt_off + p;
    const double kr = kx * path[q] + ky * path[path_ld + q] + kz * path[2 * path_ld + q];
    double s, c;
    sincos(kr, &s, &c);
    const double2 ph = make_double2(dv * c, dv * s);
    for (int ch = 0; ch < sd.nlm; ch++) {
      const long o = sd.tab_off + (long)ch * sd.npts_pad + p;
      const double2 t = table[o];
      tablek[o] = cmul(make_double2(t.x, -t.y), ph);
    }
  }
}

// ---------------------------------------------------------------------------------------
// (a5) <p_i|psi~> : sphere gather + contraction on FP64 tensor cores  [projector.c:245-272]
//   P[slot][lm_off + lm] = sum_pt Tk[lm][pt] * x[slot][idx[pt]]
// CTA = one site x 32 grid slots (bands); 4 warps, each owns 8 slots (one n-tile) and all
// m-tiles.  Sphere samples are gathered with 16-B cp.async into a KT-point stage ring
// (the gather is the HBM-bound stream: 16 B per point per band, read once); the table tile is
// staged alongside (re-read from L2 by the other band blocks of the same site).
// ---------------------------------------------------------------------------------------
constexpr int PROJ_IL = 16;       // band interleave of the pruned-FFT boxes (== FFT_B)
constexpr int PROJ_KT = 32;       // points per stage
constexpr int PROJ_NB = 32;       // grid slots per CTA
constexpr int PROJ_STAGES = 4;
constexpr int PROJ_LDB = PROJ_NB + 2;   // double2 row stride of the sample tile (bank spread)
constexpr int PROJ_LDA = PROJ_KT + 4;   // double2 row stride of the table tile

template <int MT>   // m-tiles of 8 channels
__global__ void __launch_bounds__(128)
sphere_project_kernel(const SiteDev* __restrict__ sites, const int* __restrict__ site_list,
                      const int* __restrict__ idx, const double2* __restrict__ tablek,
                      const double2* __restrict__ x, long ngrid, int nslot, double2* __restrict__ P,
                      long ldp, int slot0) {
  const SiteDev sd = sites[site_list[blockIdx.y]];   // sites with exactly MT m-tiles
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* sB = reinterpret_cast<double2*>(smem_raw);                       // [ST][KT][LDB]
  double2* sA = sB + PROJ_STAGES * PROJ_KT * PROJ_LDB;                      // [ST][8*MT][LDA]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sbase = blockIdx.x * PROJ_NB;          // first slot of this CTA (within the batch)
  const int nk = sd.npts_pad / PROJ_KT;

  // zero the channel-padding rows of every stage once (never overwritten)
  for (int e = tid; e < PROJ_STAGES * 8 * MT * PROJ_LDA; e += 128) {
    const int row = (e / PROJ_LDA) % (8 * MT);
    if (row >= sd.nlm) sA[e] = make_double2(0, 0);
  }

  auto issue = [&](int kt, int st) {
    if (kt < nk) {
      const int g = __ldg(idx + sd.pt_off + kt * PROJ_KT + lane);
      double2* dstB = sB + (st * PROJ_KT + lane) * PROJ_LDB;
#pragma unroll
      for (int q = 0; q < PROJ_NB / 4; q++) {
        const int sl = warp + 4 * q;
        int s = sbase + sl;
        if (s >= nslot) s = nslot - 1;   // tail CTA: duplicate the last slot, discarded on store
        cp_async16(dstB + sl, x + (long)s * ngrid + g);
      }
      // table tile: nlm rows x 32 points; 128 threads cover 4 rows per pass
      for (int row = warp; row < sd.nlm; row += 4)
        cp_async16(sA + (st * 8 * MT + row) * PROJ_LDA + lane,
                   tablek + sd.tab_off + (long)row * sd.npts_pad + kt * PROJ_KT + lane);
    }
    cp_async_commit();
  };

  double rr[MT][2], ii[MT][2], ri[MT][2];
#pragma unroll
  for (int m = 0; m < MT; m++) rr[m][0] = rr[m][1] = ii[m][0] = ii[m][1] = ri[m][0] = ri[m][1] = 0;

#pragma unroll
  for (int s = 0; s < PROJ_STAGES - 1; s++) issue(s, s);

  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<PROJ_STAGES - 2>();
    __syncthreads();
    issue(kt + PROJ_STAGES - 1, (kt + PROJ_STAGES - 1) % PROJ_STAGES);
    const int st = kt % PROJ_STAGES;
    const double2* tB = sB + st * PROJ_KT * PROJ_LDB;
    const double2* tA = sA + st * 8 * MT * PROJ_LDA;
#pragma unroll
    for (int kk = 0; kk < PROJ_KT / 4; kk++) {
      const double2 b = tB[(4 * kk + (lane & 3)) * PROJ_LDB + 8 * warp + (lane >> 2)];
#pragma unroll
      for (int m = 0; m < MT; m++) {
        const double2 a = tA[(8 * m + (lane >> 2)) * PROJ_LDA + 4 * kk + (lane & 3)];
        dmma884(rr[m][0], rr[m][1], a.x, b.x);
        dmma884(ii[m][0], ii[m][1], a.y, b.y);
        dmma884(ri[m][0], ri[m][1], a.x, b.y);
        dmma884(ri[m][0], ri[m][1], a.y, b.x);
      }
    }
  }
  cp_async_wait<0>();

  // D fragment: row (channel) = lane>>2, cols (slots) = 2*(lane&3)+{0,1}
#pragma unroll
  for (int m = 0; m < MT; m++) {
    const int ch = 8 * m + (lane >> 2);
    if (ch < sd.nlm) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int s = sbase + 8 * warp + 2 * (lane & 3) + c;
        if (s < nslot)
          P[(long)(slot0 + s) * ldp + sd.lm_off + ch] = make_double2(rr[m][c] - ii[m][c], ri[m][c]);
      }
    }
  }
}

// Same contraction on band-interleaved boxes X[group][grid point][IL] (IL = 16 slots = 256 B per point,
// produced by the pruned FFT of fft3d.cuh): every gathered point is two full 256-B segments, so the DRAM
// traffic equals the algorithmic 16 B per (point, band).
template <int MT>
__global__ void __launch_bounds__(128)
sphere_project_il_kernel(const SiteDev* __restrict__ sites, const int* __restrict__ site_list,
                         const int* __restrict__ idx, const double2* __restrict__ tablek,
                         const double2* __restrict__ X, long ngrid, int nslot, int ngroups,
                         double2* __restrict__ P, long ldp, int slot0, int idx_cap) {
  constexpr int IL = PROJ_IL;
  const SiteDev sd = sites[site_list[blockIdx.y]];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* sB = reinterpret_cast<double2*>(smem_raw);                       // [ST][KT][LDB]
  double2* sA = sB + PROJ_STAGES * PROJ_KT * PROJ_LDB;                      // [ST][8*MT][LDA]
  int* sIdx = reinterpret_cast<int*>(sA + PROJ_STAGES * 8 * MT * PROJ_LDA);  // [idx_cap] sphere index list
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sbase = blockIdx.x * PROJ_NB;
  const int nk = sd.npts_pad / PROJ_KT;
  // the gather addresses come from the index list: keep it in shared memory so the cp.async issue does
  // not wait on a global load per stage
  for (int e = tid; e < sd.npts_pad && e < idx_cap; e += 128) sIdx[e] = __ldg(idx + sd.pt_off + e);
  for (int e = tid; e < PROJ_STAGES * 8 * MT * PROJ_LDA; e += 128) {
    const int row = (e / PROJ_LDA) % (8 * MT);
    if (row >= sd.nlm) sA[e] = make_double2(0, 0);
  }
  __syncthreads();
  int grp = (sbase + lane) / IL;
  if (grp >= ngroups) grp = ngroups - 1;          // tail CTA: duplicate the last group, discarded on store
  const double2* xsrc = X + (long)grp * ngrid * IL + (lane & (IL - 1));

  auto issue = [&](int kt, int st) {
    if (kt < nk) {
#pragma unroll
      for (int q = 0; q < PROJ_KT / 4; q++) {
        const int pt = warp + 4 * q;
        const int ip = kt * PROJ_KT + pt;
        const int g = ip < idx_cap ? sIdx[ip] : __ldg(idx + sd.pt_off + ip);   // warp-uniform -> broadcast
        cp_async16(sB + (st * PROJ_KT + pt) * PROJ_LDB + lane, xsrc + (long)g * IL);
      }
      for (int row = warp; row < sd.nlm; row += 4)
        cp_async16(sA + (st * 8 * MT + row) * PROJ_LDA + lane,
                   tablek + sd.tab_off + (long)row * sd.npts_pad + kt * PROJ_KT + lane);
    }
    cp_async_commit();
  };

  double rr[MT][2], ii[MT][2], ri[MT][2];
#pragma unroll
  for (int m = 0; m < MT; m++) rr[m][0] = rr[m][1] = ii[m][0] = ii[m][1] = ri[m][0] = ri[m][1] = 0;
#pragma unroll
  for (int s = 0; s < PROJ_STAGES - 1; s++) issue(s, s);
  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<PROJ_STAGES - 2>();
    __syncthreads();
    issue(kt + PROJ_STAGES - 1, (kt + PROJ_STAGES - 1) % PROJ_STAGES);
    const int st = kt % PROJ_STAGES;
    const double2* tB = sB + st * PROJ_KT * PROJ_LDB;
    const double2* tA = sA + st * 8 * MT * PROJ_LDA;
#pragma unroll
    for (int kk = 0; kk < PROJ_KT / 4; kk++) {
      const double2 b = tB[(4 * kk + (lane & 3)) * PROJ_LDB + 8 * warp + (lane >> 2)];
#pragma unroll
      for (int m = 0; m < MT; m++) {
        const double2 a = tA[(8 * m + (lane >> 2)) * PROJ_LDA + 4 * kk + (lane & 3)];
        dmma884(rr[m][0], rr[m][1], a.x, b.x);
        dmma884(ii[m][0], ii[m][1], a.y, b.y);
        dmma884(ri[m][0], ri[m][1], a.x, b.y);
        dmma884(ri[m][0], ri[m][1], a.y, b.x);
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int m = 0; m < MT; m++) {
    const int ch = 8 * m + (lane >> 2);
    if (ch < sd.nlm) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int s = sbase + 8 * warp + 2 * (lane & 3) + c;
        if (s < nslot)
          P[(long)(slot0 + s) * ldp + sd.lm_off + ch] = make_double2(rr[m][c] - ii[m][c], ri[m][c]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// (a5) the same contraction with REAL tables: half the FP64 tensor work and half the table bytes
// ---------------------------------------------------------------------------------------
// The table value of a channel is R_n(r) Y_lm(r^) (projector.c:259-271 contracts conj(value) with the sample), and
// Y_l,-m = (-1)^m conj(Y_l,m) (utils.c:377-449 builds both from the same Legendre / cexp factors; they agree to
// rounding).  So per radial channel n only the 2l+1 REAL functions
//     row(m > 0) = Re T(n,+m),   row(m = 0) = T(n,0) (real),   row(m < 0) = Im T(n,+|m|)
// are contracted with the sample x' = x * dv * exp(i k.path) (the band-independent Bloch phase now multiplies the
// sample, exactly where projector.c:262 applies it): Q[row] = sum_pt U[row][pt] x'[pt] is 2 DMMA per k-step instead
// of 4, and the channel values follow in the epilogue,
//     P(+m) = Q(+m) - i Q(-m),    P(-m) = (-1)^m (Q(+m) + i Q(-m)),    P(0) = Q(0).
// U does not depend on k (built once per table set); per k only the S phase factors are rebuilt.
__global__ void __launch_bounds__(256)
real_table_kernel(const SiteDev* __restrict__ sites, const int* __restrict__ chan_m,
                  const double2* __restrict__ table, double* __restrict__ ureal) {
  const SiteDev sd = sites[blockIdx.y];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < sd.npts_pad; p += gridDim.x * blockDim.x)
    for (int c = 0; c < sd.nlm; c++) {
      const int m = chan_m[sd.lm_off + c];
      const long o = sd.tab_off + (long)c * sd.npts_pad + p;
      ureal[o] = m >= 0 ? table[o].x : table[o - 2L * m * sd.npts_pad].y;     // partner channel (n, +|m|) is c + 2|m|
    }
}

// phk[pt] = dv * exp(i k_cart . path[pt])   [projector.c:259-263]
__global__ void __launch_bounds__(256)
phase_points_kernel(const double* __restrict__ path, long path_ld, long npts, double2* __restrict__ phk, double kx,
                    double ky, double kz, double dv) {
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < npts; q += (long)gridDim.x * blockDim.x) {
    const double kr = kx * path[q] + ky * path[path_ld + q] + kz * path[2 * path_ld + q];
    double s, c;
    sincos(kr, &s, &c);
    phk[q] = make_double2(dv * c, dv * s);
  }
}

constexpr int PROJ_LDU = PROJ_KT + 4;   // double row stride of the real table tile (rows 0..3 cover all 32 banks)

constexpr int PROJ_RSTAGES = 3;         // real-table kernel: 3 stages and no index list in shared memory -> 3 CTAs/SM

template <int MT>
inline size_t sphere_project_real_smem() {
  return (size_t)PROJ_RSTAGES * (sizeof(double2) * (PROJ_KT * PROJ_LDB + PROJ_KT) + sizeof(double) * 8 * MT * PROJ_LDU);
}

// 8 warps per CTA: warps w and w + 4 work on the same 8 slots and split the k-steps of every stage between them
// (even / odd), so an SM holds twice the warps for the same shared memory - the 4-warp version left the DMMA pipe idle
// between the dependent issue slots of its two warps per scheduler (ncu r02: "wait" 44 % of the stall samples).
// The sphere indices a warp needs for a stage (4 of them) are fetched one iteration ahead into a register of lanes
// 0..3 and broadcast by shuffle, so no index list occupies shared memory: with 3 stages even the 18-channel tile
// (74.5 KB) fits three times per SM = 24 warps.
constexpr int PROJ_THREADS = 256;

template <int MT>
__global__ void __launch_bounds__(PROJ_THREADS)
sphere_project_real_kernel(const SiteDev* __restrict__ sites, const int* __restrict__ site_list,
                           const int* __restrict__ idx, const double* __restrict__ ureal,
                           const double2* __restrict__ phk, const int* __restrict__ chan_m,
                           const double2* __restrict__ X, long ngrid, int nslot, int ngroups,
                           double2* __restrict__ P, long ldp, int slot0, int) {
  constexpr int IL = PROJ_IL;
  constexpr int NS = PROJ_RSTAGES;
  const SiteDev sd = sites[site_list[blockIdx.y]];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* sB = reinterpret_cast<double2*>(smem_raw);                       // [NS][KT][LDB] samples
  double2* sPh = sB + NS * PROJ_KT * PROJ_LDB;                              // [NS][KT]      phase factors
  double* sU = reinterpret_cast<double*>(sPh + NS * PROJ_KT);               // [NS][8*MT][LDU] real table rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wq = warp & 3;                          // slot octet of this warp
  const int wh = warp >> 2;                         // which half of the k-steps of a stage
  const int sbase = blockIdx.x * PROJ_NB;
  const int nk = sd.npts_pad / PROJ_KT;
  for (int e = tid; e < NS * 8 * MT * PROJ_LDU; e += PROJ_THREADS) {
    const int row = (e / PROJ_LDU) % (8 * MT);
    if (row >= sd.nlm) sU[e] = 0.0;                 // channel padding rows, never overwritten
  }
  __syncthreads();
  int grp = (sbase + lane) / IL;
  if (grp >= ngroups) grp = ngroups - 1;          // tail CTA: duplicate the last group, discarded on store
  const double2* xsrc = X + (long)grp * ngrid * IL + (lane & (IL - 1));
  const int* myidx = idx + sd.pt_off + warp + 8 * (lane & 3);               // lane q < 4: point warp + 8 q of a stage

  // sphere indices of this warp's four points of stage kt (valid in lanes 0..3)
  auto fetch = [&](int kt) { return kt < nk ? __ldg(myidx + kt * PROJ_KT) : 0; };
  auto issue = [&](int kt, int st, int gidx) {
    if (kt < nk) {
#pragma unroll
      for (int q = 0; q < PROJ_KT / 8; q++) {
        const int pt = warp + 8 * q;
        const int g = __shfl_sync(0xffffffffu, gidx, q);
        cp_async16(sB + (st * PROJ_KT + pt) * PROJ_LDB + lane, xsrc + (long)g * IL);
      }
      // real table tile: nlm rows x 32 points (16 chunks of 16 B per row); a warp covers two rows per pass
      for (int row = 2 * warp + (lane >> 4); row < sd.nlm; row += 16)
        cp_async16(sU + (st * 8 * MT + row) * PROJ_LDU + 2 * (lane & 15),
                   ureal + sd.tab_off + (long)row * sd.npts_pad + kt * PROJ_KT + 2 * (lane & 15));
      if (warp == 7) cp_async16(sPh + st * PROJ_KT + lane, phk + sd.pt_off + kt * PROJ_KT + lane);
    }
    cp_async_commit();
  };

  double qr[MT][2], qi[MT][2];
#pragma unroll
  for (int m = 0; m < MT; m++) qr[m][0] = qr[m][1] = qi[m][0] = qi[m][1] = 0;
#pragma unroll
  for (int s = 0; s < NS - 1; s++) issue(s, s, fetch(s));
  int gnext = fetch(NS - 1);                        // indices of the stage issued in iteration 0
  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<NS - 2>();
    __syncthreads();
    const int gcur = gnext;
    gnext = fetch(kt + NS);                         // one iteration ahead of its use
    issue(kt + NS - 1, (kt + NS - 1) % NS, gcur);
    const int st = kt % NS;
    const double2* tB = sB + st * PROJ_KT * PROJ_LDB;
    const double2* tP = sPh + st * PROJ_KT;
    const double* tU = sU + st * 8 * MT * PROJ_LDU;
#pragma unroll
    for (int k2 = 0; k2 < PROJ_KT / 8; k2++) {
      const int kk = 2 * k2 + wh;
      const double2 x = tB[(4 * kk + (lane & 3)) * PROJ_LDB + 8 * wq + (lane >> 2)];
      const double2 ph = tP[4 * kk + (lane & 3)];
      const double bx = x.x * ph.x - x.y * ph.y, by = x.x * ph.y + x.y * ph.x;     // x' = x dv e^{i k.r}
#pragma unroll
      for (int m = 0; m < MT; m++) {
        const double u = tU[(8 * m + (lane >> 2)) * PROJ_LDU + 4 * kk + (lane & 3)];
        dmma884(qr[m][0], qr[m][1], u, bx);
        dmma884(qi[m][0], qi[m][1], u, by);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                 // every warp is done with the stage buffers: reuse sB for Q
  // Q fragment: row (channel) = 8m + lane>>2, cols (slots) = 2*(lane&3)+{0,1}  ->  sQ[octet][channel][8 slots]
  double2* sQ = sB + wq * (8 * MT * 8);
  if (wh == 1) {
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int c = 0; c < 2; c++)
        sQ[(8 * m + (lane >> 2)) * 8 + 2 * (lane & 3) + c] = make_double2(qr[m][c], qi[m][c]);
  }
  __syncthreads();
  if (wh == 1) return;
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int c = 0; c < 2; c++) {
      double2& e = sQ[(8 * m + (lane >> 2)) * 8 + 2 * (lane & 3) + c];      // same thread wrote / reads this element
      e = make_double2(e.x + qr[m][c], e.y + qi[m][c]);
    }
  __syncwarp();
#pragma unroll
  for (int m = 0; m < MT; m++) {
    const int ch = 8 * m + (lane >> 2);
    if (ch < sd.nlm) {
      const int mm = __ldg(chan_m + sd.lm_off + ch);
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int col = 2 * (lane & 3) + c;
        const int s = sbase + 8 * wq + col;
        if (s >= nslot) continue;
        const double2 q0 = sQ[ch * 8 + col];
        double2 out = q0;
        if (mm > 0) {                               // P(+m) = Q(+m) - i Q(-m)
          const double2 qs = sQ[(ch - 2 * mm) * 8 + col];
          out = make_double2(q0.x + qs.y, q0.y - qs.x);
        } else if (mm < 0) {                        // P(-m) = (-1)^m (Q(+m) + i Q(-m))
          const double2 qc = sQ[(ch - 2 * mm) * 8 + col];
          const double sg = (mm & 1) ? -1.0 : 1.0;
          out = make_double2(sg * (qc.x - q0.y), sg * (qc.y + q0.x));
        }
        P[(long)(slot0 + s) * ldp + sd.lm_off + ch] = out;
      }
    }
  }
}

inline size_t sphere_project_smem(int MT) {
  return sizeof(double2) * PROJ_STAGES * (PROJ_KT * PROJ_LDB + 8 * MT * PROJ_LDA);
}

// ---------------------------------------------------------------------------------------
// augmentation operands for the one-centre GEMM  [projector.c:890-959 in matrix form]
// ---------------------------------------------------------------------------------------
// Generic "gather + small block matmul":  dst[row][dst_off + i] = sum_j Mat[i][j] * src[row][src_off + j]
// (Mat == nullptr: plain copy of ni entries).  One block entry per (site pair); rows = bands.
struct BlockOp {
  int src_off, dst_off, ni, nj;
  long mat_off;             // offset into the matrix pool (double2), -1 for identity copy
};
__global__ void __launch_bounds__(128)
block_apply_kernel(const BlockOp* __restrict__ ops, const double2* __restrict__ mats,
                   const double2* __restrict__ src, long lds, double2* __restrict__ dst, long ldd,
                   int nrows) {
  const BlockOp op = ops[blockIdx.y];
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const double2* s = src + (long)row * lds + op.src_off;
    double2* d = dst + (long)row * ldd + op.dst_off;
    for (int i = threadIdx.x; i < op.ni; i += blockDim.x) {
      if (op.mat_off < 0) {
        d[i] = s[i];
      } else {
        const double2* M = mats + op.mat_off + (long)i * op.nj;
        double2 acc = make_double2(0, 0);
        for (int j = 0; j < op.nj; j++) {
          const double2 t = cmul(M[j], s[j]);
          acc.x += t.x;
          acc.y += t.y;
        }
        d[i] = acc;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// (a12) real-space AE state  [density.c:241-253, 297-304]
// ---------------------------------------------------------------------------------------
// x[g] *= exp(sign * 2 pi i k.r_frac), one thread per grid point
__global__ void __launch_bounds__(256)
bloch_phase_kernel(double2* __restrict__ x, int n0, int n1, int n2, double kx, double ky,
                   double kz, double sign, int nbox) {
  const double PI = 3.14159265359;   // density.c:13
  const long ngrid = (long)n0 * n1 * n2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < ngrid; g += stride) {
    const int i = (int)(g / ((long)n1 * n2));
    const int rem = (int)(g % ((long)n1 * n2));
    const double kr = kx * ((double)i / n0) + ky * ((double)(rem / n2) / n1) +
                      kz * ((double)(rem % n2) / n2);
    double s, c;
    sincos(sign * 2 * PI * kr, &s, &c);
    for (int b = 0; b < nbox; b++) {
      const double2 v = x[(long)b * ngrid + g];
      x[(long)b * ngrid + g] = make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
    }
  }
}

// x[idx[pt]] += e^{2 pi i k.(R + wrap)} * sum_lm A[lm][pt] * P[lm]   (one warp-wide pass per site)
// Spheres of neighbouring atoms may overlap, so the accumulation uses FP64 atomics (the
// reference races here under OpenMP, density.c:256-304).
__global__ void __launch_bounds__(256)
augment_add_kernel(const SiteDev* __restrict__ sites, const int* __restrict__ idx,
                   const int* __restrict__ wrap, long wrap_ld, const double2* __restrict__ table,
                   const double2* __restrict__ P, long ldp, int nbox, double2* __restrict__ x,
                   long ngrid, double kx, double ky, double kz) {
  const double PI = 3.14159265359;   // density.c:13
  const SiteDev sd = sites[blockIdx.y];
  extern __shared__ double2 sP[];    // [nbox][nlm]
  for (int e = threadIdx.x; e < nbox * sd.nlm; e += blockDim.x)
    sP[e] = P[(long)(e / sd.nlm) * ldp + sd.lm_off + (e % sd.nlm)];
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < sd.npts; p += gridDim.x * blockDim.x) {
    const long q = sd.pt_off + p;
    const double ph = (sd.coord[0] + wrap[q]) * kx + (sd.coord[1] + wrap[wrap_ld + q]) * ky +
                      (sd.coord[2] + wrap[2 * wrap_ld + q]) * kz;
    double s, c;
    sincos(2 * PI * ph, &s, &c);
    const int g = idx[q];
    for (int b = 0; b < nbox; b++) {
      double2 acc = make_double2(0, 0);
      for (int ch = 0; ch < sd.nlm; ch++) {
        const double2 t = cmul(table[sd.tab_off + (long)ch * sd.npts_pad + p], sP[b * sd.nlm + ch]);
        acc.x += t.x;
        acc.y += t.y;
      }
      double* dst = reinterpret_cast<double*>(x + (long)b * ngrid + g);
      atomicAdd(dst, acc.x * c - acc.y * s);
      atomicAdd(dst + 1, acc.x * s + acc.y * c);
    }
  }
}

// (f3) augmentation part of a band on the FFT grid  [projector.c:276-331, get_aug_freqs_helper]:
//   x[b][idx[pt]] += e^{-i k_cart.path[pt]} * sum_ch T[ch][pt] * P[b][full_off[site] + ch]
// T = filtered (phi - phit) tables of the listed sites, P = the band's projector overlaps (full channel axis).
__global__ void __launch_bounds__(256)
aug_freq_add_kernel(const SiteDev* __restrict__ sites, const int* __restrict__ full_off,
                    const int* __restrict__ idx, const double* __restrict__ path, long path_ld,
                    const double2* __restrict__ table, const double2* __restrict__ P, long ldp, int nbox,
                    double2* __restrict__ x, long ngrid, double kx, double ky, double kz) {
  const SiteDev sd = sites[blockIdx.y];
  const int off = full_off[blockIdx.y];
  extern __shared__ double2 sP[];    // [nbox][nlm]
  for (int e = threadIdx.x; e < nbox * sd.nlm; e += blockDim.x)
    sP[e] = P[(long)(e / sd.nlm) * ldp + off + (e % sd.nlm)];
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < sd.npts; p += gridDim.x * blockDim.x) {
    const long q = sd.pt_off + p;
    const double kr = -(kx * path[q] + ky * path[path_ld + q] + kz * path[2 * path_ld + q]);
    double s, c;
    sincos(kr, &s, &c);
    const int g = idx[q];
    for (int b = 0; b < nbox; b++) {
      double2 acc = make_double2(0, 0);
      for (int ch = 0; ch < sd.nlm; ch++) {
        const double2 t = cmul(sP[b * sd.nlm + ch], table[sd.tab_off + (long)ch * sd.npts_pad + p]);
        acc.x += t.x;
        acc.y += t.y;
      }
      double* dst = reinterpret_cast<double*>(x + (long)b * ngrid + g);
      atomicAdd(dst, acc.x * c - acc.y * s);
      atomicAdd(dst + 1, acc.x * s + acc.y * c);
    }
  }
}

// Deterministic variants: ONE site per launch (the launches of a site list run back to back on the stream), so every
// grid point is touched by exactly one thread of a launch and the sums over overlapping spheres are formed in site
// order without atomics.  `planar`: boxes x[b][g];  interleaved: x[group][g][16] for the pruned forward transform.
__global__ void __launch_bounds__(256)
aug_freq_add_site_kernel(const SiteDev* __restrict__ sites, int site, int full_off, const int* __restrict__ idx,
                         const double* __restrict__ path, long path_ld, const double2* __restrict__ table,
                         const double2* __restrict__ P, long ldp, int nbox, double2* __restrict__ x, long ngrid,
                         double kx, double ky, double kz, int interleaved) {
  const SiteDev sd = sites[site];
  extern __shared__ double2 sP[];    // [nbox][nlm]
  for (int e = threadIdx.x; e < nbox * sd.nlm; e += blockDim.x)
    sP[e] = P[(long)(e / sd.nlm) * ldp + full_off + (e % sd.nlm)];
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < sd.npts; p += gridDim.x * blockDim.x) {
    const long q = sd.pt_off + p;
    const double kr = -(kx * path[q] + ky * path[path_ld + q] + kz * path[2 * path_ld + q]);
    double s, c;
    sincos(kr, &s, &c);
    const int g = idx[q];
    for (int b = 0; b < nbox; b++) {
      double2 acc = make_double2(0, 0);
      for (int ch = 0; ch < sd.nlm; ch++) {
        const double2 t = cmul(sP[b * sd.nlm + ch], table[sd.tab_off + (long)ch * sd.npts_pad + p]);
        acc.x += t.x;
        acc.y += t.y;
      }
      double2* dst = interleaved ? x + ((long)(b >> 4) * ngrid + g) * 16 + (b & 15) : x + (long)b * ngrid + g;
      double2 v = *dst;
      v.x += acc.x * c - acc.y * s;
      v.y += acc.x * s + acc.y * c;
      *dst = v;
    }
  }
}

// batched (a3): Cout[b][w] = (complex64) scale * x[b][gidx[w]]
__global__ void __launch_bounds__(256)
gather_pw_batch_kernel(const double2* __restrict__ x, long ngrid, const int* __restrict__ gidx,
                       float2* __restrict__ Cout, long ldc, int npw, double scale) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= npw) return;
  const int g = gidx[w];
  const double2 v = x[(long)blockIdx.y * ngrid + g];
  Cout[(long)blockIdx.y * ldc + w] = make_float2((float)(v.x * scale), (float)(v.y * scale));
}

// ---- band-interleaved variants (boxes from the pruned transform: X[group][g][16]) of the real-space kernels ----
// x *= exp(sign 2 pi i k.r_frac): a warp takes 32 consecutive grid points, each lane evaluates one phase, and the
// 16 slots of two points are processed per step (512 contiguous bytes per warp access).
__global__ void __launch_bounds__(256)
bloch_phase_il_kernel(double2* __restrict__ x, int n0, int n1, int n2, double kx, double ky, double kz,
                      double sign, int ngroups) {
  const double PI = 3.14159265359;   // density.c:13
  const long ngrid = (long)n0 * n1 * n2;
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long g0 = warp * 32; g0 < ngrid; g0 += nwarps * 32) {
    const long g = g0 + lane;
    double s = 0, c = 1;
    if (g < ngrid) {
      const int i = (int)(g / ((long)n1 * n2));
      const int rem = (int)(g % ((long)n1 * n2));
      const double kr = kx * ((double)i / n0) + ky * ((double)(rem / n2) / n1) + kz * ((double)(rem % n2) / n2);
      sincos(sign * 2 * PI * kr, &s, &c);
    }
    for (int it = 0; it < 16; it++) {
      const int src = 2 * it + (lane >> 4);
      const double ss = __shfl_sync(0xffffffffu, s, src), cc = __shfl_sync(0xffffffffu, c, src);
      const long gg = g0 + src;
      if (gg >= ngrid) continue;
      for (int grp = 0; grp < ngroups; grp++) {
        double2* p = x + ((long)grp * ngrid + gg) * 16 + (lane & 15);
        const double2 v = *p;
        *p = make_double2(v.x * cc - v.y * ss, v.x * ss + v.y * cc);
      }
    }
  }
}

// augment_add_kernel on interleaved boxes: slot b of the batch lives at x[((b >> 4) * ngrid + g) * 16 + (b & 15)]
__global__ void __launch_bounds__(256)
augment_add_il_kernel(const SiteDev* __restrict__ sites, const int* __restrict__ idx,
                      const int* __restrict__ wrap, long wrap_ld, const double2* __restrict__ table,
                      const double2* __restrict__ P, long ldp, int b0, int nbox, double2* __restrict__ x,
                      long ngrid, double kx, double ky, double kz) {
  const double PI = 3.14159265359;   // density.c:13
  const SiteDev sd = sites[blockIdx.y];
  extern __shared__ double2 sP[];    // [nbox][nlm]
  for (int e = threadIdx.x; e < nbox * sd.nlm; e += blockDim.x)
    sP[e] = P[(long)(b0 + e / sd.nlm) * ldp + sd.lm_off + (e % sd.nlm)];
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < sd.npts; p += gridDim.x * blockDim.x) {
    const long q = sd.pt_off + p;
    const double ph = (sd.coord[0] + wrap[q]) * kx + (sd.coord[1] + wrap[wrap_ld + q]) * ky +
                      (sd.coord[2] + wrap[2 * wrap_ld + q]) * kz;
    double s, c;
    sincos(2 * PI * ph, &s, &c);
    const int g = idx[q];
    for (int b = 0; b < nbox; b++) {
      double2 acc = make_double2(0, 0);
      for (int ch = 0; ch < sd.nlm; ch++) {
        const double2 t = cmul(table[sd.tab_off + (long)ch * sd.npts_pad + p], sP[b * sd.nlm + ch]);
        acc.x += t.x;
        acc.y += t.y;
      }
      const int slot = b0 + b;
      double* dst = reinterpret_cast<double*>(x + ((long)(slot >> 4) * ngrid + g) * 16 + (slot & 15));
      atomicAdd(dst, acc.x * c - acc.y * s);
      atomicAdd(dst + 1, acc.x * s + acc.y * c);
    }
  }
}

// rho[g] += sum_slot w[slot] |x_slot[g]|^2 on interleaved boxes: half a warp per grid point, one lane per slot
__global__ void __launch_bounds__(256)
density_accum_il_kernel(const double2* __restrict__ x, long ngrid, int ngroups, const double* __restrict__ w,
                        double* __restrict__ rho) {
  const int lane = threadIdx.x & 31, b = lane & 15;
  const long half = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const long nhalf = ((long)gridDim.x * blockDim.x) >> 4;
  for (long g0 = 0; g0 < ngrid; g0 += nhalf) {          // uniform trip count per warp (shuffles inside)
    const long g = g0 + half;
    double a = 0;
    if (g < ngrid)
      for (int grp = 0; grp < ngroups; grp++) {
        const double2 v = x[((long)grp * ngrid + g) * 16 + b];
        a += (v.x * v.x + v.y * v.y) * w[grp * 16 + b];
      }
    a += __shfl_xor_sync(0xffffffffu, a, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (b == 0 && g < ngrid) rho[g] += a;
  }
}

// planar boxes out[j][g] (j < nslots) from the interleaved group layout x[group][g][16], slots slot0 .. slot0+nslots-1
// (slot0 is relative to the first slot of x): the caller-layout copy of single real-space states
__global__ void __launch_bounds__(256)
extract_il_kernel(const double2* __restrict__ x, long ngrid, int slot0, int nslots, double2* __restrict__ out) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < ngrid * nslots; e += stride) {
    const int j = (int)(e / ngrid);
    const long g = e % ngrid;
    const int slot = slot0 + j;
    out[e] = x[((long)(slot >> 4) * ngrid + g) * 16 + (slot & 15)];
  }
}

// the reverse: planar boxes in[j][g] into slots slot0 .. of the (pre-zeroed) interleaved layout
__global__ void __launch_bounds__(256)
insert_il_kernel(const double2* __restrict__ in, long ngrid, int slot0, int nslots, double2* __restrict__ x) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < ngrid * nslots; e += stride) {
    const int j = (int)(e / ngrid);
    const long g = e % ngrid;
    const int slot = slot0 + j;
    x[((long)(slot >> 4) * ngrid + g) * 16 + (slot & 15)] = in[e];
  }
}

// ---------------------------------------------------------------------------------------
// (f4) momentum matrix elements  [momentum.c]
// ---------------------------------------------------------------------------------------
struct MomTerm {          // one (radial function, L, M) term of an element: F = f(|G|) * Y_LM(G^) * cfac
  int slot;               // radial table: data[slot_off] = f[N], then 3N spline coefficients
  int L, M;
  int special0;           // value at |G| = 0: 1 -> Y(L,M;0,0) kept (the reference's L==0 && m1==m2 / r==0 rule), 0 -> zero
};
struct MomElem {          // per element
  int nterm, term_off;    // terms of this element in the concatenated term list
  int N;                  // radial grid size
  long ks_off;            // offset of the k grid in `radial`
};

// pseudo part: out[g] = sum_w2 conj(C1[box(G2[w2] + GP_g)]) * C2[w2]   [momentum.c:47-107], one warp per GP.
__global__ void __launch_bounds__(256)
momentum_pseudo_kernel(int numg, const int* __restrict__ igall, const float2* __restrict__ C1,
                       const float2* __restrict__ C2, const int* __restrict__ G2, int npw2,
                       const int* __restrict__ boxmap, int lo0, int lo1, int lo2, int d0, int d1, int d2,
                       double2* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= numg) return;
  const int p0 = igall[3 * g], p1 = igall[3 * g + 1], p2 = igall[3 * g + 2];
  double re = 0, im = 0;
  for (int w = lane; w < npw2; w += 32) {
    const int a = G2[3 * w] + p0 - lo0, b = G2[3 * w + 1] + p1 - lo1, c = G2[3 * w + 2] + p2 - lo2;
    if (a < 0 || a >= d0 || b < 0 || b >= d1 || c < 0 || c >= d2) continue;
    const int j = boxmap[(a * d1 + b) * d2 + c];
    if (j < 0) continue;
    const float2 x = C1[j], y = C2[w];
    re += (double)y.x * x.x + (double)y.y * x.y;       // y * conj(x)
    im += (double)y.y * x.x - (double)y.x * x.y;
  }
  for (int o = 16; o; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  if (lane == 0) out[g] = make_double2(re, im);
}

// out[g] += pref * sum_s exp(sgn 2 pi i GP.R_s) sum_t W[s][t] * F_{elem(s), t}(G),  G = gsign * (GP + dk) in Cartesian
// One CTA per GP: the F terms of every element are evaluated once into shared memory, then threads run over sites.
__global__ void __launch_bounds__(128)
momentum_site_kernel(int numg, const int* __restrict__ igall, double dk0, double dk1, double dk2, double gsign,
                     const double* __restrict__ recl, int nelem, const MomElem* __restrict__ elems,
                     const MomTerm* __restrict__ terms, const double* __restrict__ radial,
                     const long* __restrict__ slot_off, int nsites, const int* __restrict__ site_elem,
                     const double* __restrict__ coords, const long* __restrict__ w_off,
                     const double2* __restrict__ W, double phase_sign, double pref, double2* __restrict__ out) {
  extern __shared__ double2 sF[];                 // all elements' terms
  __shared__ double red[2][4];
  const double PI = 3.14159265358979323846;
  const int g = blockIdx.x;
  if (g >= numg) return;
  const double q0 = igall[3 * g], q1 = igall[3 * g + 1], q2 = igall[3 * g + 2];
  double f[3] = {gsign * (q0 + dk0), gsign * (q1 + dk1), gsign * (q2 + dk2)};
  double v[3];
  for (int d = 0; d < 3; d++) v[d] = f[0] * recl[d] + f[1] * recl[3 + d] + f[2] * recl[6 + d];   // frac_to_cartesian
  const double r = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double theta, phi;
  d_angles(v, r, &theta, &phi);
  for (int e = 0; e < nelem; e++) {
    const MomElem me = elems[e];
    const double* ks = radial + me.ks_off;
    for (int t = threadIdx.x; t < me.nterm; t += blockDim.x) {
      const MomTerm mt = terms[me.term_off + t];
      double2 val = make_double2(0, 0);
      const int am = mt.M < 0 ? -mt.M : mt.M;
      if (am <= mt.L && (r != 0 || mt.special0)) {
        const double* fr = radial + slot_off[mt.slot];
        const double rad = d_eval_log(r, me.N, ks, fr, fr + me.N);
        const double2 y = d_ylm(mt.L, mt.M, theta, phi);
        val = make_double2(rad * y.x, rad * y.y);
      }
      sF[me.term_off + t] = val;
    }
  }
  __syncthreads();
  double re = 0, im = 0;
  for (int s = threadIdx.x; s < nsites; s += blockDim.x) {
    const MomElem me = elems[site_elem[s]];
    const double2* w = W + w_off[s];
    double ar = 0, ai = 0;
    for (int t = 0; t < me.nterm; t++) {
      const double2 a = w[t], b = sF[me.term_off + t];
      ar += a.x * b.x - a.y * b.y;
      ai += a.x * b.y + a.y * b.x;
    }
    double sn, cs;
    sincos(phase_sign * 2 * PI * (q0 * coords[3 * s] + q1 * coords[3 * s + 1] + q2 * coords[3 * s + 2]), &sn, &cs);
    re += ar * cs - ai * sn;
    im += ar * sn + ai * cs;
  }
  for (int o = 16; o; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = re; red[1][threadIdx.x >> 5] = im; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0, ti = 0;
    for (int wq = 0; wq < 4; wq++) { tr += red[0][wq]; ti += red[1][wq]; }
    out[g].x += pref * tr;
    out[g].y += pref * ti;
  }
}

// plane-wave part of fullwf_reciprocal [momentum.c:483-499]: out[g] += C[box(GP_g)] when GP is inside the box
__global__ void momentum_pick_kernel(int numg, const int* __restrict__ igall, const float2* __restrict__ C,
                                     const int* __restrict__ boxmap, int lo0, int lo1, int lo2, int d0, int d1,
                                     int d2, double2* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= numg) return;
  const int a = igall[3 * g] - lo0, b = igall[3 * g + 1] - lo1, c = igall[3 * g + 2] - lo2;
  if (a < 0 || a >= d0 || b < 0 || b >= d1 || c < 0 || c >= d2) return;
  const int j = boxmap[(a * d1 + b) * d2 + c];
  if (j < 0) return;
  out[g].x += (double)C[j].x;
  out[g].y += (double)C[j].y;
}

// (a13) density accumulation  [density.c:170-173, 193-196]:  rho[g] += sum_box w[box] |x_box[g]|^2
__global__ void __launch_bounds__(256)
density_accum_kernel(const double2* __restrict__ x, long ngrid, int nbox,
                     const double* __restrict__ w, double* __restrict__ rho) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < ngrid; g += stride) {
    double a = 0;
    for (int b = 0; b < nbox; b++) {
      const double2 v = x[(long)b * ngrid + g];
      a += (v.x * v.x + v.y * v.y) * w[b];
    }
    rho[g] += a;
  }
}

// (f4) brute-force real-space overlap  [density.c:205-230]: partial[b][chunk] = sum_{g in chunk} conj(xR[b][g]) x[g]
// Deterministic two-stage reduction (fixed chunking, tree inside the block, ordered sum over chunks).
__global__ void __launch_bounds__(256)
grid_dot_partial_kernel(const double2* __restrict__ xR, const double2* __restrict__ x, long ngrid, int nchunk,
                        double2* __restrict__ partial) {
  const int b = blockIdx.y, c = blockIdx.x;
  const long per = (ngrid + nchunk - 1) / nchunk;
  const long g0 = (long)c * per, g1 = min(ngrid, g0 + per);
  double re = 0, im = 0;
  for (long g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
    const double2 a = xR[(long)b * ngrid + g], v = x[g];
    re += a.x * v.x + a.y * v.y;
    im += a.x * v.y - a.y * v.x;
  }
  __shared__ double sr[256], si[256];
  sr[threadIdx.x] = re;
  si[threadIdx.x] = im;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sr[threadIdx.x] += sr[threadIdx.x + s];
      si[threadIdx.x] += si[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(long)b * nchunk + c] = make_double2(sr[0], si[0]);
}
// same partial sums on interleaved boxes xR[group][g][16]: block = (chunk, group), thread = (point, band)
__global__ void __launch_bounds__(256)
grid_dot_partial_il_kernel(const double2* __restrict__ xR, const double2* __restrict__ x, long ngrid, int nchunk,
                           double2* __restrict__ partial) {
  const int grp = blockIdx.y, c = blockIdx.x;
  const int b = threadIdx.x & 15, pt = threadIdx.x >> 4;
  const long per = (ngrid + nchunk - 1) / nchunk;
  const long g0 = (long)c * per, g1 = min(ngrid, g0 + per);
  double re = 0, im = 0;
  for (long g = g0 + pt; g < g1; g += 16) {
    const double2 a = xR[((long)grp * ngrid + g) * 16 + b], v = x[g];
    re += a.x * v.x + a.y * v.y;
    im += a.x * v.y - a.y * v.x;
  }
  __shared__ double sr[256], si[256];
  sr[threadIdx.x] = re;
  si[threadIdx.x] = im;
  __syncthreads();
  for (int s = 128; s >= 16; s >>= 1) {          // fixed tree over the 16 point slots of each band
    if (threadIdx.x < s) {
      sr[threadIdx.x] += sr[threadIdx.x + s];
      si[threadIdx.x] += si[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x < 16) partial[(long)(grp * 16 + threadIdx.x) * nchunk + c] = make_double2(sr[threadIdx.x], si[threadIdx.x]);
}
__global__ void grid_dot_final_kernel(const double2* __restrict__ partial, int nchunk, int nb, double scale,
                                      double2* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  double re = 0, im = 0;
  for (int c = 0; c < nchunk; c++) {
    re += partial[(long)b * nchunk + c].x;
    im += partial[(long)b * nchunk + c].y;
  }
  out[b] = make_double2(re * scale, im * scale);
}

}  // namespace pawb200
