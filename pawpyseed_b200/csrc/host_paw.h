// Host-side (C++) PAW setup for the B200 band-projection engine.
//
// Everything here is O(#elements) or O(#sites * sphere box) work that feeds the GPU path:
// cubic splines, the NumSBT low-pass filter of (phi - phi~), one-centre overlap matrices,
// the bit-exact sphere index lists, off-site partial-wave overlaps, and the WAVECAR parser.
// The arithmetic follows the reference C (file:line cited per function, relative to
// pawpyseed/core/) closely enough to agree to ~1e-14; the sphere-membership test is
// evaluated in the reference's exact operation order on the host because it is discontinuous.
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

namespace pawb200 {

using cdouble = std::complex<double>;

constexpr double kPi = 3.14159265358979323846;   // utils.c:12
constexpr double kPiDensity = 3.14159265359;     // density.c:13 (truncated in the reference)
constexpr double kC = 0.262465831;               // projector.c:16

// ---- vectors ---------------------------------------------------------------------------
double determinant3(const double* m);
void frac_to_cart(double* v, const double* lattice);            // in place, utils.c:149
void cart_to_frac(double* v, const double* reclattice);         // in place, utils.c:162
double vec_mag(const double* v);                                // pow(dot,0.5), utils.c:44
void min_image_path(const double* coord, const double* center, const double* lattice,
                    double* path, double* r);                   // utils.c:51-73
void reciprocal_lattice(const double* lattice, double* rec);    // reader.c:58-64

// ---- splines ---------------------------------------------------------------------------
struct Spline {                 // y + 3 coefficient rows on grid x (not owned)
  std::vector<double> c[3];
};
Spline make_spline(const double* x, const double* y, int n);               // utils.c:699-749
double spline_integrate(const double* x, const double* a, const Spline& s, int n);  // :751-764
double eval_linear_grid(double r, double rmax, int n, const double* x, const double* f,
                        const Spline& s);                                   // utils.c:461-475
double eval_log_grid(double r, int n, const double* x, const double* f, const Spline& s);  // :477-489

// ---- spherical harmonics ------------------------------------------------------------------
double assoc_legendre(int l, int m, double x);                              // utils.c:377-386
cdouble sph_harm(int l, int m, double theta, double phi);                   // utils.c:441-449
cdouble sph_harm_cos(int l, int m, double costheta, double phi);            // utils.c:451-459
double sph_bessel_rec(double x, int l);                                     // utils.c:807-827

// ---- NumSBT (Talman 2009), sbt.c:24-215 ------------------------------------------------
class BesselTransform {
 public:
  BesselTransform(double encut, double enbuf, int lmax, int n, const double* r);
  std::vector<double> forward(const double* f, int l) const;   // input r*R(r) -> g(k)
  std::vector<double> inverse(const double* g, int l) const;   // g(k) -> R(r)
  const std::vector<double>& kgrid() const { return kgrid_; }

 private:
  void dft_backward(std::vector<cdouble>& x) const;             // unnormalised e^{+i} DFT (Bluestein)
  int n2_;
  int lmax_;
  std::vector<double> ks_, rs_, kgrid_;
  std::vector<std::vector<cdouble>> mult_;
  std::vector<cdouble> twiddle_;
  int m2_ = 0;                                                  // power-of-two convolution length
  std::vector<cdouble> chirp_, chirp_fft_;                      // w[n] = exp(+i pi n^2 / N), FFT of its conjugate kernel
};

// ---- per-element PAW data ----------------------------------------------------------------
struct Channel { int n, l, m; };     // radial index, l, m - order of utils.c:617-632

struct RadialFunc {
  int l = 0;
  std::vector<double> proj, aewave, pswave, diffwave, kwave, smooth_diffwave;
  Spline proj_s, diffwave_s, kwave_s, smooth_s;
};

struct Element {
  int num_projs = 0, total_projs = 0, lmax = 0, proj_gridsize = 0, wave_gridsize = 0;
  double rmax = 0, wave_rmax = 0;
  std::vector<double> wave_grid, kwave_grid, proj_grid, smooth_grid;
  std::vector<RadialFunc> funcs;
  std::vector<double> aeov, psov, diov;    // num_projs x num_projs
  std::vector<Channel> chan;
};

// get_projector_list, projector.c:20-171 (+ make_pwave_overlap_matrices :507-558)
std::vector<Element> build_elements(int num_els, const int* labels, const int* ls,
                                    const double* wave_grids, const double* projectors,
                                    const double* aewaves, const double* pswaves,
                                    const double* rmaxs, double grid_encut);

// ---- sphere geometry (utils.c:636-671; density.c:262-296) ------------------------------------
struct SphereGeom {
  std::vector<int32_t> index;     // wrapped linear grid index, ascending (i,j,k) box order
  std::vector<int16_t> ijk;       // 4 per point: unwrapped grid coordinates (i, j, k, 0) - the device rebuilds the
                                  // Cartesian offset, the wrapped index and the cell shifts from them
};
// radius_test: points with |r| < radius_test are kept; box from rmax_box.
SphereGeom sphere_geometry(const double* coord, const double* lattice, const int* fftg,
                           double rmax_box, double radius_test);

// ---- off-site partial-wave overlap (radial.c:116-196, SBTFACS regenerated) ----------------
double wigner3j(int j1, int j2, int j3, int m1, int m2, int m3);
double sbt_factor(int l1, int l2, int L, int m1, int m2);       // table entry of gaunt.py:17-30
cdouble offsite_overlap_recip(const double* dcoord, const double* k1, const double* f1,
                              const Spline& s1, int size1, const double* k2, const double* f2,
                              const Spline& s2, int size2, int l1, int m1, int l2, int m2);

// ---- WAVECAR parsing (reader.c:55-315) ------------------------------------------------------
struct WavecarHeader {
  long nrecl = 0;
  int nspin = 0, nwk = 0, nband = 0;
  double encut = 0;
  double lattice[9], reclattice[9];
  double nbmax[3];
  int npmax = 0;        // reader.c:26-57 upper bound on the plane-wave count
};
struct KPointInfo {
  int nplane = 0;               // coefficients per band as stored (2*ng for noncollinear)
  double k[3];
  std::vector<int32_t> G;       // 3 per plane wave (one spinor half), file order
  std::vector<int32_t> perm;    // storage position j holds file coefficient perm[j] (box-index order)
  std::vector<int32_t> pos;     // inverse of perm
  std::vector<double> energy, occ;
};
// Enumerate G vectors for one k-point in file order (reader.c:230-271); updates G_bounds.
std::vector<int32_t> enumerate_g(const WavecarHeader& h, const double* k, int* G_bounds);
void wavecar_bounds(WavecarHeader& h);   // reader.c:55-127 `setup`

}  // namespace pawb200
