// Host-side PAW setup. See host_paw.h. Compiled with -ffp-contract=off: the sphere membership
// and plane-wave cutoff tests are discontinuous and must round exactly like the reference's
// plain-C (no FMA) build.
#include "host_paw.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <stdexcept>

namespace pawb200 {

// ---------------------------------------------------------------------------------------
// vectors
// ---------------------------------------------------------------------------------------
double determinant3(const double* m) {
  return m[0] * m[4] * m[8] + m[1] * m[5] * m[6] + m[2] * m[3] * m[7] - m[2] * m[4] * m[6] -
         m[1] * m[3] * m[8] - m[0] * m[5] * m[7];
}

void frac_to_cart(double* v, const double* L) {
  const double a = v[0], b = v[1], c = v[2];
  v[0] = a * L[0] + b * L[3] + c * L[6];
  v[1] = a * L[1] + b * L[4] + c * L[7];
  v[2] = a * L[2] + b * L[5] + c * L[8];
}

void cart_to_frac(double* v, const double* R) {
  const double a = v[0], b = v[1], c = v[2];
  v[0] = (a * R[0] + b * R[1] + c * R[2]) / 2 / kPi;
  v[1] = (a * R[3] + b * R[4] + c * R[5]) / 2 / kPi;
  v[2] = (a * R[6] + b * R[7] + c * R[8]) / 2 / kPi;
}

static inline double dot3(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline void cross3(double* r, const double* t, const double* b) {
  r[0] = t[1] * b[2] - t[2] * b[1];
  r[1] = t[2] * b[0] - t[0] * b[2];
  r[2] = t[0] * b[1] - t[1] * b[0];
}

double vec_mag(const double* v) { return std::pow(dot3(v, v), 0.5); }

void min_image_path(const double* coord, const double* center, const double* lattice,
                    double* path, double* r) {
  double best = INFINITY;
  for (int i = -1; i <= 1; i++)
    for (int j = -1; j <= 1; j++)
      for (int k = -1; k <= 1; k++) {
        double t[3] = {coord[0] + i - center[0], coord[1] + j - center[1],
                       coord[2] + k - center[2]};
        frac_to_cart(t, lattice);
        const double d = vec_mag(t);
        if (d < best) {
          best = d;
          path[0] = t[0];
          path[1] = t[1];
          path[2] = t[2];
        }
      }
  *r = best;
}

void reciprocal_lattice(const double* a, double* b) {
  cross3(b + 0, a + 3, a + 6);
  cross3(b + 3, a + 6, a + 0);
  cross3(b + 6, a + 0, a + 3);
  const double vol = determinant3(a);
  for (int i = 0; i < 9; i++) b[i] *= 2.0 * kPi / vol;
}

// ---------------------------------------------------------------------------------------
// splines (VASP SPLCOF recurrences as used by the reference)
// ---------------------------------------------------------------------------------------
Spline make_spline(const double* x, const double* y, int n) {
  Spline s;
  for (auto& v : s.c) v.assign(n, 0.0);
  std::vector<double>&b = s.c[0], &c = s.c[1], &d = s.c[2];
  const double slope0 = (y[1] - y[0]) / (x[1] - x[0]);
  if (slope0 > 0.99e30) {
    c[0] = 0;
    b[0] = 0;
  } else {
    c[0] = -0.5;
    b[0] = (3 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - slope0);
  }
  for (int i = 1; i < n - 1; i++) {
    const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    const double den = sig * c[i - 1] + 2;
    c[i] = (sig - 1) / den;
    b[i] = (6 * ((y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1])) /
                (x[i + 1] - x[i - 1]) -
            sig * b[i - 1]) /
           den;
  }
  b[n - 1] = c[n - 1] = d[n - 1] = 0;
  for (int i = n - 2; i >= 0; i--) c[i] = c[i] * c[i + 1] + b[i];
  for (int i = 0; i < n - 1; i++) {
    const double h = x[i + 1] - x[i];
    const double t = (c[i + 1] - c[i]) / 6;
    d[i] = t / h;
    c[i] = c[i] / 2;
    b[i] = (y[i + 1] - y[i]) / h - (c[i] + t) * h;
  }
  return s;
}

double spline_integrate(const double* x, const double* a, const Spline& s, int n) {
  double total = 0;
  for (int i = 0; i < n - 1; i++) {
    const double dx = x[i + 1] - x[i];
    total += dx * (a[i] + dx * (s.c[0][i] / 2 + dx * (s.c[1][i] / 3 + s.c[2][i] * dx / 4)));
  }
  return total;
}

double eval_linear_grid(double r, double rmax, int n, const double* x, const double* f,
                        const Spline& s) {
  if (r > x[n - 1]) return 0;
  if (r < x[0]) return f[0];
  const int i = std::min((int)(r / rmax * n), n - 2);
  const double t = r - x[i];
  return f[i] + t * (s.c[0][i] + t * (s.c[1][i] + t * s.c[2][i]));
}

double eval_log_grid(double r, int n, const double* x, const double* f, const Spline& s) {
  if (r > x[n - 1]) return 0;
  if (r < x[0]) return f[0];
  const int i = std::min((int)(std::log(r / x[0]) / std::log(x[1] / x[0])), n - 2);
  const double t = r - x[i];
  return f[i] + t * (s.c[0][i] + t * (s.c[1][i] + t * s.c[2][i]));
}

// ---------------------------------------------------------------------------------------
// spherical harmonics
// ---------------------------------------------------------------------------------------
static double ifac(int n) {  // integer factorial like utils.c:431-439 (l <= 3 keeps it < 2^31)
  int t = 1;
  for (int m = 1; m <= n; m++) t *= m;
  return (double)t;
}

double assoc_legendre(int l, int m, double x) {
  if (m < 0) return std::pow(-1.0, m) * ifac(l + m) / ifac(l - m) * assoc_legendre(l, -m, x);
  double total = 0;
  for (int n = l; n >= 0 && 2 * n - l - m >= 0; n--)
    total += std::pow(x, 2 * n - l - m) * ifac(2 * n) / ifac(2 * n - l - m) / ifac(n) /
             ifac(l - n) * std::pow(-1, l - n);
  return total * std::pow(-1, m) * std::pow(1 - x * x, m / 2.0) / std::pow(2, l);
}

cdouble sph_harm_cos(int l, int m, double ct, double phi) {
  const double norm = std::pow((2 * l + 1) / (4 * kPi) * ifac(l - m) / ifac(l + m), 0.5);
  return norm * assoc_legendre(l, m, ct) * std::exp(cdouble(0, m * phi));
}

cdouble sph_harm(int l, int m, double theta, double phi) {
  return sph_harm_cos(l, m, std::cos(theta), phi);
}

double sph_bessel_rec(double x, int l) {
  if (x < 10e-6) return l == 0 ? 1.0 : 0.0;
  double jm = std::sin(x) / x;
  double j = std::sin(x) / (x * x) - std::cos(x) / x;
  if (l == 0) return jm;
  if (l == 1) return j;
  double jp = 0;
  for (int q = 1; q < l; q++) {
    jp = (2 * q + 1) / x * j - jm;
    jm = j;
    j = jp;
  }
  return jp;
}

// ---------------------------------------------------------------------------------------
// NumSBT
// ---------------------------------------------------------------------------------------
static void fft_pow2(std::vector<cdouble>& a, int sign, const std::vector<cdouble>& roots);
BesselTransform::BesselTransform(double encut, double enbuf, int lmax, int n, const double* r) {
  n2_ = 2 * n;
  lmax_ = lmax == 0 ? 1 : lmax;
  const int N = n2_;
  const double drho = std::log(r[1] / r[0]);
  const double dt = 2 * kPi / N / drho;
  const double rmin = r[0];
  const double kmin = std::pow((encut + enbuf) * kC, 0.5) * std::exp(-(N / 2 - 1) * drho);
  const double kappamin = std::log(kmin);
  ks_.resize(N);
  rs_.resize(N);
  for (int i = 0; i < N; i++) {
    ks_[i] = kmin * std::exp(i * drho);
    rs_[i] = rmin * std::exp((i - N / 2) * drho);
  }
  kgrid_.assign(ks_.begin(), ks_.begin() + N / 2);
  const double rhomin = std::log(rs_[0]);
  mult_.assign(lmax_ + 1, std::vector<cdouble>(N));
  for (int i = 0; i < N; i++) {
    const double t = i * dt;
    const double rad = std::pow(10.5 * 10.5 + t * t, 0.5);
    const double phi3 = (kappamin + rhomin) * t;
    double phi = std::atan((2 * t) / 21);
    double phi1 = -10 * phi - t * std::log(rad) + t + std::sin(phi) / (12 * rad) -
                  std::sin(3 * phi) / (360 * std::pow(rad, 3)) +
                  std::sin(5 * phi) / (1260 * std::pow(rad, 5)) -
                  std::sin(7 * phi) / (1680 * std::pow(rad, 7));
    for (int j = 1; j <= 10; j++) phi1 += std::atan((2 * t) / (2 * j - 1));
    const double phi2 = -std::atan(std::tanh(kPi * t / 2));
    phi = phi1 + phi2 + phi3;
    mult_[0][i] = std::pow(kPi / 2, 0.5) * std::exp(cdouble(0, phi)) / (double)N;
    if (i == 0) mult_[0][i] = 0.5 * mult_[0][i];
    phi = -phi2 - std::atan(2 * t);
    mult_[1][i] = std::exp(cdouble(0, 2 * phi)) * mult_[0][i];
    for (int l = 1; l < lmax_; l++) {
      phi = -std::atan(2 * t / (2 * l + 1));
      mult_[l + 1][i] = std::exp(cdouble(0, 2 * phi)) * mult_[l - 1][i];
    }
  }
  // chirp w[n] = exp(+i pi n^2 / N): reduce n^2 mod 2N in integers so the argument stays accurate
  chirp_.resize(N);
  for (long n = 0; n < N; n++) {
    const long q = (n * n) % (2L * N);
    const long double a = 3.141592653589793238462643383279502884L * q / N;
    chirp_[n] = cdouble((double)cosl(a), (double)sinl(a));
  }
  m2_ = 1;
  while (m2_ < 2 * N - 1) m2_ <<= 1;
  twiddle_.resize(m2_ / 2);   // roots of unity of the power-of-two transforms
  for (int k = 0; k < m2_ / 2; k++) {
    const long double a = 2.0L * 3.141592653589793238462643383279502884L * k / m2_;
    twiddle_[k] = cdouble((double)cosl(a), (double)sinl(a));
  }
  chirp_fft_.assign(m2_, cdouble(0, 0));
  chirp_fft_[0] = std::conj(chirp_[0]);
  for (int n = 1; n < N; n++) chirp_fft_[n] = chirp_fft_[m2_ - n] = std::conj(chirp_[n]);
  fft_pow2(chirp_fft_, +1, twiddle_);
}

// In-place radix-2 FFT of length m (power of two); sign = +1 -> e^{+i}, unnormalised.
// roots[k] = exp(+2 pi i k / m), k < m/2.
static void fft_pow2(std::vector<cdouble>& a, int sign, const std::vector<cdouble>& roots) {
  const int m = (int)a.size();
  for (int i = 1, j = 0; i < m; i++) {
    int bit = m >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (int len = 2; len <= m; len <<= 1) {
    const int step = m / len;
    for (int i = 0; i < m; i += len)
      for (int k = 0; k < len / 2; k++) {
        const cdouble w = sign > 0 ? roots[k * step] : std::conj(roots[k * step]);
        const cdouble u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
  }
}

// Length-2N (~650, factors 2*17*19) DFT through Bluestein's chirp-z identity
//   X[k] = w[k] * sum_n (x[n] w[n]) * conj(w)[k-n],  w[n] = exp(+i pi n^2 / N)
// evaluated as a power-of-two circular convolution: O(N log N) instead of the O(N^2) direct sum.
void BesselTransform::dft_backward(std::vector<cdouble>& x) const {
  const int N = n2_, M = m2_;
  std::vector<cdouble> a(M, cdouble(0, 0));
  for (int n = 0; n < N; n++) a[n] = x[n] * chirp_[n];
  fft_pow2(a, +1, twiddle_);
  for (int i = 0; i < M; i++) a[i] *= chirp_fft_[i];
  fft_pow2(a, -1, twiddle_);
  const double inv = 1.0 / M;
  for (int k = 0; k < N; k++) x[k] = a[k] * inv * chirp_[k];
}

std::vector<double> BesselTransform::forward(const double* f, int l) const {
  const int N = n2_, h = N / 2;
  std::vector<double> fs(N);
  const double c0 = f[0] / std::pow(rs_[h], l + 1);
  for (int i = 0; i < h; i++) fs[i] = c0 * std::pow(rs_[i], l + 1);
  for (int i = h; i < N; i++) fs[i] = f[i - h];
  std::vector<cdouble> x(N);
  for (int m = 0; m < N; m++) x[m] = std::pow(rs_[m], 0.5) * fs[m];
  dft_backward(x);
  for (int n = 0; n < N; n++) x[n] = n >= h ? cdouble(0, 0) : x[n] * mult_[l][n];
  dft_backward(x);
  std::vector<double> out(h);
  for (int p = 0; p < h; p++) out[p] = x[p].real() * (2 / std::pow(ks_[p], 1.5));
  return out;
}

std::vector<double> BesselTransform::inverse(const double* g, int l) const {
  const int N = n2_, h = N / 2;
  const std::vector<double>&kk = rs_, &rr = ks_;   // roles swap, sbt.c:172-173
  std::vector<cdouble> x(N);
  for (int m = 0; m < N; m++) x[m] = std::pow(rr[m], 1.5) * (m < h ? g[m] : 0.0);
  dft_backward(x);
  for (int n = 0; n < N; n++) x[n] = n >= h ? cdouble(0, 0) : x[n] * mult_[l][n];
  dft_backward(x);
  std::vector<double> out(h);
  for (int p = 0; p < h; p++) {
    double v = x[p + h].real() / kPi * 2;
    v *= 2 / std::pow(kk[p + h], 1.5);
    out[p] = v;
  }
  return out;
}

// ---------------------------------------------------------------------------------------
// per-element setup
// ---------------------------------------------------------------------------------------
static void overlap_matrices(Element& e) {
  const int P = e.num_projs, W = e.wave_gridsize;
  e.aeov.assign(P * P, 0.0);
  e.psov.assign(P * P, 0.0);
  e.diov.assign(P * P, 0.0);
  std::vector<double> prod(W);
  auto integral = [&](auto&& fn) {
    for (int k = 0; k < W; k++) prod[k] = fn(k);
    Spline s = make_spline(e.wave_grid.data(), prod.data(), W);
    return spline_integrate(e.wave_grid.data(), prod.data(), s, W);
  };
  for (int i = 0; i < P; i++)
    for (int j = i; j < P; j++) {
      if (e.funcs[i].l != e.funcs[j].l) continue;
      const auto &a = e.funcs[i], &b = e.funcs[j];
      const double ps = integral([&](int k) { return a.pswave[k] * b.pswave[k]; });
      const double ae = integral([&](int k) { return a.aewave[k] * b.aewave[k]; });
      const double di = integral(
          [&](int k) { return (a.aewave[k] - a.pswave[k]) * (b.aewave[k] - b.pswave[k]); });
      e.psov[i * P + j] = e.psov[j * P + i] = ps;
      e.aeov[i * P + j] = e.aeov[j * P + i] = ae;
      e.diov[i * P + j] = e.diov[j * P + i] = di;
    }
}

std::vector<Element> build_elements(int num_els, const int* labels, const int* ls,
                                    const double* wave_grids, const double* projectors,
                                    const double* aewaves, const double* pswaves,
                                    const double* rmaxs, double grid_encut) {
  std::vector<Element> out(num_els);
  long wpos = 0, ppos = 0, gpos = 0;
  int lpos = 0;
  const double cutoff_k = std::pow(kC * grid_encut, 0.5);
  for (int e = 0; e < num_els; e++) {
    Element& el = out[e];
    el.num_projs = labels[4 * e + 1];
    el.proj_gridsize = labels[4 * e + 2];
    el.wave_gridsize = labels[4 * e + 3];
    el.rmax = rmaxs[e];
    const int PG = el.proj_gridsize, WG = el.wave_gridsize;
    if (PG < 3 || WG < 3) throw std::runtime_error("radial grids need at least 3 points");
    el.wave_grid.assign(wave_grids + gpos, wave_grids + gpos + WG);
    gpos += WG;
    el.proj_grid.resize(PG);
    for (int j = 0; j < PG; j++) el.proj_grid[j] = el.rmax / PG * j;
    // the partial-wave grid regenerated by repeated multiplication (projector.c:65-70)
    std::vector<double> regrid(WG);
    regrid[0] = el.wave_grid[0];
    const double ratio = std::pow(el.wave_grid[1] / el.wave_grid[0], 1.0);
    for (int p = 1; p < WG; p++) regrid[p] = regrid[p - 1] * ratio;
    el.wave_rmax = el.wave_grid[WG - 1];
    el.smooth_grid.resize(PG);
    for (int j = 0; j < PG; j++) el.smooth_grid[j] = el.wave_rmax / PG * j;

    el.funcs.resize(el.num_projs);
    for (int k = 0; k < el.num_projs; k++) {
      RadialFunc& f = el.funcs[k];
      f.l = ls[lpos++];
      el.lmax = std::max(el.lmax, f.l);
      el.total_projs += 2 * f.l + 1;
      f.aewave.assign(aewaves + wpos, aewaves + wpos + WG);
      f.pswave.assign(pswaves + wpos, pswaves + wpos + WG);
      wpos += WG;
      f.diffwave.resize(WG);
      for (int j = 0; j < WG; j++) f.diffwave[j] = f.aewave[j] - f.pswave[j];
      f.proj.assign(projectors + ppos, projectors + ppos + PG);
      ppos += PG;
      f.proj_s = make_spline(el.proj_grid.data(), f.proj.data(), PG);
      f.diffwave_s = make_spline(el.wave_grid.data(), f.diffwave.data(), WG);
      for (int m = -f.l; m <= f.l; m++) el.chan.push_back({k, f.l, m});
    }
    if (el.lmax > 3) throw std::runtime_error("l > 3 is not supported (as in the reference)");

    BesselTransform sbt(1e7, 0, el.lmax, WG, el.wave_grid.data());
    el.kwave_grid = sbt.kgrid();
#pragma omp parallel for schedule(dynamic)
    for (int fi = 0; fi < el.num_projs; fi++) {
      RadialFunc& f = el.funcs[fi];
      f.kwave = sbt.forward(f.diffwave.data(), f.l);
      f.kwave_s = make_spline(el.kwave_grid.data(), f.kwave.data(), WG);
    }
#pragma omp parallel for schedule(dynamic)
    for (int fi = 0; fi < el.num_projs; fi++) {
      RadialFunc& f = el.funcs[fi];
      std::vector<double> lowpass(WG, 0.0);
      for (int q = 0; q < WG && el.kwave_grid[q] < cutoff_k; q++) lowpass[q] = f.kwave[q];
      std::vector<double> smooth = sbt.inverse(lowpass.data(), f.l);
      Spline ss = make_spline(regrid.data(), smooth.data(), WG);
      f.smooth_diffwave.assign(PG, 0.0);
      for (int p = 1; p < PG; p++)
        f.smooth_diffwave[p] = eval_log_grid(el.smooth_grid[p], WG, regrid.data(), smooth.data(), ss);
      f.smooth_diffwave[0] = f.l > 0 ? 0.0 : f.smooth_diffwave[1];
      f.smooth_s = make_spline(el.smooth_grid.data(), f.smooth_diffwave.data(), PG);
    }
    overlap_matrices(el);
  }
  return out;
}

// ---------------------------------------------------------------------------------------
// sphere geometry
// ---------------------------------------------------------------------------------------
SphereGeom sphere_geometry(const double* coord, const double* L, const int* fftg,
                           double rmax_box, double radius_test) {
  SphereGeom g;
  const double vol = determinant3(L);
  double res[3];
  int half[3];
  cross3(res, L + 3, L + 6);
  half[0] = (int)(vec_mag(res) * rmax_box / vol * fftg[0]) + 1;
  cross3(res, L + 0, L + 6);
  half[1] = (int)(vec_mag(res) * rmax_box / vol * fftg[1]) + 1;
  cross3(res, L + 0, L + 3);
  half[2] = (int)(vec_mag(res) * rmax_box / vol * fftg[2]) + 1;
  int cen[3];
  for (int d = 0; d < 3; d++) cen[d] = (int)std::round(coord[d] * fftg[d]);
  const int N0 = fftg[0], N1 = fftg[1], N2 = fftg[2];
  const double r2_lo = radius_test * radius_test * (1 - 1e-12), r2_hi = radius_test * radius_test * (1 + 1e-12);
  {
    const double dv = std::fabs(vol) / ((double)N0 * N1 * N2);
    const size_t guess = (size_t)(4.19 * radius_test * radius_test * radius_test / dv * 1.15) + 64;
    g.index.reserve(guess);
    g.ijk.reserve(4 * guess);
  }
  // per-axis fractional offsets of the box (exactly (double)i / N - coord like utils.c:656-658)
  std::vector<double> tz(2 * half[2] + 1);
  std::vector<int> kz(2 * half[2] + 1);
  for (int q = 0; q <= 2 * half[2]; q++) {
    const int k = -half[2] + cen[2] + q;
    tz[q] = (double)k / N2 - coord[2];
    kz[q] = (k % N2 + N2) % N2;
  }
  for (int i = -half[0] + cen[0]; i <= half[0] + cen[0]; i++) {
    const double t0 = (double)i / N0 - coord[0];
    const int ii = (i % N0 + N0) % N0;
    for (int j = -half[1] + cen[1]; j <= half[1] + cen[1]; j++) {
      const double t1 = (double)j / N1 - coord[1];
      const int jj = (j % N1 + N1) % N1;
      // frac_to_cartesian evaluates (t0*L0 + t1*L3) + t2*L6 left to right: the first two terms are hoisted
      const double p0 = t0 * L[0] + t1 * L[3], p1 = t0 * L[1] + t1 * L[4], p2 = t0 * L[2] + t1 * L[5];
      const int rowbase = ii * N1 * N2 + jj * N2;
      for (int q = 0; q <= 2 * half[2]; q++) {
        const double x = p0 + tz[q] * L[6], y = p1 + tz[q] * L[7], z = p2 + tz[q] * L[8];
        // The reference tests pow(dot, 0.5) < R0.  Away from the surface the comparison of the squares
        // decides identically; only within a relative 1e-12 shell is the exact libm expression evaluated,
        // so the index lists stay bit-identical while pow() is skipped for all but a handful of points.
        const double d2 = x * x + y * y + z * z;
        bool inside;
        if (d2 < r2_lo)
          inside = true;
        else if (d2 > r2_hi)
          inside = false;
        else
          inside = std::pow(d2, 0.5) < radius_test;
        if (inside) {
          g.index.push_back(rowbase + kz[q]);
          g.ijk.push_back((int16_t)i);
          g.ijk.push_back((int16_t)j);
          g.ijk.push_back((int16_t)(-half[2] + cen[2] + q));
          g.ijk.push_back(0);
        }
      }
    }
  }
  return g;
}

// ---------------------------------------------------------------------------------------
// off-site overlap
// ---------------------------------------------------------------------------------------
static double lfac(int n) {
  double t = 1;
  for (int m = 2; m <= n; m++) t *= m;
  return t;
}

double wigner3j(int j1, int j2, int j3, int m1, int m2, int m3) {
  // Racah's single-sum formula; integer j only, j <= 6 here.
  if (m1 + m2 + m3 != 0) return 0;
  if (j3 < std::abs(j1 - j2) || j3 > j1 + j2) return 0;
  if (std::abs(m1) > j1 || std::abs(m2) > j2 || std::abs(m3) > j3) return 0;
  const double tri = lfac(j1 + j2 - j3) * lfac(j1 - j2 + j3) * lfac(-j1 + j2 + j3) /
                     lfac(j1 + j2 + j3 + 1);
  const double pre = std::sqrt(tri * lfac(j1 + m1) * lfac(j1 - m1) * lfac(j2 + m2) *
                               lfac(j2 - m2) * lfac(j3 + m3) * lfac(j3 - m3));
  const int tmin = std::max({0, j2 - j3 - m1, j1 - j3 + m2});
  const int tmax = std::min({j1 + j2 - j3, j1 - m1, j2 + m2});
  double sum = 0;
  for (int t = tmin; t <= tmax; t++) {
    const double den = lfac(t) * lfac(j3 - j2 + t + m1) * lfac(j3 - j1 + t - m2) *
                       lfac(j1 + j2 - j3 - t) * lfac(j1 - t - m1) * lfac(j2 - t + m2);
    sum += ((t & 1) ? -1.0 : 1.0) / den;
  }
  const int ph = j1 - j2 - m3;
  return ((ph & 1) ? -1.0 : 1.0) * pre * sum;
}

double sbt_factor(int l1, int l2, int L, int m1, int m2) {
  // gaunt.py:22-28: 3j(l1 l2 L;000) 3j(l1 l2 L; -m1 m2 m1-m2) sqrt((2l1+1)(2l2+1)(2L+1)/4pi)
  return wigner3j(l1, l2, L, 0, 0, 0) * wigner3j(l1, l2, L, -m1, m2, m1 - m2) *
         std::sqrt((double)((2 * l1 + 1) * (2 * l2 + 1) * (2 * L + 1)) / 4 / kPi);
}

cdouble offsite_overlap_recip(const double* dcoord, const double* k1, const double* f1,
                              const Spline& s1, int size1, const double* k2, const double* f2,
                              const Spline& s2, int size2, int l1, int m1, int l2, int m2) {
  constexpr int NK = 500;  // radial.c:11
  int lx = l1, ly = l2, mx = m1, my = m2;
  if (l1 < l2) {
    lx = l2; ly = l1; mx = m2; my = m1;
  }
  if (my < 0) {
    mx = -mx;
    my = -my;
  }
  const double kmax = std::min(k1[size1 - 1], k2[size2 - 1]);
  const double kmin = std::max(k1[0], k2[0]);
  double theta = 0, phi = 0;
  double R = vec_mag(dcoord);
  if (R < 10e-12) {
    R = 0;
  } else {
    theta = std::acos(dcoord[2] / R);
    if (R - std::fabs(dcoord[2]) < 10e-12)
      phi = 0;
    else
      phi = std::acos(dcoord[0] / std::pow(dcoord[0] * dcoord[0] + dcoord[1] * dcoord[1], 0.5));
    if (dcoord[1] < 0) phi = 2 * kPi - phi;
  }
  std::vector<double> kg(NK), base(NK), fn(NK);
  for (int q = 0; q < NK; q++) {
    kg[q] = kmin * std::pow(kmax / kmin, (double)q / NK);
    base[q] = eval_log_grid(kg[q], size1, k1, f1, s1) * eval_log_grid(kg[q], size2, k2, f2, s2) *
              kg[q] * kg[q];
  }
  cdouble total = 0;
  const double mult = std::pow(-1, m1) * 8;
  for (int L = std::abs(l1 - l2); L <= l1 + l2; L += 2) {
    for (int q = 0; q < NK; q++) fn[q] = base[q] * sph_bessel_rec(kg[q] * R, L);
    Spline sp = make_spline(kg.data(), fn.data(), NK);
    const double integ = spline_integrate(kg.data(), fn.data(), sp, NK);
    if (R > 10e-10) {
      if (std::abs(m1 - m2) > L) continue;  // Y_L^{m1-m2} vanishes identically
      cdouble ipow;
      switch (((l2 + L - l1) % 4 + 4) % 4) {
        case 0: ipow = cdouble(1, 0); break;
        case 1: ipow = cdouble(0, 1); break;
        case 2: ipow = cdouble(-1, 0); break;
        default: ipow = cdouble(0, -1); break;
      }
      total += integ * sbt_factor(lx, ly, L, mx, my) * sph_harm(L, m1 - m2, theta, phi) * ipow * mult;
    } else if (L == 0 && l1 == l2 && m1 == m2) {
      total += integ * 2 / kPi;
    }
  }
  return total;
}

// ---------------------------------------------------------------------------------------
// WAVECAR
// ---------------------------------------------------------------------------------------
void wavecar_bounds(WavecarHeader& h) {
  reciprocal_lattice(h.lattice, h.reclattice);
  const double* b = h.reclattice;
  const double mb[3] = {vec_mag(b), vec_mag(b + 3), vec_mag(b + 6)};
  const double g = std::pow(h.encut * kC, 0.5);
  double nb[3][3];
  auto combo = [&](int i, int j, int k, double* out) {
    const double ang = std::acos(dot3(b + 3 * i, b + 3 * j) / (mb[i] * mb[j]));
    double v[3];
    cross3(v, b + 3 * i, b + 3 * j);
    const double s3 = dot3(b + 3 * k, v) / (vec_mag(v) * mb[k]);
    out[i] = g / (mb[i] * std::fabs(std::sin(ang))) + 1;
    out[j] = g / (mb[j] * std::fabs(std::sin(ang))) + 1;
    out[k] = g / (mb[k] * std::fabs(s3)) + 1;
  };
  combo(0, 1, 2, nb[0]);
  combo(0, 2, 1, nb[1]);
  combo(2, 1, 0, nb[2]);
  for (int d = 0; d < 3; d++) h.nbmax[d] = std::fmax(nb[0][d], std::fmax(nb[1][d], nb[2][d]));
  int np = 0;
  for (int c = 0; c < 3; c++) {
    const int v = (int)std::round(4.0 / 3.0 * kPi * nb[c][0] * nb[c][1] * nb[c][2]);
    if (c == 0 || v < np) np = v;
  }
  h.npmax = np;
}

std::vector<int32_t> enumerate_g(const WavecarHeader& h, const double* k, int* Gb) {
  const double *b1 = h.reclattice, *b2 = h.reclattice + 3, *b3 = h.reclattice + 6;
  const double n1 = h.nbmax[0], n2 = h.nbmax[1], n3 = h.nbmax[2];
  std::vector<int> planes;
  for (int ig3 = 0; ig3 <= 2 * n3; ig3++) planes.push_back(ig3);
  std::vector<std::vector<int32_t>> per_plane(planes.size());
#pragma omp parallel for schedule(dynamic)
  for (long pi = 0; pi < (long)planes.size(); pi++) {
    const int ig3 = planes[pi];
    int ig3p = ig3;
    if (ig3 > n3) ig3p = ig3 - 2 * n3 - 1;
    auto& out = per_plane[pi];
    for (int ig2 = 0; ig2 <= 2 * n2; ig2++) {
      int ig2p = ig2;
      if (ig2 > n2) ig2p = ig2 - 2 * n2 - 1;
      for (int ig1 = 0; ig1 <= 2 * n1; ig1++) {
        int ig1p = ig1;
        if (ig1 > n1) ig1p = ig1 - 2 * n1 - 1;
        double s[3];
        for (int j = 0; j < 3; j++)
          s[j] = (k[0] + ig1p) * b1[j] + (k[1] + ig2p) * b2[j] + (k[2] + ig3p) * b3[j];
        const double gt = vec_mag(s);
        const double et = std::pow(gt, 2.0) / kC;
        if (et <= h.encut) {
          out.push_back(ig1p);
          out.push_back(ig2p);
          out.push_back(ig3p);
        }
      }
    }
  }
  std::vector<int32_t> all;
  for (auto& v : per_plane) all.insert(all.end(), v.begin(), v.end());
  if (Gb) {
    for (size_t w = 0; w < all.size() / 3; w++)
      for (int d = 0; d < 3; d++) {
        Gb[2 * d] = std::min(Gb[2 * d], (int)all[3 * w + d]);
        Gb[2 * d + 1] = std::max(Gb[2 * d + 1], (int)all[3 * w + d]);
      }
  }
  return all;
}

}  // namespace pawb200
