"""CPU restatement (numpy, FP64) of the reference's PAW band-projection path.

TEST INFRASTRUCTURE ONLY - this is the *oracle*: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product
(pawpyseed_b200/) never does.

Every function cites the reference file:line (relative to /root/reference/pawpyseed/core)
whose arithmetic it restates.  Pinning: tests/test_oracle_vs_ref.py checks this module live
against oracle/_ref/libpawpy_ref.so (the unmodified reference C built by oracle/Makefile)
and tests/test_oracle_golden.py against the committed fixtures under tests/golden/ generated
from that library (tests/golden/make_golden.py).

One deliberate divergence, stated in DESIGN.md: `pseudoprojection` here accumulates
in FP64; the reference accumulates and returns single precision (pseudoprojector.c:86).
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np
from scipy.special import sph_harm_y

PI = 3.14159265358979323846          # utils.c:12, projector.c:17, sbt.c:14, radial.c:10
PI_DENSITY = 3.14159265359           # density.c:13 (truncated constant, kept for fidelity)
CCONST = 0.262465831                 # projector.c:16, reader.c:11, sbt.c:13
KGRID_SIZE = 500                     # radial.c:11


# --------------------------------------------------------------------------- #
# small vector helpers (utils.c:34-49, 149-173)
# --------------------------------------------------------------------------- #
def determinant(m):
    m = np.asarray(m, dtype=np.float64).reshape(9)
    return (m[0] * m[4] * m[8] + m[1] * m[5] * m[6] + m[2] * m[3] * m[7]
            - m[2] * m[4] * m[6] - m[1] * m[3] * m[8] - m[0] * m[5] * m[7])


def frac_to_cartesian(frac, lattice):
    """utils.c:149-160, vectorised over leading axes; same operation order."""
    L = np.asarray(lattice, dtype=np.float64).reshape(9)
    f = np.asarray(frac, dtype=np.float64)
    x = f[..., 0] * L[0] + f[..., 1] * L[3] + f[..., 2] * L[6]
    y = f[..., 0] * L[1] + f[..., 1] * L[4] + f[..., 2] * L[7]
    z = f[..., 0] * L[2] + f[..., 1] * L[5] + f[..., 2] * L[8]
    return np.stack([x, y, z], axis=-1)


def mag(v):
    """utils.c:44: pow(dot(x,x), 0.5) - libm pow, not sqrt."""
    v = np.asarray(v, dtype=np.float64)
    return np.power(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1] + v[..., 2] * v[..., 2], 0.5)


def reciprocal_lattice(lattice):
    """reader.c:58-64."""
    a = np.asarray(lattice, dtype=np.float64).reshape(3, 3)
    vol = determinant(a)
    b = np.stack([np.cross(a[1], a[2]), np.cross(a[2], a[0]), np.cross(a[0], a[1])])
    return b * (2.0 * PI / vol)


def min_cart_path(coord, center, lattice):
    """utils.c:51-73, vectorised over coord[...,3]; first strict minimum over the
    27 images in the loop order i,j,k = -1..1."""
    coord = np.asarray(coord, dtype=np.float64)
    center = np.asarray(center, dtype=np.float64)
    best_r = np.full(coord.shape[:-1], np.inf)
    best = np.zeros(coord.shape)
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for k in (-1, 0, 1):
                t = np.stack([coord[..., 0] + i - center[0], coord[..., 1] + j - center[1],
                              coord[..., 2] + k - center[2]], axis=-1)
                c = frac_to_cartesian(t, lattice)
                r = mag(c)
                upd = r < best_r
                best[upd] = c[upd]
                best_r = np.where(upd, r, best_r)
    return best, best_r


# --------------------------------------------------------------------------- #
# splines (utils.c:461-489, 699-764)
# --------------------------------------------------------------------------- #
def spline_coeff(x, y):
    """utils.c:699-749 (VASP SPLCOF). Returns (3, N) array."""
    x = [float(v) for v in x]
    y = [float(v) for v in y]
    N = len(x)
    c0, c1, c2 = [0.0] * N, [0.0] * N, [0.0] * N
    d1p1 = (y[1] - y[0]) / (x[1] - x[0])
    if d1p1 > 0.99e30:
        c1[0] = 0.0
        c0[0] = 0.0
    else:
        c1[0] = -0.5
        c0[0] = (3 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - d1p1)
    for i in range(1, N - 1):
        s = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1])
        r = s * c1[i - 1] + 2
        c1[i] = (s - 1) / r
        c0[i] = (6 * ((y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]))
                 / (x[i + 1] - x[i - 1]) - s * c0[i - 1]) / r
    c0[N - 1] = c1[N - 1] = c2[N - 1] = 0.0
    for i in range(N - 2, -1, -1):
        c1[i] = c1[i] * c1[i + 1] + c0[i]
    for i in range(N - 1):
        s = x[i + 1] - x[i]
        r = (c1[i + 1] - c1[i]) / 6
        c2[i] = r / s
        c1[i] = c1[i] / 2
        c0[i] = (y[i + 1] - y[i]) / s - (c1[i] + r) * s
    return np.array([c0, c1, c2])


def spline_integral(x, a, s):
    """utils.c:751-764."""
    x = np.asarray(x); a = np.asarray(a)
    dx = x[1:] - x[:-1]
    b, c, d = s[0][:-1], s[1][:-1], s[2][:-1]
    terms = dx * (a[:-1] + dx * (b / 2 + dx * (c / 3 + d * dx / 4)))
    tot = 0.0
    for t in terms:           # sequential accumulation like the C loop
        tot += float(t)
    return tot


def proj_interpolate(r, rmax, x, f, s):
    """utils.c:461-475, vectorised over r."""
    r = np.asarray(r, dtype=np.float64)
    size = len(x)
    ind = np.minimum((r / rmax * size).astype(np.int64), size - 2)
    ind = np.clip(ind, 0, size - 2)
    rem = r - x[ind]
    val = f[ind] + rem * (s[0][ind] + rem * (s[1][ind] + rem * s[2][ind]))
    val = np.where(r < x[0], f[0], val)
    val = np.where(r > x[size - 1], 0.0, val)
    return val


def wave_interpolate(r, x, f, s):
    """utils.c:477-489, vectorised over r."""
    r = np.asarray(r, dtype=np.float64)
    size = len(x)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.log(r / x[0]) / math.log(x[1] / x[0])
    q = np.where(np.isfinite(q), q, 0.0)
    ind = np.clip(np.minimum(q.astype(np.int64), size - 2), 0, size - 2)
    rem = r - x[ind]
    val = f[ind] + rem * (s[0][ind] + rem * (s[1][ind] + rem * s[2][ind]))
    val = np.where(r < x[0], f[0], val)
    val = np.where(r > x[size - 1], 0.0, val)
    return val


# --------------------------------------------------------------------------- #
# spherical harmonics (utils.c:377-449, 547-564)
# --------------------------------------------------------------------------- #
def fac(n):
    """utils.c:431-439 (int factorial)."""
    t = 1
    for m in range(1, n + 1):
        t *= m
    return float(t)


def legendre(l, m, x):
    """utils.c:377-386, vectorised over x."""
    x = np.asarray(x, dtype=np.float64)
    if m < 0:
        return (-1.0) ** m * fac(l + m) / fac(l - m) * legendre(l, -m, x)
    total = np.zeros_like(x)
    n = l
    while n >= 0 and 2 * n - l - m >= 0:
        total = total + np.power(x, 2 * n - l - m) * fac(2 * n) / fac(2 * n - l - m) / fac(n) \
            / fac(l - n) * (-1.0) ** (l - n)
        n -= 1
    return total * (-1.0) ** m * np.power(1 - x * x, m / 2.0) / 2.0 ** l


def Ylm(l, m, theta, phi):
    """utils.c:441-449 - complex Y_lm with Condon-Shortley phase."""
    pref = math.pow((2 * l + 1) / (4 * PI) * fac(l - m) / fac(l + m), 0.5)
    return pref * legendre(l, m, np.cos(theta)) * np.exp(1j * m * np.asarray(phi))


def angles(vec, r):
    """utils.c:554-562 / 535-542: (theta, phi) of Cartesian vectors (r>0 assumed)."""
    vec = np.asarray(vec, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        theta = np.arccos(vec[..., 2] / r)
        rho = np.power(vec[..., 0] * vec[..., 0] + vec[..., 1] * vec[..., 1], 0.5)
        phi = np.arccos(vec[..., 0] / rho)
    phi = np.where(r - np.abs(vec[..., 2]) == 0, 0.0, phi)
    phi = np.where(vec[..., 1] < 0, 2 * PI - phi, phi)
    return theta, phi


def radial_times_ylm(radial, l, m, vec, r):
    """proj_value_helper tail (utils.c:552-563) / wave_value2 tail (utils.c:533-544)."""
    theta, phi = angles(vec, np.where(r == 0, 1.0, r))
    theta = np.where(r == 0, 0.0, theta)
    phi = np.where(r == 0, 0.0, phi)
    return radial * Ylm(l, m, theta, phi)


# --------------------------------------------------------------------------- #
# NumSBT (sbt.c:24-215); DftiComputeBackward == unnormalised e^{+i} DFT
# --------------------------------------------------------------------------- #
def _backward(x):
    return np.fft.ifft(x) * len(x)


class SBT:
    def __init__(self, encut, enbuf, lmax, r):
        """sbt.c:24-100."""
        r = np.asarray(r, dtype=np.float64)
        N = 2 * len(r)
        if lmax == 0:
            lmax = 1
        drho = math.log(r[1] / r[0])
        dt = 2 * PI / N / drho
        rmin = r[0]
        kmin = math.pow((encut + enbuf) * CCONST, 0.5) * math.exp(-(N // 2 - 1) * drho)
        kappamin = math.log(kmin)
        i = np.arange(N)
        self.ks = kmin * np.exp(i * drho)
        self.rs = rmin * np.exp((i - N // 2) * drho)
        rhomin = math.log(self.rs[0])
        t = i * dt
        rad = np.power(10.5 * 10.5 + t * t, 0.5)
        phi3 = (kappamin + rhomin) * t
        phi = np.arctan((2 * t) / 21)
        phi1 = (-10 * phi - t * np.log(rad) + t + np.sin(phi) / (12 * rad)
                - np.sin(3 * phi) / (360 * rad ** 3) + np.sin(5 * phi) / (1260 * rad ** 5)
                - np.sin(7 * phi) / (1680 * rad ** 7))
        for j in range(1, 11):
            phi1 = phi1 + np.arctan((2 * t) / (2 * j - 1))
        phi2 = -np.arctan(np.tanh(PI * t / 2))
        M = np.zeros((lmax + 1, N), dtype=np.complex128)
        M[0] = math.pow(PI / 2, 0.5) * np.exp(1j * (phi1 + phi2 + phi3)) / N
        M[0, 0] *= 0.5
        M[1] = np.exp(2j * (-phi2 - np.arctan(2 * t))) * M[0]
        for l in range(1, lmax):
            M[l + 1] = np.exp(2j * (-np.arctan(2 * t / (2 * l + 1)))) * M[l - 1]
        self.M, self.N, self.lmax = M, N, lmax
        self.kgrid = self.ks[:N // 2].copy()

    def forward(self, f, l):
        """sbt.c:102-158: f is r*R(r) on the input grid."""
        N, h = self.N, self.N // 2
        fs = np.empty(N)
        Cc = f[0] / math.pow(self.rs[h], l + 1)
        fs[:h] = Cc * np.power(self.rs[:h], l + 1)
        fs[h:] = f
        x = _backward(np.power(self.rs, 0.5) * fs)
        x = x * self.M[l]
        x[h:] = 0
        x = _backward(x)
        return np.real(x[:h]) * 2 / np.power(self.ks[:h], 1.5)

    def inverse(self, f, l):
        """sbt.c:160-215 (ks and rs swap roles)."""
        N, h = self.N, self.N // 2
        ks, r = self.rs, self.ks
        fs = np.zeros(N)
        fs[:h] = f
        x = _backward(np.power(r, 1.5) * fs)
        x = x * self.M[l]
        x[h:] = 0
        x = _backward(x)
        return np.real(x[h:]) / PI * 2 * 2 / np.power(ks[h:], 1.5)


# --------------------------------------------------------------------------- #
# per-element PAW setup (projector.c:20-171, 507-558)
# --------------------------------------------------------------------------- #
class PPot:
    pass


def build_ppot(pp, grid_encut):
    """get_projector_list for one element (projector.c:33-167)."""
    o = PPot()
    o.ls = [int(l) for l in pp.ls]
    o.num_projs = len(o.ls)
    o.rmax = float(pp.rmax)
    o.proj_gridsize = int(pp.ndata)
    o.wave_grid = np.asarray(pp.grid, dtype=np.float64)
    o.wave_gridsize = len(o.wave_grid)
    o.total_projs = sum(2 * l + 1 for l in o.ls)
    o.lmax = max(o.ls)
    n = o.proj_gridsize
    o.proj_grid = o.rmax / n * np.arange(n)                       # :55-58
    dense = np.empty(o.wave_gridsize)                              # :65-70
    dense[0] = o.wave_grid[0]
    factor = math.pow(o.wave_grid[1] / o.wave_grid[0], 1.0)
    for p in range(1, o.wave_gridsize):
        dense[p] = dense[p - 1] * factor
    o.wave_rmax = float(o.wave_grid[-1])                           # :71
    o.smooth_grid = o.wave_rmax / n * np.arange(n)                 # :74-76
    o.proj = [np.asarray(p, dtype=np.float64) for p in pp.realprojs]
    o.aewave = [np.asarray(p, dtype=np.float64) for p in pp.aewaves]
    o.pswave = [np.asarray(p, dtype=np.float64) for p in pp.pswaves]
    o.diffwave = [a - b for a, b in zip(o.aewave, o.pswave)]       # :98
    o.proj_spline = [spline_coeff(o.proj_grid, f) for f in o.proj]
    o.diffwave_spline = [spline_coeff(o.wave_grid, f) for f in o.diffwave]
    sbt = SBT(1e7, 0, o.lmax, o.wave_grid)                         # :115-117
    o.kwave_grid = sbt.kgrid
    o.kwave = [sbt.forward(f, l) for f, l in zip(o.diffwave, o.ls)]
    o.kwave_spline = [spline_coeff(o.kwave_grid, f) for f in o.kwave]
    cut = math.pow(CCONST * grid_encut, 0.5)                       # :132
    o.smooth_diffwave, o.smooth_diffwave_spline = [], []
    for k, l in enumerate(o.ls):
        dk = np.zeros(o.wave_gridsize)
        q = 0
        while q < o.wave_gridsize and o.kwave_grid[q] < cut:
            dk[q] = o.kwave[k][q]
            q += 1
        sm = sbt.inverse(dk, l)                                    # :136-137
        sms = spline_coeff(dense, sm)
        sdw = np.zeros(n)
        sdw[1:] = wave_interpolate(o.smooth_grid[1:], dense, sm, sms)   # :145-151
        sdw[0] = 0.0 if l > 0 else sdw[1]                          # :143-153
        o.smooth_diffwave.append(sdw)
        o.smooth_diffwave_spline.append(spline_coeff(o.smooth_grid, sdw))
    # make_pwave_overlap_matrices (projector.c:507-558)
    P = o.num_projs
    o.psov, o.aeov, o.diov = np.zeros((P, P)), np.zeros((P, P)), np.zeros((P, P))
    for i in range(P):
        for j in range(i, P):
            if o.ls[i] == o.ls[j]:
                for mat, prod in ((o.psov, o.pswave[i] * o.pswave[j]),
                                  (o.aeov, o.aewave[i] * o.aewave[j]),
                                  (o.diov, (o.aewave[i] - o.pswave[i]) * (o.aewave[j] - o.pswave[j]))):
                    mat[i, j] = mat[j, i] = spline_integral(o.wave_grid, prod,
                                                            spline_coeff(o.wave_grid, prod))
    # channel table (utils.c:617-632): radial index n, then m=-l..l
    o.chan = [(j, l, m) for j, l in enumerate(o.ls) for m in range(-l, l + 1)]
    return o


def build_ppots(pps, grid_encut):
    return [build_ppot(pps[k], grid_encut) for k in sorted(pps.keys())]


# --------------------------------------------------------------------------- #
# sphere geometry + tables (utils.c:590-696)
# --------------------------------------------------------------------------- #
def _bbox(lattice, rmax, fftg):
    """utils.c:641-646."""
    L = np.asarray(lattice, dtype=np.float64).reshape(3, 3)
    vol = determinant(L)
    g = []
    for (a, b, n) in ((1, 2, 0), (0, 2, 1), (0, 1, 2)):
        res = np.cross(L[a], L[b])
        g.append(int(mag(res) * rmax / vol * fftg[n]) + 1)
    return g


def sphere_points(coord, lattice, fftg, rmax, R0, strict_radius=None):
    """Candidate loop of utils.c:647-671 (and density.c:262-296).
    Returns (i,j,k unwrapped ints [n,3], linear index [n], Cartesian offsets [n,3])."""
    g = _bbox(lattice, rmax, fftg)
    def c_round(v):  # C round(): halves away from zero
        return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))
    cen = [c_round(coord[d] * fftg[d]) for d in range(3)]
    ax = [np.arange(-g[d] + cen[d], g[d] + cen[d] + 1) for d in range(3)]
    I, J, K = np.meshgrid(*ax, indexing="ij")
    t = np.stack([I / float(fftg[0]) - coord[0], J / float(fftg[1]) - coord[1],
                  K / float(fftg[2]) - coord[2]], axis=-1)
    cart = frac_to_cartesian(t, lattice)
    r = mag(cart)
    m = r < R0
    ijk = np.stack([I[m], J[m], K[m]], axis=1)
    w = np.stack([((ijk[:, d] % fftg[d]) + fftg[d]) % fftg[d] for d in range(3)], axis=1)
    lin = w[:, 0] * fftg[1] * fftg[2] + w[:, 1] * fftg[2] + w[:, 2]
    return ijk, w, lin.astype(np.int32), cart[m]


def setup_site(ppots, site_nums, labels, coords, lattice, fftg, pr0_pw1):
    """utils.c:590-696.  One dict per listed site: indices, paths, values[lm, npts]."""
    coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
    out = []
    for s, p in enumerate(site_nums):
        pp = ppots[labels[p]]
        rmax = pp.wave_rmax if pr0_pw1 else pp.rmax
        # utils.c:651-652: divisor uses labels[s] (loop index), numerator labels[p]
        R0 = (pp.proj_gridsize - 1) * rmax / ppots[labels[s]].proj_gridsize
        ijk, w, lin, paths = sphere_points(coords[p], lattice, fftg, rmax, R0)
        frac = np.stack([w[:, d] / float(fftg[d]) for d in range(3)], axis=1)
        vec, r = min_cart_path(frac, coords[p], lattice)          # proj_value, utils.c:566-588
        x = pp.smooth_grid if pr0_pw1 else pp.proj_grid
        vals = np.zeros((pp.total_projs, len(lin)), dtype=np.complex128)
        for n, (j, l, m) in enumerate(pp.chan):
            f = pp.smooth_diffwave[j] if pr0_pw1 else pp.proj[j]
            sp = pp.smooth_diffwave_spline[j] if pr0_pw1 else pp.proj_spline[j]
            rad = proj_interpolate(r, rmax, x, f, sp)
            vals[n] = radial_times_ylm(rad, l, m, vec, r)
        out.append(dict(index=int(p), elem=int(labels[p]), indices=lin, paths=paths, values=vals))
    return out


# --------------------------------------------------------------------------- #
# FFT box (linalg.c:14-79)
# --------------------------------------------------------------------------- #
def box_index(Gs, fftg):
    g = (np.asarray(Gs, dtype=np.int64) + np.asarray(fftg)) % np.asarray(fftg)
    return g[:, 0] * fftg[1] * fftg[2] + g[:, 1] * fftg[2] + g[:, 2]


def fft3d(Gs, Cs, lattice, fftg):
    """linalg.c:14-45."""
    n = int(fftg[0]) * int(fftg[1]) * int(fftg[2])
    x = np.zeros(n, dtype=np.complex128)
    x[box_index(Gs, fftg)] = np.asarray(Cs).astype(np.complex128)
    x = np.fft.ifftn(x.reshape(tuple(int(v) for v in fftg))) * n
    return x * math.pow(determinant(lattice), -0.5)


def fwd_fft3d(x, Gs, lattice, fftg):
    """linalg.c:47-79 (returns complex64)."""
    n = int(fftg[0]) * int(fftg[1]) * int(fftg[2])
    y = np.fft.fftn(np.asarray(x).reshape(tuple(int(v) for v in fftg)))
    y = y * (math.pow(determinant(lattice), 0.5) / fftg[0] / fftg[1] / fftg[2])
    return y.reshape(n)[box_index(Gs, fftg)].astype(np.complex64)


# --------------------------------------------------------------------------- #
# <p_i|psi~>  (projector.c:223-274)
# --------------------------------------------------------------------------- #
def onto_projector_helper(x, sites, lattice, reclattice, kpt, fftg):
    dv = determinant(lattice) / fftg[0] / fftg[1] / fftg[2]
    kc = frac_to_cartesian(np.asarray(kpt, dtype=np.float64), reclattice)
    xf = np.asarray(x).reshape(-1)
    out = []
    for st in sites:
        pth = st["paths"]
        kdotr = kc[0] * pth[:, 0] + kc[1] * pth[:, 1] + kc[2] * pth[:, 2]
        xv = xf[st["indices"]] * dv * np.exp(1j * kdotr)
        out.append(np.conj(st["values"]) @ xv)
    return out


class Wavefunction:
    """Container restating pswf_t (utils.h:117-141) as arrays.
    kappa = k + s*nwk (utils.c:367-373)."""

    def __init__(self, lattice, kpts, kws, nspin, nband, Gs, Cs, occs, energies=None,
                 encut=0.0, ncl=False):
        self.lattice = np.asarray(lattice, dtype=np.float64).reshape(3, 3)
        self.reclattice = reciprocal_lattice(self.lattice)
        self.kpts = np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
        self.kws = np.asarray(kws, dtype=np.float64)
        self.nwk, self.nspin, self.nband = len(self.kpts), nspin, nband
        self.Gs = Gs            # list per k (len nwk) of int[npw,3]
        self.Cs = Cs            # list per kappa of complex64 [nband, npw_file]
        self.occs = np.asarray(occs, dtype=np.float64)   # [nkappa, nband]
        self.energies = None if energies is None else np.asarray(energies, dtype=np.float64)
        self.encut, self.ncl = encut, ncl
        self.P = None
        self.W = None

    @classmethod
    def from_image(cls, img, kws):
        """read_wavecar (reader.c:129-315) on an in-memory image."""
        from pawpyseed_b200.synth import enumerate_gvectors
        b = np.asarray(img, dtype=np.uint8)
        h = b[:24].view(np.float64)
        nrecl, nspin = int(round(h[0])), int(round(h[1]))
        h1 = b[nrecl:nrecl + 8 * 12].view(np.float64)
        nwk, nband, encut = int(round(h1[0])), int(round(h1[1])), float(h1[2])
        lattice = h1[3:12].copy().reshape(3, 3)
        kpts, Gs, Cs, occs, ens = [], [], [], [], []
        ncl = False
        for iwk in range(nwk * nspin):
            base = 2 + iwk * (1 + nband)
            kh = b[base * nrecl:(base + 1) * nrecl].view(np.float64)
            nplane = int(round(kh[0]))
            k = kh[1:4].copy()
            ens.append(kh[4:4 + 3 * nband:3].copy())
            occs.append(kh[6:6 + 3 * nband:3].copy())
            if iwk < nwk:
                g = enumerate_gvectors(lattice, encut, k)
                if 2 * len(g) == nplane:
                    ncl = True
                elif len(g) != nplane:
                    raise ValueError("plane-wave count mismatch %d vs %d" % (len(g), nplane))
                kpts.append(k)
                Gs.append(g)
            blk = b[(base + 1) * nrecl:(base + 1 + nband) * nrecl].reshape(nband, nrecl)
            Cs.append(blk[:, :8 * nplane].copy().view(np.complex64))
        return cls(lattice, kpts, kws, nspin, nband, Gs, Cs, np.array(occs), np.array(ens),
                   encut, ncl)

    # -- setup_projections (projector.c:560-602) --------------------------------
    def setup_projections(self, pps, labels, coords, dim, grid_encut):
        self.ppots = build_ppots(pps, grid_encut)
        self.labels = np.asarray(labels, dtype=np.int32)
        self.coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
        self.fftg = np.asarray(dim, dtype=np.int32)
        self.num_sites = len(self.labels)
        self.sites = setup_site(self.ppots, list(range(self.num_sites)), self.labels,
                                self.coords, self.lattice, self.fftg, 0)
        self.P = self._project_all(self.sites)
        self.chan_index = np.array([(s, j, l, m) for s in range(self.num_sites)
                                    for (j, l, m) in self.ppots[self.labels[s]].chan],
                                   dtype=np.int32).reshape(-1, 4)
        self.site_off = np.concatenate([[0], np.cumsum(
            [self.ppots[self.labels[s]].total_projs for s in range(self.num_sites)])])

    def _project_all(self, sites, lattice=None, fftg=None):
        """onto_projector / onto_projector_ncl / onto_smoothpw over all bands
        (projector.c:333-418).  Returns list per kappa of c128 [nband, nproj]
        (ncl: [nband, 2, nproj] = up, down)."""
        lattice = self.lattice if lattice is None else lattice
        fftg = self.fftg if fftg is None else fftg
        out = []
        for kap in range(self.nwk * self.nspin):
            k = kap % self.nwk
            G = self.Gs[k]
            rows = []
            for b in range(self.nband):
                c = self.Cs[kap][b]
                if self.ncl:
                    h = len(c) // 2
                    halves = []
                    for part in (c[:h], c[h:]):
                        x = fft3d(G, part, lattice, fftg)
                        halves.append(np.concatenate(onto_projector_helper(
                            x, sites, lattice, self.reclattice, self.kpts[k], fftg))
                            if sites else np.zeros(0, np.complex128))
                    rows.append(np.stack(halves))
                else:
                    x = fft3d(G, c, lattice, fftg)
                    rows.append(np.concatenate(onto_projector_helper(
                        x, sites, lattice, self.reclattice, self.kpts[k], fftg))
                        if sites else np.zeros(0, np.complex128))
            out.append(np.array(rows))
        return out

    # -- pseudoprojection (pseudoprojector.c:63-90), FP64 accumulation -----------
    def pseudoprojection(self, band_num, basis, flip_spin=False):
        NK = basis.nwk * basis.nspin
        res = np.zeros(basis.nband * NK, dtype=np.complex128)
        for kap in range(NK):
            kp = kap
            if basis.nspin == 2 and flip_spin:
                kp = kap + basis.nwk if kap < basis.nwk else kap - basis.nwk
            c1 = self.Cs[kap][band_num].astype(np.complex128)
            c2 = basis.Cs[kp].astype(np.complex128)
            res[kap::NK] = np.conj(c2) @ c1
        return res

    # -- realspace_state (density.c:232-421) --------------------------------------
    def realspace_state(self, b, kap, fftg=None):
        fftg = self.fftg if fftg is None else np.asarray(fftg, dtype=np.int32)
        k = kap % self.nwk
        kvec = self.kpts[k]
        G = self.Gs[k]
        c = self.Cs[kap][b]
        n = int(np.prod(fftg))
        parts = [c[:len(c) // 2], c[len(c) // 2:]] if self.ncl else [c]
        I, J, K = np.meshgrid(*[np.arange(fftg[d]) / float(fftg[d]) for d in range(3)], indexing="ij")
        kdotr = kvec[0] * I + kvec[1] * J + kvec[2] * K
        ph = np.exp(2 * PI_DENSITY * 1j * kdotr)
        xs = [(fft3d(G, p, self.lattice, fftg) * ph).reshape(n) for p in parts]
        for p in range(self.num_sites):
            pp = self.ppots[self.labels[p]]
            rmax = float(pp.wave_grid[-1])
            ijk, w, lin, cart = sphere_points(self.coords[p], self.lattice, fftg, rmax, rmax)
            if len(lin) == 0:
                continue
            r = mag(cart)
            wrap = np.stack([(w[:, d] - ijk[:, d]) // fftg[d] for d in range(3)], axis=1)
            pc = self.coords[p][None, :] + wrap
            phase = pc[:, 0] * kvec[0] + pc[:, 1] * kvec[1] + pc[:, 2] * kvec[2]
            eph = np.exp(2 * PI_DENSITY * 1j * phase)
            lo, hi = self.site_off[p], self.site_off[p + 1]
            for h, x in enumerate(xs):
                ov = self.P[kap][b][h, lo:hi] if self.ncl else self.P[kap][b][lo:hi]
                add = np.zeros(len(lin), dtype=np.complex128)
                for nidx, (j, l, m) in enumerate(pp.chan):
                    rad = wave_interpolate(r, pp.wave_grid, pp.diffwave[j], pp.diffwave_spline[j])
                    rad = np.where(r < pp.wave_grid[0], rad / pp.wave_grid[0],
                                   rad / np.where(r == 0, 1.0, r))
                    add += radial_times_ylm(rad, l, m, cart, r) * ov[nidx]
                np.add.at(x, lin, add * eph)
        shp = tuple(int(v) for v in fftg)
        return np.stack([x.reshape(shp) for x in xs]) if self.ncl else xs[0].reshape(shp)

    def remove_phase(self, x, kap, fftg=None):
        """density.c:312-326."""
        fftg = self.fftg if fftg is None else fftg
        kvec = self.kpts[kap % self.nwk]
        I, J, K = np.meshgrid(*[np.arange(fftg[d]) / float(fftg[d]) for d in range(3)], indexing="ij")
        return x * np.exp(-2 * PI_DENSITY * 1j * (kvec[0] * I + kvec[1] * J + kvec[2] * K))

    def chg_density(self, fftg):
        """ae_chg_density / ncl_ae_chg_density (density.c:158-203)."""
        fftg = np.asarray(fftg, dtype=np.int32)
        P = np.zeros(tuple(int(v) for v in fftg))
        mult = 1 if self.ncl else 2 // self.nspin
        for kap in range(self.nwk * self.nspin):
            for b in range(self.nband):
                occ = self.occs[kap, b]
                if occ > 0:
                    x = self.realspace_state(b, kap, fftg)
                    d = np.real(x * np.conj(x))
                    if self.ncl:
                        d = d[0] + d[1]
                    P += d * self.kws[kap % self.nwk] * occ * mult
        return P



# --------------------------------------------------------------------------- #
# k-point desymmetrisation (utils.c:829-1098 `expand_symm_wf`)
# --------------------------------------------------------------------------- #
def expand_symm_wf(rwf: "Wavefunction", maps, ops, drs, kws, trs):
    """New Wavefunction whose k-point q is ops[q] applied to rwf's k-point maps[q]
    (time-reversed if trs[q]) with fractional translation drs[q].  Coefficients and
    phase factors are complex64 like the reference's (utils.c:1040-1081)."""
    from pawpyseed_b200.synth import enumerate_gvectors
    if rwf.ncl:
        raise NotImplementedError("utils.c:994-1056 cannot map the duplicated spinor G list")
    maps = np.asarray(maps, dtype=np.int64)
    ops = np.asarray(ops, dtype=np.float64).reshape(-1, 3, 3)
    drs = np.asarray(drs, dtype=np.float64).reshape(-1, 3)
    nk = len(maps)
    kpts, Gs, Cs, occs, ens = [], [], [], [], []
    kdiffs = []
    for q in range(nk):
        k = ops[q] @ rwf.kpts[maps[q]]                     # utils.c:885-889 rotate_frac
        if trs[q] == 1:
            k = -k
        kd = np.array([math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5) for v in k])  # C round()
        k = k - kd
        for i in range(3):                                  # utils.c:900-905
            if abs(k[i] + 0.5) < 0.0001:
                kd[i] -= 1
                k[i] += 1
        kpts.append(k)
        kdiffs.append(kd)
        g = enumerate_gvectors(rwf.lattice, rwf.encut, k)   # utils.c:928-975 (same order, integer bounds)
        if len(g) != len(rwf.Gs[maps[q]]):
            raise ValueError("plane-wave count mismatch in expand_symm_wf")
        Gs.append(g)
    for kap in range(nk * rwf.nspin):
        q = kap % nk
        rnum = int(maps[q]) + (rwf.nwk if (kap >= nk and rwf.nspin == 2) else 0)
        g_new, g_old = Gs[q], rwf.Gs[maps[q]]
        lookup = {tuple(v): w for w, v in enumerate(g_new.tolist())}
        pw = (g_old.astype(np.float64) @ ops[q].T)
        if trs[q] == 1:
            pw = -pw
        pw = pw + kdiffs[q]
        gi = np.rint(pw).astype(np.int64)
        npw = len(g_new)
        gmaps = np.full(npw, -1, dtype=np.int64)
        factors = np.zeros(npw, dtype=np.complex64)
        sign = -1.0 if trs[q] == 0 else 1.0
        for g in range(npw):
            w = lookup.get((int(gi[g, 0]), int(gi[g, 1]), int(gi[g, 2])), -1)
            if w < 0:
                raise ValueError("bad plane-wave mapping")
            gmaps[w] = g
            ph = float(np.dot(kpts[q], drs[q]) + np.dot(pw[g], drs[q]))
            arg = np.float32(sign * 2 * PI * ph)
            factors[w] = np.complex64(complex(math.cos(float(arg)), math.sin(float(arg))))
        if (gmaps < 0).any():
            raise ValueError("incomplete plane-wave mapping")
        src = rwf.Cs[rnum][:, gmaps].astype(np.complex64)
        new = (factors[None, :] * src).astype(np.complex64)
        if trs[q] == 1:
            new = np.conj(new)
        Cs.append(new)
        occs.append(rwf.occs[rnum])
        ens.append(rwf.energies[rnum] if getattr(rwf, "energies", None) is not None else np.zeros(rwf.nband))
    return Wavefunction(rwf.lattice, kpts, kws, rwf.nspin, rwf.nband, Gs, Cs, np.array(occs), np.array(ens),
                        rwf.encut, False)

# --------------------------------------------------------------------------- #
# off-site partial-wave overlap (radial.c:116-196, gaunt.py:17-30)
# --------------------------------------------------------------------------- #
@lru_cache(maxsize=None)
def sbtfac(l1, l2, Lidx, m1off, m2):
    """SBTFACS[l1][l2][(L-|l1-l2|)/2][l1+m1][m2] regenerated with sympy (gaunt.py:17-30)."""
    from sympy import N as symN
    from sympy.physics.wigner import wigner_3j
    m1 = m1off - l1
    L = abs(l1 - l2) + 2 * Lidx
    if l2 > l1 or m2 > l2 or L > l1 + l2:
        return 0.0
    v = symN(wigner_3j(l1, l2, L, 0, 0, 0)) * symN(wigner_3j(l1, l2, L, -m1, m2, m1 - m2))
    return float(v) * math.sqrt((2 * l1 + 1) * (2 * l2 + 1) * (2 * L + 1) / 4 / np.pi)


def sbf(x, l):
    """utils.c:807-827, vectorised."""
    x = np.asarray(x, dtype=np.float64)
    small = x < 10e-6
    xs = np.where(small, 1.0, x)
    jm1 = np.sin(xs) / xs
    jl = np.sin(xs) / (xs * xs) - np.cos(xs) / xs
    if l == 0:
        res = jm1
    elif l == 1:
        res = jl
    else:
        jp = jl
        for ll in range(1, l):
            jp = (2 * ll + 1) / xs * jl - jm1
            jm1, jl = jl, jp
        res = jp
    return np.where(small, 1.0 if l == 0 else 0.0, res)


def reciprocal_offsite_wave_overlap(dcoord, k1, f1, s1, k2, f2, s2, l1, m1, l2, m2):
    """radial.c:116-196."""
    if l1 < l2:
        lx, ly, mx, my = l2, l1, m2, m1
    else:
        lx, ly, mx, my = l1, l2, m1, m2
    if my < 0:
        mx, my = -mx, -my
    kmax = min(k1[-1], k2[-1])
    kmin = max(k1[0], k2[0])
    d = np.asarray(dcoord, dtype=np.float64)
    R = float(mag(d))
    if R < 10e-12:
        theta = phi = R = 0.0
    else:
        theta = math.acos(d[2] / R)
        if R - abs(d[2]) < 10e-12:
            phi = 0.0
        else:
            phi = math.acos(d[0] / math.pow(d[0] * d[0] + d[1] * d[1], 0.5))
        if d[1] < 0:
            phi = 2 * PI - phi
    kgrid = kmin * np.power(kmax / kmin, np.arange(KGRID_SIZE) / float(KGRID_SIZE))
    base = wave_interpolate(kgrid, k1, f1, s1) * wave_interpolate(kgrid, k2, f2, s2) * kgrid * kgrid
    total = 0j
    mult = (-1.0) ** m1 * 8
    for L in range(abs(l1 - l2), l1 + l2 + 1, 2):
        ifunc = base * sbf(kgrid * R, L)
        integ = spline_integral(kgrid, ifunc, spline_coeff(kgrid, ifunc))
        if R > 10e-10:
            yl = complex(Ylm(L, m1 - m2, theta, phi)) if abs(m1 - m2) <= L else 0j
            total += integ * sbtfac(lx, ly, (L - abs(l1 - l2)) // 2, lx + mx, my) * yl \
                * (1j) ** (l2 + L - l1) * mult
        elif L == 0 and l1 == l2 and m1 == m2:
            total += integ * 2 / PI
    return total


# --------------------------------------------------------------------------- #
# Projector: overlap_setup_real + compensation_terms (projector.c:604-725, 850-963)
# --------------------------------------------------------------------------- #
def aug_freqs(wf: "Wavefunction", site_nums):
    """get_aug_freqs + helper over all bands (projector.c:276-331, 420-453): the (phi - phit)
    augmentation of the listed sites, weighted by the band's projector overlaps, summed on the FFT
    grid with phase e^{-i k_cart.path}, forward-transformed and narrowed to complex64.
    Returns list per kappa of complex64 [nband, npw]."""
    st = setup_site(wf.ppots, site_nums, wf.labels, wf.coords, wf.lattice, wf.fftg, 1)
    n = int(np.prod(wf.fftg))
    out = []
    for kap in range(wf.nwk * wf.nspin):
        k = kap % wf.nwk
        kc = frac_to_cartesian(np.asarray(wf.kpts[k], dtype=np.float64), wf.reclattice)
        rows = []
        for b in range(wf.nband):
            x = np.zeros(n, dtype=np.complex128)
            for site in st:
                pth = site["paths"]
                phase = np.exp(-1j * (kc[0] * pth[:, 0] + kc[1] * pth[:, 1] + kc[2] * pth[:, 2]))
                p = wf.P[kap][b][wf.site_off[site["index"]]:wf.site_off[site["index"] + 1]]
                np.add.at(x, site["indices"], (p @ site["values"]) * phase)
            rows.append(fwd_fft3d(x, wf.Gs[k], wf.lattice, wf.fftg))
        out.append(np.array(rows))
    return out


class Projector:
    def __init__(self, wf: Wavefunction, basis: Wavefunction, site_cat, recip=False):
        self.S, self.R = wf, basis
        self.cat = [list(map(int, x)) for x in site_cat]
        M_R, M_S, N_R, N_S, N_RS_R, N_RS_S = self.cat
        R, S = self.R, self.S
        self.recip = recip
        if recip:
            # overlap_setup_recip parts 1-2 (projector.c:748-792); part 3 below is shared
            self.CA_R = aug_freqs(R, N_R) if N_R else None
            self.CA_S = aug_freqs(S, N_S) if N_S else None
            N_R, N_S = [], []
        # part 1: <(phi-phit)_R sites | psi_S>  on S's lattice / grid      (:625-646)
        self.W_S = None
        if N_R:
            st = setup_site(R.ppots, N_R, R.labels, R.coords, S.lattice, S.fftg, 1)
            self.W_S = S._project_all(st, S.lattice, S.fftg)
        # part 2: <(phi-phit)_S sites | psi_R>                             (:649-671)
        self.W_R = None
        if N_S:
            st = setup_site(S.ppots, N_S, S.labels, S.coords, R.lattice, R.fftg, 1)
            self.W_R = R._project_all(st, R.lattice, R.fftg)
        # part 3: off-site partial-wave overlaps                           (:682-719)
        self.dcoords, self.omega = [], []
        for s1, s2 in zip(N_RS_R, N_RS_S):
            pp1, pp2 = R.ppots[R.labels[s1]], S.ppots[S.labels[s2]]
            path, _ = min_cart_path(S.coords[s2][None, :], R.coords[s1], R.lattice)
            d = path[0]
            om = np.zeros((pp1.total_projs, pp2.total_projs), dtype=np.complex128)
            for tj, (j, l1, m1) in enumerate(pp1.chan):
                for tk, (k, l2, m2) in enumerate(pp2.chan):
                    om[tj, tk] = np.conj(reciprocal_offsite_wave_overlap(
                        d, pp1.kwave_grid, pp1.kwave[j], pp1.kwave_spline[j],
                        pp2.kwave_grid, pp2.kwave[k], pp2.kwave_spline[k], l1, m1, l2, m2))
            self.dcoords.append(d)
            self.omega.append(om)

    def compensation_terms(self, band_num, flip_spin=False, parts=False):
        """projector.c:850-963 -> c128[nband_R * NK], index b*NK + kappa."""
        M_R, M_S, N_R, N_S, N_RS_R, N_RS_S = self.cat
        R, S = self.R, self.S
        NK = R.nwk * R.nspin
        out = np.zeros((4, R.nband * NK), dtype=np.complex128)
        for kap in range(NK):
            kr = kap
            if R.nspin == 2 and flip_spin:
                kr = kap + R.nwk if kap < R.nwk else kap - R.nwk
            PR = R.P[kr]                      # [nband_R, nproj_R]
            ps = S.P[kap][band_num]           # [nproj_S]
            t = np.zeros((4, R.nband), dtype=np.complex128)
            for s1, s2 in zip(M_R, M_S):      # O_M (:890-910)
                pp = R.ppots[R.labels[s1]]
                chR = pp.chan
                chS = S.ppots[S.labels[s2]].chan
                a = PR[:, R.site_off[s1]:R.site_off[s1 + 1]]
                bvec = ps[S.site_off[s2]:S.site_off[s2 + 1]]
                for i, (ni, li, mi) in enumerate(chR):
                    for j, (nj, lj, mj) in enumerate(chS):
                        if li == lj and mi == mj:
                            t[0] += np.conj(a[:, i]) * (pp.aeov[ni, nj] - pp.psov[ni, nj]) * bvec[j]
            for s, site in enumerate(N_R):    # O_R (:915-924)
                a = PR[:, R.site_off[site]:R.site_off[site + 1]]
                nt = a.shape[1]
                w = self.W_S[kap][band_num]
                off = sum(R.ppots[R.labels[q]].total_projs for q in N_R[:s])
                t[1] += (np.conj(a) * w[off:off + nt][None, :]).sum(axis=1)
            for s, site in enumerate(N_S):    # O_S (:929-938)
                bvec = ps[S.site_off[site]:S.site_off[site + 1]]
                nt = len(bvec)
                off = sum(S.ppots[S.labels[q]].total_projs for q in N_S[:s])
                w = self.W_R[kr][:, off:off + nt]
                t[2] += (np.conj(w) * bvec[None, :]).sum(axis=1)
            for s, (s1, s2) in enumerate(zip(N_RS_R, N_RS_S)):   # O_N (:944-959)
                a = PR[:, R.site_off[s1]:R.site_off[s1 + 1]]
                bvec = ps[S.site_off[s2]:S.site_off[s2 + 1]]
                kvec = R.kpts[kr % R.nwk]
                d = self.dcoords[s]
                ph = np.exp(2j * PI * (kvec[0] * d[0] + kvec[1] * d[1] + kvec[2] * d[2]))
                t[3] += (np.conj(a) @ (self.omega[s] @ bvec)) * ph
            out[:, kap::NK] = t
        return out if parts else out.sum(axis=0)

    def compensation_terms_recip(self, band_num, flip_spin=False, fp32_dots=False):
        """projector.c:965-1077: <CA_R|C_S> + <C_R|CA_S> + O_M + O_N.  The reference accumulates the two
        plane-wave dot products in single precision (cblas_cdotc_sub); default here is FP64."""
        M_R, M_S, _, _, N_RS_R, N_RS_S = self.cat
        R, S = self.R, self.S
        NK = R.nwk * R.nspin
        saved = self.cat
        self.cat = [M_R, M_S, [], [], N_RS_R, N_RS_S]
        out = self.compensation_terms(band_num, flip_spin)
        self.cat = saved
        acc = np.complex64 if fp32_dots else np.complex128
        for kap in range(NK):
            kr = kap
            if R.nspin == 2 and flip_spin:
                kr = kap + R.nwk if kap < R.nwk else kap - R.nwk
            if self.CA_R is not None:
                out[kap::NK] += (np.conj(self.CA_R[kr].astype(acc)) @ S.Cs[kap][band_num].astype(acc))
            if self.CA_S is not None:
                out[kap::NK] += (np.conj(R.Cs[kr].astype(acc)) @ self.CA_S[kap][band_num].astype(acc))
        return out

    def single_band_projection(self, band_num, flip_spin=False):
        """projector.py:210-236."""
        comp = self.compensation_terms_recip if self.recip else self.compensation_terms
        return self.S.pseudoprojection(band_num, self.R, flip_spin) + comp(band_num, flip_spin)


def project_realspace_state(band_num, wf: Wavefunction, wf_R: Wavefunction, fftg):
    """density.c:205-230: <psi_R,b | psi_band> by brute-force integration of the AE states on `fftg`."""
    fftg = np.asarray(fftg, dtype=np.int32)
    NK = wf.nwk * wf.nspin
    n = int(np.prod(fftg))
    vol = determinant(wf.lattice)
    out = np.zeros(wf_R.nband * NK, dtype=np.complex128)
    for k in range(NK):
        st = wf.realspace_state(band_num, k, fftg).reshape(-1)
        for b in range(wf_R.nband):
            sr = wf_R.realspace_state(b, k, fftg).reshape(-1)
            out[b * NK + k] = np.vdot(sr, st) * (vol / n)
    return out


# --------------------------------------------------------------------------- #
# MomentumMatrix (momentum.c; pawpyc.pyx:738-807)
# --------------------------------------------------------------------------- #
def _dir_angles(v):
    """theta, phi of a Cartesian vector with the reference's branch choices (momentum.c:190-202, 533-545)."""
    r = mag(v)
    theta = math.acos(v[2] / r)
    if r - abs(v[2]) == 0:
        phi = 0.0
    else:
        phi = math.acos(v[0] / math.pow(v[0] * v[0] + v[1] * v[1], 0.5))
    if v[1] < 0:
        phi = 2 * PI - phi
    return theta, phi


class MomentumMatrix:
    def __init__(self, wf: Wavefunction, encut):
        from pawpyseed_b200.synth import enumerate_gvectors
        self.wf, self.encut = wf, float(encut)
        # momentum_grid_size + get_momentum_grid (momentum.c:359-363, 402-445): the reader's enumeration at k = 0
        self.ggrid = enumerate_gvectors(wf.lattice, self.encut, np.zeros(3)).astype(np.int32)
        lo, hi = self.ggrid.min(axis=0), self.ggrid.max(axis=0)
        self.gbounds = np.array([min(lo[0], 0), max(hi[0], 0), min(lo[1], 0), max(hi[1], 0), min(lo[2], 0), max(hi[2], 0)])
        self.gdim = np.array([self.gbounds[1] - self.gbounds[0] + 1, self.gbounds[3] - self.gbounds[2] + 1,
                              self.gbounds[5] - self.gbounds[4] + 1])
        # get_all_transforms (momentum.c:225-276): SBT of (phi_i phi_j - phit_i phit_j)/r for every L
        self.trans = []
        for pp in wf.ppots:
            t = {}
            for n1, l1 in enumerate(pp.ls):
                for n2, l2 in enumerate(pp.ls):
                    rho = (pp.aewave[n1] * pp.aewave[n2] - pp.pswave[n1] * pp.pswave[n2]) / pp.wave_grid
                    sbt = SBT(1e7, 0.0, l1 + l2, pp.wave_grid)
                    for L in range(abs(l1 - l2), l1 + l2 + 1, 2):
                        f = sbt.forward(rho, L)
                        t[(n1, n2, L)] = (sbt.kgrid, f, spline_coeff(sbt.kgrid, f))
            self.trans.append(t)

    def spher_momentum(self, elem, n1, l1, m1, n2, l2, m2, G):
        """momentum.c:156-223."""
        if l1 < l2:
            lx, ly, mx, my = l2, l1, m2, m1
        else:
            lx, ly, mx, my = l1, l2, m1, m2
        if my < 0:
            mx, my = -mx, -my
        total = 0j
        magG = mag(G)
        for L in range(abs(l1 - l2), l1 + l2 + 1, 2):
            k, f, sp = self.trans[elem][(n1, n2, L)]
            if magG == 0:
                sph = complex(Ylm(L, m2 - m1, 0.0, 0.0)) if (L == 0 and m1 == m2) else 0j
            else:
                th, ph = _dir_angles(G)
                sph = complex(Ylm(L, m2 - m1, th, ph)) if abs(m2 - m1) <= L else 0j
            total += sbtfac(lx, ly, (L - abs(l1 - l2)) // 2, lx + mx, my) * sph * 4 * PI * (1j) ** L \
                * (-1.0) ** m2 * float(wave_interpolate(np.array([magG]), k, f, sp)[0])
        return total

    def pseudo_momentum(self, GP, kap1, b1, kap2, b2):
        """momentum.c:47-107 (accumulated in FP64 here; the reference uses float complex)."""
        wf = self.wf
        G1, G2 = wf.Gs[kap1 % wf.nwk], wf.Gs[kap2 % wf.nwk]
        look = {tuple(g): w for w, g in enumerate(G1.tolist())}
        c1, c2 = wf.Cs[kap1][b1].astype(np.complex128), wf.Cs[kap2][b2].astype(np.complex128)
        total = 0j
        for w, g in enumerate(G2.tolist()):
            j = look.get((g[0] + GP[0], g[1] + GP[1], g[2] + GP[2]))
            if j is not None:
                total += c2[w] * np.conj(c1[j])
        return total

    def momentum_matrix_elems(self, b1, k1, s1, b2, k2, s2):
        """get_momentum_matrix (momentum.c:278-357)."""
        wf = self.wf
        kap1, kap2 = k1 + s1 * wf.nwk, k2 + s2 * wf.nwk
        P1, P2 = wf.P[kap1][b1], wf.P[kap2][b2]
        out = np.zeros(len(self.ggrid), dtype=np.complex128)
        for gi, GP in enumerate(self.ggrid.tolist()):
            total = self.pseudo_momentum(GP, kap1, b1, kap2, b2)
            Gc = frac_to_cartesian(wf.kpts[k1] - wf.kpts[k2] + np.array(GP, dtype=np.float64), wf.reclattice)
            cache = {}
            for s in range(wf.num_sites):
                e = int(wf.labels[s])
                if e not in cache:
                    ch = wf.ppots[e].chan
                    cache[e] = np.array([[self.spher_momentum(e, n1, l1, m1, n2, l2, m2, Gc)
                                          for (n2, l2, m2) in ch] for (n1, l1, m1) in ch])
                phase = np.exp(2j * PI * np.dot(np.array(GP, dtype=np.float64), wf.coords[s]))
                a = P1[wf.site_off[s]:wf.site_off[s + 1]]
                b = P2[wf.site_off[s]:wf.site_off[s + 1]]
                total += (np.conj(a) @ cache[e] @ b) * phase
            out[gi] = total
        return out

    def reciprocal_fullfw(self, b, k, s):
        """fullwf_reciprocal (momentum.c:465-543)."""
        wf = self.wf
        kap = k + s * wf.nwk
        G = wf.Gs[k]
        look = {tuple(g): w for w, g in enumerate(G.tolist())}
        # fill_grid wraps with the G_bounds extents: later plane waves overwrite aliased earlier ones
        lo = G.min(axis=0) if len(G) else np.zeros(3, int)
        c = wf.Cs[kap][b]
        inv_sqrt_vol = math.pow(determinant(wf.lattice), -0.5)
        out = np.zeros(len(self.ggrid), dtype=np.complex128)
        gb = self._wf_gbounds()
        fd = np.array([gb[1] - gb[0] + 1, gb[3] - gb[2] + 1, gb[5] - gb[4] + 1])
        grid = {}
        for w, g in enumerate(G.tolist()):
            grid[tuple(np.mod(np.mod(g, fd) + fd, fd))] = np.complex64(c[w])
        for gi, GP in enumerate(self.ggrid.tolist()):
            val = 0j
            if gb[0] <= GP[0] <= gb[1] and gb[2] <= GP[1] <= gb[3] and gb[4] <= GP[2] <= gb[5]:
                val += complex(grid.get(tuple(np.mod(np.array(GP) + fd, fd)), 0))
            Gc = frac_to_cartesian(-np.array(GP, dtype=np.float64) - wf.kpts[k], wf.reclattice)
            r = mag(Gc)
            for site in range(wf.num_sites):
                pp = wf.ppots[wf.labels[site]]
                phase = np.exp(-2j * PI * np.dot(np.array(GP, dtype=np.float64), wf.coords[site]))
                pr = wf.P[kap][b][wf.site_off[site]:wf.site_off[site + 1]]
                for p, (n, l, m) in enumerate(pp.chan):
                    rad = float(wave_interpolate(np.array([r]), pp.kwave_grid, pp.kwave[n], pp.kwave_spline[n])[0])
                    if r == 0:
                        sph = complex(Ylm(l, m, 0.0, 0.0))
                    else:
                        th, ph = _dir_angles(Gc)
                        sph = complex(Ylm(l, m, th, ph))
                    val += pr[p] * rad * sph * (1j) ** l * phase * 4 * PI * inv_sqrt_vol
            out[gi] = val
        return out

    def _wf_gbounds(self):
        """wf->G_bounds as the reader accumulates it over all k-points (reader.c:262-268), starting from 0."""
        allg = np.concatenate(self.wf.Gs)
        lo, hi = np.minimum(allg.min(axis=0), 0), np.maximum(allg.max(axis=0), 0)
        return [lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]]

    def g_from_fullfw(self, b1, k1, s1, b2, k2, s2, dG):
        """quick_overlap (momentum.c:547-574)."""
        v1, v2 = self.reciprocal_fullfw(b1, k1, s1), self.reciprocal_fullfw(b2, k2, s2)
        look = {tuple(g): w for w, g in enumerate(self.ggrid.tolist())}
        gb = self.gbounds
        total = 0j
        for w, g in enumerate(self.ggrid.tolist()):
            q = (g[0] + dG[0], g[1] + dG[1], g[2] + dG[2])
            if gb[0] <= q[0] <= gb[1] and gb[2] <= q[1] <= gb[3] and gb[4] <= q[2] <= gb[5]:
                wp = look.get(q)
                if wp is not None:
                    total += np.conj(v1[wp]) * v2[w]
        return total


def make_site_lists(coords_R, labels_R, coords_S, labels_S, lattice, rmax_R, rmax_S, tol=0.02):
    """projector.py:115-160 with pymatgen's periodic distance restated through
    min_cart_path; element identity is label equality.  rmax_* : per-label list."""
    cR = np.asarray(coords_R).reshape(-1, 3)
    cS = np.asarray(coords_S).reshape(-1, 3)
    M_R, M_S = [], []
    dist = np.zeros((len(cR), len(cS)))
    for i in range(len(cR)):
        _, r = min_cart_path(cS % 1.0, cR[i] % 1.0, lattice)
        dist[i] = r
        for j in range(len(cS)):
            if r[j] <= tol and labels_R[i] == labels_S[j]:
                M_R.append(i)
                M_S.append(j)
    N_R = [i for i in range(len(cR)) if i not in M_R]
    N_S = [j for j in range(len(cS)) if j not in M_S]
    N_RS = [(i, j) for i in N_R for j in N_S
            if dist[i, j] < rmax_R[labels_R[i]] + rmax_S[labels_S[min(i, len(cS) - 1)]]]
    return M_R, M_S, N_R, N_S, [p[0] for p in N_RS], [p[1] for p in N_RS]
