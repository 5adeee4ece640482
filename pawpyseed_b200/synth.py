"""Synthetic inputs for the PAW band-projection path (numpy only, no GPU).

POTCAR files are licensed and not shipped with the reference, and the named
benchmark shapes (BASELINE.json configs 2-5) have no public WAVECAR, so tests
and bench.py drive the engine with

* an analytic PAW dataset (`SyntheticPseudopotential`) carrying exactly the
  attributes the reference flattens in pawpyc.pyx:368-389
  (``ls, ndata, grid, realprojs, aewaves, pswaves, rmax``), and
* WAVECAR byte images (`wavecar_image`) laid out exactly as the reference
  reader expects (reader.c:129-315): the image is handed to
  ``read_wavefunctions_from_str`` of either engine.

Nothing here is on the timed path; it only manufactures inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

CCONST = 0.262465831  # hbar^2/2m in eV A^2 units as used by reader.c:11


# --------------------------------------------------------------------------- #
# lattice / plane-wave enumeration
# --------------------------------------------------------------------------- #
def reciprocal_lattice(lattice: np.ndarray) -> np.ndarray:
    """Rows b_i = 2 pi (a_j x a_k) / V  (reader.c:58-64)."""
    a = np.asarray(lattice, dtype=np.float64).reshape(3, 3)
    vol = np.linalg.det(a)
    b = np.empty((3, 3))
    b[0] = np.cross(a[1], a[2])
    b[1] = np.cross(a[2], a[0])
    b[2] = np.cross(a[0], a[1])
    return b * (2.0 * math.pi / vol)


def _nbmax(lattice: np.ndarray, encut: float):
    """WaveTrans bounds nb{1,2,3}max as doubles (reader.c:65-102)."""
    b = reciprocal_lattice(lattice)
    mag = np.linalg.norm(b, axis=1)
    g = math.sqrt(encut * CCONST)

    def one(i, j, k):
        # angle between b_i and b_j, and b_k against their normal
        phi = math.acos(np.dot(b[i], b[j]) / (mag[i] * mag[j]))
        v = np.cross(b[i], b[j])
        sin3 = np.dot(b[k], v) / (np.linalg.norm(v) * mag[k])
        out = [0.0, 0.0, 0.0]
        out[i] = g / (mag[i] * abs(math.sin(phi))) + 1
        out[j] = g / (mag[j] * abs(math.sin(phi))) + 1
        out[k] = g / (mag[k] * abs(sin3)) + 1
        return out

    A = one(0, 1, 2)
    B = one(0, 2, 1)
    C = one(2, 1, 0)  # phi23 uses (b3, b2); normal b2 x b3 against b1
    return tuple(max(A[i], B[i], C[i]) for i in range(3))


def enumerate_gvectors(lattice, encut: float, kpt) -> np.ndarray:
    """G list for one k-point in WAVECAR coefficient order (reader.c:230-271):
    ig3 outermost ... ig1 innermost, kept iff |k+G|^2 / c <= encut."""
    b = reciprocal_lattice(lattice)
    nb1, nb2, nb3 = _nbmax(lattice, encut)

    def axis(nb):
        n = int(math.floor(2 * nb))
        ig = np.arange(0, n + 1)
        # C: ig - 2*nb - 1 evaluated in double then truncated toward zero
        neg = np.trunc(ig - 2 * nb - 1).astype(np.int64)
        return np.where(ig > nb, neg, ig).astype(np.int64)

    g1, g2, g3 = axis(nb1), axis(nb2), axis(nb3)
    k = np.asarray(kpt, dtype=np.float64)
    out = []
    # vectorise over (ig2, ig1) per ig3 plane to bound memory
    G2, G1 = np.meshgrid(g2, g1, indexing="ij")
    for z in g3:
        s = ((k[0] + G1)[..., None] * b[0] + (k[1] + G2)[..., None] * b[1]
             + (k[2] + z) * b[2])
        e = np.sum(s * s, axis=-1) / CCONST
        # the reference computes pow(mag,2)/c with mag = pow(dot,0.5); equal to
        # within an ulp, and synthetic cutoffs are chosen away from shell edges
        m = e <= encut
        if m.any():
            out.append(np.stack([G1[m], G2[m], np.full(m.sum(), z)], axis=1))
    return np.concatenate(out).astype(np.int32)


def smooth_fft_size(nmin: int) -> int:
    """Smallest N >= nmin whose only prime factors are 2, 3, 5, 7."""
    n = max(int(nmin), 2)
    while True:
        m = n
        for p in (2, 3, 5, 7):
            while m % p == 0:
                m //= p
        if m == 1:
            return n
        n += 1


def fft_grid_for(gvecs_per_k, factor: float = 3.0) -> np.ndarray:
    """VASP PREC=Normal style box: N_i >= factor * max|G_i|."""
    gmax = np.zeros(3, dtype=np.int64)
    for g in gvecs_per_k:
        gmax = np.maximum(gmax, np.abs(g).max(axis=0))
    return np.array([smooth_fft_size(int(math.ceil(factor * x))) for x in gmax],
                    dtype=np.int32)


# --------------------------------------------------------------------------- #
# WAVECAR image
# --------------------------------------------------------------------------- #
def wavecar_image(lattice, encut, kpts, nspin, nband, coeffs, occs=None,
                  energies=None, ncl=False, gvecs=None) -> np.ndarray:
    """Build an in-memory WAVECAR (uint8 array) that `read_wavefunctions_from_str`
    parses (layout: SURVEY App. D / reader.c:129-228).

    coeffs: callable (kappa, npw_file) -> complex64 [nband, npw_file] (or None to leave the
            block as zero pages), or a list indexed by kappa = s*nwk + k.  For ncl, npw_file = 2*npw.
    """
    lattice = np.asarray(lattice, dtype=np.float64).reshape(3, 3)
    kpts = np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
    nwk = len(kpts)
    if gvecs is None:
        gvecs = [enumerate_gvectors(lattice, encut, k) for k in kpts]
    mult = 2 if ncl else 1
    npw_file = [mult * len(g) for g in gvecs]
    nrecl = max(8 * max(npw_file), 8 * (4 + 3 * nband), 8 * 12)
    nrec = 2 + nwk * nspin * (1 + nband)
    img = np.zeros(nrec * nrecl, dtype=np.uint8)

    def rec(i):
        return img[i * nrecl:(i + 1) * nrecl]

    rec(0)[:24].view(np.float64)[:] = [nrecl, nspin, 45200]
    hdr = np.zeros(12 + 1)
    hdr[0], hdr[1], hdr[2] = nwk, nband, encut
    hdr[3:12] = lattice.reshape(9)
    rec(1)[:8 * 13].view(np.float64)[:] = hdr
    if occs is None:
        occs = np.where(np.arange(nband) < (nband + 1) // 2, 1.0, 0.0)
    if energies is None:
        energies = np.linspace(-5.0, 5.0, nband)
    occs = np.broadcast_to(np.asarray(occs, dtype=np.float64), (nwk * nspin, nband)) \
        if np.ndim(occs) == 1 else np.asarray(occs, dtype=np.float64)
    energies = np.broadcast_to(np.asarray(energies, dtype=np.float64), (nwk * nspin, nband)) \
        if np.ndim(energies) == 1 else np.asarray(energies, dtype=np.float64)
    for s in range(nspin):
        for k in range(nwk):
            kap = s * nwk + k
            base = 2 + kap * (1 + nband)
            h = np.zeros(4 + 3 * nband)
            h[0] = npw_file[k]
            h[1:4] = kpts[k]
            h[4::3] = energies[kap]
            h[6::3] = occs[kap]
            rec(base)[:8 * len(h)].view(np.float64)[:] = h
            c = coeffs(kap, npw_file[k]) if callable(coeffs) else coeffs[kap]
            if c is None:       # block left as untouched zero pages (another rank's shard)
                continue
            c = np.ascontiguousarray(c, dtype=np.complex64)
            assert c.shape == (nband, npw_file[k]), (c.shape, nband, npw_file[k])
            view = img[(base + 1) * nrecl:(base + 1 + nband) * nrecl].reshape(nband, nrecl)
            view[:, :8 * npw_file[k]] = c.view(np.uint8).reshape(nband, -1)
    return img


def random_coeffs(seed_base: int, nband: int, orthonormal=False):
    """Coefficient generator of SURVEY 8d: N(0,1) real/imag, L2-normalised bands."""
    def gen(kap, npw):
        rng = np.random.default_rng(seed_base + 10 * kap)
        c = rng.standard_normal((nband, npw), dtype=np.float32) \
            + 1j * rng.standard_normal((nband, npw), dtype=np.float32)
        if orthonormal:
            q, _ = np.linalg.qr(c.astype(np.complex128).T)
            return q.T.astype(np.complex64)
        c /= np.linalg.norm(c, axis=1, keepdims=True)
        return c.astype(np.complex64)
    return gen


# --------------------------------------------------------------------------- #
# analytic PAW dataset
# --------------------------------------------------------------------------- #
@dataclass
class SyntheticPseudopotential:
    """Same attribute names as reference wavefunction.py:17-140 `Pseudopotential`."""
    ls: list
    rmax: float = 1.5
    wave_rmax: float = 1.45
    ndata: int = 100
    nwave: int = 323
    rmin: float = 1e-4
    grid: np.ndarray = field(init=False)
    realprojs: list = field(init=False)
    aewaves: list = field(init=False)
    pswaves: list = field(init=False)
    augs: np.ndarray = field(init=False)

    def __post_init__(self):
        self.ls = [int(l) for l in self.ls]
        self.grid = self.rmin * (self.wave_rmax / self.rmin) ** (
            np.arange(self.nwave) / (self.nwave - 1))
        self.projgrid = np.arange(self.ndata) * self.rmax / self.ndata
        self.realprojs, self.aewaves, self.pswaves = [], [], []
        r, rp, rc = self.grid, self.projgrid, self.wave_rmax
        for n, l in enumerate(self.ls):
            a = 1.0 + 0.7 * (n % 2)
            ps = r ** (l + 1) * np.exp(-a * r * r)
            ae = ps + 0.8 * r ** (l + 1) * (1 - (r / rc) ** 2) ** 3 * np.cos(6 * r * (1 + n % 2))
            p = rp ** l * np.exp(-2 * a * rp * rp) * (1 - (rp / self.rmax) ** 2) ** 2
            self.pswaves.append(ps)
            self.aewaves.append(ae)
            self.realprojs.append(p)
        self.augs = np.zeros(1)


ELEMENT_CHANNELS = {"Si": [0, 0, 1, 1], "Ga": [0, 0, 1, 1, 2, 2], "N": [0, 0, 1, 1]}
ELEMENT_RADII = {"Si": (1.5, 1.45), "Ga": (1.6, 1.5), "N": (1.4, 1.4)}


def synthetic_pps(elements):
    """dict label(int) -> SyntheticPseudopotential, labels in the given element order
    (wavefunction.py:395-403 assigns labels in dict order)."""
    return {i: SyntheticPseudopotential(ELEMENT_CHANNELS[e], *ELEMENT_RADII[e])
            for i, e in enumerate(elements)}


def flatten_pps(pps):
    """The exact flattening of pawpyc.pyx:359-389 -> arrays for get_projector_list."""
    clabels, ls, wgrids, projectors, aewaves, pswaves, rmaxs = [], [], [], [], [], [], []
    for num in sorted(pps.keys()):
        pp = pps[num]
        clabels += [num, len(pp.ls), pp.ndata, len(pp.grid)]
        rmaxs.append(pp.rmax)
        ls += list(pp.ls)
        wgrids.append(pp.grid)
        for i in range(len(pp.ls)):
            projectors.append(pp.realprojs[i])
            aewaves.append(pp.aewaves[i])
            pswaves.append(pp.pswaves[i])
    cat = lambda x: np.ascontiguousarray(np.concatenate(x), dtype=np.float64)
    return (np.array(clabels, np.int32), np.array(ls, np.int32), cat(wgrids),
            cat(projectors), cat(aewaves), cat(pswaves), np.array(rmaxs, np.float64))


# --------------------------------------------------------------------------- #
# structures of the named configs
# --------------------------------------------------------------------------- #
def diamond_supercell(a0=5.43, n=3, vacancy=None):
    """n^3 conventional diamond cells (8 atoms each). Returns lattice, frac coords."""
    base = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                     [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    cells = np.array([[i, j, k] for i in range(n) for j in range(n) for k in range(n)])
    frac = ((cells[:, None, :] + base[None, :, :]) / n).reshape(-1, 3)
    if vacancy is not None:
        frac = np.delete(frac, vacancy, axis=0)
    return np.eye(3) * a0 * n, frac


def wurtzite_supercell(reps=(8, 4, 2), a=3.19, b=5.525, c=5.185):
    """Orthorhombic 8-atom wurtzite cell repeated; returns lattice, frac, labels (0=Ga,1=N)."""
    u = 0.377
    ga = np.array([[0, 0, 0], [.5, .5, 0], [0, 1 / 3, .5], [.5, 5 / 6, .5]])
    n_ = ga + np.array([0, 0, u])
    base = np.concatenate([ga, n_])
    lab = np.array([0] * 4 + [1] * 4, dtype=np.int32)
    r = np.array(reps)
    cells = np.array([[i, j, k] for i in range(r[0]) for j in range(r[1]) for k in range(r[2])])
    frac = ((cells[:, None, :] + base[None, :, :]) / r).reshape(-1, 3) % 1.0
    labels = np.tile(lab, len(cells))
    return np.diag([a * r[0], b * r[1], c * r[2]]), frac, labels


def grid_encut(dim, lattice):
    """wavefunction.py:411 (constant 0.262, not CCONST)."""
    abc = np.linalg.norm(np.asarray(lattice).reshape(3, 3), axis=1)
    return float(np.max((np.pi * np.asarray(dim) / abc) ** 2 / 0.262))
