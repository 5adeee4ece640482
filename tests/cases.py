"""Shared test inputs: small synthetic cases + the bundled Ga4 fixture (copied under tests/golden)."""
import gzip
import os

import numpy as np

from pawpyseed_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GA4_COORDS = np.array([
    [0.1560427174376784, 0.8439572825631387, 0.9166540580595424],
    [0.3439572825623216, 0.6560427174368613, 0.4166540580595424],
    [0.8439572825623216, 0.1560427174368613, 0.0833459419404576],
    [0.6560427174376784, 0.3439572825631387, 0.5833459419404576]])
GA4_LATTICE = np.array([[4.3515518806929165, 0.0106499131953374, 0.0],
                        [-2.0970141186059448, 3.8129574195815592, 0.0],
                        [0.0, 0.0, 4.4060081425154092]])


def small_case(seed=7, nband=12, nspin=2, encut=200.0, kpts=((0.25, 0.25, 0.25), (-0.25, 0.25, 0.25)),
               elements=("Ga", "N"), labels=(0, 1, 0, 1), perturb=0.0, lattice=None, coords=None,
               ncl=False, dim=None):
    """A two-element, spin-polarised, two-k synthetic cell small enough for the numpy oracle."""
    lattice = GA4_LATTICE * 1.15 if lattice is None else np.asarray(lattice, dtype=float)
    coords = GA4_COORDS.copy() if coords is None else np.asarray(coords, dtype=float)
    rng = np.random.default_rng(seed + 1000)
    coords = coords + perturb * rng.standard_normal(coords.shape)
    kpts = np.asarray(kpts, dtype=float)
    gv = [synth.enumerate_gvectors(lattice, encut, k) for k in kpts]
    if dim is None:
        dim = synth.fft_grid_for(gv)
    img = synth.wavecar_image(lattice, encut, kpts, nspin, nband, synth.random_coeffs(seed, nband),
                              ncl=ncl, gvecs=gv)
    pps = synth.synthetic_pps(list(elements))
    kws = np.full(len(kpts), 1.0 / len(kpts))
    return dict(image=img, kws=kws, kpts=kpts, lattice=lattice, coords=coords,
                labels=np.asarray(labels, dtype=np.int32), dim=np.asarray(dim, dtype=np.int32), pps=pps,
                grid_encut=synth.grid_encut(dim, lattice), nband=nband, nspin=nspin, ncl=ncl)


def desymm_case(seed=21, nband=5, nspin=2, encut=180.0):
    """Cubic two-element cell with an irreducible k-set and the operations (reciprocal fractional
    coordinates) that expand it: rotations, inversion, time reversal, fractional translations."""
    a = 4.6
    lattice = np.eye(3) * a
    kpts = np.array([[0.25, 0.25, 0.25], [0.0, 0.25, 0.5], [0.0, 0.0, 0.0]])
    c = small_case(seed=seed, nband=nband, nspin=nspin, encut=encut, kpts=kpts, lattice=lattice,
                   coords=[[0.0, 0.0, 0.0], [0.5, 0.5, 0.5], [0.5, 0.5, 0.0], [0.0, 0.0, 0.5]])
    E = np.eye(3)
    C4z = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    C3 = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    Mx = np.diag([-1.0, 1.0, 1.0])
    # (source k, operator, translation, time reversal); the k = (0,1/4,1/2) images exercise the -1/2 -> +1/2 fold
    spec = [(0, E, (0, 0, 0), 0), (0, C4z, (0.5, 0.0, 0.0), 0), (0, -E, (0, 0, 0), 0), (0, Mx, (0.0, 0.5, 0.25), 1),
            (1, E, (0, 0, 0), 0), (1, C3, (0.5, 0.5, 0.5), 0), (1, C4z @ C3, (0.25, 0.0, 0.5), 1),
            (1, -E, (0.0, 0.0, 0.5), 0), (2, E, (0, 0, 0), 0), (2, C3, (0.5, 0.5, 0.0), 1)]
    c["maps"] = np.array([s[0] for s in spec], dtype=np.int32)
    c["ops"] = np.array([s[1] for s in spec], dtype=np.float64)
    c["drs"] = np.array([s[2] for s in spec], dtype=np.float64)
    c["trs"] = np.array([s[3] for s in spec], dtype=np.int32)
    c["new_kws"] = np.full(len(spec), 1.0 / len(spec))
    return c


def cubic_point_group():
    """The 48 signed permutation matrices (O_h in the fractional coordinates of a cubic cell)."""
    import itertools
    mats = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1.0, -1.0), repeat=3):
            m = np.zeros((3, 3))
            for r in range(3):
                m[r, perm[r]] = signs[r]
            mats.append(m)
    return mats
