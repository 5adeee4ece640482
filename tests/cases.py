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
