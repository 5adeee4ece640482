"""BASELINE config 4 shape: noncollinear (spinor) wavefunction of a 128-site cell - setup_projections
(two transforms and two sets of projector overlaps per band) and ncl_realspace_state, timed on one B200.
Usage: python scripts/config4_ncl.py [nband]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pawpyseed_b200 import _lib, pawpyc, synth  # noqa: E402

nband = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lat, frac, lab = synth.wurtzite_supercell((4, 2, 2))            # 128 sites, 12.76 x 11.05 x 10.37 A
encut = 400.0
kpts = np.array([[0.0, 0.0, 0.0], [0.5, 0.0, 0.0]])
kws = np.array([0.5, 0.5])
gv = [synth.enumerate_gvectors(lat, encut, k) for k in kpts]
dim = synth.fft_grid_for(gv)
img = synth.wavecar_image(lat, encut, kpts, 1, nband, synth.random_coeffs(4, nband), ncl=True, gvecs=gv)
pps = synth.synthetic_pps(["Ga", "N"])
print("sites %d, npw %s (x2 spinor), grid %s, bands %d, image %.2f GB" % (
    len(lab), [len(g) for g in gv], dim, nband, img.nbytes / 1e9), flush=True)
wf = pawpyc.CNCLWavefunction(pawpyc.PWFPointer.from_arrays(img, kpts, kws))
assert wf.ncl
ge = synth.grid_encut(dim, lat)
for it in range(3):
    _lib.reset_timers()
    t = time.perf_counter()
    wf.projector_owner = 0
    wf._c_projector_setup(len(pps), len(lab), ge, lab.astype(np.int32), frac, dim, pps)
    P = wf._get_projections(0, 0, 1)       # forces completion (up-spinor projections)
    dt = time.perf_counter() - t
    tm = _lib.timers()
    nfft = 2 * nband * len(kpts)
    print("setup_projections pass %d: %.1f ms wall; %d spinor transforms + projections: fft %.1f ms, project %.1f ms, "
          "tables %.1f ms -> %.1f us per spinor component" % (it, dt * 1e3, nfft, tm["fft_ms"], tm["project_ms"],
                                                             tm["table_ms"], 1e3 * (tm["fft_ms"] + tm["project_ms"]) / nfft),
          flush=True)
t = time.perf_counter()
up, dn = wf._get_realspace_state(0, 1, 0)
first = time.perf_counter() - t
t = time.perf_counter()
for b in range(1, 9):
    up, dn = wf._get_realspace_state(b, 1, 0)
dt = (time.perf_counter() - t) / 8
print("ncl_realspace_state: first call %.1f ms (builds the AE partial-wave tables), then %.2f ms per band "
      "(two %s AE grids to the host)" % (first * 1e3, dt * 1e3, tuple(int(x) for x in dim)))
print("norm check:", float((np.abs(up) ** 2 + np.abs(dn) ** 2).sum() * abs(np.linalg.det(lat)) / up.size))
