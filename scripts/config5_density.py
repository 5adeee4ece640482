"""BASELINE config 5 shape: AE charge density on a 400^3 grid (write_density path), timed on one B200.
Consistency check: density from ae_chg_density == sum_b w |realspace_state_b|^2 computed band by band."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pawpyseed_b200 import pawpyc, synth, _lib

nband = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dim = np.array([int(sys.argv[2])] * 3 if len(sys.argv) > 2 else [200, 200, 200], np.int32)   # fine grid = 2*dim
lat, coords = synth.diamond_supercell(5.43, 2)           # 64 Si sites, a = 10.86 A
encut = 300.0
kpts = np.array([[0.0, 0.0, 0.0]]); kws = np.array([1.0])
gv = [synth.enumerate_gvectors(lat, encut, kpts[0])]
img = synth.wavecar_image(lat, encut, kpts, 1, nband, synth.random_coeffs(5, nband), gvecs=gv)
pps = synth.synthetic_pps(["Si"])
wf = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(img, kpts, kws))
labels = np.zeros(len(coords), np.int32)
t = time.time()
wf._c_projector_setup(1, len(coords), synth.grid_encut(dim, lat), labels, coords, dim, pps)
print("setup_projections %.3f s (npw %d, grid %s)" % (time.time() - t, len(gv[0]), dim))
_lib.reset_timers()
t = time.time()
rho = wf._get_realspace_density()
t1 = time.time() - t
nocc = (nband + 1) // 2
print("ae_chg_density on %s: %.3f s first call (tables + %d occupied bands)" % (tuple(wf.fdimv), t1, nocc))
t = time.time()
rho2 = wf._get_realspace_density()
t2 = time.time() - t
print("second call %.3f s -> %.1f ms per band, %.2f GB grid per band" % (t2, 1e3 * t2 / nocc, 16 * rho.size / 1e9))
print("timers", {k: round(v, 2) for k, v in _lib.timers().items()})
assert np.array_equal(rho, rho2)
# band-by-band consistency
acc = np.zeros_like(rho)
for b in range(min(nocc, 4) if "--quick" in sys.argv else nocc):
    acc += wf._get_realspace_state_density(b, 0, 0) * 2.0     # weight 1 * occ 1 * spin_mult 2
if "--quick" not in sys.argv:
    print("consistency rel err", np.abs(acc - rho).max() / np.abs(rho).max())
vol = abs(np.linalg.det(lat))
print("integral rho dV =", rho.sum() * vol / rho.size, "(2 x %d occupied bands, pseudo norm + PAW correction)" % nocc)
