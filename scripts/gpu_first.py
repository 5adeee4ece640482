import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
print(torch.cuda.get_device_name(0))
from pawpyseed_b200 import pawpyc, _lib, synth
from oracle import paw_numpy as pn
import cases
c = cases.small_case()
c2 = cases.small_case(seed=11, perturb=0.0)
def rel(a,b): return np.abs(a-b).max()/max(np.abs(b).max(),1e-300)
def mk(c):
    pwf = pawpyc.PWFPointer.from_arrays(c['image'], c['kpts'], c['kws'])
    wf = pawpyc.CWavefunction(pwf)
    wf._c_projector_setup(len(c['pps']), len(c['labels']), c['grid_encut'], c['labels'], c['coords'], c['dim'], c['pps'])
    return wf
t=time.time(); wf = mk(c); wf2 = mk(c2); print('gpu setup', time.time()-t)
nwf = pn.Wavefunction.from_image(c['image'], c['kws']); nwf.setup_projections(c['pps'], c['labels'], c['coords'], c['dim'], c['grid_encut'])
nwf2 = pn.Wavefunction.from_image(c2['image'], c2['kws']); nwf2.setup_projections(c2['pps'], c2['labels'], c2['coords'], c2['dim'], c2['grid_encut'])
print('chan idx', np.array_equal(wf._get_channel_index(), nwf.chan_index))
for s in range(4): print('site idx', s, np.array_equal(wf._get_site_indices(s), nwf.sites[s]['indices']))
worst = 0
for kap in range(4):
    for b in range(c['nband']):
        worst = max(worst, rel(wf._get_projections(b, kap), nwf.P[kap][b]))
print('projection rel err', worst)
# overlaps
for cat in ([[0,1,2,3],[0,1,2,3],[],[],[],[]], [[],[],[0,1,2,3],[0,1,2,3],[0,1,2,3],[0,1,2,3]], [[0,1],[0,1],[2,3],[2,3],[2,3,2],[2,3,3]]):
    pr = pawpyc.CProjector(wf2, wf); pr._setup_overlap(cat, False)
    npr = pn.Projector(nwf2, nwf, cat)
    for b in (0, 5):
        for flip in (False, True):
            res = wf2.pseudoprojection(b, wf, flip); ps = res.copy()
            pr._add_augmentation_terms(res, b, flip)
            nps = nwf2.pseudoprojection(b, nwf, flip); naug = npr.compensation_terms(b, flip)
            print(cat[0], b, flip, 'pseudo rel', rel(ps, nps), 'aug rel', rel(res-ps, naug), np.abs(naug).max())
    M = pr._projection_matrix()
    full = np.array([npr.single_band_projection(b) for b in range(c['nband'])])  # [bS][bR*NK+k]
    NK=4
    ref = full.reshape(c['nband'], c['nband'], NK).transpose(2,0,1)
    print('matrix rel', rel(M, ref))
x = wf._get_realspace_state(3, 1, 1); nx = nwf.realspace_state(3, 1+2)
print('realspace rel', rel(x, nx))
d = wf._get_realspace_density(); nd = nwf.chg_density(c['dim']*2)
print('density rel', rel(d, nd))
print(_lib.timers())
