"""The pruned band-interleaved FFT + interleaved projection against the oracle on awkward grids, the generic
(cuFFT) fallback for unsupported sizes, sharded and asynchronous ingest."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from oracle import paw_numpy as pn
from pawpyseed_b200 import _lib, pawpyc

pytestmark = pytest.mark.gpu
TOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


def gpu(c):
    wf = (pawpyc.CNCLWavefunction if c["ncl"] else pawpyc.CWavefunction)(
        pawpyc.PWFPointer.from_arrays(c["image"], c["kpts"], c["kws"]))
    wf._c_projector_setup(len(c["pps"]), len(c["labels"]), c["grid_encut"], c["labels"], c["coords"], c["dim"],
                          c["pps"])
    return wf


def oracle(c):
    w = pn.Wavefunction.from_image(c["image"], c["kws"])
    w.setup_projections(c["pps"], c["labels"], c["coords"], c["dim"], c["grid_encut"])
    return w


@pytest.mark.parametrize("dim", [
    (18, 20, 24),    # 3x6, 4x5, 4x6            (small radices)
    (28, 30, 32),    # 4x7, 5x6, 4x8            (radix 7, 8)
    (45, 48, 56),    # 5x9, 3x16|6x8, 7x8       (radix 9, 16)
    (60, 42, 75),    # 6x10, 6x7, 5x15          (radix 10, 15) - non-cubic
    (162, 20, 36),   # 9x18, 4x5, 6x6           (radix 18: the 320-thread class)
    (20, 200, 18),   # 4x5, 10x20, 3x6          (radix 20)
    (125, 18, 20),   # 5x5x5: three-factor x axis (no split into two radices <= 20)
    (20, 147, 18),   # 7x7x3: three-factor y axis
    (18, 20, 243),   # 9x9x3: three-factor z axis (the pass that reads the interleaved coefficients)
    (384, 18, 420),  # 8x8x6 and 10x7x6: the largest three-factor lines (512-thread CTAs, 223 KB of shared memory)
    (22, 26, 20),    # 2x11 / 2x13: unsupported radices -> generic cuFFT path
    (19, 20, 20),    # prime size -> generic path
])
def test_projections_on_awkward_grids(dim):
    c = cases.small_case(seed=5, nband=5, encut=120.0, dim=dim)
    _lib.reset_timers()
    wf, o = gpu(c), oracle(c)
    got = np.array([[wf._get_projections(b, k) for b in range(5)] for k in range(4)])
    assert rel(got, np.array(o.P)) < TOL
    # which transform ran: the hand-written pruned passes never scatter into contiguous boxes
    t = _lib.timers()
    generic = dim in ((22, 26, 20), (19, 20, 20))
    assert (t["boxes_scattered"] > 0) == generic and t["boxes_fft"] > 0
    assert rel(wf._get_realspace_state(2, 1, 0), o.realspace_state(2, 1)) < TOL
    # density on the same grid: pruned transform + interleaved accumulation where the grid factors, cuFFT otherwise
    wf.fdimv = np.array(dim, np.int32)
    wf.fgridsize = int(np.prod(dim))
    assert rel(wf._get_realspace_density(), o.chg_density(np.array(dim))) < TOL


def test_17_bands_group_tail_and_offsite_reuse_of_resident_boxes():
    # 17 bands = one full interleave group + a one-band tail group; overlap_setup_real re-projects from the
    # boxes kept in HBM (no second transform)
    cR = cases.small_case(seed=21, nband=17)
    cS = cases.small_case(seed=22, nband=17, perturb=0.02)
    R, S, oR, oS = gpu(cR), gpu(cS), oracle(cR), oracle(cS)
    cat = [[0, 1], [0, 1], [2, 3], [2, 3], [2, 3], [2, 3]]
    pr = pawpyc.CProjector(S, R)
    _lib.reset_timers()
    pr._setup_overlap(cat, False)
    t = _lib.timers()
    assert t["boxes_fft"] == 0 and t["slots_projected"] == 2 * 17 * 4      # W_S and W_R, 4 kappa, no FFT
    want = np.array([pn.Projector(oS, oR, cat).single_band_projection(b) for b in range(17)])
    assert rel(pr._projection_matrix(), want.reshape(17, 17, 4).transpose(2, 0, 1)) < TOL


def test_generic_path_env_switch_matches():
    """PAWB200_FFT=cufft forces the scatter + cuFFT + contiguous projection path; same numbers."""
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import cases; from pawpyseed_b200 import pawpyc\n"
        "c = cases.small_case(seed=5, nband=6)\n"
        "wf = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(c['image'], c['kpts'], c['kws']))\n"
        "wf._c_projector_setup(len(c['pps']), len(c['labels']), c['grid_encut'], c['labels'], c['coords'], c['dim'], c['pps'])\n"
        "np.save(sys.argv[1], np.array([[wf._get_projections(b, k) for b in range(6)] for k in range(4)]))\n"
    ) % (ROOT, os.path.join(ROOT, "tests"))
    outs = []
    for mode in ("pruned", "cufft"):
        path = os.path.join("/tmp", "pawb200_proj_%s.npy" % mode)
        env = dict(os.environ)
        if mode == "cufft":
            env["PAWB200_FFT"] = "cufft"
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env)
        outs.append(np.load(path))
    assert rel(outs[0], outs[1]) < 1e-12
    c = cases.small_case(seed=5, nband=6)
    assert rel(outs[0], np.array(oracle(c).P)) < TOL


def test_sharded_and_async_ingest():
    c = cases.small_case(seed=9, nband=6)
    o = oracle(c)
    L = _lib.lib()
    try:
        L.pawb200_set_async_ingest(1)       # the image stays alive in `c` for the whole test
        parts = []
        for rank in range(2):
            L.pawb200_set_read_shard(rank, 2)
            wf = gpu(c)
            pr = pawpyc.CProjector(wf, wf)
            pr._setup_overlap([[0, 1, 2, 3], [0, 1, 2, 3], [], [], [], []], False)
            m = pr._projection_matrix()
            own = [k for k in range(4) if k % 2 == rank]
            other = [k for k in range(4) if k % 2 != rank]
            assert np.all(m[other] == 0)                     # blocks of the other rank come back as zeros
            with pytest.raises(_lib.PAWpyError):
                wf._get_projections(0, other[0])             # not resident here
            parts.append(m)
    finally:
        L.pawb200_set_read_shard(0, 1)
        L.pawb200_set_async_ingest(0)
    full = parts[0] + parts[1]                                # disjoint blocks: SUM == all-gather
    opr = pn.Projector(o, o, [[0, 1, 2, 3], [0, 1, 2, 3], [], [], [], []])
    want = np.array([opr.single_band_projection(b) for b in range(6)]).reshape(6, 6, 4).transpose(2, 0, 1)
    assert rel(full, want) < TOL


def test_async_pair_block_ordered_ingest_matches_oracle():
    """Two different wavefunctions read asynchronously: each issues its first (k,spin) block at read time, the other
    three are deferred and go out in block order across both when the second one is read (flush_pending_ingest);
    the PAW-corrected matrix equals the oracle's, and a second projector on the same pair (nothing pending, nothing
    prelaunched) gives bit-identical numbers."""
    cR, cS = cases.small_case(seed=31, nband=12), cases.small_case(seed=32, nband=12, perturb=0.02)
    oR, oS = oracle(cR), oracle(cS)
    cat = [[0, 1], [0, 1], [2, 3], [2, 3], [2, 3], [2, 3]]
    L = _lib.lib()
    try:
        L.pawb200_set_async_ingest(1)       # the images stay alive in cR / cS
        R, S = gpu(cR), gpu(cS)
        pr = pawpyc.CProjector(S, R)
        pr._setup_overlap(cat, False)
        got = pr._projection_matrix()
        pr2 = pawpyc.CProjector(S, R)
        pr2._setup_overlap(cat, False)
        again = pr2._projection_matrix()
    finally:
        L.pawb200_set_async_ingest(0)
    want = np.array([pn.Projector(oS, oR, cat).single_band_projection(b) for b in range(12)])
    assert rel(got, want.reshape(12, 12, 4).transpose(2, 0, 1)) < TOL
    assert np.array_equal(got, again)


def test_chunked_side_stream_gemm_matches_single_gemm(tmp_path):
    """While wf coefficients are still arriving the pseudo-overlap GEMM runs per ingest chunk on its own stream
    (PAWB200_GEMM_CHUNKED=1 forces that path): 300 bands = five 64-band chunks; same matrix as the one-shot GEMM.
    Both runs ingest asynchronously, so the second (k,spin) block of each wavefunction is a deferred, block-ordered
    copy; the forced run also takes the prelaunch path (overlap_setup_real queues the GEMMs of both blocks on the GEMM
    stream, overlap_matrix joins and adds the augmentation)."""
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import cases; from pawpyseed_b200 import pawpyc, _lib\n"
        "_lib.lib().pawb200_set_async_ingest(1)\n"
        "cR, cS = cases.small_case(seed=41, nband=300, nspin=1), cases.small_case(seed=42, nband=300, nspin=1, perturb=0.02)\n"
        "ws = []\n"
        "for c in (cR, cS):\n"
        "    w = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(c['image'], c['kpts'], c['kws']))\n"
        "    w._c_projector_setup(len(c['pps']), len(c['labels']), c['grid_encut'], c['labels'], c['coords'], c['dim'], c['pps'])\n"
        "    ws.append(w)\n"
        "pr = pawpyc.CProjector(ws[1], ws[0])\n"
        "pr._setup_overlap([[0, 1, 2], [0, 1, 2], [3], [3], [3], [3]], False)\n"
        "np.save(sys.argv[1], pr._projection_matrix())\n"
    ) % (ROOT, os.path.join(ROOT, "tests"))
    outs = []
    for mode in ("single", "chunked"):
        env = dict(os.environ)
        env.pop("PAWB200_GEMM_CHUNKED", None)
        if mode == "chunked":
            env["PAWB200_GEMM_CHUNKED"] = "1"
        out = str(tmp_path / (mode + ".npy"))
        subprocess.run([sys.executable, "-c", code, out], check=True, env=env, timeout=600)
        outs.append(np.load(out))
    assert outs[0].shape == (2, 300, 300)
    assert rel(outs[1], outs[0]) < 1e-13


@pytest.mark.parametrize("world,nband", [(2, 12), (3, 10)])
def test_band_sharded_projection_equals_unsharded(world, nband):
    """SURVEY 8e level 2: the bands of one (k,spin) block split over `world` ranks.  The ranks are emulated in one
    process (pawb200_set_band_shard before each read); the NCCL all-gather of the basis rows is replaced by explicit
    copies between the ranks' device buffers (the same buffers distributed.gather_band_blocks exchanges).  Every
    rank's matrix holds only the rows of its wf bands; together they equal the unsharded matrix."""
    from pawpyseed_b200 import distributed as pd
    L = _lib.lib()
    cR, cS = cases.small_case(seed=7, nband=nband), cases.small_case(seed=11, nband=nband, perturb=0.03)
    cat = [[0, 1], [0, 1], [2, 3], [2, 3], [2, 3], [2, 3]]
    L.pawb200_set_band_shard(0, 1)
    R0, S0 = gpu(cR), gpu(cS)
    pr0 = pawpyc.CProjector(S0, R0)
    pr0._setup_overlap(cat, False)
    want = pr0._projection_matrix()
    try:
        ranks = []
        for r in range(world):
            L.pawb200_set_band_shard(r, world)
            R, S = gpu(cR), gpu(cS)
            pr = pawpyc.CProjector(S, R)
            pr._setup_overlap(cat, False)
            ranks.append((R, S, pr))
        NK = 4
        per = -(-nband // world)
        # "all-gather" of the basis rows: coefficients, projections, wave projections
        for which in (0, 1, 2):
            for kappa in range(NK):
                views = [pd._DeviceRows(R, which, kappa) for R, _, _ in ranks]
                assert all(v.rows == per * world * (1 if which == 0 else 1) for v in views)
                for dst in views:
                    for src in views:
                        if src is not dst:
                            dst.tensor[src.lo:src.hi].copy_(src.tensor[src.lo:src.hi])
        total = np.zeros_like(want)
        for r, (R, S, pr) in enumerate(ranks):
            got = pr._projection_matrix()
            lo, hi = min(nband, r * per), min(nband, (r + 1) * per)
            outside = np.ones(nband, bool)
            outside[lo:hi] = False
            assert np.all(got[:, outside, :] == 0)          # rows of other ranks' bands stay zero
            total += got
        assert rel(total, want) < 1e-12
    finally:
        L.pawb200_set_band_shard(0, 1)
