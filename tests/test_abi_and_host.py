"""C-ABI surface and the host-side (C++) PAW setup, no GPU needed:
 - libpawb200.so loads and exports every symbol include/pawpyseed_b200.h declares;
 - compute entry points fail loudly (no CPU fallback) when no B200 is present;
 - the reference's own known-answer tests for the utilities (test_core.py:54-146);
 - host C++ vs the oracle for splines, NumSBT, off-site overlaps."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.special as sp
from scipy.interpolate import CubicSpline

import cases
from oracle import paw_numpy as pn
from pawpyseed_b200 import _lib, pawpyc, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "pawpyseed_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pawb200_\w+)\s*\(", hdr)))
    assert len(declared) > 40
    L = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.exported_symbols()) <= set(declared)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    assert L.pawb200_device_check() != 0
    assert b"no CPU fallback" in L.pawb200_last_error()
    c = cases.small_case(nband=2)
    with pytest.raises(_lib.PAWpyError):
        pawpyc.PWFPointer.from_arrays(c["image"], c["kpts"], c["kws"])


def test_legendre_vs_scipy():
    # reference KAT: test_core.py:54-66
    xs = np.linspace(-0.95, 0.95, 23)
    for l in range(4):
        for m in range(0, l + 1):
            for x in xs:
                assert abs(pawpyc.legendre(l, m, x) - sp.lpmv(m, l, x)) < 1e-10


def test_ylm_vs_scipy():
    # reference KAT: test_core.py:68-90
    rng = np.random.default_rng(0)
    for l in range(4):
        for m in range(-l, l + 1):
            for _ in range(5):
                th, ph = rng.uniform(0, np.pi), rng.uniform(0, 2 * np.pi)
                want = sp.sph_harm_y(l, m, th, ph)
                assert abs(pawpyc.Ylm(l, m, th, ph) - want) < 1e-12
                assert abs(pawpyc.Ylm2(l, m, np.cos(th), ph) - want) < 1e-12


def test_frac_cart_roundtrip():
    lat = cases.GA4_LATTICE
    rec = pn.reciprocal_lattice(lat)
    c = np.array([0.3, -0.2, 0.9])
    v = c.copy()
    pawpyc.frac_to_cartesian(v, lat)
    assert np.allclose(v, c @ lat, atol=1e-14)
    pawpyc.cartesian_to_frac(v, rec)
    assert np.allclose(v, c, atol=1e-13)


def test_spline_vs_cubicspline_and_oracle():
    # reference KAT: test_core.py:106-119 (2 decimals against scipy); oracle to 1e-13
    x = np.linspace(0, 1.5, 100, endpoint=False)
    y = np.exp(-x * x) * np.cos(3 * x)
    tst = np.linspace(0.01, 1.4, 57)
    res = np.zeros_like(tst)
    pawpyc.interpolate(res, tst, x, y, 1.5, 100, len(tst))
    assert np.abs(res - CubicSpline(x, y, bc_type="natural")(tst)).max() < 1e-2
    want = pn.proj_interpolate(tst, 1.5, x, y, pn.spline_coeff(x, y))
    assert np.abs(res - want).max() < 1e-13
    assert np.abs(pawpyc._spline(x, y).reshape(3, -1) - pn.spline_coeff(x, y)).max() < 1e-12


def test_sbt_vs_oracle_and_trapezoid():
    # reference KAT: test_core.py:130-146 (3 decimals against a trapezoid transform)
    pp = synth.synthetic_pps(["Ga"])[0]
    r = pp.grid
    for n, l in enumerate(pp.ls):
        f = pp.aewaves[n] - pp.pswaves[n]
        k, fk = pawpyc.spherical_bessel_transform(1e7, l, r, f)
        o = pn.SBT(1e7, 0, l, r)
        want = o.forward(f, l)
        assert np.abs(k - o.kgrid).max() / k.max() < 1e-13
        assert np.abs(fk - want).max() / np.abs(want).max() < 1e-10
        for i in (150, 180, 200):
            direct = np.trapezoid(sp.spherical_jn(l, r * k[i]) * f * r, r)
            assert abs(fk[i] - direct) < 1e-3


def test_offsite_overlap_vs_oracle():
    pps = synth.synthetic_pps(["Ga", "N"])
    pp = pn.build_ppots(pps, 400.0)
    d = np.array([0.7, -1.1, 1.3])
    worst = 0.0
    for (pa, pb) in ((pp[0], pp[1]), (pp[1], pp[0]), (pp[0], pp[0])):
        for (j, l1, m1) in pa.chan:
            for (k, l2, m2) in pb.chan:
                want = pn.reciprocal_offsite_wave_overlap(d, pa.kwave_grid, pa.kwave[j], pa.kwave_spline[j],
                                                          pb.kwave_grid, pb.kwave[k], pb.kwave_spline[k],
                                                          l1, m1, l2, m2)
                o = np.zeros(2)
                s1 = np.ascontiguousarray(pa.kwave_spline[j].reshape(-1))
                s2 = np.ascontiguousarray(pb.kwave_spline[k].reshape(-1))
                _lib.lib().pawb200_reciprocal_offsite_wave_overlap(
                    _lib.dp(d), _lib.dp(pa.kwave_grid), _lib.dp(pa.kwave[j]), _lib.dp(s1), len(pa.kwave_grid),
                    _lib.dp(pb.kwave_grid), _lib.dp(pb.kwave[k]), _lib.dp(s2), len(pb.kwave_grid),
                    l1, m1, l2, m2, _lib.dp(o))
                worst = max(worst, abs(complex(o[0], o[1]) - want))
    assert worst < 1e-15


def test_synth_gvectors_match_reference_golden():
    g = np.load(os.path.join(cases.GOLDEN, "ga4.npz"), allow_pickle=True)
    gv = synth.enumerate_gvectors(cases.GA4_LATTICE, 320.0, g["kpts"][0])
    assert np.array_equal(gv, g["gvecs_k0"])


def test_api_error_behaviour_mirrors_reference():
    # pawpyc.pyx:296-297: NULL pointer -> Exception; projector.py:270-271 index checks live in Python
    with pytest.raises(Exception):
        pawpyc.PseudoWavefunction(pawpyc.PWFPointer())


def test_write_volumetric_matches_reference_text(tmp_path):
    # density.c:461-477: "%E   " five per line, x fastest / z slowest; golden text written by the reference C
    g = np.load(os.path.join(cases.GOLDEN, "volumetric.npz"), allow_pickle=True)
    x = np.ascontiguousarray(g["x"])
    dim = np.ascontiguousarray(g["dim"], dtype=np.int32)
    fn = str(tmp_path / "v.txt")
    _lib.lib().pawb200_write_volumetric(fn.encode(), _lib.dp(x), _lib.ip(dim), float(g["scale"]))
    _lib.check()
    assert open(fn).read() == str(g["text"])


def test_momentum_grid_helpers_vs_reference_golden():
    # host-side pieces of MomentumMatrix (momentum.c:365-400, 547-574) need no GPU: compare with the reference's
    # own outputs stored in tests/golden/momentum.npz
    g = np.load(os.path.join(cases.GOLDEN, "momentum.npz"))
    L = _lib.lib()
    ggrid = np.ascontiguousarray(g["ggrid"], dtype=np.int32)
    n = len(ggrid) // 3
    gb, gd = np.zeros(6, np.int32), np.zeros(3, np.int32)
    L.pawb200_grid_bounds(_lib.ip(gb), _lib.ip(gd), _lib.ip(ggrid), n)
    assert np.array_equal(gb, g["gbounds"]) and np.array_equal(gd, g["gdim"])
    grid3d = -np.ones(int(np.prod(gd)), dtype=np.int32)
    L.pawb200_list_to_grid_map(_lib.ip(grid3d), _lib.ip(gb), _lib.ip(gd), _lib.ip(ggrid), n)
    assert np.array_equal(grid3d, g["grid3d"])
    # quick_overlap of two stored AE expansions against a direct evaluation
    v1, v2 = np.ascontiguousarray(g["full_b1k0s0"]), np.ascontiguousarray(g["full_b3k1s1"])
    dG = np.array([1, 0, -1], np.int32)
    out = np.zeros(2)
    L.pawb200_quick_overlap(_lib.ip(dG), v1.ctypes.data_as(_lib.c_dbl_p), v2.ctypes.data_as(_lib.c_dbl_p), n,
                            _lib.ip(ggrid), _lib.ip(grid3d), _lib.ip(gb), _lib.ip(gd), _lib.dp(out))
    look = {tuple(v): i for i, v in enumerate(ggrid.reshape(-1, 3).tolist())}
    want = 0j
    for w, v in enumerate(ggrid.reshape(-1, 3).tolist()):
        q = (v[0] + 1, v[1], v[2] - 1)
        if q in look and gb[0] <= q[0] <= gb[1] and gb[2] <= q[1] <= gb[3] and gb[4] <= q[2] <= gb[5]:
            want += np.conj(v1[look[q]]) * v2[w]
    assert abs(complex(out[0], out[1]) - want) < 1e-14
