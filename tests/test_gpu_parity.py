"""GPU parity tests: the sm_100a path, called through the C ABI (via the pawpyc mirror), against
 (1) the committed golden vectors from the unmodified reference C, and
 (2) the numpy oracle on seeded synthetic inputs.
Bars: FP64 quantities 1e-10 relative (north_star); index arrays bit-exact; the pseudo overlap
against the reference's single-precision value 5e-6 absolute (it is an FP32 accumulate there)."""
import os

import numpy as np
import pytest

import cases
from oracle import paw_numpy as pn
from pawpyseed_b200 import _lib, pawpyc, synth

pytestmark = pytest.mark.gpu
TOL = 1e-10
G = cases.GOLDEN


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


def gpu_wf(image, kpts, kws, pps, labels, coords, dim, grid_encut, ncl=False):
    cls = pawpyc.CNCLWavefunction if ncl else pawpyc.CWavefunction
    wf = cls(pawpyc.PWFPointer.from_arrays(image, kpts, kws))
    wf._c_projector_setup(len(pps), len(labels), grid_encut, labels, coords, dim, pps)
    return wf


def oracle_wf(c):
    w = pn.Wavefunction.from_image(c["image"], c["kws"])
    w.setup_projections(c["pps"], c["labels"], c["coords"], c["dim"], c["grid_encut"])
    return w


def from_case(c):
    return gpu_wf(c["image"], c["kpts"], c["kws"], c["pps"], c["labels"], c["coords"], c["dim"],
                  c["grid_encut"], c["ncl"])


# ---------------------------------------------------------------------------------------------
# golden: the reference's own Ga4 fixtures (BASELINE config 1)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ga4():
    g = np.load(os.path.join(G, "ga4.npz"), allow_pickle=True)
    pps = synth.synthetic_pps(["Ga"])
    lab = np.zeros(4, np.int32)
    R = gpu_wf(g["image_R"], g["kpts"], g["kws"], pps, lab, cases.GA4_COORDS, g["dim"], float(g["grid_encut"]))
    S = gpu_wf(g["image_S"], g["kpts"], g["kws"], pps, lab, cases.GA4_COORDS, g["dim"], float(g["grid_encut"]))
    return g, R, S


def test_ga4_indices_bit_exact(ga4):
    g, R, _ = ga4
    assert np.array_equal(R._get_channel_index(), g["chan_index"])
    assert np.array_equal(R._get_site_indices(0), g["site_index_0"])
    assert [len(R._get_site_indices(s)) for s in range(4)] == list(g["site_npts"])


def test_ga4_projections(ga4):
    g, R, S = ga4
    for wf, key in ((R, "proj_R"), (S, "proj_S")):
        got = np.array([[wf._get_projections(b, k) for b in range(8)] for k in range(4)])
        assert rel(got, g[key]) < TOL


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_ga4_single_band_projection(ga4, ci):
    g, R, S = ga4
    pr = pawpyc.CProjector(S, R)
    pr._setup_overlap([list(x) for x in g["cats"][ci]], False)
    for flip in (0, 1):
        for bi, b in enumerate(g["bands"]):
            res = S.pseudoprojection(int(b), R, bool(flip))
            assert np.abs(res - g["pseudo_f%d" % flip][bi]).max() < 5e-6      # FP32 term of the reference
            ps = res.copy()
            pr._add_augmentation_terms(res, int(b), bool(flip))                 # accumulates (+=)
            assert rel(res - ps, g["aug_c%d_f%d" % (ci, flip)][bi]) < TOL


def test_ga4_realspace_and_density(ga4):
    g, R, _ = ga4
    assert rel(R._get_realspace_state(1, 1, 1), g["state_b1_k1"]) < TOL
    assert rel(R._get_realspace_state(1, 1, 1, remove_phase=True), g["state_b1_k1_nophase"]) < TOL
    assert rel(R._get_realspace_density(), g["density"]) < TOL


def test_ncl_fixture():
    g = np.load(os.path.join(G, "ncl.npz"), allow_pickle=True)
    N = gpu_wf(g["image"], g["kpts"], g["kws"], synth.synthetic_pps(["Ga"]), np.zeros(4, np.int32),
               cases.GA4_COORDS, g["dim"], float(g["grid_encut"]), ncl=True)
    assert N.ncl
    up = np.array([[N._get_projections(b, k, 1) for b in range(4)] for k in range(2)])
    dn = np.array([[N._get_projections(b, k, 2) for b in range(4)] for k in range(2)])
    assert rel(up, g["up"]) < TOL and rel(dn, g["down"]) < TOL
    s0, s1 = N._get_realspace_state(2, 1, 0)
    assert rel(np.stack([s0, s1]), g["state_b2_k1"]) < TOL
    assert rel(N._get_realspace_density(), g["density"]) < TOL


# ---------------------------------------------------------------------------------------------
# oracle: seeded synthetic cells
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gan():
    cR, cS = cases.small_case(seed=7), cases.small_case(seed=11, perturb=0.03)
    return cR, cS, from_case(cR), from_case(cS), oracle_wf(cR), oracle_wf(cS)


def test_synthetic_golden_two_element(gan):
    g = np.load(os.path.join(G, "synth_gan.npz"), allow_pickle=True)
    cR, cS = cases.small_case(seed=7, nband=6), cases.small_case(seed=11, nband=6, perturb=0.03)
    R, S = from_case(cR), from_case(cS)
    assert np.array_equal(R._get_channel_index(), g["chan_index"])
    got = np.array([[R._get_projections(b, k) for b in range(6)] for k in range(4)])
    assert rel(got, g["proj_R"]) < TOL
    pr = pawpyc.CProjector(S, R)
    pr._setup_overlap([list(x) for x in g["cats"][0]], False)
    for flip in (0, 1):
        for bi, b in enumerate(g["bands"]):
            res = np.zeros(6 * 4, complex)
            pr._add_augmentation_terms(res, int(b), bool(flip))
            assert rel(res, g["aug_c0_f%d" % flip][bi]) < TOL


@pytest.mark.parametrize("cat", [
    [[0, 1, 2, 3], [0, 1, 2, 3], [], [], [], []],                       # all matched (O_M only)
    [[], [], [0, 1, 2, 3], [0, 1, 2, 3], [0, 1, 2, 3], [0, 1, 2, 3]],    # all off-site (O_R, O_S, O_N)
    [[0, 1], [0, 1], [2, 3], [2, 3], [2, 3, 2], [2, 3, 3]],              # mixed, with a Ga-N pair
    [[0], [0], [1, 2, 3], [], [], []],                                   # N_R only (vacancy-like)
    [[], [], [], [], [], []],                                            # nothing: pure pseudo overlap
])
def test_overlap_matrix_vs_oracle(gan, cat):
    cR, cS, R, S, oR, oS = gan
    pr = pawpyc.CProjector(S, R)
    pr._setup_overlap(cat, False)
    opr = pn.Projector(oS, oR, cat)
    nb, NK = cR["nband"], 4
    for flip in (False, True):
        want = np.array([opr.single_band_projection(b, flip) for b in range(nb)]).reshape(nb, nb, NK)
        got = pr._projection_matrix(flip)                                 # [kappa][b_wf][b_basis]
        assert rel(got, want.transpose(2, 0, 1)) < TOL
        # per-band reference API (pseudoprojection + compensation_terms) agrees with the batched call
        for b in (0, nb - 1):
            res = S.pseudoprojection(b, R, flip)
            pr._add_augmentation_terms(res, b, flip)
            assert rel(res, want[b].reshape(-1)) < TOL


def test_hermiticity_and_self_projection(gan):
    cR, _, R, _, oR, _ = gan
    pr = pawpyc.CProjector(R, R)
    pr._setup_overlap([[0, 1, 2, 3], [0, 1, 2, 3], [], [], [], []], False)
    M = pr._projection_matrix()
    for k in range(4):
        assert np.abs(M[k] - M[k].conj().T).max() < 1e-12                # <i|O|j> = conj(<j|O|i>)


def test_ragged_band_counts_and_single_site():
    for nband in (1, 5, 33):
        c = cases.small_case(seed=3, nband=nband, elements=("N",), labels=(0,), coords=[[0.1, 0.2, 0.3]],
                             nspin=1, kpts=((0.0, 0.0, 0.0),))
        wf, o = from_case(c), oracle_wf(c)
        got = np.array([wf._get_projections(b, 0) for b in range(nband)])
        assert rel(got, np.array(o.P[0])) < TOL
        pr = pawpyc.CProjector(wf, wf)
        pr._setup_overlap([[0], [0], [], [], [], []], False)
        want = np.array([pn.Projector(o, o, [[0], [0], [], [], [], []]).single_band_projection(b)
                         for b in range(nband)])
        assert rel(pr._projection_matrix()[0], want) < TOL


def test_realspace_state_custom_grid_and_density(gan):
    cR, _, R, _, oR, _ = gan
    x = R._get_realspace_state(3, 1, 1)
    assert rel(x, oR.realspace_state(3, 1 + 2)) < TOL
    d = R._get_realspace_density()
    want = oR.chg_density(cR["dim"] * 2)
    assert rel(d, want) < TOL
    sd = R._get_realspace_state_density(2, 0, 1)
    xs = oR.realspace_state(2, 2, cR["dim"] * 2)
    assert rel(sd, np.abs(xs) ** 2) < TOL


def test_fft_check_known_answer():
    """The reference's own KAT (tests.c:28-80): FFT of a band equals the direct plane-wave sum to 1e-5,
    and fwd_fft3d(fft3d(C)) returns C to 1e-5."""
    g = np.load(os.path.join(G, "ga4.npz"), allow_pickle=True)
    o = pn.Wavefunction.from_image(g["image_R"], g["kws"])
    Gs, Cs = np.ascontiguousarray(o.Gs[0], dtype=np.int32), np.ascontiguousarray(o.Cs[0][0])
    dim = np.ascontiguousarray(g["dim"], dtype=np.int32)
    lat = np.ascontiguousarray(o.lattice.reshape(-1))
    L = _lib.lib()
    x = np.zeros(int(np.prod(dim)), np.complex128)
    k = np.zeros(3)
    L.pawb200_fft3d(x.ctypes.data_as(_lib.c_dbl_p), None, _lib.dp(lat), _lib.dp(k), _lib.ip(Gs.reshape(-1)),
                    Cs.ctypes.data, len(Cs), _lib.ip(dim))
    _lib.check()
    I, J, K = np.meshgrid(*[np.arange(n) / n for n in dim], indexing="ij")
    direct = np.zeros(tuple(dim), complex)
    for w in range(len(Cs)):
        direct += Cs[w] * np.exp(2j * np.pi * (I * Gs[w, 0] + J * Gs[w, 1] + K * Gs[w, 2]))
    direct *= pn.determinant(o.lattice) ** -0.5
    assert np.abs(x.reshape(tuple(dim)) - direct).max() < 1e-5
    assert rel(x.reshape(tuple(dim)), pn.fft3d(Gs, Cs, o.lattice, dim)) < 1e-13
    back = np.zeros(len(Cs), np.complex64)
    L.pawb200_fwd_fft3d(x.ctypes.data_as(_lib.c_dbl_p), None, _lib.dp(lat), _lib.dp(k), _lib.ip(Gs.reshape(-1)),
                        back.ctypes.data, len(Cs), _lib.ip(dim))
    _lib.check()
    assert np.abs(back - Cs).max() < 1e-5


def test_error_behaviour():
    c = cases.small_case(nband=4)
    wf = from_case(c)
    with pytest.raises(ValueError):
        wf._get_realspace_state(99, 0, 0)                # pawpyc.pyx:426-427
    with pytest.raises(ValueError):
        wf._get_realspace_state(0, 9, 0)
    res = np.zeros(4 * 4, complex)
    pr = pawpyc.CProjector(wf, wf)
    pr._setup_overlap([[0, 1, 2, 3], [0, 1, 2, 3], [], [], [], []], False)
    with pytest.raises(_lib.PAWpyError):
        pr._add_augmentation_terms(res, 99, False)       # the C ABI reports instead of reading out of bounds
    with pytest.raises(_lib.PAWpyError):                 # mismatched spin counts: refused
        c1 = cases.small_case(nband=4, nspin=1)
        pawpyc.CProjector(from_case(c1), wf)._projection_matrix()
    with pytest.raises(_lib.PAWpyError):
        pawpyc.PWFPointer.from_arrays(np.zeros(4096, np.uint8), c["kpts"], c["kws"])   # not a WAVECAR


def test_ga4_realspace_projection_method(ga4):
    # project_realspace_state (density.c:205-230), golden from the reference C
    g, R, S = ga4
    gp = np.load(os.path.join(G, "realspace_proj.npz"))
    pr = pawpyc.CProjector(S, R)
    got = pr._realspace_projection(int(gp["band"]), gp["dim"])
    assert rel(got, gp["res"]) < TOL


def test_desymmetrisation_vs_reference_and_oracle():
    # expand_symm_wf (utils.c:829-1098) remapped on the GPU.  The phase factors come from the
    # same cexpf and the complex64 product is unfused, so coefficients match the reference C bit for bit; the
    # FP64 projections of the expanded wavefunction then meet the 1e-10 bar.
    c = cases.desymm_case()
    g = np.load(os.path.join(G, "desymm.npz"), allow_pickle=True)
    src = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(c["image"], c["kpts"], c["kws"]))
    L = _lib.lib()
    pw = pawpyc.PWFPointer()
    pw.kpts, pw.weights, pw.band_props = g["kpts"], c["new_kws"], np.zeros(4)
    ops, drs = np.ascontiguousarray(c["ops"]).reshape(-1), np.ascontiguousarray(c["drs"]).reshape(-1)
    pw.ptr = L.pawb200_expand_symm_wf(src.wf_ptr, len(c["maps"]), _lib.ip(c["maps"]), _lib.dp(ops), _lib.dp(drs),
                                      _lib.dp(pw.weights), _lib.ip(c["trs"]))
    _lib.check()
    E = pawpyc.CWavefunction(pw)
    assert (E.nwk, E.nspin, E.nband) == (len(c["maps"]), 2, c["nband"])
    k3 = np.zeros(3)
    for kap in range(E.nwk * E.nspin):
        n = L.pawb200_get_kpoint(E.wf_ptr, kap, _lib.dp(k3), None)
        assert np.array_equal(k3, g["kpts"][kap % E.nwk]) and n == len(g["gvecs"][kap % E.nwk])
        for b in range(E.nband):
            assert np.array_equal(E._get_coefficients(b, kap), g["coeffs"][kap][b])
            assert L.pawb200_get_occ(E.wf_ptr, b, kap % E.nwk, kap // E.nwk) == g["occ"][kap][b]
    E._c_projector_setup(len(c["pps"]), len(c["labels"]), c["grid_encut"], c["labels"], c["coords"], c["dim"], c["pps"])
    P = np.array([[E._get_projections(b, kap) for b in range(E.nband)] for kap in range(E.nwk * E.nspin)])
    assert rel(P, g["proj"]) < TOL
    # the source wavefunction is untouched and still usable
    assert np.array_equal(src._get_coefficients(1, 0), pn.Wavefunction.from_image(c["image"], c["kws"]).Cs[0][1])
    # bad operator (not a symmetry of the reciprocal lattice) is reported, not silently mis-mapped
    bad = ops.copy()
    bad[:9] = np.array([[1.0, 0.3, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]).reshape(-1)
    assert not L.pawb200_expand_symm_wf(src.wf_ptr, len(c["maps"]), _lib.ip(c["maps"]), _lib.dp(bad), _lib.dp(drs),
                                        _lib.dp(pw.weights), _lib.ip(c["trs"]))
    with pytest.raises(_lib.PAWpyError):
        _lib.check()


def test_aug_recip_vs_reference_and_oracle(gan):
    # overlap_setup_recip + compensation_terms_recip (projector.c:727-848, 965-1077).
    # The reference stores CAs as complex64 and accumulates both dot products in single precision; the GPU
    # keeps complex64 storage (same layout as the plane-wave coefficients) and accumulates in FP64.  Bars:
    # FP32 round-off of the correction against the reference C, 1e-7 of it against the FP64-accumulating oracle
    # (the only difference left is which way a CA coefficient rounds to float).
    g = np.load(os.path.join(G, "aug_recip.npz"), allow_pickle=True)
    cat = [list(x) for x in g["cat"]]
    cR, cS = cases.small_case(seed=7, nband=6), cases.small_case(seed=11, nband=6, perturb=0.03)
    R, S = from_case(cR), from_case(cS)
    oR, oS = oracle_wf(cR), oracle_wf(cS)
    pr = pawpyc.CProjector(S, R)
    pr._setup_overlap(cat, True)
    opr = pn.Projector(oS, oR, cat, recip=True)
    scale = np.abs(g["aug_f0"]).max()
    for flip in (0, 1):
        for b in range(6):
            res = np.zeros(6 * 4, complex)
            pr._projection_recip(res, b, bool(flip))
            assert np.abs(res - g["aug_f%d" % flip][b]).max() < 5e-6 * scale
            assert np.abs(res - opr.compensation_terms_recip(b, bool(flip))).max() < 1e-7 * scale
        # batched call: pseudo + recip augmentation for every band pair
        want = np.array([opr.single_band_projection(b, bool(flip)) for b in range(6)]).reshape(6, 6, 4)
        got = pr._projection_matrix(bool(flip))
        assert np.abs(got - want.transpose(2, 0, 1)).max() < 1e-7 * scale
    # switching the same pair back to the real-space method drops the recip state
    pr2 = pawpyc.CProjector(S, R)
    pr2._setup_overlap(cat, False)
    g2 = np.load(os.path.join(G, "synth_gan.npz"), allow_pickle=True)
    res = np.zeros(6 * 4, complex)
    pr2._add_augmentation_terms(res, 0, False)
    assert rel(res, g2["aug_c0_f0"][0]) < TOL
    with pytest.raises(_lib.PAWpyError):
        pr._projection_recip(np.zeros(24, complex), 0, False)      # recip data was replaced by the real setup


@pytest.mark.parametrize("cat", [
    [[0], [0], [1, 2, 3], [], [], []],                                   # N_R only
    [[0, 1, 2], [0, 1, 2], [], [3], [], []],                             # N_S only
    [[], [], [0, 1, 2, 3], [0, 1, 2, 3], [0, 1, 2, 3], [0, 1, 2, 3]],    # everything unmatched
])
def test_aug_recip_site_categories_vs_oracle(gan, cat):
    cR, cS, R, S, oR, oS = gan
    pr = pawpyc.CProjector(S, R)
    pr._setup_overlap(cat, True)
    opr = pn.Projector(oS, oR, cat, recip=True)
    nb = cR["nband"]
    want = np.array([opr.compensation_terms_recip(b, False) for b in range(nb)])
    scale = max(np.abs(want).max(), 1e-30)
    for b in (0, nb - 1):
        res = np.zeros(nb * 4, complex)
        pr._projection_recip(res, b, False)
        assert np.abs(res - want[b]).max() < 1e-7 * scale


def test_momentum_matrix_vs_reference_and_oracle():
    # MomentumMatrix (momentum.c).  One-centre terms, the AE plane-wave expansion and the grids are
    # FP64 -> 1e-10 against the reference C; the plane-wave correlation is a float-complex sum in the reference
    # (FP64 here), so full matrix elements are compared at FP32 round-off with it and at 1e-10 with the FP64 oracle.
    g = np.load(os.path.join(G, "momentum.npz"))
    c = cases.small_case(seed=7, nband=4, encut=120.0)
    wf = from_case(c)
    mm = pawpyc.CMomentumMatrix(wf, float(g["encut"]))
    assert np.array_equal(mm.ggrid, g["ggrid"])                                  # G order: bit-exact
    assert np.array_equal(mm.gbounds, g["gbounds"]) and np.array_equal(mm.gdim, g["gdim"])
    assert np.array_equal(mm.grid3d, g["grid3d"])
    for name, args in (("m_00_00", (0, 0, 0, 0, 0, 0)), ("m_0k0_1k1", (0, 0, 0, 1, 1, 0)),
                       ("m_2k1s1_3k0s1", (2, 1, 1, 3, 0, 1))):
        got = mm._get_momentum_matrix_elems(*args)
        assert np.abs(got - g[name]).max() < 2e-6 * np.abs(g[name]).max()
    assert rel(mm._get_reciprocal_fullfw(1, 0, 0), g["full_b1k0s0"]) < TOL
    assert rel(mm._get_reciprocal_fullfw(3, 1, 1), g["full_b3k1s1"]) < TOL
    assert abs(mm._get_g_from_fullfw(0, 0, 0, 1, 1, 0, [1, 0, 0]) - complex(g["gfrom"])) < 1e-12
    # FP64 oracle on a sample of the grid
    o = oracle_wf(c)
    om = pn.MomentumMatrix(o, float(g["encut"]))
    sel = list(range(0, len(om.ggrid), 53))
    om.ggrid = om.ggrid[sel]
    got = mm._get_momentum_matrix_elems(2, 1, 1, 3, 0, 1)[sel]
    assert rel(got, om.momentum_matrix_elems(2, 1, 1, 3, 0, 1)) < TOL
    with pytest.raises(_lib.PAWpyError):
        mm._get_momentum_matrix_elems(9, 0, 0, 0, 0, 0)
