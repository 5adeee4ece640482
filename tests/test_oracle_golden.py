"""The oracle (oracle/paw_numpy.py) against the committed golden vectors, which were produced by the
unmodified reference C (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import cases
from oracle import paw_numpy as pn
from pawpyseed_b200 import synth

G = cases.GOLDEN
TOL = 1e-10     # the FP64 bar of BASELINE.json north_star


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def ga4():
    g = np.load(os.path.join(G, "ga4.npz"), allow_pickle=True)
    pps = synth.synthetic_pps(["Ga"])
    R = pn.Wavefunction.from_image(g["image_R"], g["kws"])
    S = pn.Wavefunction.from_image(g["image_S"], g["kws"])
    lab = np.zeros(4, np.int32)
    for w in (R, S):
        w.setup_projections(pps, lab, cases.GA4_COORDS, g["dim"], float(g["grid_encut"]))
    return g, R, S


def test_reader_g_enumeration_matches_reference(ga4):
    g, R, _ = ga4
    assert np.array_equal(R.Gs[0], g["gvecs_k0"])          # bit-exact plane-wave order
    assert R.nband == 8 and R.nwk == 2 and R.nspin == 2


def test_indices_bit_exact(ga4):
    g, R, _ = ga4
    assert np.array_equal(R.chan_index, g["chan_index"])   # (site, n, l, m) order
    assert np.array_equal(R.sites[0]["indices"], g["site_index_0"])
    assert np.array_equal([len(s["indices"]) for s in R.sites], g["site_npts"])


def test_projections(ga4):
    g, R, S = ga4
    assert rel(np.array(R.P), g["proj_R"]) < TOL
    assert rel(np.array(S.P), g["proj_S"]) < TOL


@pytest.mark.parametrize("ci", [0, 1, 2])
@pytest.mark.parametrize("flip", [0, 1])
def test_compensation_terms(ga4, ci, flip):
    g, R, S = ga4
    pr = pn.Projector(S, R, [list(x) for x in g["cats"][ci]])
    got = np.array([pr.compensation_terms(int(b), bool(flip)) for b in g["bands"]])
    assert rel(got, g["aug_c%d_f%d" % (ci, flip)]) < TOL


def test_pseudoprojection_fp64_vs_reference_fp32(ga4):
    g, R, S = ga4
    got = np.array([S.pseudoprojection(int(b), R, False) for b in g["bands"]])
    # the reference accumulates and stores this term in single precision (pseudoprojector.c:86)
    assert np.abs(got - g["pseudo_f0"]).max() < 5e-6
    got = np.array([S.pseudoprojection(int(b), R, True) for b in g["bands"]])
    assert np.abs(got - g["pseudo_f1"]).max() < 5e-6


def test_realspace_state_and_density(ga4):
    g, R, _ = ga4
    x = R.realspace_state(1, 1 + 1 * 2)
    assert rel(x, g["state_b1_k1"]) < TOL
    assert rel(R.remove_phase(x, 3), g["state_b1_k1_nophase"]) < TOL
    assert rel(R.chg_density(g["dim"] * 2), g["density"]) < TOL


def test_synthetic_two_element_offsite():
    g = np.load(os.path.join(G, "synth_gan.npz"), allow_pickle=True)
    cR, cS = cases.small_case(seed=7, nband=6), cases.small_case(seed=11, nband=6, perturb=0.03)
    R = pn.Wavefunction.from_image(cR["image"], cR["kws"])
    S = pn.Wavefunction.from_image(cS["image"], cS["kws"])
    R.setup_projections(cR["pps"], cR["labels"], cR["coords"], cR["dim"], cR["grid_encut"])
    S.setup_projections(cS["pps"], cS["labels"], cS["coords"], cS["dim"], cS["grid_encut"])
    assert np.array_equal(R.chan_index, g["chan_index"])
    assert rel(np.array(R.P), g["proj_R"]) < TOL
    pr = pn.Projector(S, R, [list(x) for x in g["cats"][0]])
    for flip in (0, 1):
        got = np.array([pr.compensation_terms(int(b), bool(flip)) for b in g["bands"]])
        assert rel(got, g["aug_c0_f%d" % flip]) < TOL


def test_noncollinear_fixture():
    g = np.load(os.path.join(G, "ncl.npz"), allow_pickle=True)
    N = pn.Wavefunction.from_image(g["image"], g["kws"])
    assert N.ncl
    N.setup_projections(synth.synthetic_pps(["Ga"]), np.zeros(4, np.int32), cases.GA4_COORDS, g["dim"],
                        float(g["grid_encut"]))
    P = np.array(N.P)                     # [k][band][2][nproj]
    assert rel(P[:, :, 0], g["up"]) < TOL
    assert rel(P[:, :, 1], g["down"]) < TOL
    assert rel(N.realspace_state(2, 1), g["state_b2_k1"]) < TOL


def test_realspace_projection_method(ga4):
    # Projector(method="realspace") -> project_realspace_state (density.c:205-230)
    g, R, S = ga4
    gp = np.load(os.path.join(G, "realspace_proj.npz"))
    got = pn.project_realspace_state(int(gp["band"]), S, R, gp["dim"])
    assert rel(got, gp["res"]) < TOL


def test_desymmetrisation_restatement_vs_reference():
    # expand_symm_wf (utils.c:829-1098); golden from the reference C (make_golden.py desymm).
    # k-points, G lists and occupations are exact; coefficients carry complex64 phase factors (cexpf), so the
    # numpy restatement may differ from the C value by float rounding.
    c = cases.desymm_case()
    g = np.load(os.path.join(G, "desymm.npz"), allow_pickle=True)
    R = pn.Wavefunction.from_image(c["image"], c["kws"])
    E = pn.expand_symm_wf(R, c["maps"], c["ops"], c["drs"], c["new_kws"], c["trs"])
    assert np.array_equal(E.kpts, g["kpts"])
    assert all(np.array_equal(E.Gs[k], g["gvecs"][k]) for k in range(E.nwk))
    assert np.array_equal(E.occs, g["occ"])
    for kap in range(E.nwk * E.nspin):
        for b in range(E.nband):
            assert rel(E.Cs[kap][b], g["coeffs"][kap][b]) < 5e-7
    E.setup_projections(c["pps"], c["labels"], c["coords"], c["dim"], c["grid_encut"])
    assert rel(np.asarray(E.P), g["proj"]) < 5e-7


def test_symmetry_kpoint_generation_round_trip():
    # host side of row f2 (symmetry.py:34-164): every generated k-point maps back to its source
    from pawpyseed_b200 import symmetry
    ops = [symmetry.SymmOp(m) for m in cases.cubic_point_group()]
    assert len(ops) == 48
    kpts = np.array([[0.25, 0.25, 0.25], [0.0, 0.25, 0.5], [0.0, 0.0, 0.0]])
    allk, orig, opn, _, trs = symmetry.get_nosym_kpoints(kpts, symmops=ops)
    assert len(allk) == len(orig) == len(opn) == len(trs)
    # stars mod the lattice: 8 x (1/4,1/4,1/4), 12 x (0,1/4,1/2), Gamma; the half-zone filter of
    # symmetry.py:58-67 keeps 4 + 9 + 1 of them (zone-boundary members are their own -k image)
    assert len(allk) == 4 + 9 + 1
    for q, k in enumerate(allk):
        img = (-1.0 if trs[q] else 1.0) * ops[opn[q]].rotation_matrix @ kpts[orig[q]]
        assert np.allclose((img - k + 0.5) % 1 - 0.5, 0, atol=1e-9) or np.allclose(np.abs((img - k) % 1), [0, 0, 0], atol=1e-9)
    o2, n2, _, t2 = symmetry.get_kpt_mapping(allk, kpts, symmops=ops)
    assert list(o2) == list(orig)
    full, *_ = symmetry.get_nosym_kpoints(kpts, symmops=ops, fil_trsym=False)
    assert len(full) == 8 + 12 + 1
    with pytest.raises(Exception):
        symmetry.get_kpt_mapping(np.array([[0.1, 0.2, 0.3]]), kpts, symmops=ops)


def test_aug_recip_restatement_vs_reference():
    # overlap_setup_recip + compensation_terms_recip; golden from the reference C.
    # The reference keeps the augmentation coefficients and both dot products in single precision, so the
    # comparison is absolute at FP32 round-off of the (small) correction.
    g = np.load(os.path.join(G, "aug_recip.npz"), allow_pickle=True)
    cR, cS = cases.small_case(seed=7, nband=6), cases.small_case(seed=11, nband=6, perturb=0.03)
    R, S = pn.Wavefunction.from_image(cR["image"], cR["kws"]), pn.Wavefunction.from_image(cS["image"], cS["kws"])
    for w, c in ((R, cR), (S, cS)):
        w.setup_projections(c["pps"], c["labels"], c["coords"], cR["dim"], cR["grid_encut"])
    pr = pn.Projector(S, R, list(g["cat"]), recip=True)
    scale = np.abs(g["aug_f0"]).max()
    for b in (0, 5):
        for flip in (0, 1):
            got = pr.compensation_terms_recip(b, bool(flip))
            assert np.abs(got - g["aug_f%d" % flip][b]).max() < 5e-6 * scale


def test_momentum_matrix_restatement_vs_reference():
    # MomentumMatrix (momentum.c); golden from the reference C.  The plane-wave correlation is a
    # float-complex sum there, so matrix elements agree to FP32 round-off; everything else is FP64.
    g = np.load(os.path.join(G, "momentum.npz"))
    c = cases.small_case(seed=7, nband=4, encut=120.0)
    o = pn.Wavefunction.from_image(c["image"], c["kws"])
    o.setup_projections(c["pps"], c["labels"], c["coords"], c["dim"], c["grid_encut"])
    mm = pn.MomentumMatrix(o, float(g["encut"]))
    assert np.array_equal(mm.ggrid.reshape(-1), g["ggrid"])
    assert np.array_equal(mm.gbounds, g["gbounds"]) and np.array_equal(mm.gdim, g["gdim"])
    sel = list(range(0, len(mm.ggrid), 37))                    # the pure-python loops are slow: sample the grid
    saved = mm.ggrid
    mm.ggrid = saved[sel]
    got = mm.momentum_matrix_elems(0, 0, 0, 1, 1, 0)
    assert np.abs(got - g["m_0k0_1k1"][sel]).max() < 2e-6 * np.abs(g["m_0k0_1k1"]).max()
    full = mm.reciprocal_fullfw(1, 0, 0)
    assert rel(full, g["full_b1k0s0"][sel]) < TOL
