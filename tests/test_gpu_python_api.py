"""The reference's Python-facing API (Wavefunction / CoreRegion / Projector / NCLWavefunction) on the GPU engine."""
import os

import numpy as np
import pytest

import cases
from oracle import paw_numpy as pn
from pawpyseed_b200 import CoreRegion, NCLWavefunction, PAWpyError, Projector, Wavefunction, synth
from pawpyseed_b200.structure import Structure

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


def build(c, symbols):
    pps = c["pps"]
    cr = CoreRegion({sym: pps[i] for i, sym in enumerate(dict.fromkeys(symbols))})
    struct = Structure(c["lattice"], symbols, c["coords"])
    return Wavefunction.from_arrays(struct, c["image"], cr, c["dim"], c["kpts"], c["kws"])


def oracle(c):
    w = pn.Wavefunction.from_image(c["image"], c["kws"])
    w.setup_projections(c["pps"], c["labels"], c["coords"], c["dim"], c["grid_encut"])
    return w


@pytest.fixture(scope="module")
def pair():
    # basis: 4 sites; wf: same cell with site 3 displaced by 0.35 A -> M = {0,1,2}, N_R = N_S = {3}, N_RS = {(3,3)}
    cR = cases.small_case(seed=31, nband=8)
    coordsS = cR["coords"].copy()
    coordsS[3] += np.linalg.solve(cR["lattice"].T, np.array([0.35, 0.0, 0.0]))
    cS = cases.small_case(seed=32, nband=8, coords=coordsS)
    sym = ["Ga", "N", "Ga", "N"]
    return cR, cS, build(cR, sym), build(cS, sym), oracle(cR), oracle(cS)


def test_projector_aug_real_matches_oracle(pair):
    cR, cS, basis, wf, oR, oS = pair
    pr = Projector(wf, basis)                                   # method="aug_real", runs setup_overlap
    M_R, M_S, N_R, N_S, N_RS = pr.make_site_lists()
    assert (M_R, M_S, N_R, N_S, N_RS) == ([0, 1, 2], [0, 1, 2], [3], [3], [(3, 3)])
    cat = [M_R, M_S, N_R, N_S, [3], [3]]
    opr = pn.Projector(oS, oR, cat)
    for b in (0, 7):
        assert rel(pr.single_band_projection(b), opr.single_band_projection(b)) < TOL
    res = pr.single_band_projection(2, flip_spin=True)
    assert rel(res, opr.single_band_projection(2, True)) < TOL
    full = pr.projection_matrix()
    assert rel(full[1, 5], opr.single_band_projection(5).reshape(8, 4)[:, 1]) < TOL
    with pytest.raises(ValueError):
        pr.single_band_projection(99)                           # projector.py:270-271


def test_proportion_conduction_and_defect_band_analysis(pair):
    cR, cS, basis, wf, oR, oS = pair
    pr = Projector(wf, basis)
    opr = pn.Projector(oS, oR, pr.site_cat)
    occs = basis._get_occs()
    nb, nwk, nspin = 8, 2, 2
    assert occs.shape == (nb * nwk * nspin,)
    res = opr.single_band_projection(3)
    v = sum(abs(res[i]) ** 2 * cS["kws"][i % nwk] / nspin for i in range(len(res)) if occs[i] > 0.5)
    c = sum(abs(res[i]) ** 2 * cS["kws"][i % nwk] / nspin for i in range(len(res)) if occs[i] <= 0.5)
    gv, gc = pr.proportion_conduction(3)
    assert abs(gv - v) < 1e-10 and abs(gc - c) < 1e-10
    sv, sc = pr.proportion_conduction(3, spinpol=True)
    assert len(sv) == 2 and abs(sum(sv) / nspin - v) < 1e-10
    out, energies = pr.defect_band_analysis(num_below_ef=1, num_above_ef=1, return_energies=True)
    assert sorted(out) == [2, 3, 4]                             # vbm = 3 (4 of 8 bands occupied)
    assert len(energies[3]) == nwk * nspin
    assert pr.defect_band_analysis(analyze_all=True).keys() == set(range(8))


def test_pseudo_method_and_errors(pair):
    cR, cS, basis, wf, oR, oS = pair
    pr = Projector(wf, basis, method="pseudo")
    assert rel(pr.single_band_projection(1), oS.pseudoprojection(1, oR)) < 1e-12
    v, c = pr.proportion_conduction(1)
    assert abs(v + c - 1) < 1e-12                               # normalised for method='pseudo' (projector.py:428-431)
    with pytest.raises(PAWpyError):
        Projector(wf, basis, method="nonsense")
    with pytest.raises(PAWpyError):
        pr.setup_overlap()                                      # projector.py:180-181: aug methods only


def test_wavefunction_realspace_and_files(pair, tmp_path):
    cR, _, basis, _, oR, _ = pair
    x = basis.get_state_realspace(2, 1, 0)
    assert rel(x, oR.realspace_state(2, 1)) < TOL
    os.chdir(tmp_path)
    rho = basis.write_density_realspace("AECCAR_test.vasp", scale=2.0)
    assert rho.shape == tuple(2 * cR["dim"])
    assert rel(rho, oR.chg_density(cR["dim"] * 2)) < TOL
    lines = open("AECCAR_test.vasp").read().split("\n")
    assert lines[0] == "AECCAR_test.vasp" and lines[7] == "Direct"
    hdr = 8 + 4 + 1
    assert lines[hdr].split() == [str(2 * d) for d in cR["dim"]]
    vals = np.array(" ".join(lines[hdr + 1:]).split(), dtype=float)
    # VASP order: x fastest, z slowest (density.c:461-477), scaled
    assert np.allclose(vals, 2.0 * rho.transpose(2, 1, 0).reshape(-1), rtol=1e-6)
    basis.write_state_realspace(1, 0, 1, fileprefix="st_")
    assert os.path.exists("st_B1K0S1_REAL.vasp") and os.path.exists("st_B1K0S1_IMAG.vasp")
    with pytest.raises(ValueError):
        basis.get_state_realspace(0, 0, 5)


def test_ncl_wavefunction_api():
    g = np.load(os.path.join(cases.GOLDEN, "ncl.npz"), allow_pickle=True)
    cr = CoreRegion({"Ga": synth.synthetic_pps(["Ga"])[0]})
    struct = Structure(cases.GA4_LATTICE, ["Ga"] * 4, cases.GA4_COORDS)
    wf = NCLWavefunction.from_arrays(struct, g["image"], cr, g["dim"], g["kpts"], g["kws"])
    s0, s1 = wf.get_state_realspace(2, 1, 0)
    assert rel(np.stack([s0, s1]), g["state_b2_k1"]) < TOL
    assert rel(wf.get_realspace_density(), g["density"]) < TOL
    with pytest.raises(PAWpyError):
        Projector(wf, wf)                                       # projector.py:74-75
    with pytest.raises(PAWpyError):
        Wavefunction.from_arrays(struct, g["image"], cr, g["dim"], g["kpts"], g["kws"])   # wavefunction.py:205-208


def test_desymmetrized_copy_matches_oracle():
    # wavefunction.py:249-279 with explicit operators (no pymatgen here): irreducible k-set -> half zone
    from pawpyseed_b200 import symmetry
    c = cases.desymm_case()
    wf = build(c, ["Ga", "N", "Ga", "N"])
    ops = [symmetry.SymmOp(m) for m in cases.cubic_point_group()]
    new = wf.desymmetrized_copy(symmops=ops)
    allk, orig, opn, _, trs = symmetry.get_nosym_kpoints(c["kpts"], symmops=ops)
    assert new.nwk == len(allk) == 14 and np.allclose(new.kpts, allk)
    assert abs(new.kws.sum() - 1) < 1e-14 and new.kws[2] == pytest.approx(0.5 * new.kws[0])
    R = pn.Wavefunction.from_image(c["image"], c["kws"])
    o_ops, o_drs = symmetry.make_c_ops(opn, ops)
    E = pn.expand_symm_wf(R, orig, o_ops.reshape(-1, 3, 3), o_drs.reshape(-1, 3), new.kws, trs)
    for kap in (0, 5, 13, 14 + 9):
        for b in (0, c["nband"] - 1):
            assert rel(new._get_coefficients(b, kap), E.Cs[kap][b]) < 5e-7
    # a projection from the expanded set onto itself is the identity on the pseudo + augmentation level
    new.check_c_projectors()
    E.setup_projections(c["pps"], c["labels"], c["coords"], c["dim"], c["grid_encut"])
    P = np.array([new._get_projections(b, 3) for b in range(new.nband)])
    assert rel(P, np.asarray(E.P)[3]) < 5e-6
    # mapping onto a given mesh
    sub = wf.desymmetrized_copy(allkpts=allk[[2, 7]], weights=np.array([0.5, 0.5]), symmops=ops)
    assert sub.nwk == 2 and np.allclose(sub.kpts, allk[[2, 7]])


def test_projector_method_aug_recip(pair):
    # projector.py:225-236 through the public class; site lists come from make_site_lists
    cR, cS, basis, wf = pair[:4]
    pr = Projector(wf, basis, method="aug_recip")
    oR, oS = oracle(cR), oracle(cS)
    opr = pn.Projector(oS, oR, pr.site_cat, recip=True)
    want = opr.single_band_projection(2)
    got = pr.single_band_projection(2)
    assert np.abs(got - want).max() < 1e-7 * np.abs(want).max()
    M = pr.projection_matrix()
    assert np.abs(M[:, 2, :].T.reshape(-1) - got).max() < 1e-12


def test_density_band_shards_sum_to_full_density(pair):
    # band-split density = one all-reduce of the shard grids; here the two shards are summed in-process
    from pawpyseed_b200 import distributed as pdist
    basis = pair[2]
    full = basis._get_realspace_density()
    nb = basis.nband
    parts = [basis._get_realspace_density_shard(*pdist.band_block(nb, r, 4)) for r in range(4)]
    assert rel(sum(parts), full) < 1e-13
    assert np.abs(parts[0]).max() > 0 and np.abs(parts[1]).max() > 0       # the occupied half is split over two shards
    assert pdist.sharded_chg_density(basis).shape == full.shape            # no process group: plain call


def test_momentum_matrix_class(pair):
    # momentum.py:4-91 through the public class; self matrix element at G = 0 is the AE norm of the band
    from pawpyseed_b200 import MomentumMatrix
    basis = pair[2]
    mm = MomentumMatrix(basis, encut=1.5 * basis.encut)
    grid = mm.momentum_grid
    assert grid.shape[1] == 3 and (grid[0] == 0).all()
    res = mm.get_momentum_matrix_elems(0, 0, 0, 0, 0, 0)
    assert res.shape == (grid.shape[0],) and abs(res[0].imag) < 1e-12 and res[0].real > 0
    full = mm.get_reciprocal_fullfw(0, 0, 0)
    assert full.shape == res.shape
    assert abs(mm.g_from_wf(0, 0, 0, 0, 0, 0, [0, 0, 0]).imag) < 1e-12
    with pytest.raises(ValueError):
        mm.get_momentum_matrix_elems(99, 0, 0, 0, 0, 0)


def test_wavecar_file_and_gz_ingest_match_memory_image(tmp_path):
    # the reference's test_projector_gz / PWFPointer(filename) path (pawpyc.pyx:217-224): plain file through the
    # staged reader, .gz through the in-memory reader; 40 bands = two ingest chunks per (k,spin) block
    import gzip
    from pawpyseed_b200 import pawpyc
    c = cases.small_case(seed=13, nband=40)
    plain = tmp_path / "WAVECAR"
    plain.write_bytes(c["image"].tobytes())
    gz = tmp_path / "WAVECAR2.gz"
    with gzip.open(gz, "wb") as f:
        f.write(c["image"].tobytes())
    mem = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(c["image"], c["kpts"], c["kws"]))
    for src in (str(plain), str(gz)):
        wf = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(src, c["kpts"], c["kws"]))
        assert (wf.nband, wf.nwk, wf.nspin, wf.ncl) == (40, 2, 2, False) and wf.encut == mem.encut
        for kap in range(4):
            for b in (0, 31, 32, 39):
                assert np.array_equal(wf._get_coefficients(b, kap), mem._get_coefficients(b, kap))
        assert np.array_equal(wf._get_occs(), mem._get_occs())
    with pytest.raises(PAWpyError):
        pawpyc.PWFPointer.from_arrays(str(tmp_path / "missing"), c["kpts"], c["kws"])


def test_projector_with_desymmetrised_pair():
    # projector.py:77-95: unsym_basis + unsym_wf bring both onto the unreduced mesh (GPU remap), then the usual
    # PAW-corrected projection; oracle = numpy expand_symm_wf + numpy Projector with the same mapping
    from pawpyseed_b200 import symmetry
    cR, cS = cases.desymm_case(seed=21), cases.desymm_case(seed=22)
    sym = ["Ga", "N", "Ga", "N"]
    basis, wf = build(cR, sym), build(cS, sym)
    ops = [symmetry.SymmOp(m) for m in cases.cubic_point_group()]
    pr = Projector(wf, basis, unsym_basis=True, unsym_wf=True, symmops=ops)
    assert pr.basis.nwk == pr.wf.nwk == 14 and np.allclose(pr.basis.kpts, pr.wf.kpts)
    allk, orig, opn, _, trs = symmetry.get_nosym_kpoints(cR["kpts"], symmops=ops)
    o_ops, o_drs = symmetry.make_c_ops(opn, ops)
    exp = []
    for c in (cR, cS):
        w = pn.Wavefunction.from_image(c["image"], c["kws"])
        e = pn.expand_symm_wf(w, orig, o_ops.reshape(-1, 3, 3), o_drs.reshape(-1, 3), pr.basis.kws, trs)
        e.setup_projections(c["pps"], c["labels"], c["coords"], c["dim"], c["grid_encut"])
        exp.append(e)
    opr = pn.Projector(exp[1], exp[0], pr.site_cat)
    want = opr.single_band_projection(1)
    got = pr.single_band_projection(1)
    # coefficients of the expanded sets carry complex64 phase factors whose last bit may differ between numpy and libm
    assert np.abs(got - want).max() < 5e-6 * np.abs(want).max()
