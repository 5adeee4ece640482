"""Generates tests/golden/*.npz from the UNMODIFIED reference C (oracle/_ref/libpawpy_ref.so).

Run in the build container (needs /root/reference for the bundled WAVECARs):
    python tests/golden/make_golden.py [base] [realspace_proj] [volumetric] [desymm] [aug_recip] [momentum]    (default: all)
Inputs are (a) the first bands of the reference's own fixtures test_files/WAVECAR,
WAVECAR2.gz and noncollinear/WAVECAR re-packed into small WAVECAR images, and (b) seeded
synthetic cells from tests/cases.py.  Every stored output comes from the reference library
run on exactly the stored input, with the synthetic PAW set of pawpyseed_b200/synth.py
(POTCARs are licensed and not shipped with the reference).
"""
import gzip
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import ref_driver as rd  # noqa: E402
from pawpyseed_b200 import synth  # noqa: E402

REF_FILES = "/root/reference/test_files"


def repack(wf, nband, nk=None, ncl=False):
    """First `nband` bands (and first nk k-points) of a loaded reference wavefunction -> WAVECAR image."""
    nk = nk or wf.nwk
    kpts = np.array([wf.kpt(k) for k in range(nk)])
    gv = [wf.gvecs(k)[: (len(wf.gvecs(k)) // 2 if ncl else len(wf.gvecs(k)))] for k in range(nk)]
    occs, ens = [], []
    L = rd.lib()
    for s in range(wf.nspin):
        for k in range(nk):
            occs.append([L.get_occ(wf.ptr, b, k, s) for b in range(nband)])
            ens.append([L.get_energy(wf.ptr, b, k, s) for b in range(nband)])

    def coeffs(kap, npw):
        s, k = divmod(kap, nk)
        return np.array([wf.coeffs(k + s * wf.nwk, b) for b in range(nband)])
    img = synth.wavecar_image(wf.lattice, wf.encut, kpts, wf.nspin, nband, coeffs, occs=np.array(occs),
                              energies=np.array(ens), ncl=ncl, gvecs=gv)
    return img, kpts


def run_pair(imgR, imgS, kws, pps, labelsR, coordsR, labelsS, coordsS, dim, cats, bands, grids=True):
    """Reference outputs for basis R / wf S built from two images."""
    lat = None
    R, S = rd.RefWavefunction(imgR, kws), rd.RefWavefunction(imgS, kws)
    ge = synth.grid_encut(dim, R.lattice)
    R.setup_projections(pps, labelsR, coordsR, dim, ge)
    S.setup_projections(pps, labelsS, coordsS, dim, ge)
    NK = R.nwk * R.nspin
    out = dict(grid_encut=ge, chan_index=R.channel_index(),
               site_index_0=R.site_tables("proj")[0]["indices"],
               site_npts=np.array([len(t["indices"]) for t in R.site_tables("proj")]),
               proj_R=np.array([[R.projections(k, b) for b in range(R.nband)] for k in range(NK)]),
               proj_S=np.array([[S.projections(k, b) for b in range(S.nband)] for k in range(NK)]))
    for ci, cat in enumerate(cats):
        pr = rd.RefProjector(S, R, cat)
        for flip in (0, 1):
            ps = np.array([S.pseudoprojection(b, R, bool(flip)) for b in bands])
            aug = np.array([pr.add_augmentation_terms(np.zeros(R.nband * NK, complex), b, bool(flip))
                            for b in bands])
            out["pseudo_f%d" % flip] = ps      # single precision accumulate in the reference
            out["aug_c%d_f%d" % (ci, flip)] = aug
    if grids:
        out["state_b1_k1"] = R.realspace_state(1, 1 % R.nwk, R.nspin - 1)
        out["state_b1_k1_nophase"] = R.realspace_state(1, 1 % R.nwk, R.nspin - 1, remove_phase=True)
        out["density"] = R.chg_density()
    R.free()
    S.free()
    return out


def make_base():
    pps_ga = synth.synthetic_pps(["Ga"])
    kws = np.array([0.5, 0.5])
    dim = np.array([20, 20, 20], np.int32)
    # ---- case A: the reference's own Ga4 fixtures (BASELINE config 1), first 8 bands ---------
    w1 = rd.RefWavefunction(os.path.join(REF_FILES, "WAVECAR"), kws)
    w2 = rd.RefWavefunction(np.frombuffer(gzip.open(os.path.join(REF_FILES, "WAVECAR2.gz")).read(), np.uint8), kws)
    nb = 8
    img1, kpts = repack(w1, nb)
    img2, _ = repack(w2, nb)
    gv0 = w1.gvecs(0)
    w1.free(); w2.free()
    cats = [[[0, 1, 2, 3], [0, 1, 2, 3], [], [], [], []],
            [[], [], [0, 1, 2, 3], [0, 1, 2, 3], [0, 1, 2, 3], [0, 1, 2, 3]],      # DummyProjector-style
            [[0, 1], [0, 1], [2, 3], [2, 3], [2, 3, 2], [2, 3, 3]]]
    labels = np.zeros(4, np.int32)
    out = run_pair(img1, img2, kws, pps_ga, labels, cases.GA4_COORDS, labels, cases.GA4_COORDS, dim, cats,
                   bands=[0, 3, 7])
    np.savez_compressed(os.path.join(HERE, "ga4.npz"), image_R=img1, image_S=img2, kpts=kpts, kws=kws,
                        dim=dim, gvecs_k0=gv0, cats=np.array(cats, dtype=object), bands=[0, 3, 7], **out)
    # ---- case B: synthetic two-element cell, displaced wf structure ------------------------------
    cR, cS = cases.small_case(seed=7, nband=6), cases.small_case(seed=11, nband=6, perturb=0.03)
    catsB = [[[0, 1], [0, 1], [2, 3], [2, 3], [2, 3, 2], [2, 3, 3]]]
    out = run_pair(cR["image"], cS["image"], cR["kws"], cR["pps"], cR["labels"], cR["coords"], cS["labels"],
                   cS["coords"], cR["dim"], catsB, bands=[0, 5])
    np.savez_compressed(os.path.join(HERE, "synth_gan.npz"), cats=np.array(catsB, dtype=object), bands=[0, 5], **out)
    # ---- case C: the reference's noncollinear fixture, first 4 bands, first 2 k-points -------------
    wn = rd.RefWavefunction(os.path.join(REF_FILES, "noncollinear", "WAVECAR"), np.full(4, 0.25))
    assert wn.ncl
    imgn, kptsn = repack(wn, 4, nk=2, ncl=True)
    wn.free()
    kwsn = np.array([0.5, 0.5])
    dimn = np.array([30, 30, 30], np.int32)
    N = rd.RefWavefunction(imgn, kwsn)
    assert N.ncl and N.nwk == 2
    gen = synth.grid_encut(dimn, N.lattice)
    N.setup_projections(pps_ga, labels, cases.GA4_COORDS, dimn, gen)
    up = np.array([[N.projections(k, b, "up_projections") for b in range(4)] for k in range(2)])
    dn = np.array([[N.projections(k, b, "down_projections") for b in range(4)] for k in range(2)])
    st = N.realspace_state(2, 1, 0)
    dens = N.chg_density()
    N.free()
    np.savez_compressed(os.path.join(HERE, "ncl.npz"), image=imgn, kpts=kptsn, kws=kwsn, dim=dimn, up=up, down=dn,
                        state_b2_k1=st, density=dens, grid_encut=gen)


def make_realspace_proj():
    """Projector(method="realspace") on the Ga4 pair: project_realspace_state (density.c:205-230), band 3, 24^3."""
    g = np.load(os.path.join(HERE, "ga4.npz"), allow_pickle=True)
    pps_ga = synth.synthetic_pps(["Ga"])
    labels = np.zeros(4, np.int32)
    R, S = rd.RefWavefunction(g["image_R"], g["kws"]), rd.RefWavefunction(g["image_S"], g["kws"])
    for w in (R, S):
        w.setup_projections(pps_ga, labels, cases.GA4_COORDS, g["dim"], float(g["grid_encut"]))
    pr = rd.RefProjector(S, R, [[], [], [], [], [], []])
    dim = np.array([24, 24, 24], np.int32)
    res = pr.realspace_projection(3, dim)
    R.free(); S.free()
    np.savez_compressed(os.path.join(HERE, "realspace_proj.npz"), band=3, dim=dim, res=res)


def make_volumetric():
    """Text written by the reference's write_volumetric (density.c:461-477) for a seeded 3x4x5 grid."""
    import tempfile
    rng = np.random.default_rng(5)
    dim = np.array([3, 4, 5], np.int32)
    x = rng.standard_normal(60) * 10.0 ** rng.integers(-4, 5, 60)
    fn = os.path.join(tempfile.mkdtemp(), "v.txt")
    rd.lib().write_volumetric(fn.encode(), rd._dp(x), rd._ip(dim), 1.5)
    np.savez_compressed(os.path.join(HERE, "volumetric.npz"), x=x, dim=dim, scale=1.5, text=open(fn).read())


def make_desymm():
    """expand_symm_wf (utils.c:829-1098) on a seeded cubic, spin-polarised cell: rotations, an inversion,
    time reversal and fractional translations; stores the new coefficients and the projections after
    setup_projections on the expanded wavefunction."""
    c = cases.desymm_case()
    R = rd.RefWavefunction(c["image"], c["kws"])
    E = R.expand_symm(c["maps"], c["ops"], c["drs"], c["new_kws"], c["trs"])
    NK = E.nwk * E.nspin
    kpts = np.array([E.kpt(k) for k in range(E.nwk)])
    gvecs = np.array([E.gvecs(k) for k in range(E.nwk)], dtype=object)
    coeffs = np.array([[E.coeffs(kap, b) for b in range(E.nband)] for kap in range(NK)], dtype=object)
    E.setup_projections(c["pps"], c["labels"], c["coords"], c["dim"], c["grid_encut"])
    proj = np.array([[E.projections(k, b) for b in range(E.nband)] for k in range(NK)])
    occ = np.array([[rd.lib().get_occ(E.ptr, b, k % E.nwk, k // E.nwk) for b in range(E.nband)] for k in range(NK)])
    np.savez_compressed(os.path.join(HERE, "desymm.npz"), kpts=kpts, gvecs=gvecs, coeffs=coeffs, proj=proj, occ=occ)
    R.free()


def make_recip():
    """method "aug_recip" (overlap_setup_recip + compensation_terms_recip, projector.c:727-848, 965-1077) on the
    synthetic two-element pair of synth_gan.npz: all three unmatched-site mechanisms (N_R, N_S, N_RS)."""
    cR, cS = cases.small_case(seed=7, nband=6), cases.small_case(seed=11, nband=6, perturb=0.03)
    cat = [[0, 1], [0, 1], [2, 3], [2, 3], [2, 3, 2], [2, 3, 3]]
    R, S = rd.RefWavefunction(cR["image"], cR["kws"]), rd.RefWavefunction(cS["image"], cS["kws"])
    for w, c in ((R, cR), (S, cS)):
        w.setup_projections(c["pps"], c["labels"], c["coords"], cR["dim"], cR["grid_encut"])
    pr = rd.RefProjector(S, R, cat, recip=True)
    NK = R.nwk * R.nspin
    out = {}
    for flip in (0, 1):
        out["aug_f%d" % flip] = np.array([pr.add_augmentation_terms(np.zeros(R.nband * NK, complex), b, bool(flip))
                                          for b in range(S.nband)])
    R.free(); S.free()
    np.savez_compressed(os.path.join(HERE, "aug_recip.npz"), cat=np.array(cat, dtype=object), **out)


def make_momentum():
    """MomentumMatrix (momentum.c) on the synthetic two-element cell, encut = 2 * wf.encut: the G grid, matrix elements
    for same-k and cross-k / cross-spin band pairs, the plane-wave expansion of an AE band and one g_from_wf value."""
    c = cases.small_case(seed=7, nband=4, encut=120.0)
    R = rd.RefWavefunction(c["image"], c["kws"])
    R.setup_projections(c["pps"], c["labels"], c["coords"], c["dim"], c["grid_encut"])
    mm = rd.RefMomentumMatrix(R, 2 * R.encut)
    out = dict(ggrid=mm.ggrid, gbounds=mm.gbounds, gdim=mm.gdim, grid3d=mm.grid3d, encut=2 * R.encut)
    for name, args in (("m_00_00", (0, 0, 0, 0, 0, 0)), ("m_0k0_1k1", (0, 0, 0, 1, 1, 0)), ("m_2k1s1_3k0s1", (2, 1, 1, 3, 0, 1))):
        out[name] = mm.momentum_matrix_elems(*args)
    out["full_b1k0s0"] = mm.reciprocal_fullfw(1, 0, 0)
    out["full_b3k1s1"] = mm.reciprocal_fullfw(3, 1, 1)
    out["gfrom"] = np.array(mm.g_from_fullfw(0, 0, 0, 1, 1, 0, [1, 0, 0]))
    R.free()
    np.savez_compressed(os.path.join(HERE, "momentum.npz"), **out)


MAKERS = {"base": make_base, "realspace_proj": make_realspace_proj, "volumetric": make_volumetric,
          "desymm": make_desymm, "aug_recip": make_recip, "momentum": make_momentum}


def main():
    os.makedirs(HERE, exist_ok=True)
    for name in (sys.argv[1:] or list(MAKERS)):
        MAKERS[name]()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
