"""Size-independent properties at benchmark scale (BASELINE config 2 shapes) on the GPU."""
import os

import numpy as np
import pytest

from pawpyseed_b200 import _lib, pawpyc, synth

pytestmark = pytest.mark.gpu


def test_pseudo_overlap_gemm_at_full_size_vs_fp64_matmul():
    """600 x 600 x 116489 stream-K DMMA GEMM against an independent FP64 complex matmul (torch)."""
    import torch
    import bench
    w = bench.workload("cfg2", nband=600)
    (imgR, _), (imgS, _) = bench.make_images(w)
    R = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(imgR, w["kpts"], w["kws"]))
    S = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(imgS, w["kpts"], w["kws"]))
    got = pawpyc.CProjector(S, R)._projection_matrix(pseudo_only=True)[0]
    npw = len(w["gvecs"][0])
    nrecl = int(round(imgR[:8].view(np.float64)[0]))

    def coeffs(img):
        blk = img[3 * nrecl:(3 + 600) * nrecl].reshape(600, nrecl)[:, :8 * npw]
        return torch.from_numpy(np.ascontiguousarray(blk).view(np.complex64)).cuda().to(torch.complex128)
    want = (coeffs(imgS) @ coeffs(imgR).conj().T).cpu().numpy()
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-12
    # unit-norm random bands: diagonal of <S|S> is 1 (a checksum of checksums)
    self_ov = pawpyc.CProjector(S, S)._projection_matrix(pseudo_only=True)[0]
    assert np.abs(np.diag(self_ov) - 1).max() < 1e-6      # complex64 normalisation of the inputs


def test_projection_linearity_at_full_grid():
    """<p|a psi1 + b psi2> = a<p|psi1> + b<p|psi2> on the 90^3 / 216-site configuration."""
    import bench
    w = bench.workload("cfg2", nband=3)
    rng = np.random.default_rng(0)
    npw = len(w["gvecs"][0])
    c = (rng.standard_normal((3, npw)) + 1j * rng.standard_normal((3, npw))).astype(np.complex64)
    a, b = np.float32(0.5), np.float32(-0.25)       # exact in complex64
    c[2] = a * c[0] + b * c[1]
    img = synth.wavecar_image(w["lattice"], w["encut"], w["kpts"], 1, 3, [c], gvecs=w["gvecs"])
    wf = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(img, w["kpts"], w["kws"]))
    wf._c_projector_setup(1, len(w["labels_R"]), w["grid_encut"], w["labels_R"], w["coords_R"], w["dim"], w["pps"])
    p = [wf._get_projections(i, 0) for i in range(3)]
    assert len(p[0]) == 216 * 8
    # c[2] carries complex64 rounding of the combination; compare against the same rounded input
    exact = c[2].astype(np.complex128) - (a * c[0].astype(np.complex128) + b * c[1].astype(np.complex128))
    scale = np.abs(p[2]).max()
    assert np.abs(p[2] - (a * p[0] + b * p[1])).max() < 1e-10 * scale + 50 * np.abs(exact).max()


def _full_shape_parity(config, nb, kappa):
    """All sites of the named config, `nb` bands of one (k,spin) block: the B200 library against the unmodified
    reference C run on the same WAVECAR bytes and the same synthetic PAW set (projector.c:223-274, 850-963,
    pseudoprojector.c:63-90).  This is the in-bench parity check of bench.py at full site count."""
    import bench
    from oracle import ref_driver as rd
    if not rd.available():
        pytest.skip("oracle/_ref/libpawpy_ref.so not built")
    w = bench.workload(config, nband=nb)
    plan = bench.sample_plan(w, ns_each=10 ** 6, nb=nb, n_pair=nb, kappa=kappa)
    assert len(plan["R_sub"]) == len(w["labels_R"]) and len(plan["S_sub"]) == len(w["labels_S"])
    own = {kappa}
    imgs = bench.make_images(w, own=own)
    simgs = bench.sample_images(w, imgs, plan)
    _, ref = bench.ref_sample(w, simgs, plan, threads=os.cpu_count() or 1, collect=True,
                              timing=False)
    got = bench.gpu_sample(w, simgs, plan)
    rep = bench.parity_report(ref, got)
    assert rep["index_exact"], rep
    assert rep["projections_max_rel"] < 1e-10, rep
    assert rep["wave_projections_max_rel"] < 1e-10, rep
    assert rep["compensation_terms_max_rel"] < 1e-10, rep
    assert rep["pseudoprojection_max_abs"] < 5e-6, rep        # the reference accumulates this term in FP32
    return rep


def test_cfg2_subsample_vs_reference_c():
    """BASELINE config 2 shape: 216/215 Si sites (8 channels, sphere_project_il_kernel<1>), 90^3 grid (radix
    class 10), Gamma; 32 of the 600 bands."""
    _full_shape_parity("cfg2", nb=32, kappa=0)


def test_cfg3_shape_vs_reference_c():
    """BASELINE config 3 shape: GaN 512/511 sites (18-channel Ga -> sphere_project_il_kernel<3>, 8-channel N),
    144x126x60 grid (radix classes 12/14/10), k = b1/2 (Bloch phases on the sphere samples), 32 bands."""
    _full_shape_parity("cfg3", nb=32, kappa=1)
