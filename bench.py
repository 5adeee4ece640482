#!/usr/bin/env python
"""bench.py - PAW-corrected band-pair projections / second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config cfg3|cfg2|tiny]

One "step" = one pass of the whole hot path over one synthetic wavefunction pair:
setup_projections(basis) + setup_projections(wf) + overlap_setup_real + every band pair
(pseudo overlap GEMM + augmentation GEMM) + result to host.

Default workload = BASELINE config 3, the north-star job: GaN 512-site cell x N-vacancy cell, ENCUT 520,
2000 bands, spin-polarised, 4 k-points = 8 (k,spin) blocks of 2000 x 2000 pairs.  `--gpus N` runs the SAME job
on N GPUs ("scaling": "strong"): rank r owns the blocks kappa % N == r (no data-path collective), and the
per-kappa result matrices are all-gathered over NCCL at the end of each step.  Config 2 (Si216 x Si215, 600
bands, Gamma) is measured as a secondary workload at N = 1 and reported under the key "cfg2".

* `value`  : pairs/s with the plane-wave coefficients already resident in HBM (device-timed, CUDA events).
* `e2e`    : pairs/s through the public API from HOST WAVECAR images in pinned memory
             (read_wavefunctions_from_str -> ... -> result matrix on host), H2D/D2H inside the timed region.
* `roofline`: the dominant kernel (stream-K DMMA complex GEMM of the pseudo overlap) against the FP64
             GEMM rate measured in this run (MEASURED_PEAKS.json has no FP64 entry); HBM-bound kernels are
             listed under `kernels` against MEASURED_PEAKS.json's copy bandwidth.
* `cpu_baseline`: the unmodified reference C (oracle/_ref) on this box's host cores, on a bounded sample of the
             workload (a band block x a site subset of one (k,spin) block).  Every coefficient of the reference's
             cost model is measured in that sample; the model's scaling laws are the reference's own loop bounds
             (see ref_model()).  `--ref-full` instead times the complete workload (config 2: ~30 s).
* `parity` : the GPU results for that same sample (projections, compensation_terms, pseudoprojection, channel and
             sphere index arrays) compared with what the reference C just computed.
* `--impl reference`: the reference arm - K steps of that sample on the host cores; `ms_per_step` is the measured
             wall time of a sample step, `value` the modelled full-workload pairs/s (`model` holds every term).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pawpyseed_b200 import synth  # noqa: E402

METRIC = "paw_band_pair_projections_per_sec"
UNIT = "pairs/s"


# --------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------
def workload(name, nk=1, nband=None, blocks=None):
    """Returns dict(lattice, encut, kpts, nspin, nband, basis/wf coords+labels, elements, site_cat, dim).
    nk is only used by config 2's weak-scaling variant (nk Gamma-like k-points, one per GPU).
    blocks (debug): keep only the first `blocks` (k,spin) blocks of a multi-k workload (1 = what one rank of an
    8-GPU config-3 job holds)."""
    if name == "cfg2":      # Si216 bulk vs Si215 vacancy, ENCUT 520, Gamma (full sphere), 600 bands
        lat, bulk = synth.diamond_supercell(5.43, 3)
        _, defect = synth.diamond_supercell(5.43, 3, vacancy=107)
        w = dict(name="Si216 bulk (basis) x Si215+vacancy (wf), ENCUT 520, 600 bands, Gamma, 90^3 grid",
                 lattice=lat, encut=520.0, nband=600, nspin=1, elements=["Si"], vac=107,
                 coords_R=bulk, labels_R=np.zeros(len(bulk), np.int32),
                 coords_S=defect, labels_S=np.zeros(len(defect), np.int32))
    elif name == "cfg3":    # BASELINE config 3: GaN 512-site cell + N vacancy, ENCUT 520, 2000 bands, 4 k x 2 spins
        lat, frac, lab = synth.wurtzite_supercell((8, 4, 2))
        vac = int(np.where(lab == 1)[0][100])
        w = dict(name="GaN512 bulk (basis) x GaN511 N-vacancy (wf), ENCUT 520, 2000 bands, spin-polarised, "
                      "4 k-points {0, b1/2, b2/2, b3/2} x 2 spins, 144x126x60 grid",
                 lattice=lat, encut=520.0, nband=2000, nspin=2,
                 elements=["Ga", "N"], vac=vac, coords_R=frac, labels_R=lab.astype(np.int32),
                 coords_S=np.delete(frac, vac, axis=0), labels_S=np.delete(lab, vac).astype(np.int32),
                 kpt_list=[[0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [0.0, 0.5, 0.0], [0.0, 0.0, 0.5]],
                 dim=np.array([144, 126, 60], np.int32))   # the PREC=Normal grid of this cell
    elif name == "tiny":    # CPU-sized smoke configuration of the same shape (two elements, 2 k x 2 spins)
        lat, frac, lab = synth.wurtzite_supercell((1, 1, 1))
        w = dict(name="GaN8 x GaN7 N-vacancy, ENCUT 200, 24 bands, 2 k-points x 2 spins", lattice=lat, encut=200.0,
                 nband=24, nspin=2, elements=["Ga", "N"], vac=5, coords_R=frac, labels_R=lab.astype(np.int32),
                 coords_S=np.delete(frac, 5, axis=0), labels_S=np.delete(lab, 5).astype(np.int32),
                 kpt_list=[[0.0, 0.0, 0.0], [0.5, 0.0, 0.0]])
    else:
        raise SystemExit("unknown config %s" % name)
    if nband:
        w["nband"] = int(nband)
    if blocks and "kpt_list" in w:
        if blocks < len(w["kpt_list"]):
            w["nspin"] = 1
        w["kpt_list"] = w["kpt_list"][:max(1, min(blocks, len(w["kpt_list"])))]
        w["name"] += " [debug: first %d block(s)]" % (len(w["kpt_list"]) * w["nspin"])
    w["key"] = name
    if "kpt_list" in w:
        w["kpts"] = np.array(w["kpt_list"])
    else:
        # k-points: Gamma for nk == 1; for the weak-scaling variant, nk distinct points along b1
        w["kpts"] = np.array([[0.0, 0.0, 0.0]] if nk == 1 else [[0.5 * i / nk, 0.0, 0.0] for i in range(nk)])
    w["nk"] = len(w["kpts"])
    w["kws"] = np.full(w["nk"], 1.0 / w["nk"])
    w["gvecs"] = [synth.enumerate_gvectors(w["lattice"], w["encut"], k) for k in w["kpts"]]
    if "dim" not in w:
        w["dim"] = synth.fft_grid_for(w["gvecs"])
    w["grid_encut"] = synth.grid_encut(w["dim"], w["lattice"])
    nR = len(w["coords_R"])
    vac = w["vac"]
    M_R = [i for i in range(nR) if i != vac]
    w["site_cat"] = [M_R, list(range(nR - 1)), [vac], [], [], []]
    w["pps"] = synth.synthetic_pps(w["elements"])
    return w


def shard_mode(w, world):
    """(k,spin) blocks round-robin over ranks when there are enough of them, else band blocks of the blocks."""
    return "kappa" if world == 1 or w["nk"] * w["nspin"] >= world else "bands"


def config_dict(w, world, scaling):
    """`config` of the JSON line - built by the same function for both arms so that they compare like with like."""
    NK = w["nk"] * w["nspin"]
    how = ("(k,spin) blocks round-robin" if shard_mode(w, world) == "kappa" else
           "band blocks of each (k,spin) block (basis rows all-gathered over NVLink)")
    return {"workload": w["name"], "nband": w["nband"], "npw": [len(g) for g in w["gvecs"]],
            "fft_grid": [int(x) for x in w["dim"]], "sites": [len(w["labels_R"]), len(w["labels_S"])],
            "kappa_blocks": NK, "pairs_per_step": w["nband"] ** 2 * NK,
            "parallelism": "%s over %d GPU(s), %s" % (how, world, scaling),
            "l2": "inputs (%.1f GB coefficients + FFT boxes per step) exceed the 126 MB L2" %
                  (2 * 8 * w["nband"] * sum(len(g) for g in w["gvecs"]) * w["nspin"] / 1e9)}


def _block_bytes(img):
    return int(round(img[:8].view(np.float64)[0]))


def make_images(w, own=None, nband=None, pinned=False, use_gpu=None):
    """WAVECAR images (basis, wf) as [(uint8 array, None)].  own: set of kappa whose coefficient records are
    filled (others stay untouched zero pages - sharded ranks never read them).  Coefficients are N(0,1) + i N(0,1),
    L2-normalised per band; with a GPU they are drawn by torch's Philox generator on the device (seed
    = 1000*structure + kappa) and copied into the image, otherwise by numpy (small CPU-side cases)."""
    import torch
    nband = nband or w["nband"]
    NK = w["nk"] * w["nspin"]
    kaps = list(range(NK)) if own is None else sorted(own)
    if use_gpu is None:
        use_gpu = torch.cuda.is_available()
    imgs = []
    for sid in (0, 1):
        if use_gpu:
            img = synth.wavecar_image(w["lattice"], w["encut"], w["kpts"], w["nspin"], nband, lambda kap, npw: None,
                                      gvecs=w["gvecs"])
        else:
            def gen(kap, npw, _sid=sid):
                if own is not None and kap not in own:
                    return None
                return synth.random_coeffs(2000 + _sid, nband)(kap, npw)
            img = synth.wavecar_image(w["lattice"], w["encut"], w["kpts"], w["nspin"], nband, gen, gvecs=w["gvecs"])
        nrecl = _block_bytes(img)
        pinned_spans = None
        if pinned:
            # page-lock only the coefficient records this rank reads (the image of a sharded job is mostly
            # other ranks' zero pages)
            rt = torch.cuda.cudart()
            spans = []          # page-aligned [a0, a1), adjacent / overlapping blocks merged (a page registers once)
            for kap in kaps:
                lo = (2 + kap * (1 + nband)) * nrecl
                hi = lo + (1 + nband) * nrecl
                a0 = (img.ctypes.data + lo) // 4096 * 4096
                a1 = min(-(-(img.ctypes.data + hi) // 4096) * 4096, img.ctypes.data + img.nbytes)
                if spans and a0 <= spans[-1][1]:
                    spans[-1][1] = max(spans[-1][1], a1)
                else:
                    spans.append([a0, a1])
            for a0, a1 in spans:
                err = rt.cudaHostRegister(a0, a1 - a0, 0)
                if int(err) != 0:
                    raise SystemExit("cudaHostRegister(%d bytes) failed: %s" % (a1 - a0, err))
            pinned_spans = spans
        if use_gpu:
            for kap in kaps:
                npw = len(w["gvecs"][kap % w["nk"]])
                g = torch.Generator(device="cuda")
                g.manual_seed(1000 * (sid + 1) + kap)
                c = torch.zeros(nband, nrecl // 8, 2, dtype=torch.float32, device="cuda")
                c[:, :npw] = torch.randn(nband, npw, 2, dtype=torch.float32, device="cuda", generator=g)
                c /= c.view(nband, -1).norm(dim=1).view(nband, 1, 1)
                lo = (3 + kap * (1 + nband)) * nrecl
                dst = torch.from_numpy(img[lo:lo + nband * nrecl].view(np.float32))
                dst.copy_(c.view(-1))
                del c
            torch.cuda.synchronize()
        imgs.append((img, pinned_spans))
    return imgs


def unpin_images(imgs):
    """Undo make_images(pinned=True) before the arrays are freed: the allocator may hand the same pages to the
    next image, and a page cannot be registered twice."""
    import torch
    torch.cuda.synchronize()
    rt = torch.cuda.cudart()
    for i, (img, spans) in enumerate(imgs):
        for a0, _ in spans or []:
            rt.cudaHostUnregister(a0)
        imgs[i] = (img, None)


def coefficient_block(w, img, kap, nband_total, nb):
    """complex64 [nb, npw] view of the first nb bands of block kappa of a WAVECAR image."""
    nrecl = _block_bytes(img)
    npw = len(w["gvecs"][kap % w["nk"]])
    lo = (3 + kap * (1 + nband_total)) * nrecl
    return img[lo:lo + nb * nrecl].reshape(nb, nrecl)[:, :8 * npw].view(np.complex64)


# --------------------------------------------------------------------------------------------
# the CPU sample: a band block x a site subset of one (k,spin) block, shared by the cpu_baseline leg, the
# reference arm and the in-bench parity check
# --------------------------------------------------------------------------------------------
def sample_plan(w, ns_each, nb, n_pair, kappa):
    """Site subset = the vacancy site of the basis + the first `ns_each` sites of every element (same element mix
    as the full cell); the wf structure carries the same sites minus the vacancy."""
    labR, vac = w["labels_R"], w["vac"]
    chosen = []
    for el in sorted(set(int(x) for x in labR)):
        chosen += [i for i in range(len(labR)) if labR[i] == el and i != vac][:ns_each]
    chosen.sort()
    R_sub = sorted(chosen + [vac])
    S_sub = [i if i < vac else i - 1 for i in chosen]
    cat = [[R_sub.index(i) for i in chosen], list(range(len(chosen))), [R_sub.index(vac)], [], [], []]
    nb = min(nb, w["nband"])
    return dict(R_sub=R_sub, S_sub=S_sub, cat=cat, nb=nb, n_pair=min(n_pair, nb), kappa=kappa,
                k_index=kappa % w["nk"])


def sample_images(w, imgs, plan):
    """Single-(k,spin)-block WAVECAR images holding the first nb bands of block `kappa` of the full images."""
    out = []
    k = plan["k_index"]
    for (img, _) in imgs:
        c = np.ascontiguousarray(coefficient_block(w, img, plan["kappa"], w["nband"], plan["nb"]))
        out.append(synth.wavecar_image(w["lattice"], w["encut"], w["kpts"][k:k + 1], 1, plan["nb"], [c],
                                       gvecs=[w["gvecs"][k]]))
    c8 = np.ascontiguousarray(coefficient_block(w, imgs[0][0], plan["kappa"], w["nband"], min(8, plan["nb"])))
    out.append(synth.wavecar_image(w["lattice"], w["encut"], w["kpts"][k:k + 1], 1, len(c8), [c8],
                                   gvecs=[w["gvecs"][k]]))     # 8-band basis: per-call overhead probe
    return out


class _Quiet:
    """The reference printf()s progress lines; keep them out of the JSON stream."""

    def __enter__(self):
        sys.stdout.flush()
        self.devnull = os.open(os.devnull, os.O_WRONLY)
        self.saved = os.dup(1)
        os.dup2(self.devnull, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.devnull)
        os.close(self.saved)


def index_check_sites(plan, count=6):
    """positions (in the basis site subset) whose sphere index lists the parity check compares"""
    n = len(plan["R_sub"])
    return sorted(set(int(round(i * (n - 1) / max(count - 1, 1))) for i in range(min(count, n))))


def ref_sample(w, simgs, plan, threads, collect=False, warm=False, timing=True):
    """One pass of the unmodified reference C over the sample.  Returns wall times per reference call and, with
    collect=True, the values the parity check compares."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import ref_driver as rd
    kws = np.ones(1)
    labR, labS = w["labels_R"][plan["R_sub"]], w["labels_S"][plan["S_sub"]]
    crdR, crdS = w["coords_R"][plan["R_sub"]], w["coords_S"][plan["S_sub"]]
    nb, n_pair = plan["nb"], plan["n_pair"]
    t, out = {}, {}
    with _Quiet():
        t0 = time.perf_counter()
        R = rd.RefWavefunction(simgs[0], kws)
        S = rd.RefWavefunction(simgs[1], kws)
        F = rd.RefWavefunction(simgs[0], kws)
        t["read"] = time.perf_counter() - t0
        if warm:    # first use of this grid size in the process: MKL descriptor set-up, OpenMP team start-up
            W = rd.RefWavefunction(simgs[0], kws)
            W.setup_projections(w["pps"], labR[:1], crdR[:1], w["dim"], w["grid_encut"])
            W.free()
        # (1) the sample proper
        t0 = time.perf_counter()
        R.setup_projections(w["pps"], labR, crdR, w["dim"], w["grid_encut"])
        t["setup_R"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        S.setup_projections(w["pps"], labS, crdS, w["dim"], w["grid_encut"])
        t["setup_S"] = time.perf_counter() - t0
        if timing:
            t["site_R"] = R.time_projector_values()    # the serial setup_site share of setup_R, measured alone
            # (2) bands x ONE site: the FFT-dominated part of setup_projections with its real OpenMP behaviour
            t0 = time.perf_counter()
            F.setup_projections(w["pps"], labR[:1], crdR[:1], w["dim"], w["grid_encut"])
            t["setup_1site"] = time.perf_counter() - t0
            t["site_1"] = F.time_projector_values()
        F.free()
        t0 = time.perf_counter()
        pr = rd.RefProjector(S, R, plan["cat"])
        t["overlap_setup"] = time.perf_counter() - t0
        # the per-pair dot-product cost carries most of the modelled time and is host-memory-bandwidth bound, so it
        # varies from run to run (r02: 1.4e-5 .. 2.8e-5 s per pair): repeat the rows and keep the FASTEST pass
        # (the baseline least favourable to the B200 arm)
        best, reps, tstart = None, 0, time.perf_counter()
        while reps == 0 or (timing and reps < 8 and time.perf_counter() - tstart < 0.8):
            t0 = time.perf_counter()
            ps = [S.pseudoprojection(b, R) for b in range(n_pair)]
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            reps += 1
        t["pseudo"] = best
        t["pseudo_reps"] = reps
        if timing:
            # per-call overhead of the two per-band entry points (OpenMP fork/join, ctypes), so that it is not
            # charged per pair: pseudoprojection against an 8-band basis, compensation_terms with empty site lists
            T8 = rd.RefWavefunction(simgs[2], kws)
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                for b in range(n_pair):
                    S.pseudoprojection(b, T8)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            t["pseudo_8band_basis"] = best
            T8.free()
            empty = rd.RefProjector.__new__(rd.RefProjector)
            empty.wf, empty.basis, empty.recip = S, R, False
            empty.cat = [np.zeros(0, np.int32)] * 6
            t0 = time.perf_counter()
            for b in range(nb):
                empty.add_augmentation_terms(np.zeros(nb, np.complex128), b)
            t["compensation_no_sites"] = time.perf_counter() - t0
        # compensation_terms is cheap per row on a site subset: time every wf band of the sample, repeated until
        # the timed region is long enough to trust
        cp, reps, tstart, best = [], 0, time.perf_counter(), None
        while reps == 0 or (timing and time.perf_counter() - tstart < 0.25 and reps < 64):
            t0 = time.perf_counter()
            cp = [pr.add_augmentation_terms(np.zeros(nb, np.complex128), b) for b in range(nb)]
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            reps += 1
        t["compensation"] = best
        cp = cp[:n_pair]
        if collect:
            nchk = min(nb, 8)
            out["pseudo"] = np.array(ps)
            out["comp"] = np.array(cp)
            out["P_R"] = np.array([R.projections(0, b) for b in range(nchk)])
            out["P_S"] = np.array([S.projections(0, b) for b in range(nchk)])
            out["W_S"] = np.array([S.projections(0, b, "wave_projections") for b in range(nchk)])
            out["chan_R"] = R.channel_index()
            # sphere index lists of a few sites, from a reference wavefunction set up on just those sites
            # (setup_site treats every site independently, utils.c:635-696)
            pos = index_check_sites(plan)
            I = rd.RefWavefunction(simgs[0], kws)
            I.setup_projections(w["pps"], labR[pos], crdR[pos], w["dim"], w["grid_encut"])
            out["idx_R"] = [st["indices"] for st in I.site_tables("proj", indices_only=True)]
            I.free()
        R.free()
        S.free()
    return t, out


def ref_model(w, plan, t, threads):
    """Full-workload time of the reference from the sample's measured walls.  The scaling laws are the reference's
    own loop bounds:
      setup_projections (projector.c:560-602) = setup_site over the sites, SERIAL (utils.c:635, once per structure,
          independent of bands and k) + an OpenMP loop over bands x (k,spin) of [fft3d + per-site projection];
      overlap_setup_real (projector.c:625-646)  = setup_site over N_R + the same band loop on the N_R sites;
      pseudoprojection (pseudoprojector.c:63-90) per wf band = basis bands x (k,spin) dot products of npw;
      compensation_terms (projector.c:872-961)  per wf band = basis bands x (k,spin) x sum over listed sites.
    """
    nb_s, n_pair = plan["nb"], plan["n_pair"]
    nsR_s, nsS_s = len(plan["R_sub"]), len(plan["S_sub"])
    nsR, nsS = len(w["labels_R"]), len(w["labels_S"])
    nb, NK = w["nband"], w["nk"] * w["nspin"]
    c_site = t["site_R"] / nsR_s                                    # s per site (element mix of the subset)
    band_R = max(t["setup_R"] - t["site_R"], 0.0)                   # band loop at nsR_s sites
    band_S = max(t["setup_S"] - c_site * nsS_s, 0.0)
    band_1 = max(t["setup_1site"] - t["site_1"], 0.0)               # band loop at 1 site (FFT dominated)
    c_ps = max(band_R + band_S - 2 * band_1, 0.0) / (nb_s * (nsR_s + nsS_s - 2))   # s per (band, site), OMP wall
    c_f = max(band_1 / nb_s - c_ps, 0.0)                            # s per band transform, OMP wall
    full = {
        "setup_site_s": c_site * (nsR + nsS),
        "setup_bands_s": NK * nb * (2 * c_f + c_ps * (nsR + nsS)),
        "overlap_setup_s": c_site * len(plan["cat"][2]) +
                           NK * nb * max(t["overlap_setup"] - c_site * len(plan["cat"][2]), 0.0) / nb_s,
    }
    # per-band calls: t(call) = overhead + basis bands x per-pair cost
    a_ps = t["pseudo_8band_basis"] / n_pair
    c_dot = max(t["pseudo"] / n_pair - a_ps, 0.0) / max(nb_s - 8, 1)          # s per dot product (OMP wall)
    a_ps = max(a_ps - 8 * c_dot, 0.0)
    a_cp = t["compensation_no_sites"] / nb_s
    c_cp = max(t["compensation"] / nb_s - a_cp, 0.0) / nb_s / max(nsS_s, 1)   # s per (pair, matched site)
    full["pseudoprojection_s"] = NK * nb * (a_ps + nb * c_dot)
    full["compensation_terms_s"] = NK * nb * (a_cp + nb * c_cp * nsS)
    total = sum(full.values())
    return {"full_workload_s": total, "stages_full_s": full, "pairs_per_s": nb * nb * NK / total,
            "coefficients": {"setup_site_s_per_site": c_site, "transform_s_per_band": c_f,
                             "projection_s_per_band_site": c_ps, "dot_product_s_per_pair": c_dot,
                             "compensation_s_per_pair_site": c_cp, "call_overhead_s": [a_ps, a_cp],
                             "threads": threads},
            "sample_s": t}


def sample_text(w, plan, threads):
    return ("unmodified reference C (oracle/_ref: gcc -O2 -fopenmp, MKL DFTI, %d threads) on the first %d bands of "
            "(k,spin) block %d restricted to %d basis / %d wf sites (vacancy + %d per element): setup_projections x2, "
            "overlap_setup_real, %d x pseudoprojection and all compensation_terms rows against %d basis bands; full-workload "
            "time from the measured per-site / per-band / per-pair costs (ref_model)"
            % (threads, plan["nb"], plan["kappa"], len(plan["R_sub"]), len(plan["S_sub"]),
               (len(plan["R_sub"]) - 1) // len(w["elements"]), plan["n_pair"], plan["nb"]))


def ref_full(w, imgs, threads):
    """The complete workload through the reference C (all bands, all sites, every (k,spin) block in the images)."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import ref_driver as rd
    t = {}
    with _Quiet():
        t0 = time.perf_counter()
        R = rd.RefWavefunction(imgs[0][0], w["kws"])
        S = rd.RefWavefunction(imgs[1][0], w["kws"])
        t["read"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        R.setup_projections(w["pps"], w["labels_R"], w["coords_R"], w["dim"], w["grid_encut"])
        S.setup_projections(w["pps"], w["labels_S"], w["coords_S"], w["dim"], w["grid_encut"])
        t["setup"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        pr = rd.RefProjector(S, R, w["site_cat"])
        t["overlap_setup"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        for b in range(S.nband):
            pr.single_band_projection(b)
        t["pairs"] = time.perf_counter() - t0
        R.free()
        S.free()
    t["total_excl_read"] = t["setup"] + t["overlap_setup"] + t["pairs"]
    return t


def gpu_sample(w, simgs, plan):
    """The same sample through the B200 library (the reference-facing C ABI via pawpyc)."""
    from pawpyseed_b200 import _lib, pawpyc
    L = _lib.lib()
    k = plan["k_index"]
    kpts, kws = w["kpts"][k:k + 1], np.ones(1)
    L.pawb200_set_read_shard(0, 1)
    labR, labS = w["labels_R"][plan["R_sub"]], w["labels_S"][plan["S_sub"]]
    crdR, crdS = w["coords_R"][plan["R_sub"]], w["coords_S"][plan["S_sub"]]

    def mk(img, lab, crd):
        wf = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(img, kpts, kws))
        wf.projector_owner = 0
        wf._c_projector_setup(len(w["pps"]), len(lab), w["grid_encut"], np.ascontiguousarray(lab),
                              np.ascontiguousarray(crd), w["dim"], w["pps"])
        return wf
    R, S = mk(simgs[0], labR, crdR), mk(simgs[1], labS, crdS)
    pr = pawpyc.CProjector(S, R)
    pr._setup_overlap(plan["cat"], False)
    nb, n_pair, nchk = plan["nb"], plan["n_pair"], min(plan["nb"], 8)
    out = {"pseudo": np.array([S.pseudoprojection(b, R) for b in range(n_pair)])}
    comp = []
    for b in range(n_pair):
        res = np.zeros(nb, np.complex128)
        pr._add_augmentation_terms(res, b, False)
        comp.append(res)
    out["comp"] = np.array(comp)
    out["P_R"] = np.array([R._get_projections(b, 0) for b in range(nchk)])
    out["P_S"] = np.array([S._get_projections(b, 0) for b in range(nchk)])
    out["W_S"] = np.array([S._get_projections(b, 0, 3) for b in range(nchk)])
    out["chan_R"] = R._get_channel_index()
    out["idx_R"] = [R._get_site_indices(s) for s in index_check_sites(plan)]
    return out


def parity_report(ref, got):
    """GPU vs reference C on the sample: FP64 stages at 1e-10 relative, the pseudo overlap at the reference's own
    single precision (it accumulates and stores the dot product as float complex, pseudoprojector.c:82-87),
    index arrays bit-exact."""
    def rel(a, b):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape:
            return float("inf")
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0
    chan_ok = np.array_equal(np.asarray(got["chan_R"]).reshape(-1), np.asarray(ref["chan_R"]).reshape(-1))
    idx_ok = len(got["idx_R"]) == len(ref["idx_R"]) and all(
        np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(got["idx_R"], ref["idx_R"]))
    rep = {"projections_max_rel": max(rel(got["P_R"], ref["P_R"]), rel(got["P_S"], ref["P_S"])),
           "wave_projections_max_rel": rel(got["W_S"], ref["W_S"]),
           "compensation_terms_max_rel": rel(got["comp"], ref["comp"]),
           "pseudoprojection_max_abs": float(np.abs(got["pseudo"] - ref["pseudo"]).max()),
           "index_exact": bool(chan_ok and idx_ok),
           "tolerance": {"fp64_stages_rel": 1e-10, "pseudo_abs": 5e-6},
           "compared": "projections of %d bands x %d sites, %d compensation_terms rows, %d pseudoprojection rows, "
                       "channel (site,n,l,m) table, %d sphere index lists" %
                       (len(ref["P_R"]), len(ref["idx_R"]), len(ref["comp"]), len(ref["pseudo"]), len(ref["idx_R"]))}
    rep["ok"] = bool(rep["index_exact"] and rep["projections_max_rel"] < 1e-10 and
                     rep["wave_projections_max_rel"] < 1e-10 and rep["compensation_terms_max_rel"] < 1e-10 and
                     rep["pseudoprojection_max_abs"] < 5e-6)
    return rep


def cpu_defaults(args, w):
    """Sample size per workload: large enough that every timed reference call runs for >= 0.1 s."""
    if args.cpu_bands:
        nb = args.cpu_bands
    else:
        nb = {"cfg3": 192, "cfg2": 192}.get(w["key"], 16)
    ns_each = args.cpu_sites or {"cfg3": 8, "cfg2": 24}.get(w["key"], 2)
    return sample_plan(w, ns_each, nb, n_pair=args.cpu_pair_bands or 24, kappa=0)


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled every 10 ms DURING the timed region (NVML)."""

    def __init__(self, gpu_index=0):
        self.idx, self.rows, self._stop, self.th = gpu_index, [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    u = pynvml.nvmlDeviceGetUUID(h)
                    u = u.decode() if isinstance(u, bytes) else u
                    if uuid in u:
                        self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, r))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()

    def stop(self):
        self._stop = True
        if self.th:
            self.th.join(timeout=1)
        if not self.nv or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median([s for s, _ in self.rows])), "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.rows)}


def fp64_gemm_peak_tflops():
    import torch
    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    torch.cuda.empty_cache()
    return best


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key, nband, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu captures (profiles/r0*_traffic.json); None when
    this run's shape is not a captured one."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            e = t.get(key, {})
            if e.get("nband", 600 if key == "cfg2" else None) == nband and kernel in e:
                return e[kernel]
        except Exception:
            pass
    return None


def bind_to_gpu_numa_node(gpu_index, world):
    """Pin this rank's threads to a slice of the CPUs NVML reports as local to its GPU, so the pinned WAVECAR
    images are first-touched on that NUMA node and eight ranks do not pull their H2D traffic across sockets."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
        handle = None
        for i in range(pynvml.nvmlDeviceGetCount()):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            u = pynvml.nvmlDeviceGetUUID(h)
            if uuid in (u.decode() if isinstance(u, bytes) else u):
                handle = h
        if handle is None:
            return
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = [64 * wi + b for wi, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if len(allowed) >= 2:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


def measure_b200(args, w, rank, world, local, steps, warmup, scaling, e2e_steps):
    """Times the hot path of workload `w` on this rank's GPU (resident + end to end).  Returns the per-rank
    measurement dict; rank 0's also carries the gathered result of the last step."""
    import torch
    import torch.distributed as dist
    from pawpyseed_b200 import _lib, pawpyc
    from pawpyseed_b200 import distributed as pdist
    L = _lib.lib()
    nband, NK = w["nband"], w["nk"] * w["nspin"]
    bands = shard_mode(w, world) == "bands"
    if bands:       # every rank holds every (k,spin) block, but only its band block of it
        own = set(range(NK))
        L.pawb200_set_read_shard(0, 1)
        L.pawb200_set_band_shard(rank, world)
    else:
        own = {k for k in range(NK) if k % world == rank}
        L.pawb200_set_read_shard(rank, world)
        L.pawb200_set_band_shard(0, 1)
    imgs = make_images(w, own=own, pinned=True)
    h2d_bytes = sum(2 * 8 * nband * len(w["gvecs"][k % w["nk"]]) for k in range(NK))   # both structures, all ranks
    d2h_bytes = 16 * nband * nband * NK
    pairs_total = nband * nband * NK
    per = -(-NK // world)
    gather_pin = torch.empty(world * per * nband * nband * 2, dtype=torch.float64).pin_memory() if world > 1 else None

    def read(i):
        pwf = pawpyc.PWFPointer.from_arrays(imgs[i][0], w["kpts"], w["kws"])
        return pawpyc.CWavefunction(pwf)

    def setup(obj, which):
        lab, crd = (w["labels_R"], w["coords_R"]) if which == 0 else (w["labels_S"], w["coords_S"])
        obj.projector_owner = 0
        obj._c_projector_setup(len(w["pps"]), len(lab), w["grid_encut"], lab, crd, w["dim"], w["pps"])

    def hot_path(basis, wf, do_setup=True):
        if do_setup:
            setup(basis, 0)
            setup(wf, 1)
        pr = pawpyc.CProjector(wf, basis)
        pr._setup_overlap(w["site_cat"], False)
        if world == 1:
            return pr._projection_matrix()     # [NK][nbS][nbR] on host
        if bands:   # all-gather the basis rows, multiply the own wf rows, all-gather the result rows (NCCL)
            return pdist.band_sharded_projection_matrix(pr, pinned_out=gather_pin, want_host=(rank == 0))
        # one process per GPU: compute only the owned (k,spin) blocks, then all-gather the per-k matrices over NCCL
        ks = sorted(own)
        return pdist.gather_projection_blocks(pr, ks, NK, pinned_out=gather_pin, want_host=(rank == 0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input measurement (value) -------------------------------------------------
    basis, wf = read(0), read(1)
    for _ in range(warmup):
        res = hot_path(basis, wf)
    sampler = ClockSampler(local)
    barrier()
    _lib.reset_timers()
    sampler.start()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        res = hot_path(basis, wf)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    tm = _lib.timers()
    ms = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / steps
    checksum = float(np.abs(res).sum()) if rank == 0 else 0.0
    diag = float(np.abs(np.diagonal(res, axis1=1, axis2=2)).mean()) if rank == 0 else 0.0

    # ---- end-to-end from host images (e2e) -----------------------------------------------------
    del basis, wf
    L.pawb200_set_async_ingest(1)   # the pinned images outlive the wavefunctions; H2D overlaps the transforms

    def e2e_step():
        # the reference flow Wavefunction(..., setup_projectors=True) for basis then wf, then Projector(wf, basis):
        # the second WAVECAR's H2D (copy stream) overlaps the first structure's kernels
        basis = read(0)
        setup(basis, 0)
        wf = read(1)
        setup(wf, 1)
        out = hot_path(basis, wf, do_setup=False)
        del basis, wf
        return out

    if warmup > 0:
        e2e_step()      # one untimed pass: side-stream / staging buffers of the asynchronous path are created here
    barrier()
    tracing = bool(os.environ.get("PAWB200_TRACE"))
    if tracing:
        _lib.timers()       # flushes the device trace of the resident-input phase to stderr
        sys.stderr.write("==== e2e steps (rank %d)\n" % rank)
    f0, f1 = torch.cuda.Event(True), torch.cuda.Event(True)
    f0.record()
    for _ in range(e2e_steps):
        e2e_step()
    f1.record()
    barrier()
    if tracing:
        _lib.timers()
    ms2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item()) / e2e_steps
    L.pawb200_set_async_ingest(0)
    L.pawb200_set_band_shard(0, 1)
    L.pawb200_set_read_shard(0, 1)

    stage = {k: v / steps for k, v in tm.items() if k.endswith("_ms")}
    if os.environ.get("PAWB200_BENCH_DEBUG") or world > 1:
        sys.stderr.write("[rank %d] blocks %s dev_ms/step %.2f host_wall_ms/step %.2f stage_ms %s\n" % (
            rank, sorted(own), dev_ms / steps, wall * 1e3 / steps, {k: round(v, 2) for k, v in stage.items()}))
    # per-rank stage split, gathered on rank 0 (N > 1): separates load imbalance from the exchange tail
    per_rank = None
    if world > 1:
        keys = sorted(stage)
        mine = torch.tensor([dev_ms / steps, wall * 1e3 / steps] + [stage[k] for k in keys], dtype=torch.float64,
                            device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        per_rank = [dict(zip(["device_ms", "host_wall_ms"] + keys, [round(float(x), 3) for x in v.tolist()]))
                    for v in allv]
    return dict(ms_per_step=ms_per_step, value=pairs_total / (ms_per_step * 1e-3), e2e_ms=e2e_ms,
                e2e_value=pairs_total / (e2e_ms * 1e-3), h2d_bytes=h2d_bytes, d2h_bytes=d2h_bytes, tm=tm, stage=stage,
                wall_ms=wall * 1e3 / steps, clocks=clocks, checksum=checksum, mean_abs_diagonal=diag, own=own,
                imgs=imgs, pairs_total=pairs_total, per_rank=per_rank, e2e_steps=e2e_steps, steps=steps,
                band_rows=(-(-nband // world) if bands else nband))


def rooflines(w, m, fp64_peak, hbm_peak, hbm_src):
    """`roofline` (dominant kernel) and `kernels` (the HBM-bound stages) from the engine's CUDA-event stage timers."""
    tm, steps = m["tm"], m["steps"]
    nband = w["nband"]
    own = sorted(m["own"])
    npw_own = [len(w["gvecs"][k % w["nk"]]) for k in own]
    npw_mean = float(np.mean(npw_own)) if npw_own else 0.0
    ngrid = int(np.prod(w["dim"]))
    gemm_launches = steps * len(own)
    mrows = m.get("band_rows", nband)                                # wf rows of this rank's GEMMs (band-sharded: 1/N)
    gemm_flops = 8.0 * mrows * nband * npw_mean                      # SURVEY 8d, per launch (4 real products)
    gemm_ms = tm["gemm_pseudo_ms"] / max(gemm_launches, 1)
    gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    kern = {}
    if tm["scatter_ms"] > 0:
        kern["scatter_pw"] = {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "boxes": tm["boxes_scattered"],
                              "achieved": (12.0 * npw_mean + 16.0 * ngrid) * tm["boxes_scattered"] /
                              (tm["scatter_ms"] * 1e-3) / 1e9}
    if tm["fft_ms"] > 0:
        # SURVEY 8d per band: scatter 12 npw + 16 N, FFT 32 N.  With the pruned, scatter-fused transform both
        # stages are one kernel sequence, so they are reported together against the sum of the two figures.
        fused = tm["scatter_ms"] == 0
        per_box = (12.0 * npw_mean + 48.0 * ngrid) if fused else 32.0 * ngrid
        kern["pruned_fft3d" if fused else "cufft_z2z_3d"] = {
            "bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "boxes": tm["boxes_fft"], "library": not fused,
            "algorithmic_bytes_per_box": per_box, "ms_per_step": tm["fft_ms"] / steps,
            "achieved": per_box * tm["boxes_fft"] / (tm["fft_ms"] * 1e-3) / 1e9}
    if tm["project_ms"] > 0:
        # each sphere sample of psi~ (16 B) read once per band
        kern["sphere_project"] = {"bound": "hbm (nlm <= 11) / FP64 tensor above", "unit": "GB/s", "peak": hbm_peak,
                                  "slots": tm["slots_projected"], "sphere_samples_gathered": tm["sphere_samples"],
                                  "ms_per_step": tm["project_ms"] / steps,
                                  "achieved": 16.0 * tm["sphere_samples"] / (tm["project_ms"] * 1e-3) / 1e9}
    for k in kern.values():
        k["frac"] = k["achieved"] / k["peak"]
    total_stage = sum(v for k, v in tm.items() if k.endswith("_ms"))
    use4m = bool(os.environ.get("PAWB200_GEMM_4M"))
    bn = 64 if use4m else 48
    pad_m, pad_n = -(-mrows // 64) * 64, -(-nband // bn) * bn
    issued = (8.0 if use4m else 6.0) * pad_m * pad_n * npw_mean
    issued_tflops = issued / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    roof = {"kernel": "zgemm_abh_kernel<float2,%s> (pseudo overlap, DMMA.8x8x4, stream-K) + fixup" %
                      ("4M" if use4m else "3M"),
            "algorithm": "4 real products" if use4m else "3M (Karatsuba): 3 real DMMA products per complex product",
            "bound": "tensor", "achieved": gemm_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": (gemm_tflops / fp64_peak) if gemm_tflops else None,
            "frac_algorithmic": (gemm_tflops / fp64_peak) if gemm_tflops else None,
            "frac_issued": (issued_tflops / fp64_peak) if issued_tflops else None,
            "dmma_flops_issued_per_launch": issued,
            "traffic": ncu_traffic(w["key"], nband, "zgemm_abh_kernel<float2,3M>"),
            "peak_source": "torch.matmul fp64 6144^3 (cuBLAS DGEMM) measured in this run; "
                           "MEASURED_PEAKS.json has no FP64 entry; HBM peak %s" % hbm_src,
            "flops_per_launch": gemm_flops, "ms_per_launch": gemm_ms, "launches": gemm_launches,
            "share_of_step": tm["gemm_pseudo_ms"] / total_stage if total_stage else None}
    return roof, kern


def cpu_leg(args, w, imgs, want_parity):
    """cpu_baseline (+ parity) on the sample of workload w."""
    from oracle import ref_driver as rd
    threads = os.cpu_count() or 1
    if not rd.available():
        return ({"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                 "sample": "oracle/_ref/libpawpy_ref.so not present"}, None)
    plan = cpu_defaults(args, w)
    simgs = sample_images(w, imgs, plan)
    t, ref = ref_sample(w, simgs, plan, threads, collect=want_parity, warm=True)
    model = ref_model(w, plan, t, threads)
    cpu = {"value": model["pairs_per_s"], "unit": UNIT, "cores": threads, "kind": "reference",
           "sample": sample_text(w, plan, threads), "model": model}
    par = None
    if want_parity:
        par = parity_report(ref, gpu_sample(w, simgs, plan))
    return cpu, par


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pawpyseed_b200 import _lib

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    N = args.gpus
    if world != N and world != 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (N, world))
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local, world)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    if L.pawb200_device_check() != 0:
        raise SystemExit("pawpyseed_b200: " + L.pawb200_last_error().decode())
    host_threads = int(os.environ.get("PAWB200_BENCH_THREADS", max(1, (os.cpu_count() or 1) // world)))
    L.pawb200_set_host_threads(host_threads)   # torchrun exports OMP_NUM_THREADS=1

    weak = args.weak and world > 1      # --weak: config 2 replicated over N k-points, one per GPU (round-1 curve)
    scaling = "weak" if weak else "strong"
    w = workload(args.config, nk=world if weak else 1, nband=args.nband, blocks=args.blocks)
    e2e_steps = max(1, min(args.steps, 3))
    m = measure_b200(args, w, rank, world, local, args.steps, args.warmup, scaling, e2e_steps)
    if rank != 0:
        unpin_images(m["imgs"])
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, hbm_src = load_peaks()
    fp64_peak = fp64_gemm_peak_tflops()
    roof, kern = rooflines(w, m, fp64_peak, hbm_peak, hbm_src)
    cpu, parity = (None, None)
    if not args.no_cpu and world == 1:
        cpu, parity = cpu_leg(args, w, m["imgs"], want_parity=True)
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(w, world, scaling),
        "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": m["h2d_bytes"],
                "d2h_bytes_per_step": m["d2h_bytes"], "steps": e2e_steps, "ms_per_step": m["e2e_ms"]},
        "gpu_launches": int(m["tm"]["launches"]),
        "clocks": m["clocks"],
        "roofline": roof,
        "kernels": kern,
        "stage_ms_per_step": m["stage"],
        "host_wall_ms_per_step": m["wall_ms"],
        "cpu_baseline": cpu,
        "parity": parity,
        "checksum": m["checksum"], "mean_abs_diagonal": m["mean_abs_diagonal"],
    }
    if m["per_rank"]:
        line["per_rank_ms_per_step"] = m["per_rank"]
    unpin_images(m["imgs"])
    del m
    # ---- secondary workload: config 2 on one GPU ----------------------------------------------------------------
    if world == 1 and args.config == "cfg3" and not args.no_secondary:
        w2 = workload("cfg2", nk=1)
        k2, w2s = min(args.steps, 10), min(args.warmup, 3)
        m2 = measure_b200(args, w2, 0, 1, local, k2, w2s, "strong", max(1, min(k2, 3)))
        roof2, kern2 = rooflines(w2, m2, fp64_peak, hbm_peak, hbm_src)
        cpu2, par2 = (None, None)
        if not args.no_cpu:
            cpu2, par2 = cpu_leg(args, w2, m2["imgs"], want_parity=True)
        unpin_images(m2["imgs"])
        line["cfg2"] = {"config": config_dict(w2, 1, "strong"), "value": m2["value"], "unit": UNIT,
                        "ms_per_step": m2["ms_per_step"], "steps": k2, "warmup": w2s,
                        "e2e": {"value": m2["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": m2["h2d_bytes"],
                                "d2h_bytes_per_step": m2["d2h_bytes"], "ms_per_step": m2["e2e_ms"]},
                        "roofline": roof2, "kernels": kern2, "stage_ms_per_step": m2["stage"],
                        "checksum": m2["checksum"], "mean_abs_diagonal": m2["mean_abs_diagonal"],
                        "cpu_baseline": cpu2, "parity": par2}
    print(json.dumps(line))
    if parity is not None and not parity["ok"]:
        sys.stderr.write("PARITY FAILURE against the reference C: %s\n" % json.dumps(parity))
        sys.exit(3)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref = unmodified reference C; never on the product path)
# --------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    weak = args.weak and world > 1
    scaling = "weak" if weak else "strong"
    w = workload(args.config, nk=world if weak else 1, nband=args.nband)
    threads = os.cpu_count() or 1
    from oracle import ref_driver as rd
    if not rd.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpawpy_ref.so missing"}))
        return
    plan = cpu_defaults(args, w)
    t_all0 = time.perf_counter()
    if args.ref_full:
        imgs = make_images(w, use_gpu=False)
        walls = []
        for i in range(args.warmup + args.steps):
            t = ref_full(w, imgs, threads)
            if i >= args.warmup:
                walls.append(t["total_excl_read"])
        step_s = float(np.mean(walls))
        value = w["nband"] ** 2 * w["nk"] * w["nspin"] / step_s
        model, sample = {"measured_full_s": t}, ("the COMPLETE workload through the reference C "
                                                 "(oracle/_ref, %d threads), no sampling" % threads)
    else:
        # images of the sample only: the first nb bands of (k,spin) block 0, same generator as the B200 arm
        ws = dict(w)
        ws["kpts"], ws["gvecs"], ws["nk"], ws["nspin"] = w["kpts"][:1], w["gvecs"][:1], 1, 1
        ws["kws"] = np.ones(1)
        imgs = make_images(ws, nband=plan["nb"], use_gpu=False)      # the reference arm never touches the GPU
        ws["nband"] = plan["nb"]
        simgs = sample_images(ws, imgs, plan)
        walls, models = [], []
        for i in range(args.warmup + args.steps):
            s0 = time.perf_counter()
            t, _ = ref_sample(w, simgs, plan, threads, warm=(i == 0))
            if i >= args.warmup:
                walls.append(time.perf_counter() - s0)
                models.append(ref_model(w, plan, t, threads))
        step_s = float(np.mean(walls))
        models.sort(key=lambda mm: mm["pairs_per_s"])
        model = models[len(models) // 2]
        value = model["pairs_per_s"]
        sample = sample_text(w, plan, threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(w, world, scaling),
            "value_is": "measured" if args.ref_full else
                        "full-workload pairs/s from the reference's cost model with every coefficient measured in "
                        "this run's sample steps; ms_per_step is the measured wall time of one sample step",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": sample, "model": model},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# BASELINE configs 4 and 5: the spinor projection path and the real-space density path (own metrics)
# --------------------------------------------------------------------------------------------
def _device_timed(fn, steps, warmup):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def run_aux(args):
    """--config cfg4: noncollinear 128-site cell, 512 spinor bands x 2 k-points - one step = setup_projections
    (two transforms + up/down projector overlaps per band, projector.c:372-418).
    --config cfg5: ae_chg_density on a 400^3 grid, 256 bands (128 occupied), 64 sites - one step = the whole
    density (density.c:158-179: realspace_state + |psi|^2 accumulation per occupied band, grid to the host)."""
    import torch
    from pawpyseed_b200 import _lib, pawpyc
    from oracle import ref_driver as rd
    torch.cuda.set_device(0)
    L = _lib.lib()
    if L.pawb200_device_check() != 0:
        raise SystemExit("pawpyseed_b200: " + L.pawb200_last_error().decode())
    L.pawb200_set_host_threads(os.cpu_count() or 1)
    hbm_peak, hbm_src = load_peaks()
    threads = os.cpu_count() or 1
    if args.config == "cfg4":
        nband = args.nband or 512
        lat, frac, lab = synth.wurtzite_supercell((4, 2, 2))            # 128 sites
        lab = lab.astype(np.int32)
        encut, kpts, kws = 400.0, np.array([[0.0, 0.0, 0.0], [0.5, 0.0, 0.0]]), np.array([0.5, 0.5])
        gv = [synth.enumerate_gvectors(lat, encut, k) for k in kpts]
        dim = synth.fft_grid_for(gv)
        ge = synth.grid_encut(dim, lat)
        pps = synth.synthetic_pps(["Ga", "N"])
        img = synth.wavecar_image(lat, encut, kpts, 1, nband, synth.random_coeffs(4, nband), ncl=True, gvecs=gv)

        def read():
            return pawpyc.CNCLWavefunction(pawpyc.PWFPointer.from_arrays(img, kpts, kws))

        def setup(wf):
            wf.projector_owner = 0
            wf._c_projector_setup(len(pps), len(lab), ge, lab, frac, dim, pps)
            return wf._get_projections(0, 0, 1)
        wf = read()
        _lib.reset_timers()
        ms, _ = _device_timed(lambda: setup(wf), args.steps, args.warmup)
        tm = _lib.timers()
        n = args.steps + args.warmup
        units = nband * len(kpts)
        e2e_ms, _ = _device_timed(lambda: setup(read()), max(1, min(args.steps, 3)), 1)
        t0 = time.perf_counter()
        wf._get_realspace_state(0, 1, 0)
        first_state = time.perf_counter() - t0
        st_ms, _ = _device_timed(lambda: wf._get_realspace_state(1, 1, 0), 4, 1)
        ngrid, npw = int(np.prod(dim)), float(np.mean([len(g) for g in gv]))
        per_box = 12.0 * npw + 48.0 * ngrid
        cpu = None
        if not args.no_cpu and rd.available():
            nb_s = 32
            simg = synth.wavecar_image(lat, encut, kpts, 1, nb_s, synth.random_coeffs(4, nb_s), ncl=True, gvecs=gv)
            os.environ["OMP_NUM_THREADS"] = str(threads)
            with _Quiet():
                R = rd.RefWavefunction(simg, kws)
                t0 = time.perf_counter()
                R.setup_projections(pps, lab, frac, dim, ge)
                t_all = time.perf_counter() - t0
                t_site = R.time_projector_values()
                R.free()
            full = t_site + (t_all - t_site) * nband / nb_s
            cpu = {"value": units / full, "unit": "spinor bands/s", "cores": threads, "kind": "reference",
                   "sample": "unmodified reference C setup_projections on %d of %d spinor bands x 2 k-points, all 128 "
                             "sites: %.2f s of which the serial setup_site %.2f s; band part scaled by %d/%d"
                             % (nb_s, nband, t_all, t_site, nband, nb_s)}
        line = {"metric": "ncl_spinor_band_projections_per_sec", "value": units / (ms * 1e-3), "unit": "spinor bands/s",
                "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "noncollinear GaN 128-site cell, ENCUT 400, %d spinor bands, 2 k-points, grid %s"
                                       % (nband, [int(x) for x in dim]), "npw_per_spinor_half": [len(g) for g in gv],
                           "l2": "FFT boxes of a launch (8 groups x 16 slots) exceed the 126 MB L2"},
                "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "spinor bands/s", "h2d_bytes_per_step": int(img.nbytes),
                        "d2h_bytes_per_step": 16 * int(L.pawb200_num_projections(wf.wf_ptr, 1)), "ms_per_step": e2e_ms},
                "gpu_launches": int(tm["launches"]),
                "roofline": {"kernel": "pruned fft3d (pass Z + Y + X), two transforms per spinor band", "bound": "hbm",
                             "achieved": per_box * tm["boxes_fft"] / (tm["fft_ms"] * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "traffic": None, "peak_source": hbm_src,
                             "share_of_step": tm["fft_ms"] / n / ms},
                "stage_ms_per_step": {k: v / n for k, v in tm.items() if k.endswith("_ms")},
                "ncl_realspace_state": {"first_call_ms": first_state * 1e3, "ms_per_state": st_ms,
                                        "bytes_to_host_per_state": 2 * 16 * ngrid},
                "cpu_baseline": cpu}
        line["roofline"]["frac"] = line["roofline"]["achieved"] / hbm_peak
        print(json.dumps(line))
        return
    # ---- cfg5 ----
    nband = args.nband or 256
    dim = np.array([args.grid // 2] * 3, np.int32)                        # fine grid = 2 * dim (pawpyc.pyx:455-461)
    lat, coords = synth.diamond_supercell(5.43, 2)                        # 64 Si sites
    encut, kpts, kws = 300.0, np.array([[0.0, 0.0, 0.0]]), np.array([1.0])
    gv = [synth.enumerate_gvectors(lat, encut, kpts[0])]
    pps = synth.synthetic_pps(["Si"])
    labels = np.zeros(len(coords), np.int32)
    ge = synth.grid_encut(dim, lat)
    img = synth.wavecar_image(lat, encut, kpts, 1, nband, synth.random_coeffs(5, nband), gvecs=gv)
    nocc = (nband + 1) // 2

    def make():
        wf = pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(img, kpts, kws))
        wf.projector_owner = 0
        wf._c_projector_setup(1, len(coords), ge, labels, coords, dim, pps)
        return wf
    wf = make()
    t0 = time.perf_counter()
    wf._get_realspace_density()
    first = time.perf_counter() - t0                                      # builds the AE partial-wave tables
    _lib.reset_timers()
    ms, rho = _device_timed(wf._get_realspace_density, args.steps, min(args.warmup, 1))
    tm = _lib.timers()
    n = args.steps + min(args.warmup, 1)
    e2e_ms, _ = _device_timed(lambda: make()._get_realspace_density(), 1, 0)
    ngrid, npw = int(rho.size), len(gv[0])
    per_state = 12.0 * npw + 48.0 * ngrid + 32.0 * ngrid                  # scatter + FFT + density RMW
    vol = abs(np.linalg.det(lat))
    cpu = None
    if not args.no_cpu and rd.available():
        nb_s = 4                                                          # 2 occupied bands
        simg = synth.wavecar_image(lat, encut, kpts, 1, nb_s, synth.random_coeffs(5, nb_s), gvecs=gv)
        os.environ["OMP_NUM_THREADS"] = str(threads)
        with _Quiet():
            R = rd.RefWavefunction(simg, kws)
            R.setup_projections(pps, labels, coords, dim, ge)
            t0 = time.perf_counter()
            R.chg_density()
            t_ref = time.perf_counter() - t0
            R.free()
        cpu = {"value": 2 / t_ref, "unit": "AE states/s", "cores": threads, "kind": "reference",
               "sample": "unmodified reference C ae_chg_density on the same cell and %d^3 grid with 2 occupied bands: "
                         "%.1f s (serial over bands, density.c:158-179)" % (args.grid, t_ref)}
    line = {"metric": "ae_states_per_sec", "value": nocc / (ms * 1e-3), "unit": "AE states/s", "n_gpus": 1,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ae_chg_density (write_density path) on a %d^3 grid, %d bands (%d occupied), "
                                   "64 Si sites, ENCUT 300, Gamma" % (args.grid, nband, nocc), "npw": npw,
                       "l2": "every box (%.2f GB) exceeds the 126 MB L2" % (16 * ngrid / 1e9)},
            "e2e": {"value": nocc / (e2e_ms * 1e-3), "unit": "AE states/s", "h2d_bytes_per_step": int(img.nbytes),
                    "d2h_bytes_per_step": 8 * ngrid, "ms_per_step": e2e_ms,
                    "note": "read_wavefunctions + setup_projections + AE tables + density, grid on the host"},
            "gpu_launches": int(tm["launches"]),
            "roofline": {"kernel": "pruned fft3d + Bloch phase + augmentation + |psi|^2 accumulation per state",
                         "bound": "hbm", "achieved": per_state * nocc / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "traffic": None, "peak_source": hbm_src,
                         "algorithmic_bytes_per_state": per_state},
            "stage_ms_per_step": {k: v / n for k, v in tm.items() if k.endswith("_ms")},
            "first_call_ms": first * 1e3,
            "density_integral": float(rho.sum() * vol / rho.size),
            "cpu_baseline": cpu}
    line["roofline"]["frac"] = line["roofline"]["achieved"] / hbm_peak
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--nband", type=int, default=None, help="override the band count (debug)")
    ap.add_argument("--blocks", type=int, default=None, help="keep only the first n (k,spin) blocks (debug)")
    ap.add_argument("--cpu-bands", type=int, default=0, help="bands per structure in the CPU sample")
    ap.add_argument("--cpu-sites", type=int, default=0, help="sites per element in the CPU sample")
    ap.add_argument("--cpu-pair-bands", type=int, default=0, help="wf bands whose pair rows the CPU sample computes")
    ap.add_argument("--ref-full", action="store_true", help="reference arm: time the complete workload (no sample)")
    ap.add_argument("--weak", action="store_true",
                    help="config 2 only: N GPUs = N k-points of the config's shape (weak scaling over (k,spin) blocks)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-2 secondary measurement")
    ap.add_argument("--grid", type=int, default=400, help="cfg5: fine grid points per axis")
    args = ap.parse_args()
    if args.config in ("cfg4", "cfg5"):
        if int(os.environ.get("RANK", 0)) == 0:
            if args.impl == "reference":
                print(json.dumps({"impl": "reference", "unavailable": "configs 4/5 report their CPU leg inside the "
                                  "b200 line (cpu_baseline)"}))
            else:
                run_aux(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
