#!/usr/bin/env python
"""bench.py - PAW-corrected band-pair projections / second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config cfg2|cfg1|tiny]

One "step" = one pass of the whole hot path over one synthetic wavefunction pair:
setup_projections(basis) + setup_projections(wf) + overlap_setup_real + every band pair
(pseudo overlap GEMM + augmentation GEMM) + result to host.

* `value`  : pairs/s with the plane-wave coefficients already resident in HBM (device-timed, CUDA events).
* `e2e`    : pairs/s through the public API from HOST WAVECAR images in pinned memory
             (read_wavefunctions_from_str -> ... -> result matrix on host), H2D/D2H inside the timed region.
* `roofline`: the dominant kernel (stream-K DMMA complex GEMM of the pseudo overlap) against the FP64
             GEMM rate measured in this run (MEASURED_PEAKS.json has no FP64 entry); HBM-bound kernels are
             listed under `kernels` against MEASURED_PEAKS.json's copy bandwidth.
* `cpu_baseline`: the unmodified reference C (oracle/_ref) on this box's host cores, bounded band sample,
             extrapolated to the full workload with the reference's own cost model (setup ~ bands, pairs ~ bands^2).
* `--impl reference`: the reference arm - the same CPU measurement as the headline line.

N > 1 (torchrun, one rank per GPU): weak scaling over k-points - the job has N (k,spin) blocks of the
config's shape, rank r owns block r (no data-path collective), and the per-k result blocks are
all-gathered over NCCL at the end of each step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
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
def workload(name, nk=1, nband=None):
    """Returns dict(lattice, encut, kpts, nspin, nband, basis/wf coords+labels, elements, site_cat, dim)."""
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
                      "k in {0, b1/2, b2/2, b3/2}", lattice=lat, encut=520.0, nband=2000, nspin=2,
                 elements=["Ga", "N"], vac=vac, coords_R=frac, labels_R=lab.astype(np.int32),
                 coords_S=np.delete(frac, vac, axis=0), labels_S=np.delete(lab, vac).astype(np.int32),
                 kpt_list=[[0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [0.0, 0.5, 0.0], [0.0, 0.0, 0.5]],
                 dim=np.array([144, 126, 60], np.int32))   # the PREC=Normal grid of this cell
    elif name == "tiny":    # CPU-sized smoke configuration of the same shape
        lat, bulk = synth.diamond_supercell(5.43, 1)
        _, defect = synth.diamond_supercell(5.43, 1, vacancy=3)
        w = dict(name="Si8 bulk x Si7+vacancy, ENCUT 250, 32 bands, Gamma", lattice=lat, encut=250.0,
                 nband=32, nspin=1, elements=["Si"], vac=3, coords_R=bulk,
                 labels_R=np.zeros(len(bulk), np.int32), coords_S=defect,
                 labels_S=np.zeros(len(defect), np.int32))
    else:
        raise SystemExit("unknown config %s" % name)
    if nband:
        w["nband"] = int(nband)
    if "kpt_list" in w:
        # weak scaling over (k,spin) blocks: `nk` is the number of blocks wanted (= GPUs); both spins of a k-point
        # first, then more k-points (8 GPUs = the full 4 k x 2 spins job)
        if nk == 1:
            w["nspin"] = 1
        nk = max(1, nk // w["nspin"])
        if nk > len(w["kpt_list"]):
            raise SystemExit("%s has only %d k-points" % (name, len(w["kpt_list"])))
        w["kpts"] = np.array(w["kpt_list"][:nk])
    else:
        # k-points: Gamma for nk == 1; for the weak-scaling job, nk distinct points along b1
        w["kpts"] = np.array([[0.0, 0.0, 0.0]] if nk == 1 else [[0.5 * i / nk, 0.0, 0.0] for i in range(nk)])
    w["nk"] = nk
    w["kws"] = np.full(nk, 1.0 / nk)
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


def ncu_traffic(config, nband, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r01_traffic.json); None when
    this run's shape is not the captured one."""
    try:
        t = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_traffic.json")))
        return t[config][kernel] if config == "cfg2" and nband == 600 else None
    except Exception:
        return None


def make_images(w, own=None, nband=None, pinned=False):
    """WAVECAR images (basis, wf).  own: set of kappa whose coefficient records are filled
    (others stay zero pages - sharded ranks never read them)."""
    import torch
    nband = nband or w["nband"]
    imgs = []
    for sid in (0, 1):
        def gen(kap, npw, _sid=sid):
            if own is not None and kap not in own:
                return None                                   # untouched zero pages
            return synth.random_coeffs(2000 + sid, nband)(kap, npw)
        img = synth.wavecar_image(w["lattice"], w["encut"], w["kpts"], w["nspin"], nband, gen, gvecs=w["gvecs"])
        if pinned:
            # page-lock only the coefficient records this rank reads (the image of a sharded job is mostly
            # other ranks' zero pages)
            nrecl = int(round(img[:8].view(np.float64)[0]))
            NK = len(w["kpts"]) * w["nspin"]
            rt = torch.cuda.cudart()
            for kap in (range(NK) if own is None else sorted(own)):
                lo = (2 + kap * (1 + nband)) * nrecl
                hi = lo + (1 + nband) * nrecl
                a0 = (img.ctypes.data + lo) // 4096 * 4096
                a1 = min(-(-(img.ctypes.data + hi) // 4096) * 4096, img.ctypes.data + img.nbytes)
                err = rt.cudaHostRegister(a0, a1 - a0, 0)
                if int(err) != 0:
                    raise SystemExit("cudaHostRegister failed: %s" % err)
        imgs.append((img, None))
    return imgs


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


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pawpyseed_b200 import _lib, pawpyc
    from pawpyseed_b200 import distributed as pdist

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

    w = workload(args.config, nk=world, nband=args.nband)
    nband, NK = w["nband"], w["nk"] * w["nspin"]
    own = {k for k in range(NK) if k % world == rank}
    L.pawb200_set_read_shard(rank, world)
    host_threads = int(os.environ.get("PAWB200_BENCH_THREADS", max(1, (os.cpu_count() or 1) // world)))
    L.pawb200_set_host_threads(host_threads)   # torchrun exports OMP_NUM_THREADS=1
    imgs = make_images(w, own=own, pinned=True)
    h2d_bytes = sum(2 * 8 * nband * len(w["gvecs"][k % w["nk"]]) for k in own)   # both structures
    d2h_bytes = 16 * nband * nband * len(own)
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
        # one process per GPU: compute only the owned (k,spin) blocks, then all-gather the per-k matrices over NCCL
        ks = sorted(own)
        blocks = [pr._projection_matrix(kappa_range=(k, k + 1)) for k in ks]
        mine = blocks[0] if len(blocks) == 1 else (np.concatenate(blocks) if blocks else
                                                   np.zeros((0, nband, nband), np.complex128))
        return pdist.all_gather_own_blocks(mine, ks, NK, pinned_out=gather_pin, want_host=(rank == 0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input measurement (value) -------------------------------------------------
    basis, wf = read(0), read(1)
    for _ in range(args.warmup):
        res = hot_path(basis, wf)
    sampler = ClockSampler(local)
    barrier()
    _lib.reset_timers()
    sampler.start()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
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
    ms_per_step = float(ms.item()) / args.steps
    value = pairs_total / (ms_per_step * 1e-3)
    checksum = float(np.abs(res).sum()) if rank == 0 else 0.0

    # ---- end-to-end from host images (e2e) -----------------------------------------------------
    del basis, wf
    e2e_steps = max(1, min(args.steps, 3))
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

    if args.warmup > 0:
        e2e_step()      # one untimed pass: side-stream / staging buffers of the asynchronous path are created here
    barrier()
    f0, f1 = torch.cuda.Event(True), torch.cuda.Event(True)
    f0.record()
    for _ in range(e2e_steps):
        res2 = e2e_step()
    f1.record()
    barrier()
    ms2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = pairs_total / (float(ms2.item()) / e2e_steps * 1e-3)

    if os.environ.get("PAWB200_BENCH_DEBUG"):
        sys.stderr.write("[rank %d] dev_ms/step %.2f wall_ms/step %.2f stages %s\n" % (
            rank, dev_ms / args.steps, wall * 1e3 / args.steps,
            {k: round(v / args.steps, 2) for k, v in tm.items() if k.endswith("_ms")}))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline ----------------------------------------------------------------------------------
    hbm_peak, hbm_src = load_peaks()
    fp64_peak = fp64_gemm_peak_tflops()
    npw = len(w["gvecs"][0])
    ngrid = int(np.prod(w["dim"]))
    steps = args.steps
    n_own = len(own)
    gemm_launches = steps * n_own
    gemm_flops = 8.0 * nband * nband * npw                      # SURVEY 8d, per launch
    gemm_ms = tm["gemm_pseudo_ms"] / max(gemm_launches, 1)
    gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    kern = {}
    if tm["scatter_ms"] > 0:
        kern["scatter_pw"] = {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "boxes": tm["boxes_scattered"],
                              "achieved": (12.0 * npw + 16.0 * ngrid) * tm["boxes_scattered"] /
                              (tm["scatter_ms"] * 1e-3) / 1e9}
    if tm["fft_ms"] > 0:
        # SURVEY 8d per band: scatter 12 npw + 16 N, FFT 32 N.  With the pruned, scatter-fused transform both
        # stages are one kernel sequence, so they are reported together against the sum of the two figures.
        fused = tm["scatter_ms"] == 0
        per_box = (12.0 * npw + 48.0 * ngrid) if fused else 32.0 * ngrid
        kern["pruned_fft3d_zyx" if fused else "cufft_z2z_3d"] = {
            "bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "boxes": tm["boxes_fft"], "library": not fused,
            "algorithmic_bytes_per_box": per_box,
            "achieved": per_box * tm["boxes_fft"] / (tm["fft_ms"] * 1e-3) / 1e9}
    if tm["project_ms"] > 0:
        # each sphere sample of psi~ (16 B) read once per band
        kern["sphere_project"] = {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "slots": tm["slots_projected"],
                                  "sphere_samples_gathered": tm["sphere_samples"],
                                  "achieved": 16.0 * tm["sphere_samples"] / (tm["project_ms"] * 1e-3) / 1e9}
    for k in kern.values():
        k["frac"] = k["achieved"] / k["peak"]
    total_stage = sum(v for k, v in tm.items() if k.endswith("_ms"))
    use4m = bool(os.environ.get("PAWB200_GEMM_4M"))
    bn = 64 if use4m else 48
    pad_m, pad_n = -(-nband // 64) * 64, -(-nband // bn) * bn
    roof = {"kernel": "zgemm_abh_kernel<float2,%s> (pseudo overlap, DMMA.8x8x4, stream-K) + fixup" %
                      ("4M" if use4m else "3M"),
            "algorithm": "4 real products" if use4m else "3M (Karatsuba): 3 real DMMA products per complex product",
            "dmma_flops_issued_per_launch": (8.0 if use4m else 6.0) * pad_m * pad_n * npw,
            "bound": "tensor", "achieved": gemm_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": (gemm_tflops / fp64_peak) if gemm_tflops else None,
            "traffic": ncu_traffic(args.config, nband, "zgemm_abh_kernel<float2,3M>"),
            "peak_source": "torch.matmul fp64 6144^3 (cuBLAS DGEMM) measured in this run; "
                           "MEASURED_PEAKS.json has no FP64 entry; HBM peak %s" % hbm_src,
            "flops_per_launch": gemm_flops, "ms_per_launch": gemm_ms,
            "share_of_step": tm["gemm_pseudo_ms"] / total_stage if total_stage else None}

    cpu = cpu_baseline(args, w) if not args.no_cpu else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"] + (" x %d (k,spin) blocks (one per GPU)" % NK if world > 1 else ""),
                   "nband": nband, "npw": npw, "fft_grid": [int(x) for x in w["dim"]],
                   "sites": [len(w["labels_R"]), len(w["labels_S"])], "kappa_blocks": NK,
                   "pairs_per_step": pairs_total, "parallelism": "kpoint-shard x%d" % world,
                   "l2": "inputs (%.2f GB coefficients + FFT boxes) exceed the 126 MB L2" %
                         (2 * 8 * nband * npw / 1e9)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps},
        "gpu_launches": int(tm["launches"]),
        "clocks": clocks,
        "roofline": roof,
        "kernels": kern,
        "stage_ms_per_step": {k: v / steps for k, v in tm.items() if k.endswith("_ms")},
        "host_wall_ms_per_step": wall * 1e3 / steps,
        "cpu_baseline": cpu,
        "checksum": checksum,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref = unmodified reference C; never on the product path)
# --------------------------------------------------------------------------------------------
def cpu_reference_measure(w, nb_sample, threads):
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import ref_driver as rd
    if not rd.available():
        return None
    ws = dict(w)
    imgs = make_images(ws, nband=nb_sample)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)                      # the reference printf()s progress lines
    try:
        t0 = time.perf_counter()
        R = rd.RefWavefunction(imgs[0][0], w["kws"])
        S = rd.RefWavefunction(imgs[1][0], w["kws"])
        t_read = time.perf_counter() - t0
        t0 = time.perf_counter()
        R.setup_projections(w["pps"], w["labels_R"], w["coords_R"], w["dim"], w["grid_encut"])
        S.setup_projections(w["pps"], w["labels_S"], w["coords_S"], w["dim"], w["grid_encut"])
        t_setup = time.perf_counter() - t0
        t0 = time.perf_counter()
        pr = rd.RefProjector(S, R, w["site_cat"])
        t_ov = time.perf_counter() - t0
        t0 = time.perf_counter()
        for b in range(nb_sample):
            pr.single_band_projection(b)
        t_pairs = time.perf_counter() - t0
        R.free()
        S.free()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return dict(read=t_read, setup=t_setup, overlap_setup=t_ov, pairs=t_pairs)


def cpu_extrapolate(w, nb_sample, t, nk_blocks):
    """Reference cost model (docs/techref.tex:223-275): setup and overlap_setup scale with the number of
    bands, the per-pair stage with bands^2; every (k,spin) block costs the same."""
    nb = w["nband"]
    scale1 = nb / nb_sample
    t_full = (t["setup"] + t["overlap_setup"]) * scale1 + t["pairs"] * scale1 ** 2
    pairs_full = nb * nb * nk_blocks
    return pairs_full / (t_full * nk_blocks), t_full


def cpu_baseline(args, w, nb_sample=None):
    threads = os.cpu_count() or 1
    nb_sample = nb_sample or min(w["nband"], args.cpu_bands)
    t = cpu_reference_measure(w, nb_sample, threads)
    if t is None:
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": "oracle/_ref/libpawpy_ref.so not present"}
    v, t_full = cpu_extrapolate(w, nb_sample, t, 1)
    return {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": "unmodified reference C (gcc -O2 -fopenmp, MKL DFTI) on %d of %d bands per structure, "
                      "1 (k,spin) block: setup_projections x2 %.2fs, overlap_setup_real %.2fs, %d x "
                      "(pseudoprojection+compensation_terms) %.3fs; extrapolated setup~bands, pairs~bands^2 "
                      "-> %.1f s per full step" % (nb_sample, w["nband"], t["setup"], t["overlap_setup"],
                                                   nb_sample, t["pairs"], t_full),
            "measured_s": t}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    w = workload(args.config, nk=1, nband=args.nband)
    threads = os.cpu_count() or 1
    nb_sample = min(w["nband"], args.cpu_bands)
    from oracle import ref_driver as rd
    if not rd.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpawpy_ref.so missing"}))
        return
    for _ in range(min(args.warmup, 1)):
        cpu_reference_measure(w, max(2, nb_sample // 4), threads)
    vals, ts = [], []
    t_all0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        t = cpu_reference_measure(w, nb_sample, threads)
        v, t_full = cpu_extrapolate(w, nb_sample, t, 1)
        vals.append(v)
        ts.append(t_full)
    value = float(np.median(vals))
    npw = len(w["gvecs"][0])
    sample = ("reference C (oracle/_ref, unmodified sources, gcc -O2 -fopenmp, %d threads) on %d of %d bands; "
              "extrapolated with setup~bands, pairs~bands^2" % (threads, nb_sample, w["nband"]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.median(ts)) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w["name"] + (" x %d k-points" % world if world > 1 else ""),
                       "nband": w["nband"], "npw": npw, "fft_grid": [int(x) for x in w["dim"]],
                       "pairs_per_step": w["nband"] ** 2 * world},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--nband", type=int, default=None, help="override the band count (debug)")
    ap.add_argument("--cpu-bands", type=int, default=24, help="bands per structure in the CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
