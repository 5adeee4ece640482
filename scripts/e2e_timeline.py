"""Host-side timeline of one end-to-end step: when each API call returns on the host, and when the device is done.
Usage on the GPU box: python scripts/e2e_timeline.py [steps] [config=cfg2] [blocks] [host threads]
(cfg3 with blocks=1 and 2 host threads is what one rank of the 8-GPU job does)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pawpyseed_b200 import _lib, pawpyc  # noqa: E402

L = _lib.lib()
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else None
w = bench.workload(cfg, blocks=blocks)
imgs = bench.make_images(w, pinned=True)
L.pawb200_set_host_threads(int(sys.argv[4]) if len(sys.argv) > 4 else os.cpu_count())
L.pawb200_set_async_ingest(1)


def read(i):
    return pawpyc.CWavefunction(pawpyc.PWFPointer.from_arrays(imgs[i][0], w["kpts"], w["kws"]))


def setup(obj, which):
    lab, crd = (w["labels_R"], w["coords_R"]) if which == 0 else (w["labels_S"], w["coords_S"])
    obj.projector_owner = 0
    obj._c_projector_setup(len(w["pps"]), len(lab), w["grid_encut"], lab, crd, w["dim"], w["pps"])


# raw pinned H2D bandwidth of this box, for reference
_src = torch.empty(1 << 29, dtype=torch.uint8).pin_memory()
_dst = torch.empty(1 << 29, dtype=torch.uint8, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); _dst.copy_(_src, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("pinned H2D 512 MiB: %.1f GB/s" % (_src.numel() / e0.elapsed_time(e1) / 1e6), flush=True)

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for step in range(nsteps):
    torch.cuda.synchronize()
    _lib.timers()            # flushes the PAWB200_TRACE marks of the previous step to stderr
    sys.stderr.write("==== step %d\n" % step)
    _lib.reset_timers()
    t0 = time.perf_counter()
    marks = []

    def mark(name):
        marks.append((name, (time.perf_counter() - t0) * 1e3))
    basis = read(0); mark("read(R)")
    setup(basis, 0); mark("setup(R)")
    wf = read(1); mark("read(S)")
    setup(wf, 1); mark("setup(S)")
    pr = pawpyc.CProjector(wf, basis)
    pr._setup_overlap(w["site_cat"], False); mark("overlap_setup")
    res = pr._projection_matrix(); mark("projection_matrix (synced)")
    del basis, wf, pr; mark("free")
    tm = _lib.timers()
    print("   timers", {k: round(v, 2) for k, v in tm.items() if k.endswith("_ms")}, flush=True)
    print("step", step, "  ".join("%s %.1f" % m for m in marks), flush=True)
