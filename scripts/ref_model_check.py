#!/usr/bin/env python
"""Validate bench.ref_model (the reference arm's cost model) against a COMPLETE run of the unmodified
reference C on a reduced workload: same cell, grid and site count as the named config, fewer bands and one
(k,spin) block, so that the full run takes about a minute.

  python scripts/ref_model_check.py [--config cfg3|cfg2] [--nband 256] [--out profiles/r02_ref_model_check.json]

Prints / writes: measured full wall time per stage, the model's prediction for the same workload from one
sample pass, and their ratio.  CPU only (oracle/_ref)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--nband", type=int, default=256)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    w = bench.workload(a.config, nband=a.nband)
    # one (k,spin) block
    w["kpts"], w["gvecs"], w["nk"], w["nspin"], w["kws"] = w["kpts"][:1], w["gvecs"][:1], 1, 1, np.ones(1)
    threads = os.cpu_count() or 1
    imgs = bench.make_images(w, use_gpu=False)
    args = argparse.Namespace(cpu_bands=0, cpu_sites=0, cpu_pair_bands=0)
    plan = bench.cpu_defaults(args, w)
    simgs = bench.sample_images(w, imgs, plan)
    t, _ = bench.ref_sample(w, simgs, plan, threads, warm=True)
    model = bench.ref_model(w, plan, t, threads)
    t0 = time.perf_counter()
    full = bench.ref_full(w, imgs, threads)
    full["wall_incl_read"] = time.perf_counter() - t0
    pred = model["stages_full_s"]
    pred_setup = pred["setup_site_s"] + pred["setup_bands_s"]
    pred_pairs = pred["pseudoprojection_s"] + pred["compensation_terms_s"]
    rep = {"workload": w["name"] + " [reduced: %d bands, 1 (k,spin) block]" % w["nband"], "threads": threads,
           "measured_full_s": full,
           "model_s": {"setup": pred_setup, "overlap_setup": pred["overlap_setup_s"], "pairs": pred_pairs,
                       "total": model["full_workload_s"]},
           "model_over_measured": {"setup": pred_setup / full["setup"],
                                   "overlap_setup": pred["overlap_setup_s"] / full["overlap_setup"],
                                   "pairs": pred_pairs / full["pairs"],
                                   "total": model["full_workload_s"] / full["total_excl_read"]},
           "model_detail": model}
    print(json.dumps(rep, indent=1))
    if a.out:
        json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
