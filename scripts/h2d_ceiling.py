#!/usr/bin/env python
"""Aggregate host->device copy ceiling of the box: every rank copies a pinned 1 GiB buffer to its GPU in a loop at
the same time (torchrun, one rank per GPU); rank 0 prints per-rank and aggregate GB/s.  This is the bound of the
end-to-end (`e2e`) figure of bench.py, whose steps re-read the WAVECAR images from pinned host memory.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      scripts/h2d_ceiling.py [--seconds 3] [--numa-bind 0|1]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--numa-bind", type=int, default=1)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if a.numa_bind and world > 1:
        import bench
        bench.bind_to_gpu_numa_node(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 30
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)                                   # first touch on the bound NUMA node
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < a.seconds:
        for _ in range(4):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        reps += 4
    dt = time.perf_counter() - t0
    gbs = torch.tensor([reps * n / dt / 1e9], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(allv, gbs)
        vals = [float(v.item()) for v in allv]
    else:
        vals = [float(gbs.item())]
    if rank == 0:
        print(json.dumps({"ranks": world, "per_rank_GBps": [round(v, 2) for v in vals],
                          "aggregate_GBps": round(sum(vals), 2), "buffer_bytes": n, "seconds": a.seconds,
                          "cpus": os.cpu_count(), "numa_bind": bool(a.numa_bind)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
