"""One-process-per-GPU sharding of the band-projection path over (k,spin) blocks.

The overlap operator is block diagonal in k-point and spin (pseudoprojector.c:71-89,
projector.c:872-887 loop over independent `kpt_num`), so rank r owns the blocks
kappa % world == r: it reads only those coefficient records into HBM, runs the whole path on
them with no data-path collective, and the per-kappa result matrices are exchanged once at the
end (NCCL over NVLink on GPUs; gloo in the CPU tests of this host logic).
"""
from __future__ import annotations

import numpy as np


def shard_kappas(nkappa: int, rank: int, world: int):
    """(k,spin) blocks owned by `rank` (round-robin, matches pawb200_set_read_shard)."""
    return [k for k in range(nkappa) if k % world == rank]


def set_read_shard(rank: int, world: int):
    from . import _lib
    _lib.lib().pawb200_set_read_shard(int(rank), int(world))


def gather_blocks(local: np.ndarray, group=None, device=None) -> np.ndarray:
    """All-gather of per-kappa result blocks.  `local` is [nkappa, nS, nR] complex128 holding this
    rank's blocks and zeros elsewhere (what pawb200_projection_matrix returns on a sharded
    wavefunction); ownership is disjoint, so a SUM all-reduce is exactly the all-gather."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    t = torch.from_numpy(np.ascontiguousarray(local).view(np.float64))
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().view(np.complex128).reshape(local.shape)


def max_over_ranks(value: float, group=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
