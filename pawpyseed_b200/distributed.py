"""One-process-per-GPU sharding of the band-projection path: (k,spin) blocks over ranks (level 1) and, for jobs
with fewer blocks than GPUs, band blocks of one (k,spin) block over ranks (level 2).

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


def all_gather_own_blocks(own: np.ndarray, own_kappas, nkappa: int, group=None, device=None, pinned_out=None,
                          want_host=True):
    """All-gather of the per-kappa result blocks each rank owns (kappa % world == rank).

    `own` is [len(own_kappas), nS, nR] complex128 - only this rank's blocks, so the exchange moves
    nkappa * nS * nR * 16 B in total (NCCL all_gather over NVLink on GPUs, gloo on CPU).  Returns the full
    [nkappa, nS, nR] array on the host (want_host=False: the gathered device tensor, rank-major).  Ranks may own different block counts; shorter ones are padded."""
    import torch
    import torch.distributed as dist
    own = np.ascontiguousarray(own)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        full = np.zeros((nkappa,) + own.shape[1:], dtype=own.dtype)
        full[list(own_kappas)] = own
        return full
    world = dist.get_world_size(group)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    per = -(-nkappa // world)                       # blocks per rank, padded
    blk = int(np.prod(own.shape[1:]))
    send = torch.zeros(per * blk * 2, dtype=torch.float64, device=device)
    if len(own_kappas):
        send[: own.size * 2] = torch.from_numpy(own.reshape(-1).view(np.float64)).to(device, non_blocking=True)
    recv = torch.empty(world * per * blk * 2, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if not want_host:
        return recv          # every rank holds all blocks in HBM; only callers that ask pay the D2H
    # rank-major [world][per] -> kappa-major (kappa = j*world + r) on the device, then ONE copy of the nkappa valid
    # blocks to the host; with a pinned buffer the returned array is a view of it (no host-side repacking)
    if per > 1:
        recv = recv.view(world, per, blk * 2).transpose(0, 1).contiguous().view(-1)
    valid = recv[: nkappa * blk * 2]
    if pinned_out is not None:
        dst = pinned_out[: valid.numel()]
        dst.copy_(valid, non_blocking=False)
        host = dst.numpy()
    else:
        host = valid.cpu().numpy()
    return host.view(np.complex128).reshape((nkappa,) + own.shape[1:])


_gather_bufs = {}


def exchange_blocks(send, nkappa: int, blk: int, group=None, recv=None):
    """All-gather of the rank-local send buffers ([per * blk] float64, block j of rank r = kappa j*world + r,
    zero padded) into a kappa-major tensor [nkappa * blk] on the same device (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    per = -(-nkappa // world)
    if world == 1:
        return send[: nkappa * blk]
    if recv is None:
        recv = torch.empty(world * per * blk, dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if per > 1:     # rank-major [world][per] -> kappa-major (kappa = j*world + r)
        recv = recv.view(world, per, blk).transpose(0, 1).contiguous().view(-1)
    return recv[: nkappa * blk]


def gather_projection_blocks(pr, own_kappas, nkappa: int, group=None, pinned_out=None, want_host=True):
    """The (k,spin) blocks this rank owns, computed straight into a device send buffer
    (`pawb200_projection_matrix_dev`) and exchanged with one NCCL all-gather - the result never visits the host
    on the way (one allgather of the per-k projection matrices over NVLink).  `pr` is a CProjector /
    Projector whose wavefunctions were read with `set_read_shard(rank, world)`.

    Returns the full [nkappa, nband_wf, nband_basis] complex128 array on the host when want_host (a view of
    `pinned_out` if given), else the gathered, kappa-major device tensor (float64 pairs)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nS, nR = pr.wf.nband, pr.basis.nband
    blk = nS * nR * 2
    per = -(-nkappa // world)
    key = (per, blk, world)
    if key not in _gather_bufs:
        _gather_bufs.clear()
        _gather_bufs[key] = (torch.zeros(per * blk, dtype=torch.float64, device="cuda"),
                             torch.empty(world * per * blk, dtype=torch.float64, device="cuda"))
    send, recv = _gather_bufs[key]
    for j, k in enumerate(sorted(own_kappas)):
        pr._projection_matrix_dev(send[j * blk:(j + 1) * blk], kappa_range=(k, k + 1))
    valid = exchange_blocks(send, nkappa, blk, group, recv)
    if not want_host:
        return valid
    if pinned_out is not None:
        dst = pinned_out[: valid.numel()]
        dst.copy_(valid, non_blocking=False)
        host = dst.numpy()
    else:
        host = valid.cpu().numpy()
    return host.view(np.complex128).reshape(nkappa, nS, nR)


# --------------------------------------------------------------------------------------------
# level 2: band blocks of ONE (k,spin) block over ranks (jobs with fewer blocks than GPUs)
# --------------------------------------------------------------------------------------------
def set_band_shard(rank: int, world: int):
    """Wavefunctions read after this call hold only the bands [rank*per, (rank+1)*per), per = ceil(nband/world)."""
    from . import _lib
    _lib.lib().pawb200_set_band_shard(int(rank), int(world))


class _DeviceRows:
    """A row buffer inside the library (coefficients / projections of one (k,spin) block) viewed as a torch tensor."""

    def __init__(self, wf, which, kappa):
        import ctypes as C
        import torch
        from . import _lib
        ld, rows, lo, hi = C.c_long(0), C.c_int(0), C.c_int(0), C.c_int(0)
        ptr = _lib.lib().pawb200_get_device_buffer(wf.wf_ptr, int(which), int(kappa), C.byref(ld), C.byref(rows),
                                                   C.byref(lo), C.byref(hi))
        _lib.check()
        self.rows, self.lo, self.hi = rows.value, lo.value, hi.value
        elem = 8 if which == 0 else 16                     # complex64 / complex128
        self.row_bytes = ld.value * elem
        self.__cuda_array_interface__ = {"shape": (self.rows * self.row_bytes,), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 2}
        self.tensor = torch.as_tensor(self, device="cuda").view(self.rows, self.row_bytes)


def exchange_rows(buf, per: int, group=None):
    """In-place all-gather of equal row blocks: `buf` is [world * per, row_bytes] on every rank and rank r has
    filled rows [r*per, (r+1)*per).  NCCL gathers in place (send = the own slice of the receive buffer)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return buf
    rank = dist.get_rank(group)
    own = buf[rank * per:(rank + 1) * per]
    if dist.get_backend(group) != "nccl":
        own = own.clone()                                   # gloo: no aliasing between input and output
    dist.all_gather_into_tensor(buf.view(-1), own.reshape(-1), group=group)
    return buf


def gather_band_blocks(wf, group=None, coefficients=True, projections=True):
    """All-gather the band blocks of a band-sharded wavefunction in place over NVLink: afterwards every rank holds
    all rows of C (plane-wave coefficients), P (projections) and W (wave projections, when present) of every
    resident (k,spin) block.  This is the "allgather of the column blocks before the GEMM" of SURVEY 8e: the
    basis side of a projection needs it, the wf side does not."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    NK = wf.nwk * wf.nspin
    for kappa in range(NK):
        todo = ([0] if coefficients else []) + ([1, 2] if projections else [])
        for which in todo:
            try:
                rows = _DeviceRows(wf, which, kappa)
            except Exception:
                if which == 2:
                    continue                                # no wave projections for this pair
                raise
            exchange_rows(rows.tensor, rows.rows // world, group)


def band_sharded_projection_matrix(pr, group=None, pinned_out=None, want_host=True):
    """PAW-corrected overlap matrix of a band-sharded pair: the basis rows are gathered, every rank multiplies the
    rows of ITS wf bands against the whole basis (`pawb200_projection_matrix_dev` fills only those rows) and the
    row blocks are all-gathered.  Call after Projector / CProjector._setup_overlap on every rank.
    Returns [nkappa, nband_wf, nband_basis] complex128 (host array, or the device tensor when not want_host)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    gather_band_blocks(pr.basis, group)
    NK = pr.basis.nwk * pr.basis.nspin
    nS, nR = pr.wf.nband, pr.basis.nband
    per = -(-nS // world)
    full = torch.zeros(NK, world * per, nR * 2, dtype=torch.float64, device="cuda")
    blk = torch.empty(nS * nR * 2, dtype=torch.float64, device="cuda")
    for kappa in range(NK):
        pr._projection_matrix_dev(blk, kappa_range=(kappa, kappa + 1))
        lo, hi = min(nS, rank * per), min(nS, (rank + 1) * per)
        full[kappa, lo:hi] = blk.view(nS, nR * 2)[lo:hi]
        exchange_rows(full[kappa], per, group)
    out = full[:, :nS]
    if not want_host:
        return out
    if pinned_out is not None:
        dst = pinned_out[: out.numel()].view(NK, nS, nR * 2)
        dst.copy_(out, non_blocking=False)
        host = dst.numpy()
    else:
        host = out.cpu().numpy()
    return np.ascontiguousarray(host).view(np.complex128).reshape(NK, nS, nR)


def band_block(nband: int, rank: int, world: int):
    """Contiguous band range [lo, hi) of `rank` when the bands of one (k,spin) block are split over ranks."""
    per = -(-nband // world)
    return min(nband, rank * per), min(nband, (rank + 1) * per)


def sharded_chg_density(wf, group=None, device=None):
    """AE charge density with the bands split over ranks (the one reduction on the path).
    Every rank holds the same wavefunction `wf` (a CWavefunction with projectors set up), accumulates the
    |psi|^2 of its band block on its own GPU, and the f64 grids are summed with one all-reduce (NCCL on GPUs)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return wf._get_realspace_density()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = band_block(wf.nband, rank, world)
    part = wf._get_realspace_density_shard(lo, hi)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(part.reshape(-1)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().reshape(part.shape)


def max_over_ranks(value: float, group=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
