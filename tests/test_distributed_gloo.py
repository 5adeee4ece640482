"""world_size-2 gloo test of the multi-GPU host logic (k-point sharding + block exchange), on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pawpyseed_b200 import distributed as pd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeWavefunction:
    """Stands in for a CWavefunction in the CPU test of the band-split density reduction: band b contributes
    (b + 1) * pattern to the grid, so the exact sum over all bands is known."""
    nband = 7
    pattern = np.arange(24, dtype=np.float64).reshape(2, 3, 4) + 1.0

    def _get_realspace_density_shard(self, lo, hi):
        return sum((b + 1) * self.pattern for b in range(lo, hi)) if hi > lo else np.zeros_like(self.pattern)

    def _get_realspace_density(self):
        return self._get_realspace_density_shard(0, self.nband)


def _worker(rank, world, port, nk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    full = rng.standard_normal((nk, 6, 5)) + 1j * rng.standard_normal((nk, 6, 5))
    own = pd.shard_kappas(nk, rank, world)
    local = np.zeros_like(full)
    local[own] = full[own]
    got = pd.gather_blocks(local)
    got2 = pd.all_gather_own_blocks(full[own], own, nk)
    # the device-resident exchange used by gather_projection_blocks (here on CPU tensors over gloo)
    blk, per = 6 * 5 * 2, -(-nk // world)
    send = torch.zeros(per * blk, dtype=torch.float64)
    for j, k in enumerate(own):
        send[j * blk:(j + 1) * blk] = torch.from_numpy(np.ascontiguousarray(full[k]).reshape(-1).view(np.float64))
    got3 = pd.exchange_blocks(send, nk, blk).numpy().view(np.complex128).reshape(full.shape)
    # level 2: in-place all-gather of band-row blocks (the C / P rows of a band-sharded wavefunction)
    per_rows, width = 3, 10
    rows = torch.zeros(world * per_rows, width, dtype=torch.uint8)
    ref_rows = torch.arange(world * per_rows * width, dtype=torch.int64).remainder(251).to(torch.uint8).view(-1, width)
    rows[rank * per_rows:(rank + 1) * per_rows] = ref_rows[rank * per_rows:(rank + 1) * per_rows]
    pd.exchange_rows(rows, per_rows)
    rows_ok = bool(torch.equal(rows, ref_rows))
    tmax = pd.max_over_ranks(float(rank + 1))
    fake = _FakeWavefunction()
    dens = pd.sharded_chg_density(fake)                       # bands split 4 + 3, one all-reduce
    dens_ok = bool(np.allclose(dens, fake._get_realspace_density(), rtol=0, atol=1e-12))
    q.put((rank, bool(np.array_equal(got, full) and np.array_equal(got2, full) and np.array_equal(got3, full)
                      and dens_ok and rows_ok), tmax, own))
    dist.destroy_process_group()


def test_shard_assignment_is_a_partition():
    for nk in (1, 3, 8):
        for world in (1, 2, 4, 8):
            seen = sorted(k for r in range(world) for k in pd.shard_kappas(nk, r, world))
            assert seen == list(range(nk))


def test_band_blocks_partition_the_bands():
    for nband in (1, 7, 600, 2000):
        for world in (1, 2, 3, 8):
            blocks = [pd.band_block(nband, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == nband
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


def test_block_exchange_world2_gloo():
    world, nk = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nk, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res)
    assert all(t == 2.0 for _, _, t, _ in res)       # max over ranks
    owns = {r: o for r, _, _, o in res}
    assert owns[0] == [0, 2, 4] and owns[1] == [1, 3]
