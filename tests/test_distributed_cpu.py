"""World-size-2 gloo test of the N>1 host logic: sharding is a partition, all-gather restores order."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mink_b200.distributed import all_gather_rows, shard, shard_bounds


def test_shard_bounds_partition():
    for total in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3)
    mine = shard(full) * 2.0            # "solve" the local shard
    gathered = all_gather_rows(mine, total)
    ok[rank] = int(torch.equal(gathered, full * 2.0))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_gloo_world2_gather(total):
    world, port = 2, _free_port()
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, port, total, ok), nprocs=world, join=True)
    assert list(ok) == [1, 1]
