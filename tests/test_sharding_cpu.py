"""World-size-2 gloo test of the multi-GPU host logic (contiguous block sharding + one all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from semdiff_b200 import sharding


def test_shard_ranges_cover_everything_once():
    for n in (0, 1, 7, 8, 10_000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(n))
            assert max(hi - lo for lo, hi in spans) <= sharding.shard_size(n, world)


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_pairs, dtype=torch.float32) * 0.5 + 1.0

    def load_pairs(lo, hi):
        return full[lo:hi].clone(), torch.zeros(hi - lo)

    scores = sharding.score_sharded(lambda gt, sr: gt * 2.0 + sr, n_pairs, load_pairs)
    q.put((rank, scores.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [7, 1])
def test_world_size_2_gloo(n_pairs):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [(i * 0.5 + 1.0) * 2.0 for i in range(n_pairs)]
    assert results[0] == expect and results[1] == expect
