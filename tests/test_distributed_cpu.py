"""CPU (gloo, world_size 2): the N > 1 host logic -- sharding with no data-path collective, max-over-ranks timing,
whole-job unit accounting.  The per-rank compute itself is covered by the GPU tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from codlad_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_units, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_bounds(n_units, rank, world)
    elapsed = 10.0 + rank                                   # rank 1 is the slow one
    t = D.max_over_ranks(elapsed)
    counts = D.gather_counts(hi - lo)
    q.put((rank, lo, hi, t, counts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, n_units = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, lo, hi, t, counts in res:
        covered += list(range(lo, hi))
        assert t == 11.0                                    # max over ranks, identical on every rank
        assert counts == [6, 5] and sum(counts) == n_units
    assert covered == list(range(n_units))                  # every unit exactly once, no overlap


def test_shard_helpers():
    assert [D.shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    parts = D.shard_by_cost([500, 100, 400, 300, 200, 600], 2)
    assert sorted(sum(parts, [])) == list(range(6))
    loads = [sum([500, 100, 400, 300, 200, 600][i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 100
    assert D.env_rank_world()[1] >= 1
