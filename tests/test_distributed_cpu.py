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


# ---- sharded sampling driver (host logic; the per-rank compute is replaced by a stand-in: no GPU here) ----------------------
def _fake_compute(batches, infos, units, lengths, generator):
    return {(f, e): torch.full((int(infos[f][0].numel()), 3), float(1000 * f + e)) for f, e in units}


def _sharded_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from codlad_b200 import sampler, synthetic
    prots = [synthetic.make_protein(n, 1, seed=200 + i) for i, n in enumerate([30, 55, 41, 38, 62])]
    batches = [synthetic.collate(p) for p in prots]
    infos = [p.info for p in prots]
    sb = sampler.ShardedBackmapper(None, rank, world)
    out = sb.backmap(batches, infos, 3, _compute=_fake_compute)
    parts = sb.plan([p.L for p in prots], 3)
    if rank == 0:
        q.put((sorted(out), [(k, tuple(v.shape), float(v[0, 0])) for k, v in sorted(out.items())], parts))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_backmapper_gathers_every_unit_on_rank0():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    keys, rows, parts = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from codlad_b200 import synthetic
    na = [synthetic.make_protein(n, 1, seed=200 + i).num_atoms for i, n in enumerate([30, 55, 41, 38, 62])]
    assert keys == [(f, e) for f in range(5) for e in range(3)]
    for (f, e), shape, v in rows:
        assert shape == (na[f], 3) and v == 1000 * f + e
    assert sorted(sum(parts, [])) == keys and all(parts)         # every unit on exactly one rank, both ranks busy
    frames = [sorted({f for f, _ in p}) for p in parts]
    assert not set(frames[0]) & set(frames[1])                   # F >= world: whole frames per rank (features built once)


def test_partition_and_grouping():
    from codlad_b200 import sampler
    # fewer frames than ranks: the members are dealt out, a frame may live on several ranks
    parts = sampler.partition_units([2000], 32, 8, k_neighbors=48)
    assert [len(p) for p in parts] == [4] * 8 and sorted(sum(parts, [])) == [(0, e) for e in range(32)]
    # ragged lengths: LPT keeps the per-rank cost within a few per cent
    import random
    rnd = random.Random(1)
    L = [rnd.randint(300, 700) for _ in range(256)]
    parts = sampler.partition_units(L, 1, 8)
    loads = [sum(L[f] for f, _ in p) for p in parts]
    assert max(loads) / (sum(loads) / 8) < 1.01
    groups = sampler.group_by_length(list(range(64)), L, max_frames=32, max_pad=0.12)
    assert sorted(sum(groups, [])) == list(range(64)) and all(len(g) <= 32 for g in groups)
    for g in groups:
        tot = sum(L[f] for f in g)
        assert (max(L[f] for f in g) * len(g) - tot) <= 0.12 * tot + 1e-9
    sb = sampler.ShardedBackmapper(None, 0, 1)
    lg = sb.local_groups([(0, 0), (0, 1), (2, 1), (2, 0), (1, 0)], [100, 50, 98])
    assert [g[0] for g in lg] == [[0, 2], [1]]
    assert lg[0][2] == [(0, 0), (2, 0), (0, 1), (2, 1)] and lg[0][1] == [0, 1, 0, 1]


# ---- gradient all-reduce of the data-parallel train_latent step ---------------------------------------
def _allreduce_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from codlad_b200.train import allreduce_flat
    n = 2449974 + 37                                           # the reference's parameter count, not a multiple of the bucket count
    g = torch.arange(n, dtype=torch.float32) * (rank + 1)
    allreduce_flat(g, n_buckets=4)
    q.put((rank, float(g[1]), float(g[-1]), bool(torch.equal(g, torch.arange(n, dtype=torch.float32) * 1.5))))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_averages_over_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_allreduce_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, b, ok in res:
        assert a == 1.5 and ok, (rank, a, b)
