"""World-size-2 gloo tests (CPU) of the frame sharder and result gathering."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from awesome_b200.sharding import gather_objects, merge_by_unit, shard_units


def test_shard_policies_cover_every_unit_once():
    for n in (0, 1, 7, 60, 61):
        for world in (1, 2, 4, 8):
            for policy in ("interleave", "chunk"):
                units = [shard_units(n, r, world, policy) for r in range(world)]
                flat = sorted(u for us in units for u in us)
                assert flat == list(range(n))
                assert max(len(u) for u in units) - min(len(u) for u in units) <= 1
    assert shard_units(60, 7, 8, "chunk") == list(range(53, 60))
    assert max(len(shard_units(60, r, 8)) for r in range(8)) == 8          # ideal 7.5x speed-up at 60 frames / 8 GPUs
    with pytest.raises(ValueError):
        shard_units(4, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_units(9, rank, world)
        local = {u: {"iou": 0.9 + 0.001 * u, "state": torch.full((3,), float(u))} for u in mine}
        merged = merge_by_unit(gather_objects(local))
        ok = list(merged.keys()) == list(range(9)) and all(float(merged[u]["state"][0]) == u for u in merged)
        t = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(float(t))
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1.0


def test_merge_rejects_double_ownership():
    with pytest.raises(ValueError):
        merge_by_unit([{0: 1}, {0: 2}])


def test_fixed_segmentation_is_independent_of_world_size():
    """fit_sequence_sharded's plan: the segments of the sequence are fixed; ranks only choose which of them they fit."""
    from awesome_b200.sharded_fit import plan_segments, segments_of_rank
    segs = plan_segments(60, 8)
    assert [len(s) for s in segs] == [8, 8, 8, 8, 7, 7, 7, 7] and [i for s in segs for i in s] == list(range(60))
    for world in (1, 2, 4, 8):
        owned = [segments_of_rank(len(segs), r, world) for r in range(world)]
        assert sorted(s for o in owned for s in o) == list(range(8))
        assert max(len(o) for o in owned) == 8 // world
    assert plan_segments(3, 8) == [[0], [1], [2]] and plan_segments(0, 4) == [[]]
    with pytest.raises(ValueError):
        segments_of_rank(8, 2, 2)
