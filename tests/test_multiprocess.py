"""CPU, world_size 2 (gloo): the N > 1 path of the hot path is clip sharding with no data-path
collective (lib/utils/subprocess.py:27-74, lib/core/test_engine.py:278-308): contiguous
np.array_split ranges per rank, results concatenated in rank order, and (bench.py) a barrier +
max-over-ranks of the per-rank device time."""
import os
import socket

import numpy as np
import pytest

from detectandtrack_b200.utils.subprocess import split_ranges


def test_split_ranges_matches_reference_array_split():
    for total, n in [(10, 2), (11, 4), (3, 8), (100, 8), (1, 1)]:
        parts = np.array_split(range(total), n)
        exp = [(int(p[0]), int(p[-1]) + 1) for p in parts if len(p)]
        got = split_ranges(total, n)
        assert got == exp
        assert sorted(i for s, e in got for i in range(s, e)) == list(range(total))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    s, e = split_ranges(total, world)[rank]
    # "detections" of this rank's clips: deterministic function of the clip index
    mine = [[i, i * i % 7] for i in range(s, e)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)           # per-rank elapsed ms
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((sum(gathered, []), t.item()))
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_over_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port, total = _free_port(), 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [[i, i * i % 7] for i in range(total)]        # rank-order concatenation == serial order
    assert tmax == 15.0


# ---- the one real exchange step of the path: the gradient all-reduce of the training config (model_builder.py:922-942) ----
def test_plan_buckets_cuts_at_parameter_boundaries():
    from detectandtrack_b200.modeling.trainer import plan_buckets
    counts = [100, 20, 300, 5, 75, 500]
    bounds, ends = plan_buckets(counts, 4)
    assert bounds == [100, 120, 420, 425, 500, 1000]
    assert ends[-1] == 1000 and all(e in bounds for e in ends) and ends == sorted(set(ends))
    assert plan_buckets(counts, 1)[1] == [1000]
    assert plan_buckets([7], 8) == ([7], [7])


def _train_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from detectandtrack_b200.modeling.trainer import plan_buckets, BucketReducer
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    counts = [64, 16, 200, 8, 40]                                  # gradients in backward order
    bounds, ends = plan_buckets(counts, 3)
    flat = torch.arange(sum(counts), dtype=torch.float32) * (rank + 1)          # rank r holds (r + 1) * g
    red = BucketReducer(flat, ends, world)
    for b in bounds:                                                # "backward": parameter gradients become ready in order
        red.ready(b)
    red.wait()
    if rank == 0:
        q.put(flat.clone())
    dist.destroy_process_group()


def test_two_rank_bucketed_gradient_allreduce():
    """world_size 2 on gloo: every bucket is reduced exactly once, launched as soon as its last gradient is ready; the
    result is the SUM over ranks (losses carry 1/NUM_GPUS in the reference, model_builder.py:484)."""
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(flat, torch.arange(328, dtype=torch.float32) * 3.0)
