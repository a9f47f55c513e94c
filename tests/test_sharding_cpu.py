"""world_size-2 gloo checks of the multi-GPU host logic (the data path itself has no
collective; ranks process disjoint image groups)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_view_stereonet_b200 import sharding, synthetic


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = sharding.shard_range(4, rank, world)
    left = synthetic.make_inputs(64, 80, 1, count, first_item=first)[0][0]
    # rank r pretends to take (r + 1) * 10 ms for 5 steps
    t = sharding.gather_timings([(rank + 1) * 10.0, float(left.double().sum())])
    value, worst = sharding.aggregate_throughput(count, 5, t[:, 0])
    q.put((rank, first, count, t.tolist(), value, worst))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = synthetic.make_inputs(64, 80, 1, 4)[0][0]
    for rank, first, count, t, value, worst in results:
        assert (first, count) == (rank * 2, 2)
        assert len(t) == world and t[0][0] == 10.0 and t[1][0] == 20.0      # every rank sees every timing
        # rank-local generation reproduces the single-process slice exactly
        assert abs(t[rank][1] - float(full[first:first + count].double().sum())) < 1e-9
        assert worst == 20.0
        assert abs(value - world * 2 * 5 / 0.020) < 1e-6                      # units of ALL ranks / slowest rank


def test_single_process_fallback():
    t = sharding.gather_timings([3.0, 4.0])
    assert tuple(t.shape) == (1, 2)
    value, worst = sharding.aggregate_throughput(1, 10, t[:, 0])
    assert worst == 3.0 and abs(value - 10 / 0.003) < 1e-6
