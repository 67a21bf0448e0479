"""CPU test of the N > 1 path: two gloo ranks each own a contiguous shard of a batch of perturbed Pyramid
worlds (stepped by the test-only host simulator), all-gather the per-world digests, and rank 0 checks
them against one process that owns the whole batch — sharding by world changes nothing."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import HOSTSIM_SO, ROOT

TOTAL, STEPS, SEED = 6, 25, 77


def _simulate(first, count):
    from box2d_rs_b200 import batch, scenes, sharding, world
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    scenes.pyramid(wg)
    wg.set_allow_sleeping(False)
    b = wg.batch(count, max_contacts=800)
    b.set_linear_velocity(211, sharding.perturbation(first, count, SEED))
    b.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, STEPS)
    d = sharding.world_digests(b.body_state())
    b.close()
    wg.close()
    ctx.close()
    return d


def _rank_main(rank, world_size, port, queue):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from box2d_rs_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    first, end = sharding.world_range(TOTAL, rank, world_size)
    digests = _simulate(first, end - first)
    gathered = sharding.allgather_digests(dist, digests)
    dist.barrier()
    if rank == 0:
        queue.put(np.concatenate(gathered))
    dist.destroy_process_group()


def test_world_range_partitions():
    from box2d_rs_b200 import sharding
    for total in (1, 7, 4096, 4099):
        for ws in (1, 2, 3, 8):
            spans = [sharding.world_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [e - f for f, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_shards_match_single_process(built):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = _simulate(0, TOTAL)
    assert sharded.shape == (TOTAL,)
    assert np.array_equal(sharded, whole)
    assert len(set(whole.tolist())) == TOTAL  # the perturbation really makes the worlds differ
