"""World-size-2 gloo test of the only multi-rank step of the path: sharding the scenario set and
reducing / gathering the closed-loop statistics."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sihmpc_b200.dist import shard_range, reduce_moments, gather_stats, moments_from_stats
from sihmpc_b200 import scenarios as sc


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    scen = sc.distillation_scenarios(hi - lo, first=lo)
    # a deterministic stand-in for the device statistics block: a function of the scenario draws
    stats = np.concatenate([scen["dist"], scen["ysp_seq"][:, 2, :], scen["dist_step"][:, None].astype(float)], axis=1)
    agg = torch.tensor(moments_from_stats(stats))
    reduce_moments(agg)
    full = gather_stats(torch.tensor(stats), total)
    if rank == 0:
        q.put((agg.numpy(), full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly():
    for total in (0, 1, 7, 4096, 1000003):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_stats_exchange_matches_single_process():
    total, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    agg, full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scen = sc.distillation_scenarios(total)
    ref = np.concatenate([scen["dist"], scen["ysp_seq"][:, 2, :], scen["dist_step"][:, None].astype(float)], axis=1)
    assert np.array_equal(full, ref)                         # scenario draws do not depend on the sharding
    assert np.allclose(agg, moments_from_stats(ref), rtol=1e-13, atol=0)
