"""world_size-2 gloo run on CPU: the host-side multi-process logic (sharding + the single
all-reduce of episode statistics).  No GPU, no kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_for_im_b200 import dist as imx_dist


def test_shard_ranges_cover_everything():
    for total, world in ((65536, 8), (1000003, 8), (7, 4), (262144, 2)):
        ranges = [imx_dist.shard_range(total, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        for a, b in zip(ranges, ranges[1:]):
            assert a[1] == b[0]
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = imx_dist.shard_config({"num_stages": 4}, total)
    lo, hi = cfg["env_offset"], cfg["env_offset"] + cfg["num_envs"]
    rng = np.random.default_rng(5)
    returns = rng.normal(size=(total, 4))[lo:hi]               # what this rank's envs produced
    tot = returns.sum(axis=1)
    stats = [float(hi - lo), tot.sum(), (tot ** 2).sum()]
    for i in range(4):
        stats += [returns[:, i].sum(), (returns[:, i] ** 2).sum()]
    st = imx_dist.allreduce_stats(torch.tensor(stats, dtype=torch.float64))
    if rank == 0:
        q.put((imx_dist.summarize(st), lo, hi))
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    total = 1001
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    summary, lo, hi = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    full = np.random.default_rng(5).normal(size=(total, 4))
    tot = full.sum(axis=1)
    assert summary["n"] == total and (lo, hi) == (0, 501)
    np.testing.assert_allclose(summary["mean"], tot.mean(), rtol=1e-12)
    np.testing.assert_allclose(summary["std"], tot.std(), rtol=1e-10)
    for i in range(4):
        np.testing.assert_allclose(summary["per_agent"][i]["mean"], full[:, i].mean(), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(summary["per_agent"][i]["std"], full[:, i].std(), rtol=1e-10)
