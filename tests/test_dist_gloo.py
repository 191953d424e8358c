"""CPU, world_size 2 over gloo: the multi-GPU path's host logic -- trajectory sharding and the
rank-ordered gather + fixed-order merge of the dataset-statistics aggregates (the only collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pipeline as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_traj, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fluid_llm_b200.compute_ds_stats import gather_stats
    from fluid_llm_b200.field_path import shard_range
    lo, hi = shard_range(n_traj, rank, world)
    rng = np.random.default_rng(0)
    data = [rng.standard_normal(500 + 37 * i) * (1 + i) + i for i in range(n_traj)]     # same on every rank
    agg = torch.zeros(6, 3, dtype=torch.float64)
    for c in range(6):
        a = (0, 0.0, 0.0)
        for i in range(lo, hi):
            a = P.update_variance_batch(a, data[i] + c)
        agg[c] = torch.tensor(a, dtype=torch.float64)
    parts = gather_stats(agg)                      # (world, 6, 3), rank order, identical on all ranks
    merged = []
    for c in range(6):
        m = (0, 0.0, 0.0)
        for r in range(world):
            m = P.chan_merge(m, tuple(parts[r, c].tolist()))
        merged.append(m)
    out[rank] = (parts.numpy().copy(), np.array(merged), (lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_gather_and_merge_world2():
    world, n_traj = 2, 7
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_traj, out), nprocs=world, join=True)
    p0, m0, s0 = out[0]
    p1, m1, s1 = out[1]
    assert s0 == (0, 4) and s1 == (4, 7)
    assert np.array_equal(p0, p1) and np.array_equal(m0, m1)          # bit-identical on every rank
    rng = np.random.default_rng(0)
    data = np.concatenate([rng.standard_normal(500 + 37 * i) * (1 + i) + i for i in range(n_traj)])
    for c in range(6):
        assert m0[c, 0] == len(data)
        np.testing.assert_allclose(m0[c, 1], (data + c).mean(), rtol=1e-12)
        np.testing.assert_allclose(np.sqrt(m0[c, 2] / m0[c, 0]), (data + c).std(), rtol=1e-12)
