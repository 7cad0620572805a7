"""N > 1 host logic on CPU: two gloo ranks shard an ensemble, "integrate" their slices with the
C oracle standing in for the GPU kernels (the product kernels need a device), and agree on
max-timing / summed work and on the gathered result (no collective in the
data path, only a final gather)."""

import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from nbkode_b200 import distributed as nd  # noqa: E402


def test_shard_bounds_partition():
    for n in (1, 7, 64, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [nd.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    a = np.arange(10).reshape(5, 2)
    assert np.array_equal(np.concatenate([nd.shard(a, r, 2) for r in range(2)]), a)


def _worker(rank, world, port, ret):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "numbakit-ode_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import nbk_oracle_c as oc
        from oracle import nbk_oracle_np as onp

        B = 37  # odd on purpose: shards of 19 and 18 rows
        y0, p = onp.lorenz_ensemble(B, seed=4)
        lo, hi = nd.shard_bounds(B, rank, world)
        mine = oc.run_ensemble("RungeKutta45", "lorenz", nd.shard(y0, rank, world), nd.shard(p, rank, world), [1.0], rtol=1e-6, atol=1e-9)
        y_all = nd.gather_batch(torch.from_numpy(mine["ys"]), B)
        (t_max,), (steps,) = nd.reduce_max_sum([10.0 + rank], [float(mine["n_accepted"].sum())])
        full = oc.run_ensemble("RungeKutta45", "lorenz", y0, p, [1.0], rtol=1e-6, atol=1e-9)
        ok = (
            np.array_equal(y_all.numpy(), full["ys"]) and t_max == 10.0 + world - 1
            and steps == float(full["n_accepted"].sum()) and (hi - lo) == mine["ys"].shape[0]
        )
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_bind_to_gpu_numa_is_best_effort():
    """Without NVML / a GPU the helper changes nothing and says so (bench.py calls it unconditionally for N > 1)."""
    import os

    before = os.sched_getaffinity(0)
    assert nd.bind_to_gpu_numa(0) in (True, False)
    assert os.sched_getaffinity(0) <= before and len(os.sched_getaffinity(0)) > 0
    os.sched_setaffinity(0, before)
