"""Host-side data-parallel logic on CPU (gloo, world_size 2): bucketed gradient averaging and the
replicated ImagePool, which must return on every rank exactly what a single process seeing the global
batch would return (bit-exact)."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cycle_depth_estimation_b200.cycle_gan_model import CycleGANModel, GradBuckets
        from cycle_depth_estimation_b200.image_pool import ImagePool
        # ---- bucketed all-reduce == mean of the per-rank gradients
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(s)) for s in ((300, 7), (5,), (64, 3, 3, 3), (11,))]
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        params[1].grad = None  # frozen / unused parameters are skipped
        GradBuckets(params, bucket_bytes=4096).all_reduce()
        ok_grad = all(torch.allclose(p.grad, torch.full_like(p, (1 + 2) / 2 * (i + 1)))
                      for i, p in enumerate(params) if p.grad is not None) and params[1].grad is None
        # ---- replicated pool
        model = CycleGANModel.__new__(CycleGANModel)
        pool = ImagePool(3)
        random.seed(77)
        outs = []
        for step in range(12):
            fake = torch.full((2, 1, 2, 2), float(100 * step + 10 * rank)) + torch.arange(2.).view(2, 1, 1, 1)
            outs.append(model._pool_query(pool, fake).clone())
        # plain lists: a tensor would travel through the worker's fd-sharing socket, which is gone once it exits
        q.put((rank, ok_grad, torch.stack(outs).tolist(), list(pool.trace)))
    finally:
        dist.destroy_process_group()


def test_grad_buckets_and_replicated_pool_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    res = [(r[0], r[1], torch.tensor(r[2]), r[3]) for r in res]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "bucketed all-reduce did not produce the mean"
    assert res[0][3] == res[1][3], "ranks diverged in their pool decisions"
    # single-process reference over the global batch (rank-major order)
    from oracle.networks_oracle import ImagePoolOracle
    ref = ImagePoolOracle(3)
    random.seed(77)
    for step in range(12):
        fakes = [torch.full((2, 1, 2, 2), float(100 * step + 10 * r)) + torch.arange(2.).view(2, 1, 1, 1)
                 for r in range(world)]
        out = ref.query(torch.cat(fakes, 0))
        for r in range(world):
            assert torch.equal(res[r][2][step], out[2 * r:2 * r + 2]), (step, r)
    assert ref.trace == res[0][3]


def _metrics_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from cycle_depth_estimation_b200 import my_eval
        from oracle import networks_oracle as O
        gts, preds = _metric_pairs()
        calls = []

        def per_image(g, p):       # the device evaluation replaced by the oracle: the test is about the sharding
            calls.append(len(g))
            rows = O.eval_metric_arrays(list(g), list(p))[1]
            return np.concatenate([np.asarray(rows, np.float64), np.ones((len(g), 1))], 1)
        means, per = my_eval.eval_metric_arrays(gts, preds, per_image_fn=per_image)
        q.put((rank, [float(m) for m in means], per.tolist(), calls))
    finally:
        dist.destroy_process_group()


def _metric_pairs(n=7):
    import numpy as np
    rng = np.random.default_rng(2019)
    gts = rng.integers(0, 80, (n, 24, 40), dtype=np.uint8)
    preds = rng.integers(0, 256, (n, 24, 40), dtype=np.uint8)
    return gts, preds


def test_depth_metrics_shard_round_robin_world2():
    """images split round-robin, float32 rows all-gathered back into image order, the reference's
    ordered float32 reduction on every rank -> means bit-identical to the single-process evaluation (7 images over
    2 ranks: uneven shards)."""
    import numpy as np
    from oracle import networks_oracle as O
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_metrics_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    gts, preds = _metric_pairs()
    ref_means, ref_rows = O.eval_metric_arrays(list(gts), list(preds))
    assert res[0][3] == [4] and res[1][3] == [3]              # images 0,2,4,6 / 1,3,5
    for r in res:
        assert np.array_equal(np.asarray(r[2], np.float32), np.asarray(ref_rows, np.float32))
        assert [np.float32(m) for m in r[1]] == [np.float32(m) for m in ref_means]


def test_no_distributed_is_a_noop():
    from cycle_depth_estimation_b200.cycle_gan_model import GradBuckets
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    GradBuckets([p]).all_reduce()
    assert torch.equal(p.grad, torch.ones(4))
