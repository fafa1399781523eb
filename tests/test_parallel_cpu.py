"""CPU, world_size 2 over gloo: the data-parallel plumbing (flat gradient bucket all-reduce, module
broadcast, identical coin-flip seeds).  One process per rank like the GPU launch."""
import os
import random
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from deepsee_b200 import parallel
    parallel.init_from_env(backend="gloo")
    assert parallel.is_dist() and parallel.world_size() == world and parallel.rank() == rank
    torch.manual_seed(100 + rank)                     # ranks start with DIFFERENT weights
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    net.register_buffer("running", torch.full((4,), float(rank)))
    parallel.broadcast_module(net)                     # ... and end up with rank 0's
    w0 = net[0].weight.detach().clone()
    # one-shot form: rank r contributes (r + 1) on layer 0 and 4 * (1 - r) on layer 1
    for p in net[0].parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    for p in net[1].parameters():
        p.grad = torch.full_like(p, 4.0 * (1 - rank))
    bucket = parallel.GradBucket(list(net.parameters()), n_chunks=2)
    bucket.allreduce_mean()
    # after the reduction every .grad IS a view of the flat bucket (no copy back)
    views_ok = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    g0_oneshot, g1_oneshot = net[0].weight.grad.clone(), net[1].weight.grad.clone()
    # hook form (what TrainerManager does): begin() -> backward accumulates straight into the bucket
    # views and chunks are all-reduced as they complete -> finish().  `extra` takes no part in the
    # loss (like the encoder branch the coin flip skipped): its .grad must come back None so Adam
    # skips it as in a single-process run.
    extra = torch.nn.Parameter(torch.zeros(7))
    bucket2 = parallel.GradBucket(list(net.parameters()) + [extra], n_chunks=3)
    bucket2.begin()
    xin = torch.full((2, 6), float(rank + 1))
    net(xin).sum().backward()
    bucket2.finish()
    with torch.no_grad():
        # d/dW1 of sum(W1 (W0 x + b0) + b1) = 1 (x) h summed over the batch; compare with the mean
        # of both ranks' analytic gradients
        want = []
        for r in range(world):
            xr = torch.full((2, 6), float(r + 1))
            h = torch.nn.functional.linear(xr, net[0].weight, net[0].bias)
            want.append(torch.ones(2, 3).t() @ h)
        want = sum(want) / world
    views_ok = views_ok and extra.grad is None and torch.allclose(net[1].weight.grad, want, atol=1e-5) \
        and net[1].weight.grad.data_ptr() == bucket2.views[2].data_ptr() \
        and all(lo < hi for lo, hi in bucket2.chunk_range) and len(bucket2.chunk_range) == 3
    # Sync-BN mode: batch-norm backward's two reductions are summed over ranks, the parameter-gradient
    # rows stay local (deepsee_models/networks/architecture.py:_sync_bwd_sums)
    from deepsee_b200.config import config
    from deepsee_b200.deepsee_models.networks import architecture as arch
    nsums = torch.arange(12, dtype=torch.float32).view(4, 3) + 100 * rank

    class _St:
        inv_count = 0.25
    config.sync_bn = False
    local = arch._sync_bwd_sums(nsums, _St)
    config.sync_bn = True
    synced = arch._sync_bwd_sums(nsums, _St)
    _St.inv_count = 0.0                                # eval-mode statistics: nothing to synchronise
    evalmode = arch._sync_bwd_sums(nsums, _St)
    config.sync_bn = False
    sync_ok = (local is nsums and evalmode is nsums and
               torch.equal(synced[:2], 2 * torch.arange(6, dtype=torch.float32).view(2, 3) + 100) and
               torch.equal(synced[2:], nsums[2:]))
    stat = parallel.allreduce_sum_(torch.full((2, 4), float(rank + 1)))
    sync_ok = sync_ok and torch.equal(stat, torch.full((2, 4), 3.0))
    parallel.seed_python_random(0)
    flips = [random.random() for _ in range(3)]
    q.put((rank, w0, g0_oneshot, g1_oneshot, net.running.clone(), flips,
           bucket.nbytes(), views_ok, sync_ok))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w_a, g0_a, g1_a, run_a, flips_a, nb, va, sa), (_, w_b, g0_b, g1_b, run_b, flips_b, _, vb, sb) = res
    assert va and vb, "bucket views / hook-driven chunked all-reduce / None-gradient restoration"
    assert sa and sb, "Sync-BN statistics all-reduce"
    assert torch.equal(w_a, w_b)                        # broadcast from rank 0
    assert torch.equal(run_a, run_b) and float(run_a[0]) == 0.0   # buffers too
    assert torch.allclose(g0_a, torch.full_like(g0_a, 1.5)) and torch.equal(g0_a, g0_b)   # mean(1, 2)
    assert torch.allclose(g1_a, torch.full_like(g1_a, 2.0)) and torch.equal(g1_a, g1_b)   # mean(4, 0)
    assert flips_a == flips_b
    assert nb == 4 * (6 * 5 + 5 + 5 * 3 + 3)


def test_single_process_is_a_noop():
    from deepsee_b200 import parallel
    assert not parallel.is_dist() and parallel.world_size() == 1 and parallel.rank() == 0
    net = torch.nn.Linear(3, 2)
    net.weight.grad = torch.ones_like(net.weight)
    parallel.GradBucket(list(net.parameters())).allreduce_mean()   # no process group: untouched
    assert torch.equal(net.weight.grad, torch.ones_like(net.weight))
    parallel.broadcast_module(net)


def test_dropin_aliases_resolve_to_this_package():
    import sys
    import deepsee_b200.dropin as dropin
    saved = {k: sys.modules.get(k) for k in dropin._ALIASES}
    try:
        for k in dropin._ALIASES:
            sys.modules.pop(k, None)
        dropin.install()
        from managers.trainer_manager import TrainerManager
        from deepsee_models.sr_model import SRModel
        import deepsee_models.networks as networks
        assert TrainerManager.__module__.startswith("deepsee_b200.")
        assert SRModel.__module__.startswith("deepsee_b200.")
        assert hasattr(networks, "define_SR") and hasattr(networks, "define_D") and hasattr(networks, "define_E")
        for m in ("run_generator_one_step", "run_discriminator_one_step", "get_latest_losses",
                  "get_latest_generated", "save", "update_learning_rate", "get_logs"):
            assert hasattr(TrainerManager, m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
