"""Data parallelism: one process per GPU, NCCL over NVLink for the gradient exchange only.

Replaces the reference's DataParallelWithCallback (managers/base_manager.py:17-21: scatter /
replicate-every-forward / gather in one process) and its Sync-BN master/slave pipes
(networks/sync_batchnorm/comm.py).  The batch is sharded across ranks by the caller (each rank
loads / synthesises its own samples); the only data-path collective is one all-reduce(sum) of a
flat fp32 gradient bucket per optimizer step (G+E after the generator backward, D after the
discriminator backward), scaled by 1/world (the reference averages replica losses:
trainer_manager.py:36,53).  Parameters without a gradient (never-used `style_conv.*`, the encoder
branch the coin flip skipped) contribute zeros so every rank reduces the same layout.
Works unchanged with the `gloo` backend on CPU tensors (used by the world_size-2 tests).
"""
import os
import random

import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size():
    return dist.get_world_size() if is_dist() else 1


def rank():
    return dist.get_rank() if is_dist() else 0


def init_from_env(backend=None):
    """Initialises the default process group from torchrun's environment (no-op without it)."""
    if dist.is_available() and dist.is_initialized():
        return
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1:
        return
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {}
    if backend == "nccl":
        # bind the communicator to this rank's GPU up front (eager NCCL init, no device guessing in barrier())
        kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend, init_method="env://", **kw)


def allreduce_sum_(t):
    """In-place sum over ranks of a small statistics tensor (Sync-BN mode); identity for one rank."""
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def seed_python_random(seed):
    """Python's `random` drives SRModel's encoder coin flips (sr_model.py:616,643); all ranks must
    draw the same sequence or their gradient buckets would disagree on which encoder ran."""
    random.seed(seed)


def broadcast_module(module, src=0):
    if not is_dist():
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)


class GradBucket:
    """Flat fp32 gradient buffer over a fixed parameter list; one all-reduce per step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def nbytes(self):
        return self.flat.numel() * 4

    @torch.no_grad()
    def allreduce_mean(self):
        """Gathers .grad into the bucket (zeros where absent), all-reduces, and leaves every
        parameter's .grad as a view of the averaged bucket (no copy back).  The gather is one
        multi-tensor copy, not one launch per parameter."""
        if not is_dist():
            return
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / dist.get_world_size())
        for p, v in zip(self.params, self.views):
            p.grad = v
