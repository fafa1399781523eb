"""Data parallelism: one process per GPU, NCCL over NVLink for the gradient exchange only.

Replaces the reference's DataParallelWithCallback (managers/base_manager.py:17-21: scatter /
replicate-every-forward / gather in one process) and its Sync-BN master/slave pipes
(networks/sync_batchnorm/comm.py).  The batch is sharded across ranks by the caller (each rank
loads / synthesises its own samples); the only data-path collective is one all-reduce(sum) of a
flat fp32 gradient bucket per optimizer step (G+E after the generator backward, D after the
discriminator backward), scaled by 1/world (the reference averages replica losses:
trainer_manager.py:36,53).  Parameters without a gradient (never-used `style_conv.*`, the encoder
branch the coin flip skipped) contribute zeros so every rank reduces the same layout, and get
`.grad = None` back afterwards so the optimizer skips them like a single-process run does.
Works unchanged with the `gloo` backend on CPU tensors (used by the world_size-2 tests).
"""
import os
import random

import torch
import torch.distributed as dist


_LOCAL = [False]


class local_mode:
    """`with local_mode():` this process behaves like a single-process run (no gradient buckets, no
    Sync-BN exchange, no broadcasts) although a process group exists - the single-process reference
    of the data-parallel parity check that bench.py runs at N > 1."""

    def __enter__(self):
        self.saved = _LOCAL[0]
        _LOCAL[0] = True

    def __exit__(self, *exc):
        _LOCAL[0] = self.saved
        return False


def is_dist():
    return (not _LOCAL[0]) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


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


class _PeerExchange:
    """Sum of small fp32 vectors across the GPUs of this node through NVLink peer memory: one kernel
    of the library per exchange (csrc/peer_kernels.cu) instead of an NCCL launch.  The symmetric
    buffer comes from torch.distributed._symmetric_memory (CUDA virtual-memory handles exchanged
    through the process group's store)."""
    RING, MAXN = 4, 4096

    def __init__(self):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.lib = _lib.load()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        nbytes = int(self.lib.dsee_peer_exchange_bytes(self.world, self.RING, self.MAXN))
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert len(ptrs) == self.world and all(ptrs)
        self.ptrs = (C.c_void_p * self.world)(*ptrs)
        torch.cuda.synchronize()
        dist.barrier()     # every rank's buffer is zeroed before anyone pushes

    def allreduce_sum_(self, t):
        import ctypes as C
        from . import _lib, ops
        stream = ops._stream()
        _lib.check(self.lib.dsee_peer_allreduce_small(self.ptrs, self.world, self.rank, self.RING, self.MAXN,
                                                      C.c_void_p(t.data_ptr()), C.c_void_p(t.data_ptr()),
                                                      t.numel(), stream))
        return t


_PEER = {"obj": None, "failed": False}


def _peer_exchange():
    """The node-local peer-memory exchange, or None (disabled, not NCCL, more than 8 ranks, or
    symmetric memory unavailable: then the statistics go through NCCL like the gradients)."""
    from .config import config
    if not config.peer_sync_bn or _PEER["failed"]:
        return None
    if _PEER["obj"] is None:
        try:
            if dist.get_backend() != "nccl" or dist.get_world_size() > 8:
                raise RuntimeError("needs the nccl backend and at most 8 ranks on one node")
            _PEER["obj"] = _PeerExchange()
        except Exception as e:  # noqa: BLE001
            import sys
            print("deepsee_b200: peer-memory statistics exchange unavailable (%r); using NCCL" % (e,),
                  file=sys.stderr)
            _PEER["failed"] = True
            return None
    return _PEER["obj"]


def peer_exchange_active():
    """True once the NVLink peer-memory exchange has been set up in this process."""
    return _PEER["obj"] is not None


def allreduce_sum_(t):
    """In-place sum over ranks of a small statistics tensor (Sync-BN mode); identity for one rank.
    fp32 CUDA vectors of up to 4096 elements go through the NVLink peer-memory kernel, everything
    else (and the gloo tests) through the process group."""
    if is_dist():
        px = None
        if t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= _PeerExchange.MAXN:
            px = _peer_exchange()
        if px is not None:
            px.allreduce_sum_(t)
        else:
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
    """Flat fp32 gradient buffer over a fixed parameter list, all-reduced in a few contiguous chunks
    that are launched from inside the backward pass.

    * `begin()` replaces `optimizer.zero_grad()`: one memset of the flat buffer and every
      parameter's `.grad` becomes its view of it, so autograd accumulates straight into the bucket
      (no gather copy afterwards).
    * A post-accumulate-grad hook per parameter counts down its chunk; when the last gradient of a
      chunk has landed, that chunk's all-reduce is issued asynchronously (NCCL's own stream) while
      the backward pass continues.  Chunks are contiguous in parameter order, so the chunk holding
      the LAST layers completes first.  The launch order is a function of the autograd graph, hence
      the same on every rank.
    * `finish()` issues whatever chunks are still open (parameters that got no gradient this step:
      never-used `style_conv.*`, the encoder branch the coin flip skipped), waits, scales by
      1/world, and restores `.grad = None` on the parameters that received no gradient, so Adam
      skips them exactly like the single-process run and the reference do (the coin flips are
      rank-synchronised, so the local "no gradient" set is the global one).
    """

    def __init__(self, params, n_chunks=4):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        offs = []
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            offs.append(off)
            off += p.numel()
        # contiguous chunks of roughly equal size
        n_chunks = max(1, min(n_chunks, len(self.params)))
        target = (n + n_chunks - 1) // n_chunks if n else 1
        self.chunk_of, self.chunk_range, self.chunk_params = [], [], []
        lo = 0
        cnt = 0
        for i, p in enumerate(self.params):
            self.chunk_of.append(len(self.chunk_range))
            cnt += 1
            end = offs[i] + p.numel()
            if end - lo >= target or i == len(self.params) - 1:
                self.chunk_range.append((lo, end))
                self.chunk_params.append(cnt)
                lo, cnt = end, 0
        self._active = False
        self._pending, self._touched, self._works, self._launched = [], [], [], []
        from .config import config
        self._overlap = config.grad_overlap
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def nbytes(self):
        return self.flat.numel() * 4

    def _make_hook(self, i):
        def hook(param):
            if not self._active:
                return
            if not self._touched[i]:
                self._touched[i] = True
                c = self.chunk_of[i]
                self._pending[c] -= 1
                if self._pending[c] == 0 and self._overlap:
                    self._launch(c)
        return hook

    def _launch(self, c):
        lo, hi = self.chunk_range[c]
        self._launched[c] = True
        if hi > lo:
            self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    @torch.no_grad()
    def begin(self):
        """Call instead of optimizer.zero_grad() before the forward pass of a step."""
        if not is_dist():
            for p in self.params:
                p.grad = None
            return
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
        self._pending = list(self.chunk_params)
        self._touched = [False] * len(self.params)
        self._launched = [False] * len(self.chunk_range)
        self._works = []
        self._active = True

    @torch.no_grad()
    def finish(self):
        """Call after backward(): completes the reduction; .grad = averaged gradient (a bucket view)
        for parameters that received one on this step, None for the others."""
        if not is_dist():
            return
        if not self._active:
            # begin() was not called (gradients live in their own tensors): gather them first
            self._gather_loose_grads()
        else:
            for i, (p, v) in enumerate(zip(self.params, self.views)):
                g = p.grad
                if g is not None and g.data_ptr() != v.data_ptr():
                    # something replaced .grad (e.g. a hook-free manual assignment): fold it in
                    v.copy_(g)
                    p.grad = v
                    self._touched[i] = True
        self._active = False
        for c in range(len(self.chunk_range)):
            if not self._launched[c]:
                self._launch(c)
        for w in self._works:
            w.wait()
        self._works = []
        self.flat.mul_(1.0 / dist.get_world_size())
        for p, v, t in zip(self.params, self.views, self._touched):
            p.grad = v if t else None

    def _gather_loose_grads(self):
        self._touched = [p.grad is not None for p in self.params]
        self._launched = [False] * len(self.chunk_range)
        self._works = []
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)

    def allreduce_mean(self):
        """One-shot form (no begin()): gathers .grad into the bucket, all-reduces, averages."""
        self.finish()
