"""TrainerManager (reference: managers/trainer_manager.py:6-99).

Same public surface as the reference (`run_generator_one_step`, `run_discriminator_one_step`,
`get_latest_losses`, `get_latest_generated`, `get_logs`, `save`, `update_learning_rate`,
`sr_model`, `sr_model_on_one_gpu`), so train.py drives it unchanged.  Differences:
  * one process per GPU; the G(+E) or D gradients are averaged across ranks through a flat
    bucket whose chunks are all-reduced over NCCL while the backward pass is still running
    (..parallel.GradBucket);
  * checkpoints are written by rank 0 only.
Both optimizer steps share one code path (`_optimize`).
"""
import random

import torch
from torch.nn.utils import clip_grad_value_

from .. import ops, parallel
from ..config import config
from .base_manager import BaseManager


def _detached(x):
    """The step's result (dict / tuple of tensors) without autograd history: the static outputs of a
    captured graph must not keep the capture-time autograd graph alive."""
    if torch.is_tensor(x):
        return x.detach()
    if isinstance(x, dict):
        return type(x)((k, _detached(v)) for k, v in x.items())
    if isinstance(x, (tuple, list)):
        return type(x)(_detached(v) for v in x)
    return x


class _ReplayRandom:
    """Stands in for `random.random` while one optimizer sub-step runs: hands out values drawn up
    front from the real generator (in order), so the global stream advances exactly as in an
    un-graphed run while the outcome of the encoder coin flips (sr_model.py:616,643) is known before
    the step starts - it selects which captured graph to replay."""

    def __init__(self, values, real):
        self.values, self.real, self.i = list(values), real, 0

    def __call__(self):
        if self.i < len(self.values):
            v = self.values[self.i]
        else:           # more draws than the probe step saw: fall back to the real stream
            v = self.real()
            self.values.append(v)
        self.i += 1
        return v


class _GraphedStep:
    """One optimizer sub-step of TrainerManager ('generator' or 'discriminator') as CUDA graphs.

    The first calls run eagerly (they create the Adam state, size NCCL buffers, set kernel
    attributes and count how many `random.random()` draws the step makes).  After that, each
    distinct outcome of those draws - full / mini style encoder, noisy / clean style - is captured
    once: preprocessing, forward, backward, gradient all-reduce, clipping and the Adam update of one
    static input batch, all kernels on the capture stream.  Later calls copy the batch into the
    static buffers and replay.  NoiseInjection seeds are baked into the captured launches; the
    device-side noise epoch, advanced by the graph's first node, makes every replay draw fresh noise
    (csrc: eff_noise_seed).  A capture that fails for any reason disables graphs for this manager and
    the step continues eagerly."""

    EAGER_CALLS = 2

    def __init__(self, mgr, mode, optimizer, bucket):
        self.mgr, self.mode, self.optimizer, self.bucket = mgr, mode, optimizer, bucket
        self.calls = 0
        self.n_draws = None
        self.static_in = None
        self.graphs = {}          # flips -> (graph, outputs)
        self.pool = None
        self.lr_sig = None
        self.disabled = False
        self.captured_launches = {}

    def _materialize_optimizer_state(self):
        """Creates the Adam state of every parameter that has none yet, exactly as torch.optim.Adam would
        on its first gradient (zeros, step 0), BEFORE anything is captured.  A parameter that received no
        gradient in the eager calls (the style-noise weights when both coin flips came out 'clean', the
        branch of the encoder the flips skipped) would otherwise get its state allocated from the graph's
        private pool in the middle of a capture: that block is scratch memory of the graphs captured
        earlier into the same pool, whose replays then overwrite the moments (NaN parameters some
        iterations later, depending on the order of the coin flips)."""
        opt = self.optimizer
        with torch.cuda.stream(self.mgr._train_stream):
            for group in opt.param_groups:
                for p in group['params']:
                    if not p.requires_grad or len(opt.state.get(p, {})) != 0:
                        continue
                    st = opt.state[p]
                    on_device = group.get('capturable') or group.get('fused')
                    st['step'] = (torch.zeros((), dtype=torch.get_default_dtype(), device=p.device) if on_device
                                  else torch.tensor(0.0, dtype=torch.get_default_dtype()))
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group.get('amsgrad'):
                        st['max_exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        self.mgr._train_stream.synchronize()

    def _lr_signature(self):
        return tuple(float(g['lr']) for g in self.optimizer.param_groups)

    def _stage(self, data):
        """Copies the caller's batch into the static device buffers (allocated on first use)."""
        if self.static_in is None:
            self.static_in = {k: torch.empty(v.shape, dtype=v.dtype, device='cuda')
                              for k, v in data.items() if torch.is_tensor(v)}
        for k, buf in self.static_in.items():
            buf.copy_(data[k], non_blocking=True)
        return dict(self.static_in)

    def _eager(self, data):
        """Eager step on the manager's training stream.  Everything that touches the parameters runs
        on that one non-default stream, eager steps and captures alike: autograd binds a parameter's
        gradient accumulator to the stream of its first use, and an accumulator living on the legacy
        default stream would invalidate a later capture."""
        mgr = self.mgr
        cur = torch.cuda.current_stream()
        mgr._train_stream.wait_stream(cur)
        with torch.cuda.stream(mgr._train_stream):
            out = mgr._step_body(data, self.mode, self.optimizer, self.bucket)
        cur.wait_stream(mgr._train_stream)
        return out

    def __call__(self, data):
        real = random.random
        self.calls += 1
        if self.disabled or self.calls <= self.EAGER_CALLS:
            # eager; the first call counts the draws
            counter = {'n': 0}

            def counting():
                counter['n'] += 1
                return real()
            random.random = counting
            try:
                out = self._eager(data)
            finally:
                random.random = real
            self.n_draws = counter['n']
            return out
        if self.static_in is not None and any(
                torch.is_tensor(v) and (k not in self.static_in or v.shape != self.static_in[k].shape)
                for k, v in data.items()):
            # a batch of another shape (the last, smaller one of an epoch): this call runs eagerly
            draws = [real() for _ in range(self.n_draws)]
            random.random = _ReplayRandom(draws, real)
            try:
                return self._eager(data)
            finally:
                random.random = real
        draws = [real() for _ in range(self.n_draws)]
        return self._run_variant(draws, data)

    def capture_all(self, data):
        """Captures every coin-flip variant now (bench warm-up: no capture inside a timed region).
        Each capture executes one real optimizer step on `data`; the global `random` stream is not
        touched."""
        import itertools
        if self.disabled or self.n_draws is None:
            return
        for flips in itertools.product((0.25, 0.75), repeat=self.n_draws):
            self._run_variant(list(flips), data)

    def _run_variant(self, draws, data):
        mgr = self.mgr
        real = random.random
        if self.lr_sig != self._lr_signature():    # learning rates are baked into the Adam nodes
            # (destroying the last graph of a private pool releases the pool: start a new one)
            self.graphs.clear()
            self.captured_launches.clear()
            self.pool = None
            self.lr_sig = self._lr_signature()
        key = tuple(v < 0.5 for v in draws)
        static = self._stage(data)
        if key not in self.graphs:
            if not self.graphs:
                self._materialize_optimizer_state()
            replay = _ReplayRandom(draws, real)
            random.random = replay
            saved_check, saved_timer = config.check_onehot, ops.KernelTimer.active
            config.check_onehot, ops.KernelTimer.active = False, None
            from .. import _lib
            n0 = _lib.launch_count()
            try:
                g = torch.cuda.CUDAGraph()
                if self.pool is None:
                    self.pool = torch.cuda.graph_pool_handle()
                kw = {"capture_error_mode": "thread_local"} if parallel.is_dist() else {}
                with torch.cuda.graph(g, pool=self.pool, stream=mgr._train_stream, **kw):
                    ops.noise_epoch_advance()
                    out = _detached(mgr._step_body(static, self.mode, self.optimizer, self.bucket))
                self.graphs[key] = (g, out)
                self.captured_launches[key] = _lib.launch_count() - n0
            except Exception as e:
                # a failed capture leaves torch's CUDA generator and caching allocator in capture mode:
                # there is no clean way back to eager execution in this process
                self.disabled = True
                raise RuntimeError("deepsee_b200: CUDA graph capture of the %s step failed (%r). Run with "
                                   "DSEE_CUDA_GRAPHS=0 (bench.py --no-graph) for eager execution."
                                   % (self.mode, e)) from e
            finally:
                random.random = real
                config.check_onehot, ops.KernelTimer.active = saved_check, saved_timer
        g, out = self.graphs[key]
        g.replay()
        mgr.graph_replays += 1
        mgr.graph_launches += self.captured_launches[key]
        return out


class TrainerManager(BaseManager):
    def __init__(self, opt):
        super().__init__(opt, create_model=True)
        assert opt.isTrain
        model = self.sr_model_on_one_gpu
        self.optimizer_G, self.optimizer_D = model.create_optimizers(opt)
        self.old_lr = opt.lr
        self.generated = None
        self.logs = {}
        self.g_losses, self.d_losses = {}, {}
        self._bucket_G = self._bucket_D = None
        self._graphed = {}
        self._train_stream = torch.cuda.Stream() if config.cuda_graphs else None
        self.graph_replays = 0       # sub-steps executed as graph replays
        self.graph_launches = 0      # kernels of this library inside those replays
        if parallel.is_dist():
            g_params = list(model.netSR.parameters()) + (list(model.netE.parameters()) if model.use_E else [])
            self._bucket_G = parallel.GradBucket(g_params)
            self._bucket_D = parallel.GradBucket(list(model.netD.parameters()))

    # ---- one optimizer step (trainer_manager.py:32-61) ------------------------------------------------
    def _optimize(self, data, mode, optimizer, bucket):
        if config.cuda_graphs:
            step = self._graphed.get(mode)
            if step is None:
                step = self._graphed[mode] = _GraphedStep(self, mode, optimizer, bucket)
            return step(data)
        return self._step_body(data, mode, optimizer, bucket)

    def _step_body(self, data, mode, optimizer, bucket):
        """zero_grad -> forward(mode) -> mean of the summed losses -> backward -> gradient all-reduce
        (multi-GPU) -> optional value clipping -> optimizer step.  Returns SRModel.forward's result."""
        optimizer.zero_grad()
        if bucket is not None:
            bucket.begin()        # .grad = views of the flat bucket; chunks all-reduce during backward
        result = self.sr_model(self.preprocess_input(data), mode=mode)
        losses = result[0] if mode == 'generator' else result
        sum(losses.values()).mean().backward()
        if bucket is not None:
            bucket.finish()
        if self.opt.gradient_clip > 0:
            clip_grad_value_(self.sr_model.parameters(), self.opt.gradient_clip)
        optimizer.step()
        return result

    def warm_graphs(self, data):
        """config.cuda_graphs: capture every variant of both sub-steps on `data` now instead of on
        first encounter (each capture is one real training step).  Call after a few ordinary steps."""
        for mode in ('generator', 'discriminator'):
            step = self._graphed.get(mode)
            if step is not None:
                step.capture_all(data)

    def graphs_active(self):
        return bool(self._graphed) and all(not s_.disabled and s_.graphs for s_ in self._graphed.values())

    def run_generator_one_step(self, data):
        self.g_losses, self.generated = self._optimize(data, 'generator', self.optimizer_G, self._bucket_G)

    def run_discriminator_one_step(self, data):
        self.d_losses = self._optimize(data, 'discriminator', self.optimizer_D, self._bucket_D)

    # ---- accessors train.py uses ----------------------------------------------------------------------
    def preprocess_input(self, data):
        return super().preprocess(data, from_dataloader=True)

    def get_logs(self):
        return {**self.logs, **self.sr_model_on_one_gpu.get_logs()}

    def get_latest_losses(self):
        return {**self.g_losses, **self.d_losses}

    def get_latest_generated(self):
        return self.generated

    def save(self, epoch):
        if parallel.rank() == 0:
            self.sr_model_on_one_gpu.save(epoch)

    def update_learning_rate(self, epoch):
        """trainer_manager.py:76-99: constant for `niter` epochs, then a linear decay by
        lr / niter_decay per epoch; with TTUR the generator runs at half and the discriminator at
        twice the base rate."""
        decay = self.opt.lr / self.opt.niter_decay if epoch > self.opt.niter else 0.0
        new_lr = self.old_lr - decay
        if new_lr == self.old_lr:
            return
        g_scale, d_scale = (1.0, 1.0) if self.opt.no_TTUR else (0.5, 2.0)
        for optimizer, scale in ((self.optimizer_G, g_scale), (self.optimizer_D, d_scale)):
            for group in optimizer.param_groups:
                group['lr'] = new_lr * scale
        print('update learning rate: %f -> %f' % (self.old_lr, new_lr))
        self.old_lr = new_lr
