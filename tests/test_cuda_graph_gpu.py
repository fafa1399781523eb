"""GPU: TrainerManager with config.cuda_graphs - every optimizer sub-step captured per encoder
coin-flip variant and replayed - must train exactly like the eager path: same losses step by step and
the same parameters afterwards (no noise injection / style noise here, so both runs are deterministic;
every reduction in the library has a fixed order).  Also: the device-side noise epoch that lets a
replayed graph draw fresh NoiseInjection noise."""
import random

import pytest
import torch

from oracle import deepsee_oracle as O
from test_generator_gpu import _mk_opt

pytestmark = pytest.mark.gpu


def _run(graphs, steps, name, over):
    from deepsee_b200.config import config
    from deepsee_b200.managers.trainer_manager import TrainerManager
    saved = config.cuda_graphs
    # both runs build their optimizers with capturable Adam (device-side step counters and bias
    # corrections), so the eager reference executes exactly the kernels the graphs replay
    config.cuda_graphs = True
    try:
        o = O.make_opt(name, is_train=True, **over)
        mgr = TrainerManager(_mk_opt(o))
        config.cuda_graphs = graphs
        m = mgr.sr_model
        m.netSR.load_state_dict(O.make_generator_state(o, 0), strict=True)
        m.netE.load_state_dict(O.make_encoder_state(o, 1), strict=True)
        m.netD.load_state_dict(O.make_discriminator_state(o, 2), strict=True)
        m.train()
        random.seed(123)
        losses = []
        for i in range(steps):
            if i == 5:      # a learning-rate change (baked into the captured Adam nodes: forces a re-capture)
                for opt_ in (mgr.optimizer_G, mgr.optimizer_D):
                    for grp in opt_.param_groups:
                        grp['lr'] = grp['lr'] * 0.5
            # step 7: a batch of another size (the last batch of an epoch) runs eagerly in graph mode
            raw = O.synthetic_batch(o, 1 if i == 7 else 2, seed=500 + i)
            data = {k: (v.float() if "label" in k else v) for k, v in raw.items()}
            mgr.run_generator_one_step(dict(data))
            mgr.run_discriminator_one_step(dict(data))
            losses.append({k: float(v.detach().mean()) for k, v in mgr.get_latest_losses().items()})
        torch.cuda.synchronize()
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        return losses, sd, mgr
    finally:
        config.cuda_graphs = saved


@pytest.mark.parametrize("name,over", [
    ("8x_independent_256x256", dict(ngf=8, nef=8, ndf=8, start_size=8, crop_size=64, load_size=64,
                                    add_noise=False, noisy_style_scale=0.0)),
    ("32x_guided_512x512", dict(ngf=8, nef=8, ndf=8, start_size=4, crop_size=128, load_size=512,
                                max_fm_size=64, noisy_style_scale=0.0)),
])
def test_graphed_training_equals_eager(name, over):
    steps = 9
    l_eager, sd_eager, _ = _run(False, steps, name, over)
    assert _.graph_replays == 0
    l_graph, sd_graph, mgr = _run(True, steps, name, over)
    assert mgr.graphs_active() and mgr.graph_replays > 0, "the graphed path did not engage"
    variants = {m: len(s.graphs) for m, s in mgr._graphed.items()}
    print("captured variants:", variants, "replayed sub-steps:", mgr.graph_replays)
    for i, (a, b) in enumerate(zip(l_eager, l_graph)):
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-6 * max(1.0, abs(a[k])), (i, k, a[k], b[k])
    for k in sd_eager:
        if sd_eager[k].dtype.is_floating_point:
            torch.testing.assert_close(sd_graph[k], sd_eager[k], rtol=1e-6, atol=1e-7, msg=k)
        else:
            assert torch.equal(sd_graph[k], sd_eager[k]), k


def test_noise_epoch_changes_the_stream():
    from deepsee_b200 import ops
    a = ops.noise_fill(1234, (1, 4, 4, 8))
    b = ops.noise_fill(1234, (1, 4, 4, 8))
    assert torch.equal(a, b)
    ops.noise_epoch_advance()
    c = ops.noise_fill(1234, (1, 4, 4, 8))
    assert not torch.equal(a, c)
    assert abs(float(c.mean())) < 1.0 and 0.3 < float(c.std()) < 2.0


def test_optimizer_state_of_late_parameters_survives_other_graphs():
    """A parameter whose first gradient arrives only after captures have begun (the style-noise weights
    when the eager calls drew 'clean' style) must not get its Adam state from a graph's private pool:
    replays of the graphs captured earlier would scribble over it.  Scripted coin flips: two clean eager
    iterations, then clean / noisy / full / mini variants in an interleaved order; the graph-replayed run
    must match the eager run parameter by parameter (style noise made a fixed function of the shape)."""
    from deepsee_b200.config import config
    from deepsee_b200.deepsee_models.networks import encoder as enc_mod
    from deepsee_b200.managers.trainer_manager import TrainerManager
    name = "8x_independent_256x256"
    over = dict(ngf=8, nef=8, ndf=8, start_size=8, crop_size=64, load_size=64, add_noise=False)
    full_clean, mini_clean, mini_noisy, full_noisy = (0.25, 0.25), (0.75, 0.25), (0.75, 0.75), (0.25, 0.75)
    g_script = [mini_clean, mini_clean, full_clean, mini_noisy, full_noisy, full_clean, mini_noisy, mini_clean,
                full_noisy, full_clean, mini_noisy, full_noisy, mini_clean, full_noisy]

    def fixed_noise(self, like):
        n = like.numel()
        return (torch.arange(n, device=like.device, dtype=torch.float32) * 0.6180339887).frac().view_as(like)

    def run(graphs):
        saved, real = config.cuda_graphs, random.random
        patched = [c for c in vars(enc_mod).values() if isinstance(c, type) and "_unit_noise" in vars(c)]
        olds = [(c, c._unit_noise) for c in patched]
        config.cuda_graphs = True
        try:
            for c in patched:
                c._unit_noise = fixed_noise
            o = O.make_opt(name, is_train=True, **over)
            mgr = TrainerManager(_mk_opt(o))
            config.cuda_graphs = graphs
            m = mgr.sr_model
            m.netSR.load_state_dict(O.make_generator_state(o, 0), strict=True)
            m.netE.load_state_dict(O.make_encoder_state(o, 1), strict=True)
            m.netD.load_state_dict(O.make_discriminator_state(o, 2), strict=True)
            m.train()
            raw = O.synthetic_batch(o, 2, seed=77)
            data = {k: (v.float() if "label" in k else v) for k, v in raw.items()}
            for flips in g_script:
                seq = iter(list(flips) + list(flips))     # generator sub-step, then discriminator sub-step
                random.random = lambda: next(seq)
                mgr.run_generator_one_step(dict(data))
                mgr.run_discriminator_one_step(dict(data))
            torch.cuda.synchronize()
            return {k: v.detach().clone() for k, v in m.state_dict().items()}, mgr
        finally:
            random.random = real
            config.cuda_graphs = saved
            for c, f in olds:
                c._unit_noise = f

    sd_eager, _ = run(False)
    sd_graph, mgr = run(True)
    assert mgr.graphs_active() and len(mgr._graphed['generator'].graphs) == 4
    noisy_keys = [k for k in sd_eager if k.endswith("noise_weights")]
    assert noisy_keys and all(float(sd_eager[k].abs().max()) > 0 for k in noisy_keys), "the noisy branch never trained"
    for k in sd_eager:
        assert bool(torch.isfinite(sd_graph[k].float()).all()), k
        if sd_eager[k].dtype.is_floating_point:
            torch.testing.assert_close(sd_graph[k], sd_eager[k], rtol=1e-6, atol=1e-7, msg=k)
