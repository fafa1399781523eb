"""GPU: one full training iteration (TrainerManager.run_generator_one_step +
run_discriminator_one_step: style encoder, generator, multi-scale discriminator, hinge + feature
matching losses, backward, Adam) through the reference-shaped managers, against the oracle's CPU
restatement of trainer_manager.py:32-61 on the same seeded inputs, noise and coin flips.

Losses (forward values) are compared tightly.  Gradients are compared by relative L2 error per
network with a loose bound: both sides run freely here, so LeakyReLU / hinge kink flips caused by
the 1e-5 forward differences contribute ~1e-2 (tests/test_backward_gpu.py pins the activation
pattern and asserts the tight bound)."""
import random
import zlib

import pytest
import torch

from oracle import deepsee_oracle as O
from test_generator_gpu import _mk_opt

pytestmark = pytest.mark.gpu


def _noise(name, k, shape):
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * k) % (2 ** 31))
    return torch.randn(shape, generator=g)


def _unit(k, shape):
    return torch.rand(shape, generator=torch.Generator().manual_seed(1000 + k))


class _OracleTrainer(O.CpuTrainer):
    def __init__(self, *a):
        super().__init__(*a)
        self.fwd = 0
        self.enc = 0

    def _noise_fn(self, name, shape):
        return _noise(name, self.fwd, shape)

    def _fake(self, d):
        self.fwd += 1
        return super()._fake(d)

    def _style_noise(self, shape):
        self.enc += 1
        return _unit(self.enc, shape)


def _grads(named):
    return {k: (p.grad.detach().cpu().clone() if p.grad is not None else None) for k, p in named}


def _rel_l2(ours, ref):
    num = den = 0.0
    for k, r in ref.items():
        if r is None:
            continue
        o = ours[k]
        assert o is not None, "missing gradient for %s" % k
        num += float((o - r).pow(2).sum())
        den += float(r.pow(2).sum())
    return (num / max(den, 1e-30)) ** 0.5


@pytest.mark.parametrize("mode", ["passes3", "bench"])
@pytest.mark.parametrize("name,over", [
    ("8x_independent_256x256", dict(ngf=8, nef=8, ndf=8, start_size=8, crop_size=64, load_size=64)),
    ("32x_guided_512x512", dict(ngf=8, nef=8, ndf=8, start_size=4, crop_size=128, load_size=512,
                                max_fm_size=64)),
])
def test_train_iteration_vs_oracle(name, over, mode):
    """mode "passes3": the library default (fp32-class).  mode "bench": the precision bench.py
    measures (1 pass, main convs of the low-resolution stages 3 passes) - forward values within
    north_star's 1e-3, losses to 5e-3 relative, gradients by the same loose rel-L2 bound."""
    from deepsee_b200.config import config
    from test_full_size_parity_gpu import bench_precision
    import contextlib
    # the discriminator's gradient is compared free-running: its hinge loss and LeakyReLUs flip on
    # the ~2e-4 differences of the bench-mode fake image, and with these toy sizes (a few hundred
    # prediction elements, 8-channel layers) a handful of flips is 10 % of the gradient norm
    with (bench_precision() if mode == "bench" else contextlib.nullcontext()):
        _train_iteration(name, over, 5e-3 if mode == "bench" else 2e-4, 1e-3 if mode == "bench" else 3e-4,
                         1e-2 if mode == "bench" else 2e-3, 2.5e-1 if mode == "bench" else 5e-2)
    assert config.passes == 3


def _train_iteration(name, over, tol_g, tol_img, tol_d, tol_dgrad):
    from deepsee_b200.managers.trainer_manager import TrainerManager
    o = O.make_opt(name, is_train=True, **over)
    sdG, sdE, sdD = O.make_generator_state(o, 0), O.make_encoder_state(o, 1), O.make_discriminator_state(o, 2)
    clone = lambda sd: {k: v.clone() for k, v in sd.items()}
    ref = _OracleTrainer(o, clone(sdG), clone(sdE), clone(sdD))
    raw = O.synthetic_batch(o, 2, seed=99)
    d = O.preprocess(o, raw)

    mgr = TrainerManager(_mk_opt(o))
    m = mgr.sr_model
    m.netSR.load_state_dict(sdG, strict=True)
    m.netE.load_state_dict(sdE, strict=True)
    m.netD.load_state_dict(sdD, strict=True)
    m.train()
    random.seed(0)  # same coin-flip stream as the oracle's random.Random(0)
    state = {"fwd": 0, "enc": 0}
    m.netSR.register_forward_pre_hook(lambda *_: state.__setitem__("fwd", state["fwd"] + 1))
    for bname, blk in m.netSR.named_modules():
        for nm in ("noise_in", "noise_skip", "noise_middle"):
            if hasattr(blk, nm):
                full = "%s.%s" % (bname, nm)
                getattr(blk, nm).sample = (lambda full: (
                    lambda B, H, W: _noise(full, state["fwd"], (B, blk_c(m), H, W)).permute(0, 2, 3, 1)
                    .contiguous().cuda()))(full)

    def unit_noise(like):
        state["enc"] += 1
        return _unit(state["enc"], tuple(like.shape)).cuda()
    m.netE._unit_noise = unit_noise

    def batch():
        return {k: v.clone() for k, v in raw.items()}

    # ---- generator step ----
    g_ref, fake_ref = ref.generator_step(d)
    ref_gG = {k: (v.grad.clone() if v.grad is not None else None) for k, v in ref.sdG.items() if v.requires_grad}
    ref_gE = {k: (v.grad.clone() if v.grad is not None else None) for k, v in ref.sdE.items() if v.requires_grad}
    data = batch()
    data["label"] = data["label"].float()
    if "guiding_label" in data:
        data["guiding_label"] = data["guiding_label"].float()
    mgr.run_generator_one_step(data)
    ours = mgr.get_latest_losses()
    for k in g_ref:
        a, b = float(ours[k].mean()), float(g_ref[k].mean())
        print("G loss %-9s ours %.6f oracle %.6f" % (k, a, b))
        assert abs(a - b) < tol_g * max(1.0, abs(b))
    img_err = (mgr.get_latest_generated().detach().cpu() - fake_ref).abs().max().item()
    print("generated image max-abs vs oracle %.3e" % img_err)
    assert img_err < tol_img
    eG = _rel_l2(_grads(m.netSR.named_parameters()), ref_gG)
    eE = _rel_l2(_grads(m.netE.named_parameters()), ref_gE)
    print("G-step gradient rel-L2: generator %.3e encoder %.3e" % (eG, eE))
    assert eG < 5e-2 and eE < 5e-2

    # ---- discriminator step ----
    d_ref = ref.discriminator_step(d)
    ref_gD = {k: (v.grad.clone() if v.grad is not None else None) for k, v in ref.sdD.items() if v.requires_grad}
    data = batch()
    data["label"] = data["label"].float()
    if "guiding_label" in data:
        data["guiding_label"] = data["guiding_label"].float()
    mgr.run_discriminator_one_step(data)
    ours = mgr.get_latest_losses()
    for k in d_ref:
        a, b = float(ours[k].mean()), float(d_ref[k].mean())
        print("D loss %-9s ours %.6f oracle %.6f" % (k, a, b))
        assert abs(a - b) < tol_d * max(1.0, abs(b))   # G weights already moved by one Adam step
    eD = _rel_l2(_grads(m.netD.named_parameters()), ref_gD)
    print("D-step gradient rel-L2: discriminator %.3e" % eD)
    assert eD < tol_dgrad
    # spectral-norm power iteration / BN running statistics advanced like the reference's
    got = m.netSR.state_dict()
    for k in ("head_0.norm_0.param_free_norm.running_mean", "G_middle_1.conv_1.weight_u"):
        torch.testing.assert_close(got[k].cpu(), ref.sdG[k].detach(), rtol=5e-3, atol=5e-4)
    assert int(got["head_0.norm_0.param_free_norm.num_batches_tracked"]) == 102


def blk_c(m):
    return m.netSR.initial.weight.shape[0]
