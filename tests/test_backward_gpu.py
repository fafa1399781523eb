"""GPU parity of the generator's backward pass (C-ABI backward kernels driven by the block-level
autograd node) against CPU autograd through the oracle on the same seeded inputs.

LeakyReLU's kink makes a raw comparison meaningless at tight tolerances: a forward difference of
1e-5 flips the side of ~1e-5 of all pre-activations, and each flip changes that element's gradient
by a factor of 5.  The oracle therefore evaluates its LeakyReLUs with the activation pattern the
CUDA forward took (``act_fn`` hook of the oracle), which makes both sides differentiate the same
piecewise-linear function.  Tolerance: per tensor, max-abs error / max(|reference gradient|,
1e-4 * largest gradient entry of the model): 3-pass operands (fp16 hi+lo planes, fp32 accumulate)
are asserted at 2e-3, the 1-pass (TF32-class) mode at 5e-2; measured figures are printed."""
import pytest
import torch

from oracle import deepsee_oracle as O
from test_generator_gpu import _build_G

pytestmark = pytest.mark.gpu


def _is_param(k, v):
    return v.is_floating_point() and not O._is_buffer(k)


def _run_case(name, over, passes, batch=2, train=True, seed=31):
    from deepsee_b200.config import config
    o = O.make_opt(name, is_train=True, **over)
    sd = O.make_generator_state(o, 0)
    sd_ref = {k: v.clone() for k, v in sd.items()}
    for k, v in sd_ref.items():
        if _is_param(k, v):
            v.requires_grad_(True)
    d = O.preprocess(o, O.synthetic_batch(o, batch, seed=seed))
    g = torch.Generator().manual_seed(seed + 1)
    z_ref = (torch.rand(batch, 19, 128, generator=g) * 2 - 1).requires_grad_(True)
    proj = torch.randn(batch, 3, o.crop_size, o.crop_size, generator=g)
    noises = {}

    def noise_fn(nm, shape):
        noises[nm] = torch.randn(shape, generator=torch.Generator().manual_seed(len(noises)))
        return noises[nm]

    if o.add_noise and train:  # draw the noise tensors (names / shapes) on a throw-away copy
        with torch.no_grad():
            O.generator_forward({k: v.clone() for k, v in sd.items()}, o, d["image_lr"],
                                d["input_semantics"], z_ref, train, noise_fn)
    from deepsee_b200 import ops
    masks = []
    orig_mod, orig_head, orig_head_tc = ops.spade_modulate, ops.head, ops.head_tc

    def rec_mod(*a, **k):
        r = orig_mod(*a, **k)
        act = r if hasattr(r, "hi") else r[0]  # (activation planes, saved G planes) when training
        masks.append((act.hi > 0).permute(0, 3, 1, 2).cpu())
        return r

    def rec_head(x, *a, **k):
        masks.append((x > 0).permute(0, 3, 1, 2).cpu())
        return orig_head(x, *a, **k)

    def rec_head_tc(planes, *a, **k):   # tensor-core head: the mask its backward uses is the hi plane's sign
        masks.append((planes.hi > 0).permute(0, 3, 1, 2).cpu())
        return orig_head_tc(planes, *a, **k)

    old = config.passes
    config.passes = passes
    try:
        G = _build_G(o, sd)
        G.train(train)
        if o.add_noise and train:
            for pfx, _, _ in O.generator_layout(o):
                blk = G.get_submodule(pfx[:-1])
                for nm in ("noise_in", "noise_skip", "noise_middle"):
                    n = noises[pfx + nm].permute(0, 2, 3, 1).contiguous().cuda()
                    getattr(blk, nm).sample = (lambda t: (lambda B, H, W: t))(n)
        z = z_ref.detach().clone().cuda().requires_grad_(True)
        ops.spade_modulate, ops.head, ops.head_tc = rec_mod, rec_head, rec_head_tc
        out = G(d["image_lr"].cuda(), seg=d["input_semantics"].cuda(), z=z)
        (out * proj.cuda()).sum().backward()
        torch.cuda.synchronize()
    finally:
        config.passes = old
        ops.spade_modulate, ops.head, ops.head_tc = orig_mod, orig_head, orig_head_tc
    keys = []
    for pfx, _, _ in O.generator_layout(o):
        keys += [pfx + "act_0", pfx + "act_1"]
    keys.append("head")
    pattern = dict(zip(keys, masks))
    assert len(masks) == len(keys)
    flips = [0, 0]

    def act_fn(key, t):
        m = pattern[key]
        flips[0] += int(((t > 0) != m).sum())
        flips[1] += t.numel()
        return torch.where(m, t, O.LRELU * t)

    ref = O.generator_forward(sd_ref, o, d["image_lr"], d["input_semantics"], z_ref, train,
                              (lambda nm, shape: noises[nm]) if noises else noise_fn, act_fn=act_fn)
    (ref * proj).sum().backward()
    print("   activation-pattern flips vs a free-running oracle: %d of %d" % tuple(flips))
    fwd_err = (out.detach().cpu() - ref.detach()).abs().max().item()
    worst = ("", 0.0)
    checked = 0
    # a gradient that is mathematically zero (a conv bias in front of a training-mode batch norm) is
    # rounding noise on both sides: errors are measured against max(|ref grad|, 1e-4 * the largest
    # gradient entry of the whole model)
    gmax = max(float(v.grad.abs().max()) for v in sd_ref.values() if getattr(v, "grad", None) is not None)
    report = []
    for k, p in G.named_parameters():
        rg = sd_ref[k].grad
        if rg is None:
            # parameters the reference never touches (style_conv, unused trailing blocks, mlp_shared
            # of a PureSEAN layer below max_fm_size)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, "missing gradient for %s" % k
        scale = rg.abs().max().item()
        err = (p.grad.cpu() - rg).abs().max().item()
        rel = err / max(scale, 1e-4 * gmax)
        report.append((rel, k, err, scale))
        checked += 1
        if rel > worst[1]:
            worst = (k, rel)
    for rel, k, err, scale in sorted(report, reverse=True)[:8]:
        print("   %-55s rel %.2e abs %.2e ref-scale %.2e (gmax %.2e)" % (k, rel, err, scale, gmax))
    zrel = (z.grad.cpu() - z_ref.grad).abs().max().item() / max(z_ref.grad.abs().max().item(), 1e-12)
    return fwd_err, worst, zrel, checked


CASES = {
    # 8x preset: SPADE head + SEAN blocks, noise injection on in training
    "8x": ("8x_independent_256x256", dict(ngf=8, start_size=8, crop_size=64, load_size=64)),
    # 32x preset, scaled down so the max_fm_size quirk and the PureSEAN tail are both exercised
    "32x": ("32x_independent_512x512", dict(ngf=8, start_size=4, crop_size=128, load_size=512,
                                            max_fm_size=64)),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("passes", [3, 1])
def test_generator_backward_vs_oracle(case, passes):
    name, over = CASES[case]
    fwd_err, worst, zrel, checked = _run_case(name, over, passes)
    print("%s passes=%d: fwd max-abs %.3e; worst param-grad rel err %.3e (%s); dz rel err %.3e; "
          "%d tensors" % (case, passes, fwd_err, worst[1], worst[0], zrel, checked))
    tol = 2e-3 if passes == 3 else 5e-2
    assert fwd_err < (3e-4 if passes == 3 else 1e-2)
    if passes == 1 and "alpha_" in worst[0]:
        # d alpha = sigma'(alpha) * (<dW, W_style> - <dW, W_seg>): a scalar difference of two large
        # sums, so the 1e-3 relative error of 1-pass operands is amplified by the cancellation
        assert worst[1] < 2e-1, worst
    else:
        assert worst[1] < tol, worst
    assert zrel < tol
    assert checked > 50


@pytest.mark.parametrize("save_gamma,fuse", [(True, False), (False, False)])
def test_generator_backward_alternate_kernel_paths(save_gamma, fuse):
    """The default backward runs backward-data and K1's backward as one kernel (saved G planes).  The
    two other routes stay selectable and must agree with the oracle just as well: streaming K1
    backward on the saved G planes (DSEE_FUSE_DGRAD_MODBWD=0) and the gamma-GEMM-recompute kernel
    (DSEE_SAVE_GAMMA=0; explicit noise tensors, which is what this harness injects)."""
    from deepsee_b200.config import config
    old = (config.save_gamma, config.fuse_dgrad_modbwd)
    config.save_gamma, config.fuse_dgrad_modbwd = save_gamma, fuse
    try:
        name, over = CASES["8x"]
        fwd_err, worst, zrel, checked = _run_case(name, over, 3)
    finally:
        config.save_gamma, config.fuse_dgrad_modbwd = old
    print("save_gamma=%s fuse=%s: worst param-grad rel err %.3e (%s); dz rel err %.3e" %
          (save_gamma, fuse, worst[1], worst[0], zrel))
    assert worst[1] < 2e-3 and zrel < 2e-3 and checked > 50


def test_generator_backward_eval_mode():
    """Gradients with running statistics (eval-mode batch norm, no noise)."""
    name, over = CASES["8x"]
    fwd_err, worst, zrel, _ = _run_case(name, over, 3, train=False)
    print("eval-mode: fwd %.3e worst %.3e (%s) dz %.3e" % (fwd_err, worst[1], worst[0], zrel))
    assert worst[1] < 2e-3 and zrel < 2e-3


def test_in_kernel_noise_matches_materialised_stream():
    """NoiseInjection in production is a seed: every kernel regenerates the tensor's elements
    (Philox4x32-7 + Box-Muller).  (a) the stream is N(0,1); (b) a training forward+backward run on
    seeds equals, bit for bit, the same run fed the tensors dsee_noise_fill materialises from those
    seeds - i.e. statistics pass, K1, K2 epilogues and the backward kernels all see identical noise."""
    from deepsee_b200 import ops
    z = ops.noise_fill(12345, (4, 32, 32, 128))
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.var()) - 1.0) < 1e-2
    assert abs(float((z ** 4).mean()) - 3.0) < 0.1          # kurtosis of a normal
    flat = z.flatten()
    assert abs(float((flat[:-1] * flat[1:]).mean())) < 5e-3  # neighbours uncorrelated
    assert not torch.equal(z, ops.noise_fill(12346, (4, 32, 32, 128)))

    name, over = CASES["8x"]
    o = O.make_opt(name, is_train=True, **over)
    sd = O.make_generator_state(o, 0)
    d = O.preprocess(o, O.synthetic_batch(o, 2, seed=3))
    zst = (torch.rand(2, 19, 128, generator=torch.Generator().manual_seed(4)) * 2 - 1).cuda()
    proj = torch.randn(2, 3, o.crop_size, o.crop_size, generator=torch.Generator().manual_seed(5)).cuda()

    def run(materialise):
        G = _build_G(o, sd).train()
        seeds = {}
        for pfx, _, _ in O.generator_layout(o):
            blk = G.get_submodule(pfx[:-1])
            for k, nm in enumerate(("noise_in", "noise_skip", "noise_middle")):
                seed = 1000 + 7 * len(seeds)
                seeds[pfx + nm] = seed
                C = getattr(blk, nm).n_channels
                if materialise:
                    getattr(blk, nm).sample = (lambda s, C: (lambda B, H, W: ops.noise_fill(s, (B, H, W, C))))(seed, C)
                else:
                    getattr(blk, nm).sample = (lambda s: (lambda B, H, W: ops.NoiseSeed(s)))(seed)
        zz = zst.clone().requires_grad_(True)
        out = G(d["image_lr"].cuda(), seg=d["input_semantics"].cuda(), z=zz)
        (out * proj).sum().backward()
        torch.cuda.synchronize()
        return out.detach(), zz.grad, {k: p.grad for k, p in G.named_parameters() if p.grad is not None}

    out_s, dz_s, g_s = run(False)
    out_t, dz_t, g_t = run(True)
    assert torch.equal(out_s, out_t)
    assert torch.equal(dz_s, dz_t)
    for k in g_t:
        assert torch.equal(g_s[k], g_t[k]), k
    assert any("noise_in.weight" in k for k in g_s)
