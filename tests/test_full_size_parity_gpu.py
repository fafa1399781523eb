"""GPU parity at the BENCHED shapes and in the BENCHED precision mode.

bench.py measures `config.passes = 1` with the default mixed-precision rule (config.passes3_upto =
"auto": the main convs of every stage up to a quarter of the output resolution run 3 passes).  These
tests run the FULL-SIZE generators of BASELINE.json's configurations - c2/c3: 8x 256x256, 512
channels, 5 blocks; c4/c5: 32x 512x512, 7 blocks incl. the PureSEAN tail with max_fm_size = 256 -
through the C ABI in exactly that mode, eval and train mode (batch statistics, running-stat updates,
noise injection with the oracle's noise tensors), against the CPU oracle on conditioned weights,
and assert north_star's bound: max-abs < 1e-3 on the tanh output.  max-abs / std is printed.

This also exercises the B*H*W-scale paths nothing else in tests/ reaches: bn_finalize over tens of
thousands of tile partials, the folded-upsample statistics count, 512 x 512 tiles.
(The oracle needs ~1-3 s per image at these sizes on the GPU box's host cores.)"""
import pytest
import torch

from oracle import deepsee_oracle as O
from test_generator_gpu import _build_G

pytestmark = pytest.mark.gpu

NORTH_STAR_TOL = 1e-3   # BASELINE.json: "outputs within 1e-3 max-abs of reference"


class bench_precision:
    """The precision mode bench.py runs in (and restores the library default afterwards)."""

    def __enter__(self):
        from deepsee_b200.config import config
        self.saved = (config.passes, config.passes3_upto)
        config.passes, config.passes3_upto = 1, "auto"

    def __exit__(self, *exc):
        from deepsee_b200.config import config
        config.passes, config.passes3_upto = self.saved
        return False


def _run_case(name, train, batch, seed):
    o = O.make_opt(name, is_train=train)
    sd = O.make_generator_state(o, seed)
    d = O.preprocess(o, O.synthetic_batch(o, batch, seed=70 + seed))
    z = torch.rand(batch, 19, 128, generator=torch.Generator().manual_seed(5 + seed)) * 2 - 1
    noises = {}

    def noise_fn(nm, shape):
        noises[nm] = torch.randn(shape, generator=torch.Generator().manual_seed(len(noises)))
        return noises[nm]

    sd_ref = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        ref = O.generator_forward(sd_ref, o, d["image_lr"], d["input_semantics"], z, train,
                                  noise_fn if train else None)
    G = _build_G(o, sd).train(train)
    if train and o.add_noise:
        for pfx, _, _ in O.generator_layout(o):
            blk = G.get_submodule(pfx[:-1])
            for nm in ("noise_in", "noise_skip", "noise_middle"):
                n = noises[pfx + nm].permute(0, 2, 3, 1).contiguous().cuda()
                getattr(blk, nm).sample = (lambda t: (lambda B, H, W: t))(n)
    with bench_precision(), torch.no_grad():
        out = G(d["image_lr"].cuda(), seg=d["input_semantics"].cuda(), z=z.cuda())
    e = (out.cpu() - ref).abs()
    std = ref.std().item()
    print("%s %s batch %d: bench-mode max-abs vs oracle %.3e (mean-abs %.3e, max-abs/std %.3e, ref std %.3f)"
          % (name, "train" if train else "eval", batch, e.max().item(), e.mean().item(),
             e.max().item() / std, std))
    return e.max().item(), G, sd_ref


@pytest.mark.parametrize("train", [False, True])
def test_c2_generator_full_size_bench_precision(train):
    """8x 256x256 (c2 / c3): 512 channels, blocks at 32, 64, 64, 128, 256."""
    err, G, sd_ref = _run_case("8x_independent_256x256", train, 2 if not train else 1, 0)
    assert err < NORTH_STAR_TOL
    if train:
        # B*H*W-scale statistics: the running stats of the last (256 x 256) norm layer
        got = G.state_dict()
        for k in ("up_list.1.norm_1.param_free_norm.running_var",
                  "up_list.1.norm_0.param_free_norm.running_mean"):
            torch.testing.assert_close(got[k].cpu(), sd_ref[k], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("train", [False, True])
def test_c4_generator_full_size_bench_precision(train):
    """32x 512x512 (c4 / c5): 7 blocks at 16 ... 512; the last one is the PureSEAN block above
    max_fm_size (normalization.py:275-277)."""
    err, G, sd_ref = _run_case("32x_independent_512x512", train, 1, 0)
    assert err < NORTH_STAR_TOL
    if train:
        got = G.state_dict()
        k = "up_list.3.norm_1.param_free_norm.running_var"
        torch.testing.assert_close(got[k].cpu(), sd_ref[k], rtol=1e-3, atol=1e-4)


def test_c2_generator_second_seed_bench_precision():
    """Another draw of weights and inputs (train mode, the worse of the two modes)."""
    err, _, _ = _run_case("8x_independent_256x256", True, 1, 3)
    assert err < NORTH_STAR_TOL
