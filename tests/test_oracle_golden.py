"""CPU: the oracle restatement against golden vectors produced by the unmodified reference
(oracle/make_golden.py, run in the build container where /root/reference exists)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import deepsee_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {
    "g8x_eval": ("8x_independent_256x256", dict(ngf=8, start_size=8, crop_size=64, load_size=64,
                                                 max_fm_size=256)),
    "g32x_eval": ("32x_independent_512x512", dict(ngf=8, start_size=4, crop_size=128, load_size=512,
                                                  max_fm_size=64)),
    "g8x_train": ("8x_independent_256x256", dict(ngf=8, start_size=8, crop_size=64, load_size=64,
                                                  max_fm_size=256)),
}


def _checksum(*sds):
    return sum(float(v.double().abs().sum()) for sd in sds for v in sd.values())


def _load(tag):
    return np.load(os.path.join(GOLD, tag + ".npz"), allow_pickle=False)


def _onehot(labels, L=19):
    return torch.from_numpy(O.preprocess_label_np(labels[:, None].astype(np.int64), L))


@pytest.mark.parametrize("tag", ["g8x_eval", "g32x_eval"])
def test_generator_eval_matches_reference(tag):
    g = _load(tag)
    name, over = CASES[tag]
    o = O.make_opt(name, **over)
    sd = O.make_generator_state(o, 0)
    assert abs(_checksum(sd) - float(g["weights_checksum"])) < 1e-6 * float(g["weights_checksum"]), \
        "seeded weights differ from the ones the golden was generated with (torch RNG drift)"
    with torch.no_grad():
        out = O.generator_forward(sd, o, torch.from_numpy(g["x_lr"]), _onehot(g["labels"]),
                                  torch.from_numpy(g["z"]), training=False)
    # same ATen kernels as the reference run -> tolerance only covers thread-count dependent
    # summation order inside mkldnn convs
    assert np.abs(out.numpy() - g["fake"]).max() < 2e-5


def test_generator_train_forward_backward_matches_reference():
    g = _load("g8x_train")
    name, over = CASES["g8x_train"]
    o = O.make_opt(name, is_train=True, **over)
    sd = O.make_generator_state(o, 0)
    for k, v in sd.items():
        if v.is_floating_point() and not O._is_buffer(k):
            v.requires_grad_(True)
    torch.manual_seed(7)  # protocol stored in the golden: randn(shape) per NoiseInjection call
    out = O.generator_forward(sd, o, torch.from_numpy(g["x_lr"]), _onehot(g["labels"]),
                              torch.from_numpy(g["z"]), training=True,
                              noise_fn=lambda name, shape: torch.randn(shape))
    assert np.abs(out.detach().numpy() - g["fake"]).max() < 5e-5
    (out * torch.linspace(-1, 1, out.numel()).view_as(out)).sum().backward()
    np.testing.assert_allclose(sd["G_middle_0.norm_0.param_free_norm.running_var"].numpy(),
                               g["running_var_G_middle_0_norm_0"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd["up_list.0.norm_1.param_free_norm.running_mean"].numpy(),
                               g["running_mean_up_list_0_norm_1"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd["head_0.conv_0.weight_u"].numpy(), g["weight_u_head_0_conv_0"],
                               rtol=1e-4, atol=1e-6)
    gs = sd["head_0.conv_0.weight_orig"].grad.flatten()[::997].numpy()
    ref = g["grad_head_0_conv_0_weight_orig_sample"]
    assert np.abs(gs - ref).max() < 1e-3 * np.abs(ref).max()
    gi = sd["conv_img.weight"].grad.numpy()
    assert np.abs(gi - g["grad_conv_img_weight"]).max() < 1e-3 * np.abs(g["grad_conv_img_weight"]).max()


def test_encoder_discriminator_losses_match_reference():
    g = _load("enc_disc_losses")
    o = O.make_opt("8x_independent_256x256", is_train=True, ngf=8, start_size=8, crop_size=64,
                   load_size=64, add_noise=False)
    sdG, sdE, sdD = O.make_generator_state(o, 0), O.make_encoder_state(o, 1), \
        O.make_discriminator_state(o, 2)
    assert abs(_checksum(sdG, sdE, sdD) - float(g["weights_checksum"])) < 1e-6 * float(g["weights_checksum"])
    data = O.preprocess(o, O.synthetic_batch(o, 2, seed=4321))
    with torch.no_grad():
        zm = O.encoder_forward(sdE, o, data["image_lr"], data["input_semantics"], "mini")
        zf = O.encoder_forward(sdE, o, data["image_hr"], data["input_semantics"], "full")
    assert np.abs(zm.numpy() - g["z_mini"]).max() < 1e-5
    assert np.abs(zf.numpy() - g["z_full"]).max() < 1e-5
    sdD2 = {k: v.clone() for k, v in sdD.items()}
    with torch.no_grad():
        pf, pr = O.discriminate(sdD2, o, data["input_semantics"], torch.from_numpy(g["fake_for_d"]),
                                data["image_hr"], True)
    assert np.abs(pf[0][-1].numpy() - g["d_fake_pred0"]).max() < 2e-5
    assert np.abs(pr[1][-1].numpy() - g["d_real_pred1"]).max() < 2e-5
    # generator-mode losses with the coin flips seeded like the reference run
    sG, sE, sD = ({k: v.clone() for k, v in sd.items()} for sd in (sdG, sdE, sdD))
    rng = random.Random(3)
    torch.manual_seed(3)
    with torch.no_grad():
        z, mode = O.encode_style(sE, o, rng, True, data["image_lr"], data["input_semantics"],
                                 data["image_hr"], None, None)
        fake = O.generator_forward(sG, o, data["image_lr"], data["input_semantics"], z, True)
        lo = O.generator_losses(sD, o, data["input_semantics"], fake, data["image_hr"])
    assert mode == str(g["encoder_mode"])
    assert np.abs(fake.numpy() - g["gen_fake"]).max() < 5e-5
    assert abs(float(lo["GAN"]) - float(g["loss_GAN"])) < 1e-4
    assert abs(float(lo["GAN_Feat"]) - float(g["loss_GAN_Feat"])) < 1e-3


def test_label_and_preprocess_goldens():
    g = _load("labels")
    lab = g["labels"].astype(np.int64)
    oh = O.preprocess_label_np(lab, 19)
    np.testing.assert_array_equal(oh.sum((0, 2, 3)), g["onehot_sum"])
    assert oh.sum(1).min() == 1.0 and oh.sum(1).max() == 1.0
    for s in (8, 12, 24, 32, 48):
        np.testing.assert_array_equal(O.resize_labels_np(lab[:, 0], s, s), g["resized_%d" % s])
    lr = O.downsample_image(torch.from_numpy(g["image_hr"]), 8)
    assert np.abs(lr.numpy() - g["image_lr"]).max() < 1e-6


def test_generator_layouts():
    o = O.make_opt("8x_independent_256x256")
    assert [b[1] for b in O.generator_layout(o)] == ["spade", "sean", "sean", "sean", "sean"]
    o = O.make_opt("32x_independent_512x512")
    lay = O.generator_layout(o)
    assert [b[1] for b in lay] == ["spade", "sean", "sean", "sean", "sean", "sean", "puresean"]
    assert [b[2] for b in lay] == [False, True, False, True, True, True, True]


TRAINER_CASES = {
    "train_iteration": ("8x_independent_256x256", dict(ngf=8, nef=8, ndf=8, start_size=8, crop_size=64,
                                                       load_size=64)),
    "train_iteration_guided": ("32x_guided_512x512", dict(ngf=8, nef=8, ndf=8, start_size=4, crop_size=128,
                                                          load_size=512, max_fm_size=64)),
}


@pytest.mark.parametrize("tag", sorted(TRAINER_CASES))
def test_cpu_trainer_vs_reference_training_iteration(tag):
    """The oracle's CpuTrainer (restatement of trainer_manager.py:32-61 + sr_model.py:469-564) against
    goldens written by the reference's own TrainerManager on CPU: the four losses of one G step + one D
    step and probes of parameters of all three networks after both Adam updates, for the independent
    8x preset and for the guided 32x preset (PureSEAN tail, style encoder on the guiding image)."""
    g = np.load(os.path.join(GOLD, tag + ".npz"))
    name, over = TRAINER_CASES[tag]
    o = O.make_opt(name, is_train=True, add_noise=False, noisy_style_scale=0.0, **over)
    sdG, sdE, sdD = O.make_generator_state(o, 0), O.make_encoder_state(o, 1), O.make_discriminator_state(o, 2)
    cs = sum(float(v.double().abs().sum()) for sd in (sdG, sdE, sdD) for v in sd.values())
    assert abs(cs - float(g["weights_checksum"])) < 1e-6 * cs, "seeded weights drifted (torch RNG changed?)"
    tr = O.CpuTrainer(o, sdG, sdE, sdD)
    d = O.preprocess(o, O.synthetic_batch(o, 2, seed=5))
    torch.manual_seed(0)
    g_l, _ = tr.generator_step(d)
    d_l = tr.discriminator_step(d)
    for k, v in {**g_l, **d_l}.items():
        ref = float(g["loss_" + k])
        assert abs(float(v.detach().mean()) - ref) < 2e-5 * max(1.0, abs(ref)), (k, float(v.detach().mean()), ref)
    probes = [k for k in g.files if k.startswith("param_")]
    assert len(probes) >= 6
    for k in probes:
        net, key = k[len("param_"):].split(".", 1)
        sd = {"G": tr.sdG, "E": tr.sdE, "D": tr.sdD}[net]
        got = sd[key].detach().flatten()[:64].numpy()
        assert np.abs(got - g[k]).max() < 2e-6, k


def test_param_free_bn_known_answers_of_the_reference_sync_bn_tests():
    """The only numerics the reference's own tests pin on this path are the vendored Sync-BN unit tests
    (Synchronized-BatchNorm-PyTorch/tests/test_numeric_batchnorm.py:30-52, test_sync_batchnorm.py:79-107):
    batch norm in the sum / sum-of-squares formulation equals nn.BatchNorm - output, input gradient,
    running mean and UNBIASED running variance - on rand(16, 10[, 16, 16]).  Same check for the oracle's
    param_free_bn (affine=False, eps 1e-5, momentum 0.1), with torch.allclose defaults like the
    reference's assertTensorClose."""
    torch.manual_seed(0)
    x = torch.rand(16, 10, 16, 16)
    sd = {"bn.running_mean": torch.zeros(10), "bn.running_var": torch.ones(10),
          "bn.num_batches_tracked": torch.zeros((), dtype=torch.long)}
    a = x.clone().requires_grad_(True)
    y = O.param_free_bn(a, sd, "bn.", True)
    y.sum().backward()

    b = x.clone().requires_grad_(True)
    n = b.numel() // b.shape[1]
    s1 = b.sum(dim=(0, 2, 3))
    s2 = (b * b).sum(dim=(0, 2, 3))
    mean = s1 / n
    sumvar = s2 - s1 * mean                       # handy_var of the reference tests
    std = torch.sqrt((sumvar / n).clamp(min=1e-5) + 0)  # bias var; eps enters as in batchnorm.py:87-93
    y2 = (b - mean.view(1, -1, 1, 1)) / torch.sqrt(sumvar.view(1, -1, 1, 1) / n + 1e-5)
    y2.sum().backward()
    assert torch.allclose(y, y2, atol=1e-5) and torch.allclose(a.grad, b.grad, atol=1e-5)
    assert torch.allclose(sd["bn.running_mean"], 0.1 * mean.detach())
    assert torch.allclose(sd["bn.running_var"], 0.9 * torch.ones(10) + 0.1 * (sumvar / (n - 1)).detach())
    assert int(sd["bn.num_batches_tracked"]) == 1 and std.min() > 0
    # eval mode uses the running statistics
    z = O.param_free_bn(x, sd, "bn.", False)
    ref = (x - sd["bn.running_mean"].view(1, -1, 1, 1)) / torch.sqrt(sd["bn.running_var"].view(1, -1, 1, 1) + 1e-5)
    assert torch.allclose(z, ref, atol=1e-6)
