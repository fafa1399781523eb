"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, CPU, fp32).

Run in the build container only (``python oracle/make_golden.py``); the GPU box has no
/root/reference, so tests replay the committed files.  For every case the script
  1. builds the reference module from an ObjectDict opt (options/demo_options.json +
     get_opt_config + train defaults),
  2. loads the oracle's seeded state_dict with strict key matching (proves the checkpoint layout),
  3. runs the reference forward (and backward where stated) on seeded inputs,
  4. stores inputs that are small, the reference outputs, and a checksum of the weights,
  5. asserts the oracle restatement reproduces the reference output (so a broken oracle never
     gets committed together with fresh goldens).
Weights are regenerated from seeds at test time (torch CPU RNG, same image on both boxes); the
checksum catches RNG drift.
"""
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("DEEPSEE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from oracle import deepsee_oracle as O  # noqa: E402

import warnings  # noqa: E402
warnings.filterwarnings("ignore")

from util.util import ObjectDict  # noqa: E402  (reference)
from options.configurations import get_opt_config  # noqa: E402  (reference)
import deepsee_models.networks as networks  # noqa: E402  (reference)
from deepsee_models.networks import normalization as ref_norm  # noqa: E402
from deepsee_models.sr_model import SRModel  # noqa: E402
from data.preprocessor import Preprocessor  # noqa: E402

torch.autograd.set_detect_anomaly(False)  # the reference turns it on at import (normalization.py:70)
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # reduced-width (ngf=8 -> 128 channels) versions of the BASELINE.json configurations
    "g8x_eval": dict(name="8x_independent_256x256", over=dict(ngf=8, start_size=8, crop_size=64,
                                                            load_size=64, max_fm_size=256), batch=2),
    "g32x_eval": dict(name="32x_independent_512x512", over=dict(ngf=8, start_size=4, crop_size=128,
                                                               load_size=512, max_fm_size=64), batch=1),
    "g8x_train": dict(name="8x_independent_256x256", over=dict(ngf=8, start_size=8, crop_size=64,
                                                             load_size=64, max_fm_size=256), batch=2),
}


def ref_opt(o):
    d = json.load(open(os.path.join(REF, "options", "demo_options.json")))
    opt = ObjectDict(d)
    if o.name != "custom":
        opt = get_opt_config(opt, o.name)
    for k, v in o.items():
        setattr(opt, k, v)
    return opt


def checksum(sd):
    s = 0.0
    for v in sd.values():
        s += float(v.double().abs().sum())
    return s


def t2n(t):
    return t.detach().cpu().numpy()


def build_ref_generator(o, sd):
    G = networks.define_SR(ref_opt(o))
    missing, unexpected = G.load_state_dict(sd, strict=True)
    return G


def gen_case(tag, spec, train=False):
    o = O.make_opt(spec["name"], is_train=train, **spec["over"])
    sdG = O.make_generator_state(o, seed=0)
    G = build_ref_generator(o, {k: v.clone() for k, v in sdG.items()})
    data = O.preprocess(o, O.synthetic_batch(o, spec["batch"], seed=1234))
    g = torch.Generator().manual_seed(99)
    z = torch.rand(spec["batch"], o.label_nc, o.regional_style_size, generator=g) * 2 - 1
    out = {"x_lr": t2n(data["image_lr"]),
           "labels": t2n(data["input_semantics"].argmax(1)).astype(np.uint8), "z": t2n(z),
           "weights_checksum": np.float64(checksum(sdG))}
    if not train:
        G.eval()
        with torch.no_grad():
            ref = G(data["image_lr"], seg=data["input_semantics"], z=z)
        sd2 = {k: v.clone() for k, v in sdG.items()}
        with torch.no_grad():
            mine = O.generator_forward(sd2, o, data["image_lr"], data["input_semantics"], z, False)
        err = (ref - mine).abs().max().item()
        print(tag, "eval: reference vs oracle max-abs", err, "out std", ref.std().item())
        assert err < 1e-5, err
        out["fake"] = t2n(ref)
    else:
        # training-mode forward: batch statistics, one spectral power iteration, noise injection.
        # The noise tensors are recorded from the reference run and replayed into the oracle.
        G.train()
        noises = []
        orig = ref_norm.NoiseInjection.forward

        def rec(self, tensor, noise=None):
            n = torch.randn(tensor.shape)
            noises.append(n)
            return orig(self, tensor, n)

        ref_norm.NoiseInjection.forward = rec
        torch.manual_seed(7)
        ref = G(data["image_lr"], seg=data["input_semantics"], z=z)
        ref_norm.NoiseInjection.forward = orig
        loss = (ref * torch.linspace(-1, 1, ref.numel()).view_as(ref)).sum()
        loss.backward()
        sd2 = {k: v.clone() for k, v in sdG.items()}
        for k, v in sd2.items():
            if v.is_floating_point() and not O._is_buffer(k):
                v.requires_grad_(True)
        it = iter(noises)
        mine = O.generator_forward(sd2, o, data["image_lr"], data["input_semantics"], z, True,
                                   lambda name, shape: next(it))
        (mine * torch.linspace(-1, 1, mine.numel()).view_as(mine)).sum().backward()
        err = (ref - mine).abs().max().item()
        print(tag, "train: reference vs oracle max-abs", err, "noises", len(noises))
        assert err < 2e-5, err
        ref_sd = G.state_dict()
        gk = "G_middle_0.norm_0.param_free_norm.running_var"
        assert (ref_sd[gk] - sd2[gk]).abs().max().item() < 1e-5
        gname = "head_0.conv_0.weight_orig"
        gref = dict(G.named_parameters())[gname].grad
        gerr = (gref - sd2[gname].grad).abs().max().item() / gref.abs().max().item()
        print(tag, "train: grad rel err", gerr)
        assert gerr < 1e-3, gerr
        out["fake"] = t2n(ref)
        out["noise_seed_note"] = np.array(0)
        out["running_var_G_middle_0_norm_0"] = t2n(ref_sd[gk])
        out["running_mean_up_list_0_norm_1"] = t2n(
            ref_sd["up_list.0.norm_1.param_free_norm.running_mean"])
        out["grad_head_0_conv_0_weight_orig_sample"] = t2n(gref.flatten()[::997])
        out["grad_conv_img_weight"] = t2n(dict(G.named_parameters())["conv_img.weight"].grad)
        out["weight_u_head_0_conv_0"] = t2n(ref_sd["head_0.conv_0.weight_u"])
        # store the noises: 15 tensors of [2,128,h,w] are too large; store the seed protocol instead
        out["noise_protocol"] = np.array("torch.manual_seed(7); torch.randn(shape) per NoiseInjection call, in call order")
    np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **out)


def enc_disc_case():
    """Style encoders, discriminator and losses through the reference SRModel (generator mode)."""
    o = O.make_opt("8x_independent_256x256", is_train=True, ngf=8, start_size=8, crop_size=64,
                   load_size=64, add_noise=False)
    ro = ref_opt(o)
    sdG, sdE, sdD = O.make_generator_state(o, 0), O.make_encoder_state(o, 1), \
        O.make_discriminator_state(o, 2)
    model = SRModel(ro)
    model.netSR.load_state_dict({k: v.clone() for k, v in sdG.items()}, strict=True)
    model.netE.load_state_dict({k: v.clone() for k, v in sdE.items()}, strict=True)
    model.netD.load_state_dict({k: v.clone() for k, v in sdD.items()}, strict=True)
    data = O.preprocess(o, O.synthetic_batch(o, 2, seed=4321))
    out = {"weights_checksum": np.float64(checksum(sdG) + checksum(sdE) + checksum(sdD))}

    # encoders, eval mode, both branches
    model.eval()
    with torch.no_grad():
        z_mini, _ = model.netE(data["image_lr"], data["input_semantics"], mode="mini", no_noise=True)
        z_full, _ = model.netE(data["image_hr"], data["input_semantics"], mode="full", no_noise=True)
        o_mini = O.encoder_forward(sdE, o, data["image_lr"], data["input_semantics"], "mini")
        o_full = O.encoder_forward(sdE, o, data["image_hr"], data["input_semantics"], "full")
    print("encoder mini/full max-abs", (z_mini - o_mini).abs().max().item(),
          (z_full - o_full).abs().max().item())
    assert (z_mini - o_mini).abs().max().item() < 1e-5 and (z_full - o_full).abs().max().item() < 1e-5
    out["z_mini"], out["z_full"] = t2n(z_mini), t2n(z_full)

    # discriminator (train mode => power iteration) on [fake | real]
    g = torch.Generator().manual_seed(5)
    fake = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    model.train()
    sdD2 = {k: v.clone() for k, v in sdD.items()}
    pf, pr = model.discriminate(data["input_semantics"], fake, data["image_hr"])
    of, orr = O.discriminate(sdD2, o, data["input_semantics"], fake, data["image_hr"], True)
    worst = 0.0
    for i in range(2):
        for j in range(5):
            worst = max(worst, (pf[i][j] - of[i][j]).abs().max().item(),
                        (pr[i][j] - orr[i][j]).abs().max().item())
    print("discriminator features max-abs", worst)
    assert worst < 2e-5
    out["d_fake_pred0"], out["d_real_pred1"] = t2n(pf[0][-1]), t2n(pr[1][-1])
    out["d_feat_0_2_mean"] = np.array([float(pf[0][2].mean()), float(pf[0][2].abs().mean())])
    out["fake_for_d"] = t2n(fake)

    # losses in generator / discriminator mode with the style coin flips seeded
    model.netD.load_state_dict({k: v.clone() for k, v in sdD.items()}, strict=True)
    model.netSR.load_state_dict({k: v.clone() for k, v in sdG.items()}, strict=True)
    model.netE.load_state_dict({k: v.clone() for k, v in sdE.items()}, strict=True)
    random.seed(3)
    torch.manual_seed(3)
    g_losses, gen = model(dict(data), mode="generator")
    sG, sE, sD = ({k: v.clone() for k, v in sd.items()} for sd in (sdG, sdE, sdD))
    rng = random.Random(3)
    torch.manual_seed(3)
    z, mode = O.encode_style(sE, o, rng, True, data["image_lr"], data["input_semantics"],
                             data["image_hr"], None, None)
    fake_o = O.generator_forward(sG, o, data["image_lr"], data["input_semantics"], z, True)
    lo = O.generator_losses(sD, o, data["input_semantics"], fake_o, data["image_hr"])
    print("generator-mode fake max-abs", (gen - fake_o).abs().max().item(), "mode", mode,
          {k: (float(g_losses[k]), float(lo[k])) for k in lo})
    assert (gen - fake_o).abs().max().item() < 5e-5
    for k in lo:
        assert abs(float(g_losses[k]) - float(lo[k])) < 1e-4 * max(1.0, abs(float(lo[k])))
    out["gen_fake"] = t2n(gen)
    out["loss_GAN"], out["loss_GAN_Feat"] = np.float64(g_losses["GAN"]), np.float64(g_losses["GAN_Feat"])
    out["encoder_mode"] = np.array(mode)
    np.savez_compressed(os.path.join(GOLD, "enc_disc_losses.npz"), **out)


TRAINER_CASES = {
    "train_iteration": ("8x_independent_256x256", dict(ngf=8, nef=8, ndf=8, start_size=8, crop_size=64,
                                                       load_size=64)),
    # guided model (style encoder on the HR guiding image), PureSEAN tail + max_fm_size quirk
    "train_iteration_guided": ("32x_guided_512x512", dict(ngf=8, nef=8, ndf=8, start_size=4, crop_size=128,
                                                          load_size=512, max_fm_size=64)),
}


def trainer_case(tag="train_iteration"):
    """One full training iteration (G step + D step, both Adam updates) through the reference's own
    TrainerManager on CPU (managers/trainer_manager.py:32-61), against the oracle's CpuTrainer."""
    from managers.trainer_manager import TrainerManager  # reference
    name, over = TRAINER_CASES[tag]
    o = O.make_opt(name, is_train=True, add_noise=False, noisy_style_scale=0.0, **over)
    ro = ref_opt(o)
    sdG, sdE, sdD = O.make_generator_state(o, 0), O.make_encoder_state(o, 1), \
        O.make_discriminator_state(o, 2)
    clone = lambda sd: {k: v.clone() for k, v in sd.items()}
    mgr = TrainerManager(ro)
    model = mgr.sr_model_on_one_gpu
    model.netSR.load_state_dict(clone(sdG), strict=True)
    model.netE.load_state_dict(clone(sdE), strict=True)
    model.netD.load_state_dict(clone(sdD), strict=True)
    model.train()
    raw = O.synthetic_batch(o, 2, seed=5)
    batch = lambda: {k: (v.clone().float() if "label" in k else v.clone()) for k, v in raw.items()}
    random.seed(0)
    torch.manual_seed(0)
    mgr.run_generator_one_step(batch())
    mgr.run_discriminator_one_step(batch())
    ref_losses = {k: float(v.detach().mean()) for k, v in mgr.get_latest_losses().items()}

    tr = O.CpuTrainer(o, clone(sdG), clone(sdE), clone(sdD))
    d = O.preprocess(o, raw)
    torch.manual_seed(0)
    g_l, _ = tr.generator_step(d)
    d_l = tr.discriminator_step(d)
    mine = {k: float(v.detach().mean()) for k, v in {**g_l, **d_l}.items()}
    print("trainer losses reference", ref_losses, "oracle", mine)
    for k in ref_losses:
        assert abs(ref_losses[k] - mine[k]) < 2e-5 * max(1.0, abs(ref_losses[k])), k
    out = {"weights_checksum": np.float64(checksum(sdG) + checksum(sdE) + checksum(sdD))}
    for k, v in ref_losses.items():
        out["loss_" + k] = np.float64(v)
    # parameters after both Adam steps (Adam's first step moves every touched weight by ~lr, so
    # these pin the sign pattern of the gradients as well as the optimizer settings)
    probes = {"G": (model.netSR.state_dict(), tr.sdG, ["conv_img.bias", "head_0.conv_1.weight_orig",
                                                        "up_list.1.norm_1.mlp_gamma.bias",
                                                        "G_middle_0.norm_0.alpha_gamma"]),
              "E": (model.netE.state_dict(), tr.sdE, ["encoder_mini.conv0.0.0.weight_orig", "down0.0.0.weight_orig",
                                                        "final.0.0.weight_orig"]),
              "D": (model.netD.state_dict(), tr.sdD, ["discriminator_0.model0.0.bias",
                                                        "discriminator_1.model2.0.0.weight_orig"])}
    for net, (ref_sd, my_sd, keys) in probes.items():
        for k in keys:
            if k not in ref_sd:
                continue
            a, b = ref_sd[k].detach().flatten()[:64], my_sd[k].detach().flatten()[:64]
            err = (a - b).abs().max().item()
            print("  %s.%s max-abs after the iteration %.2e" % (net, k, err))
            assert err < 2e-6
            out["param_%s.%s" % (net, k)] = t2n(a)
    np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **out)
    print("training-iteration golden ok:", tag)


def label_case():
    o = O.make_opt("8x_independent_256x256")
    ro = ref_opt(o)
    pre = Preprocessor(ro)
    g = torch.Generator().manual_seed(11)
    lab = torch.randint(0, 19, (2, 1, 48, 48), generator=g)
    ref = pre.preprocess_label(lab)
    assert np.array_equal(t2n(ref), O.preprocess_label_np(lab.numpy(), 19))
    sizes = [8, 12, 24, 32, 48]
    outs = {}
    for s in sizes:
        r = torch.nn.functional.interpolate(ref, size=(s, s), mode="nearest").argmax(1)
        mine = O.resize_labels_np(lab[:, 0].numpy(), s, s)
        assert np.array_equal(t2n(r), mine), s
        outs["resized_%d" % s] = t2n(r).astype(np.uint8)
    hr = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    lr = pre.downsample_image(hr, (8, 8))
    assert torch.equal(lr, O.downsample_image(hr, 8))
    np.savez_compressed(os.path.join(GOLD, "labels.npz"), labels=lab.numpy().astype(np.uint8),
                        onehot_sum=t2n(ref.sum((0, 2, 3))), image_hr=t2n(hr), image_lr=t2n(lr), **outs)
    print("label / preprocess goldens ok")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "trainer":   # add these goldens without touching the others
        for tag in TRAINER_CASES:
            trainer_case(tag)
        sys.exit(0)
    label_case()
    gen_case("g8x_eval", CASES["g8x_eval"])
    gen_case("g32x_eval", CASES["g32x_eval"])
    gen_case("g8x_train", CASES["g8x_train"], train=True)
    enc_disc_case()
    for tag in TRAINER_CASES:
        trainer_case(tag)
    print("goldens written to", GOLD)
