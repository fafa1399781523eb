"""CPU oracle for DeepSEE's hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; the product (``deepsee_b200``) never does and has no CPU path of its own.

What it is: a functional, fp32, NCHW restatement in plain ``torch`` (CPU) of the reference's
generator / style encoder / discriminator / losses, written against a flat ``state_dict`` with the
reference's parameter names.  Each function cites the reference file:line it follows
(paths under mcbuehler/DeepSEE).  The arithmetic primitives themselves (conv2d, batch_norm,
instance_norm, interpolate, spectral normalisation) live in the reference's third-party dependency
PyTorch (requirements.txt:2, unpinned ``torch>=1.0.0``; validated here with torch 2.11.0), whose
published semantics are restated where they are not a single functional call (spectral norm:
torch/nn/utils/spectral_norm.py, one power iteration per training forward).

Pinning: the reference holds no golden vectors or tests for this path, so the
oracle is pinned against outputs of the reference itself, run in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/``; ``tests/test_oracle_golden.py``
replays them (CPU, every round).

Label-map integer work (one-hot scatter, nearest resize) is restated in numpy and compared
bit-exactly.
"""
import math
import random as _random
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LRELU = 0.2
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# =================================================================================================
# options
# =================================================================================================
class Opt(dict):
    """Attribute dict like the reference's util.ObjectDict (util/util.py:439-443)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def make_opt(name="8x_independent_256x256", is_train=False, **overrides):
    """Option set the hot path reads: options/demo_options.json defaults,
    train defaults from options/train_options.py:27-74 and the name presets of
    options/configurations.py:16-43."""
    o = Opt(
        name=name, ngf=32, nef=32, ndf=32, label_nc=19, semantic_nc=19, output_nc=3,
        contain_dontcare_label=False, regional_style_size=128, max_fm_size=256,
        norm_G="spectrallateseansyncbatch3x3", norm_D="spectralinstance", norm_E="spectralinstance",
        netG="deepsee", netE="combinedstyle", netD="multiscale", netD_subarch="n_layer", num_D=2,
        n_layers_D=4, start_size=16, crop_size=128, load_size=128, aspect_ratio=1.0,
        num_upsampling_layers="normal", add_noise=False, efficient=False, model_parallel_mode=0,
        noisy_style_scale=0.2, noisy_style_dist="uniform", random_style_matrix=False,
        full_style_image=False, guiding_style_image=False, downsampling_method="bicubic",
        init_type="xavier", init_variance=0.02, gpu_ids=[], gpu_info=False, isTrain=is_train,
        batchSize=1, checkpoints_dir="./checkpoints", which_epoch="latest", continue_train=False,
        gan_mode="hinge", lambda_feat=10.0, lambda_vgg=10.0, no_ganFeat_loss=False,
        no_vgg_loss=True, lr=0.0002, beta1=0.0, beta2=0.9, no_TTUR=False, gradient_clip=-1.0,
        niter=50, niter_decay=25, dataset="celebamaskhq",
    )
    # options/configurations.py:16-43
    if "128x128" in name and "8x_" in name:
        o.update(start_size=16, crop_size=128, load_size=128, dataset="celeba", add_noise=True)
    elif "256x256" in name and "8x_" in name:
        o.update(start_size=32, crop_size=256, load_size=256, add_noise=True, max_fm_size=256)
    elif "32x_" in name:
        o.update(start_size=16, crop_size=512, load_size=512, add_noise=False, max_fm_size=256)
    elif name != "custom":
        raise ValueError("Invalid name: %r" % name)
    if "independent" in name:
        o.update(netE="combinedstyle", noisy_style_scale=0.2)
    elif "guided" in name:
        o.update(netE="fullstyle", noisy_style_scale=0.05, guiding_style_image=True)
    o.update(overrides)
    return o


# =================================================================================================
# label maps (integer work, numpy, bit-exact)
# =================================================================================================
def preprocess_label_np(label, nc):
    """data/preprocessor.py:35-41: zeros(bs,nc,h,w).scatter_(1, label, 1.0). label int [B,1,H,W]."""
    label = np.asarray(label)
    b, _, h, w = label.shape
    out = np.zeros((b, nc, h, w), dtype=np.float32)
    bi, yi, xi = np.meshgrid(np.arange(b), np.arange(h), np.arange(w), indexing="ij")
    out[bi, label[:, 0], yi, xi] = 1.0
    return out


def nearest_index_np(n_out, n_in):
    """Source index of F.interpolate(mode='nearest') (ATen upsample_nearest: floor(dst*in/out))."""
    scale = np.float32(n_in) / np.float32(n_out)
    idx = np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, n_in - 1)


def resize_labels_np(labels, h, w):
    """normalization.py:110,174,261 applied to the integer label map [B,H,W]."""
    labels = np.asarray(labels)
    yi = nearest_index_np(h, labels.shape[1])
    xi = nearest_index_np(w, labels.shape[2])
    return labels[:, yi][:, :, xi]


def downsample_image(hr, size, method="bicubic"):
    """data/preprocessor.py:17-33."""
    return F.interpolate(hr, (size, size), mode=method).clamp(min=-1, max=1)


# =================================================================================================
# building blocks
# =================================================================================================
def _normalize(v, eps=1e-12):
    return v / max(float(v.norm()), eps)


def spectral_weight(sd, pfx, training, update_uv=True, eps=1e-12):
    """torch.nn.utils.spectral_norm as used at architecture.py:40-44 and normalization.py:29-31.
    Training: one power iteration (no grad) updating weight_u / weight_v in ``sd`` in place, then
    sigma = u . (W v), W = W_orig / sigma.  Eval: stored u, v."""
    w = sd[pfx + "weight_orig"]
    u, v = sd[pfx + "weight_u"], sd[pfx + "weight_v"]
    w_mat = w.reshape(w.shape[0], -1)
    if training and update_uv:
        with torch.no_grad():
            wm = w_mat.detach()
            v_new = _normalize(torch.mv(wm.t(), u), eps)
            u_new = _normalize(torch.mv(wm, v_new), eps)
            v.copy_(v_new)
            u.copy_(u_new)
        u, v = u.clone(), v.clone()
    sigma = torch.dot(u, torch.mv(w_mat, v))
    return w / sigma


def conv_weight(sd, pfx, training):
    """Weight of a conv that may or may not be spectral-normalised."""
    if pfx + "weight_orig" in sd:
        return spectral_weight(sd, pfx, training)
    return sd[pfx + "weight"]


def param_free_bn(x, sd, pfx, training):
    """SynchronizedBatchNorm2d(affine=False) outside DataParallel == F.batch_norm
    (sync_batchnorm/batchnorm.py:63-68); training updates running stats (momentum 0.1, unbiased)."""
    rm, rv = sd[pfx + "running_mean"], sd[pfx + "running_var"]
    out = F.batch_norm(x, rm, rv, None, None, training, BN_MOMENTUM, BN_EPS)
    if training and (pfx + "num_batches_tracked") in sd:
        sd[pfx + "num_batches_tracked"] += 1
    return out


def noise_injection(x, weight, noise):
    """normalization.py:299-304."""
    return x + weight.view(1, -1, 1, 1) * noise


def _mlp_shared(seg, sd, pfx):
    return F.relu(F.conv2d(seg, sd[pfx + "mlp_shared.0.weight"], sd[pfx + "mlp_shared.0.bias"],
                           padding=1))


def _style_map(style, seg):
    """normalization.py:182-185: sum_c style[b,c,:] * seg[b,c,y,x]."""
    return torch.einsum("bcs,bchw->bshw", style, seg)


def spade(x, seg, sd, pfx, training):
    """SPADE.forward, normalization.py:105-120."""
    normalized = param_free_bn(x, sd, pfx + "param_free_norm.", training)
    seg = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    actv = _mlp_shared(seg, sd, pfx)
    gamma = F.conv2d(actv, sd[pfx + "mlp_gamma.weight"], sd[pfx + "mlp_gamma.bias"], padding=1)
    beta = F.conv2d(actv, sd[pfx + "mlp_beta.weight"], sd[pfx + "mlp_beta.bias"], padding=1)
    return normalized * (1 + gamma) + beta


def _sean_maps(x, seg, style, sd, pfx, max_fm_size):
    out_size = tuple(x.shape[2:])
    fm_size = tuple(min(s, max_fm_size) for s in out_size)
    seg = F.interpolate(seg, size=fm_size, mode="nearest")
    actv = _mlp_shared(seg, sd, pfx)
    style_map = _style_map(style, seg)
    if out_size != fm_size:
        # normalization.py:188-190 / 275-277: BOTH maps become the upsampled actv (reference quirk:
        # the style matrix is dropped for feature maps larger than max_fm_size).
        actv = F.interpolate(actv, size=out_size)
        style_map = F.interpolate(actv, size=out_size)
    return actv, style_map


def sean_block(x, seg, style, sd, pfx, training, max_fm_size):
    """SEAN_Block.forward, normalization.py:167-213."""
    normalized = param_free_bn(x, sd, pfx + "param_free_norm.", training)
    actv, style_map = _sean_maps(x, seg, style, sd, pfx, max_fm_size)
    gamma = F.conv2d(actv, sd[pfx + "mlp_gamma.weight"], sd[pfx + "mlp_gamma.bias"], padding=1)
    beta = F.conv2d(actv, sd[pfx + "mlp_beta.weight"], sd[pfx + "mlp_beta.bias"], padding=1)
    beta_s = F.conv2d(style_map, sd[pfx + "mlp_style_beta.weight"], sd[pfx + "mlp_style_beta.bias"],
                      padding=1)
    gamma_s = F.conv2d(style_map, sd[pfx + "mlp_style_gamma.weight"],
                       sd[pfx + "mlp_style_gamma.bias"], padding=1)
    a_b = torch.sigmoid(sd[pfx + "alpha_beta"])
    a_g = torch.sigmoid(sd[pfx + "alpha_gamma"])
    offset = a_b * beta_s + (1.0 - a_b) * beta
    scale = a_g * gamma_s + (1.0 - a_g) * gamma + 1
    return normalized * scale + offset


def puresean_block(x, seg, style, sd, pfx, training, max_fm_size):
    """PureSEAN_Block.forward, normalization.py:254-286 (no '+1' on gamma)."""
    normalized = param_free_bn(x, sd, pfx + "param_free_norm.", training)
    _, style_map = _sean_maps(x, seg, style, sd, pfx, max_fm_size)
    beta_s = F.conv2d(style_map, sd[pfx + "mlp_style_beta.weight"], sd[pfx + "mlp_style_beta.bias"],
                      padding=1)
    gamma_s = F.conv2d(style_map, sd[pfx + "mlp_style_gamma.weight"],
                       sd[pfx + "mlp_style_gamma.bias"], padding=1)
    return normalized * gamma_s + beta_s


def _norm(kind, x, seg, style, sd, pfx, training, opt):
    if kind == "spade":
        return spade(x, seg, sd, pfx, training)
    if kind == "sean":
        return sean_block(x, seg, style, sd, pfx, training, opt.max_fm_size)
    return puresean_block(x, seg, style, sd, pfx, training, opt.max_fm_size)


def _actvn(t, key, act_fn):
    """architecture.py:146-147 (LeakyReLU 0.2).  ``act_fn(key, t)`` lets a test pin the activation
    pattern (which side of the kink each element is on) to the one the implementation under test
    took, so gradients can be compared without measure-zero sign flips dominating the error."""
    return F.leaky_relu(t, LRELU) if act_fn is None else act_fn(key, t)


def resnet_block(x, seg, style, sd, pfx, kind, opt, training, noise_fn=None, taps=None, act_fn=None):
    """SPADEResnetBlock.forward / shortcut / actvn, architecture.py:75-147 (fin == fout: identity
    shortcut; learned shortcut never built by DeepSEESR)."""
    add_noise = opt.add_noise and training
    if add_noise:
        x = noise_injection(x, sd[pfx + "noise_in.weight"], noise_fn(pfx + "noise_in", x.shape))
        x_s = noise_injection(x, sd[pfx + "noise_skip.weight"], noise_fn(pfx + "noise_skip", x.shape))
    else:
        x_s = x
    h = _actvn(_norm(kind, x, seg, style, sd, pfx + "norm_0.", training, opt), pfx + "act_0", act_fn)
    if taps is not None:
        taps[pfx + "act_0"] = h
    dx = F.conv2d(h, conv_weight(sd, pfx + "conv_0.", training), sd[pfx + "conv_0.bias"], padding=1)
    if add_noise:
        dx = noise_injection(dx, sd[pfx + "noise_middle.weight"],
                             noise_fn(pfx + "noise_middle", dx.shape))
    h = _actvn(_norm(kind, dx, seg, style, sd, pfx + "norm_1.", training, opt), pfx + "act_1", act_fn)
    if taps is not None:
        taps[pfx + "act_1"] = h
    dx = F.conv2d(h, conv_weight(sd, pfx + "conv_1.", training), sd[pfx + "conv_1.bias"], padding=1)
    return x_s + dx


def generator_layout(opt):
    """Block list of DeepSEESR.__init__ (sr.py:21-57): [(prefix, kind, upsample_before)]."""
    n_blocks = int(np.log2(opt.crop_size) - np.log2(opt.start_size))
    early_style = "late" not in opt.norm_G
    has_sean = "sean" in opt.norm_G.replace("spectral", "")
    style_kind = "sean" if has_sean else "spade"
    blocks = [("head_0.", style_kind if early_style else "spade", False),
              ("G_middle_0.", style_kind, True), ("G_middle_1.", style_kind, False)]
    max_n_blocks = 4 if opt.load_size >= 512 else 99
    n_plain = len(range(1, min(n_blocks, max_n_blocks)))
    idx = 0
    for _ in range(n_plain):
        blocks.append(("up_list.%d." % idx, style_kind, True))
        idx += 1
    if max_n_blocks != 99:
        for _ in range(max_n_blocks, n_blocks):
            blocks.append(("up_list.%d." % idx, "puresean", True))
            idx += 1
    # forward uses up_list[0 .. n_blocks-2] (sr.py:71)
    used = 3 + (n_blocks - 1)
    return blocks[:used] if len(blocks) >= used else blocks


def generator_forward(sd, opt, x_lr, seg, z, training=False, noise_fn=None, taps=None, act_fn=None):
    """DeepSEESR.forward, sr.py:62-98."""
    x = F.conv2d(x_lr, sd["initial.weight"], sd["initial.bias"], padding=1)
    for pfx, kind, up in generator_layout(opt):
        if up:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = resnet_block(x, seg, z, sd, pfx, kind, opt, training, noise_fn, taps, act_fn)
        if taps is not None:
            taps[pfx + "out"] = x
    x = F.conv2d(_actvn(x, "head", act_fn), sd["conv_img.weight"], sd["conv_img.bias"], padding=1)
    return torch.tanh(x)


# ---- style encoder ------------------------------------------------------------------------------
def _enc_stage(x, sd, pfx, training, stride=1, upsample=False):
    """spectral conv3x3 (no bias) + InstanceNorm2d(affine=False) + LeakyReLU(0.2)
    (encoder.py:84-98,142-157; normalization.py:19-54)."""
    if upsample:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(x, conv_weight(sd, pfx, training), None, stride=stride, padding=1)
    return F.leaky_relu(F.instance_norm(x, eps=1e-5), LRELU)


def extract_style_matrix(x, seg):
    """encoder.py:36-49: mean over H*W of x masked by each region (divides by H*W, not area)."""
    if seg.shape[2:] != x.shape[2:]:
        seg = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    return torch.einsum("bchw,blhw->blc", x, seg) / float(x.shape[2] * x.shape[3])


def corrupt_style_matrix(style, noise_weights, max_range, unit_noise):
    """encoder.py:51-70 with the uniform noise supplied by the caller (unit_noise in [0,1))."""
    nw = torch.sigmoid(noise_weights).view(1, -1, 1)
    noise = (unit_noise * 2 - 1) * max_range
    return (style + noise * nw).clamp(-1, 1)


def encoder_forward(sd, opt, x, seg, mode, no_noise=True, unit_noise=None, training=False):
    """CombinedstyleEncoder.forward (encoder.py:195-210) / FullStyleEncoder.forward (:116-132)."""
    combined = opt.netE == "combinedstyle"
    if combined:
        br = "encoder_full." if mode == "full" else "encoder_mini."
    else:
        br = ""
        mode = "full"
    if mode == "full":
        x = _enc_stage(x, sd, br + "initial.0.0.", training)
        x = _enc_stage(x, sd, br + "down0.0.0.", training, stride=2)
        x = _enc_stage(x, sd, br + "down1.0.0.", training, stride=2)
        x = _enc_stage(x, sd, br + "up_conv.1.0.", training, upsample=True)
    else:
        x = _enc_stage(x, sd, br + "initial.0.0.", training)
        x = _enc_stage(x, sd, br + "conv0.0.0.", training)
        x = _enc_stage(x, sd, br + "conv1.0.0.", training)
        x = _enc_stage(x, sd, br + "conv2.1.0.", training, upsample=True)
    x = F.conv2d(x, conv_weight(sd, "final.0.0.", training), None, padding=1)
    x = torch.tanh(F.instance_norm(x, eps=1e-5))
    style = extract_style_matrix(x, seg)
    noisy = opt.noisy_style_scale > 0
    if noisy and not no_noise:
        style = corrupt_style_matrix(style, sd["noise_weights"], opt.noisy_style_scale, unit_noise)
    return style


# ---- discriminator ------------------------------------------------------------------------------
def nlayer_discriminator(sd, pfx, x, opt, training):
    """NLayerDiscriminator.forward, discriminator.py:78-120; returns the 5 intermediate outputs."""
    outs = []
    x = F.leaky_relu(F.conv2d(x, sd[pfx + "model0.0.weight"], sd[pfx + "model0.0.bias"], stride=2,
                              padding=2), LRELU)
    outs.append(x)
    for n in range(1, opt.n_layers_D):
        stride = 1 if n == opt.n_layers_D - 1 else 2
        w = conv_weight(sd, pfx + "model%d.0.0." % n, training)
        x = F.conv2d(x, w, None, stride=stride, padding=2)
        x = F.leaky_relu(F.instance_norm(x, eps=1e-5), LRELU)
        outs.append(x)
    n = opt.n_layers_D
    x = F.conv2d(x, sd[pfx + "model%d.0.weight" % n], sd[pfx + "model%d.0.bias" % n], stride=1,
                 padding=2)
    outs.append(x)
    return outs


def discriminator_forward(sd, opt, x, training=True):
    """MultiscaleDiscriminator.forward, discriminator.py:46-63."""
    result = []
    for i in range(opt.num_D):
        result.append(nlayer_discriminator(sd, "discriminator_%d." % i, x, opt, training))
        x = F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)
    return result


def divide_pred(pred):
    """sr_model.py:671-683."""
    fake = [[t[: t.size(0) // 2] for t in p] for p in pred]
    real = [[t[t.size(0) // 2:] for t in p] for p in pred]
    return fake, real


def discriminate(sdD, opt, seg, fake, real, training=True):
    """sr_model.py:655-668."""
    fake_and_real = torch.cat([torch.cat([seg, fake], 1), torch.cat([seg, real], 1)], 0)
    return divide_pred(discriminator_forward(sdD, opt, fake_and_real, training))


# ---- losses -------------------------------------------------------------------------------------
def gan_loss(preds, target_is_real, for_discriminator, mode="hinge"):
    """GANLoss.__call__/loss, loss.py:60-101 (list-of-lists input; mean over scales)."""
    total = 0
    for p in preds:
        x = p[-1]
        if mode == "hinge":
            if for_discriminator:
                l = -torch.mean(torch.min((x if target_is_real else -x) - 1, torch.zeros_like(x)))
            else:
                l = -torch.mean(x)
        elif mode == "ls":
            l = F.mse_loss(x, torch.full_like(x, 1.0 if target_is_real else 0.0))
        elif mode == "original":
            l = F.binary_cross_entropy_with_logits(x, torch.full_like(x, 1.0 if target_is_real else 0.0))
        else:
            l = -x.mean() if target_is_real else x.mean()
        total = total + l.reshape(1, -1).mean(dim=1)
    return total / len(preds)


def generator_losses(sdD, opt, seg, fake, real):
    """SRModel.compute_generator_loss, sr_model.py:518-545 (VGG term excluded)."""
    pred_fake, pred_real = discriminate(sdD, opt, seg, fake, real)
    losses = OrderedDict()
    losses["GAN"] = gan_loss(pred_fake, True, False, opt.gan_mode)
    if not opt.no_ganFeat_loss:
        feat = torch.zeros(1)
        for i in range(len(pred_fake)):
            for j in range(len(pred_fake[i]) - 1):
                feat = feat + F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * \
                    opt.lambda_feat / len(pred_fake)
        losses["GAN_Feat"] = feat
    return losses


def discriminator_losses(sdD, opt, seg, fake, real):
    """SRModel.compute_discriminator_loss, sr_model.py:547-564."""
    pred_fake, pred_real = discriminate(sdD, opt, seg, fake.detach(), real)
    return OrderedDict(D_Fake=gan_loss(pred_fake, False, True, opt.gan_mode),
                       D_Real=gan_loss(pred_real, True, True, opt.gan_mode))


# ---- model facade -------------------------------------------------------------------------------
def encode_style(sdE, opt, rng, training, image_lr, seg, image_hr, guiding_image, guiding_label,
                 no_noise=None, encode_full=False, unit_noise=None, unit_noise_fn=None):
    """SRModel.get_encoder_inputs / encode_style, sr_model.py:582-650.  ``rng`` is a
    ``random.Random`` standing in for the module-level ``random`` the reference draws from."""
    variant = "guided" if "full" in opt.netE else "independent"
    style_sem, style_img = seg, image_lr
    if variant == "guided":
        mode = "full"
        if opt.guiding_style_image:
            style_sem, style_img = guiding_label, guiding_image
        else:
            style_img = image_hr
    else:
        if encode_full or (training and rng.random() < 0.5):
            mode = "full"
            if opt.guiding_style_image:
                style_sem, style_img = guiding_label, guiding_image
            else:
                style_img = image_hr
        else:
            mode = "mini"
        if not no_noise:
            no_noise = rng.random() < 0.5
    if unit_noise is None and not no_noise and opt.noisy_style_scale > 0:
        shape = (style_img.shape[0], opt.label_nc, opt.regional_style_size)
        unit_noise = unit_noise_fn(shape) if unit_noise_fn is not None else torch.rand(shape)
    return encode_style_run(sdE, opt, style_img, style_sem, mode, no_noise, unit_noise, training), mode


def encode_style_run(sdE, opt, style_img, style_sem, mode, no_noise, unit_noise, training):
    return encoder_forward(sdE, opt, style_img, style_sem, mode, no_noise=bool(no_noise),
                           unit_noise=unit_noise, training=training)


def inference(sdG, sdE, opt, image_lr, seg, image_hr=None):
    """SRModel.forward(mode='inference'), sr_model.py:82-89 (eval mode: no noise, 'mini' encoder for
    the independent model unless full_style_image)."""
    rng = _random.Random(0)
    with torch.no_grad():
        z, _ = encode_style(sdE, opt, rng, False, image_lr, seg, image_hr, None, None,
                            no_noise=True, encode_full=opt.full_style_image)
        return generator_forward(sdG, opt, image_lr, seg, z, training=False), z


def sweep_interpolation(sdG, sdE, opt, image_lr, seg, image_hr, n, delta, region_idx=None):
    """SRModel.forward(mode='inference_interpolation'), sr_model.py:219-261: per sample, n renderings
    with the style of `region_idx` shifted by -delta ... +delta, one batch-1 generator call each, laid
    out side by side.  -> (fake [B,3,S,n*S], applied styles [B][n,L,d])."""
    rng = _random.Random(0)
    with torch.no_grad():
        style, _ = encode_style(sdE, opt, rng, False, image_lr, seg, image_hr, None, None, no_noise=True)
        ridx = list(region_idx) if region_idx else list(range(seg.size(1)))
        rows, applied = [], []
        for b in range(seg.size(0)):
            samples, styles = [], []
            for step in np.linspace(-delta, delta, num=n):
                z = style[b].clone()
                z[ridx] = (z[ridx] + step).clamp(-1, 1)
                samples.append(generator_forward(sdG, opt, image_lr[b:b + 1], seg[b:b + 1], z[None], training=False))
                styles.append(z)
            rows.append(torch.cat(samples, -1))
            applied.append(torch.stack(styles))
        return torch.cat(rows, 0), applied


def sweep_reference(sdG, sdE, opt, image_lr, seg, image_hr, region_idx=None):
    """SRModel.forward(mode='inference_reference'), sr_model.py:381-410: the HR image's own style
    (encoder_full) with the `region_idx` rows replaced by every sample's in turn."""
    rng = _random.Random(0)
    with torch.no_grad():
        full, _ = encode_style(sdE, opt, rng, False, None, seg, image_hr, None, None, no_noise=True,
                               encode_full=True)
        ridx = list(region_idx) if region_idx else list(range(seg.size(1)))
        rows = []
        for b in range(seg.size(0)):
            samples = []
            for other in range(seg.size(0)):
                z = full[b].clone()
                z[ridx] = full[other, ridx].clamp(-1, 1)
                samples.append(generator_forward(sdG, opt, image_lr[b:b + 1], seg[b:b + 1], z[None], training=False))
            rows.append(torch.cat(samples, -1))
        return torch.cat(rows, 0)


# =================================================================================================
# VGG19 perceptual loss (loss.py:104-119, architecture.py:151-181)
# =================================================================================================
VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512]
VGG19_SLICE_ENDS = (2, 7, 12, 21, 30)   # architecture.py:160-169: features[0:2], [2:7], [7:12], [12:21], [21:30]


def make_vgg19_state(seed=3):
    """Seeded stand-in for torchvision's pretrained vgg19 `features` (He-scaled so activations keep
    O(1) magnitude through 13 layers); keys `features.N.weight / bias` like torchvision's."""
    g = torch.Generator().manual_seed(seed)
    sd, cin, i = OrderedDict(), 3, 0
    for v in VGG19_CFG:
        if v == "M":
            i += 1
            continue
        sd["features.%d.weight" % i] = torch.randn(v, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9))
        sd["features.%d.bias" % i] = torch.randn(v, generator=g) * 0.05
        cin = v
        i += 2
    return sd


def vgg19_features(sd, x):
    """VGG19.forward (architecture.py:174-181): [h_relu1_1, h_relu2_1, h_relu3_1, h_relu4_1, h_relu5_1]."""
    outs, i = [], 0
    for v in VGG19_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            i += 1
        else:
            x = F.relu(F.conv2d(x, sd["features.%d.weight" % i], sd["features.%d.bias" % i], padding=1))
            i += 2
        if i in VGG19_SLICE_ENDS:
            outs.append(x)
    return outs


def vgg_loss(sd, x, y):
    """VGGLoss.forward (loss.py:114-119)."""
    weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]
    fx, fy = vgg19_features(sd, x), vgg19_features(sd, y)
    return sum(w * F.l1_loss(a, b.detach()) for w, a, b in zip(weights, fx, fy))


# =================================================================================================
# deterministic "conditioned" weights
# =================================================================================================
def _spectral_conv(sd, pfx, cout, cin, k, g, gain, bias=True):
    w = torch.randn(cout, cin, k, k, generator=g) * (gain / math.sqrt(cin * k * k))
    u = _normalize(torch.randn(cout, generator=g))
    wm = w.reshape(cout, -1)
    for _ in range(30):  # converged power iteration, like a trained checkpoint
        v = _normalize(torch.mv(wm.t(), u))
        u = _normalize(torch.mv(wm, v))
    sd[pfx + "weight_orig"], sd[pfx + "weight_u"], sd[pfx + "weight_v"] = w, u, v
    if bias:
        sd[pfx + "bias"] = torch.randn(cout, generator=g) * 0.05


def _plain_conv(sd, pfx, cout, cin, k, g, gain):
    sd[pfx + "weight"] = torch.randn(cout, cin, k, k, generator=g) * (gain / math.sqrt(cin * k * k))
    sd[pfx + "bias"] = torch.randn(cout, generator=g) * 0.05


def make_generator_state(opt, seed=0):
    """Seeded state_dict with the reference's keys whose scales resemble a
    trained model: converged spectral-norm u/v, non-trivial BN running stats, non-zero noise
    weights and alphas, O(1) activations and O(0.5) pre-tanh values.  (A freshly initialised
    reference generator saturates to +-1, SURVEY.md section 7 hard part 1.)"""
    g = torch.Generator().manual_seed(seed)
    C = 16 * opt.ngf
    nh, L, d = 128, opt.semantic_nc, opt.regional_style_size
    sd = OrderedDict()
    _plain_conv(sd, "initial.", C, 3, 3, g, 1.5)
    for pfx, kind, _ in _all_blocks(opt):
        for cv in ("conv_0.", "conv_1."):
            # spectral normalisation rescales the weight to unit spectral norm anyway
            _spectral_conv(sd, pfx + cv, C, C, 3, g, 1.0)
        for nm in ("norm_0.", "norm_1."):
            p = pfx + nm
            if kind in ("sean",):
                sd[p + "alpha_beta"] = torch.randn(1, generator=g)
                sd[p + "alpha_gamma"] = torch.randn(1, generator=g)
            sd[p + "param_free_norm.running_mean"] = torch.randn(C, generator=g) * 0.2
            sd[p + "param_free_norm.running_var"] = torch.rand(C, generator=g) * 1.0 + 0.5
            sd[p + "param_free_norm.num_batches_tracked"] = torch.tensor(100, dtype=torch.long)
            _plain_conv(sd, p + "mlp_shared.0.", nh, L, 3, g, 1.5)
            if kind in ("spade", "sean"):
                _plain_conv(sd, p + "mlp_gamma.", C, nh, 3, g, 0.7)
                _plain_conv(sd, p + "mlp_beta.", C, nh, 3, g, 0.7)
            if kind in ("sean", "puresean"):
                sd[p + "style_conv.weight"] = torch.randn(19, 19, 1, generator=g) * 0.1
                sd[p + "style_conv.bias"] = torch.zeros(19)
                gain = 1.5 if kind == "puresean" else 1.0
                _plain_conv(sd, p + "mlp_style_gamma.", C, d, 3, g, gain)
                _plain_conv(sd, p + "mlp_style_beta.", C, d, 3, g, gain)
                if kind == "puresean":
                    sd[p + "mlp_style_gamma.bias"] = sd[p + "mlp_style_gamma.bias"] + 1.0
        if opt.add_noise:
            for nm in ("noise_in.", "noise_skip.", "noise_middle."):
                sd[pfx + nm + "weight"] = torch.randn(C, generator=g) * 0.1
    _plain_conv(sd, "conv_img.", 3, C, 3, g, 0.5)
    sd["conv_img.weight"] = sd["conv_img.weight"] * 1.35
    return _reorder_like_reference(sd, opt)


def _all_blocks(opt):
    """All blocks DeepSEESR.__init__ constructs (sr.py:33-52), including unused trailing ones."""
    n_blocks = int(np.log2(opt.crop_size) - np.log2(opt.start_size))
    lay = generator_layout(opt)
    max_n_blocks = 4 if opt.load_size >= 512 else 99
    n_up = len(range(1, min(n_blocks, max_n_blocks)))
    if max_n_blocks != 99:
        n_up += len(range(max_n_blocks, n_blocks))
    names = [b[0] for b in lay]
    has_sean = "sean" in opt.norm_G.replace("spectral", "")
    out = list(lay)
    for i in range(n_up):
        nm = "up_list.%d." % i
        if nm not in names:
            kind = "puresean" if (max_n_blocks != 99 and i >= n_up - len(range(max_n_blocks, n_blocks))) \
                else ("sean" if has_sean else "spade")
            out.append((nm, kind, True))
    return out


def _reorder_like_reference(sd, opt):
    return sd  # key order is irrelevant to load_state_dict; kept as a hook


def make_encoder_state(opt, seed=1):
    g = torch.Generator().manual_seed(seed)
    nf, d = opt.nef, opt.regional_style_size
    sd = OrderedDict()

    def branch(pfx, names):
        chans = [(3, nf), (nf, 2 * nf), (2 * nf, 4 * nf), (4 * nf, 8 * nf)]
        for nm, (ci, co) in zip(names, chans):
            _spectral_conv(sd, pfx + nm, co, ci, 3, g, 1.0, bias=False)
        _spectral_conv(sd, pfx + "final.0.0.", d, 8 * nf, 3, g, 1.0, bias=False)

    full = ["initial.0.0.", "down0.0.0.", "down1.0.0.", "up_conv.1.0."]
    mini = ["initial.0.0.", "conv0.0.0.", "conv1.0.0.", "conv2.1.0."]
    if opt.netE == "combinedstyle":
        if opt.noisy_style_scale > 0:
            sd["noise_weights"] = torch.randn(opt.label_nc, generator=g) * 0.5
        _spectral_conv(sd, "final.0.0.", d, 8 * nf, 3, g, 1.0, bias=False)
        branch("encoder_full.", full)
        branch("encoder_mini.", mini)
    else:
        if opt.noisy_style_scale > 0:
            sd["noise_weights"] = torch.randn(opt.label_nc, generator=g) * 0.5
        branch("", full)
        # key order of FullStyleEncoder: final first
    return sd


def make_discriminator_state(opt, seed=2):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    nc_in = opt.label_nc + opt.output_nc + (1 if opt.contain_dontcare_label else 0)
    for i in range(opt.num_D):
        p = "discriminator_%d." % i
        nf = opt.ndf
        _plain_conv(sd, p + "model0.0.", nf, nc_in, 4, g, 1.4)
        for n in range(1, opt.n_layers_D):
            nf_prev, nf = nf, min(nf * 2, 512)
            _spectral_conv(sd, p + "model%d.0.0." % n, nf, nf_prev, 4, g, 1.0, bias=False)
        _plain_conv(sd, p + "model%d.0." % opt.n_layers_D, 1, nf, 4, g, 1.0)
    return sd


# =================================================================================================
# synthetic inputs
# =================================================================================================
def synthetic_batch(opt, batch, seed=1234, blocky=True, guided=None):
    g = torch.Generator().manual_seed(seed)
    S, L = opt.crop_size, opt.label_nc
    hr = torch.rand(batch, 3, S, S, generator=g) * 2 - 1

    def labels():
        if blocky:
            grid = torch.randint(0, L, (batch, 1, 16, 16), generator=g)
            return F.interpolate(grid.float(), size=(S, S), mode="nearest").long()
        return torch.randint(0, L, (batch, 1, S, S), generator=g)

    data = {"label": labels(), "image": hr}
    if guided if guided is not None else opt.guiding_style_image:
        data["guiding_image"] = torch.rand(batch, 3, S, S, generator=g) * 2 - 1
        data["guiding_label"] = labels()
    return data


def preprocess(opt, data):
    """BaseManager.preprocess_from_dataloader, base_manager.py:50-66."""
    nc = opt.label_nc + (1 if opt.contain_dontcare_label else 0)
    out = {
        "input_semantics": torch.from_numpy(preprocess_label_np(data["label"].numpy(), nc)),
        "image_lr": downsample_image(data["image"], opt.start_size, opt.downsampling_method),
        "image_hr": data["image"],
    }
    if opt.guiding_style_image:
        out["guiding_image"] = data["guiding_image"]
        out["guiding_label"] = torch.from_numpy(
            preprocess_label_np(data["guiding_label"].numpy(), nc))
    return out


# =================================================================================================
# one training iteration on CPU (trainer_manager.py:32-61) - used as the CPU baseline of bench.py
# =================================================================================================
class CpuTrainer:
    def __init__(self, opt, sdG, sdE, sdD):
        self.opt = opt
        self.sdG, self.sdE, self.sdD = sdG, sdE, sdD
        self.rng = _random.Random(0)
        for sd in (sdG, sdE, sdD):
            for k, v in sd.items():
                if v.is_floating_point() and not _is_buffer(k):
                    v.requires_grad_(True)
        # SRModel.create_optimizers, sr_model.py:469-495
        g_params = [v for k, v in sdG.items() if v.requires_grad]
        low = [v for k, v in sdE.items() if v.requires_grad and "mini" in k]
        g_params += [v for k, v in sdE.items() if v.requires_grad and "mini" not in k]
        lr_g, lr_d = (opt.lr, opt.lr) if opt.no_TTUR else (opt.lr / 2, opt.lr * 2)
        self.opt_G = torch.optim.Adam([{"params": g_params}, {"params": low, "lr": lr_g / 4}],
                                      lr=lr_g, betas=(opt.beta1, opt.beta2))
        self.opt_D = torch.optim.Adam([v for v in sdD.values() if v.requires_grad], lr=lr_d,
                                      betas=(opt.beta1, opt.beta2))

    def _noise_fn(self, name, shape):
        return torch.randn(shape)

    def _style_noise(self, shape):
        return torch.rand(shape)

    def _fake(self, d):
        opt = self.opt
        style_img = d.get("guiding_image") if opt.guiding_style_image else d["image_hr"]
        z, _ = encode_style(self.sdE, opt, self.rng, True, d["image_lr"], d["input_semantics"],
                            style_img, d.get("guiding_image"), d.get("guiding_label"),
                            unit_noise_fn=self._style_noise)
        return generator_forward(self.sdG, opt, d["image_lr"], d["input_semantics"], z, True,
                                 self._noise_fn)

    def generator_step(self, d):
        self.opt_G.zero_grad()
        fake = self._fake(d)
        losses = generator_losses(self.sdD, self.opt, d["input_semantics"], fake, d["image_hr"])
        sum(losses.values()).mean().backward()
        self.opt_G.step()
        return losses, fake.detach()

    def discriminator_step(self, d):
        self.opt_D.zero_grad()
        with torch.no_grad():
            fake = self._fake(d)
        losses = discriminator_losses(self.sdD, self.opt, d["input_semantics"], fake, d["image_hr"])
        sum(losses.values()).mean().backward()
        self.opt_D.step()
        return losses


def _is_buffer(key):
    return key.endswith(("weight_u", "weight_v", "running_mean", "running_var",
                         "num_batches_tracked"))
