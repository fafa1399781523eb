"""SPADEResnetBlock on the fused B200 kernels (reference: deepsee_models/networks/architecture.py:24-147).

One block = 2 x (K1 modulate -> K2 conv):

    x (fp32 NHWC, possibly at half resolution: the nn.Upsample in front of the block is folded
       into the kernels' input indexing and never materialised)
      -> K1[norm_0]: gamma/beta GEMM, epilogue BN + modulation + LeakyReLU -> fp16 planes a0
      -> K2[conv_0]: implicit GEMM + bias (+ noise_middle) -> dx fp32, BN partial sums
      -> K1[norm_1] -> a1
      -> K2[conv_1]: + bias + shortcut (x through the folded upsample, + noise_in/noise_skip)
         -> block output fp32 NHWC and its BN partial sums for the next block's norm_0

fin == fout everywhere in DeepSEESR, so the learned shortcut (conv_s / norm_s) is never built
(architecture.py:30,36); asking for it raises.
"""
import torch
import torch.nn as nn
import torch.nn.utils.spectral_norm as spectral_norm

from ... import ops
from ...config import config
from .normalization import (SPADE, SEAN_Block, PureSEAN_Block, NoiseInjection, effective_weight,
                            BN_EPS)

BN_MOMENTUM = 0.1


class SPADEResnetBlock(nn.Module):
    def __init__(self, fin, fout, opt, style=True, puresean=False):
        super().__init__()
        self.opt = opt
        self.efficient = opt.efficient
        self.learned_shortcut = (fin != fout)
        if self.learned_shortcut:
            raise NotImplementedError('learned shortcut (fin != fout) is not part of the DeepSEE '
                                      'generator and is not implemented on the B200 path')
        fmiddle = min(fin, fout)
        self.conv_0 = nn.Conv2d(fin, fmiddle, kernel_size=3, padding=1)
        self.conv_1 = nn.Conv2d(fmiddle, fout, kernel_size=3, padding=1)
        if 'spectral' in opt.norm_G:
            self.conv_0 = spectral_norm(self.conv_0)
            self.conv_1 = spectral_norm(self.conv_1)
        cfg = opt.norm_G.replace('spectral', '')
        NormBlock = self._get_block(cfg, style, puresean)
        self.norm_0 = NormBlock(cfg, fin, opt.semantic_nc, opt)
        self.norm_1 = NormBlock(cfg, fmiddle, opt.semantic_nc, opt)
        # like the reference, noise modules exist only if add_noise was on at construction time
        if self.add_noise:
            self.noise_in = NoiseInjection(fin)
            self.noise_skip = NoiseInjection(fin)
            self.noise_middle = NoiseInjection(fmiddle)
        self._wcache = {}

    @property
    def add_noise(self):
        return self.opt.add_noise and self.training

    def _get_block(self, config_text, style, puresean=False):
        if puresean:
            return PureSEAN_Block
        if style and 'sean' in config_text:
            return SEAN_Block
        return SPADE

    # -- prepared main-conv weights -------------------------------------------------------------
    def _prepared_conv(self, conv, name, want_lo):
        w = effective_weight(conv)  # spectral norm: W_orig / sigma (power iteration when training)
        if torch.is_grad_enabled() and w.requires_grad:
            return ops.prep_conv_weight(w.detach().contiguous(), want_lo=want_lo)
        src = [getattr(conv, n) for n in ('weight_orig', 'weight_u', 'weight_v') if hasattr(conv, n)]
        if not src:
            src = [conv.weight]
        key = tuple((t.data_ptr(), t._version) for t in src) + (want_lo, self.training)
        hit = self._wcache.get(name)
        if hit is not None and hit[0] == key and not self.training:
            return hit[1]
        pw = ops.prep_conv_weight(w.detach().contiguous(), want_lo=want_lo)
        self._wcache[name] = (key, pw)
        return pw

    # -- batch norm ------------------------------------------------------------------------------
    def _bn_affine(self, norm, partials, count, unbias_count):
        bn = norm.param_free_norm
        if not self.training:
            return norm.eval_affine()
        sc, sh, _, _ = ops.bn_finalize(partials, count, BN_EPS, BN_MOMENTUM, bn.running_mean,
                                       bn.running_var, unbias_count=unbias_count)
        bn.num_batches_tracked.add_(1)
        return sc, sh

    def forward_nhwc(self, x, ctx, ups=0, stats_in=None):
        """x fp32 NHWC [B, H>>ups, W>>ups, C] -> (out fp32 NHWC [B,H,W,C], stats partials of out or
        None). ``stats_in``: BN partial sums of x from the producing kernel (training only)."""
        B, Hx, Wx, C = x.shape
        H, W = Hx << ups, Wx << ups
        passes = config.passes
        want_lo = passes == 3
        training = self.training
        noisy = self.add_noise
        n_in = n_skip = n_mid = None
        w_in = w_skip = w_mid = None
        if noisy:
            n_in, w_in = self.noise_in.sample(B, H, W), self.noise_in.weight
            n_skip, w_skip = self.noise_skip.sample(B, H, W), self.noise_skip.weight
            n_mid, w_mid = self.noise_middle.sample(B, H, W), self.noise_middle.weight

        # ---- norm_0 + actvn -------------------------------------------------------------------
        part = None
        count = ucount = 0
        if training:
            if stats_in is not None and not noisy:
                # statistics of a nearest-upsampled tensor = statistics of its source
                part, count, ucount = stats_in, B * Hx * Wx, B * H * W
            else:
                part = ops.bn_stats(x, ups, n_in, w_in)
                count = ucount = B * H * W
        sc, sh = self._bn_affine(self.norm_0, part, count, ucount)
        pw, gb, bb = self.norm_0.prepared(want_lo)
        srcs = self.norm_0.build_sources(ctx, H, W, want_lo)
        a0 = ops.spade_modulate(srcs, pw, x, ups, sc, sh, gb, bb, noise=n_in, noise_w=w_in,
                                passes=passes, want_lo=want_lo)
        # ---- conv_0 (+ noise_middle) ----------------------------------------------------------
        pw0 = self._prepared_conv(self.conv_0, 'conv_0', want_lo)
        r = ops.conv3x3([a0], pw0, self.conv_0.bias, noises=[(n_mid, w_mid)] if noisy else (),
                        passes=passes, want_stats=training)
        dx, part1 = r if training else (r, None)
        del a0
        # ---- norm_1 + actvn -------------------------------------------------------------------
        sc1, sh1 = self._bn_affine(self.norm_1, part1, B * H * W, B * H * W)
        pw, gb, bb = self.norm_1.prepared(want_lo)
        srcs = self.norm_1.build_sources(ctx, H, W, want_lo)
        a1 = ops.spade_modulate(srcs, pw, dx, 0, sc1, sh1, gb, bb, passes=passes, want_lo=want_lo)
        del dx
        # ---- conv_1 + shortcut ------------------------------------------------------------------
        pw1 = self._prepared_conv(self.conv_1, 'conv_1', want_lo)
        r = ops.conv3x3([a1], pw1, self.conv_1.bias, residual=x, res_ups=ups,
                        noises=[(n_in, w_in), (n_skip, w_skip)] if noisy else (), passes=passes,
                        want_stats=training)
        return r if training else (r, None)

    def forward(self, x, seg, style=None, split_location=-1):
        raise RuntimeError('SPADEResnetBlock is driven by DeepSEESR.forward on the B200 path '
                           '(NHWC activations, fused kernels); use forward_nhwc')
