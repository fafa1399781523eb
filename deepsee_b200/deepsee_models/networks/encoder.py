"""Regional style encoders on the B200 kernels (reference: deepsee_models/networks/encoder.py).

Same class names, ``forward(x, seg, mode, no_noise)`` signatures and state_dict keys.  Each stage
``spectral conv3x3 (no bias) -> InstanceNorm2d(affine=False) -> LeakyReLU(0.2)`` runs as
``ops.conv2d_direct`` (+ folded nearest upsample / stride) and ``ops.instance_norm``; the
region-wise masked mean (encoder.py:36-49) is ``ops.region_pool`` on the uint8 label map, so the
[B,19,C,H,W] product of the reference is never formed.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ... import ops
from .base_network import BaseNetwork
from .normalization import get_nonspade_norm_layer, effective_weight, spectral_prepass


def _khwc(w, pad_cin=None):
    """[Cout,Cin,KH,KW] -> [KH,KW,Cin(p),Cout] (input channels zero-padded to pad_cin)."""
    t = w.permute(2, 3, 1, 0)
    if pad_cin is not None and pad_cin > t.shape[2]:
        t = torch.nn.functional.pad(t, (0, 0, 0, pad_cin - t.shape[2]))
    return t.contiguous()


def _stage(x, seq, stride=1, ups=0, act=1, pad_cin=None):
    """x NHWC -> act(instance_norm(conv(x))). ``seq`` = nn.Sequential(spectral conv, InstanceNorm).
    The spectral-normalised weight stays an autograd tensor, so the weight gradient the conv node
    returns flows on to ``weight_orig`` through torch's own spectral-norm graph."""
    conv = seq[0]
    y = ops.conv_layer(x, effective_weight(conv), None, stride, 1, ups=ups)
    return ops.InstanceNormFn.apply(y, act)


class AbtractStyleEncoder(BaseNetwork):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.kw = 3
        self.pw = int(np.ceil((self.kw - 1.0) / 2))
        self.nf = opt.nef
        self.out_size = opt.regional_style_size
        self.norm_layer = get_nonspade_norm_layer(opt, opt.norm_E)
        self.final = nn.Sequential(
            self.norm_layer(nn.Conv2d(self.nf * 8, self.out_size, self.kw, stride=1, padding=self.pw)),
            nn.Tanh())

    def _labels(self, seg):
        from ...data.onehot import labels_of
        return labels_of(seg)[0]

    def extract_style_matrix(self, x_nhwc, labels_full, n_regions=None):
        """encoder.py:36-49 (divides by H*W of the feature map, not by the region area).  One row per
        channel of the segmentation input (seg.size(1) = semantic_nc: with contain_dontcare_label
        that is label_nc + 1)."""
        B, H, W, _ = x_nhwc.shape
        labels = ops.resize_labels(labels_full, H, W)
        return ops.RegionPoolFn.apply(x_nhwc, labels, n_regions or self.opt.semantic_nc)

    def corrupt_style_matrix(self, style_matrix, max_range_noise, region_idx=None):
        """encoder.py:51-70 (all regions). Tiny [B,19,128] elementwise op; the uniform draw uses
        torch's CUDA generator exactly like the reference's torch.rand_like."""
        if region_idx:
            raise NotImplementedError('region_idx subsets are a demo-only feature')
        w = torch.sigmoid(self.noise_weights).view(1, -1, 1)
        if self.opt.noisy_style_dist == 'uniform':
            noise = (self._unit_noise(style_matrix) * 2 - 1) * max_range_noise
        elif self.opt.noisy_style_dist == 'normal':
            noise = (torch.randn_like(style_matrix) * 2 - 1) * max_range_noise
        else:
            raise ValueError("Does not exist: {}".format(self.opt.noisy_style_dist))
        return (style_matrix + noise * w).clamp(-1, 1)

    def _unit_noise(self, like):
        """U[0,1) draw of corrupt_style_matrix (encoder.py:62); a method so tests can replay the
        reference's draw."""
        return torch.rand_like(like)

    def _final(self, x):
        return _stage(x, self.final[0], act=2)


class FullStyleEncoder(AbtractStyleEncoder):
    """encoder.py:73-132."""

    def __init__(self, opt):
        super().__init__(opt)
        if opt.random_style_matrix:
            raise NotImplementedError('random_style_matrix is an ablation switch, not implemented')
        nf, kw, pw = self.nf, self.kw, self.pw
        self.layers = OrderedDict()
        self.layers["initial"] = nn.Sequential(
            self.norm_layer(nn.Conv2d(3, nf, kw, stride=1, padding=pw)), nn.LeakyReLU(0.2, False))
        self.layers["down0"] = nn.Sequential(
            self.norm_layer(nn.Conv2d(nf, nf * 2, kw, stride=2, padding=pw)), nn.LeakyReLU(0.2, False))
        self.layers["down1"] = nn.Sequential(
            self.norm_layer(nn.Conv2d(nf * 2, nf * 4, kw, stride=2, padding=pw)), nn.LeakyReLU(0.2, False))
        self.layers["up_conv"] = nn.Sequential(
            nn.Upsample(scale_factor=2),
            self.norm_layer(nn.Conv2d(nf * 4, nf * 8, kw, padding=pw)), nn.LeakyReLU(0.2, False))
        for name, module in self.layers.items():
            self.add_module(name, module)
        self.up = nn.Upsample(scale_factor=2, mode='bilinear')
        self.noisy_style = "fullstyle" in self.opt.netE and self.opt.noisy_style_scale > 0
        if self.noisy_style:
            self.noise_weights = nn.Parameter(torch.zeros(opt.label_nc), requires_grad=True)
            self.actv_weights = nn.Sigmoid()
            self.max_range_noise = self.opt.noisy_style_scale

    def forward_main(self, x_nchw):
        # RGB zero-padded to 16 channels: the narrowest input the tensor-core conv takes
        x = ops.nchw_to_nhwc(x_nchw.contiguous().float(), 16)
        x = _stage(x, self.initial[0])
        x = _stage(x, self.down0[0], stride=2)
        x = _stage(x, self.down1[0], stride=2)
        x = _stage(x, self.up_conv[1], ups=1)
        return x, None

    def main_convs(self):
        return [self.initial[0][0], self.down0[0][0], self.down1[0][0], self.up_conv[1][0]]

    def forward(self, x=None, seg=None, mode="full", no_noise=False):
        with spectral_prepass(self.main_convs() + [self.final[0][0]]):
            x, activations = self.forward_main(x)
            x = self._final(x)
        style_matrix = self.extract_style_matrix(x, self._labels(seg), seg.size(1))
        if self.noisy_style and not no_noise:
            style_matrix = self.corrupt_style_matrix(style_matrix, self.max_range_noise)
        return style_matrix, activations


class MinistyleEncoder(AbtractStyleEncoder):
    """encoder.py:135-175."""

    def __init__(self, opt):
        super().__init__(opt)
        nf, kw, pw = self.nf, self.kw, self.pw
        self.layers = OrderedDict()
        self.layers["initial"] = nn.Sequential(
            self.norm_layer(nn.Conv2d(3, nf, kw, stride=1, padding=pw)), nn.LeakyReLU(0.2, False))
        self.layers["conv0"] = nn.Sequential(
            self.norm_layer(nn.Conv2d(nf, nf * 2, kw, stride=1, padding=pw)), nn.LeakyReLU(0.2, False))
        self.layers["conv1"] = nn.Sequential(
            self.norm_layer(nn.Conv2d(nf * 2, nf * 4, kw, stride=1, padding=pw)), nn.LeakyReLU(0.2, False))
        self.layers["conv2"] = nn.Sequential(
            nn.Upsample(scale_factor=2),
            self.norm_layer(nn.Conv2d(nf * 4, nf * 8, kw, padding=pw)), nn.LeakyReLU(0.2, False))
        for name, module in self.layers.items():
            self.add_module(name, module)

    def forward_main(self, x_nchw):
        x = ops.nchw_to_nhwc(x_nchw.contiguous().float(), 16)
        x = _stage(x, self.initial[0])
        x = _stage(x, self.conv0[0])
        x = _stage(x, self.conv1[0])
        x = _stage(x, self.conv2[1], ups=1)
        return x, None

    def main_convs(self):
        return [self.initial[0][0], self.conv0[0][0], self.conv1[0][0], self.conv2[1][0]]

    def forward(self, x=None, seg=None, mode="mini"):
        with spectral_prepass(self.main_convs() + [self.final[0][0]]):
            x, activations = self.forward_main(x)
            x = self._final(x)
        return self.extract_style_matrix(x, self._labels(seg), seg.size(1)), activations


class CombinedstyleEncoder(AbtractStyleEncoder):
    """encoder.py:178-210."""

    def __init__(self, opt):
        super().__init__(opt)
        self.encoder_full = FullStyleEncoder(opt)
        self.encoder_mini = MinistyleEncoder(opt)
        self.final = nn.Sequential(
            self.norm_layer(nn.Conv2d(self.nf * 8, self.out_size, self.kw, stride=1, padding=self.pw)),
            nn.Tanh())
        self.noisy_style = self.opt.noisy_style_scale > 0
        if self.noisy_style:
            self.noise_weights = nn.Parameter(torch.zeros(opt.label_nc), requires_grad=True)
            self.actv_weights = nn.Sigmoid()
            self.max_range_noise = self.opt.noisy_style_scale

    def forward(self, x=None, seg=None, mode=None, no_noise=False):
        if mode == "full":
            enc = self.encoder_full
        elif mode == "mini":
            enc = self.encoder_mini
        else:
            raise NotImplementedError()
        # only the branch that runs is normalised (its u / v advance like the reference's hooks)
        with spectral_prepass(enc.main_convs() + [self.final[0][0]]):
            x, activations = enc.forward_main(x)
            x = self._final(x)
        style_matrix = self.extract_style_matrix(x, self._labels(seg), seg.size(1))
        if self.noisy_style and not no_noise:
            style_matrix = self.corrupt_style_matrix(style_matrix, self.max_range_noise)
        return style_matrix, activations
