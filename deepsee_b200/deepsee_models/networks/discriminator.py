"""Multi-scale PatchGAN discriminator on the B200 kernels (reference:
deepsee_models/networks/discriminator.py).

``forward`` accepts what ``SRModel.discriminate`` builds.  The reference concatenates
[one-hot semantics | image] on channels and fake | real on the batch (sr_model.py:655-664) into a
fp32 NCHW tensor; here that tensor may be passed as-is (it is converted to NHWC once), or the
fused ``ops.disc_input`` result (NHWC, channel-padded) can be passed through ``forward_nhwc``.
Returned feature maps are NCHW views so the loss code reads like the reference's.
"""
import numpy as np
import torch.nn as nn

from ... import ops
from .base_network import BaseNetwork

from .normalization import get_nonspade_norm_layer, effective_weight, spectral_prepass


class MultiscaleDiscriminator(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument('--netD_subarch', type=str, default='n_layer')
        parser.add_argument('--num_D', type=int, default=2)
        NLayerDiscriminator.modify_commandline_options(parser, is_train)
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        for i in range(opt.num_D):
            self.add_module('discriminator_%d' % i, self.create_single_discriminator(opt))

    def create_single_discriminator(self, opt):
        if opt.netD_subarch != 'n_layer':
            raise ValueError('unrecognized discriminator subarchitecture %s' % opt.netD_subarch)
        return NLayerDiscriminator(opt)

    def downsample(self, x_nhwc):
        """discriminator.py:46-49: avg_pool2d(3, stride 2, pad 1, count_include_pad=False)."""
        return ops.AvgPool3s2Fn.apply(x_nhwc)

    def forward_nhwc(self, x_nhwc, detach_params=False):
        """``detach_params``: the generator step only needs the gradient wrt the input; detaching
        the weights skips the discriminator's weight-gradient kernels there (the reference computes
        and then discards those gradients: trainer_manager.py:48-49 zeroes them before use)."""
        result = []
        feats = not self.opt.no_ganFeat_loss
        convs = [c for _, D in self.named_children() for c in D.sn_convs()]
        with spectral_prepass(convs):
            for name, D in self.named_children():
                out = D.forward_nhwc(x_nhwc, detach_params)
                result.append(out if feats else [out[-1]])
                x_nhwc = self.downsample(x_nhwc)
        return result

    def forward(self, input):
        cp = (input.shape[1] + 31) // 32 * 32
        res = self.forward_nhwc(ops.nchw_to_nhwc(input.contiguous().float(), cp))
        return [[t.permute(0, 3, 1, 2) for t in scale] for scale in res]


class NLayerDiscriminator(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument('--n_layers_D', type=int, default=4)
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        kw = 4
        padw = int(np.ceil((kw - 1.0) / 2))
        nf = opt.ndf
        input_nc = self.compute_D_input_nc(opt)
        norm_layer = get_nonspade_norm_layer(opt, opt.norm_D)
        sequence = [[nn.Conv2d(input_nc, nf, kernel_size=kw, stride=2, padding=padw),
                     nn.LeakyReLU(0.2, False)]]
        for n in range(1, opt.n_layers_D):
            nf_prev = nf
            nf = min(nf * 2, 512)
            stride = 1 if n == opt.n_layers_D - 1 else 2
            sequence += [[norm_layer(nn.Conv2d(nf_prev, nf, kernel_size=kw, stride=stride,
                                               padding=padw)), nn.LeakyReLU(0.2, False)]]
        sequence += [[nn.Conv2d(nf, 1, kernel_size=kw, stride=1, padding=padw)]]
        for n in range(len(sequence)):
            self.add_module('model' + str(n), nn.Sequential(*sequence[n]))
        self.n_layers = opt.n_layers_D

    def compute_D_input_nc(self, opt):
        return opt.label_nc + opt.output_nc + (1 if opt.contain_dontcare_label else 0)

    def sn_convs(self):
        return [getattr(self, 'model%d' % n)[0][0] for n in range(1, self.n_layers)]

    def forward_nhwc(self, x, detach_params=False):
        """x NHWC [B,H,W,Cp] (channels beyond input_nc are zero) -> list of NHWC feature maps."""
        det = (lambda t: t.detach() if t is not None else None) if detach_params else (lambda t: t)
        from ...config import config
        fp = config.d_fwd_passes   # 1-pass mode: the discriminator only feeds the losses
        outs = []
        conv0 = self.model0[0]
        x = ops.conv_layer(x, det(conv0.weight), det(conv0.bias), 2, 2, lrelu=True, fwd_passes=fp)
        outs.append(x)
        for n in range(1, self.n_layers):
            seq = getattr(self, 'model%d' % n)[0]  # Sequential(spectral conv, InstanceNorm2d)
            conv = seq[0]
            stride = 1 if n == self.n_layers - 1 else 2
            y = ops.conv_layer(x, det(effective_weight(conv)), None, stride, 2, fwd_passes=fp)
            x = ops.InstanceNormFn.apply(y, 1)
            outs.append(x)
        last = getattr(self, 'model%d' % self.n_layers)[0]
        x = ops.conv_layer(x, det(last.weight), det(last.bias), 1, 2, fwd_passes=fp)
        outs.append(x)
        return outs

    def forward(self, input):
        cp = (input.shape[1] + 31) // 32 * 32
        outs = self.forward_nhwc(ops.nchw_to_nhwc(input.contiguous().float(), cp))
        outs = [t.permute(0, 3, 1, 2) for t in outs]
        return outs if not self.opt.no_ganFeat_loss else outs[-1]
