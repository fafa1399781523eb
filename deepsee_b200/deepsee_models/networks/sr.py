"""DeepSEESR generator on the B200 kernels (reference: deepsee_models/networks/sr.py:10-98).

Same constructor, ``forward(x_downsized, seg, z)`` signature, module names and state_dict keys as
the reference; NCHW fp32 in (LR image, one-hot semantics), NCHW fp32 out.  Inside, activations are
NHWC and never upsampled in memory: every ``self.up`` of the reference is folded into the next
block's kernels (K1 reads x at (y>>1, x>>1), K2 reads the shortcut the same way).
The reference's model-parallel modes (sr.py:73-92) are dropped: one B200 holds the 512x512 model.
"""
import numpy as np
import torch
import torch.nn as nn

from ... import ops
from ...config import config
from .architecture import SPADEResnetBlock
from .base_network import BaseNetwork
from .normalization import GenContext, spectral_prepass


class _StemFn(torch.autograd.Function):
    """DeepSEESR.initial (sr.py:31,65). The LR image needs no gradient."""

    @staticmethod
    def forward(ctx, x_nchw, w, b):
        ctx.save_for_backward(x_nchw)
        return ops.stem(x_nchw, w.contiguous(), b)

    @staticmethod
    def backward(ctx, dy):
        (x_nchw,) = ctx.saved_tensors
        dw, db = ops.stem_bwd(x_nchw, dy.contiguous())
        return None, dw, db


class _HeadFn(torch.autograd.Function):
    """F.tanh(conv_img(F.leaky_relu(x, 0.2))) (sr.py:56,94-95), NHWC in, NCHW out."""

    @staticmethod
    def forward(ctx, x_nhwc, w, b, planes=None):
        w = w.contiguous()
        if planes is not None:
            # tensor-core form (config.head_tc): `planes` = leaky_relu(x) as fp16 hi / lo planes from the
            # epilogue of the kernel that produced x; x itself is not read again.  The forward GEMM
            # always uses both planes (fp32-class: it is the last layer in front of the 1e-3 bound)
            out = ops.head_tc(planes, w, b, passes=3)
            ctx.planes = ops.SplitPlanes(planes.hi, planes.lo if config.passes == 3 else None)
            ctx.save_for_backward(w, out)
            return out
        ctx.planes = None
        out = ops.head(x_nhwc, w, b)
        ctx.save_for_backward(x_nhwc, w, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.planes is not None:
            w, out = ctx.saved_tensors
            planes, ctx.planes = ctx.planes, None
            dx, amax, dw, db = ops.head_tc_bwd(planes, w, out, dout.contiguous(), passes=config.passes)
            dx._dsee_amax = (amax, dx._version)   # the top block's grad_prep skips its max|dout| pass
            return dx, dw, db, None
        x_nhwc, w, out = ctx.saved_tensors
        dx, dw, db = ops.head_bwd(x_nhwc, w, out, dout.contiguous())
        return dx, dw, db, None


class DeepSEESR(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument('--num_upsampling_layers', choices=('normal', 'more', 'most'),
                            default='normal')
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        nf = opt.ngf
        self.start_size = opt.start_size
        self.n_blocks = int(np.log2(opt.crop_size) - np.log2(self.start_size))
        self.sw, self.sh = self.compute_latent_vector_size(opt)
        if getattr(opt, 'model_parallel_mode', 0):
            raise NotImplementedError('model_parallel_mode is replaced by data parallelism on B200 '
                                      '(180 GB HBM holds the 512x512 model); use mode 0')
        self.initial = nn.Conv2d(3, 16 * nf, 3, padding=1)
        early_style = not ("late" in self.opt.norm_G)
        self.head_0 = SPADEResnetBlock(16 * nf, 16 * nf, opt, style=early_style)
        self.G_middle_0 = SPADEResnetBlock(16 * nf, 16 * nf, opt, style=True)
        self.G_middle_1 = SPADEResnetBlock(16 * nf, 16 * nf, opt, style=True)
        self.mp = 0
        max_n_blocks = 4 if self.opt.load_size >= 512 else 99  # sr.py:41-43
        ups = [SPADEResnetBlock(16 * nf, 16 * nf, opt, style=True)
               for _ in range(1, min(self.n_blocks, max_n_blocks))]
        if max_n_blocks != 99:
            for _ in range(max_n_blocks, self.n_blocks):  # 512 and 1024: PureSEAN tail (sr.py:48-51)
                ups.append(SPADEResnetBlock(16 * nf, 16 * nf, opt, style=True, puresean=True))
        self.up_list = nn.ModuleList(ups)
        self.conv_img = nn.Conv2d(16 * nf, 3, 3, padding=1)
        self.up = nn.Upsample(scale_factor=2)  # kept for module-tree parity; folded, never run

    def get_device(self):
        return self.initial.weight.device

    def forward(self, x_downsized, seg=None, z=None):
        if not x_downsized.is_cuda:
            raise RuntimeError('DeepSEESR (B200 path) needs CUDA tensors; there is no CPU fallback')
        from ...data.onehot import labels_of
        x_downsized = x_downsized.contiguous().float()
        labels, bad = labels_of(seg)   # free for the preprocessor's OneHotLabels
        ctx = GenContext(labels, z.contiguous().float() if z is not None else None)

        blocks = [self.head_0, self.G_middle_0, self.G_middle_1] + [self.up_list[i] for i in range(self.n_blocks - 1)]
        if config.head_tc and self.conv_img.weight.shape[1] % 64 == 0:
            ctx.head_act_block = blocks[-1]   # its last kernel also writes the head's fp16 operand planes
        x = _StemFn.apply(x_downsized, self.initial.weight, self.initial.bias)
        with spectral_prepass([c for blk in blocks for c in (blk.conv_0, blk.conv_1)]):
            x, st = self.head_0.forward_nhwc(x, ctx, ups=0)
            x, st = self.G_middle_0.forward_nhwc(x, ctx, ups=1, stats_in=st)
            x, st = self.G_middle_1.forward_nhwc(x, ctx, ups=0, stats_in=st)
            for i in range(self.n_blocks - 1):
                x, st = self.up_list[i].forward_nhwc(x, ctx, ups=1, stats_in=st)
        out = _HeadFn.apply(x, self.conv_img.weight, self.conv_img.bias, getattr(ctx, 'head_act', None))
        if config.check_onehot and bad is not None:
            self._check_onehot(bad)
        return out

    _ONEHOT_MSG = ('DeepSEESR: `seg` must be a one-hot map (exactly one 1.0 per pixel); '
                   'the B200 path consumes it as an integer label map')

    def _check_onehot(self, bad):
        """`bad` is a device flag set by labels_from_onehot.  Reading it with .item() would stall the
        launch pipeline once per forward, so in training mode the flag travels to pinned host memory
        asynchronously and is examined at the NEXT forward (by then the copy finished long ago): a bad
        batch raises one call late, a training loop keeps running ahead of the GPU.  In eval mode
        (demo / inference, one call at a time) the check is immediate."""
        pend = getattr(self, '_onehot_pending', None)
        if pend is not None:
            host, ev = pend
            ev.synchronize()
            self._onehot_pending = None
            if int(host[0]) != 0:
                raise ValueError(self._ONEHOT_MSG + ' (detected in the previous forward)')
        if not self.training:
            if int(bad.item()) != 0:
                raise ValueError(self._ONEHOT_MSG)
            return
        host = torch.empty(1, dtype=bad.dtype, pin_memory=True)
        host.copy_(bad, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._onehot_pending = (host, ev)
