"""Conditional normalisation layers of the generator (SPADE / SEAN / PureSEAN), noise injection
and the spectral+instance wrapper used by the encoder and discriminator.

Same class names, constructor signatures and ``state_dict`` keys as the reference
(deepsee_models/networks/normalization.py), but the layers do not compute with ATen: each one
only *describes* its work to the fused K1 kernel (``ops.spade_modulate``):

  * ``mlp_shared`` (conv 19->128 over a one-hot map + ReLU) becomes a 9-tap table gather on the
    uint8 label map (``ops.shared_mlp``), exact because every input pixel has one non-zero channel;
  * ``style_map`` (normalization.py:182-185) becomes a gather ``style[b, label]``
    (``ops.style_gather``) - the [B,19,128,H,W] product is never formed;
  * the alpha blend of the SEAN branches (normalization.py:208-213) is folded into one weight
    matrix over the concatenated [actv | style_map] channels, gamma/beta rows interleaved per
    128 channels so one accumulator tile holds both;
  * batch norm, ``x_hat * (1 + gamma) + beta`` and LeakyReLU run in the kernel epilogue.
"""
import re

import torch
import torch.nn as nn
import torch.nn.utils.spectral_norm as spectral_norm

from ... import ops

NHIDDEN = 128  # normalization.py:95 ("Yes, hardcoded.")
BN_EPS = 1e-5


def effective_weight(conv):
    """Weight a (possibly spectral-normalised) conv uses in its forward (architecture.py:40-44;
    torch/nn/utils/spectral_norm.py:92-113): one power iteration in training mode (u / v buffers
    updated in place), W_orig / sigma, u and v constants for autograd.  The module keeps torch's
    parametrisation (weight_orig / weight_u / weight_v in the state_dict); the arithmetic runs in
    ops.SpectralWeightFn.  DSEE_TORCH_SPECTRAL=1 runs torch's own hook instead."""
    from ...config import config
    if not hasattr(conv, 'weight_orig'):
        return conv.weight
    if config.torch_spectral or not conv.weight_orig.is_cuda:
        for hook in conv._forward_pre_hooks.values():
            hook(conv, None)
        return conv.weight
    pre = getattr(conv, '_sn_pre', None)
    if pre is not None:
        conv._sn_pre = None  # computed by spectral_prepass for this forward; consumed exactly once
    return ops.SpectralWeightFn.apply(conv.weight_orig, conv.weight_u, conv.weight_v, conv.training,
                                      _sn_eps(conv), torch.is_grad_enabled(), pre)


def _sn_eps(conv):
    eps = 1e-12
    for hook in conv._forward_pre_hooks.values():
        eps = getattr(hook, 'eps', eps)
    return eps


class spectral_prepass:
    """``with spectral_prepass(convs): <forward that calls effective_weight on each of them once>``:
    runs the spectral normalisation of all the listed layers as one batched launch sequence up front
    (same arithmetic as the per-layer path, ~5 launches per network instead of per layer) and hands
    each layer its result through ``conv._sn_pre``.  Only layers the forward really uses may be
    listed: the power iteration advances their u / v like the reference's forward-pre-hook does.
    Whatever was not consumed is dropped on exit."""

    def __init__(self, convs):
        from ...config import config
        self.convs = [c for c in convs if hasattr(c, 'weight_orig')]
        self.on = (config.batched_spectral and not config.torch_spectral and len(self.convs) > 1 and
                   all(c.weight_orig.is_cuda for c in self.convs) and
                   len({(c.training, _sn_eps(c)) for c in self.convs}) == 1)

    def __enter__(self):
        if self.on:
            c0 = self.convs[0]
            res = ops.spectral_prepass(self.convs, c0.training, _sn_eps(c0), torch.is_grad_enabled())
            for c, r in zip(self.convs, res):
                c._sn_pre = r
        return self

    def __exit__(self, *exc):
        for c in self.convs:
            c._sn_pre = None
        return False


def get_nonspade_norm_layer(opt, norm_type='instance', oneD=False):
    """normalization.py:19-54: spectral norm on the layer, bias removed, followed by a
    parameter-free norm layer. Returns nn.Sequential(layer, norm) so the state_dict keys match;
    the encoder / discriminator forward read the parameters and call the CUDA kernels."""
    def get_out_channel(layer):
        return getattr(layer, 'out_channels', None) or layer.weight.size(0)

    def add_norm_layer(layer):
        subnorm_type = norm_type
        if norm_type.startswith('spectral'):
            layer = spectral_norm(layer)
            subnorm_type = norm_type[len('spectral'):]
        if subnorm_type == 'none' or len(subnorm_type) == 0:
            return layer
        if getattr(layer, 'bias', None) is not None:
            delattr(layer, 'bias')
            layer.register_parameter('bias', None)
        if subnorm_type == 'instance':
            norm_layer = nn.InstanceNorm1d(get_out_channel(layer), affine=False) if oneD \
                else nn.InstanceNorm2d(get_out_channel(layer), affine=False)
        else:
            raise ValueError('normalization layer %s is not supported by the B200 path '
                             '(reference default is spectralinstance)' % subnorm_type)
        return nn.Sequential(layer, norm_layer)

    return add_norm_layer


class GenContext:
    """Per-forward conditioning state: the uint8 label map at every resolution it is needed at
    (always resized from the full-resolution map, like F.interpolate(segmap, ...) at
    normalization.py:110,174,261) and the style matrix."""

    def __init__(self, labels_full, style):
        self.labels_full = labels_full  # uint8 [B,S,S]
        self.style = style              # fp32 [B,L,d] or None
        self._cache = {}

    def labels_at(self, h, w):
        key = (h, w)
        if key not in self._cache:
            self._cache[key] = ops.resize_labels(self.labels_full, h, w)
        return self._cache[key]

    def onehot_at(self, h, w):
        """fp16 one-hot plane of the label map at (h, w): the second operand of the tensor-core
        weight gradient of mlp_shared (built on first use in the backward pass)."""
        key = ('onehot', h, w)
        if key not in self._cache:
            self._cache[key] = ops.onehot_planes(self.labels_at(h, w))
        return self._cache[key]


def fold_style_weight(Wm, style, nh=NHIDDEN):
    """Splits a SEAN layer's combined modulation weight [2C, nh + d, 3, 3] into the shared
    mlp_shared-activation part Wa [2C, nh, 3, 3] and the per-image style part
        Ws[b, o, l, ky, kx] = sum_s Wm[o, nh + s, ky, kx] * style[b, l, s]      ([B, 2C, L, 3, 3]),
    i.e. conv(style_map, W_sty) with style_map[b,:,y,x] = style[b, label(y,x), :]
    (normalization.py:182-185,198-201) rewritten as a conv over the one-hot label map.  Plain torch
    ops: autograd carries the gradient of Ws back into the style matrix and the weights."""
    Wa = Wm[:, :nh].contiguous()
    Ws = torch.einsum('osyx,bls->bolyx', Wm[:, nh:], style).contiguous()
    return Wa, Ws


def _interleave_gamma_beta(wg, wb):
    """[C,Cin,3,3] x2 -> [2C,Cin,3,3] with rows [g(0..127) | b(0..127) | g(128..255) | ...]."""
    C = wg.shape[0]
    t = torch.stack([wg.view(C // 128, 128, *wg.shape[1:]), wb.view(C // 128, 128, *wb.shape[1:])], 1)
    return t.reshape(2 * C, *wg.shape[1:]).contiguous()


def _param_free_norm(config_text, pattern, norm_nc):
    parsed = re.search(pattern, config_text)
    kind = str(parsed.group(1))
    ks = int(parsed.group(2))
    if ks != 3:
        raise ValueError('the B200 path implements 3x3 modulation convs (got %dx%d)' % (ks, ks))
    if 'batch' not in kind:
        raise ValueError('%s is not a supported param-free norm type (batch / syncbatch)' % kind)
    # SynchronizedBatchNorm2d(affine=False) keys: running_mean / running_var / num_batches_tracked
    return nn.BatchNorm2d(norm_nc, affine=False)


class _CondNormBase(nn.Module):
    """Shared machinery: table for mlp_shared, BN scale/shift, prepared-weight cache."""
    kind = None

    def _table(self):
        w = self.mlp_shared[0].weight  # [nh, L, 3, 3] -> [9, L, nh]
        return w.permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0]).contiguous()

    def table_and_bias(self):
        """mlp_shared as a 9-tap table [9,L,nh] + bias; an autograd view of the conv weight, so
        the table gradient K1's backward returns flows into ``mlp_shared.0.weight``."""
        return self._table(), self.mlp_shared[0].bias

    def fm_size(self, H, W):
        mx = getattr(self.opt, 'max_fm_size', 1 << 30) if self.kind != 'spade' else 1 << 30
        return min(H, mx), min(W, mx)

    def folds_style(self, H, W):
        """True when this layer's style branch runs as per-image weights over the one-hot label planes
        (config.fold_style): a SEAN layer at or below max_fm_size."""
        from ...config import config
        return (self.kind == 'sean' and config.fold_style and config.save_gamma and
                self.fm_size(H, W) == (H, W))

    def uses_subpixel(self, H, W):
        """True when this layer runs in the sub-pixel form (config.subpixel): its sources are
        convolved through a 2x upsample (feature map above max_fm_size)."""
        from ...config import config
        fh, fw = self.fm_size(H, W)
        return (config.subpixel and config.save_gamma and self.kind != 'spade' and
                (fh * 2, fw * 2) == (H, W))

    def build_sources(self, ctx, style, H, W, want_lo, table=None, bias=None, folded=False, sub=False):
        """-> (list of SplitPlanes feeding K1's A operand at resolution (H, W), meta). ``meta``
        records what each source is ('actv' = mlp_shared output, 'style' = gathered style matrix),
        the label map and the folded-upsample flag - what the backward pass needs."""
        fh, fw = self.fm_size(H, W)
        if (fh, fw) == (H, W):
            ups = 0
        elif (fh * 2, fw * 2) == (H, W):
            ups = 1
        else:
            raise NotImplementedError('feature map %dx%d vs max_fm_size %dx%d: only 1x / 2x '
                                      'supported' % (H, W, fh, fw))
        labels = ctx.labels_at(fh, fw)
        need_actv = self.kind in ('spade', 'sean') or ups == 1
        actv = None
        if need_actv:
            if table is None:
                table, bias = self.table_and_bias()
            # sub-pixel form: the activation stays at the label map's resolution (K1 reads it through
            # the collapsed 2x2 filters); otherwise the 2x upsample is materialised here
            actv = ops.shared_mlp(labels, table.detach(), bias.detach(), ups=0 if sub else ups, want_lo=want_lo)
        meta = {'labels': labels, 'ups': ups, 'actv': actv, 'fm': (fh, fw), 'ctx': ctx, 'sub': bool(sub)}
        if self.kind == 'spade':
            meta['kinds'] = ['actv']
            return [actv], meta
        if ups == 1:
            # normalization.py:188-190 / 275-277: style_map := upsampled actv (style is dropped)
            style_map, skind = actv, 'actv'
        elif folded:
            # the style branch lives in the per-image weights (fold_style_weight): K1 reads the exact
            # one-hot planes instead of a gathered style_map
            oh = ctx.onehot_at(fh, fw)
            style_map = ops.SplitPlanes(oh.hi, ops._zeros_like_cached(oh.hi) if want_lo else None)
            skind = 'onehot'
        else:
            if style is None:
                raise RuntimeError('%s needs a style matrix z' % type(self).__name__)
            style_map, skind = ops.style_gather(labels, style, want_lo=want_lo), 'style'
        if self.kind == 'sean':
            meta['kinds'] = ['actv', skind]
            return [actv, style_map], meta
        meta['kinds'] = [skind]
        return [style_map], meta

    def combined_weight(self):
        """-> (W [2C, Cin_total, 3, 3] rows interleaved, gamma_bias [C], beta_bias [C]) as one fused
        autograd node (ops.ModWeightFn); DSEE_TORCH_MODWEIGHT=1 builds it from torch ops."""
        from ...config import config
        if config.torch_modweight:
            return self.combined_weight_torch()
        g, b = getattr(self, 'mlp_gamma', None), getattr(self, 'mlp_beta', None)
        sg, sb = getattr(self, 'mlp_style_gamma', None), getattr(self, 'mlp_style_beta', None)
        if self.kind == 'spade':
            sg = sb = None
        elif self.kind == 'puresean':
            g = b = None
        w = lambda m: m.weight if m is not None else None
        bi = lambda m: m.bias if m is not None else None
        two = g is not None and sg is not None
        return ops.ModWeightFn.apply(self.kind != 'puresean', w(g), w(b), w(sg), w(sb), bi(g), bi(b),
                                     bi(sg), bi(sb), self.alpha_gamma if two else None,
                                     self.alpha_beta if two else None)

    def combined_weight_torch(self):
        """The same tensors from plain torch ops (reference for the fused node's tests)."""
        if self.kind == 'spade':
            wg, wb = self.mlp_gamma.weight, self.mlp_beta.weight
            gb, bb = self.mlp_gamma.bias + 1.0, self.mlp_beta.bias
        elif self.kind == 'sean':
            a_b = torch.sigmoid(self.alpha_beta)
            a_g = torch.sigmoid(self.alpha_gamma)
            wg = torch.cat([(1.0 - a_g) * self.mlp_gamma.weight, a_g * self.mlp_style_gamma.weight], 1)
            wb = torch.cat([(1.0 - a_b) * self.mlp_beta.weight, a_b * self.mlp_style_beta.weight], 1)
            gb = (1.0 - a_g) * self.mlp_gamma.bias + a_g * self.mlp_style_gamma.bias + 1.0
            bb = (1.0 - a_b) * self.mlp_beta.bias + a_b * self.mlp_style_beta.bias
        else:  # puresean: no "+1" (normalization.py:286)
            wg, wb = self.mlp_style_gamma.weight, self.mlp_style_beta.weight
            gb, bb = self.mlp_style_gamma.bias, self.mlp_style_beta.bias
        return _interleave_gamma_beta(wg, wb), gb.contiguous(), bb.contiguous()

    def combined_cached(self):
        """combined_weight() as plain fp32 tensors, cached while the parameters are unchanged
        (inference with folded style: the per-image weights are rebuilt from it every call)."""
        key = self._cache_key('raw')
        cached = getattr(self, '_raw_cache', None)
        if cached is not None and cached[0] == key:
            return cached[1]
        with torch.no_grad():
            val = tuple(t.detach() for t in self.combined_weight())
        self._raw_cache = (key, val)
        return val

    def _cache_key(self, want_lo):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + (want_lo,)

    def prepared(self, want_lo):
        """Prepared (fp16 split, scaled) modulation weight + biases. Cached while the parameters
        are unchanged (inference); rebuilt every call when they carry gradients."""
        key = self._cache_key(want_lo)
        cached = getattr(self, '_prep_cache', None)
        if cached is not None and cached[0] == key and not torch.is_grad_enabled():
            return cached[1]
        with torch.no_grad():
            w, gb, bb = self.combined_weight()
            val = (ops.prep_conv_weight(w, want_lo=want_lo), gb, bb)
        self._prep_cache = (key, val)
        return val

    def eval_affine(self):
        bn = self.param_free_norm
        return ops.bn_eval_affine(bn.running_mean, bn.running_var, BN_EPS)

    def forward(self, x, segmap, style=None):
        raise RuntimeError('%s is driven through SPADEResnetBlock on the B200 path; it has no '
                           'standalone ATen forward' % type(self).__name__)


class SPADE(_CondNormBase):
    """normalization.py:71-120."""
    kind = 'spade'

    def __init__(self, config_text, norm_nc, label_nc, opt):
        super().__init__()
        assert config_text.startswith('spade') or config_text.startswith('latesean')
        pattern = r'latesean(\D+)(\d)x\d' if config_text.startswith('latesean') else r'spade(\D+)(\d)x\d'
        self.opt = opt
        self.nc = label_nc
        self.param_free_norm = _param_free_norm(config_text, pattern, norm_nc)
        self.mlp_shared = nn.Sequential(nn.Conv2d(label_nc, NHIDDEN, kernel_size=3, padding=1), nn.ReLU())
        self.mlp_gamma = nn.Conv2d(NHIDDEN, norm_nc, kernel_size=3, padding=1)
        self.mlp_beta = nn.Conv2d(NHIDDEN, norm_nc, kernel_size=3, padding=1)


class SEAN_Block(_CondNormBase):
    """normalization.py:123-213."""
    kind = 'sean'

    def __init__(self, config_text, norm_nc, label_nc, opt):
        super().__init__()
        assert 'sean' in config_text
        self.opt = opt
        self.efficient = opt.efficient
        self.nc = label_nc
        self.style_size = opt.regional_style_size
        self.param_free_norm = _param_free_norm(config_text, r'sean(\D+)(\d)x\d', norm_nc)
        self.mlp_shared = nn.Sequential(nn.Conv2d(label_nc, NHIDDEN, kernel_size=3, padding=1), nn.ReLU())
        self.mlp_gamma = nn.Conv2d(NHIDDEN, norm_nc, kernel_size=3, padding=1)
        self.mlp_beta = nn.Conv2d(NHIDDEN, norm_nc, kernel_size=3, padding=1)
        self.style_conv = nn.Conv1d(19, 19, kernel_size=1)  # unused by the reference forward too
        self.mlp_style_gamma = nn.Conv2d(self.style_size, norm_nc, kernel_size=3, padding=1)
        self.mlp_style_beta = nn.Conv2d(self.style_size, norm_nc, kernel_size=3, padding=1)
        self.alpha_beta = nn.Parameter(torch.rand(1), requires_grad=True)
        self.alpha_gamma = nn.Parameter(torch.rand(1), requires_grad=True)
        self.mp = opt.model_parallel_mode


class PureSEAN_Block(_CondNormBase):
    """normalization.py:216-286."""
    kind = 'puresean'

    def __init__(self, config_text, norm_nc, label_nc, opt):
        super().__init__()
        assert 'sean' in config_text
        self.opt = opt
        self.efficient = opt.efficient
        self.nc = label_nc
        self.style_size = opt.regional_style_size
        self.param_free_norm = _param_free_norm(config_text, r'sean(\D+)(\d)x\d', norm_nc)
        self.mlp_shared = nn.Sequential(nn.Conv2d(label_nc, NHIDDEN, kernel_size=3, padding=1), nn.ReLU())
        self.style_conv = nn.Conv1d(19, 19, kernel_size=1)
        self.mlp_style_gamma = nn.Conv2d(self.style_size, norm_nc, kernel_size=3, padding=1)
        self.mlp_style_beta = nn.Conv2d(self.style_size, norm_nc, kernel_size=3, padding=1)
        self.mp = opt.model_parallel_mode


class NoiseInjection(nn.Module):
    """normalization.py:289-304. Holds the per-channel weight; the add itself is fused into the
    K1 / K2 epilogues. ``sample`` draws the noise tensor (NHWC) - tests replace it to replay the
    reference's noise."""

    def __init__(self, n_channels):
        super().__init__()
        self.n_channels = n_channels
        self.weight = nn.Parameter(torch.zeros(self.n_channels), requires_grad=True)

    def sample(self, B, H, W):
        """One draw of the [B,H,W,C] N(0,1) tensor of normalization.py:301, as a seed: the kernels
        regenerate its elements (counter-based Philox) instead of reading a materialised tensor.
        The seed comes from torch's CPU generator, so `torch.manual_seed` controls it and ranks
        seeded differently draw different noise.  DSEE_NOISE_TENSORS=1 materialises torch.randn
        tensors instead (what the parity tests replace with the oracle's noise)."""
        from ...config import config
        if config.noise_tensors:
            return torch.randn((B, H, W, self.n_channels), dtype=torch.float32,
                               device=self.weight.device)
        return ops.NoiseSeed(int(torch.randint(1, 2 ** 62, (1,)).item()))

    def forward(self, tensor, noise=None):
        raise RuntimeError('NoiseInjection is fused into the conv epilogues on the B200 path')
