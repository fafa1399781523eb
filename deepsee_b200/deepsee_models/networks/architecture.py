"""SPADEResnetBlock on the fused B200 kernels (reference: deepsee_models/networks/architecture.py:24-147).

One block = 2 x (K1 modulate -> K2 conv):

    x (fp32 NHWC, possibly at half resolution: the nn.Upsample in front of the block is folded
       into the kernels' input indexing and never materialised)
      -> K1[norm_0]: gamma/beta GEMM, epilogue BN + modulation + LeakyReLU -> fp16 planes a0
      -> K2[conv_0]: implicit GEMM + bias (+ noise_middle) -> dx fp32, BN partial sums
      -> K1[norm_1] -> a1
      -> K2[conv_1]: + bias + shortcut (x through the folded upsample, + noise_in/noise_skip)
         -> block output fp32 NHWC and its BN partial sums for the next block's norm_0

The block is ONE autograd node (`_ResBlockFn`): its backward replays the chain with the library's
backward kernels (dgrad = the same implicit GEMM on the gradient planes with the transposed filter,
wgrad = pixel-reduction GEMM, K1 backward, batch-norm backward with the folded upsample transposed)
and hands PyTorch the gradients of the *effective* tensors (spectral-normalised conv weights, the
alpha-blended modulation weight, the mlp_shared table), so spectral norm, the alpha blend and the
optimizers stay ordinary PyTorch autograd.

fin == fout everywhere in DeepSEESR, so the learned shortcut (conv_s / norm_s) is never built
(architecture.py:30,36); asking for it raises.
"""
import torch
import torch.nn as nn
import torch.nn.utils.spectral_norm as spectral_norm

from ... import ops, parallel
from ...config import config
from .normalization import (SPADE, SEAN_Block, PureSEAN_Block, NoiseInjection, effective_weight,
                            fold_style_weight, BN_EPS)

BN_MOMENTUM = 0.1


_SIDE = {}


class _SideStream:
    """Weight-gradient GEMMs are leaves of the backward graph: nothing in the block's backward chain
    reads them.  They are issued on a second CUDA stream so the HBM-bound kernels of the chain
    (K1 backward, batch-norm backward, gradient splitting) run underneath them instead of between
    tensor-core kernels.  Tensors the side stream reads are kept alive in ``keep`` until the main
    stream has waited for it (`join`)."""

    def __init__(self):
        dev = torch.cuda.current_device()
        if dev not in _SIDE:
            _SIDE[dev] = torch.cuda.Stream(device=dev)
        self.side = _SIDE[dev] if config.overlap_wgrad else None
        self.keep = []

    def run(self, fn, *tensors):
        if self.side is None:
            return fn()
        self.keep.extend(tensors)
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            return fn()

    def join(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self.keep.clear()


class _NormState:
    """What K1's backward needs from one conditional-norm layer's forward."""
    __slots__ = ("srcs", "meta", "Wm", "Ws", "gb", "sc", "sh", "inv_count", "g", "sync")


def _norm_forward(blk, norm, pre, Wm, Ws, gb, bb, tab, tb, gctx, style, x, x_ups, H, W, part, count,
                  ucount, noise, noise_w, passes, want_lo, save_g, out_lo=None, out_f8=False):
    """BN affine + sources + K1 -> (activation planes, _NormState).  ``passes`` / ``want_lo`` are
    K1's own (its GEMM operands); ``out_lo`` / ``out_f8``: whether the consumer of the activation
    planes (the main conv) runs 3 passes and needs their lo plane (default: same as want_lo) / runs
    the fp8 correction (passes == 2) and needs their e5m2 planes."""
    if out_lo is None:
        out_lo = want_lo
    st = _NormState()
    bn = norm.param_free_norm
    st.sync = parallel.is_dist() and config.sync_bn_for(blk.opt.norm_G)
    if blk.training:
        if st.sync:
            # global-batch statistics (the reference's Sync-BN, batchnorm.py:80-93): all-reduce the
            # per-channel (sum, sum of squares) before they become mean / variance
            part = parallel.allreduce_sum_(ops.reduce_partials(part)).t().contiguous().unsqueeze(0)
            count, ucount = count * parallel.world_size(), ucount * parallel.world_size()
        st.sc, st.sh, _, _ = ops.bn_finalize(part, count, BN_EPS, BN_MOMENTUM, bn.running_mean,
                                             bn.running_var, unbias_count=ucount)
        bn.num_batches_tracked.add_(1)
        st.inv_count = 1.0 / float(ucount)
    else:
        st.sc, st.sh = norm.eval_affine()
        st.inv_count = 0.0
    sub = Wm is not None and Wm.dim() == 5   # collapsed filters [4, 2C, Cin, 2, 2]: sub-pixel form
    if pre is not None:
        pw = pre
    elif sub:
        pw = ops.prep_subpixel_weight(Wm, want_lo=want_lo)
    elif Ws is not None:  # folded style: shared actv columns + per-image one-hot columns
        pw = ops.prep_mod_weight_batched(Wm.contiguous(), Ws.contiguous(), want_lo=want_lo)
    else:
        pw = ops.prep_conv_weight(Wm.contiguous(), want_lo=want_lo)
    st.srcs, st.meta = norm.build_sources(gctx, style, H, W, want_lo, table=tab, bias=tb,
                                          folded=Ws is not None, sub=sub)
    st.Wm, st.Ws, st.gb = Wm, Ws, gb
    r = ops.spade_modulate(st.srcs, pw, x, x_ups, st.sc, st.sh, gb, bb, noise=noise, noise_w=noise_w,
                           passes=passes, want_lo=out_lo, save_g=save_g, want_f8=out_f8, subpixel=sub)
    a, st.g = r if save_g else (r, None)
    return a, st


def _norm_backward(norm, st, dt, dt_amax, x, x_ups, noise, noise_w, L, passes, want_lo, ss):
    """K1 backward for one layer -> (dxhat, sums[4,C], dWm, dtab, dtb, dstyle, dWs)."""
    C = x.shape[3]
    Wm = st.Wm
    cin = Wm.shape[1]
    if st.g is not None:
        # G = gamma + gamma_bias was saved by the forward kernel: one streaming pass, no GEMM
        dxhat, dgb, sums = ops.spade_modulate_bwd_saved(st.g, x, x_ups, st.sc, st.sh, dt, dt_amax,
                                                        noise=noise, noise_w=noise_w, want_lo=want_lo)
        st.g = None
    else:
        w_gamma = Wm.view(C // 128, 2, 128, cin, 3, 3)[:, 0].reshape(C, cin, 3, 3).contiguous()
        pwg = ops.prep_conv_weight(w_gamma, want_lo=want_lo)
        dxhat, dgb, sums = ops.spade_modulate_bwd(st.srcs, pwg, x, x_ups, st.sc, st.sh, st.gb, dt,
                                                  dt_amax, noise=noise, noise_w=noise_w, passes=passes,
                                                  want_lo=want_lo)
    return (dxhat, sums) + _norm_backward_tail(st, dgb, L, passes, want_lo, ss)


def _conv_and_norm_backward(st, g, W, a, x, x_ups, noise, noise_w, L, p1, p2, ss):
    """Backward-data of a main conv (gradient planes ``g`` of its output, weight ``W``, input
    activation planes ``a``; ``p2`` passes) followed by K1's backward of the norm layer that produced
    ``a`` (``p1`` passes for its GEMMs) -> (dxhat, sums[4,C], dWm, dtab, dtb, dstyle, dWs).  With the saved G
    planes both run as ONE kernel (ops.dgrad_modulate_bwd: dt stays in TMEM / registers); otherwise
    dgrad -> dt -> K1 backward."""
    pwT = ops.prep_conv_weight(W.contiguous(), want_lo=p2 == 3, transpose=True)
    if st.g is not None and config.fuse_dgrad_modbwd:
        dxhat, dgb, sums = ops.dgrad_modulate_bwd(g, pwT, a.hi, st.g, x, x_ups, st.sc, st.sh, noise=noise,
                                                  noise_w=noise_w, passes=p2, want_lo=p1 == 3)
        st.g = None
        return (dxhat, sums) + _norm_backward_tail(st, dgb, L, p1, p1 == 3, ss)
    dt, amax = ops.conv3x3([g], pwT, None, passes=p2, act_mask=a.hi, want_amax=True, tag="dgrad")
    return _norm_backward(None, st, dt, amax, x, x_ups, noise, noise_w, L, p1, p1 == 3, ss)


def _norm_backward_tail(st, dgb, L, passes, want_lo, ss):
    """From the [dG | dB] gradient planes: modulation-weight gradient, gradient of the sources
    (mlp_shared table / bias, style matrix) -> (dWm, dtab, dtb, dstyle, dWs)."""
    Wm = st.Wm
    dWs = None
    meta = st.meta
    if meta.get('sub'):
        # sub-pixel form: Wm holds the collapsed filters [4, 2C, Cin, 2, 2]; both backward GEMMs run at
        # 4/9 of the 3x3 work and the source gradient comes out at the sources' (half) resolution
        dWm = ss.run(lambda: ops.subpixel_wgrad(dgb, st.srcs, passes=passes), dgb, *st.srcs)
        dsrc, dsrc_amax = ops.subpixel_dgrad(dgb, Wm, passes=passes, want_lo=want_lo)
        del dgb
        dtab = dtb = None
        coff = 0
        for src in st.srcs:          # every source of a layer above max_fm_size is the actv tensor
            onehot = meta['ctx'].onehot_at(*meta['fm'])
            t, b = ops.shared_mlp_bwd_tc(dsrc, dsrc_amax, coff, meta['actv'].hi, meta['labels'], onehot, 0, L,
                                         passes=passes, side=ss)
            dtab = t if dtab is None else ss.run(lambda: dtab + t)
            dtb = b if dtb is None else dtb + b
            coff += src.hi.shape[3]
        return dWm, dtab, dtb, None, None
    if st.Ws is not None:
        # folded style: one weight gradient per image; the actv columns are shared (sum over images),
        # the label columns are the gradient of Ws.  Wm holds the actv columns only, so the
        # backward-data GEMM below produces d(actv) and nothing for the (constant) one-hot planes.
        nh = Wm.shape[1]
        dfull = ss.run(lambda: ops.conv3x3_wgrad_per_image(dgb, st.srcs, passes=passes), dgb, *st.srcs)
        dWm = ss.run(lambda: dfull[:, :, :nh].sum(0))
        dWs = ss.run(lambda: dfull[:, :, nh:nh + st.Ws.shape[2]].contiguous())
    else:
        dWm = ss.run(lambda: ops.conv3x3_wgrad_multi(dgb, st.srcs, passes=passes), dgb, *st.srcs)
    pwT = ops.prep_conv_weight(Wm.contiguous(), want_lo=want_lo, transpose=True)
    dsrc, dsrc_amax = ops.conv3x3([dgb], pwT, None, passes=passes, want_amax=True, tag="dgrad_mod")
    del dgb
    dtab = dtb = dstyle = None
    coff = 0
    for src, kind in zip(st.srcs, meta['kinds']):
        d = src.hi.shape[3]
        if kind == 'onehot':
            continue
        if kind == 'actv':
            onehot = meta['ctx'].onehot_at(*meta['fm'])
            t, b = ops.shared_mlp_bwd_tc(dsrc, dsrc_amax, coff, meta['actv'].hi, meta['labels'], onehot,
                                         meta['ups'], L, passes=passes, side=ss)
            dtab = t if dtab is None else ss.run(lambda: dtab + t)  # t lives on the side stream
            dtb = b if dtb is None else dtb + b
        else:
            g = ops.style_gather_bwd(dsrc, coff, meta['labels'], L, d)
            dstyle = g if dstyle is None else dstyle + g
        coff += d
    return dWm, dtab, dtb, dstyle, dWs


def _sync_bwd_sums(nsums, st):
    """Sync-BN backward: batch-norm's two reductions (sum dxhat, sum dxhat*xhat) run over the global
    batch.  Every rank differentiates its OWN mean loss and the gradient buckets are averaged
    afterwards, so the unscaled local sums are the right summands.  Rows 2-3 (parameter gradients)
    stay local."""
    sync = getattr(st, 'sync', None)
    if sync is None:
        sync = parallel.is_dist() and config.sync_bn is True
    if not sync or st.inv_count == 0.0:
        return nsums
    head = parallel.allreduce_sum_(nsums[:2].clone())
    return torch.cat([head, nsums[2:]], 0)


class _ResBlockFn(torch.autograd.Function):
    """forward / backward of one SPADEResnetBlock (architecture.py:75-147) on the C-ABI kernels."""

    @staticmethod
    def forward(ctx, blk, gctx, ups, stats_in, noises, pre, grad_on, x, style, W0, b0, W1, b1, Wm0, gb0,
                bb0, tab0, tb0, Wm1, gb1, bb1, tab1, tb1, nw_in, nw_skip, nw_mid, Ws0=None, Ws1=None):
        B, Hx, Wx, C = x.shape
        H, W = Hx << ups, Wx << ups
        # operand passes of this block's gamma/beta GEMMs (p1), main convs forward (p2) and the main
        # convs' backward GEMMs (p2b): config.passes_for
        S = gctx.labels_full.shape[1]
        p1, p2 = config.passes_for('k1', H, S), config.passes_for('k2', H, S)
        p2b = config.passes_for('k2b', H, S)
        if p2b == 3 and p2 != 3:
            p2b = 1  # the activation planes were produced without their lo half
        p1b = config.passes_for('k1b', H, S)
        if p1b == 3 and p1 != 3:
            p1b = 1  # same for K1's sources
        training = blk.training
        n_in, n_skip, n_mid = noises if noises is not None else (None, None, None)
        noisy = noises is not None
        pre = pre or {}
        # grad mode is invisible inside Function.forward (always off) and needs_input_grad ignores
        # torch.no_grad(), so the caller passes it: under no_grad (the discriminator step's generator
        # forward) nothing is kept for a backward pass and K1 does not write its G planes
        need_bwd = grad_on and any(ctx.needs_input_grad)
        save_g = need_bwd and config.save_gamma

        # ---- norm_0 + actvn -------------------------------------------------------------------
        part = None
        count = ucount = 0
        if training:
            if stats_in is not None and not noisy:
                # statistics of a nearest-upsampled tensor = statistics of its source
                part, count, ucount = stats_in, B * Hx * Wx, B * H * W
            else:
                part = ops.bn_stats(x, ups, n_in, nw_in if noisy else None)
                count = ucount = B * H * W
        a0, st0 = _norm_forward(blk, blk.norm_0, pre.get('pwm0'), Wm0, Ws0, gb0, bb0, tab0, tb0, gctx,
                                style, x, ups, H, W, part, count, ucount, n_in,
                                nw_in if noisy else None, p1, p1 == 3, save_g, out_lo=p2 == 3, out_f8=p2 == 2)
        # ---- conv_0 (+ noise_middle) ----------------------------------------------------------
        pw0 = pre.get('pw0') or ops.prep_conv_weight(W0.contiguous(), want_lo=p2 == 3, want_f8=p2 == 2)
        r = ops.conv3x3([a0], pw0, b0, noises=[(n_mid, nw_mid)] if noisy else (), passes=p2,
                        want_stats=training)
        dx1, part1 = r if training else (r, None)
        # ---- norm_1 + actvn -------------------------------------------------------------------
        a1, st1 = _norm_forward(blk, blk.norm_1, pre.get('pwm1'), Wm1, Ws1, gb1, bb1, tab1, tb1, gctx,
                                style, dx1, 0, H, W, part1, B * H * W, B * H * W, None, None, p1,
                                p1 == 3, save_g, out_lo=p2 == 3, out_f8=p2 == 2)
        # ---- conv_1 + shortcut ------------------------------------------------------------------
        pw1 = pre.get('pw1') or ops.prep_conv_weight(W1.contiguous(), want_lo=p2 == 3, want_f8=p2 == 2)
        # the block in front of the image head also writes leaky_relu(out) as fp16 planes (sr.py:94)
        act16 = getattr(gctx, 'head_act_block', None) is blk
        r = ops.conv3x3([a1], pw1, b1, residual=x, res_ups=ups,
                        noises=[(n_in, nw_in), (n_skip, nw_skip)] if noisy else (), passes=p2,
                        want_stats=training, act16=act16)
        out, stats = r if training else (r, None)
        if act16:
            gctx.head_act = out._dsee_act16

        if need_bwd:
            # (the e5m2 planes of a passes == 2 forward are not needed again)
            a0, a1 = ops.SplitPlanes(a0.hi, a0.lo), ops.SplitPlanes(a1.hi, a1.lo)
            ctx.s = dict(blk=blk, ups=ups, noises=noises, x=x, W0=W0, W1=W1, a0=a0, a1=a1, dx1=dx1,
                         st0=st0, st1=st1, nw_in=nw_in, p1=p1b, p2=p2b)
        if stats is None:
            stats = x.new_zeros(1)
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, dout, _dstats):
        s = ctx.s
        blk, ups, p1, p2 = s['blk'], s['ups'], s['p1'], s['p2']
        want_lo = p2 == 3
        n_in, n_skip, n_mid = s['noises'] if s['noises'] is not None else (None, None, None)
        noisy = s['noises'] is not None
        L = blk.opt.semantic_nc
        dout = dout.contiguous()
        x, dx1, a0, a1, st0, st1 = s['x'], s['dx1'], s['a0'], s['a1'], s['st0'], s['st1']

        # ---- conv_1 + shortcut: out = conv(a1, W1) + b1 + x_up + w_in*n_in + w_skip*n_skip ------
        # (d noise_in.weight of the shortcut rides along in norm_0's bn_bwd below, which regenerates
        # n_in anyway; max|dout| comes with the tensor when the next block's bn_bwd produced it)
        tag = getattr(dout, "_dsee_amax", None)
        # valid only for the very tensor bn_bwd wrote: an in-place accumulation by the autograd engine
        # (a block output with a second consumer) bumps the version counter and voids it
        amax_in = tag[0] if tag is not None and tag[1] == dout._version else None
        g1, sums = ops.grad_prep(dout, n_skip, None, want_lo=want_lo, amax=amax_in)
        db1 = sums[0]
        dnw_skip = sums[1] if noisy else None
        ss = _SideStream()
        dW1 = ss.run(lambda: ops.conv3x3_wgrad(g1, a1, passes=p2), g1, a1)
        # ---- backward-data of conv_1 + norm_1 --------------------------------------------------
        dxhat, nsums, dWm1, dtab1, dtb1, dstyle1, dWs1 = _conv_and_norm_backward(
            st1, g1, s['W1'], a1, dx1, 0, None, None, L, p1, p2, ss)
        del g1
        dgb1, dbb1 = nsums[2], nsums[3]
        nsums = _sync_bwd_sums(nsums, st1)
        ddx1, _, amax1 = ops.bn_bwd(dxhat, dx1, 0, st1.sc, st1.sh, nsums, st1.inv_count, want_amax=True)
        del dxhat
        # ---- conv_0: dx1 = conv(a0, W0) + b0 + w_mid*n_mid --------------------------------------
        g0, sums = ops.grad_prep(ddx1, n_mid, None, want_lo=want_lo, amax=amax1)
        del ddx1
        db0 = sums[0]
        dnw_mid = sums[1] if noisy else None
        dW0 = ss.run(lambda: ops.conv3x3_wgrad(g0, a0, passes=p2), g0, a0)
        # ---- backward-data of conv_0 + norm_0 (reads x through the folded upsample, + noise_in) -
        nw_in = s['nw_in'] if noisy else None
        dxhat, nsums, dWm0, dtab0, dtb0, dstyle0, dWs0 = _conv_and_norm_backward(
            st0, g0, s['W0'], a0, x, ups, n_in, nw_in, L, p1, p2, ss)
        del g0
        dgb0, dbb0 = nsums[2], nsums[3]
        nsums = _sync_bwd_sums(nsums, st0)
        dx, dnw_in, dx_amax = ops.bn_bwd(dxhat, x, ups, st0.sc, st0.sh, nsums, st0.inv_count, noise=n_in,
                                         noise_w=nw_in, dskip=dout, noise_grad_with_skip=noisy,
                                         want_amax=True)
        # the previous block's backward receives this very tensor as its `dout` (a block output has one
        # consumer); if autograd hands over a different tensor the attribute is simply absent
        dx._dsee_amax = (dx_amax, dx._version)
        dstyle = dstyle0
        if dstyle1 is not None:
            dstyle = dstyle1 if dstyle is None else dstyle + dstyle1
        ss.join()
        ctx.s = None
        return (None, None, None, None, None, None, None, dx, dstyle, dW0, db0, dW1, db1, dWm0, dgb0, dbb0,
                dtab0, dtb0, dWm1, dgb1, dbb1, dtab1, dtb1, dnw_in, dnw_skip, dnw_mid, dWs0, dWs1)


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

    # -- prepared main-conv weights (inference cache) -----------------------------------------------
    def _prepared_conv(self, conv, w, name, want_lo, want_f8=False):
        src = [getattr(conv, n) for n in ('weight_orig', 'weight_u', 'weight_v') if hasattr(conv, n)]
        if not src:
            src = [conv.weight]
        key = tuple((t.data_ptr(), t._version) for t in src) + (want_lo, want_f8)
        hit = self._wcache.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        pw = ops.prep_conv_weight(w.detach().contiguous(), want_lo=want_lo, want_f8=want_f8)
        self._wcache[name] = (key, pw)
        return pw

    def forward_nhwc(self, x, ctx, ups=0, stats_in=None):
        """x fp32 NHWC [B, H>>ups, W>>ups, C] -> (out fp32 NHWC [B,H,W,C], BN partial sums of out or
        None). ``stats_in``: BN partial sums of x from the producing kernel (training only)."""
        B, Hx, Wx, C = x.shape
        H, W = Hx << ups, Wx << ups
        S = ctx.labels_full.shape[1]
        lo1, lo2 = config.passes_for('k1', H, S) == 3, config.passes_for('k2', H, S) == 3
        f82 = config.passes_for('k2', H, S) == 2
        noises = None
        nw = (None, None, None)
        if self.add_noise:
            noises = (self.noise_in.sample(B, H, W), self.noise_skip.sample(B, H, W),
                      self.noise_middle.sample(B, H, W))
            nw = (self.noise_in.weight, self.noise_skip.weight, self.noise_middle.weight)
        # spectral norm: W_orig / sigma (one power iteration when training), an autograd tensor
        W0, W1 = effective_weight(self.conv_0), effective_weight(self.conv_1)
        cached = not self.training and not torch.is_grad_enabled()
        # SEAN layers at or below max_fm_size: the style branch runs as per-image weights over the
        # one-hot label planes (config.fold_style), rebuilt from the style matrix on every call
        fold = [n.folds_style(H, W) and ctx.style is not None for n in (self.norm_0, self.norm_1)]
        # layers above max_fm_size: sub-pixel form (collapsed 2x2 filters over the half-resolution actv)
        subp = [n.uses_subpixel(H, W) for n in (self.norm_0, self.norm_1)]
        Wm, gbs, bbs, pwm = [None, None], [None, None], [None, None], [None, None]
        for i, n in enumerate((self.norm_0, self.norm_1)):
            if cached and not fold[i] and not subp[i]:
                pwm[i], gbs[i], bbs[i] = n.prepared(lo1)
            elif cached:
                Wm[i], gbs[i], bbs[i] = n.combined_cached()
            else:
                Wm[i], gbs[i], bbs[i] = n.combined_weight()
        Ws = [None, None]
        for i, n in enumerate((self.norm_0, self.norm_1)):
            if fold[i]:
                Wm[i], Ws[i] = fold_style_weight(Wm[i], ctx.style)
            elif subp[i]:
                Wm[i] = ops.collapse_subpixel(Wm[i])
        (Wm0, Wm1), (gb0, gb1), (bb0, bb1) = Wm, gbs, bbs
        pre = None
        if cached:
            pre = {'pwm0': pwm[0], 'pwm1': pwm[1],
                   'pw0': self._prepared_conv(self.conv_0, W0, 'conv_0', lo2, f82),
                   'pw1': self._prepared_conv(self.conv_1, W1, 'conv_1', lo2, f82)}
        tab0, tb0 = self.norm_0.table_and_bias()
        tab1, tb1 = self.norm_1.table_and_bias()
        out, stats = _ResBlockFn.apply(self, ctx, ups, stats_in, noises, pre, torch.is_grad_enabled(), x,
                                       ctx.style, W0,
                                       self.conv_0.bias, W1, self.conv_1.bias, Wm0, gb0, bb0, tab0,
                                       tb0, Wm1, gb1, bb1, tab1, tb1, *nw, Ws[0], Ws[1])
        return out, (stats if self.training else None)

    def forward(self, x, seg, style=None, split_location=-1):
        raise RuntimeError('SPADEResnetBlock is driven by DeepSEESR.forward on the B200 path '
                           '(NHWC activations, fused kernels); use forward_nhwc')


# VGG19 feature stack of the perceptual loss (reference: architecture.py:151-181).  The reference
# slices torchvision.models.vgg19(pretrained=True).features; the module tree below has the same
# children names (slice1.0, slice1.1, slice2.2, ... slice5.29), so a state_dict of that model - or
# torchvision's `features.N.*` keys - loads unchanged.  The forward runs on the deepsee_b200 kernels:
# NHWC activations, 3x3 convs with the ReLU fused on the tcgen05 implicit-GEMM kernel (the 3-channel
# first layer on the direct kernel), 2x2 max pooling as its own kernel.  The weights are frozen
# (architecture.py:170-172): the backward pass is backward-data only.
_VGG19_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512]
_VGG19_SLICES = ((0, 2), (2, 7), (7, 12), (12, 21), (21, 30))


class VGG19(nn.Module):
    def __init__(self, requires_grad=False, weights=None):
        """weights: a state_dict (this module's keys or torchvision's `features.N.weight/bias`), a path
        to one, or None = look for torchvision's vgg19 checkpoint in $DSEE_VGG19_WEIGHTS and the torch
        hub cache; there is no network here, so a missing checkpoint raises."""
        super().__init__()
        layers, cin = [], 3
        for v in _VGG19_CFG:
            if v == 'M':
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        for si, (a, b) in enumerate(_VGG19_SLICES):
            seq = nn.Sequential()
            for i in range(a, b):
                seq.add_module(str(i), layers[i])
            setattr(self, 'slice%d' % (si + 1), seq)
        self._load(weights)
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def _load(self, weights):
        import os
        if weights is None:
            cands = [os.environ.get('DSEE_VGG19_WEIGHTS'),
                     os.path.join(torch.hub.get_dir(), 'checkpoints', 'vgg19-dcbb9e9d.pth')]
            weights = next((c for c in cands if c and os.path.exists(c)), None)
            if weights is None:
                raise RuntimeError(
                    'VGG19: no pretrained weights found (set DSEE_VGG19_WEIGHTS to torchvision\'s '
                    'vgg19-dcbb9e9d.pth, or pass weights=...); the perceptual loss cannot be downloaded here - '
                    'train with --no_vgg_loss or provide the file')
        if isinstance(weights, str):
            weights = torch.load(weights, map_location='cpu')
        own = {}
        for si, (a, b) in enumerate(_VGG19_SLICES):
            for i in range(a, b):
                for nm in ('weight', 'bias'):
                    for key in ('slice%d.%d.%s' % (si + 1, i, nm), 'features.%d.%s' % (i, nm)):
                        if key in weights:
                            own['slice%d.%d.%s' % (si + 1, i, nm)] = weights[key]
        self.load_state_dict(own, strict=True)

    def forward(self, X):
        """X: fp32 NCHW [B,3,H,W] -> [h_relu1 ... h_relu5] as NCHW views of the NHWC feature maps."""
        if not X.is_cuda:
            raise RuntimeError('VGG19 (B200 path) needs CUDA tensors; there is no CPU fallback')
        # NHWC with the 3 colour channels padded to 4 (the direct conv kernel's input-gradient path
        # works on channel quads); plain torch layout ops so autograd reaches the fake image
        x = torch.nn.functional.pad(X.float().permute(0, 2, 3, 1), (0, 1)).contiguous()
        outs = []
        for si in range(5):
            for m in getattr(self, 'slice%d' % (si + 1)):
                if isinstance(m, nn.Conv2d):
                    x = ops.conv_layer(x, m.weight, m.bias, 1, 1, lrelu=2)   # conv + ReLU fused
                elif isinstance(m, nn.MaxPool2d):
                    x = ops.MaxPool2Fn.apply(x)
                # nn.ReLU: fused into the conv above
            outs.append(x.permute(0, 3, 1, 2))
        return outs
