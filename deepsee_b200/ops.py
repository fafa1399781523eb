"""Tensor-level wrappers over the C ABI: PyTorch tensors in, PyTorch tensors out.

PyTorch is used here only for device memory (the caching allocator owns every buffer) and for
the current CUDA stream; all arithmetic of the hot path happens inside the shared library.
Every function raises if the library is missing or the device is not a B200 - no fallback.
"""
import ctypes as C
from collections import namedtuple

import torch

from . import _lib

# fp16 NHWC [B,H,W,C]; value = hi + lo.  f8: None, or the (lo * 2^8, value) e5m2 byte planes a
# passes == 2 consumer (fp8 correction GEMM) reads next to hi
SplitPlanes = namedtuple("SplitPlanes", ["hi", "lo", "f8"], defaults=[None])
# gradient operand: fp16 NHWC planes of g * 2^e plus the device scalar 2^-e
GradPlanes = namedtuple("GradPlanes", ["hi", "lo", "inv_scale"])
# f8: None, or the e4m3 companion [N][9][2][C] of the planes (dsee_prep_conv_weight_f8)
# batched: the planes hold one [n_total, K] matrix per image (prep_mod_weight_batched)
PreparedWeight = namedtuple("PreparedWeight", ["hi", "lo", "inv_scale", "n_total", "cin", "f8", "batched"],
                            defaults=[None, False])


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """The caller's current CUDA stream as a raw handle.  torch.cuda.current_stream() costs ~16 us per
    call (device-index resolution through several Python layers), which at ~1200 launches per training
    iteration starved the GPU during the small-kernel phases; the two C entry points below return the
    same handle in well under a microsecond."""
    if _raw_stream is not None and _raw_device is not None:
        return C.c_void_p(_raw_stream(_raw_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class KernelTimer:
    """Optional CUDA-event bracket around the tensor-core launches (bench.py's live roofline):
    records (tag, algorithmic FLOPs, start event, end event) on the launching stream."""
    active = None

    def __init__(self):
        self.records = []

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *a):
        KernelTimer.active = None

    def summary(self):
        """-> {tag: (launches, total_ms, total_flops)} (call after a device synchronize)."""
        out = {}
        for tag, flops, e0, e1 in self.records:
            n, ms, fl = out.get(tag, (0, 0.0, 0.0))
            out[tag] = (n + 1, ms + e0.elapsed_time(e1), fl + flops)
        return out


def _timed(tag, flops, fn):
    t = KernelTimer.active
    if t is None:
        return fn()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    t.records.append((tag, flops, e0, e1))
    return r


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class NoiseSeed(int):
    """A NoiseInjection tensor that is never materialised: the kernels regenerate its elements from
    this 64-bit seed (counter-based Philox + Box-Muller, see dsee_noise_fill)."""


def _noise(n):
    """noise argument (None | fp32 NHWC tensor | NoiseSeed) -> (pointer, seed) for the C ABI."""
    if n is None:
        return C.c_void_p(0), 0
    if isinstance(n, NoiseSeed):
        return C.c_void_p(0), int(n)
    return _p(n), 0


def noise_epoch_advance():
    """Bumps the device-side noise epoch (see dsee_noise_epoch_advance): seeds baked into a captured
    CUDA graph then stand for a fresh noise tensor on every replay."""
    _lib.check(_lib.load().dsee_noise_epoch_advance(_stream()))


def conv_pair_mode(on=None):
    """Switches dsee_conv3x3_fwd's CTA-pair form (tcgen05 cta_group::2) on / off for the process; returns
    the previous mode.  on=None only queries."""
    return bool(_lib.load().dsee_conv_pair_mode(-1 if on is None else int(bool(on))))


def noise_fill(seed, shape, device="cuda"):
    """Materialises the noise tensor a NoiseSeed stands for (tests)."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    _lib.check(_lib.load().dsee_noise_fill(int(seed), _p(out), out.numel(), _stream()))
    return out


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not isinstance(t, NoiseSeed):
            if not t.is_cuda:
                raise RuntimeError("deepsee_b200 ops need CUDA tensors (no CPU fallback exists)")
            if not t.is_contiguous():
                raise RuntimeError("deepsee_b200 ops need contiguous tensors")


# ---------------------------------------------------------------------------------------------
# label maps
# ---------------------------------------------------------------------------------------------
def onehot_from_labels(label, num_classes):
    """Preprocessor.preprocess_label (data/preprocessor.py:35-41). label int64 [B,1,H,W]."""
    _chk_cuda(label)
    assert label.dtype == torch.int64 and label.dim() == 4 and label.size(1) == 1
    B, _, H, W = label.shape
    out = torch.empty((B, num_classes, H, W), dtype=torch.float32, device=label.device)
    bad = torch.zeros(1, dtype=torch.int32, device=label.device)
    _lib.check(_lib.load().dsee_onehot_from_labels(_p(label), _p(out), B, num_classes, H, W, _p(bad),
                                                   _stream()))
    return out, bad


def labels_u8(label, num_classes):
    """int64 label map [B,1,H,W] (or [B,H,W]) -> (uint8 [B,H,W], device flag: 1 if out of range)."""
    _chk_cuda(label)
    assert label.dtype == torch.int64
    shape = label.shape if label.dim() == 3 else (label.shape[0],) + tuple(label.shape[2:])
    assert label.dim() == 3 or label.size(1) == 1
    out = torch.empty(shape, dtype=torch.uint8, device=label.device)
    bad = torch.zeros(1, dtype=torch.int32, device=label.device)
    _lib.check(_lib.load().dsee_labels_u8(_p(label), _p(out), label.numel(), num_classes, _p(bad), _stream()))
    return out, bad


def bicubic_clamp(x, size):
    """F.interpolate(x, size, mode='bicubic').clamp(-1, 1) (data/preprocessor.py:29-32), NCHW fp32."""
    _chk_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 4
    B, Cc, Hi, Wi = x.shape
    out = torch.empty((B, Cc, size[0], size[1]), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().dsee_bicubic_clamp(_p(x), _p(out), B, Cc, Hi, Wi, size[0], size[1], _stream()))
    return out


def labels_from_onehot(onehot):
    """fp32 one-hot [B,L,H,W] -> uint8 [B,H,W] plus a device flag that is 1 if not one-hot."""
    _chk_cuda(onehot)
    assert onehot.dtype == torch.float32 and onehot.dim() == 4
    B, L, H, W = onehot.shape
    labels = torch.empty((B, H, W), dtype=torch.uint8, device=onehot.device)
    bad = torch.zeros(1, dtype=torch.int32, device=onehot.device)
    _lib.check(_lib.load().dsee_labels_from_onehot(_p(onehot), _p(labels), B, L, H, W, _p(bad),
                                                   _stream()))
    return labels, bad


def resize_labels(labels, Hout, Wout):
    _chk_cuda(labels)
    assert labels.dtype == torch.uint8 and labels.dim() == 3
    B, Hin, Win = labels.shape
    if (Hin, Win) == (Hout, Wout):
        return labels
    out = torch.empty((B, Hout, Wout), dtype=torch.uint8, device=labels.device)
    _lib.check(_lib.load().dsee_resize_labels(_p(labels), _p(out), B, Hin, Win, Hout, Wout,
                                              _stream()))
    return out


# ---------------------------------------------------------------------------------------------
# operand builders
# ---------------------------------------------------------------------------------------------
def shared_mlp(labels, table, bias, ups=0, want_lo=True, uniform_rows=True):
    """relu(conv3x3(onehot(labels), W) + b) as a table gather. table fp32 [9,L,nh].
    uniform_rows: pixels whose 3x3 window carries a single label read one precomputed row (bit-identical)."""
    _chk_cuda(labels, table, bias)
    B, Hl, Wl = labels.shape
    _, L, nh = table.shape
    H, W = Hl << ups, Wl << ups
    hi = torch.empty((B, H, W, nh), dtype=torch.float16, device=labels.device)
    lo = torch.empty_like(hi) if want_lo else None
    rows = torch.empty((L, nh), dtype=torch.float32, device=labels.device) if uniform_rows else None
    _lib.check(_lib.load().dsee_shared_mlp_fwd(_p(labels), _p(table), _p(bias), _p(hi), _p(lo), B, Hl,
                                               Wl, ups, L, nh, _p(rows), _stream()))
    return SplitPlanes(hi, lo)


def style_gather(labels, style, want_lo=True):
    """style_map[b,y,x,:] = style[b, labels[b,y,x], :]. style fp32 [B,L,d]."""
    _chk_cuda(labels, style)
    B, H, W = labels.shape
    _, L, d = style.shape
    hi = torch.empty((B, H, W, d), dtype=torch.float16, device=labels.device)
    lo = torch.empty_like(hi) if want_lo else None
    _lib.check(_lib.load().dsee_style_gather_fwd(_p(labels), _p(style), _p(hi), _p(lo), B, H, W, L, d,
                                                 _stream()))
    return SplitPlanes(hi, lo)


def prep_conv_weight(w, want_lo=True, transpose=False, want_f8=False):
    """fp32 [N,C,3,3] -> scaled fp16 split planes [N, 9*C] in (tap, c) order.
    transpose=True: the backward-data operand [C, 9*N] (transposed, 180-degree rotated).
    want_f8: also the e4m3 companion for the fp8 correction GEMM (passes == 2)."""
    _chk_cuda(w)
    assert w.dtype == torch.float32 and w.dim() == 4 and w.shape[2:] == (3, 3)
    N, Cin = w.shape[:2]
    rows, cols = (Cin, N) if transpose else (N, Cin)
    hi = torch.empty((rows, 9 * cols), dtype=torch.float16, device=w.device)
    lo = torch.empty_like(hi) if want_lo else None
    inv = torch.empty(3, dtype=torch.float32, device=w.device)  # [2^-e, max|w|, row-L1 bound (transpose)]
    _lib.check(_lib.load().dsee_prep_conv_weight(_p(w), _p(hi), _p(lo), _p(inv), N, Cin,
                                                 int(transpose), _stream()))
    f8 = None
    if want_f8:
        assert not transpose
        f8 = torch.empty((N, 9, 2, Cin), dtype=torch.uint8, device=w.device)
        _lib.check(_lib.load().dsee_prep_conv_weight_f8(_p(w), _p(inv), _p(f8), N, Cin, _stream()))
    return PreparedWeight(hi, lo, inv, rows, cols, f8)


def prep_mod_weight_batched(wa, ws, Lp=64, want_lo=True):
    """SEAN modulation weight with the style branch folded into per-image weights over the one-hot
    label planes: wa fp32 [N,Ca,3,3] (shared), ws fp32 [B,N,Ls,3,3] -> planes [B*N, 9*(Ca+Lp)]."""
    _chk_cuda(wa, ws)
    N, Ca = wa.shape[:2]
    B, N2, Ls = ws.shape[:3]
    assert N2 == N and tuple(wa.shape[2:]) == (3, 3) and tuple(ws.shape[3:]) == (3, 3)
    hi = torch.empty((B * N, 9 * (Ca + Lp)), dtype=torch.float16, device=wa.device)
    lo = torch.empty_like(hi) if want_lo else None
    inv = torch.empty(2, dtype=torch.float32, device=wa.device)
    _lib.check(_lib.load().dsee_prep_mod_weight_batched(_p(wa), _p(ws), _p(hi), _p(lo), _p(inv), B, N, Ca,
                                                        Ls, Lp, _stream()))
    return PreparedWeight(hi, lo, inv, N, Ca + Lp, None, True)


def split_f16(x, want_lo=True):
    _chk_cuda(x)
    assert x.dtype == torch.float32
    hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi) if want_lo else None
    _lib.check(_lib.load().dsee_split_f16(_p(x), _p(hi), _p(lo), x.numel(), _stream()))
    return SplitPlanes(hi, lo)


# ---------------------------------------------------------------------------------------------
# tensor-core kernels
# ---------------------------------------------------------------------------------------------
_DTYPE_CODE = {torch.float16: 0, torch.bfloat16: 1}


def _operands(sources, pw, passes, sub=None):
    """sub: None, or (py, px) - the sources are at half resolution and this call computes one parity
    class of the sub-pixel form with the class's rows of the collapsed weight (collapse_subpixel)."""
    a0 = sources[0]
    B, H, W, C0 = a0.hi.shape
    if sub is not None:
        H, W = 2 * H, 2 * W
    ops = _lib.ConvOperands()
    ops.B, ops.H, ops.W = B, H, W
    ctot = 0
    keep = []
    for i in range(2):
        if i < len(sources):
            s = sources[i]
            _chk_cuda(s.hi, s.lo)
            assert s.hi.dtype == a0.hi.dtype and tuple(s.hi.shape[:3]) == tuple(a0.hi.shape[:3])
            ops.a_hi[i] = s.hi.data_ptr()
            ops.a_lo[i] = s.lo.data_ptr() if s.lo is not None else 0
            ops.a_channels[i] = s.hi.shape[3]
            ctot += s.hi.shape[3]
            keep.append(s)
        else:
            ops.a_hi[i] = 0
            ops.a_lo[i] = 0
            ops.a_channels[i] = 0
    assert ctot == pw.cin, "weight prepared for %d input channels, operands have %d" % (pw.cin, ctot)
    woff = 0
    ops.a_sub = ops.sub_py = ops.sub_px = 0
    if sub is not None:
        ops.a_sub, ops.sub_py, ops.sub_px = 1, sub[0], sub[1]
        assert pw.hi.shape[0] == 4 * pw.n_total, "sub-pixel form needs the 4-class collapsed weight"
        if sub[0] >= 0:   # one class per launch; (-1, -1) = all four classes in one launch
            woff = (sub[0] * 2 + sub[1]) * pw.n_total * pw.hi.shape[1] * 2   # bytes to the class's rows
    ops.w_hi = pw.hi.data_ptr() + woff
    ops.w_lo = pw.lo.data_ptr() + woff if pw.lo is not None else 0
    ops.w_inv_scale = pw.inv_scale.data_ptr()
    ops.n_total = pw.n_total
    ops.passes = passes
    ops.w_batch_rows = pw.n_total if pw.batched else 0
    if pw.batched:
        assert pw.hi.shape[0] == B * pw.n_total, "batched weight prepared for a different batch size"
    ops.a8_lo = ops.a8_hi = ops.w8 = 0
    if passes == 2:
        if a0.f8 is None or pw.f8 is None or len(sources) != 1:
            raise RuntimeError("passes == 2 (fp8 correction) needs one A source with fp8 planes "
                               "(spade_modulate(want_f8=True)) and a weight prepared with want_f8=True")
        ops.a8_lo, ops.a8_hi, ops.w8 = a0.f8[0].data_ptr(), a0.f8[1].data_ptr(), pw.f8.data_ptr()
    ops.a_dtype = _DTYPE_CODE[a0.hi.dtype]
    ops.w_dtype = _DTYPE_CODE[pw.hi.dtype]
    inv = getattr(a0, "inv_scale", None)
    ops.a_inv_scale = inv.data_ptr() if inv is not None else 0
    return ops, (B, H, W)


def conv3x3(sources, pw, bias, residual=None, res_ups=0, noises=(), passes=3, want_stats=False,
            act_mask=None, want_amax=False, tag="conv3x3", act16=False):
    """K2: 3x3 conv + bias (+ residual through an optional folded 2x upsample, + up to two
    NoiseInjection terms ``(noise NHWC, weight[C])``) -> fp32 NHWC.
    Backward-data use: ``sources`` = gradient planes (bf16), ``pw`` prepared with transpose=True,
    bias None, ``act_mask`` = fp16 hi plane of the forward activation (LeakyReLU' folded in).

    act16: also write leaky_relu(out, 0.2) as fp16 hi / lo planes (for the tensor-core image head);
    they come back as ``out._dsee_act16`` (SplitPlanes).

    Returns out, or (out, stats_partial) when want_stats (partials for bn_finalize), or
    (out, amax) when want_amax (device scalar max|out|)."""
    ops, (B, H, W) = _operands(sources, pw, passes)
    _chk_cuda(bias, residual)
    dev = sources[0].hi.device
    out = torch.empty((B, H, W, pw.n_total), dtype=torch.float32, device=dev)
    stats = None
    if want_stats:
        nt = _lib.load().dsee_conv3x3_stats_tiles(B, H, W)
        stats = torch.empty((nt, pw.n_total, 2), dtype=torch.float32, device=dev)
    epi = _lib.ConvEpilogue()
    epi.bias = bias.data_ptr() if bias is not None else 0
    epi.act_mask = act_mask.data_ptr() if act_mask is not None else 0
    epi.residual = residual.data_ptr() if residual is not None else 0
    epi.res_ups = res_ups
    noises = [n for n in noises if n is not None and n[0] is not None]
    assert len(noises) <= 2
    for i in range(2):
        if i < len(noises):
            _chk_cuda(noises[i][0], noises[i][1])
            ptr, seed = _noise(noises[i][0])
            epi.noise[i] = ptr.value or 0
            epi.noise_seed[i] = seed
            epi.noise_w[i] = noises[i][1].data_ptr()
        else:
            epi.noise[i] = 0
            epi.noise_seed[i] = 0
            epi.noise_w[i] = 0
    epi.out = out.data_ptr()
    epi.stats_partial = stats.data_ptr() if stats is not None else 0
    amax = torch.empty(1, dtype=torch.float32, device=dev) if want_amax else None
    epi.amax_out = amax.data_ptr() if amax is not None else 0
    if act16:
        planes = SplitPlanes(torch.empty(out.shape, dtype=torch.float16, device=dev),
                             torch.empty(out.shape, dtype=torch.float16, device=dev))
        epi.act16_hi, epi.act16_lo = planes.hi.data_ptr(), planes.lo.data_ptr()
        out._dsee_act16 = planes
    flops = 2.0 * 9 * pw.cin * pw.n_total * B * H * W  # reference-equivalent dense conv FLOPs
    _timed("%s_%dx%d" % (tag, H, W), flops,
           lambda: _lib.check(_lib.load().dsee_conv3x3_fwd(C.byref(ops), C.byref(epi), _stream())))
    if want_amax:
        return out, amax
    return (out, stats) if want_stats else out


_SUBPIX_M = {}


def subpixel_matrix(device):
    """M[class, tap2x2, tap3x3] in {0, 1}: which taps of a 3x3 filter over a nearest-2x-upsampled
    tensor land on the same half-resolution pixel, per output parity class (py, px) = divmod(class, 2)
    and 2x2 tap (ty, tx) = divmod(tap2x2, 2): 3x3 tap (ky, kx) reads half-resolution row
    (py + ky - 1) // 2 = ty + py - 1."""
    m = _SUBPIX_M.get(device)
    if m is None:
        m = torch.zeros(4, 4, 9)
        for py in range(2):
            for px in range(2):
                for ky in range(3):
                    for kx in range(3):
                        ty, tx = (py + ky - 1) // 2 - (py - 1), (px + kx - 1) // 2 - (px - 1)
                        m[py * 2 + px, ty * 2 + tx, ky * 3 + kx] = 1.0
        m = _SUBPIX_M[device] = m.to(device)
    return m


def collapse_subpixel(w):
    """[N, C, 3, 3] -> [4, N, C, 2, 2]: the four 2x2 filters that, applied to a half-resolution tensor,
    equal the 3x3 filter applied to its nearest 2x upsampling (one per output parity class).  A torch
    einsum, so autograd turns the gradient of the collapsed filters back into the 3x3 gradient."""
    N, Cc = w.shape[:2]
    wc = torch.einsum('ncp,ktp->knct', w.reshape(N, Cc, 9), subpixel_matrix(w.device))
    return wc.contiguous().reshape(4, N, Cc, 2, 2)


def prep_subpixel_weight(wc, want_lo=True):
    """collapse_subpixel output [4, N, C, 2, 2] -> planes [4N, 4C] (one scale for the four classes)."""
    _, N, Cc = wc.shape[:3]
    pw = prep_conv_weight_ex(wc.reshape(4 * N, Cc, 2, 2).contiguous(), want_lo)
    return PreparedWeight(pw.hi, pw.lo, pw.inv_scale, N, Cc)


def spade_modulate(sources, pw, x, x_ups, bn_scale, bn_shift, gamma_bias, beta_bias, noise=None,
                   noise_w=None, passes=3, want_lo=True, save_g=False, want_f8=False, subpixel=False):
    """K1: gamma/beta conv + batch-norm apply + modulation + LeakyReLU -> fp16 split planes.
    save_g: also return G = gamma + gamma_bias as split planes (what K1's backward multiplies by).
    want_f8: also the e5m2 planes a passes == 2 main conv reads (fp8 correction GEMM).
    subpixel: the sources are at HALF the output resolution (the reference convolves their nearest 2x
    upsampling, normalization.py:188-190,275-277) and pw is prep_subpixel_weight's: the four output
    parity classes are tiles of one launch (the classes of a pixel tile run back to back, so x and the
    sources are read from HBM once), 4/9 of the FLOPs."""
    ops, (B, H, W) = _operands(sources, pw, passes, (-1, -1) if subpixel else None)
    _chk_cuda(x, bn_scale, bn_shift, gamma_bias, beta_bias, noise, noise_w)
    Cc = x.shape[3]
    assert tuple(x.shape[:3]) == (B, H >> x_ups, W >> x_ups)
    dev = x.device
    hi = torch.empty((B, H, W, Cc), dtype=torch.float16, device=dev)
    lo = torch.empty_like(hi) if want_lo else None
    m = _lib.ModulateArgs()
    m.x, m.x_ups = x.data_ptr(), x_ups
    nptr, nseed = _noise(noise)
    m.noise, m.noise_seed = nptr.value or 0, nseed
    m.noise_w = noise_w.data_ptr() if noise_w is not None else 0
    m.bn_scale, m.bn_shift = bn_scale.data_ptr(), bn_shift.data_ptr()
    m.gamma_bias, m.beta_bias = gamma_bias.data_ptr(), beta_bias.data_ptr()
    m.out_hi = hi.data_ptr()
    m.out_lo = lo.data_ptr() if lo is not None else 0
    m.C = Cc
    ghi = glo = None
    if save_g:
        ghi = torch.empty_like(hi)
        glo = torch.empty_like(hi) if want_lo else None
    m.g_hi = ghi.data_ptr() if ghi is not None else 0
    m.g_lo = glo.data_ptr() if glo is not None else 0
    f8 = None
    if want_f8:
        f8 = (torch.empty((B, H, W, Cc), dtype=torch.uint8, device=dev),
              torch.empty((B, H, W, Cc), dtype=torch.uint8, device=dev))
    m.out8_lo = f8[0].data_ptr() if f8 is not None else 0
    m.out8_hi = f8[1].data_ptr() if f8 is not None else 0
    flops = 2.0 * 9 * pw.cin * pw.n_total * B * H * W   # reference-equivalent (sub-pixel executes 4/9)

    _timed("modulate_%dx%d" % (H, W), flops,
           lambda: _lib.check(_lib.load().dsee_spade_modulate_fwd(C.byref(ops), C.byref(m), _stream())))
    if save_g:
        return SplitPlanes(hi, lo, f8), SplitPlanes(ghi, glo)
    return SplitPlanes(hi, lo, f8)


def spade_modulate_bwd_saved(g_planes, x, x_ups, bn_scale, bn_shift, dt, dt_amax, noise=None,
                             noise_w=None, want_lo=True):
    """K1 backward from the saved G planes (one streaming pass) -> (dxhat, dgb GradPlanes, sums[4,C])."""
    _chk_cuda(x, bn_scale, bn_shift, dt, dt_amax, noise, noise_w, g_planes.hi, g_planes.lo)
    B, H, W, Cc = dt.shape
    dev = x.device
    lib = _lib.load()
    dxhat = torch.empty((B, H, W, Cc), dtype=torch.float32, device=dev)
    ghi = torch.empty((B, H, W, 2 * Cc), dtype=torch.float16, device=dev)
    glo = torch.empty_like(ghi) if want_lo else None
    ginv = torch.empty(1, dtype=torch.float32, device=dev)
    part = torch.empty((lib.dsee_grad_prep_blocks(B * H * W), Cc, 4), dtype=torch.float32, device=dev)
    nptr, nseed = _noise(noise)
    _lib.check(lib.dsee_spade_modulate_bwd_saved(_p(x), x_ups, nptr, nseed, _p(noise_w), _p(bn_scale),
                                                 _p(bn_shift), _p(g_planes.hi), _p(g_planes.lo), _p(dt),
                                                 _p(dt_amax), B, H, W, Cc, _p(dxhat), _p(ghi), _p(glo),
                                                 _p(ginv), _p(part), _stream()))
    return dxhat, GradPlanes(ghi, glo, ginv), reduce_partials(part)


def dgrad_modulate_bwd(dy, pwT, act_mask, g_planes, x, x_ups, bn_scale, bn_shift, noise=None, noise_w=None,
                       passes=3, want_lo=True, tag="dgrad_modbwd"):
    """Backward-data of a main conv fused with K1's backward (dt never reaches HBM):
    dy GradPlanes of the conv output, pwT = prep_conv_weight(W, transpose=True), act_mask = fp16 hi
    plane of the conv's input activation, g_planes = G saved by K1's forward.
    -> (dxhat fp32 NHWC, dgb GradPlanes [B,H,W,2C] interleaved, sums fp32 [4,C])."""
    ops, (B, H, W) = _operands([dy], pwT, passes)
    _chk_cuda(act_mask, g_planes.hi, g_planes.lo, x, bn_scale, bn_shift, noise, noise_w)
    Cc = x.shape[3]
    assert pwT.n_total == Cc and tuple(act_mask.shape) == (B, H, W, Cc)
    assert tuple(x.shape[:3]) == (B, H >> x_ups, W >> x_ups)
    dev = x.device
    lib = _lib.load()
    dxhat = torch.empty((B, H, W, Cc), dtype=torch.float32, device=dev)
    ghi = torch.empty((B, H, W, 2 * Cc), dtype=torch.float16, device=dev)
    glo = torch.empty_like(ghi) if want_lo else None
    ginv = torch.empty(1, dtype=torch.float32, device=dev)
    part = torch.empty((lib.dsee_conv3x3_stats_tiles(B, H, W), Cc, 4), dtype=torch.float32, device=dev)
    a = _lib.DgradModBwdArgs()
    a.act_mask = act_mask.data_ptr()
    a.g_hi = g_planes.hi.data_ptr()
    a.g_lo = g_planes.lo.data_ptr() if g_planes.lo is not None else 0
    a.x, a.x_ups = x.data_ptr(), x_ups
    nptr, nseed = _noise(noise)
    a.noise, a.noise_seed = nptr.value or 0, nseed
    a.noise_w = noise_w.data_ptr() if noise_w is not None else 0
    a.bn_scale, a.bn_shift = bn_scale.data_ptr(), bn_shift.data_ptr()
    a.dy_amax = dy.inv_scale.data_ptr() + 4        # grad_prep: inv_scale[1] = max|dY|
    a.w_l1 = pwT.inv_scale.data_ptr() + 8          # prep_conv_weight(transpose): inv_scale[2]
    a.dxhat, a.dgb_hi = dxhat.data_ptr(), ghi.data_ptr()
    a.dgb_lo = glo.data_ptr() if glo is not None else 0
    a.dgb_inv_scale, a.partial, a.C = ginv.data_ptr(), part.data_ptr(), Cc
    flops = 2.0 * 9 * pwT.cin * pwT.n_total * B * H * W
    _timed("%s_%dx%d" % (tag, H, W), flops,
           lambda: _lib.check(lib.dsee_dgrad_modulate_bwd(C.byref(ops), C.byref(a), _stream())))
    return dxhat, GradPlanes(ghi, glo, ginv), reduce_partials(part)


# ---------------------------------------------------------------------------------------------
# generator backward
# ---------------------------------------------------------------------------------------------
def grad_prep(dy, noise0=None, noise1=None, want_lo=True, amax=None):
    """dY fp32 NHWC -> (scaled fp16 GradPlanes, sums fp32 [nq,C]): sums[0] = sum dY (bias
    gradient), sums[1+i] = sum dY*noise_i (NoiseInjection.weight gradients).
    amax: device scalar max|dY| when the producer already knows it (bn_bwd(want_amax=True));
    otherwise it is computed here with one extra pass over dY."""
    _chk_cuda(dy, noise0, noise1)
    Cc = dy.shape[-1]
    npix = dy.numel() // Cc
    lib = _lib.load()
    hi = torch.empty(dy.shape, dtype=torch.float16, device=dy.device)
    lo = torch.empty_like(hi) if want_lo else None
    inv = torch.empty(2, dtype=torch.float32, device=dy.device)
    nq = 1 + (noise0 is not None) + (noise1 is not None)
    nb = lib.dsee_grad_prep_blocks(npix)
    part = torch.empty((nb, Cc, nq), dtype=torch.float32, device=dy.device)
    (p0, s0), (p1, s1) = _noise(noise0), _noise(noise1)
    _lib.check(lib.dsee_grad_prep(_p(dy), _p(hi), _p(lo), _p(inv), p0, p1, s0, s1, npix, Cc, _p(part),
                                  _p(amax), _stream()))
    return GradPlanes(hi, lo, inv), reduce_partials(part)


def reduce_partials(part, scale=1.0):
    """[n,C,nq] block partials -> [nq,C] (fixed order, double accumulation)."""
    _chk_cuda(part)
    n, Cc, nq = part.shape
    out = torch.empty((nq, Cc), dtype=torch.float32, device=part.device)
    _lib.check(_lib.load().dsee_reduce_partials(_p(part), n, Cc, nq, float(scale), _p(out), _stream()))
    return out


def conv3x3_wgrad(dy, a, passes=3):
    """dW[n][c][3][3] = sum_pixels dY[.,n] * A[.+tap,c]; dy = GradPlanes, a = SplitPlanes NHWC."""
    _chk_cuda(dy.hi, dy.lo, a.hi, a.lo)
    B, H, W, N = dy.hi.shape
    Cc = a.hi.shape[3]
    assert tuple(a.hi.shape[:3]) == (B, H, W)
    lib = _lib.load()
    ws = torch.empty(lib.dsee_conv3x3_wgrad_workspace_floats(B, H, W, N, Cc), dtype=torch.float32,
                     device=dy.hi.device)
    dw = torch.empty((N, Cc, 3, 3), dtype=torch.float32, device=dy.hi.device)
    flops = 2.0 * 9 * Cc * N * B * H * W
    assert dy.hi.dtype == a.hi.dtype
    _timed("wgrad_%dx%d" % (H, W), flops, lambda: _lib.check(lib.dsee_conv3x3_wgrad(
        _p(dy.hi), _p(dy.lo), _p(getattr(dy, "inv_scale", None)), _p(a.hi), _p(a.lo),
        _p(getattr(a, "inv_scale", None)), _DTYPE_CODE[dy.hi.dtype], B, H, W, N, Cc, passes, _p(ws),
        _p(dw), 1, _stream())))
    return dw


def conv3x3_wgrad_multi(dy, sources, passes=3):
    """conv3x3_wgrad against up to two channel-concatenated activation sources in one launch:
    -> dW [N, C0 + C1, 3, 3]."""
    if len(sources) == 1:
        return conv3x3_wgrad(dy, sources[0], passes=passes)
    assert len(sources) == 2
    _chk_cuda(dy.hi, dy.lo, *[t for s_ in sources for t in (s_.hi, s_.lo)])
    B, H, W, N = dy.hi.shape
    chans = [s_.hi.shape[3] for s_ in sources]
    ctot = sum(chans)
    lib = _lib.load()
    ws = torch.empty(lib.dsee_conv3x3_wgrad_workspace_floats(B, H, W, N, ctot), dtype=torch.float32,
                     device=dy.hi.device)
    dw = torch.empty((N, ctot, 3, 3), dtype=torch.float32, device=dy.hi.device)
    a_hi = (C.c_void_p * 2)(*[s_.hi.data_ptr() for s_ in sources])
    a_lo = (C.c_void_p * 2)(*[(s_.lo.data_ptr() if s_.lo is not None else 0) for s_ in sources])
    ach = (C.c_int * 2)(*chans)
    flops = 2.0 * 9 * ctot * N * B * H * W
    _timed("wgrad_%dx%d" % (H, W), flops, lambda: _lib.check(lib.dsee_conv3x3_wgrad2(
        _p(dy.hi), _p(dy.lo), _p(getattr(dy, "inv_scale", None)), a_hi, a_lo, ach,
        _DTYPE_CODE[dy.hi.dtype], B, H, W, N, passes, _p(ws), _p(dw), 1, _stream())))
    return dw


def conv3x3_wgrad_per_image(dy, sources, passes=3):
    """conv3x3_wgrad_multi with one result per image -> dW [B, N, C0 + C1, 3, 3]."""
    srcs = list(sources) + [None] * (2 - len(sources))
    _chk_cuda(dy.hi, dy.lo, *[t for s_ in sources for t in (s_.hi, s_.lo)])
    B, H, W, N = dy.hi.shape
    chans = [s_.hi.shape[3] if s_ is not None else 0 for s_ in srcs]
    ctot = sum(chans)
    lib = _lib.load()
    ws = torch.empty(lib.dsee_conv3x3_wgrad_per_image_workspace_floats(B, H, W, N, ctot),
                     dtype=torch.float32, device=dy.hi.device)
    dw = torch.empty((B, N, ctot, 3, 3), dtype=torch.float32, device=dy.hi.device)
    a_hi = (C.c_void_p * 2)(*[(s_.hi.data_ptr() if s_ is not None else 0) for s_ in srcs])
    a_lo = (C.c_void_p * 2)(*[(s_.lo.data_ptr() if s_ is not None and s_.lo is not None else 0) for s_ in srcs])
    ach = (C.c_int * 2)(*chans)
    flops = 2.0 * 9 * ctot * N * B * H * W
    _timed("wgrad_%dx%d" % (H, W), flops, lambda: _lib.check(lib.dsee_conv3x3_wgrad2_per_image(
        _p(dy.hi), _p(dy.lo), _p(getattr(dy, "inv_scale", None)), a_hi, a_lo, ach,
        _DTYPE_CODE[dy.hi.dtype], B, H, W, N, passes, _p(ws), _p(dw), _stream())))
    return dw


def subpixel_wgrad(dy, sources, passes=3):
    """Weight gradient of the sub-pixel form: dy GradPlanes [B,H,W,N] (full resolution), sources at
    half resolution -> d(collapsed filters) [4, N, C0 + C1, 2, 2]."""
    srcs = list(sources) + [None] * (2 - len(sources))
    _chk_cuda(dy.hi, dy.lo, *[t for s_ in sources for t in (s_.hi, s_.lo)])
    B, H, W, N = dy.hi.shape
    chans = [s_.hi.shape[3] if s_ is not None else 0 for s_ in srcs]
    ctot = sum(chans)
    lib = _lib.load()
    ws = torch.empty(lib.dsee_subpixel_wgrad_workspace_floats(B, H, W, N, ctot), dtype=torch.float32,
                     device=dy.hi.device)
    dwc = torch.empty((4, N, ctot, 2, 2), dtype=torch.float32, device=dy.hi.device)
    a_hi = (C.c_void_p * 2)(*[(s_.hi.data_ptr() if s_ is not None else 0) for s_ in srcs])
    a_lo = (C.c_void_p * 2)(*[(s_.lo.data_ptr() if s_ is not None and s_.lo is not None else 0) for s_ in srcs])
    ach = (C.c_int * 2)(*chans)
    flops = 2.0 * 9 * ctot * N * B * H * W

    def launch():
        for cls in range(4):
            py, px = divmod(cls, 2)
            _lib.check(lib.dsee_subpixel_wgrad(_p(dy.hi), _p(dy.lo), _p(getattr(dy, "inv_scale", None)), a_hi,
                                               a_lo, ach, B, H, W, N, py, px, passes, _p(ws), _p(dwc[cls]),
                                               _stream()))
    _timed("wgrad_%dx%d" % (H, W), flops, launch)
    return dwc


def subpixel_dgrad(dy, wc, passes=3, want_lo=True):
    """Gradient wrt the half-resolution sources of the sub-pixel form: dy GradPlanes [B,H,W,N], wc the
    collapsed filters [4, N, C, 2, 2] -> (dsrc fp32 [B,H/2,W/2,C], device max|dsrc|)."""
    _chk_cuda(dy.hi, dy.lo)
    B, H, W, N = dy.hi.shape
    Cc = wc.shape[2]
    # [C, N, class, tap]: one "16-tap" filter bank per source channel
    pwt = prep_conv_weight_ex(wc.reshape(4, N, Cc, 4).permute(2, 1, 0, 3).contiguous(), want_lo)
    out = torch.empty((B, H // 2, W // 2, Cc), dtype=torch.float32, device=dy.hi.device)
    amax = torch.empty(1, dtype=torch.float32, device=dy.hi.device)
    flops = 2.0 * 9 * Cc * N * B * H * W
    _timed("dgrad_mod_%dx%d" % (H, W), flops, lambda: _lib.check(_lib.load().dsee_subpixel_dgrad(
        _p(dy.hi), _p(dy.lo), _p(getattr(dy, "inv_scale", None)), _p(pwt.hi), _p(pwt.lo), _p(pwt.inv_scale),
        B, H, W, N, Cc, passes, _p(out), _p(amax), _stream())))
    return out, amax


def spade_modulate_bwd(sources, pw_gamma, x, x_ups, bn_scale, bn_shift, gamma_bias, dt, dt_amax,
                       noise=None, noise_w=None, passes=3, want_lo=True):
    """K1 backward -> (dxhat fp32 NHWC, dgb GradPlanes [B,H,W,2C] interleaved,
    sums fp32 [4,C] = sum dxhat, sum dxhat*xhat, sum dG, sum dB). dt_amax: device max|dt|."""
    ops, (B, H, W) = _operands(sources, pw_gamma, passes)
    _chk_cuda(x, bn_scale, bn_shift, gamma_bias, dt, noise, noise_w)
    Cc = x.shape[3]
    assert pw_gamma.n_total == Cc and tuple(dt.shape) == (B, H, W, Cc)
    dev = x.device
    lib = _lib.load()
    dxhat = torch.empty((B, H, W, Cc), dtype=torch.float32, device=dev)
    ghi = torch.empty((B, H, W, 2 * Cc), dtype=torch.float16, device=dev)
    glo = torch.empty_like(ghi) if want_lo else None
    ginv = torch.empty(1, dtype=torch.float32, device=dev)
    part = torch.empty((lib.dsee_conv3x3_stats_tiles(B, H, W), Cc, 4), dtype=torch.float32, device=dev)
    m = _lib.ModulateBwdArgs()
    m.x, m.x_ups = x.data_ptr(), x_ups
    if isinstance(noise, NoiseSeed):
        raise RuntimeError("the GEMM-recompute K1 backward takes explicit noise tensors only; "
                           "keep DSEE_SAVE_GAMMA=1 (default) with in-kernel noise")
    m.noise = noise.data_ptr() if noise is not None else 0
    m.noise_w = noise_w.data_ptr() if noise_w is not None else 0
    m.bn_scale, m.bn_shift = bn_scale.data_ptr(), bn_shift.data_ptr()
    m.gamma_bias, m.dt = gamma_bias.data_ptr(), dt.data_ptr()
    m.dxhat, m.dgb_hi = dxhat.data_ptr(), ghi.data_ptr()
    m.dgb_lo = glo.data_ptr() if glo is not None else 0
    m.partial = part.data_ptr()
    m.C = Cc
    m.dt_amax, m.dgb_inv_scale = dt_amax.data_ptr(), ginv.data_ptr()
    flops = 2.0 * 9 * pw_gamma.cin * pw_gamma.n_total * B * H * W
    _timed("modulate_bwd_%dx%d" % (H, W), flops,
           lambda: _lib.check(lib.dsee_spade_modulate_bwd(C.byref(ops), C.byref(m), _stream())))
    return dxhat, GradPlanes(ghi, glo, ginv), reduce_partials(part)


def bn_bwd(dxhat, x, x_ups, bn_scale, bn_shift, sums, inv_count, noise=None, noise_w=None, dskip=None,
           noise_grad_with_skip=False, want_amax=False):
    """-> (dx fp32 NHWC at x's resolution, d noise_w [C] or None[, device scalar max|dx|]).
    noise_grad_with_skip: d noise_w also collects sum(dskip * noise) (the shortcut's noise term)."""
    _chk_cuda(dxhat, x, bn_scale, bn_shift, sums, noise, noise_w, dskip)
    B, Hx, Wx, Cc = x.shape
    lib = _lib.load()
    dx = torch.empty_like(x)
    nwp = None
    if noise is not None:
        nwp = torch.empty((lib.dsee_bn_bwd_blocks(B, Hx, Wx), Cc, 1), dtype=torch.float32,
                          device=x.device)
    nptr, nseed = _noise(noise)
    amax = torch.empty(1, dtype=torch.float32, device=x.device) if want_amax else None
    _lib.check(lib.dsee_bn_bwd(_p(dxhat), _p(x), x_ups, nptr, nseed, _p(noise_w), _p(bn_scale),
                               _p(bn_shift), _p(sums), float(inv_count), _p(dskip), B, Hx, Wx, Cc,
                               _p(dx), _p(nwp), int(bool(noise_grad_with_skip)), _p(amax), _stream()))
    dnw = reduce_partials(nwp)[0] if nwp is not None else None
    return (dx, dnw, amax) if want_amax else (dx, dnw)


def shared_mlp_bwd(dsrc, coff, actv_hi, labels, ups, L):
    """Gradient of the 9-tap table and the bias: -> (dtable [9,L,nh], dbias [nh])."""
    _chk_cuda(dsrc, actv_hi, labels)
    B, Hl, Wl = labels.shape
    nh = actv_hi.shape[3]
    ld = dsrc.shape[3]
    lib = _lib.load()
    rows = 9 * L + 1
    part = torch.empty((lib.dsee_shared_mlp_bwd_blocks(B, Hl, Wl), rows, nh), dtype=torch.float32,
                       device=dsrc.device)
    out = torch.empty((rows, nh), dtype=torch.float32, device=dsrc.device)
    _lib.check(lib.dsee_shared_mlp_bwd(_p(dsrc), ld, coff, _p(actv_hi), _p(labels), B, Hl, Wl, ups, L,
                                       nh, _p(part), _p(out), _stream()))
    return out[:9 * L].view(9, L, nh), out[9 * L]


def onehot_planes(labels, Lp=64):
    """uint8 [B,H,W] -> fp16 one-hot plane [B,H,W,Lp] (exact; the A operand of the table wgrad)."""
    _chk_cuda(labels)
    B, H, W = labels.shape
    out = torch.empty((B, H, W, Lp), dtype=torch.float16, device=labels.device)
    _lib.check(_lib.load().dsee_onehot_planes(_p(labels), _p(out), B * H * W, Lp, _stream()))
    return SplitPlanes(out, None)


def shared_mlp_bwd_tc(dsrc, dsrc_amax, coff, actv_hi, labels, onehot, ups, L, passes=3, side=None):
    """mlp_shared backward as a tensor-core wgrad against the one-hot plane:
    -> (dtable [9,L,nh], dbias [nh])."""
    _chk_cuda(dsrc, dsrc_amax, actv_hi, labels, onehot.hi)
    B, Hl, Wl = labels.shape
    nh = actv_hi.shape[3]
    lib = _lib.load()
    want_lo = passes == 3
    hi = torch.empty((B, Hl, Wl, nh), dtype=torch.float16, device=dsrc.device)
    lo = torch.empty_like(hi) if want_lo else None
    inv = torch.empty(1, dtype=torch.float32, device=dsrc.device)
    part = torch.empty((lib.dsee_grad_prep_blocks(B * Hl * Wl), nh, 1), dtype=torch.float32,
                       device=dsrc.device)
    _lib.check(lib.dsee_actv_grad_prep(_p(dsrc), dsrc.shape[3], coff, _p(actv_hi), _p(dsrc_amax), B, Hl,
                                       Wl, ups, nh, _p(hi), _p(lo), _p(inv), _p(part), _stream()))
    g = GradPlanes(hi, lo, inv)
    oh = onehot if not want_lo else SplitPlanes(onehot.hi, _zeros_like_cached(onehot.hi))
    def table_grad():
        dw = conv3x3_wgrad(g, oh, passes=passes)  # [nh, Lp, 3, 3]
        return dw[:, :L].permute(2, 3, 1, 0).reshape(9, L, nh)
    dtable = side.run(table_grad, g, oh) if side is not None else table_grad()
    return dtable, reduce_partials(part)[0]


_ZERO_CACHE = {}


def _zeros_like_cached(t):
    """An all-zero lo plane for exact operands (the one-hot map), shared across calls."""
    key = (tuple(t.shape), t.dtype, t.device)
    z = _ZERO_CACHE.get(key)
    if z is None:
        z = torch.zeros_like(t)
        _ZERO_CACHE[key] = z
    return z


def style_gather_bwd(dsrc, coff, labels, L, d):
    """dstyle[b,l,:] = sum over pixels with label l of dsrc[b,y,x,coff:coff+d]."""
    _chk_cuda(dsrc, labels)
    B, H, W = labels.shape
    lib = _lib.load()
    ws = torch.empty((B, lib.dsee_region_pool_chunks(H * W), L, d), dtype=torch.float32,
                     device=dsrc.device)
    out = torch.empty((B, L, d), dtype=torch.float32, device=dsrc.device)
    _lib.check(lib.dsee_style_gather_bwd(_p(dsrc), dsrc.shape[3], coff, _p(labels), _p(out), _p(ws), B,
                                         H * W, L, d, _stream()))
    return out


def stem_bwd(x_nchw, dy):
    """-> (dW [C,3,3,3], dbias [C])."""
    _chk_cuda(x_nchw, dy)
    B, _, H, W = x_nchw.shape
    Cc = dy.shape[3]
    lib = _lib.load()
    part = torch.empty((lib.dsee_stem_bwd_blocks(B, H, W), Cc, 28), dtype=torch.float32, device=dy.device)
    out = torch.empty((Cc, 28), dtype=torch.float32, device=dy.device)
    _lib.check(lib.dsee_stem_bwd(_p(x_nchw), _p(dy), B, H, W, Cc, _p(part), _p(out), _stream()))
    return out[:, :27].reshape(Cc, 3, 3, 3), out[:, 27].contiguous()


def head_bwd(x_nhwc, w, out, dout):
    """-> (dx fp32 NHWC, dW [3,C,3,3], dbias [3])."""
    _chk_cuda(x_nhwc, w, out, dout)
    B, H, W, Cc = x_nhwc.shape
    lib = _lib.load()
    dx = torch.empty_like(x_nhwc)
    part = torch.empty((lib.dsee_head_bwd_blocks(B, H, W), Cc, 28), dtype=torch.float32,
                       device=dout.device)
    res = torch.empty((Cc, 28), dtype=torch.float32, device=dout.device)
    _lib.check(lib.dsee_head_bwd(_p(x_nhwc), _p(w), _p(out), _p(dout), B, H, W, Cc, _p(dx), _p(part),
                                 _p(res), _stream()))
    dw = res[:, :27].reshape(Cc, 3, 9).permute(1, 0, 2).reshape(3, Cc, 3, 3).contiguous()
    return dx, dw, res[:3, 27].contiguous()


# ---------------------------------------------------------------------------------------------
# batch norm
# ---------------------------------------------------------------------------------------------
def bn_stats(x, x_ups=0, noise=None, noise_w=None):
    """Tile partials [n,C,2] of per-channel sum / sum of squares of x (NHWC fp32)."""
    _chk_cuda(x, noise, noise_w)
    B, Hx, Wx, Cc = x.shape
    H, W = Hx << x_ups, Wx << x_ups
    n = C.c_int(0)
    lib = _lib.load()
    _lib.check(lib.dsee_bn_stats(C.c_void_p(0), x_ups, C.c_void_p(0), 0, C.c_void_p(0), B, H, W, Cc,
                                 C.c_void_p(0), C.byref(n), _stream()))
    part = torch.empty((n.value, Cc, 2), dtype=torch.float32, device=x.device)
    nptr, nseed = _noise(noise)
    _lib.check(lib.dsee_bn_stats(_p(x), x_ups, nptr, nseed, _p(noise_w), B, H, W, Cc, _p(part),
                                 C.byref(n), _stream()))
    return part


def bn_finalize(partials, count, eps, momentum=0.1, running_mean=None, running_var=None,
                unbias_count=0):
    """-> (bn_scale, bn_shift, mean, var); updates running stats in place when given."""
    _chk_cuda(partials, running_mean, running_var)
    n, Cc, _ = partials.shape
    dev = partials.device
    sc = torch.empty(Cc, dtype=torch.float32, device=dev)
    sh = torch.empty_like(sc)
    mean = torch.empty_like(sc)
    var = torch.empty_like(sc)
    _lib.check(_lib.load().dsee_bn_finalize(_p(partials), n, Cc, float(count), float(unbias_count),
                                            float(eps),
                                            float(momentum), _p(running_mean), _p(running_var),
                                            _p(sc), _p(sh), _p(mean), _p(var), _stream()))
    return sc, sh, mean, var


def bn_eval_affine(running_mean, running_var, eps):
    _chk_cuda(running_mean, running_var)
    Cc = running_mean.numel()
    sc = torch.empty(Cc, dtype=torch.float32, device=running_mean.device)
    sh = torch.empty_like(sc)
    _lib.check(_lib.load().dsee_bn_eval_affine(_p(running_mean), _p(running_var), float(eps), Cc,
                                               _p(sc), _p(sh), _stream()))
    return sc, sh


# ---------------------------------------------------------------------------------------------
# generator ends
# ---------------------------------------------------------------------------------------------
def stem(x_nchw, w, bias):
    """DeepSEESR.initial: fp32 NCHW [B,3,H,W] -> fp32 NHWC [B,H,W,C]."""
    _chk_cuda(x_nchw, w, bias)
    B, _, H, W = x_nchw.shape
    Cc = w.shape[0]
    out = torch.empty((B, H, W, Cc), dtype=torch.float32, device=x_nchw.device)
    _lib.check(_lib.load().dsee_stem_fwd(_p(x_nchw), _p(w), _p(bias), _p(out), B, H, W, Cc,
                                         _stream()))
    return out


def head(x_nhwc, w, bias):
    """tanh(conv_img(leaky_relu(x, 0.2))): fp32 NHWC [B,H,W,C] -> fp32 NCHW [B,3,H,W]."""
    _chk_cuda(x_nhwc, w, bias)
    B, H, W, Cc = x_nhwc.shape
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=x_nhwc.device)
    _lib.check(_lib.load().dsee_head_fwd(_p(x_nhwc), _p(w), _p(bias), _p(out), B, H, W, Cc,
                                         _stream()))
    return out


def _head_w27(w):
    """conv_img weight [3,C,3,3] -> the 1x1 form [32,C,1,1]: row tap*3 + o = w[o,:,tap], rows 27..31 zero."""
    Cc = w.shape[1]
    w27 = w.permute(2, 3, 0, 1).reshape(27, Cc)
    return torch.nn.functional.pad(w27, (0, 0, 0, 5)).reshape(32, Cc, 1, 1).contiguous()


def head_tc(a, w, bias, passes=3):
    """The image head on the tensor cores: a = SplitPlanes of leaky_relu(x, 0.2) (NHWC [B,H,W,C], written by
    the last conv3x3 with act16=True), w [3,C,3,3] -> tanh(conv_img(.)) fp32 NCHW [B,3,H,W].
    One 1x1 GEMM into the 27 (tap, channel) partial products per pixel, then a 9-tap shift-add."""
    _chk_cuda(a.hi, a.lo, w, bias)
    B, H, W, Cc = a.hi.shape
    if passes == 3 and a.lo is None:
        passes = 1
    pw = prep_conv_weight_ex(_head_w27(w), want_lo=passes == 3)
    P = conv2d_tc(a if passes == 3 else SplitPlanes(a.hi, None), pw, None, 1, 1, 1, 0, (H, W), passes=passes,
                  tag="head_gemm")
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=w.device)
    _lib.check(_lib.load().dsee_head_gather_fwd(_p(P), _p(bias), _p(out), B, H, W, _stream()))
    return out


def head_tc_bwd(a, w, out, dout, passes=1):
    """Backward of head_tc -> (dx fp32 NHWC [B,H,W,C] with max|dx| as a device scalar, dW [3,C,3,3], db [3])."""
    _chk_cuda(a.hi, a.lo, w, out, dout)
    B, H, W, Cc = a.hi.shape
    if passes == 3 and a.lo is None:
        passes = 1
    want_lo = passes == 3
    dP = torch.empty((B, H, W, 32), dtype=torch.float32, device=w.device)
    _lib.check(_lib.load().dsee_head_scatter_bwd(_p(dout), _p(out), _p(dP), B, H, W, _stream()))
    g, sums = grad_prep(dP, want_lo=want_lo)
    db = sums[0][12:15].contiguous()   # centre tap: sum over all pixels of dout * (1 - out^2)
    w27 = _head_w27(w)
    ap = a if want_lo else SplitPlanes(a.hi, None)
    dw27 = conv2d_tc_wgrad(g, ap, (32, Cc, 1, 1), 1, 0, passes=passes)
    dw = dw27[:27].reshape(3, 3, 3, Cc).permute(2, 3, 0, 1).contiguous()
    pwT = prep_conv_weight_ex(w27, want_lo, transpose=True, rows=Cc)
    dx, amax = conv2d_tc(g, pwT, None, 1, 1, 1, 0, (H, W), passes=passes, transposed=True, tag="head_dgrad",
                         act_mask=a.hi, want_amax=True)
    return dx, amax, dw, db


# ---------------------------------------------------------------------------------------------
# style encoder / discriminator layers (fp32 NHWC)
# ---------------------------------------------------------------------------------------------
def conv2d_direct(x, w_khwc, bias, stride=1, pad=1, ups=0, lrelu=False):
    """x NHWC [B,Hi,Wi,Cin]; w_khwc [KH,KW,Cin,Cout] -> NHWC [B,Ho,Wo,Cout]."""
    _chk_cuda(x, w_khwc, bias)
    B, Hi, Wi, Cin = x.shape
    KH, KW, Cin2, Cout = w_khwc.shape
    assert Cin2 == Cin, (Cin2, Cin)
    Ho = ((Hi << ups) + 2 * pad - KH) // stride + 1
    Wo = ((Wi << ups) + 2 * pad - KW) // stride + 1
    out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().dsee_conv2d_direct_fwd(_p(x), _p(w_khwc), _p(bias), _p(out), B, Hi, Wi, Cin,
                                                  Cout, KH, KW, stride, pad, ups, int(lrelu),
                                                  _stream()))
    return out


def instance_norm(x, act, eps=1e-5):
    """InstanceNorm2d(affine=False) + activation (0 none, 1 lrelu 0.2, 2 tanh). x NHWC."""
    _chk_cuda(x)
    B, H, W, Cc = x.shape
    out = torch.empty_like(x)
    mean = torch.empty((B, Cc), dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    lib = _lib.load()
    ws = torch.empty(lib.dsee_instance_norm_workspace_bytes(B, H * W, Cc), dtype=torch.uint8,
                     device=x.device)
    _lib.check(lib.dsee_instance_norm_fwd(_p(x), _p(out), _p(mean), _p(rstd), _p(ws), B, H * W, Cc,
                                          float(eps), act, _stream()))
    return out, mean, rstd


def region_pool(x, labels, L):
    """extract_style_matrix: x NHWC [B,H,W,C], labels uint8 [B,H,W] -> [B,L,C]."""
    _chk_cuda(x, labels)
    B, H, W, Cc = x.shape
    lib = _lib.load()
    chunks = lib.dsee_region_pool_chunks(H * W)
    ws = torch.empty((B, chunks, L, Cc), dtype=torch.float32, device=x.device)
    style = torch.empty((B, L, Cc), dtype=torch.float32, device=x.device)
    _lib.check(lib.dsee_region_pool_fwd(_p(x), _p(labels), _p(style), _p(ws), B, H * W, Cc, L,
                                        _stream()))
    return style


def nchw_to_nhwc(x, Cp=None):
    _chk_cuda(x)
    B, Cc, H, W = x.shape
    Cp = Cp or Cc
    out = torch.empty((B, H, W, Cp), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().dsee_nchw_to_nhwc(_p(x), _p(out), B, Cc, H, W, Cp, _stream()))
    return out


def disc_input(labels, fake, real, L, Cp):
    """[[onehot|fake];[onehot|real]] NHWC [2B,H,W,Cp] from uint8 labels and NCHW images."""
    _chk_cuda(labels, fake, real)
    B, H, W = labels.shape
    out = torch.empty((2 * B, H, W, Cp), dtype=torch.float32, device=fake.device)
    _lib.check(_lib.load().dsee_disc_input(_p(labels), _p(fake), _p(real), _p(out), B, L, H, W, Cp,
                                           _stream()))
    return out


def avgpool3s2(x):
    _chk_cuda(x)
    B, Hi, Wi, Cc = x.shape
    Ho, Wo = (Hi - 1) // 2 + 1, (Wi - 1) // 2 + 1
    out = torch.empty((B, Ho, Wo, Cc), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().dsee_avgpool3s2_fwd(_p(x), _p(out), B, Hi, Wi, Cc, _stream()))
    return out


# ---------------------------------------------------------------------------------------------
# style encoder / discriminator backward, and the autograd nodes that tie forward and backward
# together (fp32 NHWC tensors, so PyTorch's autograd carries them between layers and into the
# loss code unchanged)
# ---------------------------------------------------------------------------------------------
def act_bwd(dy, out, act):
    _chk_cuda(dy, out)
    dx = torch.empty_like(dy)
    _lib.check(_lib.load().dsee_act_bwd(_p(dy), _p(out), _p(dx), dy.numel(), act, _stream()))
    return dx


def conv2d_direct_dgrad(dy, w_khwc, x_shape, stride, pad, ups):
    _chk_cuda(dy, w_khwc)
    B, Hi, Wi, Cin = x_shape
    KH, KW, _, Cout = w_khwc.shape
    dx = torch.empty(x_shape, dtype=torch.float32, device=dy.device)
    _lib.check(_lib.load().dsee_conv2d_direct_dgrad(_p(dy), _p(w_khwc), _p(dx), B, Hi, Wi, Cin, Cout,
                                                    KH, KW, stride, pad, ups, _stream()))
    return dx


def conv2d_direct_wgrad(x, dy, w_shape, stride, pad, ups):
    _chk_cuda(x, dy)
    B, Hi, Wi, Cin = x.shape
    KH, KW, _, Cout = w_shape
    _, Ho, Wo, _ = dy.shape
    lib = _lib.load()
    ws = torch.empty(lib.dsee_conv2d_direct_wgrad_workspace_floats(B, Ho, Wo, Cin, Cout, KH, KW),
                     dtype=torch.float32, device=x.device)
    dw = torch.empty(tuple(w_shape), dtype=torch.float32, device=x.device)
    _lib.check(lib.dsee_conv2d_direct_wgrad(_p(x), _p(dy), _p(dw), _p(ws), B, Hi, Wi, Cin, Cout, KH, KW,
                                            stride, pad, ups, _stream()))
    return dw


def channel_sum(x):
    """[..., C] -> [C] sums over all leading dims."""
    _chk_cuda(x)
    Cc = x.shape[-1]
    npix = x.numel() // Cc
    lib = _lib.load()
    ws = torch.empty((lib.dsee_channel_sum_chunks(npix), Cc), dtype=torch.float32, device=x.device)
    out = torch.empty(Cc, dtype=torch.float32, device=x.device)
    _lib.check(lib.dsee_channel_sum(_p(x), npix, Cc, _p(ws), _p(out), _stream()))
    return out


def split_f16_ups2(x, want_lo=True):
    """fp32 NHWC [B,H,W,C] -> fp16 planes [B,2H,2W,C] (nearest 2x upsample materialised)."""
    _chk_cuda(x)
    B, H, W, Cc = x.shape
    hi = torch.empty((B, 2 * H, 2 * W, Cc), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi) if want_lo else None
    _lib.check(_lib.load().dsee_split_f16_ups2(_p(x), _p(hi), _p(lo), B, H, W, Cc, _stream()))
    return SplitPlanes(hi, lo)


def fold2x2(x):
    _chk_cuda(x)
    B, H, W, Cc = x.shape
    out = torch.empty((B, H // 2, W // 2, Cc), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().dsee_fold2x2(_p(x), _p(out), B, H // 2, W // 2, Cc, _stream()))
    return out


def prep_conv_weight_ex(w, want_lo=True, transpose=False, rows=None):
    """fp32 [N,C,KH,KW] -> planes [N, KH*KW*Cp] (or [rows >= C, KH*KW*Np] transposed); K blocks
    padded to 64 with zeros, extra rows zero."""
    _chk_cuda(w)
    N, Cin, KH, KW = w.shape
    r, cols = (Cin, N) if transpose else (N, Cin)
    rows = max(rows or r, r)
    kp = (cols + 63) // 64 * 64
    alloc = torch.zeros if rows > r else torch.empty
    hi = alloc((rows, KH * KW * kp), dtype=torch.float16, device=w.device)
    lo = alloc((rows, KH * KW * kp), dtype=torch.float16, device=w.device) if want_lo else None
    inv = torch.empty(2, dtype=torch.float32, device=w.device)
    _lib.check(_lib.load().dsee_prep_conv_weight_ex(_p(w), _p(hi), _p(lo), _p(inv), N, Cin, KH, KW,
                                                    int(transpose), _stream()))
    return PreparedWeight(hi, lo, inv, rows, cols)


def conv2d_tc(a, pw, bias, KH, KW, stride, pad, out_hw, passes=3, lrelu=False, transposed=False,
              tag="conv2d_tc", act_mask=None, want_amax=False):
    """General strided conv (or its backward-data when transposed) on the tcgen05 kernel.
    a: SplitPlanes / GradPlanes NHWC; out_hw: output (Ho, Wo).  act_mask (stride 1): fp16 plane of the
    forward activation, the result is multiplied by LeakyReLU'(0.2) of it.  want_amax -> (out, max|out|)."""
    _chk_cuda(a.hi, a.lo, bias)
    B, Hi, Wi, Ci = a.hi.shape
    args = _lib.Conv2dTCArgs()
    args.B, args.Hi, args.Wi = B, Hi, Wi
    args.a_hi, args.a_lo, args.Ci = a.hi.data_ptr(), (a.lo.data_ptr() if a.lo is not None else 0), Ci
    inv = getattr(a, "inv_scale", None)
    args.a_inv_scale = inv.data_ptr() if inv is not None else 0
    args.KH, args.KW, args.stride, args.pad = KH, KW, stride, pad
    args.w_hi, args.w_lo = pw.hi.data_ptr(), (pw.lo.data_ptr() if pw.lo is not None else 0)
    args.w_inv_scale = pw.inv_scale.data_ptr()
    args.n_total, args.passes, args.transposed = pw.n_total, passes, int(transposed)
    args.Ho, args.Wo = out_hw
    out = torch.empty((B, out_hw[0], out_hw[1], pw.n_total), dtype=torch.float32, device=a.hi.device)
    epi = _lib.ConvEpilogue()
    epi.bias = bias.data_ptr() if bias is not None else 0
    epi.out = out.data_ptr()
    epi.lrelu = int(lrelu)
    epi.act_mask = act_mask.data_ptr() if act_mask is not None else 0
    amax = torch.empty(1, dtype=torch.float32, device=a.hi.device) if want_amax else None
    epi.amax_out = amax.data_ptr() if amax is not None else 0
    npix = B * (Hi * Wi if transposed else out_hw[0] * out_hw[1])
    flops = 2.0 * KH * KW * pw.cin * pw.n_total * npix
    _timed(tag, flops, lambda: _lib.check(_lib.load().dsee_conv2d_tc(C.byref(args), C.byref(epi),
                                                                     _stream())))
    return (out, amax) if want_amax else out


def conv2d_tc_wgrad(dy, a, w_shape, stride, pad, passes=3):
    """-> dW [N, Ci_w, KH, KW] for the forward conv with activation planes `a`."""
    _chk_cuda(dy.hi, dy.lo, a.hi, a.lo)
    N, Cw, KH, KW = w_shape
    B, Ho, Wo, _ = dy.hi.shape
    _, Hi, Wi, Ci = a.hi.shape
    lib = _lib.load()
    cpad = (Ci + 63) // 64 * 64
    ws = torch.empty(lib.dsee_conv2d_tc_wgrad_workspace_floats(B, Ho, Wo, N, Ci, KH, KW),
                     dtype=torch.float32, device=dy.hi.device)
    dw = torch.empty((N, cpad, KH, KW), dtype=torch.float32, device=dy.hi.device)
    flops = 2.0 * KH * KW * Cw * N * B * Ho * Wo
    _timed("wgrad2d", flops, lambda: _lib.check(lib.dsee_conv2d_tc_wgrad(
        _p(dy.hi), _p(dy.lo), _p(getattr(dy, "inv_scale", None)), _p(a.hi), _p(a.lo),
        _p(getattr(a, "inv_scale", None)), B, Ho, Wo, Hi, Wi, N, Ci, KH, KW, stride, pad, passes,
        _p(ws), _p(dw), _stream())))
    return dw[:, :Cw].contiguous() if cpad != Cw else dw


def tc_conv_eligible(cin_x, cout):
    """Layers the tcgen05 path takes: enough channels for a K block to make sense and an output
    width the epilogue stores in 32-column chunks (the RGB stem of the encoder and the
    discriminator's 1-channel prediction conv stay on the direct fp32 kernels)."""
    return cin_x % 8 == 0 and cin_x >= 16 and cout % 32 == 0 and cout in (32, 64, 128, 256, 512, 1024)


class Conv2dTCFn(torch.autograd.Function):
    """Encoder / discriminator conv layer (3x3 or 4x4, stride 1 / 2, optional folded 2x upsample of
    the input, optional fused bias + LeakyReLU) on the tcgen05 implicit-GEMM kernels.
    x fp32 NHWC [B,Hi,Wi,Cx]; w fp32 [N,Cw,KH,KW] (PyTorch layout, Cw <= Cx: x may carry zero
    padding channels); out fp32 NHWC."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, pad, ups, lrelu, fwd_passes=None):
        from .config import config
        # 1-pass mode: the forward GEMM runs `fwd_passes` (default config.ed_fwd_passes = 3: the style
        # encoder's output feeds the generator, whose output carries the 1e-3 bound; the discriminator,
        # which only feeds the losses, passes 1); backward GEMMs at config.passes
        passes_f = 3 if config.passes == 3 else (fwd_passes or config.ed_fwd_passes)
        passes = config.passes
        want_lo = passes_f == 3
        N, Cw, KH, KW = w.shape
        planes = split_f16_ups2(x, want_lo) if ups else split_f16(x, want_lo)
        B, Hu, Wu, Cx = planes.hi.shape
        Ho, Wo = (Hu + 2 * pad - KH) // stride + 1, (Wu + 2 * pad - KW) // stride + 1
        pw = prep_conv_weight_ex(w.contiguous(), want_lo)
        out = conv2d_tc(planes, pw, bias, KH, KW, stride, pad, (Ho, Wo), passes=passes_f, lrelu=lrelu)
        want_lo = passes == 3
        ctx.cfg = (stride, pad, ups, lrelu, bias is not None, passes, want_lo, tuple(x.shape))
        ctx.planes = planes
        ctx.save_for_backward(w, out if lrelu else None)
        return out

    @staticmethod
    def backward(ctx, dy):
        w, out = ctx.saved_tensors
        stride, pad, ups, lrelu, has_bias, passes, want_lo, xshape = ctx.cfg
        planes = ctx.planes
        ctx.planes = None
        N, Cw, KH, KW = w.shape
        dy = dy.contiguous()
        if lrelu:
            dy = act_bwd(dy, out, 3 if lrelu == 2 else 1)
        g, sums = grad_prep(dy, want_lo=want_lo)
        dx = dw = db = None
        if ctx.needs_input_grad[1]:
            dw = conv2d_tc_wgrad(g, planes, tuple(w.shape), stride, pad, passes=passes)
        if has_bias and ctx.needs_input_grad[2]:
            db = sums[0]
        if ctx.needs_input_grad[0]:
            B, Hu, Wu, Cx = planes.hi.shape
            rows = (Cx + 31) // 32 * 32
            pwT = prep_conv_weight_ex(w.contiguous(), want_lo, transpose=True, rows=rows)
            dx = conv2d_tc(g, pwT, None, KH, KW, stride, pad, (Hu, Wu), passes=passes, transposed=True,
                           tag="dgrad2d")
            if rows != Cx:
                dx = dx[..., :Cx].contiguous()
            if ups:
                dx = fold2x2(dx)
        return dx, dw, db, None, None, None, None, None


def conv_layer(x, w, bias, stride, pad, ups=0, lrelu=False, fwd_passes=None):
    """Dispatch of an encoder / discriminator conv: tcgen05 path when eligible, else the direct fp32
    kernels. w in PyTorch layout [N,Cw,KH,KW] (autograd tensor).  fwd_passes: operand passes of the
    forward GEMM in 1-pass mode (None = config.ed_fwd_passes)."""
    if tc_conv_eligible(x.shape[3], w.shape[0]):
        return Conv2dTCFn.apply(x, w, bias, stride, pad, ups, lrelu, fwd_passes)
    if tc_conv_eligible(x.shape[3], 32) and w.shape[0] < 32:
        # narrow outputs (the discriminator's 1-channel prediction conv): zero-pad the filter bank to
        # the 32 columns the tensor-core epilogue stores and slice; autograd un-pads the gradients
        n = w.shape[0]
        wp = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, 32 - n))
        bp = torch.nn.functional.pad(bias, (0, 32 - n)) if bias is not None else None
        return Conv2dTCFn.apply(x, wp, bp, stride, pad, ups, lrelu, fwd_passes)[..., :n].contiguous()
    wk = w.permute(2, 3, 1, 0)
    if x.shape[3] > wk.shape[2]:
        wk = torch.nn.functional.pad(wk, (0, 0, 0, x.shape[3] - wk.shape[2]))
    return Conv2dDirectFn.apply(x, wk.contiguous(), bias, stride, pad, ups, lrelu)


class Conv2dDirectFn(torch.autograd.Function):
    """conv2d_direct (+ fused LeakyReLU) with its backward-data / weight-gradient kernels."""

    @staticmethod
    def forward(ctx, x, w_khwc, bias, stride, pad, ups, lrelu):
        w_khwc = w_khwc.contiguous()
        out = conv2d_direct(x, w_khwc, bias, stride=stride, pad=pad, ups=ups, lrelu=lrelu)
        ctx.cfg = (stride, pad, ups, lrelu, bias is not None)
        ctx.save_for_backward(x, w_khwc, out if lrelu else None)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w, out = ctx.saved_tensors
        stride, pad, ups, lrelu, has_bias = ctx.cfg
        dy = dy.contiguous()
        if lrelu:
            dy = act_bwd(dy, out, 3 if lrelu == 2 else 1)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = conv2d_direct_dgrad(dy, w, tuple(x.shape), stride, pad, ups)
        if ctx.needs_input_grad[1]:
            dw = conv2d_direct_wgrad(x, dy, tuple(w.shape), stride, pad, ups)
        if has_bias and ctx.needs_input_grad[2]:
            db = channel_sum(dy)
        return dx, dw, db, None, None, None, None


class InstanceNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        out, mean, rstd = instance_norm(x, act)
        ctx.act = act
        ctx.save_for_backward(x, mean, rstd)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, mean, rstd = ctx.saved_tensors
        dout = dout.contiguous()
        B, H, W, Cc = x.shape
        dx = torch.empty_like(x)
        sums = torch.empty((B, Cc, 2), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        ws = torch.empty(lib.dsee_instance_norm_workspace_bytes(B, H * W, Cc), dtype=torch.uint8,
                         device=x.device)
        _lib.check(lib.dsee_instance_norm_bwd(_p(x), _p(dout), _p(mean), _p(rstd), _p(dx), _p(sums),
                                              _p(ws), B, H * W, Cc, ctx.act, _stream()))
        return dx, None


class RegionPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, labels, L):
        ctx.save_for_backward(labels)
        ctx.shape = tuple(x.shape)
        ctx.L = L
        return region_pool(x, labels, L)

    @staticmethod
    def backward(ctx, dstyle):
        (labels,) = ctx.saved_tensors
        B, H, W, Cc = ctx.shape
        dstyle = dstyle.contiguous()
        dx = torch.empty(ctx.shape, dtype=torch.float32, device=dstyle.device)
        _lib.check(_lib.load().dsee_region_pool_bwd(_p(dstyle), _p(labels), _p(dx), B, H * W, Cc, ctx.L,
                                                    _stream()))
        return dx, None, None


class AvgPool3s2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return avgpool3s2(x)

    @staticmethod
    def backward(ctx, dout):
        B, Hi, Wi, Cc = ctx.shape
        dout = dout.contiguous()
        din = torch.empty(ctx.shape, dtype=torch.float32, device=dout.device)
        _lib.check(_lib.load().dsee_avgpool3s2_bwd(_p(dout), _p(din), B, Hi, Wi, Cc, _stream()))
        return din


class MaxPool2Fn(torch.autograd.Function):
    """nn.MaxPool2d(2, 2) on NHWC fp32 (VGG19)."""

    @staticmethod
    def forward(ctx, x):
        _chk_cuda(x)
        B, Hi, Wi, Cc = x.shape
        out = torch.empty((B, Hi // 2, Wi // 2, Cc), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().dsee_maxpool2_fwd(_p(x), _p(out), B, Hi, Wi, Cc, _stream()))
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        B, Hi, Wi, Cc = x.shape
        dout = dout.contiguous()
        din = torch.empty_like(x)
        _lib.check(_lib.load().dsee_maxpool2_bwd(_p(x), _p(dout), _p(din), B, Hi, Wi, Cc, _stream()))
        return din


class DiscInputFn(torch.autograd.Function):
    """SRModel.discriminate's input assembly; the gradient goes to the fake image only."""

    @staticmethod
    def forward(ctx, labels, fake, real, L, Cp):
        ctx.meta = (tuple(fake.shape), L, Cp)
        return disc_input(labels, fake, real, L, Cp)

    @staticmethod
    def backward(ctx, dx):
        (B, _, H, W), L, Cp = ctx.meta
        dx = dx.contiguous()
        dfake = torch.empty((B, 3, H, W), dtype=torch.float32, device=dx.device)
        _lib.check(_lib.load().dsee_disc_input_bwd(_p(dx), _p(dfake), B, L, H, W, Cp, _stream()))
        return None, dfake, None, None, None


# ---------------------------------------------------------------------------------------------
# parameter-side fusions: spectral normalisation and the modulation-weight assembly
# ---------------------------------------------------------------------------------------------
class SpectralWeightFn(torch.autograd.Function):
    """W_eff = W_orig / sigma with torch.nn.utils.spectral_norm's semantics (one power iteration in
    training mode, updating the u / v buffers in place; stored u, v in eval mode; u, v are constants
    for autograd) in 3-5 kernels instead of ~12 ATen launches."""

    @staticmethod
    def forward(ctx, w_orig, u, v, training, eps, keep_for_backward=True, pre=None):
        """pre: (w_eff, sigma2, u_saved, v_saved) already computed for this layer by
        spectral_prepass (one batched launch sequence per network); then only the graph is wired."""
        if pre is not None:
            w_eff, sig, u_s, v_s = pre
            if keep_for_backward and ctx.needs_input_grad[0]:
                ctx.save_for_backward(w_eff, u_s, v_s, sig)
            return w_eff
        _chk_cuda(w_orig, u, v)
        N = w_orig.shape[0]
        K = w_orig.numel() // N
        lib = _lib.load()
        ws = torch.empty(lib.dsee_spectral_workspace_floats(N, K), dtype=torch.float32, device=w_orig.device)
        sig = torch.empty(2, dtype=torch.float32, device=w_orig.device)
        w_eff = torch.empty_like(w_orig)
        _lib.check(lib.dsee_spectral_weight_fwd(_p(w_orig), _p(u), _p(v), N, K, int(training), float(eps),
                                                _p(ws), _p(sig), _p(w_eff), _stream()))
        if keep_for_backward and ctx.needs_input_grad[0]:
            # u / v are overwritten by the next forward; the backward needs this forward's values
            # (callers pass keep_for_backward=False under no_grad: grad mode is invisible in here)
            ctx.save_for_backward(w_eff, u.clone(), v.clone(), sig)
        return w_eff

    @staticmethod
    def backward(ctx, dw_eff):
        w_eff, u, v, sig = ctx.saved_tensors
        dw_eff = dw_eff.contiguous()
        N = w_eff.shape[0]
        K = w_eff.numel() // N
        ws = torch.empty(128, dtype=torch.float64, device=w_eff.device)
        dw = torch.empty_like(w_eff)
        _lib.check(_lib.load().dsee_spectral_weight_bwd(_p(dw_eff), _p(w_eff), _p(u), _p(v), _p(sig), N, K,
                                                        _p(ws), _p(dw), _stream()))
        return dw, None, None, None, None, None, None


def spectral_prepass(convs, training, eps, keep_for_backward):
    """SpectralWeightFn's forward arithmetic for a list of spectral-normalised layers in ONE batched
    launch sequence (dsee_spectral_weight_fwd_batched): power iteration (training), sigma, W / sigma.
    -> list of (w_eff shaped like weight_orig, sigma2 [2], u_saved | None, v_saved | None), all views
    of a few flat buffers."""
    lib = _lib.load()
    n = len(convs)
    shapes = [tuple(c.weight_orig.shape) for c in convs]
    NK = [(sh[0], int(torch.Size(sh).numel() // sh[0])) for sh in shapes]
    dev = convs[0].weight_orig.device
    tot_w = sum(N * K for N, K in NK)
    tot_ws = sum(lib.dsee_spectral_workspace_floats(N, K) for N, K in NK)
    tot_uv = sum(N + K for N, K in NK) if keep_for_backward else 0
    flat = torch.empty(tot_w + tot_ws + 2 * n + tot_uv, dtype=torch.float32, device=dev)
    items = (_lib.SnItem * n)()
    base = flat.data_ptr()
    o_w, o_ws, o_sig, o_uv = 0, tot_w, tot_w + tot_ws, tot_w + tot_ws + 2 * n
    out = []
    for i, (c, (N, K)) in enumerate(zip(convs, NK)):
        _chk_cuda(c.weight_orig, c.weight_u, c.weight_v)
        it = items[i]
        it.w_orig, it.u, it.v = c.weight_orig.data_ptr(), c.weight_u.data_ptr(), c.weight_v.data_ptr()
        it.w_eff, it.sigma2, it.workspace = base + 4 * o_w, base + 4 * o_sig, base + 4 * o_ws
        it.N, it.K = N, K
        w_eff = flat[o_w:o_w + N * K].view(shapes[i])
        sig = flat[o_sig:o_sig + 2]
        u_s = v_s = None
        if keep_for_backward:
            it.u_saved, it.v_saved = base + 4 * o_uv, base + 4 * (o_uv + N)
            u_s, v_s = flat[o_uv:o_uv + N], flat[o_uv + N:o_uv + N + K]
            o_uv += N + K
        else:
            it.u_saved, it.v_saved = 0, 0
        out.append((w_eff, sig, u_s, v_s))
        o_w += N * K
        o_ws += lib.dsee_spectral_workspace_floats(N, K)
        o_sig += 2
    _lib.check(lib.dsee_spectral_weight_fwd_batched(items, n, int(training), float(eps), _stream()))
    return out


class ModWeightFn(torch.autograd.Function):
    """K1's fused modulation weight (interleaved gamma|beta rows over [seg | style] columns, alpha
    blend folded in) and its two bias vectors from the reference's separate parameters: one kernel
    forward, two backward.  Inputs that do not exist for a layer kind are None."""

    @staticmethod
    def forward(ctx, plus_one, wg, wb, wsg, wsb, bg, bb, bsg, bsb, ag, ab):
        ref = wg if wg is not None else wsg
        Cc = ref.shape[0]
        c1 = wg.shape[1] if wg is not None else 0
        c2 = wsg.shape[1] if wsg is not None else 0
        tensors = [wg, wb, wsg, wsb, bg, bb, bsg, bsb, ag, ab]
        tensors = [t.contiguous() if t is not None else None for t in tensors]
        _chk_cuda(*tensors)
        args = _modweight_args(tensors, Cc, c1, c2, plus_one)
        wm = torch.empty((2 * Cc, c1 + c2, 3, 3), dtype=torch.float32, device=ref.device)
        gbias = torch.empty(Cc, dtype=torch.float32, device=ref.device)
        bbias = torch.empty(Cc, dtype=torch.float32, device=ref.device)
        _lib.check(_lib.load().dsee_modweight_fwd(C.byref(args), _p(wm), _p(gbias), _p(bbias), _stream()))
        ctx.meta = (Cc, c1, c2, plus_one)
        ctx.save_for_backward(*[t for t in tensors if t is not None])
        ctx.present = [t is not None for t in tensors]
        return wm, gbias, bbias

    @staticmethod
    def backward(ctx, dwm, dgb, dbb):
        Cc, c1, c2, plus_one = ctx.meta
        it = iter(ctx.saved_tensors)
        tensors = [next(it) if p else None for p in ctx.present]
        args = _modweight_args(tensors, Cc, c1, c2, plus_one)
        dev = dwm.device
        grads = [torch.empty_like(t) if t is not None else None for t in tensors]
        g = _lib.ModWeightGrads()
        for i in range(2):
            g.dw_seg[i] = grads[0 + i].data_ptr() if grads[0 + i] is not None else 0
            g.dw_sty[i] = grads[2 + i].data_ptr() if grads[2 + i] is not None else 0
            g.db_seg[i] = grads[4 + i].data_ptr() if grads[4 + i] is not None else 0
            g.db_sty[i] = grads[6 + i].data_ptr() if grads[6 + i] is not None else 0
            g.dalpha[i] = grads[8 + i].data_ptr() if grads[8 + i] is not None else 0
        lib = _lib.load()
        ws = torch.empty(lib.dsee_modweight_bwd_workspace_bytes(Cc, c1, c2), dtype=torch.uint8, device=dev)
        _lib.check(lib.dsee_modweight_bwd(C.byref(args), _p(dwm.contiguous()), _p(dgb.contiguous()),
                                          _p(dbb.contiguous()), C.byref(g), _p(ws), _stream()))
        return (None, *grads)


def _modweight_args(tensors, Cc, c1, c2, plus_one):
    wg, wb, wsg, wsb, bg, bb, bsg, bsb, ag, ab = tensors
    a = _lib.ModWeightArgs()
    ptr = lambda t: t.data_ptr() if t is not None else 0
    a.w_seg[0], a.w_seg[1] = ptr(wg), ptr(wb)
    a.w_sty[0], a.w_sty[1] = ptr(wsg), ptr(wsb)
    a.b_seg[0], a.b_seg[1] = ptr(bg), ptr(bb)
    a.b_sty[0], a.b_sty[1] = ptr(bsg), ptr(bsb)
    a.alpha[0], a.alpha[1] = ptr(ag), ptr(ab)
    a.C, a.c1, a.c2, a.plus_one = Cc, c1, c2, int(plus_one)
    return a
