"""Kernel-level parity (GPU): each C-ABI kernel against a plain PyTorch fp32 statement of the same
op (TF32 disabled).  Tolerances are written next to each check."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 64, 256), (2, 16, 16, 128, 512),
                                            (1, 32, 32, 512, 512), (1, 24, 40, 64, 32)])
@pytest.mark.parametrize("passes", [1, 3])
def test_conv3x3_matches_torch(B, H, W, Cin, Cout, passes):
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + H + Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    ref = F.conv2d(x, w, bias, padding=1)
    a = ops.split_f16(_nhwc(x))
    pw = ops.prep_conv_weight(w)
    out = _nchw(ops.conv3x3([a], pw, bias, passes=passes))
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    # 3-pass split fp16 is fp32-class; 1-pass is TF32-class (10-bit mantissa operands)
    tol = 5e-5 * scale if passes == 3 else 4e-3 * scale
    assert err <= tol, (err, scale)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 128, 256), (2, 24, 40, 256, 512), (1, 32, 32, 512, 512)])
def test_conv3x3_fp8_correction_matches_torch(B, H, W, Cin, Cout):
    """passes == 2: one fp16 pass + the fp8 correction GEMM (kind::f8f6f4) for both operand-rounding
    terms.  The fp8 planes are built here with torch's float8 casts (the kernels that produce them in
    the model are checked against the same casts in test_spade_modulate_fp8_planes)."""
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + H + Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    ref = F.conv2d(x, w, bias, padding=1)
    a = ops.split_f16(_nhwc(x))
    xn = _nhwc(x)
    lo = xn - a.hi.float()
    f8 = ((lo * 256).to(torch.float8_e5m2).view(torch.uint8), xn.to(torch.float8_e5m2).view(torch.uint8))
    a = ops.SplitPlanes(a.hi, None, f8)
    pw = ops.prep_conv_weight(w, want_lo=False, want_f8=True)
    # the weight's fp8 companion against torch's casts
    scale = 1.0 / pw.inv_scale[0].item()
    ws = (w * scale).permute(0, 2, 3, 1).reshape(Cout, 9, Cin)
    want_hi = (ws / 256).to(torch.float8_e4m3fn).view(torch.uint8)
    want_lo = (ws - ws.half().float()).to(torch.float8_e4m3fn).view(torch.uint8)
    assert torch.equal(pw.f8[:, :, 0], want_hi) and torch.equal(pw.f8[:, :, 1], want_lo)
    out1 = _nchw(ops.conv3x3([ops.SplitPlanes(a.hi, None)], pw, bias, passes=1))
    out2 = _nchw(ops.conv3x3([a], pw, bias, passes=2))
    s = ref.abs().max().item()
    e1, e2 = (out1 - ref).abs().max().item() / s, (out2 - ref).abs().max().item() / s
    print("1-pass rel err %.3e, fp16 + fp8 correction rel err %.3e" % (e1, e2))
    # the correction removes >= 85 % of the 1-pass error (the fp8 operands carry ~3-6 % relative error)
    assert e2 <= 0.15 * e1 + 2e-6 and e2 <= 6e-5


def test_spade_modulate_fp8_planes():
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    B, H, W, C, nh = 1, 16, 16, 128, 128
    actv = ops.split_f16(torch.randn(B, H, W, nh, generator=g).relu().cuda())
    wm = (torch.randn(2 * C, nh, 3, 3, generator=g) / (3 * nh ** 0.5)).cuda()
    pw = ops.prep_conv_weight(wm)
    x = torch.randn(B, H, W, C, generator=g).cuda()
    one, zero = torch.ones(C).cuda(), torch.zeros(C).cuda()
    a = ops.spade_modulate([actv], pw, x, 0, one, zero, one, zero, passes=3, want_lo=True, want_f8=True)
    t = a.hi.float() + a.lo.float()          # the fp32 activation to ~2^-22
    assert a.f8 is not None and a.f8[0].dtype == torch.uint8
    got_hi = a.f8[1].view(torch.float8_e5m2).float()
    got_lo = a.f8[0].view(torch.float8_e5m2).float() / 256
    # e5m2: 2 mantissa bits -> relative error <= 2^-3 (a half-ulp); below 2^-14 the format is
    # subnormal with an absolute step of 2^-16
    assert ((got_hi - t).abs() <= 0.126 * t.abs() + 2.0 ** -16).all()
    lo = t - a.hi.float()
    assert ((got_lo - lo).abs() <= 0.126 * lo.abs() + 2.0 ** -24).all()


def test_conv3x3_residual_upsample_and_stats():
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(7)
    B, H, W, C = 2, 16, 32, 128
    x = torch.randn(B, C, H, W, generator=g).cuda()
    w = (torch.randn(256, C, 3, 3, generator=g) / (3 * C ** 0.5)).cuda()
    bias = torch.randn(256, generator=g).cuda()
    res = torch.randn(B, 256, H // 2, W // 2, generator=g).cuda()
    ref = F.conv2d(x, w, bias, padding=1) + F.interpolate(res, scale_factor=2, mode="nearest")
    a = ops.split_f16(_nhwc(x))
    pw = ops.prep_conv_weight(w)
    out, part = ops.conv3x3([a], pw, bias, residual=_nhwc(res), res_ups=1, passes=3, want_stats=True)
    assert (_nchw(out) - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    sc, sh, mean, var = ops.bn_finalize(part, B * H * W, 1e-5)
    torch.testing.assert_close(mean, ref.mean(dim=(0, 2, 3)), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(var, ref.var(dim=(0, 2, 3), unbiased=False), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("x_ups", [0, 1])
@pytest.mark.parametrize("two_sources", [False, True])
def test_spade_modulate_matches_torch(x_ups, two_sources):
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(11 + x_ups)
    B, H, W, C, nh = 2, 16, 16, 256, 128
    actv = torch.randn(B, nh, H, W, generator=g).relu().cuda()
    sty = torch.randn(B, nh, H, W, generator=g).cuda()
    Cin = nh * (2 if two_sources else 1)
    wg = (torch.randn(C, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    wb = (torch.randn(C, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    gb = torch.randn(C, generator=g).cuda()
    bb = torch.randn(C, generator=g).cuda()
    x = torch.randn(B, C, H >> x_ups, W >> x_ups, generator=g).cuda()
    rm = torch.randn(C, generator=g).cuda() * 0.1
    rv = (torch.rand(C, generator=g) + 0.5).cuda()
    noise = torch.randn(B, C, H, W, generator=g).cuda()
    nw = torch.randn(C, generator=g).cuda() * 0.1

    inp = torch.cat([actv, sty], 1) if two_sources else actv
    gamma = F.conv2d(inp, wg, gb, padding=1)
    beta = F.conv2d(inp, wb, bb, padding=1)
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if x_ups else x
    xin = xin + nw.view(1, -1, 1, 1) * noise
    xhat = F.batch_norm(xin, rm, rv, training=False, eps=1e-5)
    ref = F.leaky_relu(xhat * (1 + gamma) + beta, 0.2)

    # rows interleaved per 128 channels: [gamma(128) | beta(128)] ...
    wcat = torch.stack([wg.view(C // 128, 128, Cin, 3, 3), wb.view(C // 128, 128, Cin, 3, 3)], 1)
    wcat = wcat.reshape(2 * C, Cin, 3, 3).contiguous()
    pw = ops.prep_conv_weight(wcat)
    srcs = [ops.split_f16(_nhwc(actv))] + ([ops.split_f16(_nhwc(sty))] if two_sources else [])
    sc, sh = ops.bn_eval_affine(rm, rv, 1e-5)
    out = ops.spade_modulate(srcs, pw, _nhwc(x), x_ups, sc, sh, gb + 1.0, bb, noise=_nhwc(noise),
                             noise_w=nw, passes=3)
    got = _nchw(out.hi.float() + out.lo.float())
    err = (got - ref).abs().max().item()
    assert err <= 3e-5 * ref.abs().max().item(), err


def test_label_ops_bit_exact():
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(3)
    B, L, S = 2, 19, 64
    lab = torch.randint(0, L, (B, 1, S, S), generator=g).cuda()
    oh, bad = ops.onehot_from_labels(lab, L)
    ref = torch.zeros(B, L, S, S, device="cuda").scatter_(1, lab, 1.0)
    assert torch.equal(oh, ref) and bad.item() == 0
    back, bad2 = ops.labels_from_onehot(oh)
    assert torch.equal(back.long(), lab[:, 0]) and bad2.item() == 0
    for h in (32, 16, 48, 24):
        small = ops.resize_labels(back, h, h)
        refs = F.interpolate(oh, size=(h, h), mode="nearest").argmax(1)
        assert torch.equal(small.long(), refs)
    # a map that is not one-hot is reported
    oh2 = oh.clone()
    oh2[0, :, 0, 0] = 0.5
    _, bad3 = ops.labels_from_onehot(oh2)
    assert bad3.item() == 1


@pytest.mark.parametrize("S", [32, 72])   # 72: the shared-memory-table kernel also at ups = 0
@pytest.mark.parametrize("ups", [0, 1])
def test_shared_mlp_and_style_gather(ups, S):
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    B, L, nh, d = 2, 19, 128, 128
    lab = torch.randint(0, L, (B, 1, S, S), generator=g).cuda()
    oh = torch.zeros(B, L, S, S, device="cuda").scatter_(1, lab, 1.0)
    w = torch.randn(nh, L, 3, 3, generator=g).cuda() * 0.2
    b = torch.randn(nh, generator=g).cuda() * 0.1
    ref = F.relu(F.conv2d(oh, w, b, padding=1))
    if ups:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
    table = w.permute(2, 3, 1, 0).reshape(9, L, nh).contiguous()
    labels = lab[:, 0].to(torch.uint8).contiguous()
    out = ops.shared_mlp(labels, table, b, ups=ups)
    got = _nchw(out.hi.float() + out.lo.float())
    assert (got - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())

    style = torch.randn(B, L, d, generator=g).cuda()
    ref_s = (style[:, :, :, None, None] * oh[:, :, None]).sum(1)
    sm = ops.style_gather(labels, style)
    got_s = _nchw(sm.hi.float() + sm.lo.float())
    assert (got_s - ref_s).abs().max().item() <= 1e-6 * ref_s.abs().max().item()


def test_stem_head_and_bn_stats():
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(9)
    B, S, C = 2, 16, 512
    x = torch.rand(B, 3, S, S, generator=g).cuda() * 2 - 1
    w = torch.randn(C, 3, 3, 3, generator=g).cuda() * 0.2
    b = torch.randn(C, generator=g).cuda() * 0.1
    ref = F.conv2d(x, w, b, padding=1)
    out = ops.stem(x, w, b)
    assert (_nchw(out) - ref).abs().max().item() <= 1e-5

    feat = torch.randn(B, 24, 20, C, generator=g).cuda()
    wi = torch.randn(3, C, 3, 3, generator=g).cuda() * 0.01
    bi = torch.randn(3, generator=g).cuda() * 0.1
    refh = torch.tanh(F.conv2d(F.leaky_relu(_nchw(feat), 0.2), wi, bi, padding=1))
    outh = ops.head(feat, wi, bi)
    assert (outh - refh).abs().max().item() <= 2e-5

    noise = torch.randn(B, 48, 40, C, generator=g).cuda()
    nw = torch.randn(C, generator=g).cuda() * 0.3
    part = ops.bn_stats(feat, x_ups=1, noise=noise, noise_w=nw)
    rm = torch.zeros(C, device="cuda")
    rv = torch.ones(C, device="cuda")
    sc, sh, mean, var = ops.bn_finalize(part, B * 48 * 40, 1e-5, 0.1, rm, rv)
    full = F.interpolate(_nchw(feat), scale_factor=2, mode="nearest") + nw.view(1, -1, 1, 1) * _nchw(noise)
    bn = torch.nn.BatchNorm2d(C, affine=False).cuda().train()
    bn(full)
    torch.testing.assert_close(mean, full.mean(dim=(0, 2, 3)), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(var, full.var(dim=(0, 2, 3), unbiased=False), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rm, bn.running_mean, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv, bn.running_var, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(sc, 1 / torch.sqrt(var + 1e-5), rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# backward kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 128, 128), (2, 16, 16, 128, 256),
                                            (1, 24, 40, 512, 512), (2, 12, 20, 64, 128)])
@pytest.mark.parametrize("passes", [1, 3])
def test_wgrad_and_dgrad_match_autograd(B, H, W, Cin, Cout, passes):
    """conv3x3_wgrad (scaled fp16 gradient planes x fp16 activation planes, MN-major tcgen05 operands) and
    backward-data (the forward kernel on gradient planes with the transposed filter) against
    torch autograd of F.conv2d."""
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(H * 7 + Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda().requires_grad_(True)
    dy = torch.randn(B, Cout, H, W, generator=g).cuda()
    F.conv2d(x, w, None, padding=1).backward(dy)
    a = ops.split_f16(_nhwc(x.detach()))
    gp, sums = ops.grad_prep(_nhwc(dy))
    torch.testing.assert_close(sums[0], dy.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)
    dw = ops.conv3x3_wgrad(gp, a, passes=passes)
    tol = 5e-5 if passes == 3 else 2e-2
    s = w.grad.abs().max().item()
    err = (dw - w.grad).abs().max().item()
    assert err <= tol * s, ("wgrad", err, s)
    pwT = ops.prep_conv_weight(w.detach(), transpose=True)
    dx = _nchw(ops.conv3x3([gp], pwT, None, passes=passes))
    s = x.grad.abs().max().item()
    err = (dx - x.grad).abs().max().item()
    assert err <= tol * s, ("dgrad", err, s)


@pytest.mark.parametrize("B,H,W", [(3, 16, 16), (2, 24, 40), (8, 8, 16)])
def test_per_image_weights_and_wgrad(B, H, W):
    """The folded-style form of a SEAN layer (config.fold_style): K1 with one weight matrix per image
    over [actv | one-hot] sources equals the gathered style_map form, and the per-image weight gradient
    equals autograd of a per-image conv."""
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + H)
    C, nh, L, d = 128, 128, 19, 128
    labels = torch.randint(0, L, (B, H, W), generator=g, dtype=torch.uint8).cuda()
    actv = torch.randn(B, H, W, nh, generator=g).relu().cuda()
    style = (torch.rand(B, L, d, generator=g) * 2 - 1).cuda()
    wm = (torch.randn(2 * C, nh + d, 3, 3, generator=g) / (3 * (nh + d) ** 0.5)).cuda()
    x = torch.randn(B, H, W, C, generator=g).cuda()
    one, zero = torch.ones(C).cuda(), torch.zeros(C).cuda()
    a_pl = ops.split_f16(actv)
    # gathered form
    smap = ops.style_gather(labels, style)
    ref = ops.spade_modulate([a_pl, smap], ops.prep_conv_weight(wm), x, 0, one, zero, one, zero, passes=3)
    # folded form
    from deepsee_b200.deepsee_models.networks.normalization import fold_style_weight
    Wa, Ws = fold_style_weight(wm, style)
    oh = ops.onehot_planes(labels)
    oh = ops.SplitPlanes(oh.hi, torch.zeros_like(oh.hi))
    got = ops.spade_modulate([a_pl, oh], ops.prep_mod_weight_batched(Wa, Ws), x, 0, one, zero, one, zero,
                             passes=3)
    rv = ref.hi.float() + ref.lo.float()
    e = ((got.hi.float() + got.lo.float()) - rv).abs().max().item() / rv.abs().max().item()
    print("folded vs gathered K1 (3-pass) max-abs / max|ref| %.3e" % e)
    assert e < 2e-5   # both fp32-class; the two forms round different intermediate tensors
    # per-image weight gradient vs autograd of per-image convs over [actv | one-hot]
    dy = torch.randn(B, H, W, 2 * C, generator=g).cuda() * 1e-2
    gp, _ = ops.grad_prep(dy)
    dw = ops.conv3x3_wgrad_per_image(gp, [a_pl, oh], passes=3)
    src = torch.cat([actv, oh.hi.float()], 3).permute(0, 3, 1, 2)
    for b in range(B):
        w = torch.zeros(2 * C, nh + 64, 3, 3, device="cuda", requires_grad=True)
        F.conv2d(src[b:b + 1], w, padding=1).backward(dy[b:b + 1].permute(0, 3, 1, 2))
        err = (dw[b] - w.grad).abs().max().item() / w.grad.abs().max().item()
        assert err < 2e-4, (b, err)
    # backward-data with a 128-row transposed weight (N = 128 < BLOCK_N)
    pwT = ops.prep_conv_weight(Wa, transpose=True)
    dsrc = ops.conv3x3([gp], pwT, None, passes=3)
    a_ref = actv.permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.conv2d(a_ref, Wa, padding=1).backward(dy.permute(0, 3, 1, 2))
    err = (dsrc.permute(0, 3, 1, 2) - a_ref.grad).abs().max().item() / a_ref.grad.abs().max().item()
    assert err < 1e-4, err


@pytest.mark.parametrize("B,H,W,two", [(2, 32, 32, False), (1, 48, 80, True)])
def test_subpixel_form_matches_the_upsampled_conv(B, H, W, two):
    """Layers above max_fm_size (normalization.py:188-190,275-277): K1 over the half-resolution
    activation with the collapsed 2x2 filters (config.subpixel) equals K1 over the materialised
    nearest-2x upsample, and its two backward GEMMs equal autograd of conv(upsample(a), W)."""
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + H + W)
    C, nh = 128, 128
    cin = nh * (2 if two else 1)
    a_low = torch.randn(B, H // 2, W // 2, nh, generator=g).relu().cuda()
    wm = (torch.randn(2 * C, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).cuda()
    x = torch.randn(B, H // 2, W // 2, C, generator=g).cuda()       # read through the folded upsample too
    one, zero = torch.ones(C).cuda(), torch.zeros(C).cuda()
    pl_low = ops.split_f16(a_low)
    a_up = a_low.repeat_interleave(2, 1).repeat_interleave(2, 2).contiguous()
    pl_up = ops.split_f16(a_up)
    srcs_up, srcs_low = ([pl_up, pl_up], [pl_low, pl_low]) if two else ([pl_up], [pl_low])
    ref = ops.spade_modulate(srcs_up, ops.prep_conv_weight(wm), x, 1, one, zero, one, zero, passes=3)
    wc = ops.collapse_subpixel(wm)
    got = ops.spade_modulate(srcs_low, ops.prep_subpixel_weight(wc), x, 1, one, zero, one, zero, passes=3,
                             subpixel=True)
    rv = ref.hi.float() + ref.lo.float()
    e = ((got.hi.float() + got.lo.float()) - rv).abs().max().item() / rv.abs().max().item()
    print("sub-pixel vs upsampled K1 (3-pass) max-abs / max|ref| %.3e" % e)
    assert e < 2e-5
    # backward GEMMs vs autograd of conv(upsample(cat(sources)), W)
    dy = torch.randn(B, H, W, 2 * C, generator=g).cuda() * 1e-2
    gp, _ = ops.grad_prep(dy)
    src = torch.cat([a_low] * (2 if two else 1), 3).permute(0, 3, 1, 2).clone().requires_grad_(True)
    w_ref = wm.clone().requires_grad_(True)
    F.conv2d(F.interpolate(src, scale_factor=2, mode="nearest"), w_ref, padding=1).backward(dy.permute(0, 3, 1, 2))
    dwc = ops.subpixel_wgrad(gp, srcs_low, passes=3)
    wl = wm.clone().requires_grad_(True)          # chain rule through the collapse (what autograd does)
    ops.collapse_subpixel(wl).backward(dwc)
    err = (wl.grad - w_ref.grad).abs().max().item() / w_ref.grad.abs().max().item()
    print("sub-pixel weight gradient rel err %.3e" % err)
    assert err < 2e-4
    dsrc, amax = ops.subpixel_dgrad(gp, wc, passes=3)
    err = (dsrc.permute(0, 3, 1, 2) - src.grad).abs().max().item() / src.grad.abs().max().item()
    print("sub-pixel source gradient rel err %.3e" % err)
    assert err < 1e-4
    assert abs(float(amax) - float(dsrc.abs().max())) <= 1e-6 * float(amax)


def test_dgrad_leaky_relu_mask():
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    B, H, W, C = 1, 16, 16, 128
    t = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).cuda()
    dy = torch.randn(B, C, H, W, generator=g).cuda()
    act = F.leaky_relu(t, 0.2)
    F.conv2d(act, w, None, padding=1).backward(dy)
    a = ops.split_f16(_nhwc(act.detach()))
    gp, _ = ops.grad_prep(_nhwc(dy))
    pwT = ops.prep_conv_weight(w, transpose=True)
    dt = _nchw(ops.conv3x3([gp], pwT, None, passes=3, act_mask=a.hi))
    assert (dt - t.grad).abs().max().item() <= 3e-5 * t.grad.abs().max().item()


@pytest.mark.parametrize("passes", [1, 3])
@pytest.mark.parametrize("ups,noisy", [(0, False), (1, True)])
def test_fused_dgrad_modulate_bwd_matches_two_kernel_path(ups, noisy, passes):
    """dsee_dgrad_modulate_bwd (backward-data GEMM with K1's backward as its epilogue, plane scale from
    the a-priori bound max|dY| * row-L1(W)) against dgrad -> dt in HBM -> dsee_spade_modulate_bwd_saved
    (plane scale from the measured max|dt|): dxhat identical, [dG|dB] planes equal after un-scaling to
    fp16 rounding, per-channel sums equal to summation order."""
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(21 + ups)
    B, H, W, C = 2, 24, 32, 256
    want_lo = passes == 3
    Hx, Wx = H >> ups, W >> ups
    x = torch.randn(B, Hx, Wx, C, generator=g).cuda()
    act = ops.split_f16(torch.randn(B, H, W, C, generator=g).cuda(), want_lo)
    G = ops.split_f16((torch.randn(B, H, W, C, generator=g) * 0.5 + 1).cuda(), want_lo)
    w = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).cuda()
    dy = (torch.randn(B, H, W, C, generator=g) * 1e-3).cuda()
    sc = (torch.rand(C, generator=g) + 0.5).cuda()
    sh = (torch.randn(C, generator=g) * 0.1).cuda()
    noise = torch.randn(B, H, W, C, generator=g).cuda() if noisy else None
    nw = (torch.randn(C, generator=g) * 0.3).cuda() if noisy else None
    gp, _ = ops.grad_prep(dy, want_lo=want_lo)
    pwT = ops.prep_conv_weight(w, want_lo, transpose=True)
    # the a-priori bound really bounds dt
    dt, amax = ops.conv3x3([gp], pwT, None, passes=passes, act_mask=act.hi, want_amax=True)
    bound = float(gp.inv_scale[1]) * float(pwT.inv_scale[2])
    l1 = float(w.abs().sum(dim=(0, 2, 3)).max())
    assert abs(float(pwT.inv_scale[2]) - l1) <= 5e-3 * l1
    assert float(amax) <= bound
    dxh0, dgb0, s0 = ops.spade_modulate_bwd_saved(G, x, ups, sc, sh, dt, amax, noise=noise, noise_w=nw,
                                                  want_lo=want_lo)
    dxh1, dgb1, s1 = ops.dgrad_modulate_bwd(gp, pwT, act.hi, G, x, ups, sc, sh, noise=noise, noise_w=nw,
                                            passes=passes, want_lo=want_lo)
    assert torch.equal(dxh0, dxh1)
    val = lambda p: (p.hi.float() + (p.lo.float() if p.lo is not None else 0)) * p.inv_scale
    v0, v1 = val(dgb0), val(dgb1)
    tol = (2e-6 if want_lo else 1.2e-3) * v0.abs().max().item()
    assert (v0 - v1).abs().max().item() <= tol
    torch.testing.assert_close(s1, s0, rtol=1e-4, atol=1e-5 * s0.abs().max().item())


@pytest.mark.parametrize("ups,noisy", [(0, False), (1, True), (1, False)])
def test_bn_bwd_matches_autograd(ups, noisy):
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(11 + ups)
    B, Hx, Wx, C = 2, 8, 12, 128
    H, W = Hx << ups, Wx << ups
    x = torch.randn(B, C, Hx, Wx, generator=g).cuda().requires_grad_(True)
    noise = torch.randn(B, C, H, W, generator=g).cuda()
    nw = (torch.randn(C, generator=g) * 0.3).cuda().requires_grad_(True)
    dxhat = torch.randn(B, C, H, W, generator=g).cuda()
    dskip = torch.randn(B, C, H, W, generator=g).cuda()
    xu = F.interpolate(x, scale_factor=2, mode="nearest") if ups else x
    xin = xu + nw.view(1, -1, 1, 1) * noise if noisy else xu
    xhat = F.batch_norm(xin, None, None, None, None, True, 0.1, 1e-5)
    ((xhat * dxhat).sum() + (xu * dskip).sum()).backward()
    n_ = _nhwc(noise) if noisy else None
    part = ops.bn_stats(_nhwc(x.detach()), ups, n_, nw.detach() if noisy else None)
    sc, sh, _, _ = ops.bn_finalize(part, B * H * W, 1e-5)
    xh = _nhwc(xhat.detach())
    dxh = _nhwc(dxhat)
    sums = torch.stack([dxh.sum(dim=(0, 1, 2)), (dxh * xh).sum(dim=(0, 1, 2))]).contiguous()
    dx, dnw = ops.bn_bwd(dxh, _nhwc(x.detach()), ups, sc, sh, sums, 1.0 / (B * H * W), noise=n_,
                         noise_w=nw.detach() if noisy else None, dskip=_nhwc(dskip))
    torch.testing.assert_close(_nchw(dx), x.grad, rtol=1e-3, atol=2e-4)
    if noisy:
        torch.testing.assert_close(dnw, nw.grad, rtol=1e-3, atol=2e-3)
    # optional outputs: max|dx| for the gradient split that follows, and the shortcut's share of
    # d noise_w (sum dskip * noise) folded into the same reduction
    dx2, dnw2, amax = ops.bn_bwd(dxh, _nhwc(x.detach()), ups, sc, sh, sums, 1.0 / (B * H * W), noise=n_,
                                 noise_w=nw.detach() if noisy else None, dskip=_nhwc(dskip),
                                 noise_grad_with_skip=noisy, want_amax=True)
    assert torch.equal(dx2, dx) and float(amax) == float(dx.abs().max())
    if noisy:
        extra = (dskip * noise).sum(dim=(0, 2, 3))
        torch.testing.assert_close(dnw2, nw.grad + extra, rtol=1e-3, atol=4e-3)
    gp_a, s_a = ops.grad_prep(dx)
    gp_b, s_b = ops.grad_prep(dx, amax=amax)
    assert torch.equal(gp_a.hi, gp_b.hi) and torch.equal(gp_a.lo, gp_b.lo)
    assert torch.equal(gp_a.inv_scale, gp_b.inv_scale) and torch.equal(s_a, s_b)


def test_stem_and_head_backward():
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(3)
    B, H, W, C = 2, 12, 20, 128
    x = torch.randn(B, 3, H, W, generator=g).cuda()
    w = (torch.randn(C, 3, 3, 3, generator=g) * 0.2).cuda().requires_grad_(True)
    b = torch.randn(C, generator=g).cuda().requires_grad_(True)
    dy = torch.randn(B, C, H, W, generator=g).cuda()
    F.conv2d(x, w, b, padding=1).backward(dy)
    dw, db = ops.stem_bwd(x, _nhwc(dy))
    torch.testing.assert_close(dw, w.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db, b.grad, rtol=1e-4, atol=1e-3)

    xh = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    wh = (torch.randn(3, C, 3, 3, generator=g) / (3 * C ** 0.5)).cuda().requires_grad_(True)
    bh = torch.randn(3, generator=g).cuda().requires_grad_(True)
    dout = torch.randn(B, 3, H, W, generator=g).cuda()
    out = torch.tanh(F.conv2d(F.leaky_relu(xh, 0.2), wh, bh, padding=1))
    out.backward(dout)
    dx, dwh, dbh = ops.head_bwd(_nhwc(xh.detach()), wh.detach(), out.detach().contiguous(), dout)
    torch.testing.assert_close(_nchw(dx), xh.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dwh, wh.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dbh, bh.grad, rtol=1e-4, atol=1e-3)


def test_shared_mlp_and_style_gather_backward():
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(9)
    B, Hl, Wl, L, nh, d = 2, 10, 14, 19, 128, 128
    for ups in (0, 1):
        H, W = Hl << ups, Wl << ups
        labels = torch.randint(0, L, (B, Hl, Wl), generator=g, dtype=torch.uint8).cuda()
        wt = (torch.randn(nh, L, 3, 3, generator=g) * 0.5).cuda().requires_grad_(True)
        bias = torch.randn(nh, generator=g).cuda().requires_grad_(True)
        style = torch.randn(B, L, d, generator=g).cuda().requires_grad_(True)
        onehot = F.one_hot(labels.long(), L).permute(0, 3, 1, 2).float()
        actv = F.relu(F.conv2d(onehot, wt, bias, padding=1))
        smap = torch.einsum("bld,blhw->bdhw", style, onehot)
        if ups:
            actv = F.interpolate(actv, scale_factor=2, mode="nearest")
            smap = F.interpolate(smap, scale_factor=2, mode="nearest")
        dsrc = torch.randn(B, H, W, nh + d, generator=g).cuda()
        (actv * dsrc[..., :nh].permute(0, 3, 1, 2)).sum().backward(retain_graph=True)
        (smap * dsrc[..., nh:].permute(0, 3, 1, 2)).sum().backward()
        table = wt.detach().permute(2, 3, 1, 0).reshape(9, L, nh).contiguous()
        planes = ops.shared_mlp(labels, table, bias.detach(), ups=ups)
        dtab, dtb = ops.shared_mlp_bwd(dsrc, 0, planes.hi, labels, ups, L)
        ref_tab = wt.grad.permute(2, 3, 1, 0).reshape(9, L, nh)
        torch.testing.assert_close(dtab, ref_tab, rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(dtb, bias.grad, rtol=1e-4, atol=1e-3)
        # tensor-core route: wgrad of G against the one-hot plane
        amax = dsrc.abs().max().reshape(1)
        dtab2, dtb2 = ops.shared_mlp_bwd_tc(dsrc, amax, 0, planes.hi, labels, ops.onehot_planes(labels),
                                            ups, L, passes=3)
        assert (dtab2 - ref_tab).abs().max().item() <= 5e-5 * ref_tab.abs().max().item()
        torch.testing.assert_close(dtb2, bias.grad, rtol=1e-4, atol=1e-3)
        if not ups:
            ds = ops.style_gather_bwd(dsrc, nh, labels, L, d)
            torch.testing.assert_close(ds, style.grad, rtol=1e-4, atol=1e-3)


# ---------------------------------------------------------------------------------------------
# style encoder / discriminator layer nodes: forward + backward against torch autograd
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Cin,Cout,K,stride,pad,ups,lrelu,bias", [
    (4, 32, 3, 1, 1, 0, False, False),     # encoder initial (RGB padded to 4)
    (32, 64, 3, 2, 1, 0, False, False),    # encoder down
    (64, 128, 3, 1, 1, 1, False, False),   # encoder up_conv (folded upsample)
    (24, 32, 4, 2, 2, 0, True, True),      # discriminator model0 (+ fused LeakyReLU)
    (32, 64, 4, 2, 2, 0, False, False),    # discriminator inner
    (64, 128, 4, 1, 2, 0, False, False),
    (128, 1, 4, 1, 2, 0, False, True),     # discriminator prediction conv
])
def test_conv2d_direct_node(Cin, Cout, K, stride, pad, ups, lrelu, bias):
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(Cin * 3 + Cout + K)
    B, H, W = 2, 13, 18
    x = torch.randn(B, Cin, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(Cout, Cin, K, K, generator=g) / (K * Cin ** 0.5)).cuda().requires_grad_(True)
    b = torch.randn(Cout, generator=g).cuda().requires_grad_(True) if bias else None
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if ups else x
    ref = F.conv2d(xin, w, b, stride=stride, padding=pad)
    if lrelu:
        ref = F.leaky_relu(ref, 0.2)
    dy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1)).cuda()
    ref.backward(dy)
    x2 = _nhwc(x.detach()).requires_grad_(True)
    w2 = w.detach().permute(2, 3, 1, 0).contiguous().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True) if bias else None
    out = ops.Conv2dDirectFn.apply(x2, w2, b2, stride, pad, ups, lrelu)
    torch.testing.assert_close(_nchw(out), ref, rtol=1e-4, atol=5e-5)
    out.backward(_nhwc(dy))
    torch.testing.assert_close(_nchw(x2.grad), x.grad, rtol=1e-4, atol=5e-5)
    wref = w.grad.permute(2, 3, 1, 0)
    assert (w2.grad - wref).abs().max().item() <= 1e-4 * wref.abs().max().item()
    if bias:
        assert (b2.grad - b.grad).abs().max().item() <= 1e-4 * b.grad.abs().max().item()


@pytest.mark.parametrize("act", [0, 1, 2])
def test_instance_norm_node(act):
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(40 + act)
    x = (torch.randn(2, 64, 9, 14, generator=g) * 1.7 + 0.3).cuda().requires_grad_(True)
    y = F.instance_norm(x, eps=1e-5)
    ref = [y, F.leaky_relu(y, 0.2), torch.tanh(y)][act]
    dy = torch.randn(ref.shape, generator=g).cuda()
    ref.backward(dy)
    x2 = _nhwc(x.detach()).requires_grad_(True)
    out = ops.InstanceNormFn.apply(x2, act)
    torch.testing.assert_close(_nchw(out), ref, rtol=1e-4, atol=1e-5)
    out.backward(_nhwc(dy))
    torch.testing.assert_close(_nchw(x2.grad), x.grad, rtol=2e-4, atol=2e-5)


def test_pool_and_input_nodes():
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(77)
    B, H, W, C, L = 2, 11, 16, 128, 19
    # region pool (encoder.py:36-49)
    x = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    labels = torch.randint(0, L, (B, H, W), generator=g, dtype=torch.uint8).cuda()
    onehot = F.one_hot(labels.long(), L).permute(0, 3, 1, 2).float()
    ref = torch.einsum("bchw,blhw->blc", x, onehot) / (H * W)
    ds = torch.randn(B, L, C, generator=g).cuda()
    ref.backward(ds)
    x2 = _nhwc(x.detach()).requires_grad_(True)
    out = ops.RegionPoolFn.apply(x2, labels, L)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-6)
    out.backward(ds)
    torch.testing.assert_close(_nchw(x2.grad), x.grad, rtol=1e-5, atol=1e-7)
    # avg-pool pyramid (discriminator.py:46-49)
    for (h, w) in ((11, 16), (8, 8), (9, 5)):
        x = torch.randn(B, 8, h, w, generator=g).cuda().requires_grad_(True)
        ref = F.avg_pool2d(x, 3, stride=2, padding=[1, 1], count_include_pad=False)
        dy = torch.randn(ref.shape, generator=g).cuda()
        ref.backward(dy)
        x2 = _nhwc(x.detach()).requires_grad_(True)
        out = ops.AvgPool3s2Fn.apply(x2)
        torch.testing.assert_close(_nchw(out), ref, rtol=1e-5, atol=1e-6)
        out.backward(_nhwc(dy))
        torch.testing.assert_close(_nchw(x2.grad), x.grad, rtol=1e-5, atol=1e-6)
    # discriminator input assembly (sr_model.py:655-664)
    fake = torch.randn(B, 3, H, W, generator=g).cuda().requires_grad_(True)
    real = torch.randn(B, 3, H, W, generator=g).cuda()
    xin = ops.DiscInputFn.apply(labels, fake, real, L, 24)
    dxin = torch.randn(xin.shape, generator=g).cuda()
    xin.backward(dxin)
    torch.testing.assert_close(fake.grad, dxin[:B, :, :, L:L + 3].permute(0, 3, 1, 2))


@pytest.mark.parametrize("Cx,Cw,Cout,K,stride,pad,ups,lrelu,bias", [
    (32, 32, 64, 3, 2, 1, 0, False, False),     # encoder down0
    (64, 64, 128, 3, 2, 1, 0, False, False),    # encoder down1
    (128, 128, 256, 3, 1, 1, 1, False, False),  # encoder up_conv (upsample materialised)
    (256, 256, 128, 3, 1, 1, 0, False, False),  # encoder final
    (32, 22, 32, 4, 2, 2, 0, True, True),       # discriminator model0 (22 real + 10 zero channels)
    (24, 22, 32, 4, 2, 2, 0, True, True),
    (32, 32, 64, 4, 2, 2, 0, False, False),     # discriminator inner, stride 2
    (128, 128, 256, 4, 1, 2, 0, False, False),  # discriminator inner, stride 1
])
@pytest.mark.parametrize("passes", [3, 1])
def test_conv2d_tc_node(Cx, Cw, Cout, K, stride, pad, ups, lrelu, bias, passes):
    """The general strided tcgen05 conv (TMA element strides), its 4-parity-class backward-data
    and the strided weight gradient against torch autograd."""
    from deepsee_b200 import ops
    from deepsee_b200.config import config
    g = torch.Generator(device="cpu").manual_seed(Cx + Cout + K + stride)
    B, H, W = 2, 13, 18
    x = torch.randn(B, Cx, H, W, generator=g)
    x[:, Cw:] = 0
    x = x.cuda().requires_grad_(True)
    w = (torch.randn(Cout, Cw, K, K, generator=g) / (K * Cw ** 0.5)).cuda().requires_grad_(True)
    b = torch.randn(Cout, generator=g).cuda().requires_grad_(True) if bias else None
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if ups else x
    ref = F.conv2d(xin[:, :Cw], w, b, stride=stride, padding=pad)
    if lrelu:
        ref = F.leaky_relu(ref, 0.2)
    dy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1)).cuda()
    ref.backward(dy)
    x2 = _nhwc(x.detach()).requires_grad_(True)
    w2 = w.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True) if bias else None
    assert ops.tc_conv_eligible(Cx, Cout)
    old = config.passes
    config.passes = passes
    try:
        out = ops.conv_layer(x2, w2, b2, stride, pad, ups=ups, lrelu=lrelu)
        out.backward(_nhwc(dy))
    finally:
        config.passes = old
    tol = 5e-5 if passes == 3 else 2e-2

    def chk(a, r, what):
        if passes == 3:
            err, s = (a - r).abs().max().item(), r.abs().max().item()
            assert err <= tol * s, (what, err, s)
        else:
            # 1-pass forward differences (1e-3) flip a few fused-LeakyReLU masks; relative L2 is the
            # meaningful measure there
            rel = ((a - r).norm() / r.norm()).item()
            assert rel <= (2e-2 if lrelu else 5e-3), (what, rel)
    chk(_nchw(out), ref, "fwd")
    chk(_nchw(x2.grad)[:, :Cw], x.grad[:, :Cw], "dgrad")
    chk(w2.grad, w.grad, "wgrad")
    if bias:
        chk(b2.grad, b.grad, "dbias")


# ---------------------------------------------------------------------------------------------
# parameter-side fusions
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(128, 128, 3, 3), (512, 512, 3, 3), (64, 32, 4, 4), (128, 256, 3, 3)])
@pytest.mark.parametrize("training", [True, False])
def test_spectral_weight_node_vs_torch(shape, training):
    """ops.SpectralWeightFn against torch.nn.utils.spectral_norm's own hook: effective weight, the
    in-place u / v power-iteration update and the gradient wrt weight_orig."""
    import torch.nn as nn
    from deepsee_b200 import ops
    torch.manual_seed(shape[0] + shape[1])
    conv = nn.utils.spectral_norm(nn.Conv2d(shape[1], shape[0], shape[2], padding=1)).cuda()
    conv.train(training)
    w0, u0, v0 = conv.weight_orig.detach().clone(), conv.weight_u.clone(), conv.weight_v.clone()
    for hook in conv._forward_pre_hooks.values():
        hook(conv, None)
    ref = conv.weight
    dy = torch.randn_like(ref)
    ref.backward(dy)
    w1 = w0.clone().requires_grad_(True)
    u1, v1 = u0.clone(), v0.clone()
    out = ops.SpectralWeightFn.apply(w1, u1, v1, training, 1e-12)
    out.backward(dy)
    torch.testing.assert_close(out, ref, rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(u1, conv.weight_u, rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(v1, conv.weight_v, rtol=2e-5, atol=1e-6)
    g = conv.weight_orig.grad
    assert (w1.grad - g).abs().max().item() <= 2e-5 * g.abs().max().item()


@pytest.mark.parametrize("training", [True, False])
def test_batched_spectral_prepass_equals_per_layer_node(training):
    """spectral_prepass (one batched launch sequence for a list of layers of different shapes) against
    the per-layer SpectralWeightFn: effective weights, updated u / v and weight_orig gradients are
    bit-identical; an unlisted layer still takes the per-layer path."""
    import copy
    import torch.nn as nn
    from deepsee_b200.config import config
    from deepsee_b200.deepsee_models.networks.normalization import effective_weight, spectral_prepass
    torch.manual_seed(3)
    shapes = [(64, 27, 3), (128, 64, 3), (32, 48, 4), (512, 512, 3)]
    a = [nn.utils.spectral_norm(nn.Conv2d(ci, co, k)).cuda().train(training) for co, ci, k in shapes]
    b = copy.deepcopy(a)
    extra_a = nn.utils.spectral_norm(nn.Conv2d(8, 16, 3)).cuda().train(training)
    extra_b = copy.deepcopy(extra_a)
    dys = [torch.randn_like(c.weight_orig) for c in a + [extra_a]]
    old = config.batched_spectral
    try:
        config.batched_spectral = False
        wa = [effective_weight(c) for c in a + [extra_a]]
        config.batched_spectral = True
        with spectral_prepass(b):
            assert all(c._sn_pre is not None for c in b)
            wb = [effective_weight(c) for c in b + [extra_b]]
        assert all(c._sn_pre is None for c in b)
    finally:
        config.batched_spectral = old
    for x, y, dy in zip(wa, wb, dys):
        assert torch.equal(x, y)
        x.backward(dy)
        y.backward(dy)
    for ca, cb in zip(a + [extra_a], b + [extra_b]):
        assert torch.equal(ca.weight_u, cb.weight_u) and torch.equal(ca.weight_v, cb.weight_v)
        assert torch.equal(ca.weight_orig.grad, cb.weight_orig.grad)


@pytest.mark.parametrize("kind", ["spade", "sean", "puresean"])
def test_modweight_node_vs_torch_ops(kind):
    from deepsee_b200.deepsee_models.networks import normalization as Nz
    from deepsee_b200.options.configurations import make_opt
    opt = make_opt(None)
    cls = {"spade": Nz.SPADE, "sean": Nz.SEAN_Block, "puresean": Nz.PureSEAN_Block}[kind]
    torch.manual_seed(3)
    m = cls("spadesyncbatch3x3" if kind == "spade" else "seansyncbatch3x3", 256, 19, opt).cuda()
    for p in m.parameters():
        torch.nn.init.normal_(p, 0.0, 0.3)
    w_ref, gb_ref, bb_ref = m.combined_weight_torch()
    dw, dg, db = torch.randn_like(w_ref), torch.randn_like(gb_ref), torch.randn_like(bb_ref)
    ((w_ref * dw).sum() + (gb_ref * dg).sum() + (bb_ref * db).sum()).backward()
    ref_grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    w, gb, bb = m.combined_weight()
    torch.testing.assert_close(w, w_ref, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(gb, gb_ref, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(bb, bb_ref, rtol=1e-6, atol=1e-7)
    ((w * dw).sum() + (gb * dg).sum() + (bb * db).sum()).backward()
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(ref_grads), (sorted(got), sorted(ref_grads))
    for k, r in ref_grads.items():
        assert (got[k] - r).abs().max().item() <= 1e-4 * max(r.abs().max().item(), 1e-6), k


@pytest.mark.parametrize("B,H,W,Cin,Cout,passes", [(1, 16, 16, 128, 256, 1), (2, 32, 48, 256, 512, 3),
                                                   (3, 48, 40, 512, 512, 2), (1, 64, 64, 256, 256, 2)])
def test_conv3x3_cta_pair_is_bit_identical(B, H, W, Cin, Cout, passes):
    """The cta_group::2 form (two SMs of a TPC share one weight box) accumulates every output element in
    the same order as the single-CTA kernel: outputs and batch-norm partial sums must agree exactly."""
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(H + Cin + Cout + passes)
    x = torch.randn(B, Cin, H, W, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    xn = _nhwc(x)
    a = ops.split_f16(xn)
    if passes == 2:
        lo = xn - a.hi.float()
        a = ops.SplitPlanes(a.hi, None, ((lo * 256).to(torch.float8_e5m2).view(torch.uint8),
                                         xn.to(torch.float8_e5m2).view(torch.uint8)))
    pw = ops.prep_conv_weight(w, want_lo=passes == 3, want_f8=passes == 2)
    res = torch.randn(B, H, W, Cout, generator=g).cuda()
    prev = ops.conv_pair_mode(False)
    try:
        o0, s0 = ops.conv3x3([a], pw, bias, passes=passes, residual=res, want_stats=True)
        ops.conv_pair_mode(True)
        o1, s1 = ops.conv3x3([a], pw, bias, passes=passes, residual=res, want_stats=True)
    finally:
        ops.conv_pair_mode(prev)
    ref = F.conv2d(x, w, bias, padding=1) + _nchw(res)
    assert (_nchw(o1) - ref).abs().max().item() <= 4e-3 * ref.abs().max().item()
    assert torch.equal(o0, o1)
    # same per-tile partial sums in a different slot order (16-row tiles, two halves each)
    torch.testing.assert_close(s0.double().sum(0), s1.double().sum(0), rtol=1e-6, atol=1e-6)
    assert torch.equal(s0.sort(0).values, s1.sort(0).values)


@pytest.mark.parametrize("B,H,W,C", [(2, 12, 20, 128), (1, 40, 24, 512)])
def test_tensor_core_head_forward_and_backward(B, H, W, C):
    """tanh(conv_img(leaky_relu(x))) as a 1x1 tcgen05 GEMM over the fp16 planes the last main conv writes
    (dsee_conv_epilogue.act16_*) + the 9-tap shift-add, against torch (sr.py:94-95)."""
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(C + H)
    # the planes come out of a real conv3x3 launch: x = conv(a) + bias, planes = split(leaky_relu(x))
    a = ops.split_f16(torch.randn(B, H, W, C, generator=g).cuda())
    wk = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).cuda()
    x_nhwc = ops.conv3x3([a], ops.prep_conv_weight(wk), torch.randn(C, generator=g).cuda(), passes=3, act16=True)
    planes = x_nhwc._dsee_act16
    act = F.leaky_relu(x_nhwc, 0.2)
    assert torch.equal(planes.hi, act.half())
    torch.testing.assert_close(planes.hi.float() + planes.lo.float(), act, rtol=0, atol=2e-7 * act.abs().max().item())

    xh = _nchw(x_nhwc).detach().clone().requires_grad_(True)
    wh = (torch.randn(3, C, 3, 3, generator=g) / (3 * C ** 0.5)).cuda().requires_grad_(True)
    bh = torch.randn(3, generator=g).cuda().requires_grad_(True)
    dout = torch.randn(B, 3, H, W, generator=g).cuda() * 1e-4
    ref = torch.tanh(F.conv2d(F.leaky_relu(xh, 0.2), wh, bh, padding=1))
    ref.backward(dout)
    out = ops.head_tc(planes, wh.detach(), bh.detach(), passes=3)
    # fp32-class: both operand planes, fp32 accumulation in the tensor core (which truncates when it
    # aligns addends: ~1e-5 at |pre-activation| ~ 1; the CUDA-core head it replaces reaches 2e-6)
    assert (out - ref).abs().max().item() <= 3e-5
    out1 = ops.head_tc(ops.SplitPlanes(planes.hi, None), wh.detach(), bh.detach(), passes=1)
    assert (out1 - ref).abs().max().item() <= 2e-3
    for passes, tol in ((3, 1e-4), (1, 4e-3)):
        dx, amax, dw, db = ops.head_tc_bwd(planes, wh.detach(), out, dout, passes=passes)
        s = xh.grad.abs().max().item()
        assert (_nchw(dx) - xh.grad).abs().max().item() <= tol * s, passes
        assert abs(amax.item() - dx.abs().max().item()) <= 1e-6 * s
        assert (dw - wh.grad).abs().max().item() <= tol * wh.grad.abs().max().item(), passes
        torch.testing.assert_close(db, bh.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("ups", [0, 1])
def test_shared_mlp_uniform_window_fast_path_is_bit_identical(ups):
    """Piecewise-constant label maps (a face parse): pixels whose 3x3 window carries one label take the
    precomputed-row path of dsee_shared_mlp_fwd; planes must equal the nine-tap gather bit for bit."""
    from deepsee_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(11)
    B, L, nh, S = 2, 19, 128, 48
    grid = torch.randint(0, L, (B, 1, 6, 6), generator=g)
    lab = F.interpolate(grid.float(), size=(S, S), mode="nearest").long().cuda()   # 8 x 8 blocks
    lab[0, 0, 5, 7] = (lab[0, 0, 5, 7] + 1) % L                                    # and an isolated pixel
    w = torch.randn(nh, L, 3, 3, generator=g).cuda() * 0.2
    b = torch.randn(nh, generator=g).cuda() * 0.1
    table = w.permute(2, 3, 1, 0).reshape(9, L, nh).contiguous()
    labels = lab[:, 0].to(torch.uint8).contiguous()
    fast = ops.shared_mlp(labels, table, b, ups=ups, uniform_rows=True)
    slow = ops.shared_mlp(labels, table, b, ups=ups, uniform_rows=False)
    assert torch.equal(fast.hi, slow.hi) and torch.equal(fast.lo, slow.lo)
    oh = torch.zeros(B, L, S, S, device="cuda").scatter_(1, lab, 1.0)
    ref = F.relu(F.conv2d(oh, w, b, padding=1))
    if ups:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
    got = _nchw(fast.hi.float() + fast.lo.float())
    assert (got - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())
