#!/usr/bin/env python
"""Round-2 forms of the tensor-core kernels, each launched ONCE at the shapes of the top generator
block (c2: 256x256, batch 8; c4/c5: --S 512 --B 2), for `ncu --set full` and CUDA-event timing:

  K1 folded       gamma/beta GEMM over [actv | one-hot] with per-image weights (K = 9 x 192)
  K2 fp8corr      main conv, one fp16 pass + the fp8 correction GEMM (passes = 2), with the in-step
                  epilogue of conv_1: low-resolution residual through the folded upsample, two in-kernel
                  noise terms, BN partial sums
  K2 1-pass       the same launch with passes = 1 (what the backward-data GEMM costs)
  dgrad+K1bwd     backward-data of a main conv fused with K1's backward
  wgrad           main-conv weight gradient;  wgrad per image: the folded modulation weight gradient
  dgrad_mod       backward-data of the modulation GEMM, N = 128 (only the mlp_shared activation)
  head            the image head on the tensor cores: K2 writing leaky_relu(x) as fp16 planes, the 1x1 GEMM
                  [pixels x 512] x [512 x 27], the 9-tap shift-add + tanh, and its backward (scatter, 1x1
                  wgrad, 1x1 dgrad with the LeakyReLU' mask)
  K1 sub-pixel    (--sub) the layer above max_fm_size: four 2x2-tap GEMMs over the half-resolution actv,
                  with its backward pair (subpixel_wgrad, subpixel_dgrad)

    python profiles/kernels_microbench_r2.py [--B 8 --S 256] [--sub] [--reps N] [--no-warm]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsee_b200 import ops  # noqa: E402
from deepsee_b200.deepsee_models.networks.normalization import fold_style_weight  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--S", type=int, default=256)
    ap.add_argument("--sub", action="store_true")
    ap.add_argument("--no-warm", action="store_true")
    ap.add_argument("--pair", action="store_true", help="also time K2 in the CTA-pair (cta_group::2) form")
    a = ap.parse_args()
    B, S, C, L, nh, d = a.B, a.S, 512, 19, 128, 128
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev)  # noqa: E731
    px = B * S * S
    labels = torch.randint(0, L, (B, S, S), generator=g, device=dev, dtype=torch.uint8)
    table, tb = rn(9, L, nh), rn(nh)
    x_lo = rn(B, S // 2, S // 2, C)
    sc, sh, gb, bb = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev)
    nw = torch.full((C,), 0.1, device=dev)
    w = rn(C, C, 3, 3) / (3 * C ** 0.5)
    bias = rn(C)
    dy = rn(B, S, S, C) * 1e-3

    def timed(name, fn, flops):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not a.no_warm:
            r = fn()
            r = fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print("%-30s %8.3f ms  %7.1f TFLOP/s (reference-equivalent)" % (name, ms, flops / ms / 1e9))
        return r

    torch.cuda.profiler.start()
    if not a.sub:
        actv = ops.shared_mlp(labels, table, tb, want_lo=False)
        oh = ops.onehot_planes(labels)
        wm = rn(2 * C, nh + d, 3, 3) / (3 * (nh + d) ** 0.5)
        style = torch.rand(B, L, d, generator=g, device=dev) * 2 - 1
        Wa, Ws = fold_style_weight(wm, style)
        pwm = ops.prep_mod_weight_batched(Wa, Ws, want_lo=False)
        act, gsaved = timed("K1 folded (per-image weights)", lambda: ops.spade_modulate(
            [actv, oh], pwm, x_lo, 1, sc, sh, gb, bb, noise=ops.NoiseSeed(11), noise_w=nw, passes=1, want_lo=False,
            save_g=True, want_f8=True), 2 * 9 * (nh + d) * 2 * C * px)
    else:
        lab_lo = labels[:, ::2, ::2].contiguous()
        actv = ops.shared_mlp(lab_lo, table, tb, want_lo=False)
        wm = rn(2 * C, nh, 3, 3) / (3 * nh ** 0.5)
        wc = ops.collapse_subpixel(wm)
        pws = ops.prep_subpixel_weight(wc, want_lo=False)
        act, gsaved = timed("K1 sub-pixel (4 x 2x2 taps)", lambda: ops.spade_modulate(
            [actv], pws, x_lo, 1, sc, sh, gb, bb, passes=1, want_lo=False, save_g=True, want_f8=True, subpixel=True),
            2 * 9 * nh * 2 * C * px)
    pw2 = ops.prep_conv_weight(w, want_lo=False, want_f8=True)
    timed("K2 fp16 + fp8 correction", lambda: ops.conv3x3(
        [act], pw2, bias, residual=x_lo, res_ups=1,
        noises=[(ops.NoiseSeed(5), nw), (ops.NoiseSeed(6), nw)], passes=2, want_stats=True), 2 * 9 * C * C * px)
    act1 = ops.SplitPlanes(act.hi, None)
    timed("K2 1-pass (same epilogue)", lambda: ops.conv3x3(
        [act1], pw2, bias, residual=x_lo, res_ups=1,
        noises=[(ops.NoiseSeed(5), nw), (ops.NoiseSeed(6), nw)], passes=1, want_stats=True), 2 * 9 * C * C * px)
    if a.pair:
        ops.conv_pair_mode(True)
        timed("K2 fp16 + fp8 corr., CTA pair", lambda: ops.conv3x3(
            [act], pw2, bias, residual=x_lo, res_ups=1,
            noises=[(ops.NoiseSeed(5), nw), (ops.NoiseSeed(6), nw)], passes=2, want_stats=True), 2 * 9 * C * C * px)
        timed("K2 1-pass, CTA pair", lambda: ops.conv3x3(
            [act1], pw2, bias, residual=x_lo, res_ups=1,
            noises=[(ops.NoiseSeed(5), nw), (ops.NoiseSeed(6), nw)], passes=1, want_stats=True), 2 * 9 * C * C * px)
        ops.conv_pair_mode(False)
    gp, _ = ops.grad_prep(dy, want_lo=False)
    pwT = ops.prep_conv_weight(w, want_lo=False, transpose=True)
    dxhat, dgb, sums = timed("dgrad + K1 backward (fused)", lambda: ops.dgrad_modulate_bwd(
        gp, pwT, act.hi, gsaved, x_lo, 1, sc, sh, passes=1, want_lo=False), 2 * 9 * C * C * px)
    timed("wgrad 512x512", lambda: ops.conv3x3_wgrad(gp, act1, passes=1), 2 * 9 * C * C * px)
    if not a.sub:
        timed("wgrad per image (folded mod.)", lambda: ops.conv3x3_wgrad_per_image(dgb, [actv, oh], passes=1),
              2 * 9 * (nh + 64) * 2 * C * px)
        pwaT = ops.prep_conv_weight(Wa, want_lo=False, transpose=True)
        timed("dgrad_mod (N = 128)", lambda: ops.conv3x3([dgb], pwaT, None, passes=1, want_amax=True, tag="dgrad_mod"),
              2 * 9 * nh * 2 * C * px)
    else:
        timed("sub-pixel wgrad (4 classes)", lambda: ops.subpixel_wgrad(dgb, [actv], passes=1), 2 * 9 * nh * 2 * C * px)
        timed("sub-pixel dgrad (16 taps)", lambda: ops.subpixel_dgrad(dgb, wc, passes=1, want_lo=False),
              2 * 9 * nh * 2 * C * px)
    # ---- image head (sr.py:94-95) on the tensor cores ----------------------------------------------
    xt = timed("K2 fp8corr + act16 planes (head producer)", lambda: ops.conv3x3(
        [act], pw2, bias, residual=x_lo, res_ups=1,
        noises=[(ops.NoiseSeed(5), nw), (ops.NoiseSeed(6), nw)], passes=2, want_stats=True, act16=True),
        2 * 9 * C * C * px)[0]
    planes = xt._dsee_act16
    wh = rn(3, C, 3, 3) / (3 * C ** 0.5)
    bh = torch.zeros(3, device=dev)
    out = timed("head forward (1x1 GEMM 3-pass + gather)", lambda: ops.head_tc(planes, wh, bh, passes=3),
                2 * 27 * C * px)
    dout = rn(B, 3, S, S) * 1e-4
    timed("head backward (scatter, wgrad, dgrad)", lambda: ops.head_tc_bwd(
        ops.SplitPlanes(planes.hi, None), wh, out, dout, passes=1), 2 * 2 * 27 * C * px)
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
