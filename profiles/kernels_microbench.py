#!/usr/bin/env python
"""Launches each tensor-core kernel of the generator path ONCE at the shapes of bench.py's default
workload (c2: 256x256, batch 8, 512 channels; the top block of the generator), for
`ncu --set full` captures and for CUDA-event timing without the rest of the step around them.

    python profiles/kernels_microbench.py [--passes 1|3] [--reps N] [--B 8] [--S 256]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsee_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--S", type=int, default=256)
    ap.add_argument("--C", type=int, default=512)
    ap.add_argument("--no-warm", action="store_true",
                    help="launch every kernel exactly once (for `ncu --set full`, which replays each launch ~40 times)")
    a = ap.parse_args()
    B, S, C, L, nh, d = a.B, a.S, a.C, 19, 128, 128
    want_lo = a.passes == 3
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev)
    x = rn(B, S, S, C)
    act = ops.split_f16(rn(B, S, S, C), want_lo)
    w = rn(C, C, 3, 3) / (3 * C ** 0.5)
    pw = ops.prep_conv_weight(w, want_lo)
    pwT = ops.prep_conv_weight(w, want_lo, transpose=True)
    bias = rn(C)
    labels = torch.randint(0, L, (B, S, S), generator=g, device=dev, dtype=torch.uint8)
    table, tb = rn(9, L, nh), rn(nh)
    actv = ops.shared_mlp(labels, table, tb, want_lo=want_lo)
    smap = ops.style_gather(labels, rn(B, L, d), want_lo=want_lo)
    wm = rn(2 * C, nh + d, 3, 3) / (3 * (nh + d) ** 0.5)
    pwm = ops.prep_conv_weight(wm, want_lo)
    pwmT = ops.prep_conv_weight(wm, want_lo, transpose=True)
    wg = wm.view(C // 128, 2, 128, nh + d, 3, 3)[:, 0].reshape(C, nh + d, 3, 3).contiguous()
    pwg = ops.prep_conv_weight(wg, want_lo)
    sc, sh, gb, bb = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev)
    dy = rn(B, S, S, C) * 1e-3
    gp, _ = ops.grad_prep(dy, want_lo=want_lo)
    torch.cuda.synchronize()

    def timed(name, fn, flops):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # two warm-up calls with the first result still alive: the timed loop then recycles two
        # cached output sets and never reaches cudaMalloc
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
        print("%-28s %8.3f ms  %7.1f TFLOP/s (reference-equivalent), x%d executed" %
              (name, ms, flops / ms / 1e9, a.passes))
        return r

    px = B * S * S
    torch.cuda.profiler.start()
    timed("K2 conv3x3 fwd", lambda: ops.conv3x3([act], pw, bias, residual=x, passes=a.passes, want_stats=True),
          2 * 9 * C * C * px)
    _, gsaved = timed("K1 modulate fwd", lambda: ops.spade_modulate([actv, smap], pwm, x, 0, sc, sh, gb, bb,
                                                                    passes=a.passes, want_lo=want_lo, save_g=True),
                      2 * 9 * (nh + d) * 2 * C * px)
    dt, amax = timed("dgrad (K2 bwd-data)", lambda: ops.conv3x3([gp], pwT, None, passes=a.passes, act_mask=act.hi,
                                                                want_amax=True, tag="dgrad"), 2 * 9 * C * C * px)
    timed("wgrad 512x512", lambda: ops.conv3x3_wgrad(gp, act, passes=a.passes), 2 * 9 * C * C * px)
    dxhat, dgb, sums = timed("K1 backward (saved G)", lambda: ops.spade_modulate_bwd_saved(
        gsaved, x, 0, sc, sh, dt, amax, want_lo=want_lo), 0.0)
    timed("dgrad + K1 backward (fused)", lambda: ops.dgrad_modulate_bwd(
        gp, pwT, act.hi, gsaved, x, 0, sc, sh, passes=a.passes, want_lo=want_lo), 2 * 9 * C * C * px)
    timed("dgrad_mod", lambda: ops.conv3x3([dgb], pwmT, None, passes=a.passes, tag="dgrad_mod"),
          2 * 9 * (nh + d) * 2 * C * px)
    timed("wgrad modulation", lambda: ops.conv3x3_wgrad_multi(dgb, [actv, smap], passes=a.passes),
          2 * 9 * (nh + d) * 2 * C * px)

    # ---- HBM-bound kernels of the same block (algorithmic bytes = one read of each input + one write
    # of each output; in-kernel noise costs no bytes) -------------------------------------------
    def hbm(name, fn, nbytes):
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
        print("%-34s %8.3f ms  %7.1f GB/s (algorithmic)" % (name, ms, nbytes / ms / 1e6))
        return r

    n = px * C
    hb = 4 if want_lo else 2                  # bytes per element of a split-plane operand
    nw = torch.full((C,), 0.1, device=dev)
    seed = ops.NoiseSeed(1234567)
    x_lo = rn(B, S // 2, S // 2, C)
    hbm("grad_prep (no noise)", lambda: ops.grad_prep(dy, want_lo=want_lo), n * (4 + 4 + hb))
    hbm("grad_prep (+2 noise sums)", lambda: ops.grad_prep(dy, seed, ops.NoiseSeed(77), want_lo=want_lo),
        n * (4 + 4 + hb))
    hbm("K1 backward (saved G)", lambda: ops.spade_modulate_bwd_saved(gsaved, x, 0, sc, sh, dt, amax,
                                                                      want_lo=want_lo), n * (hb + 4 + 4 + 4 + 2 * hb))
    hbm("K1 backward (saved G, ups+noise)", lambda: ops.spade_modulate_bwd_saved(
        gsaved, x_lo, 1, sc, sh, dt, amax, noise=seed, noise_w=nw, want_lo=want_lo), n * (hb + 1 + 4 + 4 + 2 * hb))
    hbm("bn_bwd", lambda: ops.bn_bwd(dxhat, x, 0, sc, sh, sums, 1.0 / px), n * 12)
    hbm("bn_bwd (ups+noise+dskip)", lambda: ops.bn_bwd(dxhat, x_lo, 1, sc, sh, sums, 1.0 / px, noise=seed,
                                                       noise_w=nw, dskip=dy), n * (4 + 4 + 1 + 1))
    hbm("bn_stats (ups+noise)", lambda: ops.bn_stats(x_lo, 1, seed, nw), n * 1)
    hbm("bn_stats", lambda: ops.bn_stats(x, 0), n * 4)
    wh, bh = rn(3, C, 3, 3) / (3 * C ** 0.5), rn(3)
    out = hbm("head fwd", lambda: ops.head(x, wh, bh), n * 4 + px * 12)
    dout = rn(B, 3, S, S)
    hbm("head bwd", lambda: ops.head_bwd(x, wh, out, dout), n * 8 + px * 24)
    hbm("shared_mlp (table gather)", lambda: ops.shared_mlp(labels, table, tb, want_lo=want_lo),
        px * (1 + nh * hb))
    hbm("style_gather", lambda: ops.style_gather(labels, rn(B, L, d), want_lo=want_lo), px * (1 + d * hb))
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
