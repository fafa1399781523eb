#!/usr/bin/env python
"""The image head (leaky_relu -> conv_img 512 -> 3 -> tanh, sr.py:94-95) at c2's shape: the fp32 CUDA-core
kernels against the tensor-core form, piece by piece (CUDA events, warm).

    python profiles/head_microbench.py [--B 8 --S 256]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsee_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--S", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    B, S, C = a.B, a.S, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, S, S, C, generator=g, device="cuda")
    w = torch.randn(3, C, 3, 3, generator=g, device="cuda") / (3 * C ** 0.5)
    b = torch.zeros(3, device="cuda")
    dout = torch.randn(B, 3, S, S, generator=g, device="cuda") * 1e-4
    act = torch.nn.functional.leaky_relu(x, 0.2)
    planes = ops.split_f16(act)
    gb = 1e-9

    def timed(name, fn, nbytes):
        for _ in range(2):
            r = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print("%-44s %7.3f ms  %6.0f GB/s (algorithmic bytes)" % (name, ms, nbytes * gb / (ms * 1e-3)))
        return r

    nx = x.numel()
    out = timed("head fp32 kernel (reads x fp32)", lambda: ops.head(x, w, b), nx * 4)
    out2 = timed("head tensor-core, 3-pass (reads hi + lo)", lambda: ops.head_tc(planes, w, b, passes=3), nx * 4)
    print("   max |tc - fp32| = %.2e" % (out - out2).abs().max().item())
    timed("head_bwd fp32 kernel (x in, dx out)", lambda: ops.head_bwd(x, w, out, dout), nx * 8)
    timed("head backward tensor-core, 1-pass (hi in, dx out)", lambda: ops.head_tc_bwd(planes, w, out, dout, passes=1), nx * 6)
    # the pieces of the tensor-core backward
    dP = torch.empty(B, S, S, 32, device="cuda")
    from deepsee_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    timed("  scatter dP", lambda: lib.dsee_head_scatter_bwd(dout.data_ptr(), out.data_ptr(), dP.data_ptr(), B, S, S, st),
          dP.numel() * 4)
    gp, _ = timed("  grad_prep(dP)", lambda: ops.grad_prep(dP, want_lo=False), dP.numel() * 6)
    hi = ops.SplitPlanes(planes.hi, None)
    timed("  wgrad 1x1 [32 x px] x [px x 512]", lambda: ops.conv2d_tc_wgrad(gp, hi, (32, C, 1, 1), 1, 0, passes=1), nx * 2)
    pwT = ops.prep_conv_weight_ex(ops._head_w27(w), False, transpose=True, rows=C)
    timed("  dgrad 1x1 (K = 64, N = 512) + lrelu' mask", lambda: ops.conv2d_tc(
        gp, pwT, None, 1, 1, 1, 0, (S, S), passes=1, transposed=True, act_mask=planes.hi, want_amax=True), nx * 6)


if __name__ == "__main__":
    main()
