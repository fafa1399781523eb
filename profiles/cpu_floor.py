#!/usr/bin/env python
"""Host-side floor of one training iteration: the same ~1200 launches per iteration as the c2
workload, but on a model so small (8 base channels, 64x64) that the kernels take no time - what is
left is Python + ctypes + allocator + launch overhead.  If this number approaches the c2 step time,
the step is launch-bound and fusing / batching launches is what pays.

    python profiles/cpu_floor.py [--iters 30]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsee_b200 import _lib  # noqa: E402
from deepsee_b200.managers.trainer_manager import TrainerManager  # noqa: E402
from deepsee_b200.options.configurations import make_opt  # noqa: E402
from deepsee_b200.util.synthetic import synthetic_batch, settle_spectral_norm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    sys.stdout, out = sys.stderr, sys.stdout
    o = make_opt("8x_independent_256x256", isTrain=True, ngf=8, nef=8, ndf=8, start_size=8, crop_size=64,
                 load_size=64, batchSize=8)
    mgr = TrainerManager(o)
    for net in (mgr.sr_model.netSR, mgr.sr_model.netE, mgr.sr_model.netD):
        settle_spectral_norm(net)
    mgr.sr_model.train()
    raw = synthetic_batch(o, 8)
    dev = {k: (v.float() if "label" in k else v).cuda() for k, v in raw.items()}

    def it():
        mgr.run_generator_one_step(dict(dev))
        mgr.run_discriminator_one_step(dict(dev))
    for _ in range(5):
        it()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        it()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("host floor: %.1f ms per iteration to issue (%.1f ms incl. drain), %d deepsee_b200 launches per iteration"
          % (1000 * t_issue / a.iters, 1000 * t_all / a.iters, (_lib.launch_count() - n0) / a.iters), file=out)


if __name__ == "__main__":
    main()
