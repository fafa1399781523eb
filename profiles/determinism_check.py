#!/usr/bin/env python
"""Runs K training iterations of the benched configuration from a fixed seed and prints a digest of every
parameter: two invocations must print the same digest (bit-level run-to-run determinism of the step).

    python profiles/determinism_check.py [--iters 6] [--batch 8] [--graphs 0|1]
"""
import argparse
import hashlib
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--force", default=None, help="fixed coin flips for every sub-step, e.g. 10 = (full, noisy)")
    ap.add_argument("--rseed", type=int, default=0, help="seed of Python's `random` (the encoder coin flips)")
    a = ap.parse_args()
    from deepsee_b200.config import config
    config.passes = a.passes
    config.cuda_graphs = bool(a.graphs)
    from deepsee_b200.managers.trainer_manager import TrainerManager
    from deepsee_b200.options.configurations import make_opt
    from deepsee_b200.util.synthetic import settle_spectral_norm, synthetic_batch
    torch.manual_seed(0)
    random.seed(a.rseed)
    o = make_opt("8x_independent_256x256", isTrain=True, gpu_ids=[0], batchSize=a.batch)
    mgr = TrainerManager(o)
    model = mgr.sr_model
    for net in (model.netSR, model.netE, model.netD):
        settle_spectral_norm(net)
    with torch.no_grad():
        for n_, p_ in model.netSR.named_parameters():
            if ".noise_" in n_:
                p_.fill_(0.1)
    model.train(True)
    torch.manual_seed(1234)
    dev = {k: (v.float() if "label" in k else v).cuda() for k, v in synthetic_batch(o, a.batch, seed=1234).items()}
    real_random = random.random
    drawn = []

    forced = {"i": 0}

    def recording():
        v = real_random()
        if a.force:
            v = 0.25 if a.force[forced["i"] % len(a.force)] == "1" else 0.75
            forced["i"] += 1
        drawn.append(v < 0.5)
        return v
    for it in range(a.iters):
        random.random = recording
        del drawn[:]
        forced["i"] = 0
        mgr.run_generator_one_step(dict(dev))
        ng = len(drawn)
        mgr.run_discriminator_one_step(dict(dev))
        random.random = real_random
        flips = "G%s D%s" % ("".join("01"[f] for f in drawn[:ng]), "".join("01"[f] for f in drawn[ng:]))
        amax = " ".join("%s=%.3g" % (nm, max(float(p.detach().abs().max()) for p in net.parameters()))
                        for nm, net in (("SR", model.netSR), ("E", model.netE), ("D", model.netD)))
        losses = {k: float(v.detach().mean()) for k, v in mgr.get_latest_losses().items()}
        h = hashlib.md5()
        for net in (model.netSR, model.netE, model.netD):
            for _, p in sorted(net.state_dict().items()):
                h.update(p.detach().cpu().numpy().tobytes())
        bad = [nm + "." + k for nm, net in (("SR", model.netSR), ("E", model.netE), ("D", model.netD))
               for k, p in net.named_parameters() if not bool(torch.isfinite(p).all())]
        if bad:
            print("   non-finite parameters (%d): %s" % (len(bad), bad[:6]))
        print("iter %d  params %s  flips %s  max|p| %s  %s" % (
            it, h.hexdigest()[:12], flips, amax, {k: round(v, 5) for k, v in losses.items()}), flush=True)


if __name__ == "__main__":
    main()
