#!/usr/bin/env python
"""Benchmark of the DeepSEE hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c4]
                    [--passes 1|3]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one per-GPU batch of
synthetic input: BASELINE.json config c2 (8x SR, 256x256, independent model, batch 8 per GPU) by
default; c4 = 32x SR 512x512 independent, batch 2 per GPU.  Inputs of the `value` measurement are
resident in HBM; `e2e` goes through BaseManager.preprocess + SRModel.forward with pinned host
buffers and the H2D / D2H copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c2": dict(name="8x_independent_256x256", batch=8, flops_fwd_img=1725.1e9 + 5.0e9, d_pair=5.49e9),
    "c4": dict(name="32x_independent_512x512", batch=2, flops_fwd_img=5397.4e9 + 1.3e9, d_pair=20.9e9),
}
METRIC = "generator+discriminator images/sec"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["bf16_tflops"], d["hbm_gbs"], "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = max([int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()] or [0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": reasons, "samples": len(sm)}


def build_opt(o):
    from deepsee_b200.options.configurations import make_opt
    d = dict(o)
    name = d.pop("name")
    opt = make_opt(None, **d)
    opt.name = name
    return opt


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port: the reference is pure Python
    over ATen, nothing to compile) on the host cores; each step = a bounded sample (1 image)."""
    import torch
    from oracle import deepsee_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    o = O.make_opt(cfg["name"], is_train=True)
    sdG, sdE, sdD = O.make_generator_state(o, 0), O.make_encoder_state(o, 1), O.make_discriminator_state(o, 2)
    d = O.preprocess(o, O.synthetic_batch(o, 1, seed=1234))

    def step():
        with torch.no_grad():
            fake, _ = O.inference(sdG, sdE, o, d["image_lr"], d["input_semantics"], d["image_hr"])
            O.discriminate(sdD, o, d["input_semantics"], fake, d["image_hr"], training=False)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC % (cfg["name"], 1), "precision": "fp32 (ATen CPU)"},
        "cpu_baseline": {"value": v, "unit": "images/sec", "cores": cores, "kind": "port",
                         "sample": "1 image per step (batch 1), oracle port of the reference forward"},
        "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


WORKLOAD_DESC = ("%s, batch %d per GPU: style encoder + SPADE/SEAN generator forward + multi-scale "
                 "discriminator forward on [fake|real] (inference path; backward not yet native)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--passes", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from oracle import deepsee_oracle as O      # synthetic inputs / seeded weights / CPU baseline only
    from deepsee_b200 import _lib, ops, parallel
    from deepsee_b200.config import config
    from deepsee_b200.managers.base_manager import BaseManager

    if args.passes:
        config.passes = args.passes
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    parallel.init_from_env()
    cfg = CONFIGS[args.config]
    b = cfg["batch"]
    o = O.make_opt(cfg["name"], is_train=True)
    mgr = BaseManager(build_opt(o))
    model = mgr.sr_model.eval()
    model.netSR.load_state_dict(O.make_generator_state(o, 0), strict=True)
    model.netE.load_state_dict(O.make_encoder_state(o, 1), strict=True)
    model.netD.load_state_dict(O.make_discriminator_state(o, 2), strict=True)

    raw = O.synthetic_batch(o, b, seed=1234 + rank)
    host = {"label": raw["label"].float().pin_memory(), "image": raw["image"].pin_memory()}
    dev = mgr.preprocess({k: v.clone() for k, v in host.items()}, from_dataloader=True)
    torch.cuda.synchronize()

    def hot_step(data):
        with torch.no_grad():
            out = model(dict(data), "inference")
            pf, pr = model.discriminate(data["input_semantics"], out["fake_image"], data["image_hr"])
        return out["fake_image"], pf

    def e2e_step():
        data = mgr.preprocess({k: v for k, v in host.items()}, from_dataloader=True)
        fake, pf = hot_step(data)
        return fake.cpu(), float(pf[0][-1].mean().cpu())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # inputs (b x 19 x S x S fp32 one-hot etc.) and every activation exceed L2 (126 MB) at these
    # sizes, so no explicit flush is needed between iterations.
    for _ in range(args.warmup):
        hot_step(dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ops.KernelTimer() as kt:
        e0.record()
        for _ in range(args.steps):
            hot_step(dev)
        e1.record()
        barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1)
    ksum = kt.summary()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # end to end through the public API with host buffers
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    sampler.stop_flag = True
    if rank != 0:
        return

    S = o.crop_size
    imgs = b * world * args.steps
    value = imgs / (ms / 1000.0)
    sust, burst, hbm, how = peaks()
    # dominant kernel family = the tcgen05 implicit-GEMM kernel (K2 conv and K1 modulate share it)
    tot_ms = sum(v[1] for v in ksum.values())
    tot_fl = sum(v[2] for v in ksum.values())
    top = max(ksum.items(), key=lambda kv: kv[1][1])
    top_tflops = top[1][2] / (top[1][1] / 1000.0) / 1e12
    roofline = {
        "bound": "tensor", "kernel": "conv3x3_tc_kernel (%s)" % top[0],
        "achieved": top_tflops, "peak": sust, "unit": "TFLOP/s", "frac": top_tflops / sust,
        "peak_source": "%s bf16 dense sustained (MEASURED_PEAKS.json); fp16 operands, same pipe" % how,
        "executed_passes": config.passes,
        "executed_frac": top_tflops * config.passes / sust,
        "all_tc_launches": {"achieved": tot_fl / (tot_ms / 1000.0) / 1e12,
                            "share_of_step": tot_ms / ms, "launches_per_step": sum(v[0] for v in ksum.values()) / args.steps},
        "traffic": None,
    }
    line = {
        "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32-class (fp16 split x%d, fp32 accumulate)" % config.passes
        if config.passes == 3 else "fp16 operands (TF32-class), fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC % (cfg["name"], b), "global_batch": b * world,
                   "image": "%dx%d" % (S, S), "parallelism": "dp%d" % world,
                   "l2": "inputs and activations larger than L2, no flush", "passes": config.passes},
        "gpu_launches": launches,
        "clocks": sampler.result(),
        "e2e": {"value": imgs / e2e_s, "unit": "images/sec",
                "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())),
                "d2h_bytes_per_step": int(b * 3 * S * S * 4 + 4)},
        "roofline": roofline,
    }
    if not args.no_cpu_baseline:
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        sdG, sdE, sdD = O.make_generator_state(o, 0), O.make_encoder_state(o, 1), O.make_discriminator_state(o, 2)
        d1 = O.preprocess(o, O.synthetic_batch(o, 1, seed=1234))
        n = 0
        t0 = time.perf_counter()
        while n < 2 or (time.perf_counter() - t0 < 12 and n < 6):
            with torch.no_grad():
                fk, _ = O.inference(sdG, sdE, o, d1["image_lr"], d1["input_semantics"], d1["image_hr"])
                O.discriminate(sdD, o, d1["input_semantics"], fk, d1["image_hr"], training=False)
            n += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": "images/sec", "cores": cores, "kind": "port",
                                "sample": "%d images, batch 1, same workload (oracle port of the "
                                          "reference forward, torch CPU fp32)" % n}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
