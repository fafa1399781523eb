#!/usr/bin/env python
"""Benchmark of the DeepSEE hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]
                    [--batch B] [--passes 1|3] [--mode train|infer] [--no-graph] [--no-extra]

One JSON line on stdout (rank 0).  A "step" is ONE TRAINING ITERATION of the reference's loop
(train.py:60-68: TrainerManager.run_generator_one_step + run_discriminator_one_step) over one
per-GPU batch of synthetic input: style encoder + SPADE/SEAN generator forward and backward,
multi-scale discriminator forward/backward, hinge + feature-matching losses, both Adam updates and,
for N > 1, the NCCL all-reduce of the G+E and D gradients plus the Sync-BN statistics exchange.
Headline workload = BASELINE.json config c2 (8x SR, 256x256, independent model, batch 8 per GPU).
The default run also measures, into the line's `configs` block, c4 weak (32x SR 512x512 independent,
2 images per GPU), c4 strong (global batch 8 split over the GPUs) and c5 (32x guided, 4 images per
GPU), each with img/s, ms/step, whole-step fraction of the tensor peak and e2e.

`value`: the raw batch resident in HBM; every optimizer sub-step a CUDA graph replay (--no-graph:
eager).  `e2e`: pinned HOST buffers through the same TrainerManager calls (H2D copies inside the timed
region), losses read back to the host every step.  `roofline`: CUDA events around every tensor-core
launch (eager steps right after the timed region when the timed steps are graph replays).
`parity`: max-abs of this configuration's full-size generator in this run's precision mode against
the CPU oracle (rank 0, N = 1, next to `cpu_baseline`).  `ddp_parity` / `rank_isolated_eager_ms_per_step`
(N > 1): sharded-vs-single-process agreement on the real NCCL ranks, and every rank's step time alone.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic (reference-equivalent dense conv) FLOPs per sample: SURVEY.md section 8d
CONFIGS = {
    "c2": dict(name="8x_independent_256x256", batch=8, train_flops=6.95e12, fwd_flops=1725.1e9 + 5.0e9 + 5.49e9),
    "c3": dict(name="8x_guided_256x256", batch=8, train_flops=7.02e12, fwd_flops=1725.1e9 + 20.7e9 + 5.49e9),
    "c4": dict(name="32x_independent_512x512", batch=2, train_flops=21.72e12, fwd_flops=5397.4e9 + 1.3e9 + 20.9e9),
    "c5": dict(name="32x_guided_512x512", batch=4, train_flops=22.05e12, fwd_flops=5397.4e9 + 82.6e9 + 20.9e9),
}
METRIC = "generator+discriminator images/sec"
WORKLOAD = ("%s, batch %d per GPU, one full training iteration per step (style encoder + SPADE/SEAN "
            "generator fwd+bwd, multi-scale discriminator fwd+bwd on [fake|real], hinge + feature-"
            "matching losses, Adam for G+E and D%s)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["bf16_tflops"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 1400.0, 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md's clocks line): ONE
    `nvidia-smi -lms 200` process running in the background - no per-sample process spawn next to
    the launching thread."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.stop_flag = False

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def result(self):
        samples = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                samples = [[f.strip() for f in line.split(",")] for line in out.splitlines() if line.strip()]
            except Exception:
                pass
        sm = sorted(int(s_[0]) for s_ in samples if s_[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(s_) > 2 + i and s_[2 + i].lower().startswith("active") for s_ in samples)]
        mx = max([int(s_[1]) for s_ in samples if len(s_) > 1 and s_[1].isdigit()] or [0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": reasons, "samples": len(sm)}


def cpu_train_iteration_timer(cfg, threads):
    """The reference's training iteration on the host cores: the oracle's restatement of
    trainer_manager.py:32-61 (the reference is pure Python over ATen, so there is nothing to
    compile; `kind` = "port").  Returns a closure running ONE iteration on a batch of 1."""
    import torch
    from oracle import deepsee_oracle as O
    torch.set_num_threads(threads)
    o = O.make_opt(cfg["name"], is_train=True)
    tr = O.CpuTrainer(o, O.make_generator_state(o, 0), O.make_encoder_state(o, 1),
                      O.make_discriminator_state(o, 2))
    d = O.preprocess(o, O.synthetic_batch(o, 1, seed=1234))

    def step():
        tr.generator_step(d)
        tr.discriminator_step(d)
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count()
    step = cpu_train_iteration_timer(cfg, cores)
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
        "config": {"workload": WORKLOAD % (cfg["name"], 1, ""), "precision": "fp32 (ATen CPU)"},
        "cpu_baseline": {"value": v, "unit": "images/sec", "cores": cores, "kind": "port",
                         "sample": "1 image per step (batch 1): one full G+D training iteration of the "
                                   "oracle port of trainer_manager.py:32-61, torch CPU fp32"},
        "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def kernel_traffic(tag, cfg_key):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of a kernel group inside the timed
    step, from the committed `ncu --set full` capture (profiles/r2_kernel_traffic.csv: tag, config,
    read bytes, written bytes, source).  None when that launch group has not been captured."""
    import csv
    p = os.path.join(ROOT, "profiles", "r2_kernel_traffic.csv")
    if not os.path.exists(p):
        return None, None
    for row in csv.DictReader(open(p)):
        if row["tag"] == tag and row["config"] == cfg_key:
            return float(row["dram_read_bytes"]) + float(row["dram_write_bytes"]), row["source"]
    return None, None


def parity_leg(cfg):
    """One batch-1 generator forward of THIS configuration's full-size generator in THIS run's
    precision mode, train mode (batch statistics, noise injection), against the CPU oracle on the
    same conditioned weights / inputs / noise tensors -> max-abs on the tanh output.  Checker only:
    runs after the timed regions, on rank 0, next to the cpu_baseline leg."""
    import torch
    from oracle import deepsee_oracle as O
    from deepsee_b200.options.configurations import make_opt
    from deepsee_b200.deepsee_models.networks.sr import DeepSEESR
    o = O.make_opt(cfg["name"], is_train=True)
    sd = O.make_generator_state(o, 0)
    d = O.preprocess(o, O.synthetic_batch(o, 1, seed=70))
    z = torch.rand(1, 19, 128, generator=torch.Generator().manual_seed(5)) * 2 - 1
    noises = {}

    def noise_fn(nm, shape):
        noises[nm] = torch.randn(shape, generator=torch.Generator().manual_seed(len(noises)))
        return noises[nm]

    with torch.no_grad():
        ref = O.generator_forward({k: v.clone() for k, v in sd.items()}, o, d["image_lr"],
                                  d["input_semantics"], z, True, noise_fn)
    od = dict(o)
    name = od.pop("name")
    opt = make_opt(None, **od)
    opt.name = name
    G = DeepSEESR(opt).cuda()
    G.load_state_dict(sd, strict=True)
    G.train()
    if o.add_noise:
        for pfx, _, _ in O.generator_layout(o):
            blk = G.get_submodule(pfx[:-1])
            for nm in ("noise_in", "noise_skip", "noise_middle"):
                n = noises[pfx + nm].permute(0, 2, 3, 1).contiguous().cuda()
                getattr(blk, nm).sample = (lambda t: (lambda B, H, W: t))(n)
    with torch.no_grad():
        out = G(d["image_lr"].cuda(), seg=d["input_semantics"].cuda(), z=z.cuda())
    e = (out.cpu() - ref).abs()
    return {"max_abs": e.max().item(), "mean_abs": e.mean().item(), "ref_std": ref.std().item(),
            "max_abs_over_std": e.max().item() / ref.std().item(), "tolerance": 1e-3,
            "within_tolerance": bool(e.max().item() < 1e-3),
            "what": "full-size %s generator, batch 1, train mode (batch statistics, noise injection), this "
                    "run's precision mode vs the CPU oracle on conditioned weights; asserted at this shape "
                    "by tests/test_full_size_parity_gpu.py" % cfg["name"]}


def ddp_parity_leg(world, rank, local):
    """Data-parallel parity on the real NCCL ranks of this launch (N > 1): ONE training iteration of
    a small model (16x16 -> 64x64, 128 channels) on a fixed global batch, (a) sharded over the ranks
    with the gradient buckets / Sync-BN exchange of the product, (b) on the whole batch inside this
    process with the process group ignored (parallel.local_mode).  Returns the largest relative
    differences of the rank-averaged losses, the all-reduced gradients and the BN running statistics
    - what tests/test_ddp_gpu.py asserts on a 2-GPU box, recorded by every multi-GPU bench run."""
    import random
    import torch
    import torch.distributed as dist
    from deepsee_b200 import parallel
    from deepsee_b200.config import config
    from deepsee_b200.managers.trainer_manager import TrainerManager
    from deepsee_b200.options.configurations import make_opt
    from deepsee_b200.util.synthetic import synthetic_batch, settle_spectral_norm
    saved = config.passes
    config.passes = 3
    per = 2
    kw = dict(isTrain=True, gpu_ids=[local], ngf=8, nef=8, ndf=8, start_size=16, crop_size=64, load_size=64,
              add_noise=False, noisy_style_scale=0.0)

    def build(batch):
        torch.manual_seed(7)
        mgr = TrainerManager(make_opt("8x_independent_256x256", batchSize=batch, **kw))
        for net in (mgr.sr_model.netSR, mgr.sr_model.netE, mgr.sr_model.netD):
            settle_spectral_norm(net)
        parallel.broadcast_module(mgr.sr_model)
        mgr.sr_model.train()
        return mgr

    def grads(mgr):
        m = mgr.sr_model
        named = list(m.netSR.named_parameters()) + [("E." + n, p) for n, p in m.netE.named_parameters()]
        return {n: p.grad.detach().clone() for n, p in named if p.grad is not None}

    try:
        raw = synthetic_batch(make_opt("8x_independent_256x256", **kw), per * world, seed=4242)
        raw["label"] = raw["label"].float()
        # (a) sharded
        mgr = build(per)
        random.seed(0)
        shard = {k: v[rank * per:(rank + 1) * per].clone() for k, v in raw.items()}
        mgr.run_generator_one_step(dict(shard))
        g_ddp = grads(mgr)
        loss_ddp = {}
        for k, v in mgr.get_latest_losses().items():
            t = v.detach().mean().reshape(1).clone()
            dist.all_reduce(t)
            loss_ddp[k] = float(t) / world
        rs_ddp = mgr.sr_model.netSR.state_dict()["G_middle_1.norm_1.param_free_norm.running_var"].clone()
        # (b) whole batch, single-process semantics
        with parallel.local_mode():
            ref = build(per * world)
            random.seed(0)
            ref.run_generator_one_step({k: v.clone() for k, v in raw.items()})
            g_ref = grads(ref)
            loss_ref = {k: float(v.detach().mean()) for k, v in ref.get_latest_losses().items()}
            rs_ref = ref.sr_model.netSR.state_dict()["G_middle_1.norm_1.param_free_norm.running_var"]
        worst, worst_name = 0.0, None
        # (gradients that are zero in exact arithmetic - e.g. a conv bias in front of a batch norm - are
        # rounding noise on both sides: each tensor's scale is floored at 1e-4 of the largest gradient)
        gmax = max(float(g.abs().max()) for g in g_ref.values())
        for n, gr in g_ref.items():
            if n not in g_ddp:
                worst, worst_name = float("inf"), n + " (missing)"
                break
            e = float((g_ddp[n] - gr).abs().max()) / max(float(gr.abs().max()), 1e-4 * gmax)
            if e > worst:
                worst, worst_name = e, n
        res = {"ranks": world, "per_rank_batch": per, "sync_bn": bool(config.sync_bn_for("syncbatch")),
               "loss_rel_diff": max(abs(loss_ddp[k] - loss_ref[k]) / max(1.0, abs(loss_ref[k])) for k in loss_ref),
               "grad_max_rel_diff": worst, "grad_worst": worst_name, "grad_tensors": len(g_ref),
               "grads_only_where_single_process_has_them": set(g_ddp) == set(g_ref),
               "sync_bn_over_peer_memory": parallel.peer_exchange_active(),
               "bn_running_var_rel_diff": float((rs_ddp - rs_ref).abs().max() / rs_ref.abs().max()),
               "what": "1 training iteration, %d ranks x %d samples (NCCL gradient buckets%s) vs 1 process x %d "
                       "samples, passes=3" % (world, per, " + Sync-BN" if config.sync_bn_for("syncbatch") else "",
                                              per * world)}
        t = torch.tensor([res["loss_rel_diff"], res["grad_max_rel_diff"], res["bn_running_var_rel_diff"]],
                         device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["loss_rel_diff"], res["grad_max_rel_diff"], res["bn_running_var_rel_diff"] = [float(x) for x in t]
        return res
    finally:
        config.passes = saved


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch of the headline config")
    ap.add_argument("--passes", type=int, default=None)
    ap.add_argument("--passes3-upto", default=None,
                    help="main convs up to this feature-map height run 3 passes (default 'auto' = output/4; 0 = off)")
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the `configs` block (c4 weak + strong, c5) measured after the headline config")
    ap.add_argument("--extra-steps", type=int, default=5)
    ap.add_argument("--no-graph", action="store_true",
                    help="run the timed steps eagerly (default: each optimizer sub-step is a CUDA graph replay)")
    ap.add_argument("--roof-steps", type=int, default=3,
                    help="eager steps after the timed region that bracket every tensor-core launch with CUDA "
                         "events (the roofline tables; graph replays cannot be bracketed per kernel)")
    ap.add_argument("--cprofile", default=None,
                    help="write a cProfile table (main thread: forward passes, optimizers) of 3 extra steps to this file")
    ap.add_argument("--torch-profile", default=None,
                    help="write a torch.profiler (CUPTI) kernel table of 2 extra steps to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line: anything the reference-shaped host code prints while it
    # builds the model (e.g. SRModel.create_optimizers' "lr G: ..." line, sr_model.py:486) goes to stderr
    # (NCCL and other native libraries write to file descriptor 1 directly: point fd 1 at stderr for the
    # duration of the run and keep a private duplicate of the real stdout for the JSON line)
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr

    import gc
    import torch
    import torch.distributed as dist
    # this arm touches the product only; oracle/ is imported by the cpu_baseline / parity legs alone
    from deepsee_b200 import _lib, ops, parallel
    from deepsee_b200.config import config
    from deepsee_b200.managers.trainer_manager import TrainerManager
    from deepsee_b200.options.configurations import make_opt
    from deepsee_b200.util.synthetic import synthetic_batch, settle_spectral_norm

    # Precision of the measured step: 1 pass = fp16 operands with fp32 accumulation (the TF32 class,
    # 10-bit mantissa, that stock PyTorch/cuDNN runs the reference's convs in on this GPU), with the
    # main convs of the low-resolution stages (<= output/4, ~5 % of the FLOPs) at 3 passes: the cheapest
    # mode that keeps the full-size generators inside north_star's 1e-3 max-abs bound
    # (tests/test_full_size_parity_gpu.py; `parity` below is measured by this very run).
    # --passes 3 measures the fp32-class split-operand mode (the library default, DSEE_PASSES).
    config.passes = args.passes or 1
    config.cuda_graphs = (not args.no_graph) and args.mode == "train"
    if args.passes3_upto is not None:
        config.passes3_upto = args.passes3_upto if args.passes3_upto == "auto" else int(args.passes3_upto)
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    parallel.init_from_env()
    train = args.mode == "train"
    sust, burst, hbm, how = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def measure(cfg_key, b, steps, warmup, headline):
        """Builds the named model, runs `warmup` + `steps` training iterations on a per-GPU batch of
        b device-resident samples (CUDA events, max over ranks), then the same through pinned HOST
        buffers (e2e).  Returns the result dict."""
        cfg = CONFIGS[cfg_key]
        torch.manual_seed(0)                         # same random-init weights on every rank
        import random
        random.seed(1234)                            # the encoder coin flips (sr_model.py:616,643): same on every
                                                     # rank (they must agree) and the same in every run
        o = make_opt(cfg["name"], isTrain=True, gpu_ids=[local], batchSize=b)
        mgr = TrainerManager(o)                      # random-init weights of the named architecture
        model = mgr.sr_model
        for net in (model.netSR, model.netE, model.netD):
            settle_spectral_norm(net)
        with torch.no_grad():                        # NoiseInjection.weight is zero-initialised (normalization.py:297)
            for n_, p_ in model.netSR.named_parameters():
                if ".noise_" in n_:
                    p_.fill_(0.1)
        parallel.broadcast_module(model)
        model.train(train)
        torch.manual_seed(1234 + rank)               # NoiseInjection seeds differ per rank
        raw = synthetic_batch(o, b, seed=1234 + rank)
        host = {k: (v.float() if "label" in k else v).pin_memory() for k, v in raw.items()}
        dev = {k: v.cuda() for k, v in host.items()}
        torch.cuda.synchronize()

        def iteration(data):
            if train:
                mgr.run_generator_one_step(dict(data))
                mgr.run_discriminator_one_step(dict(data))
                return mgr.get_latest_losses()
            with torch.no_grad():
                d = mgr.preprocess(dict(data), from_dataloader=True)
                out = model(d, "inference")
                pf, _ = model.discriminate(d["input_semantics"], out["fake_image"], d["image_hr"])
            return {"pred": pf[0][-1].mean()}

        def e2e_iteration():
            losses = iteration(host)
            return {k: float(v.detach().mean()) for k, v in losses.items()}  # D2H read of the step's result

        # The raw batch (b x 4 x S x S fp32) plus every activation of the step (several GB) exceed the
        # 126 MB L2 many times over, so successive iterations cannot hit in L2; no explicit flush.
        for _ in range(warmup):
            iteration(dev)
        if train and config.cuda_graphs:
            mgr.warm_graphs(dict(dev))   # capture every coin-flip variant now, not inside the timed region
            iteration(dev)
        barrier()
        graphed = bool(train and config.cuda_graphs and mgr.graphs_active())
        sampler = ClockSampler(local)
        if rank == 0 and headline:
            sampler.start()
        n0 = _lib.launch_count() + (mgr.graph_launches if train else 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if headline:
            torch.cuda.profiler.start()  # `ncu --profile-from-start off` captures exactly the timed region
        with ops.KernelTimer() as kt:
            e0.record()
            for _ in range(steps):
                iteration(dev)
            e1.record()
            barrier()
        if headline:
            torch.cuda.profiler.stop()
        launches = _lib.launch_count() + (mgr.graph_launches if train else 0) - n0
        ms = max_over_ranks(e0.elapsed_time(e1))
        ksum = kt.summary()
        roof_ms, roof_steps = ms, steps
        if graphed and (headline or not args.no_extra):
            # the per-kernel CUDA-event brackets need eager launches: same model, same batch, right after
            config.cuda_graphs = False
            try:
                iteration(dev)
                barrier()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with ops.KernelTimer() as kt:
                    r0.record()
                    for _ in range(args.roof_steps):
                        iteration(dev)
                    r1.record()
                    barrier()
                ksum = kt.summary()
                roof_ms, roof_steps = max_over_ranks(r0.elapsed_time(r1)), args.roof_steps
            finally:
                config.cuda_graphs = True
        clocks = sampler.result() if (rank == 0 and headline) else None

        if headline and args.torch_profile and rank == 0:
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                t0 = time.perf_counter()
                for _ in range(2):
                    iteration(dev)
                torch.cuda.synchronize()
                wall = time.perf_counter() - t0
            with open(args.torch_profile, "w") as f:
                f.write("2 steps under the profiler: %.1f ms wall\n" % (wall * 1000))
                f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=90,
                                                  max_name_column_width=70))
        if headline and args.cprofile and rank == 0:
            import cProfile
            import io
            import pstats
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(3):
                iteration(dev)
            torch.cuda.synchronize()
            pr.disable()
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(70)
            pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(45)
            open(args.cprofile, "w").write(buf.getvalue())
        isolated = None
        if headline and world > 1 and train:
            # every rank alone: the same eager step with the process group ignored (no gradient
            # exchange, per-rank BN statistics).  The job runs at the pace of its slowest GPU, so the
            # spread of these times bounds what data parallelism can reach on this box.
            was = config.cuda_graphs
            config.cuda_graphs = False
            try:
                with parallel.local_mode():
                    iteration(dev)
                    torch.cuda.synchronize()
                    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    l0.record()
                    for _ in range(4):
                        iteration(dev)
                    l1.record()
                    torch.cuda.synchronize()
                mine = torch.tensor([l0.elapsed_time(l1) / 4], device="cuda", dtype=torch.float64)
            finally:
                config.cuda_graphs = was
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            isolated = [round(float(t.item()), 2) for t in allr]
        e2e = None
        if not args.no_e2e:
            for _ in range(2):
                e2e_iteration()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                last = e2e_iteration()
            torch.cuda.synchronize()
            e2e_s = max_over_ranks(time.perf_counter() - t0)
            import math
            if not all(math.isfinite(v) and abs(v) < 1e6 for v in last.values()):
                raise RuntimeError("bench: the training iteration diverged (losses %r)" % (last,))
            e2e = {"value": b * world * steps / e2e_s, "unit": "images/sec",
                   "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())) * (2 if train else 1),
                   "d2h_bytes_per_step": 4 * len(last), "last_losses": last}
        peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
        S = o.crop_size
        sync_bn = bool(world > 1 and config.sync_bn_for(o.norm_G))
        del mgr, model, dev, host, raw
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        value = b * world * steps / (ms / 1000.0)
        flops_per_img = cfg["train_flops"] if train else cfg["fwd_flops"]
        return dict(cfg=cfg, b=b, S=S, steps=steps, ms=ms, value=value, ksum=ksum, launches=launches,
                    isolated=isolated,
                    graphed=graphed, roof_ms=roof_ms, roof_steps=roof_steps,
                    clocks=clocks, e2e=e2e, peak_mem=peak_mem, sync_bn=sync_bn,
                    whole_step={"algorithmic_tflops_per_gpu": value / world * flops_per_img / 1e12,
                                "frac_of_peak": value / world * flops_per_img / 1e12 / sust})

    b0 = args.batch or CONFIGS[args.config]["batch"]
    try:
        R = measure(args.config, b0, args.steps, args.warmup, True)
    except RuntimeError as e:
        if config.cuda_graphs and world == 1 and "CUDA graph capture" in str(e):
            # never lose the measurement to a capture problem: same run, eager, in a fresh process
            print("bench.py: %s\nbench.py: re-running with --no-graph" % e, file=sys.stderr)
            sys.stderr.flush()
            os.dup2(real_stdout.fileno(), 1)
            os.execv(sys.executable, [sys.executable] + sys.argv + ["--no-graph"])
        raise

    # ---- the other BASELINE.json configurations, each with img/s, ms/step, whole-step fraction, e2e
    extras = []
    if train and not args.no_extra and args.config == "c2" and args.batch is None:
        plan = [("c4", 2, "weak", "BASELINE config 4 (32x 512x512 independent, 2 images per GPU: batch 8 on 4 GPUs)"),
                ("c5", 4, "weak", "BASELINE config 5 (32x 512x512 guided, 4 images per GPU: batch 32 on 8 GPUs)")]
        if 8 % world == 0:
            plan.insert(1, ("c4", 8 // world, "strong",
                            "32x 512x512 independent, GLOBAL batch 8 split over the GPUs (north_star's >= 6x at 8 GPUs "
                            "target, strong-scaling reading)"))
        for key, b, scaling, what in plan:
            r = measure(key, b, args.extra_steps, 3, False)
            extras.append({
                "config": key, "name": r["cfg"]["name"], "what": what, "scaling": scaling,
                "per_gpu_batch": b, "global_batch": b * world, "n_gpus": world, "image": "%dx%d" % (r["S"], r["S"]),
                "value": r["value"], "unit": "images/sec", "ms_per_step": r["ms"] / r["steps"], "steps": r["steps"],
                "warmup": 3, "whole_step": r["whole_step"],
                "e2e": {k: v for k, v in r["e2e"].items() if k != "last_losses"} if r["e2e"] else None,
                "tc_share_of_step": (sum(v[1] for v in r["ksum"].values()) / r["roof_steps"]) / (r["ms"] / r["steps"]),
                "cuda_graphs": r["graphed"],
                "sync_bn": r["sync_bn"], "peak_mem_gib": round(r["peak_mem"], 2),
            })
    ddp_parity = None
    if world > 1 and train and not args.no_extra:
        try:
            ddp_parity = ddp_parity_leg(world, rank, local)
        except Exception as e:  # recorded, never fatal for the measurement
            ddp_parity = {"error": repr(e)[:300]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    cfg, b, S, ms, value, ksum = R["cfg"], R["b"], R["S"], R["ms"], R["value"], R["ksum"]
    rsteps = R["roof_steps"]
    # tensor-core launches by family (CUDA events on the launching stream around each launch)
    fam = {}
    for tag, (n, t_ms, fl) in ksum.items():
        f = tag.rsplit("_", 1)[0]
        a = fam.setdefault(f, [0, 0.0, 0.0])
        a[0] += n
        a[1] += t_ms
        a[2] += fl
    tot_ms = sum(v[1] for v in ksum.values())
    tot_fl = sum(v[2] for v in ksum.values())
    top = max(ksum.items(), key=lambda kv: kv[1][1])
    top_tflops = top[1][2] / (top[1][1] / 1000.0) / 1e12
    traffic, traffic_src = kernel_traffic(top[0], args.config)
    # fp16-pass equivalents the dominant launch group executes per algorithmic FLOP: forward main convs
    # (tag conv3x3_*) run `k2_fwd_passes` of them in the benched mode (2 = one fp16 pass + the fp8
    # correction GEMM: twice the K at twice the MMA rate), everything is 3 in the fp32-class mode
    if config.passes == 3:
        exec_factor = 3
    elif top[0].startswith("conv3x3_"):
        exec_factor = config.k2_fwd_passes
    else:
        exec_factor = 1
    roofline = {
        "bound": "tensor", "kernel": "tcgen05 implicit-GEMM family; dominant launch group: %s" % top[0],
        "achieved": top_tflops, "peak": sust, "unit": "TFLOP/s", "frac": top_tflops / sust,
        "peak_source": "%s: bf16 dense sustained; kind::f16 operands run on the same pipe at the same rate" % how,
        "executed": {"fp16_pass_equivalents_per_algorithmic_flop": exec_factor,
                     "tflops_equivalent": top_tflops * exec_factor, "frac": top_tflops * exec_factor / sust,
                     "note": "`achieved` / `frac` count the reference's dense conv FLOPs once (algorithmic); the "
                             "kernel's tensor-pipe work is this many fp16-pass equivalents of them - the price of "
                             "the 1e-3 parity bound (DESIGN.md section 4)"},
        # per-launch CUDA-event brackets: over the timed steps when they run eagerly; with CUDA graphs
        # (the default) over `rsteps` eager steps of the same model / batch right after the timed region
        "timed_over": ("%d eager steps after the timed region (the timed steps are CUDA graph replays)" % rsteps)
                      if R["graphed"] else "the timed steps",
        "all_tc_launches": {"achieved": tot_fl / (tot_ms / 1000.0) / 1e12,
                            "share_of_step": (tot_ms / rsteps) / (ms / args.steps),
                            "eager_ms_per_step": R["roof_ms"] / rsteps,
                            "launches_per_step": sum(v[0] for v in ksum.values()) / rsteps},
        "by_family": {f: {"launches_per_step": a[0] / rsteps, "ms_per_step": a[1] / rsteps,
                          "tflops": a[2] / (a[1] / 1000.0) / 1e12} for f, a in sorted(fam.items())},
        "by_launch_group": {t: {"launches_per_step": v[0] / rsteps, "ms_per_step": v[1] / rsteps,
                                "tflops": v[2] / (v[1] / 1000.0) / 1e12} for t, v in sorted(ksum.items())},
        "whole_step": R["whole_step"],
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE in-step launch of the dominant launch group,
        # looked up in the committed ncu --set full summary (null if that group was not captured)
        "traffic": traffic, "traffic_source": traffic_src,
    }
    line = {
        "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": config.precision_name(),
        "data": "synthetic",
        "config": {"workload": WORKLOAD % (cfg["name"], b, ", NCCL all-reduce of G+E and D gradients" if world > 1 else "")
                   if train else "%s, batch %d per GPU, inference forward (encoder + generator + discriminator)" % (cfg["name"], b),
                   "global_batch": b * world, "image": "%dx%d" % (S, S), "parallelism": "dp%d" % world,
                   "l2": "inputs and activations larger than L2, no flush", "passes": config.passes,
                   "passes3_upto": config.passes3_upto, "sync_bn": R["sync_bn"], "mode": args.mode,
                   "cuda_graphs": R["graphed"]},
        "gpu_launches": R["launches"],
        "clocks": R["clocks"],
        "e2e": R["e2e"],
        "roofline": roofline,
        "peak_mem_gib": round(R["peak_mem"], 2),
    }
    if extras:
        line["configs"] = extras
    if ddp_parity is not None:
        line["ddp_parity"] = ddp_parity
    if R["isolated"] is not None:
        line["rank_isolated_eager_ms_per_step"] = R["isolated"]
    if not args.no_cpu_baseline and train and world == 1:
        cores = os.cpu_count()
        step = cpu_train_iteration_timer(cfg, cores)
        t0 = time.perf_counter()
        n = 0
        while n < 1 or (time.perf_counter() - t0 < 12 and n < 4):
            step()
            n += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": "images/sec", "cores": cores, "kind": "port",
                                "sample": "%d training iteration(s) at batch 1 of the same workload "
                                          "(oracle port of trainer_manager.py:32-61, torch CPU fp32)" % n}
        line["parity"] = parity_leg(cfg)
        for ex in extras:   # the 512x512 generator in the same mode (c4 and c5 share it)
            if ex["config"] == "c4" and ex["scaling"] == "weak":
                ex["parity"] = parity_leg(CONFIGS["c4"])
    print(json.dumps(line), file=real_stdout, flush=True)


if __name__ == "__main__":
    main()
