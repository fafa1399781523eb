"""Worker of tests/test_ddp_gpu.py: one training iteration of a small model on this rank's shard of a
fixed global batch (run under torch.distributed.run, one process per GPU, NCCL), or on the whole
batch when launched as a single process.  Writes losses and a few gradients to --out (rank 0)."""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--global-batch", type=int, default=4)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from oracle import deepsee_oracle as O  # seeded synthetic inputs / weights only
    from deepsee_b200 import parallel
    from deepsee_b200.managers.trainer_manager import TrainerManager
    from test_generator_gpu import _mk_opt

    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    parallel.init_from_env()
    world, rank = parallel.world_size(), parallel.rank()
    o = O.make_opt("8x_independent_256x256", is_train=True, ngf=8, nef=8, ndf=8, start_size=8,
                   crop_size=64, load_size=64, add_noise=False, noisy_style_scale=0.0)
    mgr = TrainerManager(_mk_opt(o))
    m = mgr.sr_model
    m.netSR.load_state_dict(O.make_generator_state(o, 0), strict=True)
    m.netE.load_state_dict(O.make_encoder_state(o, 1), strict=True)
    m.netD.load_state_dict(O.make_discriminator_state(o, 2), strict=True)
    m.train()
    random.seed(0)
    raw = O.synthetic_batch(o, args.global_batch, seed=11)
    per = args.global_batch // world
    shard = {k: v[rank * per:(rank + 1) * per].clone() for k, v in raw.items()}
    shard["label"] = shard["label"].float()
    mgr.run_generator_one_step(dict(shard))
    res = {}
    for k, v in mgr.get_latest_losses().items():
        t = v.detach().mean().reshape(1).clone()
        if world > 1:
            dist.all_reduce(t)
            t /= world
        res["loss_" + k] = t.cpu()
    for name, p in list(m.netSR.named_parameters()) + [("E." + n, q) for n, q in m.netE.named_parameters()]:
        if p.grad is not None and any(s in name for s in ("head_0.conv_0.weight_orig", "up_list.1.conv_1.bias",
                                                          "G_middle_0.norm_0.mlp_gamma.weight",
                                                          "initial.weight", "E.final.0.weight_orig",
                                                          "up_list.0.norm_1.mlp_shared.0.weight")):
            res["grad_" + name] = p.grad.detach().cpu().clone()
    res["running_mean"] = m.netSR.state_dict()["G_middle_1.norm_1.param_free_norm.running_mean"].cpu().clone()
    mgr.run_discriminator_one_step(dict(shard))
    for k, v in mgr.d_losses.items():
        t = v.detach().mean().reshape(1).clone()
        if world > 1:
            dist.all_reduce(t)
            t /= world
        res["loss_" + k] = t.cpu()
    res["gradD"] = next(p for n, p in m.netD.named_parameters() if p.grad is not None).grad.detach().cpu().clone()
    res["peer_exchange"] = torch.tensor([1.0 if parallel.peer_exchange_active() else 0.0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        torch.save(res, args.out)


if __name__ == "__main__":
    main()
