"""Synthetic workloads for benchmarks and smoke runs (no dataset, no checkpoint: there is no network).

`synthetic_batch` follows SURVEY.md section 8d: HR images ~ U(-1, 1); label maps are "blocky" (an
iid 16x16 grid of classes, nearest-upsampled to the crop size: piecewise-constant regions like a face
parse) or fully iid per pixel (the worst case for any label-uniformity shortcut); guided models get
an independent guiding image / label pair.  The dict has the reference dataloader's keys
(data/celebamaskhq_dataset.py: "label" int64 [B,1,S,S], "image" fp32 [B,3,S,S]).

`settle_spectral_norm` runs the power iteration of torch.nn.utils.spectral_norm on freshly
initialised weights until u / v have converged: a freshly constructed network has random u / v, so
its first forwards would divide by a sigma estimate near zero (hard part 1) and
overflow; training from a checkpoint never sees that state.
"""
import torch
import torch.nn.functional as F


def synthetic_batch(opt, batch, seed=1234, blocky=True):
    g = torch.Generator().manual_seed(seed)
    S, L = opt.crop_size, opt.label_nc

    def image():
        return torch.rand(batch, 3, S, S, generator=g) * 2 - 1

    def labels():
        if blocky:
            grid = torch.randint(0, L, (batch, 1, 16, 16), generator=g)
            return F.interpolate(grid.float(), size=(S, S), mode="nearest").long()
        return torch.randint(0, L, (batch, 1, S, S), generator=g)

    data = {"image": image(), "label": labels()}
    if getattr(opt, "guiding_style_image", False):
        data["guiding_image"] = image()
        data["guiding_label"] = labels()
    return data


@torch.no_grad()
def settle_spectral_norm(module, iters=30, eps=1e-12):
    """In-place power iteration on every spectral-normalised layer of `module` (weight_orig / weight_u /
    weight_v as registered by torch.nn.utils.spectral_norm)."""
    n = 0
    for m in module.modules():
        if hasattr(m, "weight_orig") and hasattr(m, "weight_u") and hasattr(m, "weight_v"):
            w = m.weight_orig.detach().flatten(1)
            u, v = m.weight_u.clone(), m.weight_v.clone()
            for _ in range(iters):
                v = F.normalize(torch.mv(w.t(), u), dim=0, eps=eps)
                u = F.normalize(torch.mv(w, v), dim=0, eps=eps)
            m.weight_u.copy_(u)
            m.weight_v.copy_(v)
            n += 1
    return n
