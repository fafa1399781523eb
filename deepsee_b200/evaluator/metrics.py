"""PSNR / SSIM / MS-SSIM / RMSE of (fake, real) image batches, on the device.

Reference: evaluator/evaluation.py:79-138 (MetricsEvaluator.collect_samples),
evaluator/calculate_PSNR_SSIM.py:70-120 (PSNR and the 11x11 Gaussian-window SSIM on uint8 images),
evaluator/ssim.py:24-118 (MS-SSIM on the [0, 255] float images).  The reference pulls every image
to the host, converts it to uint8 with numpy and runs OpenCV filters per sample; here the whole batch
stays on the GPU: the uint8 quantisation is emulated exactly (clip + truncation, util/util.py:97-103)
and the Gaussian filters are separable depthwise convolutions.  These are small scalar reductions on
[B,3,S,S] images - plain torch ops, like the losses; the kernels of the hot
path are not involved.  LPIPS and FID need pretrained networks that cannot be downloaded here: they
are reported as None unless a feature extractor is supplied.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


def to_uint8_values(x):
    """util.tensor2im (util/util.py:97-103) without leaving the device: [-1, 1] float ->
    the uint8 VALUES (as float64) numpy's clip + astype(uint8) produces."""
    # (fp32 arithmetic like numpy's on the fp32 image array, so the truncation falls the same way)
    return ((x.float() + 1) / 2.0 * 255.0).clamp(0, 255).floor().double()


def psnr(fake, real):
    """calculate_PSNR_SSIM.calculate_psnr on the uint8 images, per sample -> [B] float64."""
    a, b = to_uint8_values(fake), to_uint8_values(real)
    mse = ((a - b) ** 2).mean(dim=(1, 2, 3))
    return torch.where(mse == 0, torch.full_like(mse, float("inf")),
                       20 * torch.log10(255.0 / mse.sqrt().clamp_min(1e-300)))


def rmse(fake, real):
    """evaluation.py:104-106: root mean squared error on the [-1, 1] images, per sample."""
    return ((fake.float() - real.float()) ** 2).mean(dim=(1, 2, 3)).sqrt()


def _gauss1d(size, sigma, dtype, device):
    x = torch.arange(size, dtype=torch.float64, device=device) - size // 2
    g = torch.exp(-x ** 2 / (2 * sigma ** 2))
    return (g / g.sum()).to(dtype)


def _filter_valid(x, k1d):
    """'valid' 2-D Gaussian filtering of [B,C,H,W] as two depthwise 1-D passes."""
    C = x.shape[1]
    kh = k1d.view(1, 1, -1, 1).expand(C, 1, -1, 1)
    kw = k1d.view(1, 1, 1, -1).expand(C, 1, 1, -1)
    return F.conv2d(F.conv2d(x, kh, groups=C), kw, groups=C)


def _ssim_maps(a, b, k1d, C1, C2):
    mu1, mu2 = _filter_valid(a, k1d), _filter_valid(b, k1d)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _filter_valid(a * a, k1d) - mu1_sq
    s2 = _filter_valid(b * b, k1d) - mu2_sq
    s12 = _filter_valid(a * b, k1d) - mu12
    cs = (2 * s12 + C2) / (s1 + s2 + C2)
    return ((2 * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs, cs


def ssim(fake, real):
    """calculate_PSNR_SSIM.calculate_ssim on the uint8 images (11x11 Gaussian, sigma 1.5, 'valid'
    region, C1 = (0.01*255)^2, C2 = (0.03*255)^2), per sample -> [B] float64.  (The reference's RGB
    branch evaluates `ssim(img1, img2)` on the whole H x W x 3 array three times and averages -
    calculate_PSNR_SSIM.py:108-112: cv2.filter2D filters each channel, so that equals the mean over
    all three channels' maps, which is what is computed here.)"""
    a, b = to_uint8_values(fake), to_uint8_values(real)
    k = _gauss1d(11, 1.5, torch.float64, a.device)
    m, _ = _ssim_maps(a, b, k, (0.01 * 255) ** 2, (0.03 * 255) ** 2)
    return m.mean(dim=(1, 2, 3))


def msssim(fake, real, window_size=11):
    """evaluator/ssim.py:90-118 with val_range=255 on (x + 1) * 127.5 (evaluation.py:111,124-126):
    five scales, fp32, avg-pool 2x2 between them, `prod(mcs[:-1] ** w[:-1] * mssim[-1] ** w[-1])`
    exactly as the reference combines them (per channel means first, ssim.py:44-80).  Per sample."""
    out = []
    w = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=torch.float32, device=fake.device)
    for i in range(fake.size(0)):
        a, b = (fake[i:i + 1].float() + 1.0) * 127.5, (real[i:i + 1].float() + 1.0) * 127.5
        sims, css = [], []
        for _ in range(5):
            size = min(window_size, a.shape[2], a.shape[3])
            k = _gauss1d(size, 1.5, torch.float32, a.device)
            # the reference builds the 2-D window in double and casts it to float (ssim.py:16-21)
            k2 = torch.outer(_gauss1d(size, 1.5, torch.float64, a.device),
                             _gauss1d(size, 1.5, torch.float64, a.device)).float()
            del k
            win = k2.view(1, 1, size, size).expand(a.shape[1], 1, size, size)
            f = lambda t: F.conv2d(t, win, groups=t.shape[1])  # noqa: E731
            mu1, mu2 = f(a), f(b)
            mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
            s1, s2, s12 = f(a * a) - mu1_sq, f(b * b) - mu2_sq, f(a * b) - mu12
            C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
            v1, v2 = 2.0 * s12 + C2, s1 + s2 + C2
            smap = ((2 * mu12 + C1) * v1) / ((mu1_sq + mu2_sq + C1) * v2)
            # per channel means, then the mean over channels (ssim.py:44-80)
            sims.append(smap.mean(dim=(0, 2, 3)).mean())
            css.append((v1 / v2).mean(dim=(0, 2, 3)).mean())
            a, b = F.avg_pool2d(a, (2, 2)), F.avg_pool2d(b, (2, 2))
        mssim, mcs = torch.stack(sims), torch.stack(css)
        out.append(torch.prod((mcs ** w)[:-1] * (mssim ** w)[-1]))
    return torch.stack(out)


class MetricsEvaluator:
    """evaluation.py:15-160: accumulates per-sample scores; `get_result` returns the same keys."""
    columns = ["ID", "PSNR", "SSIM", "MSSSIM", "RMSE", "LPIPS"]

    def __init__(self, write_details=False, folder_out=None, cuda=True, lpips_fn=None):
        self.lpips_fn = lpips_fn
        self.write_details, self.folder_out = write_details, folder_out
        self.clear()

    def clear(self):
        self.buf = {k: [] for k in ("psnr", "ssim", "ms_ssim", "rmse", "lpips")}
        self.ids = []
        self.n_samples = 0

    def collect_samples(self, fake, real, name=None):
        assert fake.size(0) == real.size(0)
        fake, real = fake.detach(), real.detach()
        self.buf["psnr"] += psnr(fake, real).tolist()
        self.buf["ssim"] += ssim(fake, real).tolist()
        self.buf["ms_ssim"] += msssim(fake, real).tolist()
        self.buf["rmse"] += rmse(fake, real).tolist()
        if self.lpips_fn is not None:
            self.buf["lpips"] += [float(v) for v in self.lpips_fn(fake, real).flatten()]
        if name is not None:
            self.ids += [str(n) for n in name]
        self.n_samples += fake.size(0)

    def get_result(self):
        out = OrderedDict()
        for stat, fn in (("mean", np.mean), ("std", np.std)):
            for k in ("psnr", "ssim", "ms_ssim", "rmse", "lpips"):
                out["%s/%s" % (k, stat)] = float(fn(self.buf[k])) if self.buf[k] else None
        out["n_samples"] = self.n_samples
        if self.write_details and self.folder_out:
            import csv
            import os
            os.makedirs(self.folder_out, exist_ok=True)
            with open(os.path.join(self.folder_out, "metrics.csv"), "w", newline="") as f:
                wr = csv.writer(f)
                wr.writerow(self.columns)
                for i in range(self.n_samples):
                    wr.writerow([self.ids[i] if i < len(self.ids) else i] +
                                [self.buf[k][i] if i < len(self.buf[k]) else "" for k in
                                 ("psnr", "ssim", "ms_ssim", "rmse", "lpips")])
        return out
