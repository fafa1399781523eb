"""Writes tests/golden/metrics.npz: PSNR / SSIM / MS-SSIM / RMSE of seeded image pairs computed by the
UNMODIFIED reference (evaluator/calculate_PSNR_SSIM.py, evaluator/ssim.py, util/util.py:tensor2im;
the per-sample recipe of evaluator/evaluation.py:104-131).  Run once in the build container:

    PYTHONPATH=/root/reference python oracle/make_metrics_golden.py

TEST INFRASTRUCTURE - the product never imports this."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("DEEPSEE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from evaluator.calculate_PSNR_SSIM import calculate_psnr, calculate_ssim  # noqa: E402
from evaluator.ssim import msssim  # noqa: E402
from util.util import tensor2im  # noqa: E402


def pair(seed, n, size):
    g = torch.Generator().manual_seed(seed)
    real = torch.rand(n, 3, size, size, generator=g) * 2 - 1
    # a smooth image plus noise, so SSIM is neither 0 nor 1
    real = torch.nn.functional.avg_pool2d(real, 5, 1, 2)
    real = real / real.abs().max()
    fake = (real + 0.08 * torch.randn(n, 3, size, size, generator=g)).clamp(-1.2, 1.2)
    return fake, real


def main():
    out = {}
    for tag, seed, n, size in (("a", 1, 3, 96), ("b", 2, 2, 200)):
        fake, real = pair(seed, n, size)
        fnp, rnp = tensor2im(fake), tensor2im(real)
        f255, r255 = (fake + 1.0) * 127.5, (real + 1.0) * 127.5
        out[tag + "_psnr"] = np.array([calculate_psnr(fnp[i], rnp[i]) for i in range(n)])
        out[tag + "_ssim"] = np.array([calculate_ssim(fnp[i], rnp[i]) for i in range(n)])
        out[tag + "_msssim"] = np.array([float(msssim(f255[i:i + 1], r255[i:i + 1], size_average=True,
                                                       val_range=255)) for i in range(n)])
        out[tag + "_rmse"] = torch.nn.MSELoss(reduction="none")(fake, real).mean(dim=[1, 2, 3]).sqrt().numpy()
        out[tag + "_meta"] = np.array([seed, n, size])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "metrics.npz")
    np.savez_compressed(path, **out)
    print({k: v for k, v in out.items()})


if __name__ == "__main__":
    main()
