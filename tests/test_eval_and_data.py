"""The evaluation loop and data loader: metrics pinned to outputs of
the unmodified reference (tests/golden/metrics.npz, written by oracle/make_metrics_golden.py), the
folder dataset's item format (CPU), and - on the GPU - the prefetching loader and
InferenceManager.run over a small model."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _pair(seed, n, size):
    g = torch.Generator().manual_seed(seed)
    real = torch.rand(n, 3, size, size, generator=g) * 2 - 1
    real = torch.nn.functional.avg_pool2d(real, 5, 1, 2)
    real = real / real.abs().max()
    fake = (real + 0.08 * torch.randn(n, 3, size, size, generator=g)).clamp(-1.2, 1.2)
    return fake, real


def _metrics():
    # the metrics module is plain torch: importable without the CUDA library
    import importlib.util
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("dsee_metrics", os.path.join(here, "deepsee_b200", "evaluator",
                                                                                "metrics.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("tag", ["a", "b"])
def test_metrics_match_reference_golden(tag):
    m = _metrics()
    g = np.load(os.path.join(GOLD, "metrics.npz"))
    seed, n, size = [int(v) for v in g[tag + "_meta"]]
    fake, real = _pair(seed, n, size)
    assert np.abs(m.psnr(fake, real).numpy() - g[tag + "_psnr"]).max() < 1e-9       # same uint8 images
    assert np.abs(m.ssim(fake, real).numpy() - g[tag + "_ssim"]).max() < 1e-9
    assert np.abs(m.msssim(fake, real).numpy() - g[tag + "_msssim"]).max() < 1e-5   # fp32 filters
    assert np.abs(m.rmse(fake, real).numpy() - g[tag + "_rmse"]).max() < 1e-7


def _write_folder(tmp_path, n=5, size=40, L=19):
    from PIL import Image
    rng = np.random.RandomState(0)
    (tmp_path / "image").mkdir()
    (tmp_path / "label").mkdir()
    for i in range(n):
        Image.fromarray(rng.randint(0, 256, (size, size, 3), dtype=np.uint8)).save(str(tmp_path / "image" / ("%d.jpg" % i)))
        lab = rng.randint(0, L, (size, size), dtype=np.uint8)
        lab[0, 0] = 255                                   # 'unknown' -> label_nc
        Image.fromarray(lab, mode="L").save(str(tmp_path / "label" / ("%d.png" % i)))


def _opt(tmp_path, **kw):
    from deepsee_b200.options.configurations import make_opt
    return make_opt("8x_independent_256x256", ngf=8, nef=8, ndf=8, start_size=4, crop_size=32, load_size=32,
                    image_dir=str(tmp_path / "image"), label_dir=str(tmp_path / "label"), batchSize=2,
                    dataset_mode="celebamaskhq", serial_batches=True, nThreads=0, no_flip=True,
                    max_dataset_size=1 << 30, **kw)


def test_folder_dataset_item_format(tmp_path):
    from deepsee_b200.data import create_dataloader
    _write_folder(tmp_path)
    loader = create_dataloader(_opt(tmp_path, isTrain=False))
    batches = list(loader)
    assert len(batches) == 3 and [b["image"].shape[0] for b in batches] == [2, 2, 1]
    b = batches[0]
    assert tuple(b["image"].shape) == (2, 3, 32, 32) and tuple(b["label"].shape) == (2, 1, 32, 32)
    assert b["image"].dtype == torch.float32 and -1.0 <= float(b["image"].min()) and float(b["image"].max()) <= 1.0
    assert b["label"].dtype == torch.float32 and float(b["label"].max()) <= 19 and (b["label"] == b["label"].round()).all()
    assert (b["label"] == 19).any()                        # the 255 pixel became label_nc
    assert b["path"][0].endswith("0.jpg") and b["path"][1].endswith("1.jpg")   # natural sort, serial order


@pytest.mark.gpu
def test_prefetcher_and_inference_manager(tmp_path):
    from oracle import deepsee_oracle as O
    from deepsee_b200.data import create_dataloader, DevicePrefetcher
    from deepsee_b200.managers.base_manager import BaseManager
    from deepsee_b200.managers.inference_manager import InferenceManager
    _write_folder(tmp_path, n=5, size=40, L=19)
    opt = _opt(tmp_path, isTrain=False, checkpoints_dir=str(tmp_path))
    o = O.make_opt("8x_independent_256x256", ngf=8, nef=8, ndf=8, start_size=4, crop_size=32, load_size=32)
    ck = tmp_path / opt.name
    ck.mkdir()
    torch.save({"model": O.make_generator_state(o, 0)}, str(ck / "latest_net_SR.pth"))
    torch.save({"model": O.make_encoder_state(o, 1)}, str(ck / "latest_net_E.pth"))
    loader = create_dataloader(opt)
    host = list(loader)
    dev = list(DevicePrefetcher(loader))
    assert len(dev) == len(host)
    for h, d in zip(host, dev):
        assert d["image"].is_cuda and torch.equal(d["image"].cpu(), h["image"]) and d["path"] == h["path"]
    # labels with the 'unknown' class need contain_dontcare_label; keep the known classes here
    for d in dev:
        d["label"].clamp_(max=18)
    model = BaseManager(opt).sr_model
    mgr = InferenceManager(opt, num_samples=5, fid_features=lambda x: x.mean(dim=(2, 3)))
    res = mgr.run(model, dev)
    assert res["n_samples"] == 5
    for k in ("psnr/mean", "ssim/mean", "ms_ssim/mean", "rmse/mean", "psnr/std"):
        assert res[k] is not None and np.isfinite(res[k]), k
    assert res["lpips/mean"] is None and res["FID"] is not None and res["FID"] >= -1e-6
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()})
