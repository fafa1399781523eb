"""InferenceManager (reference: managers/inference_manager.py:22-147): runs `num_samples` samples of a
dataloader through SRModel's 'inference' mode and aggregates PSNR / SSIM / MS-SSIM / RMSE (on the
device, ..evaluator.metrics) and, when a feature extractor is supplied, the Frechet distance of its
activations (the reference's FID uses a pretrained InceptionV3 that cannot be downloaded here)."""
import datetime
import os
import time
from collections import OrderedDict

import numpy as np
import torch

from ..evaluator.metrics import MetricsEvaluator
from .base_manager import BaseManager


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """evaluator/pytorch_fid/fid_score.py:calculate_frechet_distance."""
    from scipy import linalg
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


class InferenceManager(BaseManager):
    def __init__(self, opt, num_samples, write_details=False, folder_out=None, save_images=False, cuda=True,
                 fid_features=None, lpips_fn=None):
        super().__init__(opt, create_model=False)
        self.num_samples, self.batch_size = num_samples, opt.batchSize
        self.write, self.save_image, self.folder_out = write_details, save_images, folder_out
        if (self.save_image or self.write) and folder_out:
            os.makedirs(folder_out, exist_ok=True)
        self.metrics = MetricsEvaluator(write_details, folder_out, cuda=cuda, lpips_fn=lpips_fn)
        self.fid_features = fid_features   # callable: image batch in [-1, 1] -> [B, D] activations

    def run_batch(self, data, model):
        data = super().preprocess(data, from_dataloader=True)
        with torch.no_grad():
            return model(data, "inference")

    def run(self, model, dataloader):
        it = iter(dataloader)
        was_training = model.training
        model = model.eval()
        start = time.time()
        feats_fake, feats_real = [], []
        skipped = 0
        for _ in range(self.num_samples // self.batch_size + 1):
            try:
                data_i = next(it)
            except StopIteration:
                break
            try:
                paths = data_i.get('path')
                out = self.run_batch({k: v for k, v in data_i.items() if torch.is_tensor(v)}, model)
            except ValueError:
                skipped += 1
                continue
            fake, real = out['fake_image'].detach(), out['image_hr'].detach()
            self.metrics.collect_samples(fake, real, paths)
            if self.fid_features is not None:
                feats_fake.append(self.fid_features(fake).float().cpu().numpy())
                feats_real.append(self.fid_features(real).float().cpu().numpy())
        fid = None
        if feats_fake:
            af, ar = np.concatenate(feats_fake, 0), np.concatenate(feats_real, 0)
            fid = frechet_distance(af.mean(0), np.cov(af, rowvar=False), ar.mean(0), np.cov(ar, rowvar=False))
        result = OrderedDict([("FID", fid)])
        result.update(self.metrics.get_result())
        self.metrics.clear()
        model.train(was_training)
        print("Evaluation finished in %s; samples skipped: %d" %
              (datetime.timedelta(seconds=time.time() - start), skipped))
        return result
