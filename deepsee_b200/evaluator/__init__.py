"""Evaluation metrics of the reference's InferenceManager (evaluator/evaluation.py) on the GPU."""
from .metrics import MetricsEvaluator, psnr, rmse, ssim, msssim  # noqa: F401
