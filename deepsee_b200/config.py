"""Runtime knobs of the B200 path (process-wide)."""
import os


class _Config:
    # Tensor-core operand passes of the fused kernels:
    #   3 = hi*hi + lo*hi + hi*lo  (fp32-class accuracy; the parity default)
    #   1 = hi*hi                  (fp16 operands = TF32-class accuracy, 3x fewer MMAs)
    passes = int(os.environ.get("DSEE_PASSES", "3"))
    # Training: K1 saves G = gamma + gamma_bias (fp16 planes, +1-2 B per activation element) so its
    # backward is one streaming pass instead of re-running the gamma GEMM (0 = recompute).
    save_gamma = os.environ.get("DSEE_SAVE_GAMMA", "1") != "0"
    # Backward: run the backward-data GEMM of a main conv and K1's backward of the norm layer in front
    # of it as one kernel (needs save_gamma); 0 = dgrad -> dt in HBM -> streaming K1 backward.
    fuse_dgrad_modbwd = os.environ.get("DSEE_FUSE_DGRAD_MODBWD", "1") != "0"
    # Issue the weight-gradient GEMMs of the generator on a second stream (they are leaves of the
    # backward graph) so the HBM-bound kernels of the chain overlap with them.
    # (measured gain on B200: ~1 %; off by default so per-kernel CUDA-event timings stay clean)
    overlap_wgrad = os.environ.get("DSEE_OVERLAP_WGRAD", "0") == "1"
    # NoiseInjection: 0 = in-kernel counter-based noise (never in HBM), 1 = torch.randn tensors.
    noise_tensors = os.environ.get("DSEE_NOISE_TENSORS", "0") == "1"
    # Debug switches: run torch's own spectral-norm hook / build the modulation weight from torch ops
    # instead of the fused kernels (ops.SpectralWeightFn / ops.ModWeightFn).
    torch_spectral = os.environ.get("DSEE_TORCH_SPECTRAL", "0") == "1"
    torch_modweight = os.environ.get("DSEE_TORCH_MODWEIGHT", "0") == "1"
    # Multi-GPU batch-norm statistics: 0 = per-rank statistics (north_star: NCCL all-reduce "for G/D
    # gradients only"), 1 = global-batch statistics like the reference's DataParallel mode
    # (sync_batchnorm/batchnorm.py:63-93): one [2,C] all-reduce per norm layer forward and one
    # [2,C] all-reduce per norm layer backward.
    sync_bn = os.environ.get("DSEE_SYNC_BN", "0") == "1"
    # Spectral normalisation of a network's layers in one batched launch sequence per forward
    # (ops.spectral_prepass) instead of five launches per layer; 0 = per-layer kernels.
    batched_spectral = os.environ.get("DSEE_BATCHED_SPECTRAL", "1") != "0"
    # Verify (one device->host read per generator forward) that the semantic input is one-hot.
    check_onehot = os.environ.get("DSEE_CHECK_ONEHOT", "1") != "0"


config = _Config()
