"""Runtime knobs of the B200 path (process-wide)."""
import os


class _Config:
    # Tensor-core operand passes of the fused kernels:
    #   3 = hi*hi + lo*hi + hi*lo  (fp32-class accuracy; the parity default)
    #   1 = hi*hi                  (fp16 operands = TF32-class accuracy, 3x fewer MMAs)
    passes = int(os.environ.get("DSEE_PASSES", "3"))
    # Mixed precision (only matters when passes == 1).  Measured on the full-size generators
    # (profiles/r2_precision_probe_v{1,2}.json, train mode, vs the CPU oracle): pure 1-pass (fp16
    # operands = TF32-class) lands at 1.6e-3 ... 2.4e-3 max-abs on the tanh output - over north_star's
    # 1e-3 - and the error comes overwhelmingly from the MAIN convs of the FORWARD pass (K2); the
    # gamma/beta GEMMs (K1) matter only in the low-resolution stages, whose rounding errors are
    # upsampled into spatially coherent patches that every later block and the 3x3 image head sum up.
    # So in 1-pass mode:
    #   * forward main convs run `k2_fwd_passes` (2 = one fp16 pass + an fp8 correction GEMM for both
    #     rounding terms, at the cost of two fp16 passes; 3 = the full hi/lo split);
    #   * K1 runs 3 passes in stages up to `passes3_upto` pixels high ("auto" = a quarter of the output
    #     size: ~5 % of the FLOPs), 1 pass above;
    #   * every backward GEMM (dgrad, wgrad) runs 1 pass: gradients carry no 1e-3 forward bound.
    #   * style encoder convs (their output feeds the generator): forward `ed_fwd_passes` (3), backward
    #     1 pass; discriminator convs (they only feed the losses): `d_fwd_passes` (1).
    k2_fwd_passes = int(os.environ.get("DSEE_K2_FWD_PASSES", "2"))
    ed_fwd_passes = int(os.environ.get("DSEE_ED_FWD_PASSES", "3"))
    d_fwd_passes = int(os.environ.get("DSEE_D_FWD_PASSES", "1"))
    passes3_upto = os.environ.get("DSEE_PASSES3_UPTO", "auto")
    # test / probe hook: {("k1" | "k2" | "k2b", H): passes} overrides of the rule above
    pass_overrides = {}

    def passes_for(self, kind, H, S=None):
        """Tensor-core operand passes of one generator kernel at feature-map height H of a generator
        whose output is S pixels high.  kind: "k1" = the gamma/beta GEMM of a conditional norm layer,
        forward; "k1b" = the backward GEMMs that share its operands; "k2" = a main 3x3 conv, forward;
        "k2b" = the backward GEMMs of a main conv (backward-data, weight gradient)."""
        o = self.pass_overrides.get((kind, H))
        if o is not None:
            return o
        if self.passes == 3:
            return 3
        if kind == "k2":
            return self.k2_fwd_passes
        if kind == "k1":  # forward only; "k1b" (its backward GEMMs) falls through to self.passes
            upto = self.passes3_upto
            if upto == "auto":
                upto = (S // 4) if S else 0
            return 3 if H <= int(upto) else self.passes
        return self.passes

    def precision_name(self):
        if self.passes == 3:
            return "fp16 hi+lo split operands x3 passes, fp32 accumulate (fp32-class)"
        k2 = {1: "1 pass", 2: "1 fp16 pass + fp8 correction of both operand-rounding terms",
              3: "hi+lo split operands x3 passes"}[self.k2_fwd_passes]
        return ("fp16 operands, fp32 accumulate (TF32-class) for the gamma/beta GEMMs and all backward GEMMs; "
                "forward main convs: %s; forward gamma/beta GEMMs of stages <= %s and forward style-encoder "
                "convs: x3 passes" %
                (k2, "1/4 of the output size" if self.passes3_upto == "auto" else "%s px" % self.passes3_upto))

    # Training: capture each optimizer sub-step (forward, backward, gradient all-reduce, Adam) of
    # TrainerManager as a CUDA graph per encoder coin-flip variant and replay it (managers/
    # trainer_manager.py).  Removes the host from the step: the small-kernel phases (style encoder,
    # discriminator, parameter-side kernels) are launch-bound otherwise.  Off in the library (eager
    # semantics, `.grad` readable after a step); bench.py turns it on.
    cuda_graphs = os.environ.get("DSEE_CUDA_GRAPHS", "0") == "1"
    # DemoManager.run: replay the batch-1 generator forward as a CUDA graph (per input shape).
    demo_graphs = os.environ.get("DSEE_DEMO_GRAPHS", "1") != "0"
    # SEAN layers: fold the style branch into per-image modulation weights over the exact one-hot
    # label planes (normalization.py:182-185,198-201: conv(style_map, W) = conv(onehot, W x style_b)),
    # so K1's K per tap drops from 256 to 192 channels, the backward-data GEMM of the modulation
    # halves (no gradient flows into a one-hot plane) and the gathered style_map is never built.
    # Needs save_gamma.  0 = the gathered 128-channel style_map.
    fold_style = os.environ.get("DSEE_FOLD_STYLE", "1") != "0"
    # Conditional-norm layers above max_fm_size convolve the nearest-2x-upsampled mlp_shared activation
    # (normalization.py:188-190,275-277: the 512x512 PureSEAN block of the 32x models).  Sub-pixel
    # form: keep the activation at half resolution and run four 2x2-tap GEMMs, one per output parity
    # class (ops.collapse_subpixel): 4/9 of the FLOPs forward and backward, and the upsampled tensor is
    # never built.  Needs save_gamma.  0 = materialise the upsampled activation.
    subpixel = os.environ.get("DSEE_SUBPIXEL", "1") != "0"
    # Training: K1 saves G = gamma + gamma_bias (fp16 planes, +1-2 B per activation element) so its
    # backward is one streaming pass instead of re-running the gamma GEMM (0 = recompute).
    # torch.optim.Adam(fused=True): the whole update of a parameter group in one multi-tensor kernel
    fused_adam = os.environ.get("DSEE_FUSED_ADAM", "1") != "0"
    # the image head (leaky_relu -> conv_img -> tanh, sr.py:94-95) as a 1x1 tensor-core GEMM over fp16
    # planes written by the last main conv's epilogue + a 9-tap shift-add; 0 = the fp32 CUDA-core kernels
    head_tc = os.environ.get("DSEE_HEAD_TC", "1") != "0"
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
    # Multi-GPU batch-norm statistics.  "auto" (default): global-batch statistics whenever the
    # generator's norm type says so (`norm_G` contains "syncbatch", the reference default) and more
    # than one rank runs - the reference's multi-GPU semantics (sync_batchnorm/batchnorm.py:63-93):
    # one [2,C] all-reduce per norm layer forward and one per norm layer backward; running
    # statistics are then identical on every rank.  DSEE_SYNC_BN=0: per-rank statistics (gradient
    # all-reduce is the only collective; running stats diverge across ranks, rank 0's are saved);
    # DSEE_SYNC_BN=1: always synchronise.
    sync_bn = {"0": False, "1": True}.get(os.environ.get("DSEE_SYNC_BN", "auto"), "auto")

    # Gradient buckets: launch each chunk's all-reduce from the backward pass as soon as its last
    # gradient has landed (1), or all chunks after the backward pass (0).
    grad_overlap = os.environ.get("DSEE_GRAD_OVERLAP", "1") != "0"
    # Sync-BN statistics over NVLink peer memory (one library kernel per exchange, parallel._PeerExchange)
    # instead of an NCCL all-reduce per norm layer; 0 = NCCL.
    peer_sync_bn = os.environ.get("DSEE_PEER_SYNC_BN", "1") != "0"

    def sync_bn_for(self, norm_G):
        """Whether the conditional-norm layers of a generator with this norm_G string exchange their
        batch statistics across ranks (the caller still checks that a process group exists)."""
        if self.sync_bn == "auto":
            return "syncbatch" in (norm_G or "")
        return bool(self.sync_bn)
    # Spectral normalisation of a network's layers in one batched launch sequence per forward
    # (ops.spectral_prepass) instead of five launches per layer; 0 = per-layer kernels.
    batched_spectral = os.environ.get("DSEE_BATCHED_SPECTRAL", "1") != "0"
    # Verify (one device->host read per generator forward) that the semantic input is one-hot.
    check_onehot = os.environ.get("DSEE_CHECK_ONEHOT", "1") != "0"


config = _Config()
