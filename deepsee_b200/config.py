"""Runtime knobs of the B200 path (process-wide)."""
import os


class _Config:
    # Tensor-core operand passes of the fused kernels:
    #   3 = hi*hi + lo*hi + hi*lo  (fp32-class accuracy; the parity default)
    #   1 = hi*hi                  (fp16 operands = TF32-class accuracy, 3x fewer MMAs)
    passes = int(os.environ.get("DSEE_PASSES", "3"))
    # Verify (one device->host read per generator forward) that the semantic input is one-hot.
    check_onehot = os.environ.get("DSEE_CHECK_ONEHOT", "1") != "0"


config = _Config()
