"""Input preprocessing (reference: data/preprocessor.py:5-41)."""
import torch
import torch.nn.functional as F

from .. import ops
from .onehot import OneHotLabels


class Preprocessor:
    def __init__(self, opt):
        self.opt = opt

    def use_gpu(self):
        return len(self.opt.gpu_ids) > 0

    def downsample_image(self, hr_image, shape=None):
        """preprocessor.py:17-33: F.interpolate(mode=opt.downsampling_method) + clamp(-1, 1).
        Bicubic (the reference default, options/base_options.py) runs as one kernel with ATen's
        arithmetic (ops.bicubic_clamp: source index, A = -0.75 coefficients, clamped border taps,
        the clamp fused); other modes stay on torch's interpolate."""
        if shape is None:
            shape = (self.opt.start_size, self.opt.start_size)
        if (self.opt.downsampling_method == 'bicubic' and hr_image.is_cuda and
                hr_image.dtype == torch.float32 and hr_image.dim() == 4):
            return ops.bicubic_clamp(hr_image.contiguous(), tuple(shape))
        return F.interpolate(hr_image, shape, mode=self.opt.downsampling_method).clamp(min=-1, max=1)

    def preprocess_label(self, label_map):
        """preprocessor.py:35-41: the one-hot form of the integer label map.  Returned as an
        OneHotLabels: a [B, nc, H, W] float32 tensor to any caller (materialised bit-exactly on first
        use), a uint8 label map to the deepsee_b200 networks - the fp32 one-hot tensor never reaches
        HBM on the training path."""
        nc = self.opt.label_nc + 1 if self.opt.contain_dontcare_label else self.opt.label_nc
        if not label_map.is_cuda:
            raise RuntimeError('Preprocessor.preprocess_label (B200 path) needs a CUDA tensor')
        labels, bad = ops.labels_u8(label_map.long().contiguous(), nc)
        return OneHotLabels(labels, nc, bad)
