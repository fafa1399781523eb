"""Input preprocessing (reference: data/preprocessor.py:5-41)."""
import torch.nn.functional as F

from .. import ops


class Preprocessor:
    def __init__(self, opt):
        self.opt = opt

    def use_gpu(self):
        return len(self.opt.gpu_ids) > 0

    def downsample_image(self, hr_image, shape=None):
        """preprocessor.py:17-33: F.interpolate(mode=opt.downsampling_method) + clamp(-1, 1).
        One small resampling per batch on the input side of the path; kept on torch's own
        interpolate so the LR image is bit-identical to what the reference feeds its generator
        (SURVEY.md section 8f rank 1 lists a fused version as a follow-up)."""
        if shape is None:
            shape = (self.opt.start_size, self.opt.start_size)
        return F.interpolate(hr_image, shape, mode=self.opt.downsampling_method).clamp(min=-1, max=1)

    def preprocess_label(self, label_map):
        """preprocessor.py:35-41: one-hot scatter of the integer label map (bit-exact)."""
        nc = self.opt.label_nc + 1 if self.opt.contain_dontcare_label else self.opt.label_nc
        if not label_map.is_cuda:
            raise RuntimeError('Preprocessor.preprocess_label (B200 path) needs a CUDA tensor')
        onehot, bad = ops.onehot_from_labels(label_map.long().contiguous(), nc)
        return onehot
