"""One-hot semantic maps that stay a uint8 label map until someone needs the dense tensor.

The reference materialises `input_semantics` as a fp32 one-hot tensor [B, L, S, S]
(data/preprocessor.py:35-41: 5 MB per 256x256 image, 20 MB per 512x512 image) and every network then
multiplies with it.  The B200 kernels consume the uint8 label map instead.  `OneHotLabels` is what
`Preprocessor.preprocess_label` returns: a torch.Tensor subclass that *is* a [B, L, H, W] float32
tensor to every caller (shape / dtype / device are answered from metadata; any torch operation on it
materialises the dense one-hot once, with the bit-exact scatter kernel, and proceeds on that), while
the deepsee_b200 networks take `.labels` and never touch the dense form.
"""
import torch
from torch.utils._pytree import tree_map

from .. import ops


class OneHotLabels(torch.Tensor):
    @staticmethod
    def __new__(cls, labels, num_classes, bad=None):
        assert labels.dtype == torch.uint8 and labels.dim() == 3
        B, H, W = labels.shape
        r = torch.Tensor._make_wrapper_subclass(cls, (B, num_classes, H, W), dtype=torch.float32,
                                                device=labels.device, requires_grad=False)
        r.labels = labels            # uint8 [B, H, W]
        r.num_classes = num_classes
        r.bad = bad                  # device int32 flag: 1 if a label was out of range, or None
        r._dense = None
        return r

    def dense(self):
        """The fp32 one-hot tensor [B, L, H, W] (materialised on first use)."""
        if self._dense is None:
            self._dense, _ = ops.onehot_from_labels(self.labels.long().unsqueeze(1).contiguous(),
                                                    self.num_classes)
        return self._dense

    def __repr__(self):
        return "OneHotLabels(shape=%s, device=%s, dense=%s)" % (tuple(self.shape), self.device,
                                                                self._dense is not None)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        unwrap = lambda t: t.dense() if isinstance(t, OneHotLabels) else t  # noqa: E731
        return func(*tree_map(unwrap, args), **tree_map(unwrap, kwargs or {}))


def labels_of(seg):
    """-> (uint8 label map [B, H, W], device flag or None) of a semantic input: free for an
    OneHotLabels, one pass over the dense tensor (with its one-hot check) otherwise."""
    if isinstance(seg, OneHotLabels):
        return seg.labels, seg.bad
    return ops.labels_from_onehot(seg.contiguous().float())
