"""GPU: the input side of the path - the uint8 label map end to end
(`OneHotLabels`: a lazily materialised one-hot tensor) and the bicubic low-resolution image kernel -
against the reference's preprocessing restated in the oracle (data/preprocessor.py:17-41)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import deepsee_oracle as O
from test_generator_gpu import _mk_opt

pytestmark = pytest.mark.gpu


def test_onehot_labels_is_a_bit_exact_lazy_onehot():
    from deepsee_b200.data.onehot import OneHotLabels, labels_of
    from deepsee_b200.data.preprocessor import Preprocessor
    o = O.make_opt("8x_independent_256x256", ngf=8, start_size=8, crop_size=64, load_size=64)
    raw = O.synthetic_batch(o, 3, seed=9)
    pre = Preprocessor(_mk_opt(o))
    seg = pre.preprocess_label(raw["label"].cuda())
    assert isinstance(seg, OneHotLabels) and seg._dense is None
    # metadata without materialising
    assert tuple(seg.shape) == (3, 19, 64, 64) and seg.size(1) == 19 and seg.dtype == torch.float32 and seg.is_cuda
    labels, bad = labels_of(seg)
    assert seg._dense is None and labels.dtype == torch.uint8
    assert torch.equal(labels.cpu(), raw["label"][:, 0].to(torch.uint8)) and int(bad.item()) == 0
    # any torch op sees the reference's one-hot tensor, bit for bit (integer indexing work)
    ref = torch.from_numpy(O.preprocess_label_np(raw["label"].numpy().astype(np.int64), 19))
    assert torch.equal(seg.cpu(), ref)
    assert seg._dense is not None
    assert torch.equal((seg[1:2] * 2.0).cpu(), ref[1:2] * 2.0)
    assert torch.equal(torch.cat([seg, seg], 0).sum((0, 2, 3)).cpu(), 2 * ref.sum((0, 2, 3)))
    # out-of-range labels are flagged (the reference's scatter_ raises)
    lab = raw["label"].clone()
    lab[0, 0, 0, 0] = 19
    assert int(labels_of(pre.preprocess_label(lab.cuda()))[1].item()) == 1


def test_generator_takes_onehot_labels_like_a_dense_onehot():
    from deepsee_b200.data.preprocessor import Preprocessor
    from test_generator_gpu import _build_G
    o = O.make_opt("8x_independent_256x256", ngf=8, start_size=8, crop_size=64, load_size=64)
    G = _build_G(o, O.make_generator_state(o, 0)).eval()
    raw = O.synthetic_batch(o, 2, seed=10)
    seg = Preprocessor(_mk_opt(o)).preprocess_label(raw["label"].cuda())
    d = O.preprocess(o, raw)
    z = torch.rand(2, 19, 128, generator=torch.Generator().manual_seed(1)).cuda() * 2 - 1
    with torch.no_grad():
        a = G(d["image_lr"].cuda(), seg=seg, z=z)
        assert seg._dense is None, "the generator materialised the one-hot tensor"
        b = G(d["image_lr"].cuda(), seg=d["input_semantics"].cuda(), z=z)
    assert torch.equal(a, b)


@pytest.mark.parametrize("S,s", [(256, 32), (512, 16), (64, 8), (100, 24)])
def test_bicubic_clamp_matches_interpolate(S, s):
    from deepsee_b200 import ops
    g = torch.Generator().manual_seed(S + s)
    hr = (torch.rand(2, 3, S, S, generator=g) * 2.4 - 1.2).cuda()   # overshoots so the clamp matters
    ref = F.interpolate(hr, (s, s), mode="bicubic").clamp(-1, 1)
    out = ops.bicubic_clamp(hr, (s, s))
    err = (out - ref).abs().max().item()
    exact = (out == ref).float().mean().item()
    print("bicubic %d -> %d: max-abs vs F.interpolate %.2e, bit-identical elements %.1f %%" % (S, s, err, 100 * exact))
    assert err <= 1e-6     # same taps and coefficients; fp32 contraction order may differ by an ulp
    # the CPU oracle's LR image (what the reference feeds its generator) agrees to the same tolerance
    assert (out.cpu() - O.downsample_image(hr.cpu(), s)).abs().max().item() <= 2e-6
