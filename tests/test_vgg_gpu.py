"""GPU: the VGG19 perceptual loss (loss.py:104-119, architecture.py:151-181) on the deepsee_b200
conv kernels against the oracle's torch statement of the same network, with a seeded stand-in for
the pretrained weights (the real checkpoint cannot be downloaded here; the architecture, slicing,
loss weights and gradient path are what is checked)."""
import pytest
import torch

from oracle import deepsee_oracle as O

pytestmark = pytest.mark.gpu


def test_vgg19_features_loss_and_gradient_vs_oracle():
    from deepsee_b200.deepsee_models.networks.loss import VGGLoss
    sd = O.make_vgg19_state(3)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).requires_grad_(True)
    y = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    ref_feats = O.vgg19_features(sd, x)
    ref_loss = O.vgg_loss(sd, x, y)
    ref_loss.backward()

    crit = VGGLoss([0], weights=sd)          # torchvision-style keys load into the sliced module tree
    assert not any(p.requires_grad for p in crit.parameters())
    assert set(crit.vgg.state_dict()) >= {"slice1.0.weight", "slice2.5.bias", "slice5.28.weight"}
    xg = x.detach().clone().cuda().requires_grad_(True)
    feats = crit.vgg(xg)
    assert [tuple(f.shape) for f in feats] == [tuple(f.shape) for f in ref_feats]
    for i, (a, b) in enumerate(zip(feats, ref_feats)):
        err = (a.detach().cpu() - b.detach()).abs().max().item() / b.abs().max().item()
        print("relu%d_1 max-abs / max|ref| %.2e" % (i + 1, err))
        # library default (fp32-class split operands): ~1e-5 per conv, accumulated over up to 13 layers
        assert err < 5e-4
    loss = crit(xg, y.cuda())
    loss.backward()
    print("VGG loss ours %.6f oracle %.6f" % (float(loss.detach()), float(ref_loss.detach())))
    # a mean of |feature differences|: inherits the features' ~1e-4 relative accuracy
    assert abs(float(loss.detach()) - float(ref_loss.detach())) < 2e-4 * abs(float(ref_loss.detach()))
    gerr = (xg.grad.cpu() - x.grad).abs().max().item() / x.grad.abs().max().item()
    print("d loss / d fake image: max-abs / max|ref| %.2e" % gerr)
    assert gerr < 5e-3    # ReLU / max-pool / L1 kinks: a 1e-6 forward difference flips isolated elements


def test_vgg_loss_without_weights_fails_loudly(monkeypatch):
    from deepsee_b200.deepsee_models.networks.loss import VGGLoss
    monkeypatch.delenv("DSEE_VGG19_WEIGHTS", raising=False)
    monkeypatch.setattr(torch.hub, "get_dir", lambda: "/nonexistent")
    with pytest.raises(RuntimeError, match="no pretrained weights"):
        VGGLoss([0])
