"""GAN / feature-matching losses (reference: deepsee_models/networks/loss.py:19-101).

Scalar reductions over the discriminator's small output maps; SURVEY.md section 8 (a13) keeps
them in PyTorch ("negligible").  The VGG perceptual loss (loss.py:104-119) runs its VGG19 feature
stack on the deepsee_b200 conv kernels (networks/architecture.py:VGG19); its pretrained torchvision
weights cannot be downloaded here and must be supplied as a file (DSEE_VGG19_WEIGHTS).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class GANLoss(nn.Module):
    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0,
                 tensor=torch.FloatTensor, opt=None):
        super().__init__()
        if gan_mode not in ('ls', 'original', 'w', 'hinge'):
            raise ValueError('Unexpected gan_mode {}'.format(gan_mode))
        self.real_label = target_real_label
        self.fake_label = target_fake_label
        self.gan_mode = gan_mode
        self.opt = opt

    def loss(self, input, target_is_real, for_discriminator=True):
        if self.gan_mode == 'original':
            target = torch.full_like(input, self.real_label if target_is_real else self.fake_label)
            return F.binary_cross_entropy_with_logits(input, target)
        if self.gan_mode == 'ls':
            target = torch.full_like(input, self.real_label if target_is_real else self.fake_label)
            return F.mse_loss(input, target)
        if self.gan_mode == 'hinge':
            if for_discriminator:
                sgn = input if target_is_real else -input
                return -torch.mean(torch.min(sgn - 1, torch.zeros_like(input)))
            assert target_is_real, "The generator's hinge loss must be aiming for real"
            return -torch.mean(input)
        return -input.mean() if target_is_real else input.mean()

    def __call__(self, input, target_is_real, for_discriminator=True):
        # |input| is a list (scales) of lists (layers) for the multiscale discriminator (loss.py:87-101)
        if isinstance(input, list):
            loss = 0
            for pred_i in input:
                if isinstance(pred_i, list):
                    pred_i = pred_i[-1]
                loss_tensor = self.loss(pred_i, target_is_real, for_discriminator)
                bs = 1 if len(loss_tensor.size()) == 0 else loss_tensor.size(0)
                loss = loss + torch.mean(loss_tensor.view(bs, -1), dim=1)
            return loss / len(input)
        return self.loss(input, target_is_real, for_discriminator)


class VGGLoss(nn.Module):
    """Perceptual loss (loss.py:104-119): L1 distances between the VGG19 feature maps of the fake and
    the real image at relu1_1 ... relu5_1, weights 1/32, 1/16, 1/8, 1/4, 1.  Both images go through
    the feature stack as ONE batch of 2B (half the launches); only the fake half carries a gradient."""

    def __init__(self, gpu_ids, weights=None):
        super().__init__()
        from .architecture import VGG19
        self.vgg = VGG19(weights=weights)
        if gpu_ids or torch.cuda.is_available():
            self.vgg.cuda()
        self.criterion = nn.L1Loss()
        self.weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]

    def forward(self, x, y):
        B = x.size(0)
        feats = self.vgg(torch.cat([x, y.detach()], 0))
        loss = 0
        for w, f in zip(self.weights, feats):
            loss = loss + w * self.criterion(f[:B], f[B:].detach())
        return loss
