"""GAN / feature-matching losses (reference: deepsee_models/networks/loss.py:19-101).

Scalar reductions over the discriminator's small output maps; SURVEY.md section 8 (a13) keeps
them in PyTorch ("negligible").  VGG perceptual loss needs pretrained torchvision weights that
cannot be downloaded offline and is out of scope: constructing it raises.
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
    def __init__(self, gpu_ids):
        super().__init__()
        raise NotImplementedError(
            'VGGLoss needs pretrained torchvision VGG19 weights (network download) and is outside '
            'the B200 hot path; run with --no_vgg_loss')
