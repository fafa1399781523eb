"""SRModel facade (reference: deepsee_models/sr_model.py).

Same constructor, ``forward(data, mode)`` dispatch, return structures, optimizer setup and
checkpoint naming as the reference, so managers / train.py / demo.py drive it unchanged; the
networks underneath run on the deepsee_b200 CUDA kernels.  Differences, all deliberate:
  * one process per GPU: the model lives on the current CUDA device, `model_parallel_mode` and
    DataParallel are gone (data parallelism = one SRModel per rank + NCCL gradient all-reduce,
    see managers/base_manager.py);
  * the [seg | image] / fake | real concatenations of `discriminate` are one fused kernel;
  * the demo-time style-manipulation modes (sr_model.py:116-444) batch all their variants into one
    generator call instead of one batch-1 call per variant (`_style_sweep`);
    'inference_replace_semantics' is broken in the reference (calls a method that does not exist)
    and raises NotImplementedError here.
"""
import random
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import networks
from .. import ops
from ..util import util


_SWEEP_MODES = ("inference_multi_modal", "inference_reference_semantics", "inference_interpolation",
                "inference_interpolation_style", "inference_reference", "inference_reference_interpolation")
_CONSISTENT_REGIONS = [4, 6, 8, 11]   # left/right pairs kept consistent (sr_model.py:134,314)


class SRModel(torch.nn.Module):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        networks.modify_commandline_options(parser, is_train)
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.use_E = opt.netE is not None and len(opt.netE)
        self.netSR, self.netD, self.netE = self.initialize_networks(opt)
        self.mp = 0
        self.model_variant = "guided" if "full" in self.opt.netE else "independent"
        if opt.isTrain:
            self.criterionGAN = networks.GANLoss(opt.gan_mode, opt=self.opt)
            self.criterionFeat = torch.nn.L1Loss()
            if not opt.no_vgg_loss:
                self.criterionVGG = networks.VGGLoss(self.opt.gpu_ids)
        self.logs = OrderedDict()
        self.last_encoded_style_is_full = True
        self.last_encoded_style_is_noisy = False

    def load_weights(self):
        opt = self.opt
        if not opt.isTrain or opt.continue_train:
            self.netSR = util.load_network(self.netSR, 'SR', opt.which_epoch, opt)
            if opt.isTrain:
                self.netD = util.load_network(self.netD, 'D', opt.which_epoch, opt)
            if self.use_E:
                self.netE = util.load_network(self.netE, 'E', opt.which_epoch, opt)

    def get_logs(self):
        return self.logs

    def forward(self, data, mode, **kwargs):
        input_semantics = data.get("input_semantics", None)
        image_lr = data.get("image_lr", None)
        image_hr = data.get("image_hr", None)
        guiding_image = data.get("guiding_image", None)
        guiding_label = data.get("guiding_label", None)
        encoded_style = data.get("encoded_style", None)
        if mode == 'generator':
            g_loss, generated = self.compute_generator_loss(
                input_semantics, image_hr, image_lr, guiding_image, guiding_label)
            self.logs['image/downsized'] = image_lr
            return g_loss, generated
        elif mode == 'discriminator':
            return self.compute_discriminator_loss(
                input_semantics, image_hr, image_lr, guiding_image, guiding_label)
        elif mode == 'inference':
            with torch.no_grad():
                fake_image, _, _ = self.generate_fake(
                    input_semantics=input_semantics, image_downsized=image_lr, full_image=image_hr,
                    no_noise=True, guiding_image=guiding_image, guiding_label=guiding_label)
            data["fake_image"] = fake_image
            return util.filter_none(data)
        elif mode == 'encode_only':
            encoded_style, _ = self.encode_style(
                downscaled_image=image_lr, input_semantics=input_semantics, full_image=image_hr,
                no_noise=True, guiding_image=guiding_image, guiding_label=guiding_label,
                encode_full=self.opt.full_style_image)
            return encoded_style
        elif mode == 'demo':
            with torch.no_grad():
                fake_image = self.netSR(image_lr, seg=input_semantics, z=encoded_style)
            out = data
            out["fake_image"] = fake_image
            return util.filter_none(out)
        elif mode == 'baseline':
            image_baseline = F.interpolate(image_lr, (image_hr.shape[-2:]), mode='bicubic').clamp(-1, 1)
            return OrderedDict([("input_label", input_semantics), ("image_downsized", image_lr),
                                ("fake_image", image_baseline), ("image_full", image_hr)])
        elif mode == "inference_noise":
            with torch.no_grad():
                n = self.opt.batchSize
                lr_rep = image_lr.repeat_interleave(n, 0)
                sem_rep = input_semantics.repeat_interleave(n, 0)
                fake_image, _, _ = self.generate_fake(input_semantics=sem_rep, image_downsized=lr_rep,
                                                      encoded_style=None)
                fake_image = torch.stack([fake_image[i * n:i * n + n] for i in range(n)], dim=0)
                return OrderedDict([("input_label", input_semantics), ("image_downsized", image_lr),
                                    ("fake_image", fake_image), ("image_full", image_hr)])
        elif mode in _SWEEP_MODES:
            with torch.no_grad():
                return self._style_sweep(mode, data)
        elif mode == "inference_particular_combined":
            # sr_model.py:298-345: LR-only ("mini") style, optionally perturbed in opt.region_idx
            with torch.no_grad():
                style, _ = self.encode_style(input_semantics=input_semantics, downscaled_image=image_lr,
                                             no_noise=True, encode_full=False)
                if self.opt.noise_delta > 0:
                    ridx = self._region_idx(input_semantics)
                    style[:, ridx] = (style[:, ridx] + self.get_noise(style[:, ridx].shape,
                                                                      self.opt.noise_delta)).clamp(-1, 1)
                    style[:, _CONSISTENT_REGIONS] = style[:, [r + 1 for r in _CONSISTENT_REGIONS]]
                fake, _, _ = self.generate_fake(input_semantics=input_semantics, image_downsized=image_lr,
                                                encoded_style=style)
                out = OrderedDict([("input_label", input_semantics), ("image_downsized", image_lr),
                                   ("fake_image_original", fake), ("image_full", image_hr)])
                return self._with_guiding(out, data)
        elif mode == "inference_particular_full":
            # sr_model.py:347-380: style of the HR image itself, and (guided models) of the guiding image
            with torch.no_grad():
                style, _ = self.encode_style(no_noise=True, encode_full=True, guiding_image=image_hr,
                                             guiding_label=input_semantics)
                fake, _, _ = self.generate_fake(input_semantics=input_semantics, image_downsized=image_lr,
                                                encoded_style=style)
                out = OrderedDict([("input_label", input_semantics), ("image_downsized", image_lr),
                                   ("fake_image_original", fake), ("image_full", image_hr)])
                if self.opt.guiding_style_image:
                    gstyle, _ = self.encode_style(no_noise=True, encode_full=True, guiding_image=guiding_image,
                                                  guiding_label=guiding_label)
                    out["fake_image_guiding"], _, _ = self.generate_fake(
                        input_semantics=input_semantics, image_downsized=image_lr, encoded_style=gstyle)
                return self._with_guiding(out, data)
        elif mode == "inference_replace_semantics":
            raise NotImplementedError(
                "mode 'inference_replace_semantics' calls SRModel.preprocess_input, which does not exist "
                "in the reference either (sr_model.py:185); edit the label map before preprocessing and "
                "use mode 'inference'")
        else:
            raise ValueError("|mode| is invalid")

    # ---- demo-time style manipulation (sr_model.py:130-444) ------------------------------------------
    def _region_idx(self, input_semantics):
        r = getattr(self.opt, "region_idx", None)
        return list(r) if r else list(range(input_semantics.size(1)))

    def _with_guiding(self, out, data):
        if self.opt.guiding_style_image:
            for k_out, k_in in (("guiding_image_id", "guiding_image_id"), ("guiding_image", "guiding_image"),
                                ("guiding_input_label", "guiding_label")):
                if k_in in data:
                    out[k_out] = data[k_in]
        return out

    def get_noise(self, shape, delta):
        """sr_model.py:448-457."""
        dist = getattr(self.opt, "noise_dist", "normal")
        if dist == "normal":
            noise = torch.randn(shape).clamp(-1, 1) * delta
        elif dist == "uniform":
            noise = torch.rand(shape).clamp(-1, 1) * delta
        else:
            raise ValueError("Invalid noise distribution: {}".format(dist))
        return noise.cuda()

    def _style_sweep(self, mode, data):
        """The reference's style-manipulation modes (sr_model.py:130-166,198-296,381-444) all do the
        same thing: build, per sample b, a list of variant style matrices, generate one image per
        variant and lay the variants out side by side (or stacked with --dont_merge_fake).  The
        reference runs one batch-1 generator call per variant; here all variants of all samples go
        through ONE generator call (batch = B x variants), which is what the B200 kernels want."""
        import numpy as np
        opt = self.opt
        sem, lr, hr = data["input_semantics"], data["image_lr"], data.get("image_hr")
        gi, gl = data.get("guiding_image"), data.get("guiding_label")
        B = sem.size(0)
        ridx = self._region_idx(sem)
        n = getattr(opt, "n_interpolation", 5)
        delta = getattr(opt, "noise_delta", 0.0)
        sem_of = list(range(B))           # which sample's semantics / LR image each row uses
        variants = [[] for _ in range(B)]
        if mode == "inference_reference_semantics":
            # sr_model.py:198-218: every sample rendered with every sample's label map, own style
            style = None
        elif mode == "inference_interpolation_style":
            s_from, s_to = data["style_from"].to(sem.device), data["style_to"].to(sem.device)
            assert n % 2 == 1, "Please use an odd n such that the middle image has delta=0"
            for b in range(B):
                for t in np.linspace(0, 1, num=n):
                    variants[b].append((1 - t) * s_from[b] + t * s_to[b])
        else:
            if mode == "inference_interpolation" and "style_matrix" in data:
                style = data["style_matrix"]
            elif mode in ("inference_reference", "inference_reference_interpolation"):
                style, _ = self.encode_style(input_semantics=sem, full_image=hr, no_noise=True, encode_full=True,
                                             guiding_image=gi if mode == "inference_reference" else None,
                                             guiding_label=gl if mode == "inference_reference" else None)
            else:
                style, _ = self.encode_style(downscaled_image=lr, input_semantics=sem, full_image=hr,
                                             no_noise=True, guiding_image=gi, guiding_label=gl)
            for b in range(B):
                if mode == "inference_multi_modal":        # :130-166 random perturbations
                    for _ in range(n):
                        v = style[b].clone()
                        v[ridx] = (v[ridx] + self.get_noise(v[ridx].shape, delta)).clamp(-1, 1)
                        v[_CONSISTENT_REGIONS] = v[[r + 1 for r in _CONSISTENT_REGIONS]]
                        variants[b].append(v)
                elif mode == "inference_interpolation":    # :219-261 shift the style by -delta..delta
                    assert n % 2 == 1, "Please use an odd n such that the middle image has delta=0"
                    for step in np.linspace(-delta, delta, num=n):
                        v = style[b].clone()
                        v[ridx] = (v[ridx] + float(step)).clamp(-1, 1)
                        variants[b].append(v)
                elif mode == "inference_reference":        # :381-410 regions of every other sample
                    for other in range(B):
                        v = style[b].clone()
                        v[ridx] = style[other, ridx].clamp(-1, 1)
                        variants[b].append(v)
                elif mode == "inference_reference_interpolation":   # :411-444 towards the next sample
                    a = style[b].clone()
                    tgt = style[(b + 1) % B].clone() * getattr(opt, "manipulate_scale", 1.0)
                    for t in np.linspace(0, 1, num=n):
                        # (the reference updates style_a in place, so each step starts from the last)
                        a[ridx] = ((1 - float(t)) * a[ridx] + float(t) * tgt[ridx]).clamp(-1, 1)
                        variants[b].append(a.clone())
        if mode == "inference_reference_semantics":
            # NOTE reference quirk (:205-207): the inner loop overwrites current_semantics[b] with every
            # sample's map in turn, so what is rendered is sample b <- the LAST sample's semantics
            rows_sem, rows_lr = [], []
            for b in range(B):
                cur = sem.clone()
                cur[b] = sem[B - 1]
                rows_sem.append(cur)
                rows_lr.append(lr)
            fake, _, _ = self.generate_fake(input_semantics=torch.cat(rows_sem, 0),
                                            image_downsized=torch.cat(rows_lr, 0))
            fake_out = torch.cat(list(fake.split(B, 0)), -1)
            applied = []
        else:
            per = len(variants[0])
            z = torch.stack([v for b in range(B) for v in variants[b]], 0)
            rep = torch.tensor([sem_of[b] for b in range(B) for _ in range(per)], device=sem.device)
            fake, _, _ = self.generate_fake(input_semantics=sem[rep], image_downsized=lr[rep], encoded_style=z)
            fake = fake.view(B, per, *fake.shape[1:])
            if getattr(opt, "dont_merge_fake", False):
                fake_out = fake                                              # [B, variants, 3, H, W]
                applied = [torch.stack(variants[b]) for b in range(B)]
            else:
                fake_out = torch.cat([fake[:, i] for i in range(per)], -1)  # variants side by side
                applied = []
        out = OrderedDict([("input_label", sem), ("image_downsized", lr), ("fake_image", fake_out),
                           ("image_full", hr)])
        if mode in ("inference_interpolation", "inference_interpolation_style", "inference_multi_modal"):
            out["style"] = applied
        return self._with_guiding(out, data)

    def create_optimizers(self, opt):
        """sr_model.py:469-495: Adam; TTUR lr/2 for G (+E), 2*lr for D; "mini" params at lr_G/4."""
        SR_params = list(self.netSR.parameters())
        SR_params_low_lr = list()
        if self.use_E:
            for name, param in self.netE.named_parameters():
                (SR_params_low_lr if "mini" in name else SR_params).append(param)
        D_params = list(self.netD.parameters()) if opt.isTrain else []
        beta1, beta2 = opt.beta1, opt.beta2
        SR_lr, D_lr = (opt.lr, opt.lr) if opt.no_TTUR else (opt.lr / 2, opt.lr * 2)
        print("lr G: {}, lr D: {}".format(SR_lr, D_lr))
        from ..config import config
        # capturable: the step counters live on the device, so optimizer.step() can be part of a
        # captured CUDA graph (config.cuda_graphs); same arithmetic
        kw = {"capturable": True} if config.cuda_graphs else {}
        if config.fused_adam:
            kw["fused"] = True   # one multi-tensor kernel per parameter group instead of ~10 foreach passes
        optimizer_SR = torch.optim.Adam([{"params": SR_params},
                                         {"params": SR_params_low_lr, "lr": SR_lr / 4}],
                                        lr=SR_lr, betas=(beta1, beta2), **kw)
        optimizer_D = torch.optim.Adam(D_params, lr=D_lr, betas=(beta1, beta2), **kw)
        return optimizer_SR, optimizer_D

    def save(self, epoch):
        util.save_network(self.netSR, 'SR', epoch, self.opt)
        util.save_network(self.netD, 'D', epoch, self.opt)
        if self.use_E:
            util.save_network(self.netE, 'E', epoch, self.opt)

    # ------------------------------------------------------------------------------------------
    def initialize_networks(self, opt):
        self.netSR = networks.define_SR(opt)
        self.netD = networks.define_D(opt) if opt.isTrain else None
        self.netE = networks.define_E(opt) if self.use_E else None
        self.load_weights()
        return self.netSR, self.netD, self.netE

    def compute_generator_loss(self, input_semantics, image_full, image_downsized, guiding_image,
                               guiding_label):
        """sr_model.py:518-545."""
        SR_losses = {}
        style_image = guiding_image if self.opt.guiding_style_image else image_full
        fake_image, _, _ = self.generate_fake(
            input_semantics=input_semantics, image_downsized=image_downsized,
            full_image=style_image, guiding_image=guiding_image, guiding_label=guiding_label)
        # (for_generator=True: the discriminator runs with detached parameters - the generator step only
        # needs its INPUT gradient; the reference computes the weight gradients too and discards them,
        # trainer_manager.py:49)
        pred_fake, pred_real = self.discriminate(input_semantics, fake_image, image_full,
                                                 for_generator=True)
        SR_losses['GAN'] = self.criterionGAN(pred_fake, True, for_discriminator=False)
        if not self.opt.no_ganFeat_loss:
            num_D = len(pred_fake)
            GAN_Feat_loss = torch.zeros(1, device=fake_image.device)
            for i in range(num_D):
                for j in range(len(pred_fake[i]) - 1):  # last output is the prediction itself
                    unweighted = self.criterionFeat(pred_fake[i][j], pred_real[i][j].detach())
                    GAN_Feat_loss = GAN_Feat_loss + unweighted * self.opt.lambda_feat / num_D
            SR_losses['GAN_Feat'] = GAN_Feat_loss
        if not self.opt.no_vgg_loss:
            SR_losses['VGG'] = self.criterionVGG(fake_image, image_full) * self.opt.lambda_vgg
        return SR_losses, fake_image

    def compute_discriminator_loss(self, input_semantics, image_full, image_downsized,
                                   guiding_image, guiding_label):
        """sr_model.py:547-564."""
        D_losses = {}
        with torch.no_grad():
            fake_image, _, _ = self.generate_fake(
                input_semantics=input_semantics, image_downsized=image_downsized,
                full_image=image_full, guiding_image=guiding_image, guiding_label=guiding_label)
            fake_image = fake_image.detach()
        # (the reference also calls fake_image.requires_grad_() here, sr_model.py:556; nothing
        # reads that gradient, so the input-gradient kernels of the first layer are skipped)
        pred_fake, pred_real = self.discriminate(input_semantics, fake_image, image_full)
        D_losses['D_Fake'] = self.criterionGAN(pred_fake, False, for_discriminator=True)
        D_losses['D_Real'] = self.criterionGAN(pred_real, True, for_discriminator=True)
        return D_losses

    def generate_fake(self, input_semantics, image_downsized, encoded_style=None, full_image=None,
                      no_noise=False, guiding_image=None, guiding_label=None):
        """sr_model.py:566-580."""
        encoder_activations = None
        if encoded_style is None and "style" in self.opt.netE:
            encoded_style, encoder_activations = self.encode_style(
                downscaled_image=image_downsized, input_semantics=input_semantics,
                full_image=full_image, no_noise=no_noise, guiding_image=guiding_image,
                guiding_label=guiding_label, encode_full=self.opt.full_style_image)
        fake_image = self.netSR(image_downsized, seg=input_semantics, z=encoded_style)
        return fake_image, encoder_activations, encoded_style

    def get_encoder_inputs(self, downscaled_image=None, input_semantics=None, full_image=None,
                           encode_full=False, guiding_image=None, guiding_label=None):
        """sr_model.py:582-632. The coin flip uses Python's `random` like the reference; under
        data parallelism every rank seeds it identically (managers/base_manager.py)."""
        style_semantics = input_semantics
        style_image = downscaled_image
        if self.model_variant == "guided":
            mode = "full"
            if self.opt.guiding_style_image:
                style_semantics, style_image = guiding_label, guiding_image
            else:
                style_image = full_image
        elif self.model_variant == "independent":
            if encode_full or (self.training and random.random() < 0.5):
                mode = "full"
                self.last_encoded_style_is_full = True
                if self.opt.guiding_style_image:
                    style_semantics, style_image = guiding_label, guiding_image
                else:
                    style_image = full_image
            else:
                mode = "mini"
                self.last_encoded_style_is_full = False
        else:
            raise NotImplementedError()
        return style_image, style_semantics, mode

    def encode_style(self, downscaled_image=None, input_semantics=None, full_image=None,
                     encode_full=False, no_noise=None, guiding_image=None, guiding_label=None):
        """sr_model.py:634-650."""
        style_image, style_semantics, mode = self.get_encoder_inputs(
            downscaled_image=downscaled_image, input_semantics=input_semantics,
            full_image=full_image, encode_full=encode_full, guiding_image=guiding_image,
            guiding_label=guiding_label)
        if self.model_variant == "independent" and not no_noise:
            no_noise = random.random() < 0.5
            self.last_encoded_style_is_noisy = not no_noise
        return self.netE(style_image, style_semantics, mode=mode, no_noise=no_noise)

    def discriminate(self, input_semantics, fake_image, real_image, for_generator=False):
        """sr_model.py:655-668: D sees [seg | fake] and [seg | real] in one batch. The two cats and
        the NCHW->NHWC conversion are one kernel (ops.disc_input) on the uint8 label map."""
        from ..data.onehot import labels_of
        labels, _ = labels_of(input_semantics)
        L = input_semantics.shape[1]
        cp = (L + 3 + 31) // 32 * 32  # zero channels up to a multiple of 32 (tcgen05 epilogue width)
        x = ops.DiscInputFn.apply(labels, fake_image.contiguous().float(),
                                  real_image.contiguous().float(), L, cp)
        out = self.netD.forward_nhwc(x, detach_params=for_generator)
        out = [[t.permute(0, 3, 1, 2) for t in scale] for scale in out]
        return self.divide_pred(out)

    def divide_pred(self, pred):
        """sr_model.py:671-683."""
        if type(pred) == list:
            fake = [[t[:t.size(0) // 2] for t in p] for p in pred]
            real = [[t[t.size(0) // 2:] for t in p] for p in pred]
        else:
            fake = pred[:pred.size(0) // 2]
            real = pred[pred.size(0) // 2:]
        return fake, real

    def use_gpu(self):
        return True
