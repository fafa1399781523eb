"""TrainerManager (reference: managers/trainer_manager.py:6-99).

Same public surface as the reference (`run_generator_one_step`, `run_discriminator_one_step`,
`get_latest_losses`, `get_latest_generated`, `get_logs`, `save`, `update_learning_rate`,
`sr_model`, `sr_model_on_one_gpu`), so train.py drives it unchanged.  Differences:
  * one process per GPU; after each backward the G(+E) or D gradients are averaged across ranks
    with ONE flat NCCL all-reduce (..parallel.GradBucket) - the only data-path collective;
  * checkpoints are written by rank 0 only.
"""
from torch.nn.utils import clip_grad_value_

from .. import parallel
from .base_manager import BaseManager


class TrainerManager(BaseManager):
    def __init__(self, opt):
        super().__init__(opt, create_model=True)
        assert opt.isTrain
        self.optimizer_G, self.optimizer_D = self.sr_model_on_one_gpu.create_optimizers(opt)
        self.old_lr = opt.lr
        self.generated = None
        self.logs = {}
        self.g_losses = {}
        self.d_losses = {}
        m = self.sr_model_on_one_gpu
        g_params = list(m.netSR.parameters()) + (list(m.netE.parameters()) if m.use_E else [])
        self._bucket_G = parallel.GradBucket(g_params) if parallel.is_dist() else None
        self._bucket_D = parallel.GradBucket(list(m.netD.parameters())) if parallel.is_dist() else None

    def get_logs(self):
        return {**self.logs, **self.sr_model_on_one_gpu.get_logs()}

    def preprocess_input(self, data):
        return super().preprocess(data, from_dataloader=True)

    def run_generator_one_step(self, data):
        """trainer_manager.py:32-46."""
        self.optimizer_G.zero_grad()
        data_preprocessed = self.preprocess_input(data)
        g_losses, generated = self.sr_model(data_preprocessed, mode='generator')
        g_loss = sum(g_losses.values()).mean()
        g_loss.backward()
        if self._bucket_G is not None:
            self._bucket_G.allreduce_mean()
        if self.opt.gradient_clip > 0:
            clip_grad_value_(self.sr_model.parameters(), self.opt.gradient_clip)
        self.optimizer_G.step()
        self.g_losses = g_losses
        self.generated = generated

    def run_discriminator_one_step(self, data):
        """trainer_manager.py:48-61."""
        self.optimizer_D.zero_grad()
        data_preprocessed = self.preprocess_input(data)
        d_losses = self.sr_model(data_preprocessed, mode='discriminator')
        d_loss = sum(d_losses.values()).mean()
        d_loss.backward()
        if self._bucket_D is not None:
            self._bucket_D.allreduce_mean()
        if self.opt.gradient_clip > 0:
            clip_grad_value_(self.sr_model.parameters(), self.opt.gradient_clip)
        self.optimizer_D.step()
        self.d_losses = d_losses

    def get_latest_losses(self):
        return {**self.g_losses, **self.d_losses}

    def get_latest_generated(self):
        return self.generated

    def save(self, epoch):
        if parallel.rank() == 0:
            self.sr_model_on_one_gpu.save(epoch)

    def update_learning_rate(self, epoch):
        """trainer_manager.py:76-99."""
        if epoch > self.opt.niter:
            lrd = self.opt.lr / self.opt.niter_decay
            new_lr = self.old_lr - lrd
        else:
            new_lr = self.old_lr
        if new_lr != self.old_lr:
            if self.opt.no_TTUR:
                new_lr_G, new_lr_D = new_lr, new_lr
            else:
                new_lr_G, new_lr_D = new_lr / 2, new_lr * 2
            for param_group in self.optimizer_D.param_groups:
                param_group['lr'] = new_lr_D
            for param_group in self.optimizer_G.param_groups:
                param_group['lr'] = new_lr_G
            print('update learning rate: %f -> %f' % (self.old_lr, new_lr))
            self.old_lr = new_lr
