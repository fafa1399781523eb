"""TrainerManager (reference: managers/trainer_manager.py:6-99).

Same public surface as the reference (`run_generator_one_step`, `run_discriminator_one_step`,
`get_latest_losses`, `get_latest_generated`, `get_logs`, `save`, `update_learning_rate`,
`sr_model`, `sr_model_on_one_gpu`), so train.py drives it unchanged.  Differences:
  * one process per GPU; the G(+E) or D gradients are averaged across ranks through a flat
    bucket whose chunks are all-reduced over NCCL while the backward pass is still running
    (..parallel.GradBucket);
  * checkpoints are written by rank 0 only.
Both optimizer steps share one code path (`_optimize`).
"""
from torch.nn.utils import clip_grad_value_

from .. import parallel
from .base_manager import BaseManager


class TrainerManager(BaseManager):
    def __init__(self, opt):
        super().__init__(opt, create_model=True)
        assert opt.isTrain
        model = self.sr_model_on_one_gpu
        self.optimizer_G, self.optimizer_D = model.create_optimizers(opt)
        self.old_lr = opt.lr
        self.generated = None
        self.logs = {}
        self.g_losses, self.d_losses = {}, {}
        self._bucket_G = self._bucket_D = None
        if parallel.is_dist():
            g_params = list(model.netSR.parameters()) + (list(model.netE.parameters()) if model.use_E else [])
            self._bucket_G = parallel.GradBucket(g_params)
            self._bucket_D = parallel.GradBucket(list(model.netD.parameters()))

    # ---- one optimizer step (trainer_manager.py:32-61) ------------------------------------------------
    def _optimize(self, data, mode, optimizer, bucket):
        """zero_grad -> forward(mode) -> mean of the summed losses -> backward -> gradient all-reduce
        (multi-GPU) -> optional value clipping -> optimizer step.  Returns SRModel.forward's result."""
        optimizer.zero_grad()
        if bucket is not None:
            bucket.begin()        # .grad = views of the flat bucket; chunks all-reduce during backward
        result = self.sr_model(self.preprocess_input(data), mode=mode)
        losses = result[0] if mode == 'generator' else result
        sum(losses.values()).mean().backward()
        if bucket is not None:
            bucket.finish()
        if self.opt.gradient_clip > 0:
            clip_grad_value_(self.sr_model.parameters(), self.opt.gradient_clip)
        optimizer.step()
        return result

    def run_generator_one_step(self, data):
        self.g_losses, self.generated = self._optimize(data, 'generator', self.optimizer_G, self._bucket_G)

    def run_discriminator_one_step(self, data):
        self.d_losses = self._optimize(data, 'discriminator', self.optimizer_D, self._bucket_D)

    # ---- accessors train.py uses ----------------------------------------------------------------------
    def preprocess_input(self, data):
        return super().preprocess(data, from_dataloader=True)

    def get_logs(self):
        return {**self.logs, **self.sr_model_on_one_gpu.get_logs()}

    def get_latest_losses(self):
        return {**self.g_losses, **self.d_losses}

    def get_latest_generated(self):
        return self.generated

    def save(self, epoch):
        if parallel.rank() == 0:
            self.sr_model_on_one_gpu.save(epoch)

    def update_learning_rate(self, epoch):
        """trainer_manager.py:76-99: constant for `niter` epochs, then a linear decay by
        lr / niter_decay per epoch; with TTUR the generator runs at half and the discriminator at
        twice the base rate."""
        decay = self.opt.lr / self.opt.niter_decay if epoch > self.opt.niter else 0.0
        new_lr = self.old_lr - decay
        if new_lr == self.old_lr:
            return
        g_scale, d_scale = (1.0, 1.0) if self.opt.no_TTUR else (0.5, 2.0)
        for optimizer, scale in ((self.optimizer_G, g_scale), (self.optimizer_D, d_scale)):
            for group in optimizer.param_groups:
                group['lr'] = new_lr * scale
        print('update learning rate: %f -> %f' % (self.old_lr, new_lr))
        self.old_lr = new_lr
