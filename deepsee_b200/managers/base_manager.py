"""BaseManager (reference: managers/base_manager.py:8-66).

The reference wraps SRModel in a single-process, multi-thread DataParallelWithCallback
(base_manager.py:17-21).  Here every GPU has its own process (torchrun: RANK / LOCAL_RANK /
WORLD_SIZE); the model is a plain SRModel on the local device and the managers all-reduce the
G and D gradients over NCCL (see ..parallel).  `sr_model` and `sr_model_on_one_gpu` are the same
object, which is what the reference exposes when it does not wrap.
"""

from ..data.preprocessor import Preprocessor
from ..deepsee_models.sr_model import SRModel
from .. import parallel


class BaseManager:
    def __init__(self, opt, create_model=True):
        self.opt = opt
        self.preprocessor = Preprocessor(opt)
        if create_model:
            self.create_model(opt)

    def create_model(self, opt):
        parallel.init_from_env()           # no-op when launched as a single process
        if parallel.is_dist():
            # the encoder coin flips (Python's global `random`, like the reference) must agree across
            # ranks; a single process keeps whatever seeding the caller did
            parallel.seed_python_random(0)
        self.sr_model = SRModel(opt)
        parallel.broadcast_module(self.sr_model)  # identical initial weights on every rank
        self.sr_model_on_one_gpu = self.sr_model

    def use_gpu(self):
        return True

    def preprocess(self, data, from_dataloader=False):
        data = self.preprocess_datatypes(data)
        data = self.preprocess_gpu(data)
        if from_dataloader:
            data = self.preprocess_from_dataloader(data)
        return data

    def preprocess_datatypes(self, data):
        for k in data:
            if 'label' in k or 'semantics' in k:
                data[k] = data[k].long()
        return data

    def preprocess_gpu(self, data):
        for k, v in data.items():
            if hasattr(v, "cuda"):
                data[k] = v.cuda(non_blocking=True)
        return data

    def preprocess_from_dataloader(self, data):
        """base_manager.py:50-66."""
        out = {
            "input_semantics": self.preprocessor.preprocess_label(data['label']),
            "image_lr": self.preprocessor.downsample_image(data['image']),
            "image_hr": data["image"],
        }
        if self.opt.guiding_style_image:
            out["guiding_image"] = data['guiding_image']
            out["guiding_label"] = self.preprocessor.preprocess_label(data['guiding_label'].long())
        return out
