"""DemoManager (reference: managers/demo_manager.py:4-51).

`run` is the notebook's hot call: one batch-1 generator forward per widget interaction.  At batch 1
the forward is ~200 kernel launches of a few microseconds each, so the host, not the GPU, sets the
latency; with config.demo_graphs (default on) the forward is captured once per input shape as a CUDA
graph - label-map resizes, per-image style weights, every conv - and replayed on static buffers.
"""
import torch

from ..config import config
from ..data.onehot import OneHotLabels
from .base_manager import BaseManager


class DemoManager(BaseManager):
    def __init__(self, opt):
        super().__init__(opt)
        self.sr_model = self.sr_model.eval()
        self.sr_model_on_one_gpu = self.sr_model

    def compute_style_from_hr(self, inputs_hr):
        """demo_manager.py:12-29. The reference version calls a method that does not exist
        (BaseManager.preprocess_input) and passes keys SRModel never reads; this one does what its
        comments describe: encode every HR image with the full-resolution branch, then replace the
        listed regions of the first style matrix."""
        print("Encoding style from {} HR images...".format(len(inputs_hr)))
        styles = []
        for inp in inputs_hr:
            d = super().preprocess({"image_hr": inp["image_hr"], "semantics": inp["semantics"]})
            sem = self.preprocessor.preprocess_label(d["semantics"])
            full = self.opt.full_style_image
            self.opt.full_style_image = True
            try:
                styles.append(self.sr_model.forward(
                    {"image_hr": d["image_hr"], "input_semantics": sem}, "encode_only"))
            finally:
                self.opt.full_style_image = full
        encoded_style = styles[0]
        for i in range(1, len(inputs_hr)):
            for region_index in inputs_hr[i]["regions"]:
                encoded_style[:, region_index] = styles[i][:, region_index].detach()
        return encoded_style.clone()

    def compute_style_from_lr(self, data):
        print("Encoding style from LR image...")
        data = super().preprocess(data, from_dataloader=False)
        pre = {"image_lr": data["image_lr"],
               "input_semantics": self.preprocessor.preprocess_label(data["input_semantics"])}
        return self.sr_model.forward(pre, "encode_only")

    def run(self, data):
        assert "image_lr" in data.keys()
        assert "semantics" in data.keys()
        assert "encoded_style" in data.keys()
        data = super().preprocess(data, from_dataloader=False)
        pre = {"image_lr": data["image_lr"],
               "input_semantics": self.preprocessor.preprocess_label(data["semantics"]),
               "encoded_style": data["encoded_style"]}
        if config.demo_graphs and isinstance(pre["input_semantics"], OneHotLabels):
            return self._run_graphed(pre)
        return self.sr_model.forward(pre, "demo")

    def _run_graphed(self, pre):
        """Mode 'demo' (sr_model.py:116-122: netSR(image_lr, seg, z) under no_grad) as a CUDA graph
        replay.  First call per input shape: eager (fills the prepared-weight caches, sets kernel
        attributes); second: capture; afterwards: copy the three inputs into the static buffers and
        replay.  Out-of-range labels are still reported (one flag read per call)."""
        seg, lr, z = pre["input_semantics"], pre["image_lr"].contiguous().float(), pre["encoded_style"].float()
        if seg.bad is not None and int(seg.bad.item()) != 0:
            raise ValueError("DemoManager.run: label map holds values outside [0, %d)" % seg.num_classes)
        # (the captured launches point at the prepared-weight planes cached for the current parameter
        # versions: reloading weights makes a new key)
        wsig = sum(p._version for p in self.sr_model.netSR.parameters())
        key = (tuple(lr.shape), tuple(seg.labels.shape), tuple(z.shape), wsig)
        graphs = self.__dict__.setdefault("_demo_graphs", {})
        ent = graphs.get(key)
        if ent is None:
            graphs[key] = "warm"
            return self.sr_model.forward(pre, "demo")
        if ent == "warm":
            static = {"lr": lr.clone(), "labels": seg.labels.clone(), "z": z.contiguous().clone()}
            sseg = OneHotLabels(static["labels"], seg.num_classes, None)
            saved = config.check_onehot
            config.check_onehot = False
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g), torch.no_grad():
                    out = self.sr_model.netSR(static["lr"], seg=sseg, z=static["z"])
            finally:
                config.check_onehot = saved
            ent = graphs[key] = (g, static, out)
        g, static, out = ent
        static["lr"].copy_(lr, non_blocking=True)
        static["labels"].copy_(seg.labels, non_blocking=True)
        static["z"].copy_(z, non_blocking=True)
        g.replay()
        res = dict(pre)
        res["fake_image"] = out.clone()
        from ..util import util
        return util.filter_none(res)
