"""DemoManager (reference: managers/demo_manager.py:4-51)."""
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
        return self.sr_model.forward(pre, "demo")
