"""Checkpoint I/O and small helpers the hot path uses (reference: util/util.py:217-237,433-443)."""
import os
from collections import OrderedDict

import torch


def save_network(net, label, epoch, opt):
    """util/util.py:217-225: {"model": state_dict} -> <epoch>_net_<label>.pth (state on CPU)."""
    save_filename = '%s_net_%s.pth' % (epoch, label)
    save_dir = os.path.join(opt.checkpoints_dir, opt.name)
    os.makedirs(save_dir, exist_ok=True)
    state = OrderedDict((k, v.detach().cpu()) for k, v in net.state_dict().items())
    torch.save({"model": state}, os.path.join(save_dir, save_filename))


def load_network(net, label, epoch, opt):
    """util/util.py:228-237: accepts wrapped ({"model": ...}) or bare state_dicts."""
    save_filename = '%s_net_%s.pth' % (epoch, label)
    save_path = os.path.join(opt.checkpoints_dir, opt.name, save_filename)
    checkpoint = torch.load(save_path, map_location='cpu')
    net.load_state_dict(checkpoint["model"] if "model" in checkpoint else checkpoint)
    return net


def filter_none(ordered_dict):
    return OrderedDict([(k, v) for k, v in ordered_dict.items() if v is not None])


class ObjectDict(dict):
    def __init__(self, d):
        super().__init__()
        for k, v in d.items():
            setattr(self, k, v)


def gpu_info(message, opt):
    """The reference prints GPUtil utilisation here (util/util.py:426-430); a no-op on this path
    (profiling is done with ncu / CUDA events, see bench.py)."""
    return None
