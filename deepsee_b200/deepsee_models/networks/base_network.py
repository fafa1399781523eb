"""BaseNetwork: initialisation / bookkeeping shared by the networks (reference:
deepsee_models/networks/base_network.py:11-80; same public methods and the same initialisation
rules, expressed as a dispatch table)."""
import torch.nn as nn
from torch.nn import init

# init_type -> initialiser of a conv / linear weight tensor
_WEIGHT_INIT = {
    'normal': lambda w, gain: init.normal_(w, 0.0, gain),
    'xavier': lambda w, gain: init.xavier_normal_(w, gain=gain),
    'xavier_uniform': lambda w, gain: init.xavier_uniform_(w, gain=1.0),
    'kaiming': lambda w, gain: init.kaiming_normal_(w, a=0, mode='fan_in'),
    'orthogonal': lambda w, gain: init.orthogonal_(w, gain=gain),
}
_UPSAMPLING_LEVELS = {'normal': 5, 'more': 6, 'most': 7}


def _zero_bias(module):
    if getattr(module, 'bias', None) is not None:
        init.constant_(module.bias.data, 0.0)


class BaseNetwork(nn.Module):
    def __init__(self):
        super().__init__()

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def print_network(self):
        millions = sum(p.numel() for p in self.parameters()) / 1e6
        print('Network [%s] was created. Total number of parameters: %.1f million. '
              'To see the architecture, do print(network).' % (type(self).__name__, millions))

    def init_weights(self, init_type='normal', gain=0.02):
        """base_network.py:28-59.  Batch-norm scales ~ N(1, gain), conv / linear weights by `init_type`,
        biases zero.  Like the reference this writes ``module.weight.data``, which for a
        spectral-normalised conv is the derived tensor and leaves ``weight_orig`` at torch's default."""
        if init_type != 'none' and init_type not in _WEIGHT_INIT:
            raise NotImplementedError('initialization method [%s] is not implemented' % init_type)

        def visit(module):
            kind = type(module).__name__
            if 'BatchNorm2d' in kind:
                if getattr(module, 'weight', None) is not None:
                    init.normal_(module.weight.data, 1.0, gain)
                _zero_bias(module)
            elif hasattr(module, 'weight') and ('Conv' in kind or 'Linear' in kind):
                if init_type == 'none':
                    module.reset_parameters()
                else:
                    _WEIGHT_INIT[init_type](module.weight.data, gain)
                _zero_bias(module)

        self.apply(visit)
        for child in self.children():
            if hasattr(child, 'init_weights'):
                child.init_weights(init_type, gain)

    def compute_latent_vector_size(self, opt):
        """base_network.py:61-80: side of the coarsest feature map, rounded up."""
        if opt.num_upsampling_layers not in _UPSAMPLING_LEVELS:
            raise ValueError('opt.num_upsampling_layers [%s] not recognized' % opt.num_upsampling_layers)
        factor = 2 ** _UPSAMPLING_LEVELS[opt.num_upsampling_layers]
        self.output_size = opt.crop_size
        sw = -(-self.output_size // factor)
        return sw, round(sw / opt.aspect_ratio)
