"""BaseNetwork: init / bookkeeping shared by the networks (reference:
deepsee_models/networks/base_network.py:11-80)."""
import torch.nn as nn
from torch.nn import init


class BaseNetwork(nn.Module):
    def __init__(self):
        super().__init__()

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def print_network(self):
        n = sum(p.numel() for p in self.parameters())
        print('Network [%s] was created. Total number of parameters: %.1f million. '
              'To see the architecture, do print(network).' % (type(self).__name__, n / 1000000))

    def init_weights(self, init_type='normal', gain=0.02):
        """Same rules as base_network.py:28-59. Like the reference it writes ``m.weight.data``,
        which for spectral-normalised convs is the derived tensor, not ``weight_orig``."""
        def init_func(m):
            classname = m.__class__.__name__
            if classname.find('BatchNorm2d') != -1:
                if getattr(m, 'weight', None) is not None:
                    init.normal_(m.weight.data, 1.0, gain)
                if getattr(m, 'bias', None) is not None:
                    init.constant_(m.bias.data, 0.0)
            elif hasattr(m, 'weight') and (classname.find('Conv') != -1 or
                                           classname.find('Linear') != -1):
                if init_type == 'normal':
                    init.normal_(m.weight.data, 0.0, gain)
                elif init_type == 'xavier':
                    init.xavier_normal_(m.weight.data, gain=gain)
                elif init_type == 'xavier_uniform':
                    init.xavier_uniform_(m.weight.data, gain=1.0)
                elif init_type == 'kaiming':
                    init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
                elif init_type == 'orthogonal':
                    init.orthogonal_(m.weight.data, gain=gain)
                elif init_type == 'none':
                    m.reset_parameters()
                else:
                    raise NotImplementedError(
                        'initialization method [%s] is not implemented' % init_type)
                if getattr(m, 'bias', None) is not None:
                    init.constant_(m.bias.data, 0.0)

        self.apply(init_func)
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights(init_type, gain)

    def compute_latent_vector_size(self, opt):
        """base_network.py:61-80."""
        levels = {'normal': 5, 'more': 6, 'most': 7}
        if opt.num_upsampling_layers not in levels:
            raise ValueError('opt.num_upsampling_layers [%s] not recognized' %
                             opt.num_upsampling_layers)
        n = levels[opt.num_upsampling_layers]
        self.output_size = opt.crop_size
        sw = self.output_size // (2 ** n)
        if self.output_size % 2 ** n != 0:
            sw += 1
        sh = round(sw / opt.aspect_ratio)
        return sw, sh
