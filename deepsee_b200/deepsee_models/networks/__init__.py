"""Network registry (reference: deepsee_models/networks/__init__.py:12-58)."""
import torch

from .base_network import BaseNetwork
from .discriminator import MultiscaleDiscriminator, NLayerDiscriminator
from .encoder import FullStyleEncoder, MinistyleEncoder, CombinedstyleEncoder
from .loss import GANLoss, VGGLoss
from .sr import DeepSEESR

_REGISTRY = {
    'sr': {'deepsee': DeepSEESR},
    'discriminator': {'multiscale': MultiscaleDiscriminator, 'nlayer': NLayerDiscriminator},
    'encoder': {'fullstyle': FullStyleEncoder, 'ministyle': MinistyleEncoder,
                'combinedstyle': CombinedstyleEncoder},
}


def find_network_using_name(target_network_name, filename):
    key = target_network_name.replace('_', '').lower()
    try:
        network = _REGISTRY[filename][key]
    except KeyError:
        raise ValueError('In %s there is no network named %s' % (filename, target_network_name))
    assert issubclass(network, BaseNetwork)
    return network


def modify_commandline_options(parser, is_train):
    opt, _ = parser.parse_known_args()
    parser = find_network_using_name(opt.netG, 'sr').modify_commandline_options(parser, is_train)
    if is_train:
        parser = find_network_using_name(opt.netD, 'discriminator').modify_commandline_options(
            parser, is_train)
    parser = find_network_using_name(opt.netE, 'encoder').modify_commandline_options(parser, is_train)
    return parser


def create_network(cls, opt):
    """networks/__init__.py:37-43: construct, move to the GPU, init weights. On the B200 path the
    networks always live on the current CUDA device (one process per GPU); `gpu_ids == []`
    (the reference's CPU mode) is rejected when the model is used, not here, so that option
    handling and state_dict tooling still work on a CPU-only box."""
    net = cls(opt)
    if torch.cuda.is_available():
        net.cuda()
    net.init_weights(opt.init_type, opt.init_variance)
    return net


def define_D(opt):
    return create_network(find_network_using_name(opt.netD, 'discriminator'), opt)


def define_E(opt):
    return create_network(find_network_using_name(opt.netE, 'encoder'), opt)


def define_SR(opt):
    return create_network(find_network_using_name('deepsee', 'sr'), opt)
