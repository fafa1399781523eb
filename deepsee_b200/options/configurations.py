"""Name-based option presets (reference: options/configurations.py:3-43) plus the defaults of the
fields the hot path reads (options/base_options.py:32-142, options/train_options.py:11-77,
options/demo_options.json), so a model can be built without the reference's argparse stack."""
from ..util.util import ObjectDict

DEFAULTS = dict(
    name='8x_independent_128x128', model='sr', ngf=32, nef=32, ndf=32, label_nc=19, semantic_nc=19,
    output_nc=3, contain_dontcare_label=False, regional_style_size=128, max_fm_size=256,
    norm_G='spectrallateseansyncbatch3x3', norm_D='spectralinstance', norm_E='spectralinstance',
    netG='deepsee', netE='combinedstyle', netD='multiscale', netD_subarch='n_layer', num_D=2,
    n_layers_D=4, start_size=16, crop_size=128, load_size=128, aspect_ratio=1.0,
    num_upsampling_layers='normal', add_noise=False, efficient=False, model_parallel_mode=0,
    noisy_style_scale=0.2, noisy_style_dist='uniform', random_style_matrix=False,
    full_style_image=False, guiding_style_image=False, downsampling_method='bicubic',
    init_type='xavier', init_variance=0.02, gpu_ids=[0], gpu_info=False, isTrain=False,
    batchSize=1, checkpoints_dir='./checkpoints', which_epoch='latest', continue_train=False,
    gan_mode='hinge', lambda_feat=10.0, lambda_vgg=10.0, no_ganFeat_loss=False, no_vgg_loss=True,
    lr=0.0002, beta1=0.0, beta2=0.9, no_TTUR=False, gradient_clip=-1.0, niter=50, niter_decay=25,
    dataset='celebamaskhq', noise_dist='normal', noise_delta=0.0, n_interpolation=5,
    region_idx=None, dont_merge_fake=False, manipulate_scale=1.0,
)


# Presets keyed by what the model name contains (same rules as the reference's if/elif chain,
# options/configurations.py:16-43).  A name needs one resolution preset and one variant preset.
_RESOLUTION_PRESETS = (
    (lambda n: "8x_" in n and "128x128" in n,
     dict(start_size=16, crop_size=128, load_size=128, dataset="celeba", add_noise=True)),
    (lambda n: "8x_" in n and "256x256" in n,
     dict(start_size=32, crop_size=256, load_size=256, dataset="celebamaskhq", add_noise=True,
          max_fm_size=256)),
    (lambda n: "32x_" in n,
     dict(start_size=16, crop_size=512, load_size=512, dataset="celebamaskhq", add_noise=False,
          max_fm_size=256)),
)
_VARIANT_PRESETS = (
    ("independent", dict(netE="combinedstyle", noisy_style_scale=0.2)),
    ("guided", dict(netE="fullstyle", noisy_style_scale=0.05, guiding_style_image=True)),
)


def _apply(opt, fields):
    for key, value in fields.items():
        setattr(opt, key, value)
    return opt


def get_config_independent(opt):
    return _apply(opt, dict(_VARIANT_PRESETS)["independent"])


def get_config_guided(opt):
    return _apply(opt, dict(_VARIANT_PRESETS)["guided"])


def get_opt_config(opt, name):
    """Fills `opt` from the presets a model name such as "8x_independent_256x256" selects."""
    resolution = next((fields for matches, fields in _RESOLUTION_PRESETS if matches(name)), None)
    variant = next((fields for key, fields in _VARIANT_PRESETS if key in name), None)
    if resolution is None or variant is None:
        raise ValueError("Invalid name: '{}'. Please specify your options yourself.".format(name))
    return _apply(_apply(opt, resolution), variant)


def make_opt(name=None, **overrides):
    """ObjectDict with every field the hot path reads; `name` applies the reference's preset."""
    opt = ObjectDict(dict(DEFAULTS))
    if name is not None:
        opt.name = name
        opt = get_opt_config(opt, name)
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt
