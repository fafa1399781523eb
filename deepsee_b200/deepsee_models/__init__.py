"""Model registry (reference: deepsee_models/__init__.py:10-44)."""


def find_model_using_name(model_name):
    if model_name.replace('_', '').lower() != 'sr':
        raise ValueError("There is no model named %s (only 'sr')" % model_name)
    from .sr_model import SRModel
    return SRModel


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    model = find_model_using_name(opt.model)
    instance = model(opt)
    print("model [%s] was created" % (type(instance).__name__))
    return instance
