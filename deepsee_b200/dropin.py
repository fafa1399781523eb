"""Makes the reference's own import paths resolve to the B200 implementation.

train.py / demo.py of mcbuehler/DeepSEE import `managers.trainer_manager.TrainerManager`,
`managers.demo_manager.DemoManager`, `deepsee_models.sr_model.SRModel`, `deepsee_models.networks`
and `data.preprocessor.Preprocessor` (train.py:7-15, demo.py:12-17).  `install()` registers this
package's modules under those names in `sys.modules`, so the reference's entry scripts run
unchanged on top of the CUDA library:

    import deepsee_b200.dropin; deepsee_b200.dropin.install()
    import train            # the reference's train.py, untouched
"""
import importlib
import sys

_ALIASES = {
    "managers": "deepsee_b200.managers",
    "managers.base_manager": "deepsee_b200.managers.base_manager",
    "managers.trainer_manager": "deepsee_b200.managers.trainer_manager",
    "managers.demo_manager": "deepsee_b200.managers.demo_manager",
    "managers.inference_manager": "deepsee_b200.managers.inference_manager",
    "deepsee_models": "deepsee_b200.deepsee_models",
    "deepsee_models.sr_model": "deepsee_b200.deepsee_models.sr_model",
    "deepsee_models.networks": "deepsee_b200.deepsee_models.networks",
    "deepsee_models.networks.sr": "deepsee_b200.deepsee_models.networks.sr",
    "deepsee_models.networks.architecture": "deepsee_b200.deepsee_models.networks.architecture",
    "deepsee_models.networks.normalization": "deepsee_b200.deepsee_models.networks.normalization",
    "deepsee_models.networks.encoder": "deepsee_b200.deepsee_models.networks.encoder",
    "deepsee_models.networks.discriminator": "deepsee_b200.deepsee_models.networks.discriminator",
    "deepsee_models.networks.loss": "deepsee_b200.deepsee_models.networks.loss",
    "data": "deepsee_b200.data",
    "data.preprocessor": "deepsee_b200.data.preprocessor",
}


def install(overwrite=False):
    """Registers the aliases; returns the list of names that were installed."""
    done = []
    for alias, target in _ALIASES.items():
        if alias in sys.modules and not overwrite:
            continue
        sys.modules[alias] = importlib.import_module(target)
        done.append(alias)
    return done
