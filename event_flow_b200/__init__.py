"""
event_flow_b200 -- B200 (sm_100a) kernels behind the Python model / loss API of tudelft/event_flow.

The sub-packages mirror the reference's module tree so existing drivers keep working:
    event_flow_b200.models.model            <-> models/model.py            (LIFFireNet, PLIFFireNet, ...)
    event_flow_b200.models.spiking_submodules <-> models/spiking_submodules.py (ConvLIF, ConvLIFRecurrent, ...)
    event_flow_b200.loss.flow               <-> loss/flow.py               (EventWarping, ...)
    event_flow_b200.utils.iwe               <-> utils/iwe.py               (compute_pol_iwe, ...)
    event_flow_b200.dataloader.encodings    <-> dataloader/encodings.py    (events_to_channels, ...)
`install_dropin()` registers them under the reference's top-level names (`models`, `loss`, `utils`, `dataloader`) so that
`from models.model import LIFFireNet` in train_flow.py / eval_flow.py, and pickled checkpoints, resolve to this package.
All compute goes through libeventflow.so (see include/eventflow.h); there is no CPU or stock-PyTorch fallback.
"""

import importlib
import os
import sys

__version__ = "0.1.0"

_DROPIN = (
    "models",
    "models.base",
    "models.model",
    "models.model_util",
    "models.spiking_submodules",
    "models.spiking_util",
    "models.submodules",
    "models.unet",
    "loss",
    "loss.flow",
    "utils",
    "utils.iwe",
    "dataloader",
    "dataloader.encodings",
)


def install_dropin(force=False, reference_root=None):
    """
    Alias this package's modules under the reference's import names.  Returns the list of names installed.

    The reference's drivers also import modules this package does not replace (`utils.utils`, `utils.gradients`,
    `utils.visualization`, `dataloader.h5`, `configs.parser`, ...).  With `reference_root` (default: the current directory when it
    holds `train_flow.py`) the reference's own `models/`, `loss/`, `utils/`, `dataloader/` directories are appended to the search
    path of the aliased packages, so every submodule that is NOT replaced here still resolves to the reference's file.
    """
    if reference_root is None and os.path.isfile(os.path.join(os.getcwd(), "train_flow.py")):
        reference_root = os.getcwd()
    done = []
    for name in _DROPIN:
        if name in sys.modules and not force and not sys.modules[name].__name__.startswith("event_flow_b200"):
            raise ImportError(f"module '{name}' is already imported from {getattr(sys.modules[name], '__file__', '?')}")
        mod = importlib.import_module("event_flow_b200." + name)
        sys.modules[name] = mod
        if reference_root is not None and "." not in name:  # a package: fall back to the reference's directory for the rest
            ref_dir = os.path.join(reference_root, name)
            if os.path.isdir(ref_dir) and ref_dir not in mod.__path__:
                mod.__path__.append(ref_dir)
        done.append(name)
    return done
