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
    "loss",
    "loss.flow",
    "utils",
    "utils.iwe",
    "dataloader",
    "dataloader.encodings",
)


def install_dropin(force=False):
    """Alias this package's modules under the reference's import names.  Returns the list of names installed."""
    done = []
    for name in _DROPIN:
        if name in sys.modules and not force and not sys.modules[name].__name__.startswith("event_flow_b200"):
            raise ImportError(f"module '{name}' is already imported from {getattr(sys.modules[name], '__file__', '?')}")
        sys.modules[name] = importlib.import_module("event_flow_b200." + name)
        done.append(name)
    return done
