"""Drop-in shim for model/camera_model.py (imported as ``model.camera_model``)."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.camera_model",
          leak=("numpy as np", "torch", "torch.nn as nn", "wandb", "sys"))   # model/camera_model.py:1-9
