"""Drop-in shim for the reference's ``camera_model`` module: with this directory ahead of the reference's on
sys.path, ``import camera_model`` / ``from camera_model import ...`` resolves to the B200 implementation
(scnerf_b200.camera_model).  See INTEGRATION.md §1."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.camera_model",
          leak=("numpy as np", "torch", "torch.nn as nn", "wandb", "sys"))   # model/camera_model.py:1-9
