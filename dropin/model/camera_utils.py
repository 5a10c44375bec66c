"""Drop-in shim for model/camera_utils.py (imported as ``model.camera_utils``)."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.camera_utils")
