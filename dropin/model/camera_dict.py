"""Drop-in shim for model/camera_dict.py (imported as ``model.camera_dict``)."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.camera_dict")
