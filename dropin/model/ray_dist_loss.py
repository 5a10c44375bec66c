"""Drop-in shim for model/ray_dist_loss.py (imported as ``model.ray_dist_loss``)."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.ray_dist_loss")
