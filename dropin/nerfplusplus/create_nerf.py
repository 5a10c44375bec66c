"""Drop-in shim for nerfplusplus/create_nerf.py: resolves to scnerf_b200.nerfplusplus.create_nerf."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.nerfplusplus.create_nerf")
