"""Drop-in shim for nerfplusplus/nerf_network.py: resolves to scnerf_b200.nerfplusplus.nerf_network."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.nerfplusplus.nerf_network")
