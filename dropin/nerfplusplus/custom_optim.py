"""Drop-in shim for nerfplusplus/custom_optim.py: the fused multi-tensor CustomAdamOptimizer."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.custom_optim")
