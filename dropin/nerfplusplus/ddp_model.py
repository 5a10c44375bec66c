"""Drop-in shim for nerfplusplus/ddp_model.py: resolves to scnerf_b200.nerfplusplus.ddp_model."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.nerfplusplus.ddp_model")
