"""Drop-in shim for the reference's ``get_rays`` module: with this directory ahead of the reference's on
sys.path, ``import get_rays`` / ``from get_rays import ...`` resolves to the B200 implementation
(scnerf_b200.get_rays).  See INTEGRATION.md §1."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.get_rays")
