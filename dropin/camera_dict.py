"""Drop-in shim for the reference's ``camera_dict`` module: with this directory ahead of the reference's on
sys.path, ``import camera_dict`` / ``from camera_dict import ...`` resolves to the B200 implementation
(scnerf_b200.camera_dict).  See INTEGRATION.md §1."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.camera_dict")
