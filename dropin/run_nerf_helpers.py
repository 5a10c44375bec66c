"""Drop-in shim for the reference's ``run_nerf_helpers`` module: with this directory ahead of the reference's on
sys.path, ``import run_nerf_helpers`` / ``from run_nerf_helpers import ...`` resolves to the B200 implementation
(scnerf_b200.run_nerf_helpers).  See INTEGRATION.md §1."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.run_nerf_helpers")
