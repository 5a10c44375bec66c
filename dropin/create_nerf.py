"""Drop-in shim for the reference's ``create_nerf`` module: with this directory ahead of the reference's on
sys.path, ``import create_nerf`` / ``from create_nerf import ...`` resolves to the B200 implementation
(scnerf_b200.create_nerf).  See INTEGRATION.md §1."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.create_nerf")
