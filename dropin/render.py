"""Drop-in shim for the reference's ``render`` module: with this directory ahead of the reference's on
sys.path, ``import render`` / ``from render import ...`` resolves to the B200 implementation
(scnerf_b200.render).  See INTEGRATION.md §1."""
from _scnerf_shim import reexport as _reexport
_reexport(globals(), "scnerf_b200.render")
