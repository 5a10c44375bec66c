"""Drop-in shim: put this directory FIRST on PYTHONPATH and NeRF/run_nerf.py's
``import run_nerf_helpers`` / ``from run_nerf_helpers import ...`` resolves to the B200 implementation."""
from scnerf_b200.run_nerf_helpers import *  # noqa: F401,F403
from scnerf_b200 import run_nerf_helpers as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
