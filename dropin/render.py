"""Drop-in shim: put this directory FIRST on PYTHONPATH and NeRF/run_nerf.py's
``import render`` / ``from render import ...`` resolves to the B200 implementation."""
from scnerf_b200.render import *  # noqa: F401,F403
from scnerf_b200 import render as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
