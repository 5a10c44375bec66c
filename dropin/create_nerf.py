"""Drop-in shim: put this directory FIRST on PYTHONPATH and NeRF/run_nerf.py's
``import create_nerf`` / ``from create_nerf import ...`` resolves to the B200 implementation."""
from scnerf_b200.create_nerf import *  # noqa: F401,F403
from scnerf_b200 import create_nerf as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
