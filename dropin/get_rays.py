"""Drop-in shim: put this directory FIRST on PYTHONPATH and NeRF/run_nerf.py's
``import get_rays`` / ``from get_rays import ...`` resolves to the B200 implementation."""
from scnerf_b200.get_rays import *  # noqa: F401,F403
from scnerf_b200 import get_rays as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
