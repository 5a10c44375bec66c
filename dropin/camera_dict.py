"""Drop-in shim: put this directory FIRST on PYTHONPATH and NeRF/run_nerf.py's
``import camera_dict`` / ``from camera_dict import ...`` resolves to the B200 implementation."""
from scnerf_b200.camera_dict import *  # noqa: F401,F403
from scnerf_b200 import camera_dict as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
