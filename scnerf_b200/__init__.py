"""scnerf_b200 — B200-native (sm_100a) implementation of SCNeRF's volumetric-rendering hot path.

Python here is the host-side mirror of the reference's API (same module / function names as
NeRF/get_rays.py, NeRF/render.py, NeRF/create_nerf.py, NeRF/run_nerf_helpers.py,
model/camera_model.py); the work is done by csrc/libscnerf_b200.so through the C ABI in
include/scnerf_b200.h.  Importing the package does not load the library; the first call does,
and fails loudly if it is missing.
"""
__version__ = "0.1.0"
