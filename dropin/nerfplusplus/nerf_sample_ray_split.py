"""Drop-in shim for nerfplusplus/nerf_sample_ray_split.py.

The reference module mixes dataset code (``RaySamplerSingleImage``, ``get_rays_single_image`` — out of scope,
kept) with the hot-path ray generator ``render_ray_from_camera`` (:196-258).  The reference's own file is executed
in this namespace and ``render_ray_from_camera`` is then rebound to the CUDA implementation, so
``RaySamplerSingleImage.random_sample`` / ``get_all`` (:139-188, :100-137) call it too."""
from _scnerf_shim import load_next as _load_next
_load_next("nerf_sample_ray_split", globals())
from scnerf_b200.nerfplusplus.nerf_sample_ray_split import render_ray_from_camera  # noqa: E402,F401
__scnerf_impl__ = "scnerf_b200.nerfplusplus.nerf_sample_ray_split"
