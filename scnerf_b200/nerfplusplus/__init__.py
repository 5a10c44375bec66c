"""Host-side mirror of the reference's ``nerfplusplus/`` hot-path modules:
same module names, function names, argument meaning and error behaviour; every computation is a call
into ``libscnerf_b200.so`` (include/scnerf_b200_nerfpp.h)."""
from .nerf_network import Embedder, MLPNet                                   # noqa: F401
from .ddp_model import NerfNet, NerfNetWithAutoExpo, depth2pts_outside, remap_name   # noqa: F401
from .nerf_sample_ray_split import render_ray_from_camera                    # noqa: F401
from .ddp_train_nerf import intersect_sphere, perturb_samples, sample_pdf    # noqa: F401
