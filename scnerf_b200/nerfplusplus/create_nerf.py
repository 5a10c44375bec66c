"""nerfplusplus/create_nerf.py:14-154 — ``create_nerf(rank, args, camera_info) -> (start, models, camera_model)``.

Same contract as the reference: ``models`` is an OrderedDict with ``cascade_level``, ``cascade_samples``,
``net_0`` … ``net_{L-1}`` (wrapped so state-dict keys carry the ``module.`` prefix of the reference's DDP
wrapper, checkpoints interchange), ``optim`` (``CustomAdamOptimizer`` on ``--use_custom_optim`` else
``torch.optim.Adam``, over cascade parameters then camera parameters — the order the positional weight
decay indexes); the newest ``*.pth`` under ``basedir/expname`` (or ``args.ckpt_path``) is reloaded with the
reference's ``load_camera`` / ``load_test`` rules; the camera curriculum (``add_ie`` / ``add_radial`` /
``add_od``) switches ``requires_grad`` off for the stages not reached yet.

Multi-GPU: one process per GPU as in the reference.  When ``torch.distributed`` is initialised with more than
one rank the networks are wrapped in ``DistributedDataParallel`` exactly like ``create_nerf.py:54-57`` (our
fused ``NerfNet`` autograd node leaves ordinary ``.grad`` tensors, so DDP's bucketed all-reduce works unchanged);
with one rank a prefix-preserving pass-through wrapper is used.  Deliberate deviation: the camera
parameters' gradients are averaged across ranks too (the reference leaves them rank-local, so its cameras
drift apart); ``SCNERF_SYNC_CAMERA=0`` restores the reference behaviour.

When the caller is the reference trainer itself (``nerfplusplus/ddp_train_nerf.py:359``), its module-level
``intersect_sphere`` / ``perturb_samples`` / ``sample_pdf`` (defined in the trainer, ``:50-132``) are rebound
to the CUDA implementations of this package — see ``patch_trainer``.
"""
import json
import logging
import os
import sys
from collections import OrderedDict

import torch
import torch.nn as nn

from ..camera_model import (PinholeModelRotNoiseLearning10kRayoRayd,
                            PinholeModelRotNoiseLearning10kRayoRaydDistortion)
from ..custom_optim import CustomAdamOptimizer
from .ddp_model import NerfNetWithAutoExpo

logger = logging.getLogger(__package__)

HOT_PATH_NAMES = ("intersect_sphere", "perturb_samples", "sample_pdf", "render_ray_from_camera",
                  "render_single_image", "proj_ray_dist_loss_single", "preprocess_match")


class SingleProcessWrapper(nn.Module):
    """Stand-in for ``DistributedDataParallel`` at world size 1: forwards to ``module`` and keeps the
    ``module.`` prefix in ``state_dict()`` keys (create_nerf.py:54-57, checkpoints at :108-110)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def patch_trainer(namespace):
    """Rebind the hot-path functions that the reference DEFINES INSIDE its trainer module
    (nerfplusplus/ddp_train_nerf.py:50-132,135-256) and the ones it star-imports (:25-26) to this
    package's CUDA implementations.  ``namespace`` is the trainer module or its ``__dict__``.
    Returns the list of names that were rebound."""
    from . import ddp_train_nerf as impl
    from .nerf_sample_ray_split import render_ray_from_camera
    from .. import ray_dist_loss
    ns = namespace if isinstance(namespace, dict) else vars(namespace)
    table = {"intersect_sphere": impl.intersect_sphere, "perturb_samples": impl.perturb_samples,
             "sample_pdf": impl.sample_pdf, "render_single_image": impl.render_single_image,
             "render_ray_from_camera": render_ray_from_camera,
             "proj_ray_dist_loss_single": ray_dist_loss.proj_ray_dist_loss_single,
             "preprocess_match": ray_dist_loss.preprocess_match}
    done = []
    for name, fn in table.items():
        if name in ns and ns[name] is not fn:
            ns[name] = fn
            done.append(name)
    return done


def _patch_calling_trainer():
    """create_nerf() is the first hot-path call the reference trainer makes (ddp_train_nerf.py:359), in every
    spawned process: rebind the trainer-local samplers in the caller's module globals there."""
    if os.environ.get("SCNERF_PATCH_TRAINER", "1") == "0":
        return []
    f = sys._getframe(2)
    while f is not None:
        g = f.f_globals
        if "intersect_sphere" in g and "ddp_train_nerf" in g and g.get("__name__") != __name__:
            return patch_trainer(g)
        f = f.f_back
    return []


def _wrap(net, rank):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        return DDP(net, device_ids=[rank], output_device=rank, find_unused_parameters=True)
    return SingleProcessWrapper(net)


def _sync_camera_grads(camera_model):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    if os.environ.get("SCNERF_SYNC_CAMERA", "1") == "0":
        return
    world = dist.get_world_size()

    def hook(p):
        dist.all_reduce(p.grad)
        p.grad.div_(world)
    for p in camera_model.parameters():
        if p.requires_grad or p.is_leaf:
            p.register_post_accumulate_grad_hook(hook)


def path2iter(path):
    tmp = os.path.basename(path)[:-4]
    return int(tmp[tmp.rfind('_') + 1:])


def create_nerf(rank, args, camera_info):
    """nerfplusplus/create_nerf.py:14-154."""
    patched = _patch_calling_trainer()
    if patched:
        logger.info("scnerf_b200: rebound trainer functions %s", patched)
    torch.manual_seed(777)                 # identical initial weights on every process (:18)
    torch.cuda.set_device(rank)

    camera_model, H, W = None, None, None
    if args.use_camera:
        intrinsics, extrinsics = camera_info["intrinsics"], camera_info["extrinsics"]
        H, W = camera_info["H"], camera_info["W"]
        if args.camera_model == "pinhole_rot_noise_10k_rayo_rayd":
            camera_model = PinholeModelRotNoiseLearning10kRayoRayd(intrinsics, extrinsics, args, H, W).to(rank)
        else:
            camera_model = PinholeModelRotNoiseLearning10kRayoRaydDistortion(
                intrinsics, extrinsics, args, H, W, camera_info["k"]).to(rank)
        _sync_camera_grads(camera_model)

    models = OrderedDict()
    models['cascade_level'] = args.cascade_level
    models['cascade_samples'] = [int(x.strip()) for x in args.cascade_samples.split(',')]

    parameters = []
    for m in range(models['cascade_level']):
        img_names = None
        if args.optim_autoexpo:
            with open(os.path.join(args.basedir, args.expname, 'train_images.json')) as file:
                img_names = json.load(file)
        net = NerfNetWithAutoExpo(args, optim_autoexpo=args.optim_autoexpo, img_names=img_names).to(rank)
        net = _wrap(net, rank)
        parameters = [*parameters, *net.parameters()]
        models['net_{}'.format(m)] = net
    if camera_model is not None:
        parameters = [*parameters, *camera_model.parameters()]

    if getattr(args, "use_custom_optim", False):
        optim = CustomAdamOptimizer(params=nn.ParameterList(parameters), lr=args.lrate, betas=(0.9, 0.999),
                                    weight_decay=args.non_linear_weight_decay, H=H, W=W, args=args)
    else:
        optim = torch.optim.Adam(nn.ParameterList(parameters), lr=args.lrate)
    models["optim"] = optim

    start = -1
    ckpt_path = getattr(args, "ckpt_path", None)
    if ckpt_path is not None and os.path.isfile(ckpt_path):
        ckpts = [ckpt_path]
    else:
        d = os.path.join(args.basedir, args.expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith('.pth')]
    ckpts = sorted(ckpts, key=path2iter)
    logger.info('Found ckpts: {}'.format(ckpts))
    if len(ckpts) > 0 and not args.no_reload:
        fpath = ckpts[-1]
        logger.info('Reloading from: {}'.format(fpath))
        start = path2iter(fpath)
        to_load = torch.load(fpath, map_location={'cuda:%d' % 0: 'cuda:%d' % rank}, weights_only=False)
        for m in range(models['cascade_level']):
            name = 'net_{}'.format(m)
            models[name].load_state_dict(to_load[name])
        model_dict = models["optim"].state_dict()
        model_dict["state"].update(to_load["optim"]["state"])
        models["optim"].load_state_dict(model_dict)
        if getattr(args, "load_camera", False):
            assert not args.load_test
            keep = {k: v for k, v in to_load["camera_model"].items()
                    if k not in ("extrinsics_noise", "extrinsics_initial")}
            origin = camera_model.state_dict()
            origin.update(keep)
            camera_model.load_state_dict(origin)
        if getattr(args, "load_test", False):
            assert not args.load_camera
            origin = camera_model.state_dict()
            origin.update(to_load["camera_model"])
            camera_model.load_state_dict(origin)

    if not getattr(args, "load_test", False):                  # camera curriculum (:131-152)
        has = lambda *names: all(hasattr(camera_model, n) for n in names)   # noqa: E731
        if start < args.add_ie and args.use_camera and has("intrinsics_noise", "extrinsics_noise"):
            camera_model.intrinsics_noise.requires_grad_(False)
            camera_model.extrinsics_noise.requires_grad_(False)
            logger.info("Deactivated learnable intrinsic and extrinsic")
        if start < args.add_radial and args.use_camera and has("distortion_noise"):
            camera_model.distortion_noise.requires_grad_(False)
            logger.info("Deactivated learnable radial distortion")
        if start < args.add_od and args.use_camera and has("ray_o_noise", "ray_d_noise"):
            camera_model.ray_o_noise.requires_grad_(False)
            camera_model.ray_d_noise.requires_grad_(False)
            logger.info("Deactivated learnable ray offset and direction noise")
    return start, models, camera_model
