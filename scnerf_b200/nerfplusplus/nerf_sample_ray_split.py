"""nerfplusplus/nerf_sample_ray_split.py:196-258 — render_ray_from_camera."""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _pp_args(camera_model, camera_idx, sel, extrinsic, keep):
    cam = camera_model.c_struct()
    a = _lib.PPRaygenArgs()
    keep.append(cam)
    a.cam = C.pointer(cam)
    if hasattr(camera_model, "distortion_noise"):
        a.distortion_initial = _lib.ptr(camera_model.distortion_initial.detach())
        a.distortion_noise = _lib.ptr(camera_model.distortion_noise.detach())
        a.distortion_noise_scale = float(camera_model.args.distortion_noise_scale)
    a.select_inds = _lib.ptr(sel)
    if camera_idx is not None:
        a.camera_idx = int(camera_idx)
    else:
        a.camera_idx = -1
        keep.append(extrinsic)
        a.extrinsic = _lib.ptr(extrinsic)
    a.N = sel.numel()
    return a


class _PPRayGen(torch.autograd.Function):
    NAMES = ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise", "distortion_noise")

    @staticmethod
    def forward(ctx, camera_model, camera_idx, sel, extrinsic, *learnables):
        lib = _lib.load()
        dev = camera_model.intrinsics_initial.device
        keep = [sel]
        a = _pp_args(camera_model, camera_idx, sel, extrinsic, keep)
        N = sel.numel()
        o = torch.empty(N, 3, device=dev, dtype=torch.float32)
        d = torch.empty(N, 3, device=dev, dtype=torch.float32)
        depth = torch.empty(N, device=dev, dtype=torch.float32)
        _lib.check(lib.scnerf_pp_raygen_fwd(C.byref(a), _lib.ptr(o), _lib.ptr(d), _lib.ptr(depth), _lib.stream()),
                   "pp_raygen_fwd")
        ctx.cm, ctx.a, ctx.keep = camera_model, a, keep
        ctx.mark_non_differentiable(depth)
        return o, d, depth

    @staticmethod
    def backward(ctx, g_o, g_d, _g_depth):
        lib = _lib.load()
        cm = ctx.cm
        N = ctx.a.N
        dev = cm.intrinsics_initial.device
        g = _lib.CameraGrads()
        outs, g_dist = [], None
        for i, name in enumerate(_PPRayGen.NAMES):
            p = getattr(cm, name, None)
            if p is not None and ctx.needs_input_grad[4 + i]:
                t = torch.zeros_like(p, dtype=torch.float32)
                if name == "distortion_noise":
                    g_dist = t
                else:
                    setattr(g, name, _lib.ptr(t))
                outs.append(t)
            else:
                outs.append(None)
        g_o = _lib.f32(g_o) if g_o is not None else torch.zeros(N, 3, device=dev)
        g_d = _lib.f32(g_d) if g_d is not None else torch.zeros(N, 3, device=dev)
        _lib.check(lib.scnerf_pp_raygen_bwd(C.byref(ctx.a), _lib.ptr(g_o), _lib.ptr(g_d), C.byref(g),
                                            _lib.ptr(g_dist), _lib.stream()), "pp_raygen_bwd")
        return (None, None, None, None, *outs)


def render_ray_from_camera(camera_model, camera_idx, select_inds, rank, extrinsic=None):
    """-> (rays_o[N,3], rays_d[N,3], depth[N]) on the camera model's device.  ``select_inds``: flat pixel
    indices (numpy or tensor); exactly as the reference, ``camera_idx=None`` needs ``extrinsic`` (4x4 numpy)."""
    dev = camera_model.intrinsics_initial.device
    if camera_idx is None:
        assert extrinsic is not None
        extrinsic = torch.as_tensor(np.asarray(extrinsic), dtype=torch.float32, device=dev).contiguous()
    if isinstance(select_inds, torch.Tensor):
        sel = select_inds.detach().to(device=dev, dtype=torch.int64).contiguous()
    else:
        sel = torch.as_tensor(np.asarray(select_inds), dtype=torch.int64, device=dev).contiguous()
    learn = [getattr(camera_model, n, None) for n in _PPRayGen.NAMES]
    return _PPRayGen.apply(camera_model, camera_idx, sel, extrinsic, *learn)
