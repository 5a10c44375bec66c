"""Pixel -> world ray generation with the reference's function names and argument meaning
(NeRF/get_rays.py).  Each call is ONE CUDA kernel (csrc/raygen.cuh) forward and one backward,
through the C ABI (scnerf_raygen_fwd / scnerf_raygen_bwd); there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _pose_source(args, idx_in_camera_param, extrinsic, N, keep, n_cams=None):
    """Fill the pose fields of a RaygenArgs: per-ray idx tensor | scalar idx | [4,4] | [N,4,4]."""
    if extrinsic is not None:
        e = _lib.f32(extrinsic)
        if e.dim() == 3:
            if e.shape[0] != N:
                raise ValueError(f"extrinsic batch {e.shape[0]} != number of rays {N}")
            args.extrinsic_per_ray = 1
        keep.append(e)
        args.extrinsic = _lib.ptr(e)
        args.idx_scalar = -1
        return
    if torch.is_tensor(idx_in_camera_param) and idx_in_camera_param.dim() >= 1:
        i = _lib.i64(idx_in_camera_param).reshape(-1)
        if i.numel() != N:
            raise ValueError(f"idx_in_camera_param has {i.numel()} entries for {N} rays")
        if n_cams is not None:
            # Python indexing semantics of `camera_model.get_extrinsic()[idx]` (get_rays.py:120): negative
            # indices wrap; out-of-range ones are poisoned to NaN rays by the kernel (no host sync here)
            i = torch.where(i < 0, i + n_cams, i)
        keep.append(i)
        args.idx = _lib.ptr(i)
        args.idx_scalar = -1
    else:
        s = int(idx_in_camera_param)
        args.idx_scalar = s + n_cams if (s < 0 and n_cams is not None) else s


class _RayGen(torch.autograd.Function):
    """rays_o, rays_d = f(camera parameters; pixels, pose selector)."""

    @staticmethod
    def forward(ctx, camera_model, kps, idx, extrinsic, N, *learnables):
        lib = _lib.load()
        dev = camera_model.intrinsics_initial.device
        cam = camera_model.c_struct()
        args = _lib.RaygenArgs()
        keep = [cam]
        args.cam = C.pointer(cam)
        if kps is not None:
            # integer pixels -> int64; sub-pixel keypoints (SIFT / SuperGlue matches for the PRD loss) stay float:
            # the direction uses the float value, the ray_o / ray_d residual lookup its truncation
            # (get_rays.py:112-123 `.float()`, :134,140 `.long()`)
            if torch.is_floating_point(kps):
                k = _lib.f32(kps[:, :2])
                args.kps_f32 = _lib.ptr(k)
            else:
                k = _lib.i64(kps[:, :2])
                args.kps = _lib.ptr(k)
            keep.append(k)
        _pose_source(args, idx, extrinsic, N, keep, int(camera_model.extrinsics_initial.shape[0]))
        args.N = N
        rays_o = torch.empty(N, 3, device=dev, dtype=torch.float32)
        rays_d = torch.empty(N, 3, device=dev, dtype=torch.float32)
        _lib.check(lib.scnerf_raygen_fwd(C.byref(args), _lib.ptr(rays_o), _lib.ptr(rays_d),
                                         _lib.stream()), "raygen_fwd")
        ctx.camera_model, ctx.args, ctx.keep = camera_model, args, keep
        return rays_o, rays_d

    @staticmethod
    def backward(ctx, g_o, g_d):
        lib = _lib.load()
        cm = ctx.camera_model
        g = _lib.CameraGrads()
        outs = []
        for i, name in enumerate(cm.LEARNABLE):
            p = getattr(cm, name, None)
            if p is not None and ctx.needs_input_grad[5 + i]:
                t = torch.zeros_like(p, dtype=torch.float32)
                setattr(g, name, _lib.ptr(t))
                outs.append(t)
            else:
                outs.append(None)
        g_o = _lib.f32(g_o) if g_o is not None else torch.zeros(ctx.args.N, 3, device=g_d.device)
        g_d = _lib.f32(g_d) if g_d is not None else torch.zeros(ctx.args.N, 3, device=g_o.device)
        _lib.check(lib.scnerf_raygen_bwd(C.byref(ctx.args), _lib.ptr(g_o), _lib.ptr(g_d), C.byref(g),
                                         _lib.stream()), "raygen_bwd")
        return (None, None, None, None, None, *outs)


def _run_camera(H, W, camera_model, kps_list, idx_in_camera_param, extrinsic, N):
    if (camera_model.H, camera_model.W) != (H, W):
        raise ValueError("H, W differ from the camera model's image size")
    return _RayGen.apply(camera_model, kps_list, idx_in_camera_param, extrinsic, N,
                         *camera_model.learnable_tensors())


def get_rays_kps_use_camera(H, W, camera_model, kps_list, idx_in_camera_param=None, extrinsic=None):
    """NeRF/get_rays.py:93-148.  ``kps_list`` is [N,2] (x, y); exactly one of
    ``idx_in_camera_param`` (int | 0-d | [N] tensor) and ``extrinsic`` ([4,4] | [N,4,4])."""
    assert kps_list[:, 0].max() < W
    assert kps_list[:, 1].max() < H
    assert (idx_in_camera_param is None) != (extrinsic is None)
    return _run_camera(H, W, camera_model, kps_list, idx_in_camera_param, extrinsic,
                       int(kps_list.shape[0]))


def get_rays_full_image_use_camera(H, W, camera_model, idx_in_camera_param=None, extrinsic=None):
    """NeRF/get_rays.py:26-72: every pixel, row-major, [H*W,3] each.  With ``extrinsic=None`` the
    reference reads the origin from the rotation block (latent bug, SURVEY.md §7.7, never hit);
    here the camera's own pose is used."""
    assert (idx_in_camera_param is None) != (extrinsic is None)
    return _run_camera(H, W, camera_model, None, idx_in_camera_param, extrinsic, H * W)


def _run_pinhole(H, W, focal, extrinsic, kps_list, N):
    lib = _lib.load()
    assert extrinsic.dim() == 2
    e = _lib.f32(extrinsic)
    if e.shape[0] == 3:        # [3,4] poses (run_nerf.py passes c2w[:3,:4])
        e = torch.cat([e, e.new_tensor([[0., 0., 0., 1.]])], 0).contiguous()
    args = _lib.RaygenArgs()
    args.focal, args.H, args.W, args.N = float(focal), int(H), int(W), N
    args.extrinsic, args.idx_scalar = _lib.ptr(e), -1
    k = None
    if kps_list is not None:
        k = _lib.i64(kps_list[:, :2])              # get_rays.py:82 truncates first: kps_list.long()
        args.kps = _lib.ptr(k)
    rays_o = torch.empty(N, 3, device=e.device, dtype=torch.float32)
    rays_d = torch.empty(N, 3, device=e.device, dtype=torch.float32)
    _lib.check(lib.scnerf_raygen_fwd(C.byref(args), _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.stream()),
               "raygen_fwd(pinhole)")
    return rays_o, rays_d


def get_rays_kps_no_camera(H, W, focal, extrinsic, kps_list):
    """NeRF/get_rays.py:75-90 (fixed pinhole; kps rows are (x, y[, 1]))."""
    assert kps_list[:, 0].max() < W
    assert kps_list[:, 1].max() < H
    return _run_pinhole(H, W, focal, extrinsic, kps_list, int(kps_list.shape[0]))


def get_rays_full_image_no_camera(H, W, focal, extrinsic):
    """NeRF/get_rays.py:5-23: [H,W,3] each."""
    o, d = _run_pinhole(H, W, focal, extrinsic, None, H * W)
    return o.reshape(H, W, 3), d.reshape(H, W, 3)


def get_rays_np(H, W, focal, extrinsic):
    """NeRF/get_rays.py:151-165: host-side numpy pinhole rays (dataset preparation, never on the
    device path)."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -np.ones_like(i)], -1)
    rays_d = (dirs[..., None, :] * extrinsic[:3, :3]).sum(-1)
    rays_o = np.broadcast_to(extrinsic[:3, -1], rays_d.shape)
    return rays_o, rays_d
