"""model/ray_dist_loss.py — the projected ray distance (PRD) loss of SCNeRF.

``proj_ray_dist_loss_single`` keeps the reference's signature, mode / camera-model branches, asserts and
return convention; the per-match computation (closest points of the two rays, projection into the other
image, chirality and threshold masks, the two masked means, and the whole backward) is two CUDA kernels
(csrc/prd_loss.cuh) instead of ~60 eager launches with boolean-mask indexing."""
import numpy as np
import torch

from . import _lib


def preprocess_match(match_result):
    """model/ray_dist_loss.py:6-19."""
    match_result = match_result[0]
    kps0, kps1, matches = match_result["kps0"], match_result["kps1"], match_result["matches"]
    if len(matches) == 0:
        return None, None
    kps0 = torch.stack([kps0[m[0]] for m in matches])
    kps1 = torch.stack([kps1[m[1]] for m in matches])
    return torch.stack([kps0, kps1])


class _PRDLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r0o, r0d, r1o, r1d, K4, E2, kps0, kps1, eps, threshold, train):
        lib = _lib.load()
        t = [_lib.f32(x).reshape(-1, 3) for x in (r0o, r0d, r1o, r1d)]
        N = t[0].shape[0]
        K4c, E2c = _lib.f32(K4), _lib.f32(E2)
        k0, k1 = _lib.f32(kps0).reshape(-1, 2), _lib.f32(kps1).reshape(-1, 2)
        dev = t[0].device
        acc = torch.empty(5, device=dev, dtype=torch.float32)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        nm = torch.empty(1, device=dev, dtype=torch.float32)
        _lib.check(lib.scnerf_prd_loss_fwd(*[_lib.ptr(x) for x in t], _lib.ptr(k0), _lib.ptr(k1), _lib.ptr(K4c),
                                           _lib.ptr(E2c), float(eps), float(threshold), int(train), N, _lib.ptr(acc),
                                           _lib.ptr(loss), _lib.ptr(nm), _lib.stream()), "prd_loss_fwd")
        ctx.save_for_backward(*t, k0, k1, K4c, E2c, acc)
        ctx.cfg = (float(eps), float(threshold), N, r0o.shape)
        ctx.mark_non_differentiable(nm)
        return loss.reshape(()), nm.reshape(())

    @staticmethod
    def backward(ctx, g_loss, _g_nm):
        lib = _lib.load()
        r0o, r0d, r1o, r1d, k0, k1, K4c, E2c, acc = ctx.saved_tensors
        eps, threshold, N, shape = ctx.cfg
        g = _lib.f32(g_loss).reshape(1)
        outs = [torch.empty_like(r0o) for _ in range(4)]
        gK = torch.zeros_like(K4c) if ctx.needs_input_grad[4] else None
        gE = torch.zeros_like(E2c) if ctx.needs_input_grad[5] else None
        _lib.check(lib.scnerf_prd_loss_bwd(_lib.ptr(r0o), _lib.ptr(r0d), _lib.ptr(r1o), _lib.ptr(r1d), _lib.ptr(k0),
                                           _lib.ptr(k1), _lib.ptr(K4c), _lib.ptr(E2c), eps, threshold, N, _lib.ptr(acc),
                                           _lib.ptr(g), *[_lib.ptr(x) for x in outs], _lib.ptr(gK), _lib.ptr(gE),
                                           _lib.stream()), "prd_loss_bwd")
        return (*[x.reshape(shape) for x in outs], gK, gE, None, None, None, None, None)


class _CameraPair(torch.autograd.Function):
    """(K4, E2) of an image pair straight from the learnable camera parameters — one launch forward, one backward —
    instead of get_intrinsic() / get_extrinsic() over all cameras in eager torch (model/ray_dist_loss.py:59-65)."""

    @staticmethod
    def forward(ctx, camera_model, i0, i1, fx_sign, intr_noise, extr_noise):
        import ctypes as C
        lib = _lib.load()
        cam = camera_model.c_struct()
        dev = camera_model.intrinsics_initial.device
        K4 = torch.empty(4, device=dev, dtype=torch.float32)
        E2 = torch.empty(2, 3, 4, device=dev, dtype=torch.float32)
        _lib.check(lib.scnerf_camera_pair_fwd(C.byref(cam), int(i0), int(i1), float(fx_sign), _lib.ptr(K4), _lib.ptr(E2),
                                              _lib.stream()), "camera_pair_fwd")
        ctx.cm, ctx.idx, ctx.fx_sign = camera_model, (int(i0), int(i1)), float(fx_sign)
        return K4, E2

    @staticmethod
    def backward(ctx, gK, gE):
        import ctypes as C
        lib = _lib.load()
        cm = ctx.cm
        cam = cm.c_struct()
        g = _lib.CameraGrads()
        g_intr = torch.zeros_like(cm.intrinsics_noise) if ctx.needs_input_grad[4] else None
        g_extr = torch.zeros_like(cm.extrinsics_noise) if ctx.needs_input_grad[5] else None
        g.intrinsics_noise, g.extrinsics_noise = _lib.ptr(g_intr), _lib.ptr(g_extr)
        gK = _lib.f32(gK) if gK is not None else None
        gE = _lib.f32(gE) if gE is not None else None
        _lib.check(lib.scnerf_camera_pair_bwd(C.byref(cam), ctx.idx[0], ctx.idx[1], ctx.fx_sign, _lib.ptr(gK), _lib.ptr(gE),
                                              C.byref(g), _lib.stream()), "camera_pair_bwd")
        return None, None, None, None, g_intr, g_extr


def proj_ray_dist_loss_single(kps0_list, kps1_list, img_idx0, img_idx1, rays0, rays1, mode, device, H, W, args,
                              camera_model=None, intrinsic=None, extrinsic=None, eps=1e-10, i_map=None,
                              method="NeRF"):
    """model/ray_dist_loss.py:22-246.  -> (loss, num_matches) in train mode, (loss, None) otherwise."""
    assert mode in ["train", "val", "test"]
    assert method in ["NeRF", "NeRF++"]
    assert kps0_list[:, 0].max() < W and kps1_list[:, 0].max() < W
    assert kps0_list[:, 1].max() < H and kps1_list[:, 1].max() < H
    fused_KE = None
    if mode == "train":
        if camera_model is not None:                       # :51-65
            assert intrinsic is None
            assert extrinsic is None
            assert i_map is not None
            i0 = np.where(i_map == img_idx0)[0][0]
            i1 = np.where(i_map == img_idx1)[0][0]
            if hasattr(camera_model, "c_struct") and camera_model.intrinsics_initial.is_cuda:
                fused_KE = _CameraPair.apply(camera_model, i0, i1, -1.0 if method == "NeRF" else 1.0,
                                             camera_model.intrinsics_noise, camera_model.extrinsics_noise)
            else:
                intrinsic = camera_model.get_intrinsic().to(device)
                extrinsic = camera_model.get_extrinsic()[[i0, i1]].to(device)
        else:                                              # :67-76
            assert intrinsic is not None
            assert extrinsic is not None
            assert isinstance(intrinsic, torch.Tensor)
            assert isinstance(extrinsic, torch.Tensor)
            intrinsic = intrinsic.to(device)
            extrinsic = extrinsic[[img_idx0, img_idx1]].to(device)
    else:
        if camera_model is not None:                       # :80-86
            assert intrinsic is None
            assert extrinsic is not None
            intrinsic = camera_model.get_intrinsic().to(device)
        else:                                              # :88-94
            assert intrinsic is not None
            assert extrinsic is not None
            intrinsic = intrinsic.to(device)
        extrinsic = extrinsic[[img_idx0, img_idx1]].to(device)
    rays0_o, rays0_d = rays0
    rays1_o, rays1_d = rays1
    if fused_KE is not None:
        K4, E2 = fused_KE
    else:
        fx = -intrinsic[0][0] if method == "NeRF" else intrinsic[0][0]        # :113-118 (NeRF's flipped x axis)
        K4 = torch.stack([fx, intrinsic[1][1], intrinsic[0][2], intrinsic[1][2]]).to(torch.float32)
        E2 = extrinsic[:, :3, :4].to(torch.float32).contiguous()
    loss, n_match = _PRDLoss.apply(rays0_o, rays0_d, rays1_o, rays1_d, K4, E2, kps0_list, kps1_list, eps,
                                   args.proj_ray_dist_threshold, mode == "train")
    if mode == "train":
        return loss, n_match.item()                        # :226-231 (the reference syncs here too)
    return loss, None
