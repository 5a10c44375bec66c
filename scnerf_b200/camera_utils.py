"""Camera parameterisation helpers with the reference's names (model/camera_utils.py).

These small differentiable helpers exist for API parity (PRD loss, logging, checkpoints call
them on [n,9]/[4] tensors).  The ray-generation hot path does NOT go through them: the CUDA
kernels in csrc/raygen.cuh recompute K^-1 and the Gram-Schmidt rotation per ray in registers.
"""
import torch


def _unit(v):
    # model/camera_utils.py:88-95
    mag = torch.sqrt((v * v).sum(1, keepdim=True)).clamp(min=1e-8)
    return v / (mag + 1e-10)


def ortho2rotation(poses):
    """[B,6] -> [B,3,3], columns (x, y, z).  model/camera_utils.py:78-133."""
    a, b = poses[:, 0:3], poses[:, 3:6]
    x = _unit(a)
    coef = (x * b).sum(1, keepdim=True) / ((x * x).sum(1, keepdim=True).clamp(min=1e-8) + 1e-10)
    y = _unit(b - coef * x)
    z = torch.cross(x, y, dim=1)
    return torch.stack([x, y, z], 2)


def rotation2orth(rot):
    """model/camera_utils.py:136-137."""
    return torch.cat([rot[:, :, 0], rot[:, :, 1]], dim=-1)


def get_44_rotation_matrix_from_33_rotation_matrix(m):
    """model/camera_utils.py:184-188."""
    out = torch.zeros((m.shape[0], 4, 4), device=m.device, dtype=m.dtype)
    out[:, :3, :3] = m
    out[:, 3, 3] = 1
    return out


def intrinsic_param_to_K(intrinsics):
    """[fx,fy,cx,cy] -> 4x4 K.  model/camera_utils.py:191-195."""
    K = torch.eye(4, device=intrinsics.device, dtype=intrinsics.dtype)
    rows = torch.tensor([0, 1, 0, 1], device=intrinsics.device)
    cols = torch.tensor([0, 1, 2, 2], device=intrinsics.device)
    return K.index_put((rows, cols), intrinsics)
