"""Camera parameterisation helpers with the reference's names (model/camera_utils.py).

These small differentiable helpers exist for API parity (PRD loss, logging, checkpoints call
them on [n,9]/[4] tensors).  The ray-generation hot path does NOT go through them: the CUDA
kernels in csrc/raygen.cuh recompute K^-1 and the Gram-Schmidt rotation per ray in registers.
"""
import numpy as np
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


# ---- dataset / logging helpers the reference's loaders import from here (model/camera_utils.py:11-75,140-182) ----
def make_rand_axis(batch_size):
    """Random unit axes [B,3] (numpy; pose-noise injection in the data loaders, load_llff.py:329)."""
    v = np.random.rand(batch_size, 3) - 0.5
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def R_axis_angle(axis, angle):
    """Rodrigues rotation matrices [B,3,3] from unit axes [B,3] and angles [B,1] (numpy)."""
    axis, angle = np.asarray(axis, dtype=np.float64), np.asarray(angle, dtype=np.float64).reshape(-1, 1, 1)
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    zero = np.zeros_like(x)
    Kx = np.stack([np.stack([zero, -z, y], -1), np.stack([z, zero, -x], -1), np.stack([-y, x, zero], -1)], 1)
    outer = axis[:, :, None] * axis[:, None, :]
    eye = np.eye(3)[None]
    return np.cos(angle) * eye + np.sin(angle) * Kx + (1.0 - np.cos(angle)) * outer


def _to_hwc_numpy(array):
    if isinstance(array, torch.Tensor):
        array = array.detach().cpu()
        if array.dim() > 3 and array.shape[2] != 3:
            array = array.permute(1, 2, 0)
        array = array.numpy()
    return np.asarray(array)


def to_pil(array):
    from PIL import Image
    return Image.fromarray(np.uint8(_to_hwc_numpy(array) * 255))


def to_pil_normalize(array):
    from PIL import Image
    a = _to_hwc_numpy(array)
    return Image.fromarray(np.uint8((a - a.min()) / (a.max() - a.min()) * 255))


def rot_from_angle(euler):
    """[B,3] Euler angles (x, y, z) -> R = Rz(z)^T... as the reference composes it: bmm(bmm(RZ, RY), RX) with each
    factor written column-wise (model/camera_utils.py:140-175)."""
    ax, ay, az = euler[:, 0], euler[:, 1], euler[:, 2]
    o, l = torch.zeros_like(ax), torch.ones_like(ax)
    cols = lambda *c: torch.stack([torch.stack(list(v), -1) for v in c], -1)      # noqa: E731  (each arg = one column)
    RX = cols((l, o, o), (o, torch.cos(ax), -torch.sin(ax)), (o, torch.sin(ax), torch.cos(ax)))
    RY = cols((torch.cos(ay), o, torch.sin(ay)), (o, l, o), (-torch.sin(ay), o, torch.cos(ay)))
    RZ = cols((torch.cos(az), -torch.sin(az), o), (torch.sin(az), torch.cos(az), o), (o, o, l))
    return torch.bmm(torch.bmm(RZ, RY), RX)


def angle_from_rot(R):
    """Inverse of rot_from_angle (model/camera_utils.py:177-181)."""
    x = -torch.atan2(R[:, 2, 1], R[:, 2, 2])
    y = -torch.atan2(-R[:, 2, 0], torch.sqrt(R[:, 2, 1] ** 2 + R[:, 2, 2] ** 2))
    z = -torch.atan2(R[:, 1, 0], R[:, 0, 0])
    return torch.stack([x, y, z], dim=1)
