"""Learnable camera modules with the reference's class names, parameter names, shapes and
state_dict keys (model/camera_model.py:12-312), so checkpoints and the trainers' curriculum
(`hasattr(camera_model, "ray_o_noise")`, `.requires_grad_()`) work unchanged.

The modules only OWN the parameters.  Ray generation reads them through ``c_struct()`` inside
the CUDA kernels (csrc/raygen.cuh); the matrix getters below are API-parity helpers.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .camera_utils import *  # noqa: F401,F403  (the reference re-exports camera_utils here: model/camera_model.py:9)
from .camera_utils import (get_44_rotation_matrix_from_33_rotation_matrix, intrinsic_param_to_K,
                           ortho2rotation, rotation2orth, to_pil_normalize)


class CameraModel(nn.Module):
    """model/camera_model.py:12-52."""

    def __init__(self, intrinsics, extrinsics, args, H, W):
        super().__init__()
        self.args = args
        self.H, self.W = H, W
        self.model_name = args.camera_model
        self.ray_o_noise_scale = args.ray_o_noise_scale
        self.ray_d_noise_scale = args.ray_d_noise_scale
        self.extrinsics_noise_scale = args.extrinsics_noise_scale
        self.intrinsics_noise_scale = args.intrinsics_noise_scale
        self.multiplicative_noise = bool(getattr(args, "multiplicative_noise", False))

    # ---- parameter construction shared by both concrete classes --------------------------------
    def _register_common(self, intrinsics, extrinsics):
        intrinsics = torch.as_tensor(np.asarray(intrinsics), dtype=torch.float32)
        ext = torch.from_numpy(np.stack([np.asarray(e) for e in extrinsics])).float()
        p4 = torch.stack([intrinsics[0, 0], intrinsics[1, 1], intrinsics[0, 2], intrinsics[1, 2]])
        e9 = torch.cat([rotation2orth(ext[:, :3, :3]), ext[:, :3, 3]], dim=-1)
        self.register_parameter("intrinsics_initial", nn.Parameter(p4, requires_grad=False))
        self.register_parameter("extrinsics_initial", nn.Parameter(e9, requires_grad=False))
        return e9

    def _grid_shape(self):
        g = self.args.grid_size
        return (self.H // g, self.W // g, 3)

    # ---- reference API ---------------------------------------------------------------------------
    def _field(self, grid, scale):
        up = F.interpolate(grid.permute(2, 0, 1)[None], (self.H, self.W), mode="bilinear",
                           align_corners=False)
        return up.permute(0, 2, 3, 1).reshape(-1, 3) * scale

    def get_ray_d_noise(self):
        """[H*W,3] upsampled residual field (model/camera_model.py:24-34). API parity only."""
        return self._field(self.ray_d_noise, self.ray_d_noise_scale)

    def get_ray_o_noise(self):
        """model/camera_model.py:36-46."""
        return self._field(self.ray_o_noise, self.ray_o_noise_scale)

    def _intrinsic_params(self):
        if self.multiplicative_noise:
            return self.intrinsics_initial + (
                self.intrinsics_noise * self.intrinsics_noise_scale * self.intrinsics_initial)
        return self.intrinsics_initial + self.intrinsics_noise * self.intrinsics_noise_scale

    def get_intrinsic(self):
        """4x4 K (model/camera_model.py:166-177)."""
        return intrinsic_param_to_K(self._intrinsic_params())

    def _extrinsic_of(self, init, noise):
        R = ortho2rotation(init[:, :6] + self.extrinsics_noise_scale * noise[:, :6])
        E = get_44_rotation_matrix_from_33_rotation_matrix(R)
        t = init[:, 6:] + self.extrinsics_noise_scale * noise[:, 6:]
        return torch.cat([torch.cat([E[:, :3, :3], t[:, :, None]], 2), E[:, 3:, :]], 1)

    def get_extrinsic(self):
        """[n,4,4] c2w (model/camera_model.py:179-190)."""
        return self._extrinsic_of(self.extrinsics_initial, self.extrinsics_noise)

    def forward(self, idx):
        """(K, c2w[idx]) (model/camera_model.py:192-206)."""
        E = self._extrinsic_of(self.extrinsics_initial[idx, None], self.extrinsics_noise[idx, None])
        return self.get_intrinsic(), E.squeeze()

    def log_noises(self, gt_intrinsic, gt_extrinsic):
        """Scalars / images the trainers push to wandb (model/camera_model.py:54-117): same keys.  Logging only."""
        K = self.get_intrinsic()
        scal = {"camera/intrinsic_noise_mean": K.abs().mean(), "camera/intrinsic_noise_std": K.abs().mean()}
        for name, (r, c) in (("fx", (0, 0)), ("fy", (1, 1)), ("cx", (0, 2)), ("cy", (1, 2))):
            scal["camera/" + name] = K[r][c]
            scal["camera/" + name + "_err"] = (K[r][c] - gt_intrinsic[r][c]).abs()
        imgs = {}
        if hasattr(self, "extrinsics_noise"):
            E = self.get_extrinsic()
            scal["camera/extrinsic_noise_mean"] = E.abs().mean()
            scal["camera/extrinsic_noise_std"] = E.abs().std()
            scal["camera/extrinisic_err"] = (E - gt_extrinsic).abs().mean()
        for tag, getter in (("ray_o_noise", self.get_ray_o_noise), ("ray_d_noise", self.get_ray_d_noise)):
            if hasattr(self, tag):
                f = getter()
                scal[f"camera/{tag}_mean"], scal[f"camera/{tag}_std"] = f.abs().mean(), f.abs().std()
                imgs[f"camera/{tag}"] = to_pil_normalize(f.reshape(self.H, self.W, 3))
        if hasattr(self, "distortion_noise"):
            scal["camera/k1"], scal["camera/k2"] = self.get_distortion()
        return scal, imgs

    # ---- C-ABI view ------------------------------------------------------------------------------
    LEARNABLE = ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise")

    def c_struct(self):
        """ctypes view of the parameters for the CUDA kernels.  Keeps tensors alive on ``self``."""
        c = _lib.Camera()
        for name in ("intrinsics_initial", "extrinsics_initial") + self.LEARNABLE:
            t = getattr(self, name, None)
            setattr(c, name, _lib.ptr(t.detach()) if t is not None else None)
        c.intrinsics_noise_scale = float(self.intrinsics_noise_scale)
        c.extrinsics_noise_scale = float(self.extrinsics_noise_scale)
        c.ray_o_noise_scale = float(self.ray_o_noise_scale)
        c.ray_d_noise_scale = float(self.ray_d_noise_scale)
        c.multiplicative_noise = int(self.multiplicative_noise)
        c.n_cams = int(self.extrinsics_initial.shape[0])
        c.H, c.W = int(self.H), int(self.W)
        c.gh, c.gw = int(self.ray_o_noise.shape[0]), int(self.ray_o_noise.shape[1])
        return c

    def learnable_tensors(self):
        return [getattr(self, n) for n in self.LEARNABLE]


class PinholeModelRotNoiseLearning10kRayoRayd(CameraModel):
    """model/camera_model.py:120-206."""

    def __init__(self, intrinsics, extrinsics, args, H, W):
        super().__init__(intrinsics, extrinsics, args, H, W)
        e9 = self._register_common(intrinsics, extrinsics)
        self.register_parameter("intrinsics_noise", nn.Parameter(torch.zeros(4)))
        self.register_parameter("extrinsics_noise", nn.Parameter(torch.zeros_like(e9)))
        self.register_parameter("ray_o_noise", nn.Parameter(torch.zeros(self._grid_shape())))
        self.register_parameter("ray_d_noise", nn.Parameter(torch.zeros(self._grid_shape())))
        self.multiplicative_noise = bool(args.multiplicative_noise)


class PinholeModelRotNoiseLearning10kRayoRaydDistortion(CameraModel):
    """model/camera_model.py:209-312 (adds 2-coefficient radial distortion used by the NeRF++
    ray generator).  Unlike the reference on CPU, ray_o_noise / ray_d_noise never alias."""

    def __init__(self, intrinsics, extrinsics, args, H, W, k=None):
        super().__init__(intrinsics, extrinsics, args, H, W)
        e9 = self._register_common(intrinsics, extrinsics)
        k0 = torch.zeros(2) if k is None else torch.tensor([float(k[0]), float(k[1])])
        self.register_parameter("distortion_initial", nn.Parameter(k0, requires_grad=False))
        self.register_parameter("intrinsics_noise", nn.Parameter(torch.zeros(4)))
        self.register_parameter("extrinsics_noise", nn.Parameter(torch.zeros_like(e9)))
        self.register_parameter("ray_o_noise", nn.Parameter(torch.zeros(self._grid_shape())))
        self.register_parameter("ray_d_noise", nn.Parameter(torch.zeros(self._grid_shape())))
        self.register_parameter("distortion_noise", nn.Parameter(torch.zeros(2)))

    def get_distortion(self):
        """model/camera_model.py:310-312."""
        return self.distortion_initial + self.distortion_noise * self.args.distortion_noise_scale
