"""Golden vectors for the PRD loss by RUNNING THE REFERENCE's proj_ray_dist_loss_single
(model/ray_dist_loss.py:22-246) on synthetic matches.  Build container only:  python tests/golden/make_golden_prd.py

Matches: random pixels of image i, a random depth along their rays gives 3-D points, which are
projected into image j with the reference's own projection (rounded to integer pixels = ~0.3 px noise); a few
matches are corrupted so the threshold and chirality masks are exercised."""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path[:0] = [ROOT, REF + "/NeRF", REF, REF + "/model"]
sys.modules.setdefault("imageio", mock.MagicMock())
from scnerf_b200 import synth  # noqa: E402
import get_rays as ref_get_rays  # noqa: E402
import run_nerf_helpers  # noqa: E402,F401
torch.autograd.set_detect_anomaly(False)
from camera_dict import camera_dict  # noqa: E402
from model.ray_dist_loss import proj_ray_dist_loss_single  # noqa: E402

H, W, NCAM = synth.FERN_H, synth.FERN_W, synth.FERN_NCAM
T = torch.from_numpy


def make_camera(seed, requires_grad):
    args = synth.camera_args()
    cam = camera_dict[args.camera_model](intrinsics=T(synth.intrinsic_init()), extrinsics=list(synth.camera_poses(seed)),
                                         args=args, H=H, W=W)
    with torch.no_grad():
        for k, v in synth.camera_noise_state(seed).items():
            getattr(cam, k).copy_(T(v))
    for k in ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise"):
        getattr(cam, k).requires_grad_(requires_grad)
    return cam


def synth_matches(cam, i, j, N, seed):
    rng = np.random.default_rng(seed)
    kps0 = np.stack([rng.integers(20, W - 20, 4 * N), rng.integers(20, H - 20, 4 * N)], -1).astype(np.int64)
    with torch.no_grad():
        o0, d0 = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps0), idx_in_camera_param=i)
        K, E = cam.get_intrinsic(), cam.get_extrinsic()
        P = o0 + T(rng.uniform(2.0, 6.0, (4 * N, 1)).astype(np.float32)) * d0
        q = (P - E[j, :3, 3]) @ E[j, :3, :3]                     # R^T (P - t)
        u = (-K[0, 0] * q[:, 0] + K[0, 2] * q[:, 2]) / q[:, 2]
        v = (K[1, 1] * q[:, 1] + K[1, 2] * q[:, 2]) / q[:, 2]
    kps1 = np.stack([np.rint(u.numpy()), np.rint(v.numpy())], -1).astype(np.int64)
    ok = (kps1[:, 0] >= 0) & (kps1[:, 0] < W) & (kps1[:, 1] >= 0) & (kps1[:, 1] < H)
    kps0, kps1 = kps0[ok][:N], kps1[ok][:N]
    assert len(kps0) == N, len(kps0)
    kps1[:6] = np.stack([rng.integers(0, W, 6), rng.integers(0, H, 6)], -1)      # outliers: over the threshold
    return kps0, kps1


args = types.SimpleNamespace(proj_ray_dist_threshold=5.0)
out = {}
i, j, N = 2, 5, 256
cam = make_camera(7, True)
kps0, kps1 = synth_matches(cam, i, j, N, 0)
i_map = np.arange(NCAM)
# ---- train mode, learnable camera: gradients reach every camera parameter
r0 = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps0), idx_in_camera_param=i)
r1 = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps1), idx_in_camera_param=j)
loss, n_match = proj_ray_dist_loss_single(T(kps0), T(kps1), i, j, r0, r1, "train", "cpu", H, W, args, camera_model=cam,
                                          i_map=i_map, method="NeRF")
loss.backward()
out.update(kps0=kps0, kps1=kps1, i=i, j=j, train_loss=loss.detach(), train_n_match=n_match)
for k in ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise"):
    out["train_g_" + k] = getattr(cam, k).grad
# ---- train mode without a camera model: fixed K and poses, gradient w.r.t. the rays only
cam2 = make_camera(7, False)
with torch.no_grad():
    K, E = cam2.get_intrinsic(), cam2.get_extrinsic()
    q0 = ref_get_rays.get_rays_kps_use_camera(H, W, cam2, T(kps0), idx_in_camera_param=i)
    q1 = ref_get_rays.get_rays_kps_use_camera(H, W, cam2, T(kps1), idx_in_camera_param=j)
rays = [t.clone().requires_grad_(True) for t in (*q0, *q1)]
loss, n_match = proj_ray_dist_loss_single(T(kps0), T(kps1), i, j, (rays[0], rays[1]), (rays[2], rays[3]), "train", "cpu",
                                          H, W, args, intrinsic=K, extrinsic=E, method="NeRF")
loss.backward()
out.update(nocam_loss=loss.detach(), nocam_n_match=n_match, K=K, E=E)
for name, t in zip(("o0", "d0", "o1", "d1"), rays):
    out["rays_" + name] = t.detach()
    out["nocam_g_" + name] = t.grad
# ---- val mode (clamped errors), NeRF++ convention (fx not negated) for the branch coverage
with torch.no_grad():
    loss_v, none = proj_ray_dist_loss_single(T(kps0), T(kps1), i, j, q0, q1, "val", "cpu", H, W, args, camera_model=cam2,
                                             extrinsic=E, method="NeRF")
    loss_pp, _ = proj_ray_dist_loss_single(T(kps0), T(kps1), i, j, q0, q1, "val", "cpu", H, W, args, intrinsic=K,
                                           extrinsic=E, method="NeRF++")
assert none is None
out.update(val_loss=loss_v, val_loss_pp=loss_pp)
np.savez_compressed(os.path.join(HERE, "prd_loss.npz"), **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                                           for k, v in out.items()})
print("prd_loss:", os.path.getsize(os.path.join(HERE, "prd_loss.npz")) // 1024, "KiB; train loss", float(out["train_loss"]),
      "n_match", out["train_n_match"], "val", float(loss_v), float(loss_pp))
