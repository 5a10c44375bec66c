"""Deterministic synthetic inputs for the SCNeRF hot path.

Pure numpy (PCG64 streams are stable across machines), so the golden-vector
generator (run next to the reference), the CPU oracle tests, the GPU parity
tests and ``bench.py`` all see byte-identical inputs without shipping weights.

Nothing here is on the product path: it only manufactures inputs.
"""
from __future__ import annotations

import types
import numpy as np

# LLFF fern at factor=8: H=378, W=504, 20 images -> 17 train views.
FERN_H, FERN_W, FERN_NCAM, FERN_FOCAL = 378, 504, 17, 407.5


def camera_args(camera_model="pinhole_rot_noise_10k_rayo_rayd", ray_o_noise_scale=1e-3,
                ray_d_noise_scale=1e-3, extrinsics_noise_scale=1.0,
                intrinsics_noise_scale=1.0, grid_size=10, multiplicative_noise=True,
                distortion_noise_scale=1e-2):
    """Flag namespace the camera classes read (NeRF/config_argparse.py:166,270-304)."""
    return types.SimpleNamespace(
        camera_model=camera_model, ray_o_noise_scale=ray_o_noise_scale,
        ray_d_noise_scale=ray_d_noise_scale, extrinsics_noise_scale=extrinsics_noise_scale,
        intrinsics_noise_scale=intrinsics_noise_scale, grid_size=grid_size,
        multiplicative_noise=multiplicative_noise,
        distortion_noise_scale=distortion_noise_scale)


def _axis_angle_matrix(axis, angle):
    axis = axis / np.linalg.norm(axis)
    x, y, z = axis
    c, s = np.cos(angle), np.sin(angle)
    C = 1.0 - c
    return np.array([[c + x * x * C, x * y * C - z * s, x * z * C + y * s],
                     [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
                     [z * x * C - y * s, z * y * C + x * s, c + z * z * C]])


def camera_poses(seed=0, n_cams=FERN_NCAM, max_deg=10.0, t_range=0.3):
    """[n,4,4] float32 camera-to-world poses: small rotations about identity,
    translations U(-t_range, t_range)^3 (post-recenter LLFF scale)."""
    rng = np.random.default_rng(seed)
    out = np.tile(np.eye(4), (n_cams, 1, 1))
    for i in range(n_cams):
        axis = rng.standard_normal(3)
        ang = np.deg2rad(rng.uniform(-max_deg, max_deg))
        out[i, :3, :3] = _axis_angle_matrix(axis, ang)
        out[i, :3, 3] = rng.uniform(-t_range, t_range, 3)
    return out.astype(np.float32)


def intrinsic_init(H=FERN_H, W=FERN_W, focal=FERN_FOCAL):
    """4x4 K as NeRF/create_nerf.py:101-108 builds it (fx=fy=focal, principal point at centre)."""
    return np.array([[focal, 0, W / 2, 0], [0, focal, H / 2, 0],
                     [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)


def camera_noise_state(seed, n_cams=FERN_NCAM, H=FERN_H, W=FERN_W, grid_size=10,
                       sigma_ie=1e-3, sigma_od=1.0, with_distortion=False):
    """Non-trivial values for the learnable camera parameters so every gradient is exercised."""
    rng = np.random.default_rng(seed + 1000)
    st = {
        "intrinsics_noise": rng.standard_normal(4) * sigma_ie,
        "extrinsics_noise": rng.standard_normal((n_cams, 9)) * sigma_ie,
        "ray_o_noise": rng.standard_normal((H // grid_size, W // grid_size, 3)) * sigma_od,
        "ray_d_noise": rng.standard_normal((H // grid_size, W // grid_size, 3)) * sigma_od,
    }
    if with_distortion:
        st["distortion_noise"] = rng.standard_normal(2) * 1e-2
    return {k: v.astype(np.float32) for k, v in st.items()}


def mlp_layer_shapes(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5,
                     skips=(4,), use_viewdirs=True):
    """(name, out, in, activation) in ``NeRF.parameters()`` order."""
    layers = [("pts_linears.0", W, input_ch, "relu")]
    for i in range(D - 1):
        fan_in = W + input_ch if i in skips else W
        layers.append((f"pts_linears.{i + 1}", W, fan_in, "relu"))
    layers.append(("views_linears.0", W // 2, input_ch_views + W, "relu"))
    if use_viewdirs:
        layers += [("feature_linear", W, W, "linear"), ("alpha_linear", 1, W, "linear"),
                   ("rgb_linear", 3, W // 2, "linear")]
    else:
        layers.append(("output_linear", output_ch, W, "linear"))
    return layers


def mlp_state(seed, bias_sigma=0.05, **kw):
    """Xavier-uniform weights (gain sqrt2 for relu layers, 1 for linear heads —
    NeRF/run_nerf_helpers.py:18-21) and small NON-zero biases (the reference
    zero-inits them; non-zero exercises the bias path)."""
    rng = np.random.default_rng(seed + 2000)
    st = {}
    for name, fo, fi, act in mlp_layer_shapes(**kw):
        gain = np.sqrt(2.0) if act == "relu" else 1.0
        a = gain * np.sqrt(6.0 / (fi + fo))
        st[name + ".weight"] = rng.uniform(-a, a, (fo, fi)).astype(np.float32)
        st[name + ".bias"] = (rng.standard_normal(fo) * bias_sigma).astype(np.float32)
    return st


def pixel_batch(seed, N, H=FERN_H, W=FERN_W, n_cams=FERN_NCAM):
    """kps[N,2] int64 as (x, y), per-ray camera index idx[N] int64, target rgb[N,3] f32."""
    rng = np.random.default_rng(seed + 3000)
    kps = np.stack([rng.integers(0, W, N), rng.integers(0, H, N)], -1).astype(np.int64)
    idx = rng.integers(0, n_cams, N).astype(np.int64)
    target = rng.uniform(0, 1, (N, 3)).astype(np.float32)
    return kps, idx, target


def reference_pytest_rand(shape):
    """What the reference draws when ``pytest=True``: ``np.random.seed(0); np.random.rand(*shape)``
    cast to float32 (NeRF/render.py:252-255, 333-336, 432-440)."""
    np.random.seed(0)
    return np.random.rand(*shape).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# NeRF++ (inverted sphere) configs — BASELINE.json configs[3], configs[4]
# ---------------------------------------------------------------------------------------------
PP_H, PP_W, PP_NCAM, PP_FOCAL = 120, 200, 8, 160.0


def pp_camera_poses(seed=0, n_cams=PP_NCAM, radius=0.5):
    """[n,4,4] float32 camera-to-world, OpenCV convention (+z forward): cameras on a jittered ring of
    radius ~0.5 looking at the origin, i.e. inside the unit sphere as `intersect_sphere` requires
    (nerfplusplus/ddp_train_nerf.py:50-68)."""
    rng = np.random.default_rng(seed + 4000)
    out = np.tile(np.eye(4), (n_cams, 1, 1))
    for i in range(n_cams):
        ang = 2 * np.pi * i / n_cams + rng.uniform(-0.1, 0.1)
        pos = np.array([radius * np.cos(ang), rng.uniform(-0.1, 0.1), radius * np.sin(ang)])
        z = -pos / np.linalg.norm(pos)                       # forward: towards the origin
        x = np.cross(np.array([0.0, 1.0, 0.0]), z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        out[i, :3, 0], out[i, :3, 1], out[i, :3, 2], out[i, :3, 3] = x, y, z, pos
    return out.astype(np.float32)


def pp_camera_args(**kw):
    kw.setdefault("camera_model", "pinhole_rot_noise_10k_rayo_rayd_dist")
    return camera_args(**kw)


def pp_mlp_layer_shapes(input_ch, input_ch_views=27, D=8, W=256, skips=(4,)):
    """(key, out, in) of nerfplusplus/nerf_network.py:68-118 MLPNet in parameters() order."""
    layers = [("base_layers.0.0", W, input_ch)]
    for i in range(D - 1):
        layers.append((f"base_layers.{i + 1}.0", W, W + input_ch if i in skips else W))
    layers += [("sigma_layers.0", 1, W), ("base_remap_layers.0", 256, W),
               ("rgb_layers.0", W // 2, 256 + input_ch_views), ("rgb_layers.2", 3, W // 2)]
    return layers


def pp_mlp_state(seed, input_ch, input_ch_views=27, bias_sigma=0.05):
    """PyTorch-default-like init (the reference keeps nn.Linear defaults, nerf_network.py:98-118):
    U(-1/sqrt(in), 1/sqrt(in)) weights; widened a little so sigma is not vanishing; small biases."""
    rng = np.random.default_rng(seed + 5000)
    st = {}
    for name, fo, fi in pp_mlp_layer_shapes(input_ch, input_ch_views):
        a = 1.7 / np.sqrt(fi)
        st[name + ".weight"] = rng.uniform(-a, a, (fo, fi)).astype(np.float32)
        st[name + ".bias"] = (rng.standard_normal(fo) * bias_sigma).astype(np.float32)
    return st


def pp_pixel_batch(seed, N, H=PP_H, W=PP_W, n_cams=PP_NCAM):
    """select_inds[N] int64 (flat y*W+x, no replacement like np.random.choice(..., replace=False),
    nerf_sample_ray_split.py:147), camera index, target rgb[N,3]."""
    rng = np.random.default_rng(seed + 6000)
    sel = rng.choice(H * W, size=N, replace=False).astype(np.int64)
    cam_idx = int(rng.integers(0, n_cams))
    target = rng.uniform(0, 1, (N, 3)).astype(np.float32)
    return sel, cam_idx, target


def adam_case(seed, n_steps=3):
    """Parameter shapes (a slice of the real grad_vars order: MLP tensors then the camera's noise tensors) and
    per-step gradients for the optimiser tests."""
    rng = np.random.default_rng(seed + 7000)
    shapes = [(32, 63), (32,), (32, 32), (1, 32), (3, 16), (4,), (17, 9), (7, 10, 3), (7, 10, 3), (2,)]
    params = [rng.standard_normal(s).astype(np.float32) * 0.1 for s in shapes]
    grads = [[(rng.standard_normal(s) * 10.0 ** rng.uniform(-4, 0)).astype(np.float32) for s in shapes]
             for _ in range(n_steps)]
    return params, grads


# ---------------------------------------------------------------------------------------------
# Round 2: sub-pixel keypoints and the composed configs[2] step (render + PRD + custom Adam)
# ---------------------------------------------------------------------------------------------
def subpixel_kps(seed, N, H=FERN_H, W=FERN_W, n_cams=FERN_NCAM):
    """kps[N,2] float32 (x, y) with fractional parts (SIFT / SuperGlue keypoints), idx[N] int64."""
    rng = np.random.default_rng(seed + 5000)
    kps = np.stack([rng.uniform(0, W - 1, N), rng.uniform(0, H - 1, N)], -1).astype(np.float32)
    kps[: N // 8] = np.floor(kps[: N // 8])          # some exactly on the pixel grid
    return kps, rng.integers(0, n_cams, N).astype(np.int64)


def c3_case():
    """Sizes / hyper-parameters of the composed configs[2] step (tests/golden/make_golden_r2.py, bench --workload c3)."""
    return dict(seed=31, N_rays=48, Nc=64, Nf=128, pair=(2, 5), n_matches=192, threshold=5.0, prd_weight=1e-4,
                lrate=5e-4, lrate_decay=250, weight_decay=0.1, global_step0=1001, n_steps=2)


def c3_matches(seed, N=None, H=FERN_H, W=FERN_W, pose_seed=None):
    """Sub-pixel matches (kps0, kps1) [N,2] float32 of the image pair ``c3_case()["pair"]``: random sub-pixel keypoints
    of image i, lifted to a random depth along their (noise-free pinhole) rays and projected into image j, plus 0.3 px of
    detector noise; the first 8 are outliers (threshold / chirality masks of the PRD loss)."""
    C = c3_case()
    N = N or C["n_matches"]
    i, j = C["pair"]
    rng = np.random.default_rng(seed + 7000)
    poses = camera_poses(C["seed"] if pose_seed is None else pose_seed).astype(np.float64)
    K = intrinsic_init(H, W).astype(np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    kps0 = np.stack([rng.uniform(30, W - 30, 4 * N), rng.uniform(30, H - 30, 4 * N)], -1)
    dirs = np.stack([(kps0[:, 0] - cx) / fx, -(kps0[:, 1] - cy) / fy, -np.ones(4 * N)], -1)
    P = poses[i, :3, 3] + rng.uniform(2.0, 6.0, (4 * N, 1)) * (dirs @ poses[i, :3, :3].T)
    q = (P - poses[j, :3, 3]) @ poses[j, :3, :3]                  # R_j^T (P - t_j)
    kps1 = np.stack([cx - fx * q[:, 0] / q[:, 2], cy + fy * q[:, 1] / q[:, 2]], -1) + rng.normal(0, 0.3, (4 * N, 2))
    ok = (kps1[:, 0] >= 0) & (kps1[:, 0] < W - 1) & (kps1[:, 1] >= 0) & (kps1[:, 1] < H - 1)
    kps0, kps1 = kps0[ok][:N], kps1[ok][:N]
    assert len(kps0) == N, len(kps0)
    kps1[:8] = np.stack([rng.uniform(0, W - 1, 8), rng.uniform(0, H - 1, 8)], -1)
    return kps0.astype(np.float32), kps1.astype(np.float32)


# ---------------------------------------------------------------------------------------------
# Seeded scenes as CUDA modules (tests, smoke(), bench.py): camera + networks loaded from the states above
# ---------------------------------------------------------------------------------------------
def build_modules(seed, device, mult=True, n_cams=FERN_NCAM):
    """NeRF/ scene: learnable camera (pinhole_rot_noise_10k_rayo_rayd) + coarse/fine 8x256 MLPs."""
    import torch
    from .camera_dict import camera_dict
    from .run_nerf_helpers import NeRF
    args = camera_args(multiplicative_noise=mult)
    cam = camera_dict[args.camera_model](intrinsics=intrinsic_init(), extrinsics=list(camera_poses(seed, n_cams)),
                                         args=args, H=FERN_H, W=FERN_W)
    with torch.no_grad():
        for k, v in camera_noise_state(seed, n_cams).items():
            getattr(cam, k).copy_(torch.from_numpy(v))
    nets = []
    for s in (seed, seed + 1):
        net = NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in mlp_state(s).items()})
        nets.append(net.to(device))
    return dict(cam=cam.to(device), coarse=nets[0], fine=nets[1], args=args)


def pp_net_args():
    return types.SimpleNamespace(max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8, netwidth=256, use_viewdirs=True)


def build_pp_modules(seed, device, levels=2, precision=None, distortion=True):
    """NeRF++ scene: ring of cameras inside the unit sphere (learnable, with radial distortion unless
    ``distortion=False``) + one NerfNet (fg + bg 8x256 MLPs) per cascade level."""
    import torch
    from .camera_dict import camera_dict
    from .nerfplusplus import NerfNet
    args = pp_camera_args() if distortion else pp_camera_args(camera_model="pinhole_rot_noise_10k_rayo_rayd")
    kw = dict(k=(-0.05, 0.01)) if distortion else {}
    cam = camera_dict[args.camera_model](intrinsics=intrinsic_init(PP_H, PP_W, PP_FOCAL), extrinsics=list(pp_camera_poses(seed)),
                                         args=args, H=PP_H, W=PP_W, **kw)
    with torch.no_grad():
        for k, v in camera_noise_state(seed, n_cams=PP_NCAM, H=PP_H, W=PP_W, with_distortion=distortion).items():
            getattr(cam, k).copy_(torch.from_numpy(v))
    nets = []
    for m in range(levels):
        net = NerfNet(pp_net_args(), precision=precision)
        net.fg_net.load_state_dict({k: torch.from_numpy(v) for k, v in pp_mlp_state(seed + 10 + 2 * m, 63).items()})
        net.bg_net.load_state_dict({k: torch.from_numpy(v) for k, v in pp_mlp_state(seed + 11 + 2 * m, 84).items()})
        nets.append(net.to(device))
    return dict(cam=cam.to(device), nets=nets, args=args)
