"""CPU oracle for the NeRF++ (inverted-sphere) rows of the hot path: SURVEY.md §8 a6, a14, a15.

TEST INFRASTRUCTURE ONLY (same rules as ``oracle/scnerf_oracle.py``): nothing under
``scnerf_b200/`` imports this module.

Restates, in plain PyTorch CPU tensor algebra, the reference's ``nerfplusplus/`` functions on the
path; each function cites the file:line it follows.  Random draws are arguments (the reference calls
``torch.rand`` / ``torch.rand_like`` in place), so the oracle, the CUDA path and the live reference
(seeded with ``torch.manual_seed``) can replay the same numbers.

Pinning: ``tests/golden/make_golden.py`` runs the live reference modules (``nerf_sample_ray_split``,
``ddp_train_nerf``, ``ddp_model``, ``nerf_network``) on ``scnerf_b200.synth`` inputs and commits
``tests/golden/pp_*.npz``; ``tests/test_oracle_golden.py`` replays them against this file.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import scnerf_oracle as base

TINY_NUMBER = 1e-6     # nerfplusplus/utils.py:8
HUGE_NUMBER = 1e10     # nerfplusplus/utils.py:7


# --------------------------------------------------------------------------------------
# a6: ray generation (nerfplusplus/nerf_sample_ray_split.py:196-258)
# --------------------------------------------------------------------------------------


class CameraPP(base.Camera):
    """PinholeModelRotNoiseLearning10kRayoRaydDistortion (model/camera_model.py:209-312)."""

    LEARNABLE = base.Camera.LEARNABLE + ("distortion_noise",)

    def __init__(self, K4x4, poses, args, H, W, k=None, dtype=torch.float32):
        super().__init__(K4x4, poses, args, H, W, dtype=dtype)
        self.distortion_initial = torch.tensor([0.0, 0.0] if k is None else [k[0], k[1]], dtype=dtype)
        self.distortion_noise = torch.zeros(2, dtype=dtype)

    def distortion(self):
        # camera_model.py:310-312
        return self.distortion_initial + self.distortion_noise * self.args.distortion_noise_scale


def rays_from_camera(cam, camera_idx, select_inds, extrinsic=None, radial=True):
    """render_ray_from_camera, nerf_sample_ray_split.py:196-258 -> (rays_o[N,3], rays_d[N,3], depth[N]).
    ``select_inds`` are flat pixel indices y*W+x (int64)."""
    dtype = cam.intrinsics_initial.dtype
    W, H = cam.W, cam.H
    K = cam.intrinsic()
    if camera_idx is not None:
        c2w = cam.extrinsic()[camera_idx]                                       # :208
    else:
        c2w = torch.as_tensor(extrinsic, dtype=dtype)                           # :211-212
    u = (select_inds % W).to(torch.float32) + 0.5                               # :219-221
    v = (select_inds // W).to(torch.float32) + 0.5
    pixels = torch.stack([u, v, torch.ones_like(u)], 0).to(dtype)              # [3,N]
    cx, cy = K[0, 2], K[1, 2]
    if radial and hasattr(cam, "distortion_noise"):                             # :227-232
        k0, k1 = cam.distortion()
        center = torch.stack([cx, cy]).view(2, -1)
        r2 = (pixels[:2] - center) / center
        xy = (pixels[:2] - center) * (1 + r2 ** 2 * k0 + r2 ** 4 * k1) + center
        pixels = torch.cat([xy, pixels[2:]], 0)
    Kinv = torch.zeros(3, 3, dtype=dtype)                                       # :234-241
    one = torch.tensor(1.0, dtype=dtype)
    Kinv = Kinv.index_put((torch.tensor([0, 1, 0, 1, 2]), torch.tensor([0, 1, 2, 2, 2])),
                          torch.stack([1.0 / K[0][0], 1.0 / K[1][1], -K[0][2] / K[0][0],
                                       -K[1][2] / K[1][1], one]), accumulate=True)
    rays_d = Kinv @ pixels                                                      # :243
    rays_d = c2w[:3, :3] @ rays_d                                               # :244
    rays_d = rays_d.transpose(1, 0)
    rays_o = c2w[:3, 3].view(1, 3).repeat(select_inds.shape[0], 1)              # :247
    rays_o = rays_o + cam.ray_o_field()[select_inds]                            # :249-250
    rays_d = rays_d + cam.ray_d_field()[select_inds]                            # :252-254
    rays_d = rays_d / rays_d.norm(dim=1, keepdim=True)
    depth = c2w.T[2, 3] * torch.ones(rays_o.shape[0], dtype=dtype)              # :256
    return rays_o, rays_d, depth


# --------------------------------------------------------------------------------------
# a14: sampling (nerfplusplus/ddp_train_nerf.py:50-132, 437-474)
# --------------------------------------------------------------------------------------


def intersect_sphere(ray_o, ray_d):
    """ddp_train_nerf.py:50-68."""
    d1 = -torch.sum(ray_d * ray_o, dim=-1) / torch.sum(ray_d * ray_d, dim=-1)
    p = ray_o + d1.unsqueeze(-1) * ray_d
    ray_d_cos = 1.0 / torch.norm(ray_d, dim=-1)
    p_norm_sq = torch.sum(p * p, dim=-1)
    if (p_norm_sq >= 1.0).any():
        raise Exception("Not all your cameras are bounded by the unit sphere; please make sure "
                        "the cameras are normalized properly!")
    d2 = torch.sqrt(1.0 - p_norm_sq) * ray_d_cos
    return d1 + d2


def perturb_samples(z_vals, t_rand):
    """ddp_train_nerf.py:71-80 with the ``torch.rand_like`` draw passed in."""
    mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
    upper = torch.cat([mids, z_vals[..., -1:]], dim=-1)
    lower = torch.cat([z_vals[..., 0:1], mids], dim=-1)
    return lower + (upper - lower) * t_rand


def level0_depths(fg_near, fg_far, N_samples, t_fg=None, t_bg=None):
    """ddp_train_nerf.py:437-449: fg depths between min_depth and the sphere, bg inverse depths on
    linspace(0,1); both jittered when a draw is given."""
    step = (fg_far - fg_near) / (N_samples - 1)
    fg = torch.stack([fg_near + i * step for i in range(N_samples)], dim=-1)
    bg = torch.linspace(0.0, 1.0, N_samples, dtype=fg.dtype).expand(fg.shape).clone()
    if t_fg is not None:
        fg = perturb_samples(fg, t_fg)
    if t_bg is not None:
        bg = perturb_samples(bg, t_bg)
    return fg, bg


def sample_pdf(bins, weights, N_samples, det=False, u=None):
    """ddp_train_nerf.py:83-132 (NeRF++ flavour: count-based inversion, TINY_NUMBER = 1e-6, and the
    `+ TINY_NUMBER` inside the lerp)."""
    weights = weights + TINY_NUMBER
    pdf = weights / torch.sum(weights, dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[..., 0:1]), cdf], dim=-1)
    M = weights.shape[-1]
    if det:
        u = torch.linspace(0.0, 1.0, N_samples, dtype=bins.dtype).expand(list(weights.shape[:-1]) + [N_samples])
    above = torch.sum(u.unsqueeze(-1) >= cdf[..., :M].unsqueeze(-2), dim=-1).long()
    below = torch.clamp(above - 1, min=0)
    cdf0, cdf1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    b0, b1 = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf1 - cdf0
    denom = torch.where(denom < TINY_NUMBER, torch.ones_like(denom), denom)
    t = (u - cdf0) / denom
    return b0 + t * (b1 - b0 + TINY_NUMBER), above


def level1_depths(depth, weights, N_samples, u=None, det=False):
    """ddp_train_nerf.py:451-467: bins = depth mid-points, weights[1:-1] detached, sort-merge."""
    mid = 0.5 * (depth[..., 1:] + depth[..., :-1])
    samples, _ = sample_pdf(mid, weights.detach()[..., 1:-1], N_samples, det=det, u=u)
    out, _ = torch.sort(torch.cat((depth, samples), dim=-1))
    return out


# --------------------------------------------------------------------------------------
# a15: the field (nerfplusplus/nerf_network.py, ddp_model.py)
# --------------------------------------------------------------------------------------


def embed(x, L):
    """Embedder.forward, nerf_network.py:42-60: [x, sin(2^0 x), cos(2^0 x), ...] for any input width."""
    out = [x]
    for i in range(L):
        f = float(2.0 ** i)
        out += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(out, -1)


def mlp_forward(st, x, input_ch, input_ch_views, D=8, skips=(4,)):
    """MLPNet.forward, nerf_network.py:120-142.  ``st`` uses the reference's state-dict keys
    (base_layers.i.0.*, sigma_layers.0.*, base_remap_layers.0.*, rgb_layers.{0,2}.*)."""
    lin = lambda h, k: h @ st[k + ".weight"].T + st[k + ".bias"]
    pts = x[..., :input_ch]
    base = torch.relu(lin(pts, "base_layers.0.0"))
    for i in range(D - 1):
        if i in skips:
            base = torch.cat((pts, base), dim=-1)
        base = torch.relu(lin(base, f"base_layers.{i + 1}.0"))
    sigma = torch.abs(lin(base, "sigma_layers.0"))
    remap = lin(base, "base_remap_layers.0")
    views = x[..., -input_ch_views:]
    h = torch.relu(lin(torch.cat((remap, views), dim=-1), "rgb_layers.0"))
    rgb = torch.sigmoid(lin(h, "rgb_layers.2"))
    return rgb, sigma.squeeze(-1)


def depth2pts_outside(ray_o, ray_d, depth):
    """ddp_model.py:16-45."""
    d1 = -torch.sum(ray_d * ray_o, dim=-1) / torch.sum(ray_d * ray_d, dim=-1)
    p_mid = ray_o + d1.unsqueeze(-1) * ray_d
    p_mid_norm = torch.norm(p_mid, dim=-1)
    ray_d_cos = 1.0 / torch.norm(ray_d, dim=-1)
    d2 = torch.sqrt(1.0 - p_mid_norm * p_mid_norm) * ray_d_cos
    p_sphere = ray_o + (d1 + d2).unsqueeze(-1) * ray_d
    rot_axis = torch.cross(ray_o, p_sphere, dim=-1)
    rot_axis = rot_axis / torch.norm(rot_axis, dim=-1, keepdim=True)
    phi = torch.asin(p_mid_norm)
    theta = torch.asin(p_mid_norm * depth)
    rot_angle = (phi - theta).unsqueeze(-1)
    p_new = p_sphere * torch.cos(rot_angle) + \
        torch.cross(rot_axis, p_sphere, dim=-1) * torch.sin(rot_angle) + \
        rot_axis * torch.sum(rot_axis * p_sphere, dim=-1, keepdim=True) * (1.0 - torch.cos(rot_angle))
    p_new = p_new / torch.norm(p_new, dim=-1, keepdim=True)
    pts = torch.cat((p_new, depth.unsqueeze(-1)), dim=-1)
    depth_real = 1.0 / (depth + TINY_NUMBER) * torch.cos(theta) * ray_d_cos + d1
    return pts, depth_real


def nerfnet_forward(st_fg, st_bg, ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals, L_pos=10, L_dir=4):
    """NerfNet.forward, ddp_model.py:74-143.  st_fg / st_bg: MLPNet state dicts."""
    ray_d_norm = torch.norm(ray_d, dim=-1, keepdim=True)
    viewdirs = ray_d / ray_d_norm
    S = fg_z_vals.shape[-1]
    sh = list(ray_d.shape[:-1])
    ex = lambda v, n: v.unsqueeze(-2).expand(sh + [n, 3])
    fg_pts = ex(ray_o, S) + fg_z_vals.unsqueeze(-1) * ex(ray_d, S)
    inp = torch.cat((embed(fg_pts, L_pos), embed(ex(viewdirs, S), L_dir)), dim=-1)
    fg_rgb, fg_sigma = mlp_forward(st_fg, inp, 3 + 6 * L_pos, 3 + 6 * L_dir)
    fg_dists = fg_z_vals[..., 1:] - fg_z_vals[..., :-1]
    fg_dists = ray_d_norm * torch.cat((fg_dists, fg_z_max.unsqueeze(-1) - fg_z_vals[..., -1:]), dim=-1)
    fg_alpha = 1.0 - torch.exp(-fg_sigma * fg_dists)
    T = torch.cumprod(1.0 - fg_alpha + TINY_NUMBER, dim=-1)
    bg_lambda = T[..., -1]
    T = torch.cat((torch.ones_like(T[..., 0:1]), T[..., :-1]), dim=-1)
    fg_weights = fg_alpha * T
    fg_rgb_map = torch.sum(fg_weights.unsqueeze(-1) * fg_rgb, dim=-2)
    fg_depth_map = torch.sum(fg_weights * fg_z_vals, dim=-1)

    S = bg_z_vals.shape[-1]
    bg_pts, _ = depth2pts_outside(ex(ray_o, S), ex(ray_d, S), bg_z_vals)
    inp = torch.cat((embed(bg_pts, L_pos), embed(ex(viewdirs, S), L_dir)), dim=-1)
    inp = torch.flip(inp, dims=[-2])
    bg_z = torch.flip(bg_z_vals, dims=[-1])
    bg_dists = bg_z[..., :-1] - bg_z[..., 1:]
    bg_dists = torch.cat((bg_dists, HUGE_NUMBER * torch.ones_like(bg_dists[..., 0:1])), dim=-1)
    bg_rgb, bg_sigma = mlp_forward(st_bg, inp, 4 + 8 * L_pos, 3 + 6 * L_dir)
    bg_alpha = 1.0 - torch.exp(-bg_sigma * bg_dists)
    T = torch.cumprod(1.0 - bg_alpha + TINY_NUMBER, dim=-1)[..., :-1]
    T = torch.cat((torch.ones_like(T[..., 0:1]), T), dim=-1)
    bg_weights = bg_alpha * T
    bg_rgb_map = torch.sum(bg_weights.unsqueeze(-1) * bg_rgb, dim=-2)
    bg_depth_map = torch.sum(bg_weights * bg_z, dim=-1)
    bg_rgb_map = bg_lambda.unsqueeze(-1) * bg_rgb_map
    bg_depth_map = bg_lambda * bg_depth_map
    return OrderedDict([("rgb", fg_rgb_map + bg_rgb_map), ("fg_weights", fg_weights),
                        ("bg_weights", bg_weights), ("fg_rgb", fg_rgb_map), ("fg_depth", fg_depth_map),
                        ("bg_rgb", bg_rgb_map), ("bg_depth", bg_depth_map), ("bg_lambda", bg_lambda)])


def train_step(cam, camera_idx, select_inds, target, nets, cascade_samples, rand, min_depth=1e-4,
               level1_override=None):
    """One optimisation step's forward of ddp_train_nerf.py:421-488 (no auto-exposure):
    loss = sum over cascade levels of img2mse(rgb, target).  ``nets`` = [(st_fg, st_bg), ...];
    ``rand`` = dict(t_fg, t_bg, u_fg, u_bg) of injected draws (None = deterministic).
    ``level1_override`` = (fg_depth, d(fg_depth)/d(far), bg_depth): evaluate level 1 AT these sample positions
    (still differentiable w.r.t. the sphere depth through the coefficient).  The parity tests use it to compare
    gradients at the CUDA path's own samples: inverse-CDF sampling is discontinuous in the weights, so two fp32
    implementations occasionally put a sample in neighbouring bins."""
    ray_o, ray_d, _ = rays_from_camera(cam, camera_idx, select_inds)
    loss = 0.0
    rets = []
    fg_far = intersect_sphere(ray_o, ray_d)
    fg_near = min_depth * torch.ones_like(ray_d[..., 0])
    for m, (st_fg, st_bg) in enumerate(nets):
        Ns = cascade_samples[m]
        if m == 0:
            fg_depth, bg_depth = level0_depths(fg_near, fg_far, Ns, rand.get("t_fg"), rand.get("t_bg"))
        elif level1_override is not None:
            fg1, coef1, bg1 = level1_override
            fg_depth = fg1 + coef1 * (fg_far - fg_far.detach()).unsqueeze(-1)
            bg_depth = bg1
        else:
            fg_depth = level1_depths(fg_depth, ret["fg_weights"], Ns, u=rand.get("u_fg"), det=rand.get("u_fg") is None)
            bg_depth = level1_depths(bg_depth, ret["bg_weights"], Ns, u=rand.get("u_bg"), det=rand.get("u_bg") is None)
        ret = nerfnet_forward(st_fg, st_bg, ray_o, ray_d, fg_far, fg_depth, bg_depth)
        rets.append(ret)
        loss = loss + torch.mean((ret["rgb"] - target) ** 2)          # utils.py:12-14
    return loss, rets, (fg_depth, bg_depth)
