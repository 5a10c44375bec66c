"""CPU oracle for the SCNeRF volumetric-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``scnerf_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  The product path is the CUDA library
(``scnerf_b200/csrc``) and fails loudly when it is missing.

What it is: a restatement, in plain PyTorch CPU tensor algebra (the reference itself is
eager PyTorch, so this is the faithful "same language" oracle), of every function on the
path of SURVEY.md §8(a).  Each function cites the reference file:line it follows.  It runs
in float32 (parity target) or float64 (tolerance calibration) and is differentiable through
``torch.autograd`` exactly where the reference graph is.

Pinning: the reference has NO golden vectors for this path.  The oracle
is pinned against outputs of the reference itself, imported from ``/root/reference`` by
``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them.  ``searchsorted`` is additionally pinned
against numpy on the reference's own test grid
(NeRF/torchsearchsorted/test/test_searchsorted.py:27-44).
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# camera (model/camera_model.py, model/camera_utils.py)
# --------------------------------------------------------------------------------------


def effective_intrinsics(init4, noise4, scale, multiplicative):
    """[fx, fy, cx, cy] after the learnable perturbation.
    model/camera_model.py:166-177 (and :272-282 for the distortion variant)."""
    if multiplicative:
        return init4 + noise4 * scale * init4
    return init4 + noise4 * scale


def intrinsic_matrix(p4):
    """4x4 K with fx,fy on the diagonal and cx,cy in the third column.
    model/camera_utils.py:191-195."""
    K = torch.eye(4, dtype=p4.dtype)
    rows = torch.tensor([0, 1, 0, 1])
    cols = torch.tensor([0, 1, 2, 2])
    K = K.index_put((rows, cols), p4)
    return K


def _unit(v):
    # model/camera_utils.py:88-95: v / (clamp(|v|, 1e-8) + 1e-10)
    mag = torch.sqrt((v * v).sum(1, keepdim=True)).clamp(min=1e-8)
    return v / (mag + 1e-10)


def rot6d_to_matrix(p6):
    """Gram-Schmidt 6-D -> rotation, columns [x y z].  model/camera_utils.py:78-133."""
    a, b = p6[:, 0:3], p6[:, 3:6]
    x = _unit(a)
    # proj_u2a(x, b): ((x.b) / (clamp(x.x,1e-8)+1e-10)) x      (:112-122)
    coef = (x * b).sum(1, keepdim=True) / ((x * x).sum(1, keepdim=True).clamp(min=1e-8) + 1e-10)
    y = _unit(b - coef * x)
    z = torch.stack([x[:, 1] * y[:, 2] - x[:, 2] * y[:, 1],
                     x[:, 2] * y[:, 0] - x[:, 0] * y[:, 2],
                     x[:, 0] * y[:, 1] - x[:, 1] * y[:, 0]], 1)
    return torch.stack([x, y, z], 2)


def extrinsic_matrices(init9, noise9, scale):
    """[n,4,4] camera-to-world.  model/camera_model.py:179-190, camera_utils.py:184-188."""
    R = rot6d_to_matrix(init9[:, :6] + scale * noise9[:, :6])
    t = init9[:, 6:] + scale * noise9[:, 6:]
    n = init9.shape[0]
    top = torch.cat([R, t[:, :, None]], 2)
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=init9.dtype).expand(n, 1, 4)
    return torch.cat([top, bottom], 1)


def matrix_to_rot6d(R):
    """model/camera_utils.py:136-137: first two columns, concatenated."""
    return torch.cat([R[:, :, 0], R[:, :, 1]], -1)


def ray_noise_field(grid, H, W, scale):
    """Bilinear upsample of the coarse residual grid to H x W, flattened row-major.
    model/camera_model.py:24-46 (F.interpolate, align_corners=False)."""
    up = F.interpolate(grid.permute(2, 0, 1)[None], (H, W), mode="bilinear", align_corners=False)
    return up[0].permute(1, 2, 0).reshape(-1, 3) * scale


def bilinear_grid_lookup(grid, ys, xs, H, W, scale):
    """Closed form of ``ray_noise_field(...)[y*W+x]``: the form the
    CUDA kernel implements; verified equal to F.interpolate in tests."""
    gh, gw = grid.shape[0], grid.shape[1]

    def axis(p, n_in, n_out):
        s = torch.tensor(n_in / n_out, dtype=torch.float32)
        f = ((p.to(torch.float32) + 0.5) * s - 0.5).clamp(min=0.0)
        i0 = f.floor().long().clamp(max=n_in - 1)
        i1 = (i0 + 1).clamp(max=n_in - 1)
        lam = (f - i0.to(torch.float32)).to(grid.dtype)
        return i0, i1, lam

    y0, y1, ly = axis(ys, gh, H)
    x0, x1, lx = axis(xs, gw, W)
    lx, ly = lx[:, None], ly[:, None]
    top = grid[y0, x0] * (1 - lx) + grid[y0, x1] * lx
    bot = grid[y1, x0] * (1 - lx) + grid[y1, x1] * lx
    return (top * (1 - ly) + bot * ly) * scale


class Camera:
    """State holder mirroring PinholeModelRotNoiseLearning10kRayoRayd
    (model/camera_model.py:120-206); tensors may require grad."""

    def __init__(self, K4x4, poses, args, H, W, dtype=torch.float32):
        K4x4 = torch.as_tensor(K4x4, dtype=dtype)
        poses = torch.as_tensor(poses, dtype=dtype)
        self.H, self.W, self.args = H, W, args
        self.intrinsics_initial = torch.stack([K4x4[0, 0], K4x4[1, 1], K4x4[0, 2], K4x4[1, 2]])
        self.extrinsics_initial = torch.cat([matrix_to_rot6d(poses[:, :3, :3]), poses[:, :3, 3]], -1)
        g = args.grid_size
        self.intrinsics_noise = torch.zeros(4, dtype=dtype)
        self.extrinsics_noise = torch.zeros_like(self.extrinsics_initial)
        self.ray_o_noise = torch.zeros(H // g, W // g, 3, dtype=dtype)
        self.ray_d_noise = torch.zeros(H // g, W // g, 3, dtype=dtype)

    LEARNABLE = ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise")

    def load(self, state, requires_grad=False):
        for k, v in state.items():
            t = torch.as_tensor(v, dtype=self.intrinsics_initial.dtype).clone()
            setattr(self, k, t.requires_grad_(requires_grad))
        return self

    def learnables(self):
        return [getattr(self, k) for k in self.LEARNABLE]

    def intrinsic(self):
        return intrinsic_matrix(effective_intrinsics(
            self.intrinsics_initial, self.intrinsics_noise,
            self.args.intrinsics_noise_scale, self.args.multiplicative_noise))

    def extrinsic(self):
        return extrinsic_matrices(self.extrinsics_initial, self.extrinsics_noise,
                                  self.args.extrinsics_noise_scale)

    def ray_o_field(self):
        return ray_noise_field(self.ray_o_noise, self.H, self.W, self.args.ray_o_noise_scale)

    def ray_d_field(self):
        return ray_noise_field(self.ray_d_noise, self.H, self.W, self.args.ray_d_noise_scale)


# --------------------------------------------------------------------------------------
# ray generation (NeRF/get_rays.py)
# --------------------------------------------------------------------------------------


def rays_pixels_camera(H, W, cam: Camera, kps, idx=None, extrinsic=None):
    """NeRF/get_rays.py:93-148.  kps is (x, y), int64 pixels or float sub-pixel keypoints: the direction uses
    ``kps.float()`` (:112-123), the ray_o / ray_d residual lookup ``kps.long()`` (:134,140).  Exactly one of idx /
    extrinsic."""
    assert (idx is None) != (extrinsic is None)
    dtype = cam.intrinsics_initial.dtype
    pix = torch.stack([kps[:, 0], kps[:, 1], torch.ones_like(kps[:, 0])], -1).to(dtype)
    Kinv = torch.inverse(cam.intrinsic()[:3, :3])                       # :119
    c2w = cam.extrinsic()[idx] if extrinsic is None else extrinsic       # :120
    dirs = pix @ Kinv.T                                                 # :123
    dirs = dirs * torch.tensor([1.0, -1.0, -1.0], dtype=dtype)          # :125 OpenGL flip
    if c2w.dim() == 3:
        rays_d = (dirs[:, None, :] * c2w[:, :3, :3]).sum(-1)            # :128
        rays_o = c2w[:, :3, 3]
    else:
        rays_d = (dirs[:, None, :] * c2w[:3, :3]).sum(-1)               # :131
        rays_o = c2w[:3, 3].expand(rays_d.shape)
    kl = kps.long()
    flat = kl[:, 1] * W + kl[:, 0]
    rays_o = rays_o + cam.ray_o_field()[flat]                           # :134-138
    rays_d = rays_d + cam.ray_d_field()[flat]                           # :140-145
    rays_d = rays_d / (rays_d.norm(dim=1)[:, None] + 1e-10)             # :146
    return rays_o, rays_d


def full_image_pixels(H, W):
    """(x, y) for every pixel, row-major over (y, x) — NeRF/get_rays.py:37-44 after the .t()."""
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)


def rays_full_image_camera(H, W, cam: Camera, extrinsic):
    """NeRF/get_rays.py:26-72 with ``extrinsic`` given (every caller passes it: render.py:45-67)."""
    return rays_pixels_camera(H, W, cam, full_image_pixels(H, W), extrinsic=extrinsic)


def rays_pixels_pinhole(H, W, focal, c2w, kps):
    """NeRF/get_rays.py:75-90 (fixed pinhole, integer pixel coordinates, OpenGL axes)."""
    kps = kps.long()
    dirs = torch.stack([(kps[:, 0] - W * .5) / focal, -(kps[:, 1] - H * .5) / focal,
                        -torch.ones_like(kps[:, 0])], -1)
    rays_d = (dirs[:, None, :] * c2w[:3, :3]).sum(-1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def rays_full_image_pinhole(H, W, focal, c2w):
    """NeRF/get_rays.py:5-23; returns [H,W,3] pairs."""
    o, d = rays_pixels_pinhole(H, W, focal, c2w, full_image_pixels(H, W))
    return o.reshape(H, W, 3), d.reshape(H, W, 3)


# --------------------------------------------------------------------------------------
# render driver pieces (NeRF/render.py)
# --------------------------------------------------------------------------------------


def ndc_project(H, W, fx, fy, near, rays_o, rays_d):
    """NeRF/render.py:357-374 (fx == fy == focal) and :376-396 (learnable fx, fy)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    ox_oz = rays_o[..., 0] / rays_o[..., 2]
    oy_oz = rays_o[..., 1] / rays_o[..., 2]
    sx = -1. / (W / (2. * fx))
    sy = -1. / (H / (2. * fy))
    o = torch.stack([sx * ox_oz, sy * oy_oz, 1. + 2. * near / rays_o[..., 2]], -1)
    d = torch.stack([sx * (rays_d[..., 0] / rays_d[..., 2] - ox_oz),
                     sy * (rays_d[..., 1] / rays_d[..., 2] - oy_oz),
                     -2. * near / rays_o[..., 2]], -1)
    return o, d


def pack_rays(H, W, rays_o, rays_d, near, far, use_viewdirs, ndc, fx=None, fy=None):
    """The [N, 8|11] ray batch of NeRF/render.py:105-130: [o, d, near, far, viewdirs]."""
    viewdirs = None
    if use_viewdirs:
        viewdirs = (rays_d / torch.norm(rays_d, dim=-1, keepdim=True)).reshape(-1, 3)
    if ndc:
        rays_o, rays_d = ndc_project(H, W, fx, fy, 1., rays_o, rays_d)
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    cols = [rays_o, rays_d, near * torch.ones_like(rays_d[:, :1]), far * torch.ones_like(rays_d[:, :1])]
    if use_viewdirs:
        cols.append(viewdirs)
    return torch.cat(cols, -1)


def stratified_depths(near, far, n, lindisp=False, t_rand=None):
    """NeRF/render.py:235-257.  near/far are [N,1]; t_rand [N,n] in [0,1) or None (perturb=0)."""
    t = torch.linspace(0., 1., steps=n, dtype=near.dtype)
    if lindisp:
        z = 1. / (1. / near * (1. - t) + 1. / far * t)
    else:
        z = near * (1. - t) + far * t
    z = z.expand(near.shape[0], n)
    if t_rand is not None:
        mids = .5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mids, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


def posenc(x, n_freqs):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)].
    NeRF/run_nerf_helpers.py:24-72 (log-sampled bands = exact powers of two)."""
    out = [x]
    for k in range(n_freqs):
        f = float(2 ** k)
        out += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(out, -1)


def mlp_forward(P, x_pts, x_views=None, skips=(4,)):
    """NeRF.forward, NeRF/run_nerf_helpers.py:105-128.  ``P`` maps reference state_dict names
    to tensors.  x_views None <=> use_viewdirs=False (output_linear head)."""
    D = sum(1 for k in P if k.startswith("pts_linears.") and k.endswith(".weight"))
    h = x_pts
    for i in range(D):
        h = F.relu(F.linear(h, P[f"pts_linears.{i}.weight"], P[f"pts_linears.{i}.bias"]))
        if i in skips:
            h = torch.cat([x_pts, h], -1)
    if x_views is None:
        return F.linear(h, P["output_linear.weight"], P["output_linear.bias"])
    alpha = F.linear(h, P["alpha_linear.weight"], P["alpha_linear.bias"])
    feat = F.linear(h, P["feature_linear.weight"], P["feature_linear.bias"])
    hv = F.relu(F.linear(torch.cat([feat, x_views], -1),
                         P["views_linears.0.weight"], P["views_linears.0.bias"]))
    rgb = F.linear(hv, P["rgb_linear.weight"], P["rgb_linear.bias"])
    return torch.cat([rgb, alpha], -1)


def query_field(P, pts, viewdirs, L_pos=10, L_dir=4):
    """run_network, NeRF/create_nerf.py:18-32: PE of points, PE of per-ray dirs broadcast
    over samples, MLP.  pts [N,S,3] -> raw [N,S,4|5]."""
    N, S, _ = pts.shape
    e = posenc(pts.reshape(-1, 3), L_pos)
    ev = None
    if viewdirs is not None:
        ev = posenc(viewdirs[:, None, :].expand(N, S, 3).reshape(-1, 3), L_dir)
    return mlp_forward(P, e, ev).reshape(N, S, -1)


def composite(raw, z, rays_d, noise=None, white_bkgd=False):
    """raw2outputs, NeRF/render.py:302-355.  ``noise`` is the already-scaled additive
    sigma noise [N,S] (randn*raw_noise_std, or rand*raw_noise_std when pytest=True)."""
    dists = z[:, 1:] - z[:, :-1]
    dists = torch.cat([dists, torch.full_like(dists[:, :1], 1e10)], -1)
    dists = dists * torch.norm(rays_d[:, None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1. - torch.exp(-F.relu(sigma) * dists)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb_map = (weights[..., None] * rgb).sum(-2)
    depth_map = (weights * z).sum(-1)
    acc_map = weights.sum(-1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / (acc_map + 1e-10))
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[:, None])
    return rgb_map, disp_map, acc_map, weights, depth_map


def pdf_to_cdf(weights):
    """NeRF/render.py:419-422: +1e-5, normalise, cumulative sum, leading zero."""
    w = weights + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)


def pdf_to_cdf_sequential(weights):
    """Same as pdf_to_cdf but with a strictly sequential fp32 sum and running sum in index order
    (numpy, one add at a time).  torch.sum's vectorised order is build/ISA dependent (its CDF
    differs from this one by up to ~2e-6 on 62 bins, and the reference's CUDA run differs again),
    so this is the order the CUDA kernel fixes; tests use it to bit-match the kernel's logic."""
    import numpy as np
    w = (np.asarray(weights, dtype=np.float32) + np.float32(1e-5)).astype(np.float32)
    tot = np.zeros(w.shape[0], np.float32)
    for i in range(w.shape[1]):
        tot = (tot + w[:, i]).astype(np.float32)
    run = np.zeros(w.shape[0], np.float32)
    cols = [run.copy()]
    for i in range(w.shape[1]):
        run = (run + (w[:, i] / tot).astype(np.float32)).astype(np.float32)
        cols.append(run.copy())
    return torch.from_numpy(np.stack(cols, 1))


def inverse_cdf_sample(bins, weights, u, return_inds=False, cdf=None):
    """sample_pdf, NeRF/render.py:417-460, with the uniforms ``u`` [N,Nf] supplied
    (linspace(0,1,Nf) when det, rand otherwise).  ``cdf`` overrides the CDF (summation-order
    studies)."""
    cdf = pdf_to_cdf(weights) if cdf is None else cdf
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    samples = b0 + (u - c0) / denom * (b1 - b0)
    return (samples, inds) if return_inds else samples


def render_rays(rays, P_coarse, P_fine, N_samples, N_importance=0, lindisp=False,
                white_bkgd=False, t_rand=None, u=None, noise0=None, noise1=None,
                retraw=False, L_pos=10, L_dir=4, sequential_cdf=False):
    """NeRF/render.py:186-300 with all randomness injected:
    t_rand [N,Nc] (None = perturb 0), u [N,Nf] (None = deterministic linspace),
    noise0 [N,Nc] / noise1 [N,Nc+Nf] already multiplied by raw_noise_std."""
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    viewdirs = rays[:, -3:] if rays.shape[-1] > 8 else None
    near, far = rays[:, 6:7], rays[:, 7:8]
    z = stratified_depths(near, far, N_samples, lindisp, t_rand)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]
    raw = query_field(P_coarse, pts, viewdirs, L_pos, L_dir)
    rgb, disp, acc, weights, depth = composite(raw, z, rays_d, noise0, white_bkgd)
    out = {}
    if N_importance > 0:
        out.update(rgb0=rgb, disp0=disp, acc0=acc)
        z_mid = .5 * (z[:, 1:] + z[:, :-1])
        if u is None:
            u = torch.linspace(0., 1., steps=N_importance, dtype=z.dtype).expand(z.shape[0], N_importance)
        cdf = pdf_to_cdf_sequential(weights[:, 1:-1].detach().float().numpy()).to(z.dtype) if sequential_cdf else None
        z_samples = inverse_cdf_sample(z_mid, weights[:, 1:-1], u, cdf=cdf).detach()
        z, _ = torch.sort(torch.cat([z, z_samples], -1), -1)
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]
        raw = query_field(P_fine if P_fine is not None else P_coarse, pts, viewdirs, L_pos, L_dir)
        rgb, disp, acc, weights, depth = composite(raw, z, rays_d, noise1, white_bkgd)
        out["z_std"] = torch.std(z_samples, dim=-1, unbiased=False)
    out.update(rgb_map=rgb, disp_map=disp, acc_map=acc)
    if retraw:
        out["raw"] = raw
    out["_z_vals"], out["_weights"], out["_depth"] = z, weights, depth
    return out


def clamp_rgb_(ret):
    """batchify_rays' in-place saturation, NeRF/render.py:404-406 (kills the gradient of
    saturated channels: autograd sees the in-place masked write)."""
    for k in ("rgb0", "rgb1", "rgb_map"):
        if k in ret:
            ret[k] = torch.where(ret[k] >= 1.0, torch.ones_like(ret[k]), ret[k])
    return ret


def img2mse(x, y):
    """NeRF/run_nerf_helpers.py:10."""
    return torch.mean((x - y) ** 2)


def train_step(cam: Camera, P_coarse, P_fine, kps, idx, target, H, W, N_samples, N_importance,
               near=0., far=1., t_rand=None, u=None, noise0=None, noise1=None, white_bkgd=False,
               sequential_cdf=False):
    """One optimisation step's forward, NeRF/run_nerf.py:385-506 (camera branch, NDC,
    use_viewdirs): pixels -> rays -> render -> loss = mse(rgb) + mse(rgb0).  Caller runs
    ``loss.backward()``; gradients land on every tensor that requires grad."""
    rays_o, rays_d = rays_pixels_camera(H, W, cam, kps, idx=idx)
    K = cam.intrinsic()
    rays = pack_rays(H, W, rays_o, rays_d, near, far, True, True, K[0, 0], K[1, 1])
    ret = clamp_rgb_(render_rays(rays, P_coarse, P_fine, N_samples, N_importance,
                                 white_bkgd=white_bkgd, t_rand=t_rand, u=u,
                                 noise0=noise0, noise1=noise1, sequential_cdf=sequential_cdf))
    loss = img2mse(ret["rgb_map"], target)
    if "rgb0" in ret:
        loss = loss + img2mse(ret["rgb0"], target)
    return loss, ret, rays


def c3_train_step(cam: Camera, P_coarse, P_fine, kps, idx, target, H, W, N_samples, N_importance, matches, pair,
                  prd_weight, threshold, **rand):
    """BASELINE configs[2] forward, NeRF/run_nerf.py:482-598: the render loss of ``train_step`` plus
    ``ray_dist_loss_weight`` x the projected-ray-distance loss on the sub-pixel matches (kps0, kps1) of the image
    pair ``pair`` = (i, j), whose rays come from the same learnable camera (:535-548, :571-586).
    -> (total, dict(loss_render, prd, n_match, rgb))."""
    loss, ret, _ = train_step(cam, P_coarse, P_fine, kps, idx, target, H, W, N_samples, N_importance, **rand)
    kps0, kps1 = matches
    i, j = pair
    rays_i = rays_pixels_camera(H, W, cam, kps0, idx=i)
    rays_j = rays_pixels_camera(H, W, cam, kps1, idx=j)
    E2 = cam.extrinsic()[[i, j]]
    prd, n_match = proj_ray_dist_loss(kps0, kps1, rays_i, rays_j, cam.intrinsic(), E2, threshold, train=True,
                                      method="NeRF")
    return loss + prd_weight * prd, dict(loss_render=loss, prd=prd, n_match=n_match, rgb=ret["rgb_map"])


def state_to_tensors(state, dtype=torch.float32, requires_grad=False):
    return {k: torch.as_tensor(v, dtype=dtype).clone().requires_grad_(requires_grad)
            for k, v in state.items()}


# --------------------------------------------------------------------------------------
# searchsorted primitive (NeRF/torchsearchsorted/src/cpu/searchsorted_cpu_wrapper.cpp:82-126;
# semantics pinned by test/test_searchsorted.py against numpy)
# --------------------------------------------------------------------------------------


def searchsorted_rows(a, v, right):
    """Per-row numpy searchsorted with the extension's broadcasting (one row broadcasts)."""
    import numpy as np
    a, v = np.asarray(a), np.asarray(v)
    nrow = max(a.shape[0], v.shape[0])
    out = np.empty((nrow, v.shape[1]), dtype=np.int64)
    for r in range(nrow):
        out[r] = np.searchsorted(a[r if a.shape[0] > 1 else 0], v[r if v.shape[0] > 1 else 0],
                                 side="right" if right else "left")
    return out


# --------------------------------------------------------------------------------------
# optimiser step: NeRF/create_nerf.py:199-258, nerfplusplus/custom_optim.py:11-70
# --------------------------------------------------------------------------------------


def custom_adam_step(params, grads, exp_avgs, exp_avg_sqs, max_exp_avg_sqs, steps, camera_model_name, *,
                     amsgrad, beta1, beta2, lr, weight_decay, eps):
    """`f_custom_adam` restated: Adam where only the LAST parameters (ray_o / ray_d / distortion, chosen by
    substring of the camera-model name, :222-230) receive weight decay.  Updates the lists in place."""
    decay_from = len(params)
    if camera_model_name != "none":
        decay_from -= "rayo" in camera_model_name
        decay_from -= "rayd" in camera_model_name
        decay_from -= "dist" in camera_model_name
    for i, p in enumerate(params):
        g = grads[i]
        bc1, bc2 = 1 - beta1 ** steps[i], 1 - beta2 ** steps[i]
        if weight_decay != 0 and i >= decay_from:
            g = g + weight_decay * p
        exp_avgs[i] = exp_avgs[i] * beta1 + (1 - beta1) * g
        exp_avg_sqs[i] = exp_avg_sqs[i] * beta2 + (1 - beta2) * g * g
        if amsgrad:
            max_exp_avg_sqs[i] = torch.maximum(max_exp_avg_sqs[i], exp_avg_sqs[i])
            denom = max_exp_avg_sqs[i].sqrt() / math.sqrt(bc2) + eps
        else:
            denom = exp_avg_sqs[i].sqrt() / math.sqrt(bc2) + eps
        params[i] = p - (lr / bc1) * exp_avgs[i] / denom
    return params


# --------------------------------------------------------------------------------------
# projected ray distance loss: model/ray_dist_loss.py:96-246
# --------------------------------------------------------------------------------------


def proj_ray_dist_loss(kps0, kps1, rays0, rays1, K4x4, E2, threshold, train=True, method="NeRF", eps=1e-10):
    """Restatement of `proj_ray_dist_loss_single` after its mode/camera dispatch: K4x4 is the 4x4 intrinsic,
    E2 = [2,4,4] camera-to-world of image 0 and image 1.  -> (loss, n_matches or None)."""
    (o0, d0), (o1, d1) = rays0, rays1
    K = K4x4.clone()
    if method == "NeRF":
        K = torch.cat([torch.cat([-K[0:1, 0:1], K[0:1, 1:]], 1), K[1:]], 0)     # :116-118
    Rt = E2[:, :3, :3].transpose(1, 2)
    tinv = -(Rt @ E2[:, :3, 3, None]).squeeze(-1)                                # :120-126
    d0 = d0 / (d0.norm(p=2, dim=-1, keepdim=True) + eps)                        # :128-129
    d1 = d1 / (d1.norm(p=2, dim=-1, keepdim=True) + eps)
    c = (d0 * d1).sum(-1)
    den = c ** 2 - 1 + eps
    t0 = ((d0 * (o0 - o1)).sum(-1) - c * (d1 * (o0 - o1)).sum(-1)) / den        # :145-157
    t1 = ((d1 * (o1 - o0)).sum(-1) - c * (d0 * (o1 - o0)).sum(-1)) / den        # :159-171
    p0 = t0[:, None] * d0 + o0
    p1 = t1[:, None] * d1 + o1

    def project(p, cam):                                                         # :176-186
        q = p @ Rt[cam].T + tinv[cam]
        x = q @ K[:3, :3].T
        return x[:, :2] / (x[:, 2:3] + eps)

    uv0, uv1 = project(p0, 1), project(p1, 0)
    valid = (t0 > 0) & (t1 > 0)                                                  # :188-190
    loss0 = ((uv1[valid] - kps0[valid]) ** 2).sum(-1)                            # :207-212
    loss1 = ((uv0[valid] - kps1[valid]) ** 2).sum(-1)
    if train:
        k0 = (loss0 < threshold) & torch.isfinite(loss0)
        k1 = (loss1 < threshold) & torch.isfinite(loss1)
        return 0.5 * (loss0[k0].mean() + loss1[k1].mean()), float((k0 & k1).sum())
    bad0 = (loss0 > threshold) | ~torch.isfinite(loss0)
    bad1 = (loss1 > threshold) | ~torch.isfinite(loss1)
    loss0 = torch.where(bad0, torch.full_like(loss0, threshold), loss0)
    loss1 = torch.where(bad1, torch.full_like(loss1, threshold), loss1)
    return 0.5 * (loss0.mean() + loss1.mean()), None
