"""nerfplusplus/ddp_train_nerf.py:50-132 — intersect_sphere, perturb_samples, sample_pdf — plus the fused
cascade helpers the train-step mirror uses (``level0_depths`` / ``level1_depths``, :437-467)."""
import torch

from .. import _lib


class _IntersectSphere(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ray_o, ray_d):
        lib = _lib.load()
        o, d = _lib.f32(ray_o).reshape(-1, 3), _lib.f32(ray_d).reshape(-1, 3)
        N = o.shape[0]
        far = torch.empty(N, device=o.device, dtype=torch.float32)
        miss = torch.zeros(1, device=o.device, dtype=torch.int32)
        _lib.check(lib.scnerf_pp_intersect_sphere_fwd(_lib.ptr(o), _lib.ptr(d), N, _lib.ptr(far), _lib.ptr(miss),
                                                      _lib.stream()), "intersect_sphere")
        if int(miss.item()) > 0:        # the reference raises too (:61-65); same device sync as its `.any()`
            raise Exception("Not all your cameras are bounded by the unit sphere; please make sure "
                            "the cameras are normalized properly!")
        ctx.save_for_backward(o, d)
        ctx.shape = ray_o.shape
        return far.reshape(ray_o.shape[:-1])

    @staticmethod
    def backward(ctx, g_far):
        lib = _lib.load()
        o, d = ctx.saved_tensors
        N = o.shape[0]
        g = _lib.f32(g_far).reshape(-1)
        g_o, g_d = torch.zeros_like(o), torch.zeros_like(d)
        _lib.check(lib.scnerf_pp_intersect_sphere_bwd(_lib.ptr(o), _lib.ptr(d), _lib.ptr(g), N, _lib.ptr(g_o),
                                                      _lib.ptr(g_d), _lib.stream()), "intersect_sphere_bwd")
        return g_o.reshape(ctx.shape), g_d.reshape(ctx.shape)


def intersect_sphere(ray_o, ray_d):
    """Depth of the intersection with the unit sphere (:50-68)."""
    return _IntersectSphere.apply(ray_o, ray_d)


class _Perturb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_vals, t_rand):
        lib = _lib.load()
        S = z_vals.shape[-1]
        z, t = _lib.f32(z_vals).reshape(-1, S), _lib.f32(t_rand).reshape(-1, S)
        out = torch.empty_like(z)
        _lib.check(lib.scnerf_pp_perturb_samples_fwd(_lib.ptr(z), _lib.ptr(t), z.shape[0], S, _lib.ptr(out),
                                                     _lib.stream()), "perturb_samples")
        ctx.save_for_backward(t)
        ctx.shape = z_vals.shape
        return out.reshape(z_vals.shape)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (t,) = ctx.saved_tensors
        S = t.shape[-1]
        g = _lib.f32(g).reshape(-1, S)
        gz = torch.empty_like(g)
        _lib.check(lib.scnerf_pp_perturb_samples_bwd(_lib.ptr(g), _lib.ptr(t), g.shape[0], S, _lib.ptr(gz),
                                                     _lib.stream()), "perturb_samples_bwd")
        return gz.reshape(ctx.shape), None


def perturb_samples(z_vals, t_rand=None):
    """:71-80.  ``t_rand`` injects the ``torch.rand_like(z_vals)`` draw (tests); default draws it."""
    if t_rand is None:
        t_rand = torch.rand_like(z_vals)
    return _Perturb.apply(z_vals, t_rand)


class _SamplePdf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bins, weights, u, Nf):
        lib = _lib.load()
        M1 = bins.shape[-1]
        b, w = _lib.f32(bins).reshape(-1, M1), _lib.f32(weights).reshape(-1, M1 - 1)
        N = b.shape[0]
        uu = _lib.f32(u).reshape(N, Nf) if u is not None else None
        s = torch.empty(N, Nf, device=b.device, dtype=torch.float32)
        above = torch.empty(N, Nf, device=b.device, dtype=torch.int64)
        t = torch.empty(N, Nf, device=b.device, dtype=torch.float32)
        _lib.check(lib.scnerf_pp_sample_pdf_bins(_lib.ptr(b), _lib.ptr(w), _lib.ptr(uu), N, M1, Nf, _lib.ptr(s),
                                                 _lib.ptr(above), _lib.ptr(t), _lib.stream()), "sample_pdf")
        ctx.save_for_backward(above, t)
        ctx.shape, ctx.M1 = bins.shape, M1
        return s.reshape(*bins.shape[:-1], Nf)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        above, t = ctx.saved_tensors
        N, Nf = above.shape
        g = _lib.f32(g).reshape(N, Nf)
        gb = torch.zeros(N, ctx.M1, device=g.device, dtype=torch.float32)
        _lib.check(lib.scnerf_pp_sample_pdf_bins_bwd(_lib.ptr(g), _lib.ptr(above), _lib.ptr(t), N, ctx.M1, Nf,
                                                     _lib.ptr(gb), _lib.stream()), "sample_pdf_bwd")
        return gb.reshape(ctx.shape), None, None, None


def sample_pdf(bins, weights, N_samples, det=False, u=None):
    """:83-132.  bins[..., M+1], weights[..., M] -> [..., N_samples].  Differentiable w.r.t. ``bins`` (the
    reference's only caller detaches ``weights``, :452,461; no gradient flows to them here either).
    ``u`` injects the ``torch.rand`` draw (tests)."""
    if not det and u is None:
        u = torch.rand(*weights.shape[:-1], N_samples, device=bins.device)
    return _SamplePdf.apply(bins, weights.detach(), None if det else u, N_samples)


# ---- fused cascade helpers (one kernel each; carry d(depth)/d(far) as `coef`) --------------------------
class _DepthsOfFar(torch.autograd.Function):
    """depth[N,S] as an affine function of far[N]: backward = row-dot with coef."""

    @staticmethod
    def forward(ctx, far, depth, coef):
        ctx.save_for_backward(coef)
        return depth.view_as(depth)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (coef,) = ctx.saved_tensors
        N, S = coef.shape
        g = _lib.f32(g)
        g_far = torch.zeros(N, device=g.device, dtype=torch.float32)
        _lib.check(lib.scnerf_pp_depth_bwd(_lib.ptr(g), _lib.ptr(coef), N, S, _lib.ptr(g_far), _lib.stream()),
                   "pp_depth_bwd")
        return g_far, None, None


def level0_depths(fg_far_depth, N_samples, min_depth=1e-4, t_fg=None, t_bg=None, perturb=True):
    """ddp_train_nerf.py:437-449 in one kernel -> (fg_depth[N,S] differentiable w.r.t. fg_far_depth,
    fg_coef[N,S], bg_depth[N,S]).  ``min_depth``: scalar, or an [N] tensor = ``ray_batch['min_depth']`` (:438; datasets
    with per-pixel min-depth maps, nerf_sample_ray_split.py:166-171)."""
    lib = _lib.load()
    far = _lib.f32(fg_far_depth)
    N, S = far.shape[0], int(N_samples)
    dev = far.device
    if perturb:
        t_fg = torch.rand(N, S, device=dev) if t_fg is None else _lib.f32(t_fg)
        t_bg = torch.rand(N, S, device=dev) if t_bg is None else _lib.f32(t_bg)
    fg, coef, bg = (torch.empty(N, S, device=dev, dtype=torch.float32) for _ in range(3))
    if torch.is_tensor(min_depth) and min_depth.numel() > 1:
        near = _lib.f32(min_depth).reshape(-1).to(dev)
        if near.numel() != N:
            raise ValueError(f"min_depth has {near.numel()} entries for {N} rays")
        _lib.check(lib.scnerf_pp_level0_depths_rays(_lib.ptr(far), _lib.ptr(near), N, S, _lib.ptr(t_fg), _lib.ptr(t_bg),
                                                    _lib.ptr(fg), _lib.ptr(coef), _lib.ptr(bg), _lib.stream()),
                   "level0_depths")
    else:
        _lib.check(lib.scnerf_pp_level0_depths(_lib.ptr(far), float(min_depth), N, S, _lib.ptr(t_fg), _lib.ptr(t_bg),
                                               _lib.ptr(fg), _lib.ptr(coef), _lib.ptr(bg), _lib.stream()), "level0_depths")
    return _DepthsOfFar.apply(fg_far_depth, fg, coef), coef, bg


def level1_depths(depth, weights, N_samples, fg_far_depth=None, coef=None, u=None, det=False):
    """ddp_train_nerf.py:451-467 in one kernel: sample_pdf on the mid-points with weights[..., 1:-1] and
    sort(cat(depth, samples)).  With ``coef`` (fg) the result stays differentiable w.r.t. ``fg_far_depth``."""
    lib = _lib.load()
    dep, w = _lib.f32(depth), _lib.f32(weights)
    N, S = dep.shape
    Nf = int(N_samples)
    dev = dep.device
    if not det and u is None:
        u = torch.rand(N, Nf, device=dev)
    uu = None if det else _lib.f32(u)
    merged = torch.empty(N, S + Nf, device=dev, dtype=torch.float32)
    mcoef = torch.empty(N, S + Nf, device=dev, dtype=torch.float32) if coef is not None else None
    _lib.check(lib.scnerf_pp_sample_pdf(_lib.ptr(dep), _lib.ptr(coef), _lib.ptr(w), _lib.ptr(uu), N, S, Nf, None, None,
                                        _lib.ptr(merged), _lib.ptr(mcoef), _lib.stream()), "pp_sample_pdf")
    if coef is None:
        return merged, None
    return _DepthsOfFar.apply(fg_far_depth, merged, mcoef), mcoef


# ---- full-image inference: ddp_train_nerf.py:135-256 ---------------------------------
def render_single_image(rank, world_size, models, ray_sampler, chunk_size, camera_model, camera_idx=None,
                        min_depth=1e-4):
    """Render every pixel of one image through the cascade (deterministic sampling), each rank taking a
    contiguous slice of the pixels.  ``models`` = {'cascade_level', 'cascade_samples', 'net_0', ...} as in the
    reference; ``ray_sampler`` only needs ``.H``, ``.W`` and, when ``camera_idx`` is None, ``.c2w_mat``.

    ``min_depth``: scalar (1e-4, what the reference uses with a camera model, nerf_sample_ray_split.py:168-169) or a
    [H*W] tensor (the dataset's min-depth map, :166-167).
    Returns (rank 0 only, like the reference) a list over cascade levels of OrderedDicts of [H, W(, 3)] tensors.
    Deviation: results stay on the device and ranks are merged with one NCCL all_gather per
    key instead of per-chunk ``.cpu()`` copies and a Gloo gather."""
    from collections import OrderedDict
    from .nerf_sample_ray_split import render_ray_from_camera
    H, W = ray_sampler.H, ray_sampler.W
    n_pix = H * W
    if (n_pix // world_size) * world_size != n_pix:
        raise Exception('Number of pixels in the image is not divisible by the number of GPUs!\n\t# pixels: {}\n\t# GPUs: {}'
                        .format(n_pix, world_size))
    per = n_pix // world_size
    dev = camera_model.intrinsics_initial.device
    all_idx = torch.arange(rank * per, (rank + 1) * per, device=dev, dtype=torch.int64)
    levels = models['cascade_level']
    merged = [OrderedDict() for _ in range(levels)]
    with torch.no_grad():
        for s0 in range(0, per, chunk_size):
            sel = all_idx[s0:s0 + chunk_size]
            if camera_idx is not None:
                ray_o, ray_d, _ = render_ray_from_camera(camera_model, camera_idx, sel, rank)
            else:
                ray_o, ray_d, _ = render_ray_from_camera(camera_model, None, sel, rank, ray_sampler.c2w_mat)
            for m in range(levels):
                net = models['net_{}'.format(m)]
                Ns = models['cascade_samples'][m]
                if m == 0:
                    fg_far_depth = intersect_sphere(ray_o, ray_d)
                    near = min_depth
                    if torch.is_tensor(min_depth) and min_depth.numel() > 1:
                        near = min_depth.reshape(-1).to(dev)[rank * per + s0: rank * per + s0 + sel.numel()]
                    fg_depth, _, bg_depth = level0_depths(fg_far_depth, Ns, near, perturb=False)
                else:
                    fg_depth, _ = level1_depths(fg_depth, ret['fg_weights'], Ns, det=True)
                    bg_depth, _ = level1_depths(bg_depth, ret['bg_weights'], Ns, det=True)
                ret = net(ray_o, ray_d, fg_far_depth, fg_depth, bg_depth)
                for key in ret:
                    if key not in ('fg_weights', 'bg_weights') and torch.is_tensor(ret[key]):
                        merged[m].setdefault(key, []).append(ret[key])
    for m in range(levels):
        for key in merged[m]:
            merged[m][key] = torch.cat(merged[m][key], dim=0)
    if world_size > 1:
        import torch.distributed as dist
        for m in range(levels):
            for key in merged[m]:
                parts = [torch.empty_like(merged[m][key]) for _ in range(world_size)]
                dist.all_gather(parts, merged[m][key].contiguous())
                merged[m][key] = torch.cat(parts, dim=0)
    if rank != 0:
        return None
    return [OrderedDict((k, v.reshape(H, W, -1).squeeze()) for k, v in merged[m].items()) for m in range(levels)]
