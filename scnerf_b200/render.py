"""Render driver, per-ray sampler and compositor with the reference's names, arguments and return
conventions (NeRF/render.py), executing on the CUDA library through the C ABI.

The differentiable path is three autograd nodes, each one C-ABI call forward and one backward:
  get_rays.*            -> scnerf_raygen_fwd/_bwd      (csrc/raygen.cuh)
  render(): ray packing -> scnerf_rayprep_fwd/_bwd     (viewdirs + NDC, csrc/raygen.cuh)
  render_rays()         -> scnerf_render_rays_fwd/_bwd (sampling, field, compositing)
There is no CPU path; tensors must be CUDA tensors.
"""
import ctypes as C
import itertools
import os

import numpy as np
import torch

from . import _lib
from .get_rays import (get_rays_full_image_no_camera, get_rays_full_image_use_camera,
                       get_rays_kps_no_camera, get_rays_kps_use_camera)
from .run_nerf_helpers import NeRF, unwrap

to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)  # noqa: E731

_call_counter = itertools.count()


# --------------------------------------------------------------------------------------------------
# ray packing (viewdirs + NDC)
# --------------------------------------------------------------------------------------------------
class _RayPrep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, camera_model, focal, H, W, ndc, use_viewdirs, near, far,
                intr_noise):
        lib = _lib.load()
        ro, rd = _lib.f32(rays_o).reshape(-1, 3), _lib.f32(rays_d).reshape(-1, 3)
        N = ro.shape[0]
        a = _lib.RayprepArgs()
        keep = []
        if camera_model is not None:
            cam = camera_model.c_struct()
            keep.append(cam)
            a.cam = C.pointer(cam)
        a.focal = float(focal) if focal is not None else 0.0
        a.H, a.W, a.ndc, a.use_viewdirs = int(H), int(W), int(bool(ndc)), int(bool(use_viewdirs))
        a.near_, a.far_, a.N = float(near), float(far), N
        rays = torch.empty(N, 11 if use_viewdirs else 8, device=ro.device, dtype=torch.float32)
        _lib.check(lib.scnerf_rayprep_fwd(C.byref(a), _lib.ptr(ro), _lib.ptr(rd), _lib.ptr(rays),
                                          _lib.stream()), "rayprep_fwd")
        ctx.a, ctx.keep, ctx.ro, ctx.rd, ctx.shape = a, keep, ro, rd, rays_o.shape
        ctx.intr_noise = intr_noise
        return rays

    @staticmethod
    def backward(ctx, g_rays):
        lib = _lib.load()
        g_rays = _lib.f32(g_rays)
        d_ro, d_rd = torch.empty_like(ctx.ro), torch.empty_like(ctx.rd)
        g_intr = None
        if ctx.intr_noise is not None and ctx.needs_input_grad[10]:
            g_intr = torch.zeros_like(ctx.intr_noise, dtype=torch.float32)
        _lib.check(lib.scnerf_rayprep_bwd(C.byref(ctx.a), _lib.ptr(ctx.ro), _lib.ptr(ctx.rd),
                                          _lib.ptr(g_rays), _lib.ptr(d_ro), _lib.ptr(d_rd),
                                          _lib.ptr(g_intr), _lib.stream()), "rayprep_bwd")
        return (d_ro.reshape(ctx.shape), d_rd.reshape(ctx.shape), None, None, None, None, None, None,
                None, None, g_intr)


def _pack_rays(H, W, rays_o, rays_d, camera_model, focal, ndc, use_viewdirs, near, far):
    intr = getattr(camera_model, "intrinsics_noise", None) if camera_model is not None else None
    return _RayPrep.apply(rays_o, rays_d, camera_model, focal, H, W, ndc, use_viewdirs, near, far, intr)


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """NeRF/render.py:357-374 (the kernel implements the near-plane-1 case the callers use)."""
    assert float(near) == 1.0, "only near=1 (the value every caller passes, render.py:113,116)"
    r = _pack_rays(H, W, rays_o, rays_d, None, focal, True, False, 0., 1.)
    return r[:, 0:3].reshape(rays_o.shape), r[:, 3:6].reshape(rays_d.shape)


def ndc_rays_camera(H, W, camera_model, near, rays_o, rays_d):
    """NeRF/render.py:376-396 (learnable fx, fy; differentiable w.r.t. intrinsics_noise)."""
    assert float(near) == 1.0
    r = _pack_rays(H, W, rays_o, rays_d, camera_model, None, True, False, 0., 1.)
    return r[:, 0:3].reshape(rays_o.shape), r[:, 3:6].reshape(rays_d.shape)


# --------------------------------------------------------------------------------------------------
# render_rays
# --------------------------------------------------------------------------------------------------
class _RenderRays(torch.autograd.Function):
    """(rays, coarse params, fine params) -> (rgb, disp, acc, rgb0, disp0, acc0, z_std, raw)."""

    @staticmethod
    def forward(ctx, rays, opt, net_c, net_f, *params):
        lib = _lib.load()
        rays = _lib.f32(rays)
        N, cols = rays.shape
        dev = rays.device
        nc = len(net_c.field_tensors())
        p_c, p_f = params[:nc], params[nc:]
        # needs_input_grad is filled from the inputs' requires_grad even under torch.no_grad(), and grad mode is always
        # off inside forward: the caller's grad mode is captured by render_rays (opt["grad"]).  Without it every no-grad
        # render ran the TRAINING forward (tile images, 4 MB of workspace per ray, no fused composite).
        training = bool(opt.get("grad", True)) and any(ctx.needs_input_grad)
        cfg = _lib.RenderCfg()
        cfg.N_samples, cfg.N_importance, cfg.ray_cols = opt["N_samples"], opt["N_importance"], cols
        cfg.lindisp, cfg.white_bkgd = int(opt["lindisp"]), int(opt["white_bkgd"])
        cfg.perturb, cfg.raw_noise_std = int(opt["perturb"] > 0), float(opt["raw_noise_std"])
        cfg.retraw, cfg.training = int(opt["retraw"]), int(training)
        cfg.precision, cfg.seed = _lib.PRECISION[opt["precision"]], opt["seed"]
        mc = net_c.c_struct(p_c)
        mf = net_f.c_struct(p_f) if net_f is not None else None
        nbytes = lib.scnerf_render_workspace_bytes(C.byref(cfg), C.byref(mc), N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        two = cfg.N_importance > 0
        S_last = cfg.N_samples + cfg.N_importance
        rc = 4 if net_c.use_viewdirs else net_c.output_ch
        new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        rgb, disp, acc = new(N, 3), new(N), new(N)
        rgb0, disp0, acc0, z_std = (new(N, 3), new(N), new(N), new(N)) if two else (None,) * 4
        raw = new(N, S_last, rc) if opt["retraw"] else None
        out = _lib.RenderOut()
        out.rgb_map, out.disp_map, out.acc_map = _lib.ptr(rgb), _lib.ptr(disp), _lib.ptr(acc)
        out.rgb0, out.disp0, out.acc0 = _lib.ptr(rgb0), _lib.ptr(disp0), _lib.ptr(acc0)
        out.z_std, out.raw = _lib.ptr(z_std), _lib.ptr(raw)
        rnd = _lib.RenderRand()
        inj = [opt.get(k) for k in ("t_rand", "u", "noise0", "noise1")]
        inj = [None if t is None else _lib.f32(t).to(dev) for t in inj]
        rnd.t_rand, rnd.u, rnd.noise0, rnd.noise1 = [_lib.ptr(t) for t in inj]
        _lib.check(lib.scnerf_render_rays_fwd(
            C.byref(cfg), _lib.ptr(rays), N, C.byref(mc), C.byref(mf) if mf is not None else None,
            C.byref(rnd), C.byref(out), _lib.ptr(ws), nbytes, _lib.stream()), "render_rays_fwd")
        if training:
            ctx.state = (cfg, rays, mc, mf, rnd, inj, ws, nbytes, net_c, net_f, nc, two)
        nd = [t for t in (z_std, raw) if t is not None]
        if nd:
            ctx.mark_non_differentiable(*nd)
        return rgb, disp, acc, rgb0, disp0, acc0, z_std, raw

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_rgb0, g_disp0, g_acc0, _gz, _graw):
        lib = _lib.load()
        if ctx.state is None:
            raise RuntimeError("render_rays: backward ran twice through the same forward — the fused node releases its "
                               "workspace (tile images of every layer) after the first pass; sum the losses and call "
                               "backward once (retain_graph is not supported here)")
        cfg, rays, mc, mf, rnd, inj, ws, nbytes, net_c, net_f, nc, two = ctx.state
        N, dev = rays.shape[0], rays.device
        gi = _lib.RenderGradsIn()
        held = []
        for name, g in (("rgb_map", g_rgb), ("disp_map", g_disp), ("acc_map", g_acc),
                        ("rgb0", g_rgb0), ("disp0", g_disp0), ("acc0", g_acc0)):
            if g is not None:
                g = _lib.f32(g)
                held.append(g)
                setattr(gi, name, _lib.ptr(g))
        # one flat, zero-initialised gradient buffer for both networks; per-parameter grads are
        # views of it, so a data-parallel all-reduce can move it in one piece (parallel.py)
        tensors = net_c.field_tensors() + (net_f.field_tensors() if net_f is not None else [])
        flat = torch.zeros(sum(t.numel() for t in tensors), device=dev, dtype=torch.float32)
        views, off = [], 0
        for t in tensors:
            views.append(flat[off:off + t.numel()].view(t.shape))
            off += t.numel()
        gc = net_c.c_struct(views[:nc])
        gf = net_f.c_struct(views[nc:]) if net_f is not None else None
        d_rays = torch.empty_like(rays)
        _lib.check(lib.scnerf_render_rays_bwd(
            C.byref(cfg), _lib.ptr(rays), N, C.byref(mc), C.byref(mf) if mf is not None else None,
            C.byref(rnd), C.byref(gi), C.byref(gc), C.byref(gf) if gf is not None else None,
            _lib.ptr(d_rays), _lib.ptr(ws), nbytes, _lib.stream()), "render_rays_bwd")
        ctx.state = None
        grads = [v if ctx.needs_input_grad[4 + i] else None for i, v in enumerate(views)]
        return (d_rays if ctx.needs_input_grad[0] else None, None, None, None, *grads)


def _seed():
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + next(_call_counter)) & 0xFFFFFFFFFFFFFFFF


def _np_rand(shape, device):
    """The reference's ``pytest=True`` draws: np.random.seed(0); np.random.rand(*shape)
    (NeRF/render.py:252-255, 333-336, 432-440)."""
    np.random.seed(0)
    return torch.from_numpy(np.random.rand(*shape).astype(np.float32)).to(device)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False,
                perturb=0., N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.,
                verbose=False, pytest=False, precision=None):
    """NeRF/render.py:186-300.  Same arguments and returned dict keys
    (rgb_map, disp_map, acc_map[, raw][, rgb0, disp0, acc0, z_std]).

    ``network_query_fn`` is accepted for signature parity; when ``network_fn`` is this package's
    ``NeRF`` the field is evaluated by the fused CUDA kernels straight from (rays, z) — positional
    encoding included — so the [N,S,90] embedding and the [N,S,256] activations of the reference
    are never formed.  ``precision`` (extension): "bf16x3" (default, tcgen05 split-bf16) | "fp32" | "bf16"
    (``_lib.default_precision``)."""
    net_c = unwrap(network_fn)
    net_f = unwrap(network_fine) if network_fine is not None else None
    if not isinstance(net_c, NeRF) or (net_f is not None and not isinstance(net_f, NeRF)):
        raise NotImplementedError("render_rays: network_fn must be scnerf_b200's NeRF module "
                                  "(arbitrary Python fields have no CUDA path here)")
    N = ray_batch.shape[0]
    dev = ray_batch.device
    opt = dict(N_samples=int(N_samples), N_importance=int(N_importance), lindisp=bool(lindisp),
               white_bkgd=bool(white_bkgd), perturb=float(perturb), raw_noise_std=float(raw_noise_std),
               retraw=bool(retraw), precision=_lib.resolve_precision(precision, net_c, net_f), seed=_seed(),
               grad=torch.is_grad_enabled())
    if pytest:
        if perturb > 0.:
            opt["t_rand"] = _np_rand((N, N_samples), dev)
        if N_importance > 0:
            if perturb == 0.:   # det: float64 np.linspace cast to f32 (render.py:436-437)
                opt["u"] = torch.from_numpy(np.broadcast_to(
                    np.linspace(0., 1., N_importance), (N, N_importance)).astype(np.float32)).to(dev)
            else:
                opt["u"] = _np_rand((N, N_importance), dev)
        if raw_noise_std > 0.:
            opt["noise0"] = _np_rand((N, N_samples), dev)
            opt["noise1"] = _np_rand((N, N_samples + N_importance), dev)
    params = net_c.field_tensors() + (net_f.field_tensors() if net_f is not None else [])
    rgb, disp, acc, rgb0, disp0, acc0, z_std, raw = _RenderRays.apply(ray_batch, opt, net_c, net_f,
                                                                      *params)
    ret = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc}
    if retraw:
        ret["raw"] = raw
    if N_importance > 0:
        ret.update(rgb0=rgb0, disp0=disp0, acc0=acc0, z_std=z_std)
    # The reference prints on NaN/Inf here (render.py:296-298) at the cost of a device sync per
    # chunk; opt in with SCNERF_CHECK_FINITE=1.
    if os.environ.get("SCNERF_CHECK_FINITE") == "1":
        for k in ret:
            if not torch.isfinite(ret[k]).all():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """NeRF/render.py:398-413: chunk loop + saturation of rgb >= 1 (same values and the same
    zero gradient on saturated channels as the reference's in-place masked write)."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for key in ("rgb0", "rgb1", "rgb_map"):
            if key in ret:
                ret[key] = torch.where(ret[key] >= 1.0, torch.ones_like(ret[key]), ret[key])
        for k, v in ret.items():
            all_ret.setdefault(k, []).append(v)
    return {k: v[0] if len(v) == 1 else torch.cat(v, 0) for k, v in all_ret.items()}


def render(H, W, chunk, rays=None, noisy_focal=None, noisy_extrinsic=None, ndc=True, near=0.,
           far=1., use_viewdirs=False, mode=None, camera_model=None, image_idx=None, i_map=None,
           gt_intrinsic=None, gt_extrinsic=None, transform_align=None, **kwargs):
    """NeRF/render.py:18-141: same five ray-source branches, same return
    ``[rgb_map, disp_map, acc_map, {everything else}]``."""
    assert mode is not None
    focal = None
    if rays is not None:                                   # training: precomputed rays (:27-31)
        if camera_model is None:
            focal = noisy_focal
        rays_o, rays_d = rays
    elif camera_model is not None and mode == "train":     # (:33-50)
        assert i_map is not None and image_idx in i_map
        assert gt_intrinsic is None and gt_extrinsic is None
        idx_in_camera_param = np.where(i_map == image_idx)[0][0]
        rays_o, rays_d = get_rays_full_image_use_camera(
            H=H, W=W, camera_model=camera_model, extrinsic=noisy_extrinsic[idx_in_camera_param])
    elif camera_model is not None and mode in ("val", "test"):   # (:52-67)
        assert noisy_focal is None and noisy_extrinsic is None
        rays_o, rays_d = get_rays_full_image_use_camera(
            H=H, W=W, camera_model=camera_model, extrinsic=transform_align)
    elif camera_model is None and mode == "train":         # (:69-83)
        assert noisy_focal is not None and noisy_extrinsic is not None
        focal = noisy_focal
        rays_o, rays_d = get_rays_full_image_no_camera(H=H, W=W, focal=focal,
                                                       extrinsic=noisy_extrinsic[image_idx])
    elif camera_model is None and mode in ("val", "test"):  # (:85-101)
        assert gt_extrinsic is not None and noisy_focal is None and noisy_extrinsic is None
        focal = gt_intrinsic[0][0].item()
        rays_o, rays_d = get_rays_full_image_no_camera(H=H, W=W, focal=focal,
                                                       extrinsic=gt_extrinsic[image_idx])
    else:
        assert False, "This message should not appear."

    sh = rays_d.shape
    rays_flat = _pack_rays(H, W, rays_o, rays_d, camera_model, focal, ndc, use_viewdirs, near, far)
    all_ret = batchify_rays(rays_flat, chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ["rgb_map", "disp_map", "acc_map"]
    return [all_ret[k] for k in k_extract] + [{k: v for k, v in all_ret.items() if k not in k_extract}]


def render_path(render_poses, hwf, chunk, render_kwargs, mode, gt_imgs=None, args=None, savedir=None,
                camera_model=None, noisy_extrinsic=None, gt_intrinsic=None, gt_extrinsic=None,
                i_map=None, transform_align=None):
    """NeRF/render.py:143-183."""
    H, W, noisy_focal = hwf
    rgbs, disps = [], []
    for i, _pose in enumerate(render_poses):
        image_idx = i_map[i] if i_map is not None else i
        with torch.no_grad():
            rgb, disp, _acc, _ = render(
                H=H, W=W, noisy_focal=noisy_focal, chunk=chunk, noisy_extrinsic=noisy_extrinsic,
                gt_intrinsic=gt_intrinsic, gt_extrinsic=gt_extrinsic, mode=mode,
                camera_model=camera_model, image_idx=image_idx, i_map=i_map,
                transform_align=transform_align[i] if transform_align is not None else None,
                **render_kwargs)
        rgbs.append(rgb.reshape(H, W, 3).cpu().numpy())
        disps.append(disp.reshape(H, W).cpu().numpy())
        if savedir is not None:
            import imageio
            imageio.imwrite(os.path.join(savedir, "{:03d}.png".format(i)), to8b(rgbs[-1]))
    return np.stack(rgbs, 0), np.stack(disps, 0)


# --------------------------------------------------------------------------------------------------
# stage-level functions of the reference (forward only; the differentiable path is render_rays)
# --------------------------------------------------------------------------------------------------
def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """NeRF/render.py:302-355 -> (rgb_map, disp_map, acc_map, weights, depth_map)."""
    lib = _lib.load()
    raw, z, d = _lib.f32(raw), _lib.f32(z_vals), _lib.f32(rays_d)
    N, S = z.shape
    noise = None
    if raw_noise_std > 0.:
        base = _np_rand((N, S), raw.device) if pytest else torch.randn(N, S, device=raw.device)
        noise = (base * raw_noise_std).contiguous()
    new = lambda *s: torch.empty(*s, device=raw.device, dtype=torch.float32)  # noqa: E731
    rgb, disp, acc, w, depth = new(N, 3), new(N), new(N), new(N, S), new(N)
    _lib.check(lib.scnerf_raw2outputs_fwd(
        _lib.ptr(raw), raw.shape[-1], _lib.ptr(z), _lib.ptr(d), 3, _lib.ptr(noise), int(white_bkgd), N,
        S, _lib.ptr(rgb), _lib.ptr(disp), _lib.ptr(acc), _lib.ptr(w), _lib.ptr(depth), _lib.stream()),
        "raw2outputs_fwd")
    return rgb, disp, acc, w, depth


def sample_pdf(bins, weights, N_samples, det=False, pytest=False, return_inds=False):
    """NeRF/render.py:417-460."""
    lib = _lib.load()
    bins, weights = _lib.f32(bins), _lib.f32(weights)
    N, M = bins.shape
    if det:
        u = None
        if pytest:
            u = torch.from_numpy(np.broadcast_to(np.linspace(0., 1., N_samples), (N, N_samples))
                                 .astype(np.float32)).to(bins.device).contiguous()
    else:
        u = _np_rand((N, N_samples), bins.device) if pytest else torch.rand(N, N_samples, device=bins.device)
    out = torch.empty(N, N_samples, device=bins.device, dtype=torch.float32)
    inds = torch.empty(N, N_samples, device=bins.device, dtype=torch.int64) if return_inds else None
    _lib.check(lib.scnerf_sample_pdf_fwd(_lib.ptr(bins), _lib.ptr(weights), _lib.ptr(u), N, M,
                                         N_samples, _lib.ptr(out), _lib.ptr(inds), _lib.stream()),
               "sample_pdf_fwd")
    return (out, inds) if return_inds else out


def searchsorted(a, v, right=True):
    """torch.searchsorted(a, v, right=...) for 2-D float32 rows (NeRF/render.py:444; broadcasting
    as NeRF/torchsearchsorted/src/torchsearchsorted/searchsorted.py:20-53)."""
    lib = _lib.load()
    a, v = _lib.f32(a), _lib.f32(v)
    nrow = max(a.shape[0], v.shape[0])
    out = torch.empty(nrow, v.shape[1], device=a.device, dtype=torch.int64)
    _lib.check(lib.scnerf_searchsorted_f32(_lib.ptr(a), _lib.ptr(v), _lib.ptr(out), a.shape[0],
                                           v.shape[0], a.shape[1], v.shape[1], int(right),
                                           _lib.stream()), "searchsorted")
    return out
