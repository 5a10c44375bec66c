"""nerfplusplus/ddp_model.py — depth2pts_outside (:16-45), NerfNet (:48-143), NerfNetWithAutoExpo (:160-188).

``NerfNet.forward`` is one autograd.Function: foreground field (PE + MLP, tensor-core or fp32 kernels) ->
fg compositing -> background sphere points -> background field -> bg compositing, all in CUDA
(include/scnerf_b200_nerfpp.h).  Gradient flows from ``ret['rgb']`` to both networks' parameters, to
ray_o / ray_d (through the points, view directions, |d| and the sphere re-parametrisation), to
``fg_z_vals`` and ``fg_z_max`` — the reference's graph.  The other entries of the
returned dict (weights and the per-layer maps the trainer only logs or detaches) carry no gradient.
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import _lib
from .nerf_network import Embedder, MLPNet



class _Depth2Pts(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o, d, depth):
        lib = _lib.load()
        N = depth.numel()
        oo, dd, q = _lib.f32(o).reshape(N, 3), _lib.f32(d).reshape(N, 3), _lib.f32(depth).reshape(N, 1)
        pts = torch.empty(N, 1, 4, device=q.device, dtype=torch.float32)
        real = torch.empty(N, 1, device=q.device, dtype=torch.float32)
        _lib.check(lib.scnerf_pp_bg_points_fwd(_lib.ptr(oo), _lib.ptr(dd), _lib.ptr(q), N, 1, _lib.ptr(pts),
                                               _lib.ptr(real), _lib.stream()), "depth2pts_outside")
        ctx.save_for_backward(oo, dd, q)
        ctx.shape = o.shape
        real = real.reshape(depth.shape)
        ctx.mark_non_differentiable(real)
        return pts.reshape(*depth.shape, 4), real

    @staticmethod
    def backward(ctx, g_pts, _g_real):
        lib = _lib.load()
        oo, dd, q = ctx.saved_tensors
        N = q.shape[0]
        g = _lib.f32(g_pts).reshape(N, 1, 4)
        g_o, g_d = torch.zeros_like(oo), torch.zeros_like(dd)
        _lib.check(lib.scnerf_pp_bg_points_bwd(_lib.ptr(oo), _lib.ptr(dd), _lib.ptr(q), _lib.ptr(g), N, 1,
                                               _lib.ptr(g_o), _lib.ptr(g_d), _lib.stream()), "depth2pts_outside_bwd")
        return g_o.reshape(ctx.shape), g_d.reshape(ctx.shape), None


def depth2pts_outside(ray_o, ray_d, depth):
    """ray_o, ray_d: [..., 3]; depth: [...] inverse distance to the sphere origin -> (pts[..., 4], depth_real[...]).
    Differentiable w.r.t. ray_o / ray_d (depth_real is returned for API parity, without a gradient)."""
    return _Depth2Pts.apply(ray_o, ray_d, depth)


class _NerfNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, grad_enabled, ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals, *params):
        lib = _lib.load()
        st = _lib.stream()
        lead = ray_d.shape[:-1]
        o, d = _lib.f32(ray_o).reshape(-1, 3), _lib.f32(ray_d).reshape(-1, 3)
        N = o.shape[0]
        Sf, Sb = fg_z_vals.shape[-1], bg_z_vals.shape[-1]
        fz, bz = _lib.f32(fg_z_vals).reshape(N, Sf), _lib.f32(bg_z_vals).reshape(N, Sb)
        zmax = _lib.f32(fg_z_max).reshape(N)
        dev = o.device
        E = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        prec = _lib.PRECISION[net.precision]
        # (needs_input_grad reflects requires_grad even under torch.no_grad(): the caller's grad mode comes in explicitly)
        training = bool(grad_enabled) and any(ctx.needs_input_grad)
        ws_bytes = lib.scnerf_field_train_workspace_bytes if training else lib.scnerf_field_infer_rays_workspace_bytes
        field_fwd = lib.scnerf_field_train_fwd if training else lib.scnerf_field_infer_fwd   # no-grad: nothing is kept
        m_fg, m_bg = net.fg_net.c_struct(pts_dim=3), net.bg_net.c_struct(pts_dim=4)
        rays = E(N, 11)
        _lib.check(lib.scnerf_pp_pack_rays(_lib.ptr(o), _lib.ptr(d), N, _lib.ptr(rays), st), "pp_pack_rays")
        # ---- foreground
        nb = ws_bytes(m_fg, N, Sf, prec)
        ws_fg = torch.empty(nb, device=dev, dtype=torch.uint8)
        raw_fg = E(N, Sf, 4)
        _lib.check(field_fwd(m_fg, _lib.ptr(rays), 11, _lib.ptr(fz), None, None, N, Sf, _lib.ptr(raw_fg),
                                              prec, _lib.ptr(ws_fg), nb, st), "field_train_fwd(fg)")
        fg_w, fg_rgb, fg_depth, lam = E(N, Sf), E(N, 3), E(N), E(N)
        _lib.check(lib.scnerf_pp_composite_fg_fwd(_lib.ptr(raw_fg), _lib.ptr(fz), _lib.ptr(zmax), _lib.ptr(d), N, Sf,
                                                  _lib.ptr(fg_w), _lib.ptr(fg_rgb), _lib.ptr(fg_depth), _lib.ptr(lam), st),
                   "pp_composite_fg")
        # ---- background (4-D points: its own tensor-core slab plans, or the fp32 CUDA-core kernels)
        pts4 = E(N, Sb, 4)
        _lib.check(lib.scnerf_pp_bg_points_fwd(_lib.ptr(o), _lib.ptr(d), _lib.ptr(bz), N, Sb, _lib.ptr(pts4), None, st),
                   "pp_bg_points")
        vd = rays[:, 8:11].contiguous()
        nbb = ws_bytes(m_bg, N, Sb, prec)
        ws_bg = torch.empty(nbb, device=dev, dtype=torch.uint8)
        raw_bg = E(N, Sb, 4)
        _lib.check(field_fwd(m_bg, None, 0, None, _lib.ptr(pts4), _lib.ptr(vd), N, Sb, _lib.ptr(raw_bg), prec,
                                              _lib.ptr(ws_bg), nbb, st), "field_train_fwd(bg)")
        bg_w, bg_rgb, bg_depth, rgb = E(N, Sb), E(N, 3), E(N), E(N, 3)
        _lib.check(lib.scnerf_pp_composite_bg_fwd(_lib.ptr(raw_bg), _lib.ptr(bz), _lib.ptr(lam), _lib.ptr(fg_rgb), N, Sb,
                                                  _lib.ptr(bg_w), _lib.ptr(bg_rgb), _lib.ptr(bg_depth), _lib.ptr(rgb), st),
                   "pp_composite_bg")
        if training:
            ctx.net, ctx.N, ctx.Sf, ctx.Sb, ctx.prec = net, N, Sf, Sb, prec
            ctx.t = (o, d, fz, bz, zmax, rays, raw_fg, lam, pts4, vd, raw_bg, ws_fg, ws_bg)
            ctx.shapes = (ray_o.shape, ray_d.shape, fg_z_max.shape, fg_z_vals.shape)
        outs = (rgb.reshape(*lead, 3), fg_w.reshape(*lead, Sf), bg_w.reshape(*lead, Sb), fg_rgb.reshape(*lead, 3),
                fg_depth.reshape(lead), bg_rgb.reshape(*lead, 3), bg_depth.reshape(lead), lam.reshape(lead))
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, g_rgb, *_unused):
        lib = _lib.load()
        st = _lib.stream()
        if ctx.t is None:
            raise RuntimeError("NerfNet: backward ran twice through the same forward — the fused node releases its "
                               "workspace after the first pass; sum the losses and call backward once")
        net, N, Sf, Sb, prec = ctx.net, ctx.N, ctx.Sf, ctx.Sb, ctx.prec
        o, d, fz, bz, zmax, rays, raw_fg, lam, pts4, vd, raw_bg, ws_fg, ws_bg = ctx.t
        dev = o.device
        Z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        E = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        g_rgb = _lib.f32(g_rgb).reshape(N, 3)
        fg_params, bg_params = net.fg_net.field_tensors(), net.bg_net.field_tensors()
        g_fg = [Z(*p.shape) for p in fg_params]
        g_bg = [Z(*p.shape) for p in bg_params]
        m_fg, m_bg = net.fg_net.c_struct(pts_dim=3), net.bg_net.c_struct(pts_dim=4)
        gm_fg, gm_bg = net.fg_net.c_struct(g_fg, pts_dim=3), net.bg_net.c_struct(g_bg, pts_dim=4)
        # ---- background
        d_raw_bg, d_lam = E(N, Sb, 4), E(N)
        _lib.check(lib.scnerf_pp_composite_bg_bwd(_lib.ptr(raw_bg), _lib.ptr(bz), _lib.ptr(lam), N, Sb, _lib.ptr(g_rgb),
                                                  _lib.ptr(d_raw_bg), _lib.ptr(d_lam), st), "pp_composite_bg_bwd")
        d_pts4, d_vd = E(N, Sb, 4), Z(N, 3)
        _lib.check(lib.scnerf_field_train_bwd(m_bg, gm_bg, None, 0, None, _lib.ptr(pts4), _lib.ptr(vd), N, Sb,
                                              _lib.ptr(d_raw_bg), None, None, _lib.ptr(d_pts4), _lib.ptr(d_vd), prec,
                                              _lib.ptr(ws_bg), ws_bg.numel(), st), "field_train_bwd(bg)")
        g_o, g_d = Z(N, 3), Z(N, 3)
        _lib.check(lib.scnerf_pp_bg_points_bwd(_lib.ptr(o), _lib.ptr(d), _lib.ptr(bz), _lib.ptr(d_pts4), N, Sb,
                                               _lib.ptr(g_o), _lib.ptr(g_d), st), "pp_bg_points_bwd")
        # ---- foreground
        d_raw_fg, d_fz, d_zmax = E(N, Sf, 4), E(N, Sf), Z(N)
        _lib.check(lib.scnerf_pp_composite_fg_bwd(_lib.ptr(raw_fg), _lib.ptr(fz), _lib.ptr(zmax), _lib.ptr(d), N, Sf,
                                                  _lib.ptr(g_rgb), _lib.ptr(d_lam), _lib.ptr(d_raw_fg), _lib.ptr(d_fz),
                                                  _lib.ptr(d_zmax), _lib.ptr(g_d), st), "pp_composite_fg_bwd")
        d_rays, d_z2 = Z(N, 11), E(N, Sf)
        _lib.check(lib.scnerf_field_train_bwd(m_fg, gm_fg, _lib.ptr(rays), 11, _lib.ptr(fz), None, None, N, Sf,
                                              _lib.ptr(d_raw_fg), _lib.ptr(d_rays), _lib.ptr(d_z2), None, None, prec,
                                              _lib.ptr(ws_fg), ws_fg.numel(), st), "field_train_bwd(fg)")
        d_fz += d_z2
        d_rays[:, 8:11] += d_vd
        _lib.check(lib.scnerf_pp_pack_rays_bwd(_lib.ptr(d), _lib.ptr(d_rays), N, _lib.ptr(g_o), _lib.ptr(g_d), st),
                   "pp_pack_rays_bwd")
        so, sd, sm, sz = ctx.shapes
        ctx.t = None
        return (None, None, g_o.reshape(so), g_d.reshape(sd), d_zmax.reshape(sm), d_fz.reshape(sz), None, *g_fg, *g_bg)


class NerfNet(nn.Module):
    """ddp_model.py:48-143.  ``args``: max_freq_log2, max_freq_log2_viewdirs, netdepth, netwidth, use_viewdirs."""

    def __init__(self, args, precision=None):
        super().__init__()
        self.fg_embedder_position = Embedder(3, args.max_freq_log2 - 1, args.max_freq_log2)
        self.fg_embedder_viewdir = Embedder(3, args.max_freq_log2_viewdirs - 1, args.max_freq_log2_viewdirs)
        self.fg_net = MLPNet(D=args.netdepth, W=args.netwidth, input_ch=self.fg_embedder_position.out_dim,
                             input_ch_viewdirs=self.fg_embedder_viewdir.out_dim, use_viewdirs=args.use_viewdirs)
        self.bg_embedder_position = Embedder(4, args.max_freq_log2 - 1, args.max_freq_log2)
        self.bg_embedder_viewdir = Embedder(3, args.max_freq_log2_viewdirs - 1, args.max_freq_log2_viewdirs)
        self.bg_net = MLPNet(D=args.netdepth, W=args.netwidth, input_ch=self.bg_embedder_position.out_dim,
                             input_ch_viewdirs=self.bg_embedder_viewdir.out_dim, use_viewdirs=args.use_viewdirs)
        # "bf16x3" (default: tcgen05, split-bf16, parity-grade) | "fp32" (CUDA cores) | "bf16" (tcgen05, single pass)
        self.precision = _lib.resolve_precision(precision, self.fg_net, self.bg_net)

    def forward(self, ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals):
        params = self.fg_net.field_tensors() + self.bg_net.field_tensors()
        outs = _NerfNetFn.apply(self, torch.is_grad_enabled(), ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals, *params)
        keys = ("rgb", "fg_weights", "bg_weights", "fg_rgb", "fg_depth", "bg_rgb", "bg_depth", "bg_lambda")
        return OrderedDict(zip(keys, outs))


def remap_name(name):
    """ddp_model.py:146-154."""
    name = name.replace('.', '-')
    if name[-1] == '/':
        name = name[:-1]
    idx = name.rfind('/')
    for _ in range(2):
        if idx >= 0:
            idx = name[:idx].rfind('/')
    return name[idx + 1:]


class NerfNetWithAutoExpo(nn.Module):
    """ddp_model.py:157-188."""

    def __init__(self, args, optim_autoexpo=False, img_names=None, precision=None):
        super().__init__()
        self.nerf_net = NerfNet(args, precision=precision)
        self.optim_autoexpo = optim_autoexpo
        if self.optim_autoexpo:
            assert img_names is not None
            self.img_names = [remap_name(x) for x in img_names]
            self.autoexpo_params = nn.ParameterDict(OrderedDict(
                [(x, nn.Parameter(torch.Tensor([0.5, 0.]))) for x in self.img_names]))

    def forward(self, ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals, img_name=None):
        ret = self.nerf_net(ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals)
        if img_name is not None:
            img_name = remap_name(img_name)
        if self.optim_autoexpo and (img_name in self.autoexpo_params):
            autoexpo = self.autoexpo_params[img_name]
            scale = torch.abs(autoexpo[0]) + 0.5      # two scalars per image: bookkeeping, not the hot path
            shift = autoexpo[1]
            ret['autoexpo'] = (scale, shift)
        return ret
