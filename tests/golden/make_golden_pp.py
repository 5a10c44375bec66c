"""Golden vectors for the NeRF++ rows by RUNNING THE REFERENCE
(`/root/reference/nerfplusplus`, CPU fp32).  Build container only:

    python tests/golden/make_golden_pp.py

The reference draws its randomness in place (`torch.rand_like`, `torch.rand`); each call is made under
`torch.manual_seed(s)` and the same draw is regenerated right after with the same seed and shape, and
stored, so the oracle / CUDA path can be fed the identical numbers.
"""
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
sys.path[:0] = [ROOT, REF + "/nerfplusplus", REF, REF + "/model"]
for name in ("imageio", "configargparse", "matplotlib", "matplotlib.backends", "matplotlib.backends.backend_agg",
             "matplotlib.figure", "matplotlib.cm", "matplotlib.pyplot", "piqa", "piqa.ssim", "piqa.lpips",
             "tensorboardX", "lpips", "cv2"):
    sys.modules.setdefault(name, mock.MagicMock())

from scnerf_b200 import synth  # noqa: E402

import nerf_network as ref_net  # noqa: E402
import ddp_model as ref_model  # noqa: E402
import nerf_sample_ray_split as ref_rays  # noqa: E402
import ddp_train_nerf as ref_train  # noqa: E402
from camera_dict import camera_dict  # noqa: E402

torch.autograd.set_detect_anomaly(False)
H, W, NCAM, FOCAL = synth.PP_H, synth.PP_W, synth.PP_NCAM, synth.PP_FOCAL
T = torch.from_numpy


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")


def make_camera(seed, requires_grad=False, k=(-0.05, 0.01)):
    args = synth.pp_camera_args()
    poses = synth.pp_camera_poses(seed)
    cam = camera_dict[args.camera_model](
        intrinsics=T(synth.intrinsic_init(H, W, FOCAL)), extrinsics=list(poses), args=args, H=H, W=W, k=k)
    st = synth.camera_noise_state(seed, n_cams=NCAM, H=H, W=W, with_distortion=True)
    # model/camera_model.py:229,258-262 builds ray_o_noise and ray_d_noise from ONE zeros tensor, so the two
    # Parameters alias the same storage; give them separate storage so each can hold its own values
    cam.ray_o_noise = torch.nn.Parameter(torch.zeros_like(cam.ray_o_noise))
    cam.ray_d_noise = torch.nn.Parameter(torch.zeros_like(cam.ray_d_noise))
    with torch.no_grad():
        for name, v in st.items():
            getattr(cam, name).copy_(T(v))
    for name in st:
        getattr(cam, name).requires_grad_(requires_grad)
    return cam


def net_args():
    return types.SimpleNamespace(max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8, netwidth=256,
                                 use_viewdirs=True)


def make_nerfnet(seed):
    net = ref_model.NerfNet(net_args())
    net.fg_net.load_state_dict({k: T(v) for k, v in synth.pp_mlp_state(seed, 63).items()})
    net.bg_net.load_state_dict({k: T(v) for k, v in synth.pp_mlp_state(seed + 1, 84).items()})
    return net


def golden_raygen():
    out = {}
    for tag, seed in (("a", 30), ("b", 31)):
        cam = make_camera(seed, requires_grad=True)
        sel, ci, _ = synth.pp_pixel_batch(seed, 96)
        o, d, depth = ref_rays.render_ray_from_camera(cam, ci, sel, "cpu")
        rng = np.random.default_rng(seed)
        wo, wd = (T(rng.standard_normal((96, 3)).astype(np.float32)) for _ in range(2))
        ((o * wo).sum() + (d * wd).sum()).backward()
        out.update({f"{tag}_o": o, f"{tag}_d": d, f"{tag}_depth": depth, f"{tag}_wo": wo, f"{tag}_wd": wd,
                    f"{tag}_cam_idx": ci, f"{tag}_sel": sel})
        for name in ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise", "distortion_noise"):
            out[f"{tag}_g_{name}"] = getattr(cam, name).grad
    # extrinsic given as a numpy matrix (test-time path, :210-212)
    cam = make_camera(32)
    sel, _, _ = synth.pp_pixel_batch(32, 64)
    E = synth.pp_camera_poses(99)[3]
    with torch.no_grad():
        o, d, depth = ref_rays.render_ray_from_camera(cam, None, sel, "cpu", extrinsic=E)
    out.update(dict(c_o=o, c_d=d, c_depth=depth, c_sel=sel, c_E=E))
    save("pp_raygen", **out)


def rays_for(seed, N):
    cam = make_camera(seed)
    sel, ci, target = synth.pp_pixel_batch(seed, N)
    with torch.no_grad():
        o, d, _ = ref_rays.render_ray_from_camera(cam, ci, sel, "cpu")
    return o, d, T(target)


def golden_sampling():
    o, d, _ = rays_for(33, 128)
    far = ref_train.intersect_sphere(o, d)
    near = 1e-4 * torch.ones_like(far)
    Ns = 32
    step = (far - near) / (Ns - 1)
    fg = torch.stack([near + i * step for i in range(Ns)], dim=-1)
    torch.manual_seed(5); fg_p = ref_train.perturb_samples(fg)
    torch.manual_seed(5); t_fg = torch.rand_like(fg)
    bg = torch.linspace(0., 1., Ns).view(1, Ns).expand(128, Ns)
    torch.manual_seed(6); bg_p = ref_train.perturb_samples(bg)
    torch.manual_seed(6); t_bg = torch.rand_like(bg)
    rng = np.random.default_rng(7)
    w = T((rng.uniform(0, 1, (128, Ns)) ** 4).astype(np.float32))
    w[:8] = 0.0                                         # empty rays: the TINY_NUMBER paths
    w[8:12, 5:] = 0.0                                   # all mass in a few bins: flat cdf tail
    mid = .5 * (fg_p[..., 1:] + fg_p[..., :-1])
    torch.manual_seed(8); s_rand = ref_train.sample_pdf(mid, w[..., 1:-1], 64, det=False)
    torch.manual_seed(8); u = torch.rand(128, 64)
    s_det = ref_train.sample_pdf(mid, w[..., 1:-1], 64, det=True)
    merged, _ = torch.sort(torch.cat((fg_p, s_rand), dim=-1))
    save("pp_sampling", o=o, d=d, far=far, fg=fg, fg_p=fg_p, t_fg=t_fg, bg_p=bg_p, t_bg=t_bg, w=w,
         s_rand=s_rand, u=u, s_det=s_det, merged=merged)


def golden_field():
    net = make_nerfnet(40)
    o, d, target = rays_for(34, 48)
    o.requires_grad_(True); d.requires_grad_(True)
    far = ref_train.intersect_sphere(o, d)
    near = 1e-4 * torch.ones_like(far)
    Ns = 24
    step = (far - near) / (Ns - 1)
    fg = torch.stack([near + i * step for i in range(Ns)], dim=-1)
    torch.manual_seed(9); fg = ref_train.perturb_samples(fg)
    torch.manual_seed(9); t_fg = torch.rand_like(fg)
    bg = torch.linspace(0., 1., Ns).view(1, Ns).expand(48, Ns)
    torch.manual_seed(10); bg = ref_train.perturb_samples(bg)
    torch.manual_seed(10); t_bg = torch.rand_like(bg)
    pts4, depth_real = ref_model.depth2pts_outside(o.detach().unsqueeze(-2).expand(48, Ns, 3),
                                                   d.detach().unsqueeze(-2).expand(48, Ns, 3), bg)
    ret = net(o, d, far, fg, bg)
    loss = torch.mean((ret["rgb"] - target) ** 2)
    loss.backward()
    out = dict(o=o, d=d, target=target, far=far, fg=fg, bg=bg, t_fg=t_fg, t_bg=t_bg, pts4=pts4,
               depth_real=depth_real, loss=loss, g_o=o.grad, g_d=d.grad)
    out.update({"ret_" + k: v for k, v in ret.items()})
    for name in ("fg_net.base_layers.0.0.weight", "fg_net.base_layers.5.0.weight", "fg_net.sigma_layers.0.weight",
                 "fg_net.rgb_layers.0.weight", "fg_net.rgb_layers.2.bias", "bg_net.base_layers.0.0.weight",
                 "bg_net.base_layers.5.0.weight", "bg_net.base_remap_layers.0.bias", "bg_net.sigma_layers.0.bias",
                 "bg_net.rgb_layers.2.weight"):
        out["g_" + name] = dict(net.named_parameters())[name].grad[:8]      # first rows only: keeps the fixture small
    save("pp_field", **out)


def golden_train_step():
    """Two cascade levels (24, 48 samples), 40 rays, learnable camera: ddp_train_nerf.py:421-488."""
    cam = make_camera(35, requires_grad=True)
    nets = [make_nerfnet(50), make_nerfnet(52)]
    sel, ci, target = synth.pp_pixel_batch(35, 40)
    target = T(target)
    cascade = [24, 48]
    rand = {}
    loss = 0.0
    for m in range(2):
        o, d, _ = ref_rays.render_ray_from_camera(cam, ci, sel, "cpu")
        Ns = cascade[m]
        if m == 0:
            far = ref_train.intersect_sphere(o, d)
            near = 1e-4 * torch.ones_like(d[..., 0])
            step = (far - near) / (Ns - 1)
            fg = torch.stack([near + i * step for i in range(Ns)], dim=-1)
            torch.manual_seed(11); fg = ref_train.perturb_samples(fg)
            torch.manual_seed(11); rand["t_fg"] = torch.rand_like(fg)
            bg = torch.linspace(0., 1., Ns).view(1, Ns).expand(40, Ns)
            torch.manual_seed(12); bg = ref_train.perturb_samples(bg)
            torch.manual_seed(12); rand["t_bg"] = torch.rand_like(bg)
        else:
            fw = ret["fg_weights"].clone().detach()[..., 1:-1]
            mid = .5 * (fg[..., 1:] + fg[..., :-1])
            torch.manual_seed(13); s = ref_train.sample_pdf(mid, fw, Ns, det=False)
            torch.manual_seed(13); rand["u_fg"] = torch.rand(40, Ns)
            fg, _ = torch.sort(torch.cat((fg, s), dim=-1))
            bw = ret["bg_weights"].clone().detach()[..., 1:-1]
            mid = .5 * (bg[..., 1:] + bg[..., :-1])
            torch.manual_seed(14); s = ref_train.sample_pdf(mid, bw, Ns, det=False)
            torch.manual_seed(14); rand["u_bg"] = torch.rand(40, Ns)
            bg, _ = torch.sort(torch.cat((bg, s), dim=-1))
        ret = nets[m](o, d, far, fg, bg)
        loss = loss + torch.mean((ret["rgb"] - target) ** 2)
        if m == 0:
            rgb0 = ret["rgb"]
    loss.backward()
    out = dict(loss=loss, rgb0=rgb0, rgb1=ret["rgb"], fg1=fg, bg1=bg, sel=sel, cam_idx=ci, target=target, **rand)
    for name in ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise", "distortion_noise"):
        out["g_cam_" + name] = getattr(cam, name).grad
    for m in range(2):
        for name in ("fg_net.base_layers.0.0.weight", "fg_net.base_layers.7.0.weight", "fg_net.sigma_layers.0.weight",
                     "bg_net.base_layers.0.0.weight", "bg_net.rgb_layers.0.weight"):
            out[f"g_net{m}_{name}"] = dict(nets[m].named_parameters())[name].grad[:8]
    save("pp_train_step", **out)


if __name__ == "__main__":
    golden_raygen()
    golden_sampling()
    golden_field()
    golden_train_step()
