"""Generate golden vectors by RUNNING THE REFERENCE (imported from /root/reference, CPU fp32).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Inputs come from ``scnerf_b200.synth`` (seeded numpy), so the .npz files hold only what
cannot be regenerated: the reference's outputs (plus small inputs for convenience).  The
reference's own ``pytest=True`` hooks (NeRF/render.py:252-255,333-336,432-440) supply the
"random" draws, so oracle / CUDA replay them with ``synth.reference_pytest_rand``.
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
sys.path[:0] = [ROOT, REF + "/NeRF", REF, REF + "/model"]
sys.modules.setdefault("imageio", mock.MagicMock())

from scnerf_b200 import synth  # noqa: E402

import get_rays as ref_get_rays  # noqa: E402
import run_nerf_helpers as ref_helpers  # noqa: E402
torch.autograd.set_detect_anomaly(False)          # undo the import side effect (run_nerf_helpers.py:7)
import render as ref_render  # noqa: E402
import create_nerf as ref_create  # noqa: E402
from camera_dict import camera_dict  # noqa: E402

H, W, NCAM, FOCAL = synth.FERN_H, synth.FERN_W, synth.FERN_NCAM, synth.FERN_FOCAL
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


def make_camera(seed, multiplicative=True, requires_grad=False):
    args = synth.camera_args(multiplicative_noise=multiplicative)
    poses = synth.camera_poses(seed)
    cam = camera_dict[args.camera_model](
        intrinsics=T(synth.intrinsic_init()), extrinsics=list(poses), args=args, H=H, W=W)
    st = synth.camera_noise_state(seed)
    with torch.no_grad():
        for k, v in st.items():
            getattr(cam, k).copy_(T(v))
    for k in st:
        getattr(cam, k).requires_grad_(requires_grad)
    return cam


def make_nerf(seed, use_viewdirs=True):
    kw = dict(D=8, W=256, input_ch=63, output_ch=5, skips=[4],
              input_ch_views=27 if use_viewdirs else 0, use_viewdirs=use_viewdirs)
    net = ref_helpers.NeRF(**kw)
    st = synth.mlp_state(seed, use_viewdirs=use_viewdirs, input_ch_views=kw["input_ch_views"])
    net.load_state_dict({k: T(v) for k, v in st.items()})
    return net


def query_fn():
    embed_fn, _ = ref_helpers.get_embedder(10, 0)
    embeddirs_fn, _ = ref_helpers.get_embedder(4, 0)
    return lambda inputs, viewdirs, fn: ref_create.run_network(
        inputs, viewdirs, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)


def golden_camera():
    out = {}
    for mult in (True, False):
        cam = make_camera(1, mult)
        tag = "mult" if mult else "add"
        out[f"K_{tag}"] = cam.get_intrinsic()
        out[f"E_{tag}"] = cam.get_extrinsic()
    cam = make_camera(1)
    sel = np.random.default_rng(11).integers(0, H * W, 512)
    out["field_sel"] = sel
    out["ray_o_field"] = cam.get_ray_o_noise()[T(sel)]
    out["ray_d_field"] = cam.get_ray_d_noise()[T(sel)]
    K5, E5 = cam(5)
    out["fwd5_K"], out["fwd5_E"] = K5, E5
    save("camera", **out)


def golden_raygen():
    cam = make_camera(2)
    kps, idx, _ = synth.pixel_batch(2, 256)
    out = {}
    o, d = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), idx_in_camera_param=T(idx))
    out["kps_idx_o"], out["kps_idx_d"] = o, d
    o, d = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), idx_in_camera_param=3)
    out["kps_int_o"], out["kps_int_d"] = o, d
    ext = T(synth.camera_poses(7)[4])
    o, d = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), extrinsic=ext)
    out["kps_ext_o"], out["kps_ext_d"] = o, d
    extN = T(synth.camera_poses(8, n_cams=256))
    o, d = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), extrinsic=extN)
    out["kps_extN_o"], out["kps_extN_d"] = o, d
    sel = np.random.default_rng(12).integers(0, H * W, 512)
    out["full_sel"] = sel
    o, d = ref_get_rays.get_rays_full_image_use_camera(H, W, cam, extrinsic=ext)
    out["full_cam_o"], out["full_cam_d"] = o[T(sel)], d[T(sel)]
    o, d = ref_get_rays.get_rays_full_image_no_camera(H, W, FOCAL, ext)
    out["full_pin_o"], out["full_pin_d"] = o.reshape(-1, 3)[T(sel)], d.reshape(-1, 3)[T(sel)]
    o, d = ref_get_rays.get_rays_kps_no_camera(H, W, FOCAL, ext, T(kps))
    out["kps_pin_o"], out["kps_pin_d"] = o, d
    # NDC on the camera rays (render.py:357-396)
    o, d = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), idx_in_camera_param=T(idx))
    no, nd = ref_render.ndc_rays_camera(H, W, cam, 1., o, d)
    out["ndc_cam_o"], out["ndc_cam_d"] = no, nd
    o, d = ref_get_rays.get_rays_kps_no_camera(H, W, FOCAL, ext, T(kps))
    no, nd = ref_render.ndc_rays(H, W, FOCAL, 1., o, d)
    out["ndc_pin_o"], out["ndc_pin_d"] = no, nd
    save("raygen", **out)


def golden_field():
    rng = np.random.default_rng(13)
    pts = rng.uniform(-1.5, 1.5, (64, 3)).astype(np.float32)
    dirs = rng.standard_normal((64, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    e10, n10 = ref_helpers.get_embedder(10, 0)
    e4, n4 = ref_helpers.get_embedder(4, 0)
    assert (n10, n4) == (63, 27)
    x, v = e10(T(pts)), e4(T(dirs))
    net = make_nerf(3)
    raw = net(torch.cat([x, v], -1))
    net_nv = make_nerf(4, use_viewdirs=False)
    raw_nv = net_nv(x)
    save("field", pts=pts, dirs=dirs, pe_pts=x, pe_dirs=v, raw=raw, raw_noview=raw_nv)


def golden_composite():
    rng = np.random.default_rng(14)
    N, S = 32, 64
    raw = (rng.standard_normal((N, S, 4)) * 2).astype(np.float32)
    z = np.sort(rng.uniform(0, 1, (N, S)), -1).astype(np.float32)
    d = rng.standard_normal((N, 3)).astype(np.float32)
    out = dict(raw=raw, z=z, d=d)
    for std, wb, tag in ((0., False, "plain"), (1., False, "noise"), (0., True, "white"), (0.5, True, "noise_white")):
        r = ref_render.raw2outputs(T(raw), T(z), T(d), std, wb, pytest=True)
        for name, val in zip(("rgb", "disp", "acc", "weights", "depth"), r):
            out[f"{tag}_{name}"] = val
    save("composite", **out)


def golden_sample_pdf():
    rng = np.random.default_rng(15)
    N, M, Nf = 64, 63, 128
    bins = np.sort(rng.uniform(0, 1, (N, M)), -1).astype(np.float32)
    w = rng.uniform(0, 1, (N, M - 1)).astype(np.float32) ** 4
    w[:8] = 0.0                      # all-zero weights: uniform pdf after the +1e-5
    w[8:16, 5:] = 0.0                # mass concentrated in few bins -> tiny denominators
    w[16:24] *= 1e-7
    out = dict(bins=bins, weights=w)
    captured = []
    real = torch.searchsorted

    def spy(*a, **k):
        r = real(*a, **k)
        captured.append(r)
        return r

    with mock.patch.object(torch, "searchsorted", spy):
        out["det_samples"] = ref_render.sample_pdf(T(bins), T(w), Nf, det=True, pytest=True)
        out["det_inds"] = captured[-1]
        out["rand_samples"] = ref_render.sample_pdf(T(bins), T(w), Nf, det=False, pytest=True)
        out["rand_inds"] = captured[-1]
    save("sample_pdf", **out)


def golden_render_c1():
    """BASELINE.json configs[0]: 256 rays x 64 coarse, fixed pinhole, CPU reference."""
    kps, _, _ = synth.pixel_batch(5, 256)
    c2w = T(synth.camera_poses(5)[0])
    net = make_nerf(5)
    o, d = ref_get_rays.get_rays_kps_no_camera(H, W, FOCAL, c2w, T(kps))
    out = dict(kps=kps)
    for perturb, std, tag in ((0., 0., "det"), (1., 1., "rand")):
        with torch.no_grad():
            rgb, disp, acc, extras = ref_render.render(
                H, W, 1024 * 32, rays=(o, d), noisy_focal=FOCAL, ndc=True, near=0., far=1.,
                use_viewdirs=True, mode="train", network_query_fn=query_fn(), perturb=perturb,
                N_importance=0, network_fine=None, N_samples=64, network_fn=net,
                white_bkgd=False, raw_noise_std=std, retraw=True, pytest=True)
        out[f"{tag}_rgb"], out[f"{tag}_disp"], out[f"{tag}_acc"] = rgb, disp, acc
        out[f"{tag}_raw"] = extras["raw"][:16]
    save("render_c1", **out)


def golden_render_c2mini():
    """BASELINE.json configs[1] shape at 64 rays: (64c + 128f), learnable camera, NDC."""
    N = 64
    cam = make_camera(6)
    kps, idx, _ = synth.pixel_batch(6, N)
    net, fine = make_nerf(6), make_nerf(7)
    out = {}
    for perturb, std, wb, tag in ((0., 0., False, "det"), (1., 1., False, "rand"), (1., 0., True, "white")):
        with torch.no_grad():
            o, d = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), idx_in_camera_param=T(idx))
            rgb, disp, acc, ex = ref_render.render(
                H, W, 1024 * 32, rays=torch.stack([o, d]), camera_model=cam, ndc=True, near=0.,
                far=1., use_viewdirs=True, mode="train", network_query_fn=query_fn(),
                perturb=perturb, N_importance=128, network_fine=fine, N_samples=64,
                network_fn=net, white_bkgd=wb, raw_noise_std=std, retraw=True, pytest=True)
        out[f"{tag}_rgb"], out[f"{tag}_disp"], out[f"{tag}_acc"] = rgb, disp, acc
        for k in ("rgb0", "disp0", "acc0", "z_std"):
            out[f"{tag}_{k}"] = ex[k]
        out[f"{tag}_raw"] = ex["raw"][:8]
    save("render_c2mini", **out)


def _pin(g, rng):
    g = g.detach().double().reshape(-1)
    probe = torch.from_numpy(rng.standard_normal(g.numel()))
    return np.array([g.norm().item(), (g * probe).sum().item(), g.abs().max().item()])


def golden_train_step():
    """Forward + backward through camera -> rays -> NDC -> coarse+fine -> loss
    (NeRF/run_nerf.py:385-506,600).  Pins loss and every gradient."""
    N = 96
    cam = make_camera(8, requires_grad=True)
    kps, idx, target = synth.pixel_batch(8, N)
    net, fine = make_nerf(8), make_nerf(9)
    o, d = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), idx_in_camera_param=T(idx))
    rgb, disp, acc, ex = ref_render.render(
        H, W, 1024 * 32, rays=torch.stack([o, d]), camera_model=cam, ndc=True, near=0., far=1.,
        use_viewdirs=True, mode="train", network_query_fn=query_fn(), perturb=1.,
        N_importance=128, network_fine=fine, N_samples=64, network_fn=net, white_bkgd=False,
        raw_noise_std=1., retraw=True, pytest=True)
    loss = ref_helpers.img2mse(rgb, T(target)) + ref_helpers.img2mse(ex["rgb0"], T(target))
    loss.backward()
    out = dict(loss=loss, rgb=rgb, rgb0=ex["rgb0"], rays_o=o, rays_d=d)
    for k in ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise"):
        out["g_cam_" + k] = getattr(cam, k).grad
    rng = np.random.default_rng(99)
    for tag, m in (("coarse", net), ("fine", fine)):
        for name, p in m.named_parameters():
            out[f"gpin_{tag}_{name}"] = _pin(p.grad, rng)
            if p.numel() <= 768:
                out[f"g_{tag}_{name}"] = p.grad
    save("train_step", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    golden_camera()
    golden_raygen()
    golden_field()
    golden_composite()
    golden_sample_pdf()
    golden_render_c1()
    golden_render_c2mini()
    golden_train_step()
