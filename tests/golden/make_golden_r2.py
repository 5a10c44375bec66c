"""Round-2 golden vectors, by RUNNING THE REFERENCE (imported from /root/reference, CPU fp32).  Build container only:

    python tests/golden/make_golden_r2.py

  raygen_subpixel.npz  get_rays_kps_use_camera with FRACTIONAL keypoints (SIFT / SuperGlue matches, as the PRD
                       loss feeds them: NeRF/run_nerf.py:535-548) — forward and every camera gradient.
  c3_step.npz          BASELINE.json configs[2] composed: one whole optimisation step as NeRF/run_nerf.py:482-621
                       runs it with the full camera — render (64c+128f) -> img2mse(rgb)+img2mse(rgb0), + weight x
                       PRD loss on sub-pixel matches of an image pair, backward, CustomAdamOptimizer.step() (positional
                       weight decay on ray_o / ray_d), learning-rate decay.  Pins losses, every camera gradient, the
                       updated camera parameters and probes of every MLP gradient / updated MLP parameter.
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
torch.autograd.set_detect_anomaly(False)
import render as ref_render  # noqa: E402
import create_nerf as ref_create  # noqa: E402
from camera_dict import camera_dict  # noqa: E402
from model.ray_dist_loss import proj_ray_dist_loss_single  # noqa: E402

H, W, NCAM = synth.FERN_H, synth.FERN_W, synth.FERN_NCAM
T = torch.from_numpy
CAM_KEYS = ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise")


def save(name, **arrays):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")


def make_camera(seed, args=None):
    args = args or synth.camera_args()
    cam = camera_dict[args.camera_model](intrinsics=T(synth.intrinsic_init()), extrinsics=list(synth.camera_poses(seed)),
                                         args=args, H=H, W=W)
    with torch.no_grad():
        for k, v in synth.camera_noise_state(seed).items():
            getattr(cam, k).copy_(T(v))
    for k in CAM_KEYS:
        getattr(cam, k).requires_grad_(True)
    return cam


def make_nerf(seed):
    net = ref_helpers.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
    net.load_state_dict({k: T(v) for k, v in synth.mlp_state(seed).items()})
    return net


def query_fn():
    embed_fn, _ = ref_helpers.get_embedder(10, 0)
    embeddirs_fn, _ = ref_helpers.get_embedder(4, 0)
    return lambda inputs, viewdirs, fn: ref_create.run_network(
        inputs, viewdirs, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)


def golden_raygen_subpixel():
    cam = make_camera(12)
    kps, idx = synth.subpixel_kps(12, 256)
    o, d = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), idx_in_camera_param=T(idx))
    rng = np.random.default_rng(120)
    wo, wd = (rng.standard_normal((256, 3)).astype(np.float32) for _ in range(2))
    ((o * T(wo)).sum() + (d * T(wd)).sum()).backward()
    out = dict(o=o, d=d, wo=wo, wd=wd)
    for k in CAM_KEYS:
        out["g_" + k] = getattr(cam, k).grad
    # scalar camera index and a fixed extrinsic, forward only
    with torch.no_grad():
        o1, d1 = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), idx_in_camera_param=4)
        o2, d2 = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), extrinsic=T(synth.camera_poses(13)[2]))
    out.update(int_o=o1, int_d=d1, ext_o=o2, ext_d=d2)
    save("raygen_subpixel", **out)


def _pin(g, rng):
    g = g.detach().double().reshape(-1)
    probe = torch.from_numpy(rng.standard_normal(g.numel()))
    return np.array([g.norm().item(), (g * probe).sum().item(), g.abs().max().item()])


def golden_c3_step():
    C = synth.c3_case()
    cam = make_camera(C["seed"])
    net, fine = make_nerf(C["seed"]), make_nerf(C["seed"] + 1)
    kps, idx, target = synth.pixel_batch(C["seed"], C["N_rays"])
    grad_vars = list(net.parameters()) + list(fine.parameters()) + list(cam.parameters())   # create_nerf.py:57,65,123
    args = types.SimpleNamespace(camera_model=synth.camera_args().camera_model, proj_ray_dist_threshold=C["threshold"])
    opt = ref_create.CustomAdamOptimizer(params=grad_vars, lr=C["lrate"], betas=(0.9, 0.999),
                                         weight_decay=C["weight_decay"], H=H, W=W, args=args)
    out = {}
    for step in range(C["n_steps"]):
        global_step = C["global_step0"] + step
        o, d = ref_get_rays.get_rays_kps_use_camera(H, W, cam, T(kps), idx_in_camera_param=T(idx))
        rgb, disp, acc, ex = ref_render.render(
            H, W, 1024 * 32, rays=torch.stack([o, d]), camera_model=cam, ndc=True, near=0., far=1., use_viewdirs=True,
            mode="train", network_query_fn=query_fn(), perturb=1., N_importance=C["Nf"], network_fine=fine,
            N_samples=C["Nc"], network_fn=net, white_bkgd=False, raw_noise_std=1., retraw=True, pytest=True)
        opt.zero_grad()
        loss1 = ref_helpers.img2mse(rgb, T(target))
        loss0 = ref_helpers.img2mse(ex["rgb0"], T(target))
        train_loss = loss1 + loss0
        # PRD on sub-pixel matches of the pair (i, j): run_nerf.py:535-598
        i, j = C["pair"]
        kps0, kps1 = synth.c3_matches(C["seed"] + step)
        rays_i = ref_get_rays.get_rays_kps_use_camera(H=H, W=W, camera_model=cam, idx_in_camera_param=i, kps_list=T(kps0))
        rays_j = ref_get_rays.get_rays_kps_use_camera(H=H, W=W, camera_model=cam, idx_in_camera_param=j, kps_list=T(kps1))
        prd, n_match = proj_ray_dist_loss_single(
            kps0_list=T(kps0), kps1_list=T(kps1), img_idx0=i, img_idx1=j, rays0=rays_i, rays1=rays_j, mode="train",
            device="cpu", H=H, W=W, args=args, camera_model=cam, method="NeRF", i_map=np.arange(NCAM))
        train_loss = train_loss + C["prd_weight"] * prd
        train_loss.backward()
        if step == 0:
            rng = np.random.default_rng(77)
            out.update(loss1=loss1.detach(), loss0=loss0.detach(), prd=prd.detach(), n_match=n_match, total=train_loss.detach(),
                       rgb=rgb.detach())
            for k in CAM_KEYS:
                out["g_cam_" + k] = getattr(cam, k).grad.clone()
            for tag, m in (("coarse", net), ("fine", fine)):
                for name, p in m.named_parameters():
                    out[f"gpin_{tag}_{name}"] = _pin(p.grad, rng)
        opt.step()
        new_lrate = C["lrate"] * (0.1 ** (global_step / (C["lrate_decay"] * 1000)))       # run_nerf.py:617-621
        for group in opt.param_groups:
            group["lr"] = new_lrate
        rng = np.random.default_rng(78 + step)
        for k in CAM_KEYS:
            out[f"s{step}_cam_" + k] = getattr(cam, k).detach().clone()
        for tag, m in (("coarse", net), ("fine", fine)):
            for name, p in m.named_parameters():
                out[f"s{step}_ppin_{tag}_{name}"] = _pin(p, rng)
                if p.numel() <= 768:
                    out[f"s{step}_p_{tag}_{name}"] = p.detach().clone()
        out[f"s{step}_total"] = train_loss.detach()
        print(f"c3 step {step}: loss1 {float(loss1):.6f} loss0 {float(loss0):.6f} prd {float(prd):.6f} n_match {n_match}")
    save("c3_step", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    golden_raygen_subpixel()
    golden_c3_step()
