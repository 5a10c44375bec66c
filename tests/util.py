"""Shared builders for the parity tests, smoke() and bench.py: the same seeded scene as CUDA
modules (scnerf_b200) and as oracle tensors."""
import numpy as np
import torch

from scnerf_b200 import synth

H, W, FOCAL = synth.FERN_H, synth.FERN_W, synth.FERN_FOCAL
T = torch.from_numpy


def build_modules(seed, device, mult=True, n_cams=synth.FERN_NCAM):
    from scnerf_b200.camera_dict import camera_dict
    from scnerf_b200.run_nerf_helpers import NeRF
    args = synth.camera_args(multiplicative_noise=mult)
    cam = camera_dict[args.camera_model](intrinsics=synth.intrinsic_init(),
                                         extrinsics=list(synth.camera_poses(seed, n_cams)),
                                         args=args, H=H, W=W)
    with torch.no_grad():
        for k, v in synth.camera_noise_state(seed, n_cams).items():
            getattr(cam, k).copy_(T(v))
    nets = []
    for s in (seed, seed + 1):
        net = NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
        net.load_state_dict({k: T(v) for k, v in synth.mlp_state(s).items()})
        nets.append(net.to(device))
    return dict(cam=cam.to(device), coarse=nets[0], fine=nets[1], args=args)


def pytest_rand(N, Nc, Nf, perturb, std):
    r = synth.reference_pytest_rand
    return dict(t_rand=T(r((N, Nc))) if perturb else None,
                u=(T(r((N, Nf))) if perturb else T(np.broadcast_to(np.linspace(0., 1., Nf), (N, Nf))
                                                    .astype(np.float32).copy())) if Nf else None,
                noise0=T(r((N, Nc))) * std if std > 0 else None,
                noise1=T(r((N, Nc + Nf))) * std if std > 0 else None)


def oracle_step(seed, kps, idx, target, Nc, Nf, dtype=torch.float32, perturb=1, std=1.0, mult=True,
                white_bkgd=False):
    """loss, rgb, {name: grad} from the CPU oracle with the reference's pytest=True draws."""
    from oracle import scnerf_oracle as O
    N = kps.shape[0]
    cam = O.Camera(synth.intrinsic_init(), synth.camera_poses(seed),
                   synth.camera_args(multiplicative_noise=mult), H, W, dtype=dtype)
    cam.load(synth.camera_noise_state(seed), True)
    Pc = O.state_to_tensors(synth.mlp_state(seed), dtype, True)
    Pf = O.state_to_tensors(synth.mlp_state(seed + 1), dtype, True)
    rnd = {k: (v.to(dtype) if v is not None else None) for k, v in pytest_rand(N, Nc, Nf, perturb, std).items()}
    loss, ret, _ = O.train_step(cam, Pc, Pf if Nf else None, T(kps), T(idx), T(target).to(dtype), H, W, Nc,
                                Nf, white_bkgd=white_bkgd, **rnd)
    loss.backward()
    grads = {"camera." + k: getattr(cam, k).grad.numpy() for k in O.Camera.LEARNABLE}
    grads.update({"coarse." + k: v.grad.numpy() for k, v in Pc.items()})
    if Nf:
        grads.update({"fine." + k: v.grad.numpy() for k, v in Pf.items()})
    return float(loss), ret["rgb_map"].detach().numpy(), grads


def cuda_step(mods, kps, idx, target, Nc, Nf, precision="fp32", perturb=1.0, std=1.0, white_bkgd=False):
    """Same step through the public Python API (get_rays -> render -> loss -> backward)."""
    from scnerf_b200.get_rays import get_rays_kps_use_camera
    from scnerf_b200.render import render
    from scnerf_b200.run_nerf_helpers import img2mse
    cam, net, fine = mods["cam"], mods["coarse"], mods["fine"]
    dev = cam.intrinsics_initial.device
    for m in (cam, net, fine):
        m.zero_grad(set_to_none=True)
    o, d = get_rays_kps_use_camera(H, W, cam, T(kps).to(dev), idx_in_camera_param=T(idx).to(dev))
    rgb, disp, acc, ex = render(H, W, 1024 * 32, rays=torch.stack([o, d]), camera_model=cam, ndc=True,
                                near=0., far=1., use_viewdirs=True, mode="train", network_query_fn=None,
                                perturb=perturb, N_importance=Nf, network_fine=fine if Nf else None,
                                N_samples=Nc, network_fn=net, white_bkgd=white_bkgd, raw_noise_std=std,
                                retraw=True, pytest=True, precision=precision)
    tgt = T(target).to(dev)
    loss = img2mse(rgb, tgt)
    if Nf:
        loss = loss + img2mse(ex["rgb0"], tgt)
    loss.backward()
    grads = {"camera." + k: getattr(cam, k).grad.cpu().numpy() for k in cam.LEARNABLE}
    grads.update({"coarse." + k: p.grad.cpu().numpy() for k, p in net.named_parameters() if p.grad is not None})
    if Nf:
        grads.update({"fine." + k: p.grad.cpu().numpy() for k, p in fine.named_parameters() if p.grad is not None})
    return float(loss), rgb.detach().cpu().numpy(), grads
