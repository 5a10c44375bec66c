"""Shared builders for the parity tests, smoke() and bench.py: the same seeded scene as CUDA
modules (scnerf_b200) and as oracle tensors."""
import numpy as np
import torch

from scnerf_b200 import synth

H, W, FOCAL = synth.FERN_H, synth.FERN_W, synth.FERN_FOCAL
T = torch.from_numpy


build_modules = synth.build_modules      # (lives next to the other synthetic-input builders; bench.py uses it too)


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


# ---- BASELINE configs[2] composed: render step + PRD loss + CustomAdamOptimizer (NeRF/run_nerf.py:482-621) ----------
CAM_KEYS = ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise")


def pin(t, rng):
    """(norm, random projection, max|.|) of a tensor — the golden files hold these for the big MLP tensors."""
    g = np.asarray(t, dtype=np.float64).reshape(-1)
    probe = rng.standard_normal(g.size)
    return np.array([np.linalg.norm(g), float((g * probe).sum()), np.abs(g).max()])


def oracle_c3_steps(dtype=torch.float32):
    """The composed step on the CPU oracle; returns the same keys tests/golden/c3_step.npz holds."""
    from oracle import scnerf_oracle as O
    C = synth.c3_case()
    cam = O.Camera(synth.intrinsic_init(), synth.camera_poses(C["seed"]), synth.camera_args(), H, W, dtype=dtype)
    cam.load(synth.camera_noise_state(C["seed"]), True)
    Pc = O.state_to_tensors(synth.mlp_state(C["seed"]), dtype, True)
    Pf = O.state_to_tensors(synth.mlp_state(C["seed"] + 1), dtype, True)
    kps, idx, target = synth.pixel_batch(C["seed"], C["N_rays"])
    names = [("coarse", k) for k in Pc] + [("fine", k) for k in Pf] + [("camera", k) for k in
             ("intrinsics_initial", "extrinsics_initial") + O.Camera.LEARNABLE]
    # create_nerf.py:57,65,123: grad_vars = coarse, fine, camera.parameters() (frozen *_initial included, no grad)
    state = dict(m={}, v={}, t={})
    out = {}
    lr = C["lrate"]
    for step in range(C["n_steps"]):
        global_step = C["global_step0"] + step
        rnd = {k: (v.to(dtype) if v is not None else None) for k, v in pytest_rand(C["N_rays"], C["Nc"], C["Nf"], 1, 1.0).items()}
        kps0, kps1 = synth.c3_matches(C["seed"] + step)
        total, info = O.c3_train_step(cam, Pc, Pf, T(kps), T(idx), T(target).to(dtype), H, W, C["Nc"], C["Nf"],
                                      (T(kps0).to(dtype), T(kps1).to(dtype)), C["pair"], C["prd_weight"], C["threshold"], **rnd)
        for t in list(Pc.values()) + list(Pf.values()) + cam.learnables():
            t.grad = None
        total.backward()
        tensors = {("coarse", k): v for k, v in Pc.items()}
        tensors.update({("fine", k): v for k, v in Pf.items()})
        tensors.update({("camera", k): getattr(cam, k) for k in O.Camera.LEARNABLE})
        if step == 0:
            rng = np.random.default_rng(77)
            out.update(loss_render=float(info["loss_render"].detach()), prd=float(info["prd"].detach()), n_match=info["n_match"],
                       total=float(total), rgb=info["rgb"].detach().numpy())
            for k in CAM_KEYS:
                out["g_cam_" + k] = getattr(cam, k).grad.numpy().copy()
            for tag, P in (("coarse", Pc), ("fine", Pf)):
                for name, p in P.items():
                    out[f"gpin_{tag}_{name}"] = pin(p.grad.numpy(), rng)
        # CustomAdamOptimizer.step over the tensors that received a gradient, in grad_vars order (:294-336)
        with_grad = [n for n in names if n in tensors and tensors[n].grad is not None]
        params = [tensors[n].detach() for n in with_grad]
        grads = [tensors[n].grad for n in with_grad]
        for n, p in zip(with_grad, params):
            if n not in state["m"]:
                state["m"][n], state["v"][n], state["t"][n] = torch.zeros_like(p), torch.zeros_like(p), 0
            state["t"][n] += 1
        m = [state["m"][n] for n in with_grad]; v = [state["v"][n] for n in with_grad]
        new = O.custom_adam_step(params, grads, m, v, [None] * len(params), [state["t"][n] for n in with_grad],
                                 synth.camera_args().camera_model, amsgrad=False, beta1=0.9, beta2=0.999, lr=lr,
                                 weight_decay=C["weight_decay"], eps=1e-8)
        for n, p_new, mm, vv in zip(with_grad, new, m, v):
            state["m"][n], state["v"][n] = mm, vv
            with torch.no_grad():
                tensors[n].copy_(p_new)
        lr = C["lrate"] * (0.1 ** (global_step / (C["lrate_decay"] * 1000)))
        rng = np.random.default_rng(78 + step)
        for k in CAM_KEYS:
            out[f"s{step}_cam_" + k] = getattr(cam, k).detach().numpy().copy()
        for tag, P in (("coarse", Pc), ("fine", Pf)):
            for name, p in P.items():
                out[f"s{step}_ppin_{tag}_{name}"] = pin(p.detach().numpy(), rng)
        out[f"s{step}_total"] = float(total.detach())
    return out


def cuda_c3_steps(device="cuda:0", precision="bf16x3"):
    """The same composed step through this repo's public Python API: get_rays -> render -> img2mse, PRD loss via
    get_rays (sub-pixel kps) + proj_ray_dist_loss_single, backward, CustomAdamOptimizer.step, lr decay."""
    import types
    from scnerf_b200.custom_optim import CustomAdamOptimizer, update_lrate
    from scnerf_b200.get_rays import get_rays_kps_use_camera
    from scnerf_b200.ray_dist_loss import proj_ray_dist_loss_single
    from scnerf_b200.render import render
    from scnerf_b200.run_nerf_helpers import img2mse
    C = synth.c3_case()
    mods = build_modules(C["seed"], device)
    cam, net, fine = mods["cam"], mods["coarse"], mods["fine"]
    args = types.SimpleNamespace(camera_model=synth.camera_args().camera_model, proj_ray_dist_threshold=C["threshold"])
    grad_vars = list(net.parameters()) + list(fine.parameters()) + list(cam.parameters())
    opt = CustomAdamOptimizer(params=grad_vars, lr=C["lrate"], betas=(0.9, 0.999), weight_decay=C["weight_decay"],
                              H=H, W=W, args=args)
    kps, idx, target = synth.pixel_batch(C["seed"], C["N_rays"])
    kps, idx, tgt = T(kps).to(device), T(idx).to(device), T(target).to(device)
    out = {}
    for step in range(C["n_steps"]):
        global_step = C["global_step0"] + step
        o, d = get_rays_kps_use_camera(H, W, cam, kps, idx_in_camera_param=idx)
        rgb, disp, acc, ex = render(H, W, 1024 * 32, rays=torch.stack([o, d]), camera_model=cam, ndc=True, near=0., far=1.,
                                    use_viewdirs=True, mode="train", network_query_fn=None, perturb=1.,
                                    N_importance=C["Nf"], network_fine=fine, N_samples=C["Nc"], network_fn=net,
                                    white_bkgd=False, raw_noise_std=1., retraw=True, pytest=True, precision=precision)
        opt.zero_grad()
        loss_render = img2mse(rgb, tgt) + img2mse(ex["rgb0"], tgt)
        i, j = C["pair"]
        kps0, kps1 = (T(x).to(device) for x in synth.c3_matches(C["seed"] + step))
        rays_i = get_rays_kps_use_camera(H=H, W=W, camera_model=cam, idx_in_camera_param=i, kps_list=kps0)
        rays_j = get_rays_kps_use_camera(H=H, W=W, camera_model=cam, idx_in_camera_param=j, kps_list=kps1)
        prd, n_match = proj_ray_dist_loss_single(kps0_list=kps0, kps1_list=kps1, img_idx0=i, img_idx1=j, rays0=rays_i,
                                                 rays1=rays_j, mode="train", device=device, H=H, W=W, args=args,
                                                 camera_model=cam, method="NeRF", i_map=np.arange(synth.FERN_NCAM))
        total = loss_render + C["prd_weight"] * prd
        total.backward()
        if step == 0:
            rng = np.random.default_rng(77)
            out.update(loss_render=float(loss_render), prd=float(prd), n_match=float(n_match), total=float(total.detach()),
                       rgb=rgb.detach().cpu().numpy())
            for k in CAM_KEYS:
                out["g_cam_" + k] = getattr(cam, k).grad.cpu().numpy().copy()
            for tag, m in (("coarse", net), ("fine", fine)):
                for name, p in m.named_parameters():
                    out[f"gpin_{tag}_{name}"] = pin(p.grad.cpu().numpy(), rng)
        opt.step()
        update_lrate(opt, C["lrate"], C["lrate_decay"], global_step)
        rng = np.random.default_rng(78 + step)
        for k in CAM_KEYS:
            out[f"s{step}_cam_" + k] = getattr(cam, k).detach().cpu().numpy().copy()
        for tag, m in (("coarse", net), ("fine", fine)):
            for name, p in m.named_parameters():
                out[f"s{step}_ppin_{tag}_{name}"] = pin(p.detach().cpu().numpy(), rng)
                if p.numel() <= 768:
                    out[f"s{step}_p_{tag}_{name}"] = p.detach().cpu().numpy().copy()
        out[f"s{step}_total"] = float(total.detach())
    return out


def oracle_step_chunked(seed, kps, idx, target, Nc, Nf, dtype=torch.float32, chunk=512, perturb=1, std=1.0):
    """``oracle_step`` for batches too large to hold the reference's activation graph at once: the loss is a mean over
    rays, so the rays are processed in chunks whose losses are weighted n_chunk/N and the gradients accumulate —
    the same numbers as one big step up to the summation order.  Randomness: the reference's pytest draws for the
    FULL batch (np.random.seed(0); np.random.rand(N, ...)), sliced per chunk."""
    from oracle import scnerf_oracle as O
    N = kps.shape[0]
    cam = O.Camera(synth.intrinsic_init(), synth.camera_poses(seed), synth.camera_args(), H, W, dtype=dtype)
    cam.load(synth.camera_noise_state(seed), True)
    Pc = O.state_to_tensors(synth.mlp_state(seed), dtype, True)
    Pf = O.state_to_tensors(synth.mlp_state(seed + 1), dtype, True)
    rnd = pytest_rand(N, Nc, Nf, perturb, std)
    total, outs = 0.0, {"rgb_map": [], "rgb0": [], "acc_map": [], "disp_map": []}
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        r = {k: (v[s:e].to(dtype) if v is not None else None) for k, v in rnd.items()}
        loss, ret, _ = O.train_step(cam, Pc, Pf, T(kps[s:e]), T(idx[s:e]), T(target[s:e]).to(dtype), H, W, Nc, Nf, **r)
        w = (e - s) / N
        (loss * w).backward()
        total += float(loss.detach()) * w
        for k in outs:
            outs[k].append(ret[k].detach().numpy())
    grads = {"camera." + k: getattr(cam, k).grad.numpy() for k in O.Camera.LEARNABLE}
    grads.update({"coarse." + k: v.grad.numpy() for k, v in Pc.items()})
    grads.update({"fine." + k: v.grad.numpy() for k, v in Pf.items()})
    return total, {k: np.concatenate(v) for k, v in outs.items()}, grads
