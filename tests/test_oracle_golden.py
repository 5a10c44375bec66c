"""Pin the CPU oracle against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py importing /root/reference).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import scnerf_oracle as O
from scnerf_b200 import synth

H, W, FOCAL = synth.FERN_H, synth.FERN_W, synth.FERN_FOCAL
T = torch.from_numpy


def close(a, b, rtol=1e-5, atol=1e-6):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def make_cam(seed, mult=True, requires_grad=False):
    cam = O.Camera(synth.intrinsic_init(), synth.camera_poses(seed),
                   synth.camera_args(multiplicative_noise=mult), H, W)
    return cam.load(synth.camera_noise_state(seed), requires_grad)


def test_camera(golden):
    g = golden("camera")
    for mult, tag in ((True, "mult"), (False, "add")):
        cam = make_cam(1, mult)
        close(cam.intrinsic(), g[f"K_{tag}"], 1e-6, 0)
        close(cam.extrinsic(), g[f"E_{tag}"], 1e-6, 1e-7)
    cam = make_cam(1)
    sel = T(g["field_sel"])
    close(cam.ray_o_field()[sel], g["ray_o_field"], 1e-6, 1e-9)
    close(cam.ray_d_field()[sel], g["ray_d_field"], 1e-6, 1e-9)
    # closed-form bilinear lookup (what the CUDA kernel implements) == F.interpolate path
    ys, xs = sel // W, sel % W
    close(O.bilinear_grid_lookup(cam.ray_o_noise, ys, xs, H, W, cam.args.ray_o_noise_scale),
          g["ray_o_field"], 1e-4, 2e-9)
    close(cam.extrinsic()[5], g["fwd5_E"], 1e-6, 1e-7)


def test_raygen(golden):
    g = golden("raygen")
    cam = make_cam(2)
    kps, idx, _ = synth.pixel_batch(2, 256)
    kps, idx = T(kps), T(idx)
    o, d = O.rays_pixels_camera(H, W, cam, kps, idx=idx)
    close(o, g["kps_idx_o"]); close(d, g["kps_idx_d"])
    o, d = O.rays_pixels_camera(H, W, cam, kps, idx=3)
    close(o, g["kps_int_o"]); close(d, g["kps_int_d"])
    ext = T(synth.camera_poses(7)[4])
    o, d = O.rays_pixels_camera(H, W, cam, kps, extrinsic=ext)
    close(o, g["kps_ext_o"]); close(d, g["kps_ext_d"])
    o, d = O.rays_pixels_camera(H, W, cam, kps, extrinsic=T(synth.camera_poses(8, n_cams=256)))
    close(o, g["kps_extN_o"]); close(d, g["kps_extN_d"])
    sel = T(g["full_sel"])
    o, d = O.rays_full_image_camera(H, W, cam, ext)
    close(o[sel], g["full_cam_o"]); close(d[sel], g["full_cam_d"])
    o, d = O.rays_full_image_pinhole(H, W, FOCAL, ext)
    close(o.reshape(-1, 3)[sel], g["full_pin_o"]); close(d.reshape(-1, 3)[sel], g["full_pin_d"])
    o, d = O.rays_pixels_pinhole(H, W, FOCAL, ext, kps)
    close(o, g["kps_pin_o"]); close(d, g["kps_pin_d"])
    no, nd = O.ndc_project(H, W, FOCAL, FOCAL, 1., o, d)
    close(no, g["ndc_pin_o"], 1e-5, 1e-6); close(nd, g["ndc_pin_d"], 1e-5, 1e-6)
    o, d = O.rays_pixels_camera(H, W, cam, kps, idx=idx)
    K = cam.intrinsic()
    no, nd = O.ndc_project(H, W, K[0, 0], K[1, 1], 1., o, d)
    close(no, g["ndc_cam_o"], 1e-5, 1e-6); close(nd, g["ndc_cam_d"], 1e-5, 1e-6)


def test_field(golden):
    g = golden("field")
    x, v = O.posenc(T(g["pts"]), 10), O.posenc(T(g["dirs"]), 4)
    close(x, g["pe_pts"], 0, 0); close(v, g["pe_dirs"], 0, 0)
    P = O.state_to_tensors(synth.mlp_state(3))
    close(O.mlp_forward(P, x, v), g["raw"], 1e-5, 1e-6)
    P = O.state_to_tensors(synth.mlp_state(4, use_viewdirs=False, input_ch_views=0))
    close(O.mlp_forward(P, x, None), g["raw_noview"], 1e-5, 1e-6)


def test_composite(golden):
    g = golden("composite")
    raw, z, d = T(g["raw"]), T(g["z"]), T(g["d"])
    for std, wb, tag in ((0., False, "plain"), (1., False, "noise"), (0., True, "white"), (0.5, True, "noise_white")):
        noise = T(synth.reference_pytest_rand(z.shape)) * std if std > 0 else None
        r = O.composite(raw, z, d, noise, wb)
        for name, val in zip(("rgb", "disp", "acc", "weights", "depth"), r):
            close(val, g[f"{tag}_{name}"], 1e-6, 1e-7)


def test_sample_pdf(golden):
    g = golden("sample_pdf")
    bins, w = T(g["bins"]), T(g["weights"])
    # pytest=True + det draws u from float64 np.linspace cast to f32 (NeRF/render.py:436-437)
    u = T(np.linspace(0., 1., 128).astype(np.float32)).expand(64, 128)
    s, inds = O.inverse_cdf_sample(bins, w, u, return_inds=True)
    assert np.array_equal(inds.numpy(), g["det_inds"])
    close(s, g["det_samples"], 0, 0)
    u = T(synth.reference_pytest_rand((64, 128)))
    s, inds = O.inverse_cdf_sample(bins, w, u, return_inds=True)
    assert np.array_equal(inds.numpy(), g["rand_inds"])
    close(s, g["rand_samples"], 0, 0)


def test_render_c1(golden):
    """BASELINE.json configs[0]."""
    g = golden("render_c1")
    kps = T(g["kps"])
    c2w = T(synth.camera_poses(5)[0])
    P = O.state_to_tensors(synth.mlp_state(5))
    o, d = O.rays_pixels_pinhole(H, W, FOCAL, c2w, kps)
    rays = O.pack_rays(H, W, o, d, 0., 1., True, True, FOCAL, FOCAL)
    for tag in ("det", "rand"):
        rnd = T(synth.reference_pytest_rand((256, 64))) if tag == "rand" else None
        with torch.no_grad():
            r = O.clamp_rgb_(O.render_rays(rays, P, None, 64, 0, t_rand=rnd, noise0=rnd, retraw=True))
        close(r["rgb_map"], g[f"{tag}_rgb"], 1e-5, 1e-6)
        close(r["disp_map"], g[f"{tag}_disp"], 1e-5, 1e-6)
        close(r["acc_map"], g[f"{tag}_acc"], 1e-5, 1e-6)
        close(r["raw"][:16], g[f"{tag}_raw"], 1e-4, 1e-5)


def test_render_c2mini(golden):
    g = golden("render_c2mini")
    N = 64
    cam = make_cam(6)
    kps, idx, _ = synth.pixel_batch(6, N)
    Pc, Pf = O.state_to_tensors(synth.mlp_state(6)), O.state_to_tensors(synth.mlp_state(7))
    for perturb, std, wb, tag in ((0, 0., False, "det"), (1, 1., False, "rand"), (1, 0., True, "white")):
        t_rand = T(synth.reference_pytest_rand((N, 64))) if perturb else None
        u = T(synth.reference_pytest_rand((N, 128))) if perturb else None
        n0 = T(synth.reference_pytest_rand((N, 64))) * std if std > 0 else None
        n1 = T(synth.reference_pytest_rand((N, 192))) * std if std > 0 else None
        with torch.no_grad():
            o, d = O.rays_pixels_camera(H, W, cam, T(kps), idx=T(idx))
            K = cam.intrinsic()
            rays = O.pack_rays(H, W, o, d, 0., 1., True, True, K[0, 0], K[1, 1])
            r = O.clamp_rgb_(O.render_rays(rays, Pc, Pf, 64, 128, white_bkgd=wb, t_rand=t_rand, u=u,
                                           noise0=n0, noise1=n1, retraw=True))
        for k, gk in (("rgb_map", "rgb"), ("disp_map", "disp"), ("acc_map", "acc"), ("rgb0", "rgb0"),
                      ("disp0", "disp0"), ("acc0", "acc0"), ("z_std", "z_std")):
            close(r[k], g[f"{tag}_{gk}"], 2e-5, 2e-6)
        close(r["raw"][:8], g[f"{tag}_raw"], 1e-4, 1e-5)


def test_train_step_gradients(golden):
    g = golden("train_step")
    N = 96
    cam = make_cam(8, requires_grad=True)
    kps, idx, target = synth.pixel_batch(8, N)
    Pc = O.state_to_tensors(synth.mlp_state(8), requires_grad=True)
    Pf = O.state_to_tensors(synth.mlp_state(9), requires_grad=True)
    loss, ret, _ = O.train_step(
        cam, Pc, Pf, T(kps), T(idx), T(target), H, W, 64, 128,
        t_rand=T(synth.reference_pytest_rand((N, 64))), u=T(synth.reference_pytest_rand((N, 128))),
        noise0=T(synth.reference_pytest_rand((N, 64))), noise1=T(synth.reference_pytest_rand((N, 192))))
    loss.backward()
    close(loss, g["loss"], 1e-5, 0)
    close(ret["rgb_map"], g["rgb"], 2e-5, 2e-6)
    for k in O.Camera.LEARNABLE:
        ref = g["g_cam_" + k]
        # fp32 noise floor: the reference's own fp32 camera gradients sit 3e-3..9e-3 (of max|g|)
        # from an fp64 evaluation (PE frequencies up to 2^9 amplify round-off), and two fp32
        # evaluations with different accumulation order differ by up to ~2e-3.
        close(getattr(cam, k).grad, ref, 0, 5e-3 * np.abs(ref).max())
    rng = np.random.default_rng(99)
    for tag, P in (("coarse", Pc), ("fine", Pf)):
        for name, _fo, _fi, _a in synth.mlp_layer_shapes():
            for suffix in (".weight", ".bias"):
                key = name + suffix
                gr = P[key].grad.double().reshape(-1)
                probe = T(rng.standard_normal(gr.numel()))
                pin = g[f"gpin_{tag}_{key}"]
                got = np.array([gr.norm().item(), (gr * probe).sum().item(), gr.abs().max().item()])
                np.testing.assert_allclose(got, pin, rtol=5e-3, atol=5e-3 * pin[0])
                if f"g_{tag}_{key}" in g:
                    ref = g[f"g_{tag}_{key}"]
                    close(P[key].grad, ref, 0, 5e-3 * np.abs(ref).max())


@pytest.mark.parametrize("right", [False, True])
def test_searchsorted_rows_matches_torch(right):
    """Same grid as NeRF/torchsearchsorted/test/test_searchsorted.py:27-44 (subset of repeats)."""
    rng = np.random.default_rng(5)
    for Ba, Bv in ((1, 100), (100, 1), (100, 100)):
        for A in (1, 50, 500):
            for V in (1, 12, 120):
                a = np.sort(rng.random((Ba, A)).astype(np.float32), -1)
                v = rng.random((Bv, V)).astype(np.float32)
                got = O.searchsorted_rows(a, v, right)
                nrow = max(Ba, Bv)
                ref = torch.searchsorted(T(a).expand(nrow, A).contiguous(), T(v).expand(nrow, V).contiguous(), right=right)
                assert np.array_equal(got, ref.numpy())


# ---------------------------------------------------------------------------------------------
# NeRF++ rows: oracle/scnerf_pp_oracle.py vs tests/golden/pp_*.npz
# ---------------------------------------------------------------------------------------------
from oracle import scnerf_pp_oracle as OP   # noqa: E402

PH, PW, PF, PN = synth.PP_H, synth.PP_W, synth.PP_FOCAL, synth.PP_NCAM


def make_cam_pp(seed, requires_grad=False, dtype=torch.float32):
    cam = OP.CameraPP(synth.intrinsic_init(PH, PW, PF), synth.pp_camera_poses(seed), synth.pp_camera_args(),
                      PH, PW, k=(-0.05, 0.01), dtype=dtype)
    return cam.load(synth.camera_noise_state(seed, n_cams=PN, H=PH, W=PW, with_distortion=True), requires_grad)


def pp_nets(seed, dtype=torch.float32):
    cv = lambda st: {k: T(v).to(dtype) for k, v in st.items()}
    return cv(synth.pp_mlp_state(seed, 63)), cv(synth.pp_mlp_state(seed + 1, 84))


def test_pp_raygen(golden):
    g = golden("pp_raygen")
    for tag, seed in (("a", 30), ("b", 31)):
        cam = make_cam_pp(seed, requires_grad=True)
        o, d, depth = OP.rays_from_camera(cam, int(g[f"{tag}_cam_idx"]), T(g[f"{tag}_sel"]))
        close(o, g[f"{tag}_o"], 1e-6, 1e-7)
        close(d, g[f"{tag}_d"], 1e-5, 2e-7)
        close(depth, g[f"{tag}_depth"], 0, 0)
        ((o * T(g[f"{tag}_wo"])).sum() + (d * T(g[f"{tag}_wd"])).sum()).backward()
        for name in OP.CameraPP.LEARNABLE:
            ref = g[f"{tag}_g_{name}"]
            close(getattr(cam, name).grad, ref, 2e-3, 2e-5 * np.abs(ref).max())
    cam = make_cam_pp(32)
    o, d, depth = OP.rays_from_camera(cam, None, T(g["c_sel"]), extrinsic=g["c_E"])
    close(o, g["c_o"], 1e-6, 1e-7)
    close(d, g["c_d"], 1e-5, 2e-7)


def test_pp_sampling(golden):
    g = golden("pp_sampling")
    o, d = T(g["o"]), T(g["d"])
    far = OP.intersect_sphere(o, d)
    close(far, g["far"], 1e-6, 1e-7)
    fg, bg = OP.level0_depths(1e-4 * torch.ones_like(far), far, 32, T(g["t_fg"]), T(g["t_bg"]))
    close(fg, g["fg_p"], 1e-6, 1e-7)
    close(bg, g["bg_p"], 1e-6, 1e-7)
    mid = 0.5 * (T(g["fg_p"])[..., 1:] + T(g["fg_p"])[..., :-1])
    w = T(g["w"])[..., 1:-1]
    s, _ = OP.sample_pdf(mid, w, 64, u=T(g["u"]))
    close(s, g["s_rand"], 1e-6, 1e-7)
    s, _ = OP.sample_pdf(mid, w, 64, det=True)
    close(s, g["s_det"], 1e-6, 1e-7)
    merged = OP.level1_depths(T(g["fg_p"]), T(g["w"]), 64, u=T(g["u"]))
    close(merged, g["merged"], 1e-6, 1e-7)
    with pytest.raises(Exception, match="unit sphere"):
        OP.intersect_sphere(torch.tensor([[2.0, 0.0, 0.0]]), torch.tensor([[0.0, 1.0, 0.0]]))   # misses the sphere


def test_pp_field(golden):
    g = golden("pp_field")
    st_fg, st_bg = pp_nets(40)
    for v in list(st_fg.values()) + list(st_bg.values()):
        v.requires_grad_(True)
    o, d = T(g["o"]).requires_grad_(True), T(g["d"]).requires_grad_(True)
    far = OP.intersect_sphere(o, d)
    fg, bg = OP.level0_depths(1e-4 * torch.ones_like(far), far, 24, T(g["t_fg"]), T(g["t_bg"]))
    close(fg, g["fg"], 1e-6, 1e-7)
    pts4, depth_real = OP.depth2pts_outside(o.detach()[:, None, :].expand(48, 24, 3),
                                            d.detach()[:, None, :].expand(48, 24, 3), T(g["bg"]))
    close(pts4, g["pts4"], 1e-5, 1e-6)
    close(depth_real, g["depth_real"], 1e-4, 1e-4)
    ret = OP.nerfnet_forward(st_fg, st_bg, o, d, far, fg, bg)
    for k, v in ret.items():
        close(v, g["ret_" + k], 1e-4, 2e-6)
    loss = torch.mean((ret["rgb"] - T(g["target"])) ** 2)
    close(loss, g["loss"], 1e-5, 0)
    loss.backward()
    for k in list(g):
        if k.startswith("g_fg_net.") or k.startswith("g_bg_net."):
            st = st_fg if k.startswith("g_fg") else st_bg
            ref = g[k]
            close(st[k[9:]].grad[:8], ref, 5e-3, 5e-5 * np.abs(ref).max())
    close(o.grad, g["g_o"], 5e-3, 5e-5 * np.abs(g["g_o"]).max())
    close(d.grad, g["g_d"], 5e-3, 5e-5 * np.abs(g["g_d"]).max())


def test_pp_train_step(golden):
    g = golden("pp_train_step")
    cam = make_cam_pp(35, requires_grad=True)
    nets = [pp_nets(50), pp_nets(52)]
    for fgst, bgst in nets:
        for v in list(fgst.values()) + list(bgst.values()):
            v.requires_grad_(True)
    rand = {k: T(g[k]) for k in ("t_fg", "t_bg", "u_fg", "u_bg")}
    loss, rets, (fg1, bg1) = OP.train_step(cam, int(g["cam_idx"]), T(g["sel"]), T(g["target"]), nets, [24, 48], rand)
    close(rets[0]["rgb"], g["rgb0"], 1e-4, 2e-6)
    close(fg1, g["fg1"], 1e-5, 1e-6)
    close(bg1, g["bg1"], 1e-5, 1e-6)
    close(rets[1]["rgb"], g["rgb1"], 1e-4, 2e-6)
    close(loss, g["loss"], 1e-5, 0)
    loss.backward()
    for name in OP.CameraPP.LEARNABLE:
        ref = g["g_cam_" + name]
        close(getattr(cam, name).grad, ref, 1e-2, 1e-4 * np.abs(ref).max())
    for k in list(g):
        if k.startswith("g_net"):
            m, name = int(k[5]), k[7:]
            st = nets[m][0] if name.startswith("fg_net.") else nets[m][1]
            ref = g[k]
            close(st[name[7:]].grad[:8], ref, 1e-2, 1e-4 * np.abs(ref).max())


ADAM_CASES = (("dist", "pinhole_rot_noise_10k_rayo_rayd_dist", False, 0.1),
              ("od", "pinhole_rot_noise_10k_rayo_rayd", True, 0.05), ("none", "none", False, 0.1))


def test_custom_adam(golden):
    """oracle restatement of f_custom_adam vs the reference's CustomAdamOptimizer."""
    g = golden("adam")
    for tag, cam_name, amsgrad, wd in ADAM_CASES:
        p0, grads = synth.adam_case(1)
        params = [T(p.copy()) for p in p0]
        m = [torch.zeros_like(p) for p in params]
        v = [torch.zeros_like(p) for p in params]
        vmax = [torch.zeros_like(p) for p in params]
        for step, gs in enumerate(grads):
            lr = 5e-4 * (0.1 ** (step / 250000))
            params = O.custom_adam_step(params, [T(x) for x in gs], m, v, vmax, [step + 1] * len(params), cam_name,
                                        amsgrad=amsgrad, beta1=0.9, beta2=0.999, lr=lr, weight_decay=wd, eps=1e-8)
            for i, p in enumerate(params):
                close(p, g[f"{tag}_s{step}_p{i}"], 2e-6, 1e-7)


def test_prd_loss(golden):
    """oracle restatement of proj_ray_dist_loss_single vs the live reference."""
    g = golden("prd_loss")
    rays = [T(g["rays_" + k]).requires_grad_(True) for k in ("o0", "d0", "o1", "d1")]
    K, E = T(g["K"]), T(g["E"])
    E2 = E[[int(g["i"]), int(g["j"])]]
    kps0, kps1 = T(g["kps0"]).float(), T(g["kps1"]).float()
    loss, n = O.proj_ray_dist_loss(kps0, kps1, (rays[0], rays[1]), (rays[2], rays[3]), K, E2, 5.0, train=True)
    close(loss, g["nocam_loss"], 1e-5, 0)
    assert n == float(g["nocam_n_match"])
    loss.backward()
    for name, t in zip(("o0", "d0", "o1", "d1"), rays):
        ref = g["nocam_g_" + name]
        close(t.grad, ref, 1e-3, 1e-5 * np.abs(ref).max())
    with torch.no_grad():
        r = [t.detach() for t in rays]
        lv, none = O.proj_ray_dist_loss(kps0, kps1, (r[0], r[1]), (r[2], r[3]), K, E2, 5.0, train=False)
        lpp, _ = O.proj_ray_dist_loss(kps0, kps1, (r[0], r[1]), (r[2], r[3]), K, E2, 5.0, train=False, method="NeRF++")
    assert none is None
    close(lv, g["val_loss"], 1e-5, 0)
    close(lpp, g["val_loss_pp"], 1e-5, 0)


def test_raygen_subpixel(golden):
    """sub-pixel keypoints — direction from the float value, residual lookup from its truncation
    (NeRF/get_rays.py:112-123,134,140).  Oracle vs the live reference, forward and camera gradients."""
    g = golden("raygen_subpixel")
    cam = make_cam(12, requires_grad=True)
    kps, idx = synth.subpixel_kps(12, 256)
    o, d = O.rays_pixels_camera(H, W, cam, T(kps), idx=T(idx))
    close(o, g["o"]); close(d, g["d"])
    ((o * T(g["wo"])).sum() + (d * T(g["wd"])).sum()).backward()
    for k in O.Camera.LEARNABLE:
        ref = g["g_" + k]
        close(getattr(cam, k).grad, ref, 2e-3, 2e-5 * np.abs(ref).max())
    with torch.no_grad():
        o, d = O.rays_pixels_camera(H, W, cam, T(kps), idx=4)
        close(o, g["int_o"]); close(d, g["int_d"])
        o, d = O.rays_pixels_camera(H, W, cam, T(kps), extrinsic=T(synth.camera_poses(13)[2]))
        close(o, g["ext_o"]); close(d, g["ext_d"])


def test_c3_composed_step(golden):
    """BASELINE configs[2]: render + PRD (sub-pixel matches) + CustomAdamOptimizer + lr decay, two steps, composed as
    NeRF/run_nerf.py:482-621.  Oracle restatement vs the live reference."""
    from tests.util import oracle_c3_steps, CAM_KEYS
    g = golden("c3_step")
    r = oracle_c3_steps()
    close(r["loss_render"], float(g["loss1"]) + float(g["loss0"]), 2e-5, 0)
    close(r["prd"], g["prd"], 1e-4, 0)
    assert r["n_match"] == float(g["n_match"])
    close(r["total"], g["total"], 2e-5, 0)
    for k in CAM_KEYS:
        ref = g["g_cam_" + k]
        close(r["g_cam_" + k], ref, 5e-3, 2e-4 * np.abs(ref).max())
    for key in list(g):
        if key.startswith("gpin_"):
            ref = g[key]
            assert abs(r[key][0] - ref[0]) <= 2e-3 * ref[0] + 1e-12, (key, r[key], ref)
            assert abs(r[key][1] - ref[1]) <= 2e-3 * ref[0] * 30, (key, r[key], ref)       # projection on a unit-variance probe
    for step in range(synth.c3_case()["n_steps"]):
        for k in CAM_KEYS:
            # Adam's first steps move every element by ~lr regardless of the gradient's size: compare the UPDATE
            ref = g[f"s{step}_cam_" + k]
            # (step 0 moves by exactly lr x sign(g); from step 1 on m/sqrt(v) amplifies fp32 noise of small gradients)
            tol = (0.002 if step == 0 else 0.05) * synth.c3_case()["lrate"]
            assert np.abs(r[f"s{step}_cam_" + k] - ref).max() <= tol + 1e-6 * np.abs(ref).max(), k
        for key in list(g):
            if key.startswith(f"s{step}_ppin_"):
                ref = g[key]
                assert abs(r[key][0] - ref[0]) <= 1e-5 * ref[0], (key, r[key], ref)   # the update itself is ~1e-2 of the norm
