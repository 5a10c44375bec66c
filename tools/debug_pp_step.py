"""NeRF++ train-step camera gradients on the GPU vs the fp64 oracle for several cascade configurations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_parity_pp import make_cam, make_net, CAM_NAMES, relmax, PH, PW, PF, PN, T, DEV
from scnerf_b200 import synth
from scnerf_b200.nerfplusplus import intersect_sphere, render_ray_from_camera
from scnerf_b200.nerfplusplus.ddp_train_nerf import level0_depths, level1_depths
from oracle import scnerf_pp_oracle as OP

def run(cascade, N=40, seed=35):
    sel, ci, target = synth.pp_pixel_batch(seed, N)
    rng = np.random.default_rng(5)
    rand = {"t_fg": rng.uniform(0, 1, (N, cascade[0])).astype(np.float32), "t_bg": rng.uniform(0, 1, (N, cascade[0])).astype(np.float32)}
    if len(cascade) > 1:
        rand["u_fg"] = rng.uniform(0, 1, (N, cascade[1])).astype(np.float32); rand["u_bg"] = rng.uniform(0, 1, (N, cascade[1])).astype(np.float32)
    cam = make_cam(seed); nets = [make_net(50), make_net(52)][:len(cascade)]
    tgt = T(target).to(DEV); loss = 0.0
    for m in range(len(cascade)):
        o, d, _ = render_ray_from_camera(cam, ci, sel, DEV)
        if m == 0:
            far = intersect_sphere(o, d)
            fg, coef, bg = level0_depths(far, cascade[0], 1e-4, T(rand["t_fg"]).to(DEV), T(rand["t_bg"]).to(DEV))
        else:
            fg, coef = level1_depths(fg, ret["fg_weights"], cascade[1], fg_far_depth=far, coef=coef, u=T(rand["u_fg"]).to(DEV))
            bg, _ = level1_depths(bg, ret["bg_weights"], cascade[1], u=T(rand["u_bg"]).to(DEV))
        ret = nets[m](o, d, far, fg, bg)
        loss = loss + torch.mean((ret["rgb"] - tgt) ** 2)
    loss.backward()
    def oracle(dt, ov):
        c = OP.CameraPP(synth.intrinsic_init(PH, PW, PF), synth.pp_camera_poses(seed), synth.pp_camera_args(), PH, PW, k=(-0.05, 0.01), dtype=dt)
        c.load(synth.camera_noise_state(seed, n_cams=PN, H=PH, W=PW, with_distortion=True), True)
        cv = lambda st: {k: T(v).to(dt) for k, v in st.items()}
        ns = [(cv(synth.pp_mlp_state(s, 63)), cv(synth.pp_mlp_state(s + 1, 84))) for s in (50, 52)][:len(cascade)]
        l, _, _ = OP.train_step(c, ci, T(sel), T(target).to(dt), ns, cascade, {k: T(v).to(dt) for k, v in rand.items()}, level1_override=ov)
        l.backward(); return {k: getattr(c, k).grad.double().numpy() for k in CAM_NAMES}, float(l)
    ov64 = tuple(x.detach().cpu().double() for x in (fg, coef, bg)) if len(cascade) > 1 else None
    ov32 = tuple(x.float() for x in ov64) if ov64 else None
    g64, l64 = oracle(torch.float64, ov64); g32, l32 = oracle(torch.float32, ov32)
    print(f"cascade {cascade}: loss cuda {float(loss):.7f} fp64 {l64:.7f} fp32-oracle {l32:.7f}")
    for k in CAM_NAMES:
        print(f"   {k:18s} cuda-vs-fp64 {relmax(getattr(cam, k).grad, g64[k]):.2e}   fp32-oracle-vs-fp64 {relmax(g32[k], g64[k]):.2e}")

for c in ([24], [72], [24, 48], [12, 12], [16, 8]):
    run(c)
