"""NeRF++ training step (ddp_train_nerf.py:421-488 through the host mirror) at BASELINE configs[3]-like sizes:
N rays x cascade (64, 128), learnable distortion camera, fg and bg fields on the tensor-core kernels (bf16x3 / bf16)
or the fp32 CUDA-core kernels.  Prints ms/step and a per-phase breakdown (CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_parity_pp import make_cam, make_net, DEV, T
from scnerf_b200 import synth, _lib
from scnerf_b200.nerfplusplus import intersect_sphere, render_ray_from_camera
from scnerf_b200.nerfplusplus.ddp_train_nerf import level0_depths, level1_depths

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
cascade = [64, 128]
cam = make_cam(35)
nets = [make_net(50, prec), make_net(52, prec)]
sel, ci, target = synth.pp_pixel_batch(35, N)
tgt = T(target).to(DEV)
lib = _lib.load()

def step():
    for m in (cam, *nets):
        m.zero_grad(set_to_none=True)
    loss = 0.0
    for m in range(2):
        o, d, _ = render_ray_from_camera(cam, ci, sel, DEV)
        if m == 0:
            far = intersect_sphere(o, d)
            fg, coef, bg = level0_depths(far, cascade[0], 1e-4)
        else:
            fg, coef = level1_depths(fg, ret["fg_weights"], cascade[1], fg_far_depth=far, coef=coef)
            bg, _ = level1_depths(bg, ret["bg_weights"], cascade[1])
        ret = nets[m](o, d, far, fg, bg)
        loss = loss + torch.mean((ret["rgb"] - tgt) ** 2)
    loss.backward()
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
lib.scnerf_launch_count(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for _ in range(K):
    loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"NeRF++ step: {N} rays x (64,128) fg/bg, fg precision {prec}: {ms:.2f} ms/step = {N / ms * 1e3:.0f} rays/s, "
      f"{lib.scnerf_launch_count(0) // K} kernel launches/step, loss {float(loss):.5f}")
