"""Host-side breakdown of the composed configs[2] step (bench.py --workload c3): where the time between two render
steps goes.  Every section is bracketed by a device synchronise, so the numbers are latency, not overlap."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scnerf_b200 import synth
from scnerf_b200.custom_optim import update_lrate
from scnerf_b200.get_rays import get_rays_kps_use_camera
from scnerf_b200.ray_dist_loss import proj_ray_dist_loss_single
wl = bench.NerfWorkload("c3", 4096, 0, "cuda:0", "bf16x3")
for _ in range(3):
    wl.step(False)
torch.cuda.synchronize()
T = {}
def sec(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return r
C, cam, eng = wl.C, wl.mods["cam"], wl.eng
H, W = synth.FERN_H, synth.FERN_W
i, j = C["pair"]
R = 10
for _ in range(R):
    sec("zero", wl.grads.zero_)
    ri = sec("get_rays i", lambda: get_rays_kps_use_camera(H=H, W=W, camera_model=cam, idx_in_camera_param=i, kps_list=wl.kps0))
    rj = sec("get_rays j", lambda: get_rays_kps_use_camera(H=H, W=W, camera_model=cam, idx_in_camera_param=j, kps_list=wl.kps1))
    prd = sec("prd fwd", lambda: proj_ray_dist_loss_single(kps0_list=wl.kps0, kps1_list=wl.kps1, img_idx0=i, img_idx1=j, rays0=ri, rays1=rj,
              mode="train", device=wl.kps0.device, H=H, W=W, args=wl.args, camera_model=cam, method="NeRF", i_map=np.arange(synth.FERN_NCAM))[0])
    sec("prd bwd", lambda: (C["prd_weight"] * prd).backward())
    sec("render step", lambda: eng.step_device(zero=False))
    sec("all_reduce", wl.grads.all_reduce_mean)
    sec("adam", wl.opt.step)
    sec("lr", lambda: update_lrate(wl.opt, C["lrate"], C["lrate_decay"], wl.global_step))
for k, v in T.items():
    print(f"  {k:14s} {v / R:8.3f} ms")
print(f"  total          {sum(T.values()) / R:8.3f} ms")
