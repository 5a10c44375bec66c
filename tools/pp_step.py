"""A few fused NeRF++ steps (scnerf_pp_train_step) for profilers: python tools/pp_step.py [precision] [rays] [Nc] [Nf]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scnerf_b200 import synth
from scnerf_b200.nerfplusplus.engine import PPTrainStep
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
N, Nc, Nf = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((2, 4096), (3, 64), (4, 128)))
mods = synth.build_pp_modules(0, "cuda:0", levels=2, precision=prec)
sel, cam_idx, target = synth.pp_pixel_batch(1000, N)
eng = PPTrainStep(mods["cam"], mods["nets"], N, [Nc, Nf], camera_idx=cam_idx, precision=prec)
eng.step_device(torch.from_numpy(sel).cuda(), torch.from_numpy(target).cuda())
for _ in range(2):
    eng.step_device()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.step_device()
e1.record(); torch.cuda.synchronize()
print(prec, f"NeRF++ step {N} x ({Nc},{Nf}): {e0.elapsed_time(e1) / 3:.3f} ms, loss {float(eng.loss_dev):.6f}")
