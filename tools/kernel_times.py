"""Per-kernel durations of a training step from CUPTI (torch.profiler): no replay, kernels run back to back as in the bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from tests.util import build_modules
from scnerf_b200 import synth
from scnerf_b200.engine import TrainStep
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
mods = build_modules(0, "cuda:0")
kps, idx, target = (torch.from_numpy(x).cuda() for x in synth.pixel_batch(1000, 4096))
eng = TrainStep(mods["cam"], mods["coarse"], mods["fine"], 4096, 64, 128, precision=prec)
for _ in range(3):
    eng.step_device(kps, idx, target)
torch.cuda.synchronize()
STEPS = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        eng.step_device()
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0].replace("scnerf::", "")
        rows.setdefault(name, []).append(ev.device_time if hasattr(ev, "device_time") else ev.cuda_time)
tot = sum(sum(v) for v in rows.values()) / STEPS
print(f"{prec}: sum of kernel times per step {tot / 1000:.3f} ms")
for name, v in sorted(rows.items(), key=lambda kv: -sum(kv[1])):
    if sum(v) / STEPS < 20:
        continue
    per = sorted(v)
    print(f"  {sum(v) / STEPS / 1000:8.3f} ms/step  x{len(v) // STEPS}  each: " + " ".join(f"{x / 1000:.3f}" for x in sorted(set(round(y) for y in v[:len(v) // STEPS]))) + f"   {name[:90]}")
