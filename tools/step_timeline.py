"""Start/end timeline of one training step from CUPTI (torch.profiler): shows the gaps between kernels
(launch latency, memsets, pack kernels) that the per-kernel sums hide.
  python tools/step_timeline.py [precision]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from scnerf_b200 import synth
from scnerf_b200.engine import TrainStep
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
from tests.util import build_modules
mods = build_modules(0, "cuda:0")
kps, idx, target = (torch.from_numpy(x).cuda() for x in synth.pixel_batch(1000, 4096))
eng = TrainStep(mods["cam"], mods["coarse"], mods["fine"], 4096, 64, 128, precision=prec)
for _ in range(3):
    eng.step_device(kps, idx, target)
torch.cuda.synchronize()
STEPS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        eng.step_device()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
n = len(evs) // STEPS
sel = evs[n:2 * n]            # the middle step
prev_end = sel[0].time_range.start
busy = gap = 0.0
print(f"{prec}: {n} device activities per step; middle step:")
for e in sel:
    s, t = e.time_range.start, e.time_range.end
    g = s - prev_end
    gap += max(g, 0); busy += t - s
    print(f"  +{(s - sel[0].time_range.start):9.1f} us  gap {g:7.1f}  dur {t - s:8.1f}  {e.name.split('(')[0].replace('scnerf::', '')[:70]}")
    prev_end = max(prev_end, t)
span = sel[-1].time_range.end - sel[0].time_range.start
print(f"span {span:.1f} us, busy {busy:.1f} us, gaps {gap:.1f} us; step period {(evs[2 * n].time_range.start - evs[n].time_range.start):.1f} us")
