"""Time the big kernels of one training step with CUDA events around the three C-ABI phases is not
possible from outside; instead run the engine step N times and report ms/step for a precision."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.step_device()
e1.record(); torch.cuda.synchronize()
print(prec, "train step ms", e0.elapsed_time(e1) / 5, "loss", float(eng.loss_dev))
