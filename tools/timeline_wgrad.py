"""In-kernel timeline of the wgrad kernel (CTA 0 = job 0 slice 0) during a training step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import build_modules
from scnerf_b200 import synth, _lib
from scnerf_b200.engine import TrainStep
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
lib = _lib.load()
mods = build_modules(0, "cuda:0")
kps, idx, target = (torch.from_numpy(x).cuda() for x in synth.pixel_batch(1000, 4096))
eng = TrainStep(mods["cam"], mods["coarse"], mods["fine"], 4096, 64, 128, precision=prec)
eng.step_device(kps, idx, target)
T = 16
buf = torch.zeros(T * 8, 4, dtype=torch.int64, device="cuda")   # fwd kernel also writes [tile][10][4] into it: ignore
lib.scnerf_debug_timeline(_lib.ptr(torch.zeros(T * 10 * 4 + 8, dtype=torch.int64, device="cuda")), 0)
big = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib.scnerf_debug_timeline(_lib.ptr(big), T)
eng.step_device()
torch.cuda.synchronize()
lib.scnerf_debug_timeline(None, 0)
b = big.cpu().numpy()[:T * 8 * 4].reshape(T * 8, 4)
t0 = b[b > 0].min()
print("slot: prod_free  mma_full  mma_commit  helper_release | full-free  commit-full  next_free-free")
for i in range(40, 72):
    f, m, c, h = b[i] - t0
    print(f"{i:3d}: {f:8d} {m:8d} {c:8d} {h:8d} | {m - f:6d} {c - m:6d} {b[i + 1, 0] - b[i, 0]:6d}")
