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
b = big.cpu().numpy()
tl = b[:T * 8 * 4].reshape(T * 8, 4)
t0 = tl[tl > 0].min()
print("every 32nd slot: prod_free  mma_full  mma_commit helper_release | full-free  commit-full release-full cycles/slot since previous stamp")
for i in range(0, T * 8 - 1):
    if tl[i, 0] == 0 or tl[i + 1, 0] == 0: break
    f, m, c, h = tl[i] - t0
    if i % 8 == 0: print(f"{i * 32:5d}: {f:9d} {m:9d} {c:9d} {h:9d} | {m - f:6d} {c - m:6d} {h - m:6d} {(tl[i + 1, 0] - tl[i, 0]) / 32:8.0f}")
nct = 148
c = b[T * 8 * 4: T * 8 * 4 + nct * 4].reshape(nct, 4)
g0 = c[:, 0].min()
import numpy as np
alloc = [int(x) for x in os.environ.get("SCNERF_WGRAD_ALLOC", "16,14,14,14,14,14,14,14,20,14").split(",")]   # field_tc.cuh: CTAs per job on a 148-SM part
print("per job: start_us  loop_end_us  end_us  sm_cycles (min..max over the job's CTAs)")
o = 0
for j, n in enumerate(alloc):
    rows = c[o:o + n]; o += n
    print(f"job {j} ({n:2d} CTAs): start {np.min(rows[:,0]-g0)/1e3:7.1f}..{np.max(rows[:,0]-g0)/1e3:7.1f}  loop_end {np.min(rows[:,1]-g0)/1e3:7.1f}..{np.max(rows[:,1]-g0)/1e3:7.1f}"
          f"  end {np.min(rows[:,2]-g0)/1e3:7.1f}..{np.max(rows[:,2]-g0)/1e3:7.1f}  cycles {rows[:,3].min()}..{rows[:,3].max()}")
