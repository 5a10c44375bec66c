"""In-kernel timeline of the fused forward in TRAINING mode (image dumps on): last fwd launch of a step = fine pass."""
import sys, os
os.environ.setdefault("SCNERF_FWD_PIPE", "0")   # this tool reads the serial kernel's 4-stamp layout
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
T = 4
buf = torch.zeros(T, 10, 4, dtype=torch.int64, device="cuda")
lib.scnerf_debug_timeline(_lib.ptr(buf), T)
eng.step_device()
torch.cuda.synchronize()
lib.scnerf_debug_timeline(None, 0)
b = buf.cpu().numpy(); t0 = b[b > 0].min()
print(prec, "TRAIN stage: mma_issue_span  issue->epi_start  epi_span  epi_done->next_mma_start")
for t in (2,):
    for s in range(10):
        m0, m1, e0, e1 = (b[t, s] - t0)
        nxt = (b[t, s + 1, 0] - t0) if s < 9 else (b[t + 1, 0, 0] - t0)
        print(f"tile {t} stage {s}: {m1 - m0:6d} {e0 - m1:6d} {e1 - e0:6d} {nxt - e1:6d}")
    print(f"tile {t} total cycles: {b[t + 1, 0, 0] - b[t, 0, 0]}")
