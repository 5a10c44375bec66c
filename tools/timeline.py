"""Print the in-kernel timeline of the fused forward (CTA 0, first tiles)."""
import sys, os
os.environ.setdefault("SCNERF_FWD_PIPE", "0")   # this tool reads the serial kernel's 4-stamp layout
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tests.util import build_modules
from scnerf_b200.create_nerf import run_network
from scnerf_b200 import _lib
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
lib = _lib.load()
mods = build_modules(0, "cuda:0")
N = 4096
pts = torch.rand(N, 192, 3, device="cuda") * 2 - 1
vd = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda"), dim=-1)
run_network(pts, vd, mods["fine"], None, None, precision=prec)
T = 4
buf = torch.zeros(T, 10, 4, dtype=torch.int64, device="cuda")
lib.scnerf_debug_timeline(_lib.ptr(buf), T)
run_network(pts, vd, mods["fine"], None, None, precision=prec)
torch.cuda.synchronize()
lib.scnerf_debug_timeline(None, 0)
b = buf.cpu().numpy()
t0 = b[b > 0].min()
print(prec, "stage: mma_start  mma_issued  epi_start  epi_done | mma_issue_span  issue->epi_start  epi_span  epi_done->next_mma_start")
for t in range(1, 3):
    for s in range(10):
        m0, m1, e0, e1 = (b[t, s] - t0)
        nxt = (b[t, s + 1, 0] - t0) if s < 9 else (b[t + 1, 0, 0] - t0)
        print(f"tile {t} stage {s}: {m0:8d} {m1:8d} {e0:8d} {e1:8d} | {m1 - m0:6d} {e0 - m1:6d} {e1 - e0:6d} {nxt - e1:6d}")
    print(f"tile {t} total cycles: {b[t + 1, 0, 0] - b[t, 0, 0]}")
