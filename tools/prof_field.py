"""Profile target: the fused field forward alone (inference), fine-pass size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import build_modules
from scnerf_b200.create_nerf import run_network
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mods = build_modules(0, "cuda:0")
pts = torch.rand(N, 192, 3, device="cuda") * 2 - 1
vd = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda"), dim=-1)
for _ in range(3):
    run_network(pts, vd, mods["fine"], None, None, precision=prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run_network(pts, vd, mods["fine"], None, None, precision=prec)
e1.record(); torch.cuda.synchronize()
print(prec, "field fwd ms", e0.elapsed_time(e1) / 5)
