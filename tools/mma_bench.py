import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scnerf_b200 import _lib
lib = _lib.load()
for nb in (1, 148):
    for mode, name in ((0, "SS K-major"), (1, "TS (A in TMEM)"), (2, "SS MN/MN (wgrad layout)"), (3, "SS MN-A / K-B")):
        out = torch.zeros(nb, dtype=torch.int64, device="cuda")
        iters = 256
        _lib.check(lib.scnerf_debug_mma_bench(mode, iters, _lib.ptr(out), nb, _lib.stream()))
        torch.cuda.synchronize()
        c = out.float().mean().item() / (iters * 16)
        print(f"blocks={nb:3d} {name:26s}: {c:7.1f} cycles per 128x256x16 MMA (nominal 128)")
