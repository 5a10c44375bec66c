import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scnerf_b200 import _lib
lib = _lib.load()
# (mode, name, nominal cycles, MMAs per iteration)
MODES = [(0, "SS K-major N=256", 128, 16), (1, "TS (A in TMEM) N=256", 128, 16), (2, "SS MN/MN (wgrad layout)", 128, 16),
         (3, "SS MN-A / K-B", 128, 16)]
PAT = {0: "TS same acc", 1: "TS alt acc", 2: "SS", 3: "TS/SS/TS", 4: "SS/TS alternating"}
for n in (256, 128, 64, 32):
    for pat in (0, 1, 2, 3, 4):
        if pat == 1 and n > 128:
            continue
        MODES.append((256 + ((n // 8) << 4) + pat, f"probe N={n} {PAT[pat]}", n // 2, 24))
for nb in (1, 148):
    for mode, name, nominal, per_it in MODES:
        if nb == 148 and mode >= 256 and (mode & 15) not in (0, 2, 3):
            continue
        out = torch.zeros(nb + 2 + 8192, dtype=torch.int64, device="cuda")
        iters = 256
        _lib.check(lib.scnerf_debug_mma_bench(mode, iters, _lib.ptr(out), nb, _lib.stream()))
        torch.cuda.synchronize()
        c = out[:nb].float().mean().item() / (iters * per_it)
        print(f"blocks={nb:3d} {name:38s}: {c:7.1f} cycles per MMA (nominal {nominal})", flush=True)
