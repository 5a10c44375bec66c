"""In-kernel timeline of the N-half pipelined fused forward (CTA 0, first tiles): [tile][stage][8] clock64 stamps
0 MMA passed a1, 1 MMA passed a2, 2 MMA committed h0, 3 MMA committed h1,
4 epilogue saw accf0, 5 epilogue arrived a1, 6 epilogue saw accf1, 7 epilogue arrived a2."""
import sys, os
os.environ["SCNERF_DBG_MIN_TILES"] = "1000000000"   # keep the wgrad kernel (same debug pointer, other layout) from stamping
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tests.util import build_modules
from scnerf_b200.create_nerf import run_network
from scnerf_b200 import _lib
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
lib = _lib.load()
mods = build_modules(0, "cuda:0")
TRAIN = len(sys.argv) > 2 and sys.argv[2] == "train"     # training step: the fine-pass forward (image dumps on) stamps last
T = 4
buf = torch.zeros(T * 10 * 16 + 2 * 160, dtype=torch.int64, device="cuda")     # + per-CTA (end time, tiles) pairs
if TRAIN:
    from scnerf_b200 import synth
    from scnerf_b200.engine import TrainStep
    kps, idx, target = (torch.from_numpy(x).cuda() for x in synth.pixel_batch(1000, 4096))
    eng = TrainStep(mods["cam"], mods["coarse"], mods["fine"], 4096, 64, 128, precision=prec)
    eng.step_device(kps, idx, target)
    lib.scnerf_debug_timeline(_lib.ptr(buf), T)
    eng.step_device()
else:
    N = 4096
    pts = torch.rand(N, 192, 3, device="cuda") * 2 - 1
    vd = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda"), dim=-1)
    run_network(pts, vd, mods["fine"], None, None, precision=prec)
    lib.scnerf_debug_timeline(_lib.ptr(buf), T)
    run_network(pts, vd, mods["fine"], None, None, precision=prec)
torch.cuda.synchronize()
lib.scnerf_debug_timeline(None, 0)
raw = buf.cpu().numpy()
b = raw[:T * 10 * 16].reshape(T, 10, 16)
ends = raw[T * 10 * 16:].reshape(-1, 2)
ends = ends[ends[:, 0] > 0]
stamps = np.delete(b, [11, 15], axis=2)      # slots 11 / 15 are cycle counts, not stamps
t0 = stamps[stamps > 0].min()
print(prec, "TRAIN" if TRAIN else "INFER", "stage: mma_a1 mma_a2 commit0 commit1 | epi_accf0 epi_a1 epi_accf1 epi_a2   (cycles from the first stamp)")
for t in range(1, 3):
    for s in range(10):
        r = b[t, s] - t0
        print(f"tile {t} stage {s}: " + " ".join(f"{int(x):8d}" for x in r[:4]) + " | " + " ".join(f"{int(x):8d}" for x in r[4:11]) +
              f" | mma h0 {int(r[2] - r[0]):6d} h1 {int(r[3] - r[2]):6d} epi0 {int(r[5] - r[4]):6d} epi1 {int(r[7] - r[6]):6d}"
              f" | epi0 quarters handed over at +{int(r[9] - r[4]):5d} +{int(r[10] - r[4]):5d}"
              f" | epi1 at +{int(r[13] - r[6]):5d} +{int(r[14] - r[6]):5d} | weight-ring waits h0 {int(b[t, s, 11]):5d} h1 {int(b[t, s, 15]):5d}")
    print(f"tile {t} total cycles: {b[t + 1, 0, 0] - b[t, 0, 0]}")
if len(ends):
    e = (ends[:, 0] - ends[:, 0].min()) / 1e3
    for n in sorted(set(ends[:, 1])):
        m = ends[:, 1] == n
        print(f"CTAs with {int(n)} tiles: {int(m.sum()):3d}; end time relative to the first CTA to finish: min {e[m].min():7.1f}  mean {e[m].mean():7.1f}  max {e[m].max():7.1f} us")
    print(f"all {len(e)} CTAs: last - first = {e.max():.1f} us, mean - first = {e.mean():.1f} us")
