"""Diagnostic: per-tensor gradient error of the split-bf16 NeRF++ step (cascade 64,128) against the fp64 oracle at the
CUDA path's own samples, next to the fp32 CUDA-core path and the fp32 oracle floor.
  python tools/pp_grad_precision.py     (round 2 ran it with the serial and the pipelined dgrad chain; the serial one is retired)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests.test_gpu_parity_pp import _pp_cuda_step, _pp_oracle_step, relmax
torch.set_num_threads(32)
N, cascade, seed = 256, [64, 128], 70
res = {}
for prec in ("bf16x3", "fp32"):
    loss, rgbs, grads, own = _pp_cuda_step(seed, N, cascade, prec)
    _, _, g64 = _pp_oracle_step(seed, N, cascade, torch.float64, level1_override=own)
    res[prec] = {k: relmax(grads[k], g64[k]) for k in g64}
_, _, g32 = _pp_oracle_step(seed, N, cascade, torch.float32)
_, _, g64f = _pp_oracle_step(seed, N, cascade, torch.float64)
floor = {k: relmax(g32[k], g64f[k]) for k in g64f}
print("tensor: bf16x3 | fp32 cuda | fp32 oracle floor")
for k in sorted(floor):
    if "weight" in k and ("base_layers" in k or "remap" in k or "rgb_layers.0" in k):
        print(f"  {k:42s} {res['bf16x3'][k]:.2e}  {res['fp32'][k]:.2e}  {floor[k]:.2e}")
