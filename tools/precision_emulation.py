"""CPU emulation of the operand-precision schemes of the tensor-core field (float64 accumulation of rounded operands):
which split meets the 1e-4 gate?  Run on the CPU: python tools/precision_emulation.py  (numbers quoted in DESIGN.md section 3)."""
import sys, torch, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import build_modules
from oracle import scnerf_oracle as O
torch.manual_seed(0)
mods = build_modules(3, "cpu")
net = mods["coarse"]
sd = {k: v.detach().double() for k, v in net.state_dict().items()}
print(list(sd.keys())[:6])
N = 20000
pts = (torch.rand(N, 3) * 3 - 1.5).double()
dirs = torch.nn.functional.normalize(torch.randn(N, 3), dim=-1).double()
def pe(x, L):
    out = [x]
    for f in range(L):
        out += [torch.sin(x * 2.0 ** f), torch.cos(x * 2.0 ** f)]
    return torch.cat(out, -1)
X = pe(pts, 10); V = pe(dirs, 4)
def rnd(t, fmt):
    if fmt == "exact": return t
    if fmt == "bf16": return t.float().bfloat16().double()
    if fmt == "fp16": return t.float().half().double()
    if fmt == "bf16x2":
        h = t.float().bfloat16().float(); l = (t.float() - h).bfloat16().float(); return (h + l).double()
    if fmt == "fp16x2":
        h = t.float().half().float(); l = (t.float() - h).half().float(); return (h + l).double()
def lin(a, W, b, fa, fw):
    # fp32 accumulate emulated in float64 of rounded operands (accumulation error is not the question here)
    return rnd(a, fa) @ rnd(W, fw).t() + b
def forward(fa, fw):
    h = X
    for i in range(8):
        W, b = sd[f"pts_linears.{i}.weight"], sd[f"pts_linears.{i}.bias"]
        h = torch.relu(lin(h, W, b, fa, fw))
        if i == 4: h = torch.cat([X, h], -1)
    alpha = h @ sd["alpha_linear.weight"].t() + sd["alpha_linear.bias"]
    feat = lin(h, sd["feature_linear.weight"], sd["feature_linear.bias"], fa, fw)
    hv = torch.relu(lin(torch.cat([feat, V], -1), sd["views_linears.0.weight"], sd["views_linears.0.bias"], fa, fw))
    rgb = hv @ sd["rgb_linear.weight"].t() + sd["rgb_linear.bias"]
    return torch.cat([rgb, alpha], -1)
ref = forward("exact", "exact")
for fa, fw, name in (("bf16", "bf16", "bf16 single pass (1 MMA)"), ("bf16x2", "bf16x2", "bf16x3 (hi*hi + lo*hi + hi*lo; lo*lo dropped ~ both 16 bits)"),
                     ("fp16x2", "fp16", "fp16 split activations x fp16 weights (2 MMAs)"), ("fp16", "fp16x2", "fp16 activations x fp16 split weights (2 MMAs)"),
                     ("fp16", "fp16", "fp16 single pass (1 MMA)")):
    out = forward(fa, fw)
    e = (out - ref).abs().max() / ref.abs().max()
    er = (torch.sigmoid(out[:, :3]) - torch.sigmoid(ref[:, :3])).abs().max()
    print(f"{name:70s}: raw rel-to-max err {e:.2e}   max |d sigmoid(rgb)| {er:.2e}")

# ---- correction terms (lo x hi, hi x lo) as FP8 MMAs (2x rate): main term bf16 -----------------------------------
def bf(t): return t.float().bfloat16().double()
def f8(t, dt=torch.float8_e4m3fn):
    sc = 2.0 ** torch.floor(torch.log2(t.abs().max().clamp_min(1e-30) / 200.0))
    return (t / sc).float().to(dt).float().double() * sc
def mm(a, W, mode):
    ah, wh = bf(a), bf(W); al, wl = bf(a - ah), bf(W - wh)
    if mode == "exact": return a @ W.t()
    if mode == "x3": return ah @ wh.t() + al @ wh.t() + ah @ wl.t()
    if mode == "fp8corr": return ah @ wh.t() + f8(al) @ f8(wh).t() + f8(ah) @ f8(wl).t()
    if mode == "fp8corr_rowscale":   # per-row scales for the activation operands (one scale per sample)
        def f8r(t):
            sc = 2.0 ** torch.floor(torch.log2(t.abs().amax(-1, keepdim=True).clamp_min(1e-30) / 200.0))
            return (t / sc).float().to(torch.float8_e4m3fn).float().double() * sc
        return ah @ wh.t() + f8r(al) @ f8(wh).t() + f8r(ah) @ f8(wl).t()
def forward(mode):
    h = X
    for i in range(8):
        h = torch.relu(mm(h, sd[f"pts_linears.{i}.weight"], mode) + sd[f"pts_linears.{i}.bias"])
        if i == 4: h = torch.cat([X, h], -1)
    alpha = h @ sd["alpha_linear.weight"].t() + sd["alpha_linear.bias"]
    feat = mm(h, sd["feature_linear.weight"], mode) + sd["feature_linear.bias"]
    hv = torch.relu(mm(torch.cat([feat, V], -1), sd["views_linears.0.weight"], mode) + sd["views_linears.0.bias"])
    rgb = hv @ sd["rgb_linear.weight"].t() + sd["rgb_linear.bias"]
    return torch.cat([rgb, alpha], -1)
ref = forward("exact")
print("split-bf16 with the two correction terms in fp8 (e4m3):")
for mode in ("x3", "fp8corr", "fp8corr_rowscale"):
    out = forward(mode)
    print(f"{mode:18s}: raw rel-to-max err {((out - ref).abs().max() / ref.abs().max()):.2e}   max |d sigmoid(rgb)| {((torch.sigmoid(out[:, :3]) - torch.sigmoid(ref[:, :3])).abs().max()):.2e}")
