"""CPU emulation: how much accuracy the weight-gradient GEMM dW = dZ^T X loses if the tile images drop an operand half
or store the lo halves as fp8 (float64 sums of rounded operands, 60 K samples of the test network).  Quoted in DESIGN.md section 9."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import build_modules
torch.manual_seed(0)
mods = build_modules(3, "cpu")
net = mods["fine"]
sd = {k: v.detach().double() for k, v in net.state_dict().items()}
N = 60000
pts = (torch.rand(N, 3) * 2 - 1).double()
dirs = torch.nn.functional.normalize(torch.randn(N, 3), dim=-1).double()
def pe(x, L):
    out = [x]
    for f in range(L):
        out += [torch.sin(x * 2.0 ** f), torch.cos(x * 2.0 ** f)]
    return torch.cat(out, -1)
X = pe(pts, 10); V = pe(dirs, 4)
W = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
acts = {}
h = X
for i in range(8):
    acts[i] = h.detach()
    h = torch.relu(h @ W[f"pts_linears.{i}.weight"].t() + W[f"pts_linears.{i}.bias"])
    if i == 4: h = torch.cat([X, h], -1)
alpha = h @ W["alpha_linear.weight"].t() + W["alpha_linear.bias"]
feat = h @ W["feature_linear.weight"].t() + W["feature_linear.bias"]
hv = torch.relu(torch.cat([feat, V], -1) @ W["views_linears.0.weight"].t() + W["views_linears.0.bias"])
rgb = hv @ W["rgb_linear.weight"].t() + W["rgb_linear.bias"]
raw = torch.cat([rgb, alpha], -1)
# a loss whose d(raw) looks like a rendering loss: random per-sample weights of mixed sign and wide dynamic range
g_raw = torch.randn(N, 4).double() * torch.exp(torch.randn(N, 1).double() * 2) * 1e-4
# pre-activation gradients dZ_l via autograd hooks: recompute layer by layer
zs = {}
h = X
pre = []
for i in range(8):
    z = h @ sd[f"pts_linears.{i}.weight"].t() + sd[f"pts_linears.{i}.bias"]
    z.requires_grad_(True); z.retain_grad(); pre.append(z)
    h = torch.relu(z)
    if i == 4: h = torch.cat([X, h], -1)
alpha = h @ sd["alpha_linear.weight"].t() + sd["alpha_linear.bias"]
feat = h @ sd["feature_linear.weight"].t() + sd["feature_linear.bias"]
hv = torch.relu(torch.cat([feat, V], -1) @ sd["views_linears.0.weight"].t() + sd["views_linears.0.bias"])
rgb = hv @ sd["rgb_linear.weight"].t() + sd["rgb_linear.bias"]
raw2 = torch.cat([rgb, alpha], -1)
(raw2 * g_raw).sum().backward()
def bf(t): return t.float().bfloat16().double()
def split(t):
    h_ = t.float().bfloat16().float(); l_ = (t.float() - h_).bfloat16().float(); return h_.double(), l_.double()
print("layer : rel-to-max error of dW   [x3: dZ(hi+lo) x X(hi+lo), lo*lo dropped]   [dZ(hi+lo) x X(hi)]   [dZ(hi) x X(hi): bf16 single]")
for i in (1, 3, 5, 7):
    dZ = pre[i].grad; Xl = acts[i]
    exact = dZ.t() @ Xl
    dh, dl = split(dZ); xh, xl = split(Xl)
    x3 = dh.t() @ xh + dl.t() @ xh + dh.t() @ xl
    x2 = dh.t() @ xh + dl.t() @ xh
    x1 = dh.t() @ xh
    m = exact.abs().max()
    print(f"  {i}   : {((x3 - exact).abs().max() / m):.2e}   {((x2 - exact).abs().max() / m):.2e}   {((x1 - exact).abs().max() / m):.2e}    (max |dW| {m:.3e}, N = {N} samples)")
print("compact images: lo half stored as fp8 (e4m3 / e5m2), expanded to bf16 before the MMA; same three MMAs")
def f8(t, dt):
    return t.float().to(dt).float().double()
for dt, name in ((torch.float8_e4m3fn, "e4m3"), (torch.float8_e5m2, "e5m2")):
    for i in (1, 5, 7):
        dZ = pre[i].grad; Xl = acts[i]
        exact = dZ.t() @ Xl
        dh, dl = split(dZ); xh, xl = split(Xl)
        # fp8 has a narrow exponent range: scale lo by 2^k per tensor so that its max sits near the top (a per-image scale would do the same)
        def q(l):
            sc = 2.0 ** torch.floor(torch.log2(l.abs().max() / 200.0))
            return f8(l / sc, dt) * sc
        xl8, dl8 = q(xl), q(dl)
        m = exact.abs().max()
        a = dh.t() @ xh + dl.t() @ xh + dh.t() @ xl8
        b = dh.t() @ xh + dl8.t() @ xh + dh.t() @ xl8
        print(f"  {name} layer {i}: X lo in fp8 {((a - exact).abs().max() / m):.2e}   X lo and dZ lo in fp8 {((b - exact).abs().max() / m):.2e}")
