"""Piecewise gradient checks of the NeRF++ kernels against autograd through the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scnerf_b200 import _lib, synth
from scnerf_b200.nerfplusplus import depth2pts_outside
from oracle import scnerf_pp_oracle as OP
lib = _lib.load()
g = np.load("tests/golden/pp_field.npz")
T = torch.from_numpy
DEV = "cuda:0"
rel = lambda a, b: float((a.detach().cpu() - b.detach().cpu()).abs().max() / b.detach().abs().max())
N, S = 48, 24
rng = np.random.default_rng(0)
# ---- depth2pts_outside
o_c, d_c = T(g["o"]).requires_grad_(True), T(g["d"]).requires_grad_(True)
bg = T(g["bg"])
wp = T(rng.standard_normal((N, S, 4)).astype(np.float32))
p_c, _ = OP.depth2pts_outside(o_c[:, None, :].expand(N, S, 3), d_c[:, None, :].expand(N, S, 3), bg)
(p_c * wp).sum().backward()
o_g, d_g = T(g["o"]).to(DEV).requires_grad_(True), T(g["d"]).to(DEV).requires_grad_(True)
p_g, _ = depth2pts_outside(o_g[:, None, :].expand(N, S, 3), d_g[:, None, :].expand(N, S, 3), bg.to(DEV))
(p_g * wp.to(DEV)).sum().backward()
print("depth2pts fwd", rel(p_g, p_c), " g_o", rel(o_g.grad, o_c.grad), " g_d", rel(d_g.grad, d_c.grad))
# ---- fg compositing
raw = T(rng.standard_normal((N, S, 4)).astype(np.float32)); raw[..., 3] *= 3.0; raw = raw.contiguous()
fz_c = T(g["fg"]).requires_grad_(True); zmax_c = T(g["far"]).requires_grad_(True)
d_c = T(g["d"]).requires_grad_(True); raw_c = raw.clone().requires_grad_(True)
wr, wl = T(rng.standard_normal((N, 3)).astype(np.float32)), T(rng.standard_normal(N).astype(np.float32))
dn = torch.norm(d_c, dim=-1, keepdim=True)
dists = dn * torch.cat((fz_c[..., 1:] - fz_c[..., :-1], zmax_c.unsqueeze(-1) - fz_c[..., -1:]), -1)
alpha = 1 - torch.exp(-raw_c[..., 3].abs() * dists)
Tc = torch.cumprod(1 - alpha + 1e-6, -1); lam = Tc[..., -1]
Tc = torch.cat((torch.ones_like(Tc[..., :1]), Tc[..., :-1]), -1)
wts = alpha * Tc
rgb = (wts.unsqueeze(-1) * torch.sigmoid(raw_c[..., :3])).sum(-2)
((rgb * wr).sum() + (lam * wl).sum()).backward()
E = lambda *s: torch.empty(*s, device=DEV); Z = lambda *s: torch.zeros(*s, device=DEV)
raw_g, fz_g, zm_g, d_g2 = (x.contiguous() for x in (raw.to(DEV), T(g["fg"]).to(DEV), T(g["far"]).to(DEV), T(g["d"]).to(DEV)))
w_o, rgb_o, dep_o, lam_o = E(N, S), E(N, 3), E(N), E(N)
st = _lib.stream()
_lib.check(lib.scnerf_pp_composite_fg_fwd(_lib.ptr(raw_g), _lib.ptr(fz_g), _lib.ptr(zm_g), _lib.ptr(d_g2), N, S, _lib.ptr(w_o), _lib.ptr(rgb_o), _lib.ptr(dep_o), _lib.ptr(lam_o), st))
d_raw, d_fz, d_zm, g_d = E(N, S, 4), E(N, S), Z(N), Z(N, 3)
_lib.check(lib.scnerf_pp_composite_fg_bwd(_lib.ptr(raw_g), _lib.ptr(fz_g), _lib.ptr(zm_g), _lib.ptr(d_g2), N, S, _lib.ptr(wr.to(DEV)), _lib.ptr(wl.to(DEV)), _lib.ptr(d_raw), _lib.ptr(d_fz), _lib.ptr(d_zm), _lib.ptr(g_d), st))
print("fg comp fwd rgb", rel(rgb_o, rgb), "lam", rel(lam_o, lam), "| d_raw", rel(d_raw, raw_c.grad), "d_fz", rel(d_fz, fz_c.grad), "d_zmax", rel(d_zm, zmax_c.grad), "g_d", rel(g_d, d_c.grad))
# ---- bg compositing
bz = T(g["bg"]); raw_c = raw.clone().requires_grad_(True); lam_in = T(rng.uniform(0.1, 1, N).astype(np.float32)).requires_grad_(True)
fgr = T(rng.uniform(0, 1, (N, 3)).astype(np.float32))
zf = torch.flip(bz, [-1]); dists = torch.cat((zf[..., :-1] - zf[..., 1:], 1e10 * torch.ones_like(zf[..., :1])), -1)
alpha = 1 - torch.exp(-raw_c[..., 3].abs() * dists)
Tc = torch.cumprod(1 - alpha + 1e-6, -1)[..., :-1]; Tc = torch.cat((torch.ones_like(Tc[..., :1]), Tc), -1)
wts = alpha * Tc
rgbt = fgr + lam_in.unsqueeze(-1) * (wts.unsqueeze(-1) * torch.sigmoid(raw_c[..., :3])).sum(-2)
(rgbt * wr).sum().backward()
w_o, brgb, bdep, rgb_o = E(N, S), E(N, 3), E(N), E(N, 3)
_lib.check(lib.scnerf_pp_composite_bg_fwd(_lib.ptr(raw_g), _lib.ptr(bz.to(DEV)), _lib.ptr(lam_in.detach().to(DEV)), _lib.ptr(fgr.to(DEV)), N, S, _lib.ptr(w_o), _lib.ptr(brgb), _lib.ptr(bdep), _lib.ptr(rgb_o), st))
d_raw, d_lam = E(N, S, 4), E(N)
_lib.check(lib.scnerf_pp_composite_bg_bwd(_lib.ptr(raw_g), _lib.ptr(bz.to(DEV)), _lib.ptr(lam_in.detach().to(DEV)), N, S, _lib.ptr(wr.to(DEV)), _lib.ptr(d_raw), _lib.ptr(d_lam), st))
print("bg comp fwd rgb", rel(rgb_o, rgbt), "| d_raw", rel(d_raw, raw_c.grad), "d_lam", rel(d_lam, lam_in.grad))
