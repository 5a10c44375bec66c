"""Evidence for the gradient gate of tests/test_gpu_parity_pp.py: where the fp32 oracle loses a gradient tensor against the
fp64 oracle, the error matrix is rank-1 (one sample whose ReLU pre-activation changes sign).  CPU only.
  python tools/relu_sign_events.py > profiles/r2c_relu_sign_events.txt"""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
torch.set_num_threads(8)
import importlib
import types
# reuse test helpers without the gpu marker side effects
src=open('/root/repo/tests/test_gpu_parity_pp.py').read()
ns={}
exec(compile(src,'pp','exec'),ns)
N,cascade,seed=256,[64,128],70
l32,r32,g32=ns['_pp_oracle_step'](seed,N,cascade,torch.float32)
l64,r64,g64=ns['_pp_oracle_step'](seed,N,cascade,torch.float64)
for k in ('net1.bg_net.base_layers.3.0.weight','net0.bg_net.base_layers.3.0.weight','net1.bg_net.base_layers.1.0.weight','net0.fg_net.base_layers.3.0.weight'):
    a,b=g32[k],g64[k]; d=np.abs(a-b); m=np.abs(b).max()
    rows=d.max(1)/m; cols=d.max(0)/m
    print(k, 'max err %.2e'%(d.max()/m), 'rows>1e-4:', (rows>1e-4).sum(), 'top rows', np.argsort(-rows)[:4], np.sort(rows)[-4:], 'cols>1e-4:', (cols>1e-4).sum(), 'median row err %.1e'%np.median(rows))
for k in ('net1.bg_net.base_layers.3.0.weight','net1.bg_net.base_layers.1.0.weight','net0.bg_net.base_layers.3.0.weight','net0.fg_net.base_layers.3.0.weight'):
    d=(g32[k].astype(np.float64)-g64[k]); s=np.linalg.svd(d,compute_uv=False)
    print(k,'singular values of the fp32-fp64 error: top5/total_fro', s[:5]/np.linalg.norm(s), ' rel fro err %.2e'%(np.linalg.norm(d)/np.linalg.norm(g64[k])))
