"""Golden vectors for the optimiser row by RUNNING THE REFERENCE's CustomAdamOptimizer
(NeRF/create_nerf.py:259-336).  Build container only:  python tests/golden/make_golden_adam.py"""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path[:0] = [ROOT, REF + "/NeRF", REF, REF + "/model"]
sys.modules.setdefault("imageio", mock.MagicMock())
from scnerf_b200 import synth  # noqa: E402
import run_nerf_helpers  # noqa: E402,F401
torch.autograd.set_detect_anomaly(False)
import create_nerf as ref_create  # noqa: E402

out = {}
for tag, cam_name, amsgrad, wd in (("dist", "pinhole_rot_noise_10k_rayo_rayd_dist", False, 0.1),
                                   ("od", "pinhole_rot_noise_10k_rayo_rayd", True, 0.05),
                                   ("none", "none", False, 0.1)):
    p0, grads = synth.adam_case(1)
    params = [torch.nn.Parameter(torch.from_numpy(p.copy())) for p in p0]
    args = types.SimpleNamespace(camera_model=cam_name)
    opt = ref_create.CustomAdamOptimizer(params=params, lr=5e-4, betas=(0.9, 0.999), weight_decay=wd, H=378, W=504,
                                         args=args, amsgrad=amsgrad)
    for step, gs in enumerate(grads):
        for p, g in zip(params, gs):
            p.grad = torch.from_numpy(g.copy())
        for group in opt.param_groups:                      # the trainer's schedule, run_nerf.py:617-621
            group["lr"] = 5e-4 * (0.1 ** (step / 250000))
        opt.step()
        for i, p in enumerate(params):
            out[f"{tag}_s{step}_p{i}"] = p.detach().numpy().copy()
np.savez_compressed(os.path.join(HERE, "adam.npz"), **out)
print("adam:", os.path.getsize(os.path.join(HERE, "adam.npz")) // 1024, "KiB")
