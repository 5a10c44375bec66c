"""Worker of tests/test_gpu_multi.py (one process per GPU under torch.distributed.run, NCCL): the multi-rank paths that
a single-GPU test cannot reach — render_single_image's pixel sharding + all_gather (nerfplusplus/ddp_train_nerf.py:135-256),
the flat-gradient all-reduce of both engines, strong-scaling additivity of the fused steps."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from scnerf_b200 import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    out = {}
    # ---- f3: full-image NeRF++ inference, pixels sharded over ranks, one all_gather per key --------------------------
    from scnerf_b200.nerfplusplus.ddp_train_nerf import render_single_image
    mods = synth.build_pp_modules(36, dev, levels=2, precision="bf16x3")
    for p in mods["cam"].parameters():
        p.requires_grad_(False)
    models = {"cascade_level": 2, "cascade_samples": [16, 16], "net_0": mods["nets"][0], "net_1": mods["nets"][1]}
    sampler = types.SimpleNamespace(H=synth.PP_H, W=synth.PP_W, c2w_mat=None)
    multi = render_single_image(rank, world, models, sampler, 4096, mods["cam"], camera_idx=3)
    single = render_single_image(0, 1, models, sampler, 4096, mods["cam"], camera_idx=3)     # every rank: whole image alone
    if rank == 0:
        assert multi is not None and len(multi) == 2
        for m in range(2):
            for k in single[m]:
                assert multi[m][k].shape == single[m][k].shape, (k, multi[m][k].shape)
                d = float((multi[m][k] - single[m][k]).abs().max())
                out[f"render_single_image level{m} {k} max|multi-single|"] = d
                assert d == 0.0, (m, k, d)        # same kernels on the same pixels: sharding must not change a bit
    else:
        assert multi is None                      # only rank 0 returns (ddp_train_nerf.py:252-256)
    # ---- e: data-parallel step = the mean over ranks of the per-rank gradients (NeRF/ and NeRF++ engines) -----------
    from scnerf_b200.engine import TrainStep
    N = 512
    nm = synth.build_modules(40, dev)
    T = torch.from_numpy
    kps, idx, target = synth.pixel_batch(40, N * world)
    sl = slice(rank * N, (rank + 1) * N)
    eng = TrainStep(nm["cam"], nm["coarse"], nm["fine"], N, 64, 128, perturb=0., raw_noise_std=0., precision="bf16x3")
    eng.step_device(T(kps[sl]).to(dev), T(idx[sl]).to(dev), T(target[sl]).to(dev))
    eng.grads.all_reduce_mean()
    full = TrainStep(nm["cam"], nm["coarse"], nm["fine"], N * world, 64, 128, perturb=0., raw_noise_std=0., precision="bf16x3")
    full.step_device(T(kps).to(dev), T(idx).to(dev), T(target).to(dev))       # the same rays as ONE batch on one GPU
    torch.cuda.synchronize()
    a, b = eng.grads.flat, full.grads.flat
    e = float((a - b).abs().max() / b.abs().max())
    out["nerf strong-scaling additivity (all-reduced shards vs one batch), rel-to-max"] = e
    assert e <= 2e-3, e
    from scnerf_b200.nerfplusplus.engine import PPTrainStep
    pm = synth.build_pp_modules(41, dev, levels=2, precision="bf16x3")
    sel, cam_idx, tgt = synth.pp_pixel_batch(41, 256 * world)
    rng = np.random.default_rng(5)
    R = {k: T(rng.random((256 * world, n), dtype=np.float32)).to(dev) for k, n in (("t_fg", 32), ("t_bg", 32), ("u_fg", 64), ("u_bg", 64))}
    s2 = slice(rank * 256, (rank + 1) * 256)
    pe = PPTrainStep(pm["cam"], pm["nets"], 256, [32, 64], camera_idx=cam_idx)
    pe.step_device(T(sel[s2]).to(dev), T(tgt[s2]).to(dev), rand={k: v[s2].contiguous() for k, v in R.items()})
    pe.grads.all_reduce_mean()
    pf = PPTrainStep(pm["cam"], pm["nets"], 256 * world, [32, 64], camera_idx=cam_idx)
    pf.step_device(T(sel).to(dev), T(tgt).to(dev), rand=R)
    torch.cuda.synchronize()
    e = float((pe.grads.flat - pf.grads.flat).abs().max() / pf.grads.flat.abs().max())
    out["nerfpp strong-scaling additivity, rel-to-max"] = e
    assert e <= 2e-3, e
    if rank == 0:
        print("MULTI_GPU_OK " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
