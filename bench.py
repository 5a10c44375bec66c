#!/usr/bin/env python
"""bench.py — SCNeRF hot-path training step on B200 (BASELINE.json metric: rays/sec at
4096 rays x (64c+128f) samples, 8x256 MLP; % of tensor-core roofline).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp32|bf16x3|bf16]

A "step" = one pass of the hot path over one synthetic batch: pixel indices -> learnable-camera
rays -> NDC -> stratified + importance sampling -> coarse+fine PE/MLP -> composite -> loss ->
gradients of both MLPs and all camera parameters (no optimiser), NeRF/run_nerf.py:385-506,600.
Weak scaling: every rank renders its own 4096 rays; one all-reduce of the flat gradient buffer.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_RAYS, NC, NF = 4096, 64, 128
FLOP_PER_SAMPLE = 1_186_816            # 593,408 MAC per sample evaluation
EVALS_PER_RAY = NC + (NC + NF)         # coarse net on 64, fine net on 192
FWD_FLOP_PER_RAY = EVALS_PER_RAY * FLOP_PER_SAMPLE
TRAIN_FLOP_PER_RAY = 3 * FWD_FLOP_PER_RAY   # fwd + dgrad + wgrad (activations kept, no recompute)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), d["hbm_gbs"], "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_cpu_rays_per_s(n_rays, steps, warmup, threads):
    """The reference algorithm (CPU PyTorch restatement, oracle/) on the host cores: same scene,
    same step (fwd+bwd, perturb=1, raw_noise_std=1), bounded sample of the 4096-ray batch."""
    from oracle import scnerf_oracle as O
    from scnerf_b200 import synth
    torch.set_num_threads(threads)
    H, W = synth.FERN_H, synth.FERN_W
    cam = O.Camera(synth.intrinsic_init(), synth.camera_poses(0), synth.camera_args(), H, W)
    cam.load(synth.camera_noise_state(0), True)
    Pc = O.state_to_tensors(synth.mlp_state(0), requires_grad=True)
    Pf = O.state_to_tensors(synth.mlp_state(1), requires_grad=True)
    kps, idx, target = synth.pixel_batch(0, n_rays)
    kps, idx, target = torch.from_numpy(kps), torch.from_numpy(idx), torch.from_numpy(target)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _, _ = O.train_step(cam, Pc, Pf, kps, idx, target, H, W, NC, NF,
                                  t_rand=torch.rand(n_rays, NC), u=torch.rand(n_rays, NF),
                                  noise0=torch.randn(n_rays, NC), noise1=torch.randn(n_rays, NC + NF))
        loss.backward()
        for t in list(Pc.values()) + list(Pf.values()) + cam.learnables():
            t.grad = None
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return n_rays / (sum(times) / len(times)), sum(times) / len(times)


def best_thread_count(n_probe=32):
    """Eager CPU PyTorch oversubscribes on many-core hosts (128 threads ran 10x slower than 16 on
    the B200 box): probe a few thread counts on a tiny batch and keep the fastest."""
    cores = os.cpu_count() or 1
    best, best_t = 1, float("inf")
    for t in sorted({min(cores, c) for c in (8, 16, 32, 64, cores)}):
        _, sec = oracle_cpu_rays_per_s(n_probe, 1, 1, t)
        if sec < best_t:
            best, best_t = t, sec
    return best


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = best_thread_count()
    n = 256
    rps, sec = oracle_cpu_rays_per_s(n, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "rays/sec", "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: LLFF fern full, 4096 rays x (64c+128f), learnable intrinsics+extrinsics",
                   "note": f"each step = {n}-ray sample of the 4096-ray batch on the host CPU"},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{n} of 4096 rays, fwd+bwd, CPU PyTorch restatement of the reference (oracle/); "
                                   f"{threads} of {os.cpu_count()} host threads (fastest of a probe)"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SCNERF_PRECISION", "bf16x3"),
                    help="bf16x3 (default: split-bf16 tensor-core path, the parity-grade mode) | bf16 | fp32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch.distributed as dist
    from scnerf_b200 import _lib, synth
    from scnerf_b200.engine import TrainStep
    from tests.util import build_modules
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lib = _lib.load()
    mods = build_modules(0, dev)                          # identical replicas on every rank
    kps, idx, target = synth.pixel_batch(1000 + rank, N_RAYS)   # each rank draws its own rays
    eng = TrainStep(mods["cam"], mods["coarse"], mods["fine"], N_RAYS, NC, NF, perturb=1.0,
                    raw_noise_std=1.0, precision=args.precision, seed=rank)
    eng.kps_dev.copy_(torch.from_numpy(kps)); eng.idx_dev.copy_(torch.from_numpy(idx))
    eng.target_dev.copy_(torch.from_numpy(target))
    eng.kps_host.copy_(torch.from_numpy(kps)); eng.idx_host.copy_(torch.from_numpy(idx))
    eng.target_host.copy_(torch.from_numpy(target))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
            eng.grads.all_reduce_mean()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(max(args.warmup, 3)):
        eng.step_device(); eng.grads.all_reduce_mean()
    lib.scnerf_launch_count(1)
    with ClockSampler(local) as clk:
        ms_dev = timed(eng.step_device, args.steps)
    launches = int(lib.scnerf_launch_count(1))
    for _ in range(2):
        eng.step_host()
    ms_e2e = timed(eng.step_host, args.steps)
    loss = float(eng.loss_host)

    # roofline of the dominant stage: the fine-network field evaluation (PE + 8x256 MLP) —
    # timed alone with CUDA events on the launching stream
    from scnerf_b200.create_nerf import run_network
    P_rays = N_RAYS
    pts = torch.rand(P_rays, NC + NF, 3, device=dev) * 2 - 1
    vd = torch.nn.functional.normalize(torch.randn(P_rays, 3, device=dev), dim=-1)
    for _ in range(2):
        run_network(pts, vd, mods["fine"], None, None, precision=args.precision)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run_network(pts, vd, mods["fine"], None, None, precision=args.precision)
    e1.record()
    torch.cuda.synchronize()
    ms_field = e0.elapsed_time(e1) / reps
    burst, sustained, hbm, src = measured_peaks()
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/traffic.json,
    # written by tools/ncu_traffic.py); null when the capture for this precision is missing
    traffic = None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        pipe = os.environ.get("SCNERF_FWD_PIPE", "1") != "0"      # which forward kernel this process launches
        traffic = tj.get(f"{'field_fwd_pipe_kernel' if pipe else 'field_fused_fwd_kernel'}/{args.precision}/inference", {}).get("dram_bytes")
    field_flop = P_rays * (NC + NF) * FLOP_PER_SAMPLE
    achieved = field_flop / (ms_field * 1e-3) / 1e12

    # single-pass bf16 throughput mode, reported next to the headline (not parity-grade: see DESIGN.md §3)
    alt = None
    if args.precision == "bf16x3":
        eng2 = TrainStep(mods["cam"], mods["coarse"], mods["fine"], N_RAYS, NC, NF, perturb=1.0,
                         raw_noise_std=1.0, precision="bf16", seed=rank)
        eng2.kps_dev.copy_(eng.kps_dev); eng2.idx_dev.copy_(eng.idx_dev); eng2.target_dev.copy_(eng.target_dev)
        eng_saved, eng = eng, eng2
        for _ in range(3):
            eng.step_device(); eng.grads.all_reduce_mean()
        ms_alt = timed(eng.step_device, args.steps)
        eng = eng_saved
        alt = {"dtype": "bf16", "ms_per_step": ms_alt, "value": N_RAYS * world / (ms_alt * 1e-3), "unit": "rays/s",
               "note": "single-pass bf16 tensor-core path, fp32 accumulate; ~1e-2 relative error on raw, not parity-grade"}
        del eng2
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    rays_total = N_RAYS * world
    line = {
        "metric": "rays/sec", "value": rays_total / (ms_dev * 1e-3), "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "bf16x3": "bf16x3", "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "configs[1]: LLFF fern full, 4096 rays x (64c+128f), learnable intrinsics+extrinsics"
                               " (+ray_o/ray_d residual grids), fwd+bwd, per GPU",
                   "rays_per_gpu": N_RAYS, "N_samples": NC, "N_importance": NF, "mlp": "8x256 coarse + 8x256 fine",
                   "parallelism": f"dp{world}", "precision": args.precision, "perturb": 1, "raw_noise_std": 1.0,
                   "l2": "no flush: per-step working set (bf16 tile images of every layer input and dZ, ~20 GB) >> 126 MB L2",
                   "train_flop_per_ray": TRAIN_FLOP_PER_RAY},
        "e2e": {"value": rays_total / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": eng.h2d_bytes, "d2h_bytes_per_step": eng.d2h_bytes,
                "api": "scnerf_train_step(inputs_on_host=1): pinned host pixel/target buffers in, loss out"},
        "gpu_launches": launches,
        "loss": loss,
        "clocks": clk.summary(),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                     "frac": achieved / burst, "traffic": traffic, "peak_source": f"{src} cuBLAS bf16 burst",
                     "kernel": f"fine-network field forward (PE + 8x256 MLP, {args.precision}) on {P_rays * (NC + NF)} samples",
                     "ms": ms_field, "algorithmic_flop": field_flop,
                     "step_frac_of_sustained": (TRAIN_FLOP_PER_RAY * N_RAYS / (ms_dev * 1e-3) / 1e12) / sustained},
    }
    if alt is not None:
        line["throughput_mode"] = alt
    if world == 1 and not args.no_cpu_baseline:
        threads = best_thread_count()
        n = 256
        rps, sec = oracle_cpu_rays_per_s(n, 2, 1, threads)
        line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": f"{n} of 4096 rays, fwd+bwd, {sec:.2f} s/step, CPU PyTorch restatement (oracle/); "
                                          f"{threads} of {os.cpu_count()} host threads (fastest of a probe)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
