"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares; host
logic (flat gradient buffer, all-reduce over gloo with world_size 2, ray sharding) works; the
product path refuses CPU tensors instead of falling back."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from scnerf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def test_header_symbols_exported():
    _lib = _ensure_built()
    lib = _lib.load()
    header = "".join(open(os.path.join(ROOT, "include", h)).read()
                     for h in ("scnerf_b200.h", "scnerf_b200_nerfpp.h"))
    declared = set(re.findall(r"\b(scnerf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.scnerf_abi_version() == 1
    assert lib.scnerf_built_for_sm() == 100


def test_struct_sizes_match_header(tmp_path):
    """ctypes mirrors must have the C layout: sizes AND the offset of every struct's last member, as the C compiler
    lays the header's structs out (gcc on include/*.h — the header is plain C)."""
    import ctypes as C
    import subprocess
    _lib = _ensure_built()
    pairs = [("scnerf_camera", _lib.Camera), ("scnerf_camera_grads", _lib.CameraGrads), ("scnerf_mlp", _lib.Mlp),
             ("scnerf_raygen_args", _lib.RaygenArgs), ("scnerf_rayprep_args", _lib.RayprepArgs),
             ("scnerf_render_cfg", _lib.RenderCfg), ("scnerf_render_rand", _lib.RenderRand),
             ("scnerf_render_out", _lib.RenderOut), ("scnerf_render_grads_in", _lib.RenderGradsIn),
             ("scnerf_step_io", _lib.StepIO), ("scnerf_adam_tensor", _lib.AdamTensor),
             ("scnerf_pp_raygen_args", _lib.PPRaygenArgs)] + list(getattr(_lib, "EXTRA_STRUCTS", []))
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "scnerf_b200.h"', '#include "scnerf_b200_nerfpp.h"',
           'int main(void) {']
    for cname, cls in pairs:
        last = cls._fields_[-1][0]
        src.append(f'  printf("{cname} %zu %zu\\n", sizeof({cname}), offsetof({cname}, {last}));')
    src += ['  return 0;', '}']
    c = tmp_path / "sizes.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out if l.strip()}
    for cname, cls in pairs:
        last = cls._fields_[-1][0]
        assert got[cname] == (C.sizeof(cls), getattr(cls, last).offset), (cname, got[cname], C.sizeof(cls))


def test_no_cpu_fallback():
    _lib = _ensure_built()
    with pytest.raises(RuntimeError, match="CUDA"):
        _lib.ptr(torch.zeros(3))
    from scnerf_b200.render import searchsorted
    with pytest.raises(RuntimeError):
        searchsorted(torch.zeros(1, 4), torch.zeros(1, 2))


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from scnerf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_modules_have_reference_state_dict_keys():
    """checkpoint compatibility."""
    from tests.util import build_modules
    from scnerf_b200.run_nerf_helpers import SingleDeviceParallel
    m = build_modules(0, "cpu")
    keys = list(m["coarse"].state_dict())
    assert keys[:4] == ["pts_linears.0.weight", "pts_linears.0.bias", "pts_linears.1.weight", "pts_linears.1.bias"]
    assert m["coarse"].state_dict()["pts_linears.5.weight"].shape == (256, 319)
    assert m["coarse"].state_dict()["views_linears.0.weight"].shape == (128, 283)
    assert sum(p.numel() for p in m["coarse"].parameters()) == 595844
    assert list(SingleDeviceParallel(m["coarse"]).state_dict())[0] == "module.pts_linears.0.weight"
    cam = m["cam"]
    assert [n for n, _ in cam.named_parameters()] == [
        "intrinsics_initial", "extrinsics_initial", "intrinsics_noise", "extrinsics_noise",
        "ray_o_noise", "ray_d_noise"]
    assert cam.ray_o_noise.shape == (37, 50, 3) and not cam.intrinsics_initial.requires_grad
    # API-parity getters agree with the oracle
    from oracle import scnerf_oracle as O
    from scnerf_b200 import synth
    oc = O.Camera(synth.intrinsic_init(), synth.camera_poses(0), synth.camera_args(), 378, 504)
    oc.load(synth.camera_noise_state(0))
    np.testing.assert_allclose(cam.get_intrinsic().detach().numpy(), oc.intrinsic().numpy(), rtol=1e-6)
    np.testing.assert_allclose(cam.get_extrinsic().detach().numpy(), oc.extrinsic().numpy(), rtol=1e-5, atol=1e-7)
    K5, E5 = cam(5)
    np.testing.assert_allclose(E5.detach().numpy(), oc.extrinsic()[5].numpy(), rtol=1e-5, atol=1e-7)


def test_shard_rays_partition():
    from scnerf_b200.parallel import shard_rays
    for n, ws in ((190512, 8), (4096, 3), (7, 4)):
        spans = [shard_rays(n, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scnerf_b200.parallel import FlatGrads
    named = [("a", torch.zeros(3, 4)), ("b", torch.zeros(5))]
    fg = FlatGrads(named, "cpu")
    fg.views["a"].fill_(float(rank + 1))
    fg.views["b"].copy_(torch.arange(5.) * (rank + 1))
    fg.all_reduce_mean()
    q.put((rank, fg.flat.clone().numpy()))
    dist.destroy_process_group()


def test_flat_grad_allreduce_gloo_world2():
    """The N>1 path on CPU: world_size-2 gloo, one collective over the flat buffer, mean."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    want = np.concatenate([np.full(12, 1.5), np.arange(5.) * 1.5]).astype(np.float32)
    for r in range(2):
        np.testing.assert_allclose(res[r], want)


def test_synth_is_deterministic():
    from scnerf_b200 import synth
    a, b = synth.mlp_state(3), synth.mlp_state(3)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    k1, i1, t1 = synth.pixel_batch(9, 64)
    k2, i2, t2 = synth.pixel_batch(9, 64)
    assert np.array_equal(k1, k2) and np.array_equal(i1, i2) and np.array_equal(t1, t2)
    assert k1[:, 0].max() < synth.FERN_W and k1[:, 1].max() < synth.FERN_H


def test_checkpoint_layout(tmp_path):
    """checkpoints use the reference's dict keys and state-dict keys (run_nerf.py:626-641)."""
    import types
    from scnerf_b200.ray_batch import save_checkpoint
    from scnerf_b200.run_nerf_helpers import NeRF, SingleDeviceParallel
    mk = lambda: SingleDeviceParallel(NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27,
                                           use_viewdirs=True))
    kw = {"network_fn": mk(), "network_fine": mk()}
    opt = torch.optim.Adam(list(kw["network_fn"].parameters()) + list(kw["network_fine"].parameters()), lr=5e-4)
    path = save_checkpoint(str(tmp_path / "000100.tar"), 100, kw, opt)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"global_step", "network_fn_state_dict", "network_fine_state_dict", "optimizer_state_dict"}
    assert "module.pts_linears.0.weight" in ck["network_fn_state_dict"]          # nn.DataParallel prefix kept
    fresh = mk()
    fresh.load_state_dict(ck["network_fine_state_dict"])


# ---- slab plans of the N-half pipelined tensor-core kernels (host-side tables, no GPU needed) ----------------
F_ZERO_ACC, F_STAGE_END, F_HI_ONLY_A, F_STAGE_BEGIN, F_WAIT_Q1, F_WAIT_Q2, F_WAIT_Q3, F_COMMIT_BOTH = 1, 2, 4, 8, 16, 32, 64, 128
F_COMMIT_SIDE, F_WAIT_SIDE = 256, 512
A_TMEM, A_SMEM, A_MIX = 0, 1, 2


def _plan(which):
    import ctypes as C
    _lib = _ensure_built()
    lib = _lib.load()
    out = (C.c_int64 * 9)()
    _lib.check(lib.scnerf_debug_slab_plan(which, -1, out))
    n_slabs, n_stages, img_bytes = out[0], out[1], out[2]
    slabs = []
    for i in range(n_slabs):
        _lib.check(lib.scnerf_debug_slab_plan(which, i, out))
        slabs.append(dict(zip(("n", "acc_col", "a_off", "a_lo", "a_kind", "flags", "stage", "pad", "img_off"), list(out))))
    return n_slabs, n_stages, img_bytes, slabs


@pytest.mark.parametrize("which,name,ring_multiples", [(1, "pipelined forward (3-D points)", (8, 24)),
                                                       (2, "pipelined forward (4-D points)", (6, 24)),
                                                       (4, "pipelined dgrad (3-D points)", (8, 24)),
                                                       (3, "pipelined dgrad (4-D points)", (8, 24))])
def test_pipelined_slab_plans(which, name, ring_multiples):
    """Hazard invariants of the N-half pipelined schedules (csrc/field_tc_fwd_pipe.cuh, field_tc_dgrad_pipe.cuh):
    every K-slab that reads operand quarter q is issued after this stage's wait on quarter barrier q; an accumulator
    half is only re-initialised after the barrier that says it was drained; each barrier is waited exactly once per
    stage (the kernels derive every mbarrier parity from `stage & 1`); the weight image is one contiguous run in
    issue order; ring geometry divides the slab count."""
    n_slabs, n_stages, img_bytes, slabs = _plan(which)
    assert n_stages == 10 and n_stages % 2 == 0
    for m in ring_multiples:                       # GROUP x NSLOT of the split-bf16 / bf16 builds
        assert n_slabs % m == 0, (name, n_slabs, m)
    off = 0
    for s in slabs:                                # packed in issue order, no holes
        assert s["img_off"] == off
        off += s["n"] * 32
    assert img_bytes == 2 * off
    for st in range(n_stages):
        mine = [s for s in slabs if s["stage"] == st]
        waited = set()
        zeroed = set()
        commits = []
        side = "none"                                  # side pass borrowing acc1 (4-D dgrad, T4): none -> open -> committed -> parked
        for s in mine:
            fl = s["flags"]
            if fl & F_WAIT_SIDE:
                assert side == "committed" and (fl & F_ZERO_ACC) and s["acc_col"] == 128, (name, st, "side wait misplaced", s)
                side = "parked"
            elif s["acc_col"] == 128 and side in ("open", "committed"):
                assert side == "open" and not (fl & F_ZERO_ACC), (name, st, "acc1 touched while it holds the side result", s)
            if (fl & F_ZERO_ACC) and s["acc_col"] == 128 and s["n"] == 96:
                assert side == "none" and 0 in zeroed, (name, st, "side pass must follow the h0 pass")
                side = "open"
            if fl & F_COMMIT_SIDE:
                assert side == "open" and not (fl & F_STAGE_END), (name, st, s)
                side = "committed"
            for q, bit in enumerate((F_STAGE_BEGIN, F_WAIT_Q1, F_WAIT_Q2, F_WAIT_Q3)):
                if fl & bit:
                    assert q not in waited, (name, st, "barrier waited twice", q)
                    waited.add(q)
            acc = s["acc_col"]
            assert acc in (0, 128, 448) and (acc + s["n"] <= 256 or (acc == 448 and s["n"] <= 64)), (name, st, s)
            if fl & F_ZERO_ACC:
                # acc0 (columns 0..127) is drained by the owners of quarters 0 and 1: quarter 0's arrival follows
                # the loads of BOTH chunks of a thread, so waiting quarter 0 is enough; acc1 needs quarter 2
                if acc == 0:
                    assert 0 in waited, (name, st, "acc0 re-initialised before it was drained")
                if acc == 128:
                    assert 2 in waited, (name, st, "acc1 re-initialised before it was drained")
                zeroed.add(acc)
            else:
                assert acc in zeroed or s["n"] == 16, (name, st, "accumulate into an uninitialised accumulator", s)
            if s["a_kind"] == A_MIX:               # hidden-state operand: hi in TMEM, lo in smem
                col = s["a_off"]
                assert 0 <= col < 192 and col % 8 == 0
                if col < 128:                      # first operand half: buffer P[stage & 1], quarters 0 / 1
                    assert col // 64 == st % 2, (name, st, "wrong half of the P double buffer", col)
                    q = (col % 64) // 32
                    assert s["a_lo"] * 16 == (st % 2) * 32768 + (col % 64) // 8 * 4096
                else:                              # second half: Q, quarters 2 / 3
                    q = 2 + (col - 128) // 32
                    assert s["a_lo"] * 16 == 65536 + (col - 128) // 8 * 4096
                assert q in waited, (name, st, f"slab reads operand quarter {q} before its barrier", s)
            if fl & F_STAGE_END:
                commits.append(s["pad"])
                if fl & F_COMMIT_BOTH:
                    commits.append(s["pad"] ^ 1)
        assert side in ("none", "parked"), (name, st, side)        # a side pass is always drained inside its stage
        assert (side == "parked") == (which == 3 and st == 4)
        assert waited == {0, 1, 2, 3}, (name, st, waited)          # every barrier advances once per stage
        assert sorted(commits) == [0, 1], (name, st, commits)      # both accumulator-half barriers fire once
        assert mine[-1]["flags"] & F_STAGE_END                     # nothing is issued after the stage's last commit


def test_default_precision_follows_the_network_shape():
    """The default precision is the tensor-core path only for the network shape those kernels are specialised for; other
    shapes (the reference's own default has no view directions) run on the fp32 CUDA-core kernels, announced once; an
    explicit precision is never overridden."""
    import types
    import warnings
    from scnerf_b200 import _lib
    from scnerf_b200.nerfplusplus.ddp_model import NerfNet
    from scnerf_b200.run_nerf_helpers import NeRF
    std = NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
    plain = NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=0, use_viewdirs=False)
    narrow = NeRF(D=8, W=128, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
    _lib._warned_shapes.clear()
    assert _lib.resolve_precision(None, std, std) == _lib.default_precision()
    assert _lib.resolve_precision("bf16x3", plain) == "bf16x3"          # explicit: the library will refuse it loudly
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert _lib.resolve_precision(None, std, plain) == "fp32"
        assert _lib.resolve_precision(None, narrow, None) == "fp32"
        assert _lib.resolve_precision(None, plain) == "fp32"            # second time: no second warning
    assert len(w) == 2 and "fp32 CUDA-core" in str(w[0].message)
    args = types.SimpleNamespace(max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8, netwidth=256, use_viewdirs=True)
    assert NerfNet(args).precision == _lib.default_precision()
