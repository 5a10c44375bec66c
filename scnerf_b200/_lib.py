"""ctypes binding of ``csrc/libscnerf_b200.so`` (C ABI: ``include/scnerf_b200.h``, ``include/scnerf_b200_nerfpp.h``).

There is NO fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is
raised.  PyTorch is used only for device memory (``tensor.data_ptr()``) and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCNERF_LIB") or os.path.join(_HERE, "csrc", "libscnerf_b200.so")   # SCNERF_LIB: e.g. a -DSCNERF_TIMELINE build
MAX_DEPTH = 16
PRECISION = {"fp32": 0, "bf16x3": 1, "bf16": 2}


def default_precision():
    """Field arithmetic when the caller does not say: ``SCNERF_PRECISION`` if set, else ``bf16x3`` — the split-bf16
    tcgen05 path (three MMAs per product, fp32 accumulate), which meets the 1e-4 forward gate and the fp32-noise-floor
    gradient gate and is the path the library is built for (sm_100a only).  ``fp32`` selects the
    CUDA-core kernels (exact fp32 arithmetic, ~8x slower), ``bf16`` the single-pass throughput mode (~1e-2 error)."""
    p = os.environ.get("SCNERF_PRECISION", "bf16x3")
    if p not in PRECISION:
        raise ValueError(f"SCNERF_PRECISION={p!r}: expected one of {sorted(PRECISION)}")
    return p


_warned_shapes = set()


def resolve_precision(precision, *nets):
    """The precision a call runs at.  An explicit ``precision`` is taken as is (an unsupported network shape then fails
    loudly in the library).  The DEFAULT adapts to the network: the tcgen05 kernels are specialised for the reference's
    standard 8 x 256 / skip 4 / view-direction network, every other shape (``--netdepth``, ``--netwidth``, no
    ``--use_viewdirs``, other ``--multires``) runs on the fp32 CUDA-core kernels, announced once per shape."""
    if precision is not None:
        return precision
    p = default_precision()
    if p != "fp32":
        for net in nets:
            if net is not None and not net.tensor_core_shape():
                icv = getattr(net, "input_ch_views", getattr(net, "input_ch_viewdirs", None))
                key = (net.D, net.W, tuple(net.skips), bool(net.use_viewdirs), net.input_ch, icv)
                if key not in _warned_shapes:
                    _warned_shapes.add(key)
                    import warnings
                    warnings.warn(f"scnerf_b200: network shape D={net.D} W={net.W} skips={list(net.skips)} "
                                  f"use_viewdirs={bool(net.use_viewdirs)} input_ch={net.input_ch}/{icv} is not "
                                  f"the one the tensor-core kernels are specialised for: running it on the fp32 CUDA-core "
                                  f"kernels (exact fp32, slower); pass precision= explicitly to override", stacklevel=3)
                return "fp32"
    return p

_f = C.POINTER(C.c_float)
vp = C.c_void_p


class Camera(C.Structure):
    _fields_ = [("intrinsics_initial", vp), ("intrinsics_noise", vp), ("extrinsics_initial", vp),
                ("extrinsics_noise", vp), ("ray_o_noise", vp), ("ray_d_noise", vp),
                ("intrinsics_noise_scale", C.c_float), ("extrinsics_noise_scale", C.c_float),
                ("ray_o_noise_scale", C.c_float), ("ray_d_noise_scale", C.c_float),
                ("multiplicative_noise", C.c_int32), ("n_cams", C.c_int32), ("H", C.c_int32),
                ("W", C.c_int32), ("gh", C.c_int32), ("gw", C.c_int32)]


class CameraGrads(C.Structure):
    _fields_ = [("intrinsics_noise", vp), ("extrinsics_noise", vp), ("ray_o_noise", vp),
                ("ray_d_noise", vp)]


class RaygenArgs(C.Structure):
    _fields_ = [("cam", C.POINTER(Camera)), ("focal", C.c_float), ("H", C.c_int32), ("W", C.c_int32),
                ("kps", vp), ("idx", vp), ("idx_scalar", C.c_int64), ("extrinsic", vp),
                ("extrinsic_per_ray", C.c_int32), ("N", C.c_int64), ("kps_f32", vp)]


class RayprepArgs(C.Structure):
    _fields_ = [("cam", C.POINTER(Camera)), ("focal", C.c_float), ("H", C.c_int32), ("W", C.c_int32),
                ("ndc", C.c_int32), ("use_viewdirs", C.c_int32), ("near_", C.c_float),
                ("far_", C.c_float), ("N", C.c_int64)]


class Mlp(C.Structure):
    _fields_ = [("D", C.c_int32), ("W", C.c_int32), ("input_ch", C.c_int32),
                ("input_ch_views", C.c_int32), ("skip", C.c_int32), ("use_viewdirs", C.c_int32),
                ("output_ch", C.c_int32), ("L_pos", C.c_int32), ("L_dir", C.c_int32),
                ("pts_w", vp * MAX_DEPTH), ("pts_b", vp * MAX_DEPTH),
                ("views_w", vp), ("views_b", vp), ("feature_w", vp), ("feature_b", vp),
                ("alpha_w", vp), ("alpha_b", vp), ("rgb_w", vp), ("rgb_b", vp),
                ("output_w", vp), ("output_b", vp), ("pts_dim", C.c_int32)]


class PPRaygenArgs(C.Structure):          # include/scnerf_b200_nerfpp.h
    _fields_ = [("cam", C.POINTER(Camera)), ("distortion_initial", vp), ("distortion_noise", vp),
                ("distortion_noise_scale", C.c_float), ("select_inds", vp), ("camera_idx", C.c_int64),
                ("extrinsic", vp), ("N", C.c_int64)]


class PPStepCfg(C.Structure):            # scnerf_pp_step_cfg
    _fields_ = [("cascade_level", C.c_int32), ("cascade_samples", C.c_int32 * 2), ("precision", C.c_int32),
                ("perturb", C.c_int32), ("min_depth", C.c_float), ("seed", C.c_uint64)]


class PPStepRand(C.Structure):           # scnerf_pp_step_rand
    _fields_ = [("t_fg", vp), ("t_bg", vp), ("u_fg", vp), ("u_bg", vp)]


class PPNets(C.Structure):               # scnerf_pp_nets
    _fields_ = [("fg", C.POINTER(Mlp) * 2), ("bg", C.POINTER(Mlp) * 2)]


class PPStepIO(C.Structure):             # scnerf_pp_step_io
    _fields_ = [("select_inds_host", vp), ("target_host", vp), ("loss_host", vp), ("select_inds_dev", vp),
                ("target_dev", vp), ("loss_dev", vp), ("min_depth_dev", vp), ("miss_dev", vp), ("rgb_dev", vp)]


EXTRA_STRUCTS = [("scnerf_pp_step_cfg", PPStepCfg), ("scnerf_pp_step_rand", PPStepRand), ("scnerf_pp_nets", PPNets),
                 ("scnerf_pp_step_io", PPStepIO)]


class AdamTensor(C.Structure):
    _fields_ = [("param", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp), ("max_exp_avg_sq", vp),
                ("numel", C.c_int64), ("step", C.c_int32), ("decay", C.c_int32)]


class RenderCfg(C.Structure):
    _fields_ = [("N_samples", C.c_int32), ("N_importance", C.c_int32), ("ray_cols", C.c_int32),
                ("lindisp", C.c_int32), ("white_bkgd", C.c_int32), ("perturb", C.c_int32),
                ("raw_noise_std", C.c_float), ("retraw", C.c_int32), ("training", C.c_int32),
                ("precision", C.c_int32), ("seed", C.c_uint64)]


class RenderRand(C.Structure):
    _fields_ = [("t_rand", vp), ("u", vp), ("noise0", vp), ("noise1", vp)]


class RenderOut(C.Structure):
    _fields_ = [("rgb_map", vp), ("disp_map", vp), ("acc_map", vp), ("rgb0", vp), ("disp0", vp),
                ("acc0", vp), ("z_std", vp), ("raw", vp), ("z_vals", vp), ("weights", vp),
                ("depth_map", vp)]


class RenderGradsIn(C.Structure):
    _fields_ = [("rgb_map", vp), ("disp_map", vp), ("acc_map", vp), ("rgb0", vp), ("disp0", vp),
                ("acc0", vp)]


class StepIO(C.Structure):
    _fields_ = [("kps_host", vp), ("idx_host", vp), ("target_host", vp), ("loss_host", vp),
                ("kps_dev", vp), ("idx_dev", vp), ("target_dev", vp), ("loss_dev", vp)]


# name -> (restype, argtypes); every symbol include/scnerf_b200.h declares
_I, _I64, _SZ, _P = C.c_int, C.c_int64, C.c_size_t, C.POINTER
SIGNATURES = {
    "scnerf_last_error": (C.c_char_p, []),
    "scnerf_abi_version": (_I, []),
    "scnerf_built_for_sm": (_I, []),
    "scnerf_device_sm": (_I, []),
    "scnerf_launch_count": (_I64, [C.c_int32]),
    "scnerf_kernel_timing": (_I, [C.c_int32]),
    "scnerf_kernel_timing_report": (_I64, [C.c_char_p, _I64]),
    "scnerf_debug_mma_bench": (_I, [C.c_int32, C.c_int32, vp, C.c_int32, vp]),
    "scnerf_debug_timeline": (_I, [vp, C.c_int32]),
    "scnerf_debug_slab_plan": (_I, [C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "scnerf_tc_selftest": (_I, [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, _SZ, vp]),
    "scnerf_searchsorted_f32": (_I, [vp, vp, vp, _I64, _I64, _I64, _I64, _I, vp]),
    "scnerf_camera_matrices": (_I, [_P(Camera), vp, vp, vp]),
    "scnerf_camera_pair_fwd": (_I, [_P(Camera), _I64, _I64, C.c_float, vp, vp, vp]),
    "scnerf_camera_pair_bwd": (_I, [_P(Camera), _I64, _I64, C.c_float, vp, vp, _P(CameraGrads), vp]),
    "scnerf_raygen_fwd": (_I, [_P(RaygenArgs), vp, vp, vp]),
    "scnerf_raygen_bwd": (_I, [_P(RaygenArgs), vp, vp, _P(CameraGrads), vp]),
    "scnerf_rayprep_fwd": (_I, [_P(RayprepArgs), vp, vp, vp, vp]),
    "scnerf_rayprep_bwd": (_I, [_P(RayprepArgs), vp, vp, vp, vp, vp, vp, vp]),
    "scnerf_render_workspace_bytes": (_SZ, [_P(RenderCfg), _P(Mlp), _I64]),
    "scnerf_render_rays_fwd": (_I, [_P(RenderCfg), vp, _I64, _P(Mlp), _P(Mlp), _P(RenderRand),
                                    _P(RenderOut), vp, _SZ, vp]),
    "scnerf_render_rays_bwd": (_I, [_P(RenderCfg), vp, _I64, _P(Mlp), _P(Mlp), _P(RenderRand),
                                    _P(RenderGradsIn), _P(Mlp), _P(Mlp), vp, vp, _SZ, vp]),
    "scnerf_posenc_fwd": (_I, [vp, _I64, C.c_int32, vp, vp]),
    "scnerf_field_workspace_bytes": (_SZ, [_P(Mlp), _I64, C.c_int32]),
    "scnerf_field_infer_workspace_bytes": (_SZ, [_P(Mlp), _I64, C.c_int32]),
    "scnerf_field_fwd": (_I, [_P(Mlp), vp, vp, _I64, _I64, vp, C.c_int32, vp, _SZ, vp]),
    "scnerf_raw2outputs_fwd": (_I, [vp, C.c_int32, vp, vp, C.c_int32, vp, C.c_int32, _I64, _I64,
                                    vp, vp, vp, vp, vp, vp]),
    "scnerf_sample_pdf_fwd": (_I, [vp, vp, vp, _I64, _I64, _I64, vp, vp, vp]),
    "scnerf_sort_merge": (_I, [vp, vp, _I64, _I64, _I64, vp, vp]),
    "scnerf_train_step_workspace_bytes": (_SZ, [_P(RenderCfg), _P(Mlp), _I64]),
    "scnerf_train_step": (_I, [_P(Camera), _P(CameraGrads), _P(RenderCfg), C.c_int32, C.c_float,
                               C.c_float, _P(Mlp), _P(Mlp), _P(Mlp), _P(Mlp), _P(StepIO), C.c_int32,
                               _I64, vp, _SZ, vp]),
    "scnerf_prd_loss_fwd": (_I, [vp] * 8 + [C.c_float, C.c_float, C.c_int32, _I64, vp, vp, vp, vp]),
    "scnerf_prd_loss_bwd": (_I, [vp] * 8 + [C.c_float, C.c_float, _I64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "scnerf_ray_batch": (_I, [vp, _I64, vp, vp, _I64, C.c_int32, C.c_int32, vp, vp, vp, vp]),
    "scnerf_adam_step": (_I, [_P(AdamTensor), C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, vp]),
    # ---- include/scnerf_b200_nerfpp.h ----
    "scnerf_pp_raygen_fwd": (_I, [_P(PPRaygenArgs), vp, vp, vp, vp]),
    "scnerf_pp_raygen_bwd": (_I, [_P(PPRaygenArgs), vp, vp, _P(CameraGrads), vp, vp]),
    "scnerf_pp_intersect_sphere_fwd": (_I, [vp, vp, _I64, vp, vp, vp]),
    "scnerf_pp_intersect_sphere_bwd": (_I, [vp, vp, vp, _I64, vp, vp, vp]),
    "scnerf_pp_level0_depths": (_I, [vp, C.c_float, _I64, _I64, vp, vp, vp, vp, vp, vp]),
    "scnerf_pp_level0_depths_rays": (_I, [vp, vp, _I64, _I64, vp, vp, vp, vp, vp, vp]),
    "scnerf_pp_train_step_workspace_bytes": (_SZ, [_P(PPStepCfg), _P(Mlp), _P(Mlp), _I64]),
    "scnerf_pp_train_step": (_I, [_P(PPRaygenArgs), _P(CameraGrads), vp, _P(PPStepCfg), _P(PPNets), _P(PPNets),
                                  _P(PPStepRand), _P(PPStepIO), C.c_int32, _I64, vp, _SZ, vp]),
    "scnerf_pp_sample_pdf": (_I, [vp, vp, vp, vp, _I64, _I64, _I64, vp, vp, vp, vp, vp]),
    "scnerf_pp_sample_pdf_bins": (_I, [vp, vp, vp, _I64, _I64, _I64, vp, vp, vp, vp]),
    "scnerf_pp_sample_pdf_bins_bwd": (_I, [vp, vp, vp, _I64, _I64, _I64, vp, vp]),
    "scnerf_pp_perturb_samples_fwd": (_I, [vp, vp, _I64, _I64, vp, vp]),
    "scnerf_pp_perturb_samples_bwd": (_I, [vp, vp, _I64, _I64, vp, vp]),
    "scnerf_pp_depth_bwd": (_I, [vp, vp, _I64, _I64, vp, vp]),
    "scnerf_field_train_workspace_bytes": (_SZ, [_P(Mlp), _I64, _I64, C.c_int32]),
    "scnerf_field_infer_rays_workspace_bytes": (_SZ, [_P(Mlp), _I64, _I64, C.c_int32]),
    "scnerf_field_infer_fwd": (_I, [_P(Mlp), vp, C.c_int32, vp, vp, vp, _I64, _I64, vp, C.c_int32, vp, _SZ, vp]),
    "scnerf_field_train_fwd": (_I, [_P(Mlp), vp, C.c_int32, vp, vp, vp, _I64, _I64, vp, C.c_int32, vp, _SZ, vp]),
    "scnerf_field_train_bwd": (_I, [_P(Mlp), _P(Mlp), vp, C.c_int32, vp, vp, vp, _I64, _I64, vp, vp, vp, vp, vp,
                                    C.c_int32, vp, _SZ, vp]),
    "scnerf_pp_pack_rays": (_I, [vp, vp, _I64, vp, vp]),
    "scnerf_pp_pack_rays_bwd": (_I, [vp, vp, _I64, vp, vp, vp]),
    "scnerf_pp_bg_points_fwd": (_I, [vp, vp, vp, _I64, _I64, vp, vp, vp]),
    "scnerf_pp_bg_points_bwd": (_I, [vp, vp, vp, vp, _I64, _I64, vp, vp, vp]),
    "scnerf_pp_composite_fg_fwd": (_I, [vp, vp, vp, vp, _I64, _I64, vp, vp, vp, vp, vp]),
    "scnerf_pp_composite_fg_bwd": (_I, [vp, vp, vp, vp, _I64, _I64, vp, vp, vp, vp, vp, vp, vp]),
    "scnerf_pp_composite_bg_fwd": (_I, [vp, vp, vp, vp, _I64, _I64, vp, vp, vp, vp, vp]),
    "scnerf_pp_composite_bg_bwd": (_I, [vp, vp, vp, _I64, _I64, vp, vp, vp, vp]),
}

_lib = None


def load():
    """Load the CUDA library (once).  Loud failure by design — there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"scnerf_b200: CUDA library not built ({LIB_PATH} missing). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or scnerf_b200/csrc/build.sh. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def kernel_times():
    """[(launch site, grid size, ms)] recorded since ``scnerf_kernel_timing(1)`` (launch order)."""
    lib = load()
    need = lib.scnerf_kernel_timing_report(None, 0)
    buf = C.create_string_buffer(int(need) + 16)
    lib.scnerf_kernel_timing_report(buf, len(buf))
    out = []
    for line in buf.value.decode().splitlines():
        name, grid, ms = line.rsplit("\t", 2)
        out.append((name, int(grid), float(ms)))
    return out


def check(rc, what=""):
    if rc != 0:
        msg = load().scnerf_last_error()
        raise RuntimeError(f"scnerf_b200 {what} failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL); enforces the ABI's layout contract."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("scnerf_b200: tensors must live on a CUDA device (no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("scnerf_b200: tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32(t):
    return t.detach().to(dtype=torch.float32).contiguous()


def i64(t):
    return t.detach().to(dtype=torch.int64).contiguous()
