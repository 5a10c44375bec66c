"""Whole-step engine: one C-ABI call (scnerf_train_step) runs pixels -> rays -> render -> loss ->
all parameter gradients (NeRF/run_nerf.py:385-506,600), with device-resident or pinned-host
inputs.  This is what bench.py times; the trainers can use it instead of the autograd path.
"""
import ctypes as C

import torch

from . import _lib
from .parallel import FlatGrads
from .run_nerf_helpers import unwrap


class TrainStep:
    def __init__(self, camera_model, network_fn, network_fine, N_rays, N_samples, N_importance,
                 perturb=1.0, raw_noise_std=1.0, white_bkgd=False, lindisp=False, ndc=True, near=0.,
                 far=1., precision=None, seed=0):
        self.lib = _lib.load()
        self.cam = camera_model
        self.net_c, self.net_f = unwrap(network_fn), unwrap(network_fine) if network_fine is not None else None
        dev = self.cam.intrinsics_initial.device
        self.dev, self.N = dev, int(N_rays)
        self.ndc, self.near, self.far = int(ndc), float(near), float(far)
        cfg = _lib.RenderCfg()
        cfg.N_samples, cfg.N_importance = int(N_samples), int(N_importance)
        cfg.ray_cols = 11 if self.net_c.use_viewdirs else 8
        cfg.lindisp, cfg.white_bkgd = int(lindisp), int(white_bkgd)
        cfg.perturb, cfg.raw_noise_std = int(perturb > 0), float(raw_noise_std)
        cfg.training, cfg.precision, cfg.seed = 1, _lib.PRECISION[_lib.resolve_precision(precision, self.net_c, self.net_f)], int(seed)
        self.cfg = cfg
        named = [("coarse." + str(i), t) for i, t in enumerate(self.net_c.field_tensors())]
        if self.net_f is not None:
            named += [("fine." + str(i), t) for i, t in enumerate(self.net_f.field_tensors())]
        named += [("camera." + n, getattr(self.cam, n)) for n in self.cam.LEARNABLE]
        self.grads = FlatGrads(named, dev)
        nc = len(self.net_c.field_tensors())
        gviews = [self.grads.views[n] for n, _ in named]
        self.g_c = self.net_c.c_struct(gviews[:nc])
        self.g_f = self.net_f.c_struct(gviews[nc:nc + len(self.net_f.field_tensors())]) if self.net_f is not None else None
        self.g_cam = _lib.CameraGrads()
        for n in self.cam.LEARNABLE:
            setattr(self.g_cam, n, _lib.ptr(self.grads.views["camera." + n]))
        m = self.net_c.c_struct()
        self.ws_bytes = self.lib.scnerf_train_step_workspace_bytes(C.byref(cfg), C.byref(m), self.N)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        N = self.N
        self.kps_dev = torch.empty(N, 2, dtype=torch.int64, device=dev)
        self.idx_dev = torch.empty(N, dtype=torch.int64, device=dev)
        self.target_dev = torch.empty(N, 3, dtype=torch.float32, device=dev)
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.kps_host = torch.empty(N, 2, dtype=torch.int64).pin_memory()
        self.idx_host = torch.empty(N, dtype=torch.int64).pin_memory()
        self.target_host = torch.empty(N, 3, dtype=torch.float32).pin_memory()
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        io = _lib.StepIO()
        io.kps_host, io.idx_host = self.kps_host.data_ptr(), self.idx_host.data_ptr()
        io.target_host, io.loss_host = self.target_host.data_ptr(), self.loss_host.data_ptr()
        io.kps_dev, io.idx_dev = self.kps_dev.data_ptr(), self.idx_dev.data_ptr()
        io.target_dev, io.loss_dev = self.target_dev.data_ptr(), self.loss_dev.data_ptr()
        self.io = io
        self.h2d_bytes = N * (16 + 8 + 12)
        self.d2h_bytes = 4

    def _call(self, on_host, zero=True):
        cam = self.cam.c_struct()
        mc = self.net_c.c_struct()
        mf = self.net_f.c_struct() if self.net_f is not None else None
        self.cfg.seed = (self.cfg.seed + 1) & 0xFFFFFFFFFFFFFFFF
        if zero:
            self.grads.zero_()
        _lib.check(self.lib.scnerf_train_step(
            C.byref(cam), C.byref(self.g_cam), C.byref(self.cfg), self.ndc, self.near, self.far,
            C.byref(mc), C.byref(mf) if mf is not None else None, C.byref(self.g_c),
            C.byref(self.g_f) if self.g_f is not None else None, C.byref(self.io), int(on_host),
            self.N, _lib.ptr(self.ws), self.ws_bytes, _lib.stream()), "train_step")

    def assign_grads(self):
        """Point every parameter's ``.grad`` at its view of the flat buffer, so autograd terms computed outside the
        fused step (the PRD loss) accumulate into the same buffer and torch optimisers read it without copies."""
        for i, p in enumerate(self.net_c.field_tensors()):
            p.grad = self.grads.views[f"coarse.{i}"]
        if self.net_f is not None:
            for i, p in enumerate(self.net_f.field_tensors()):
                p.grad = self.grads.views[f"fine.{i}"]
        for n in self.cam.LEARNABLE:
            getattr(self.cam, n).grad = self.grads.views["camera." + n]

    def step_device(self, kps=None, idx=None, target=None, zero=True):
        """Inputs already in HBM (copied into the staging tensors if given).  Returns the device
        loss tensor; gradients are in ``self.grads`` (flat, per-parameter views).  ``zero=False`` accumulates on top of
        what the buffer holds (e.g. the PRD term's camera gradients, computed first)."""
        if kps is not None:
            self.kps_dev.copy_(kps); self.idx_dev.copy_(idx); self.target_dev.copy_(target)
        self._call(False, zero)
        return self.loss_dev

    def step_host(self, kps=None, idx=None, target=None, zero=True):
        """Inputs in pinned host memory; H2D copies and the D2H loss read are part of the call.
        The loss is valid in ``self.loss_host`` after the stream is synchronised."""
        if kps is not None:
            self.kps_host.copy_(kps); self.idx_host.copy_(idx); self.target_host.copy_(target)
        self._call(True, zero)
        return self.loss_host
