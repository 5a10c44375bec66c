"""Whole-step engine of the NeRF++ trainer: ONE C-ABI call (scnerf_pp_train_step) runs pixel indices -> rays ->
sphere exit -> cascade {depths, fg field, fg composite, sphere points, bg field, bg composite, img2mse} -> loss ->
every gradient (both networks of every cascade level + the camera), nerfplusplus/ddp_train_nerf.py:421-488,552, with
device-resident or pinned-host inputs.  What ``bench.py --workload c4|c5`` times; a trainer can use it in place of
the per-stage autograd path (``NerfNet.forward`` etc.), which remains the drop-in face.

Gradients land in one flat fp32 buffer ``[net_0.fg | net_0.bg | net_1.fg | net_1.bg | camera]`` (``self.grads``) that
``FlatGrads.all_reduce_mean()`` moves with one NCCL all-reduce (9.7 MB for two levels).
"""
import ctypes as C

import torch

from .. import _lib
from ..parallel import FlatGrads

CAM_NAMES = ("intrinsics_noise", "extrinsics_noise", "ray_o_noise", "ray_d_noise")


def _nerf_net(net):
    """NerfNetWithAutoExpo / DDP / wrapper -> the NerfNet holding fg_net and bg_net."""
    for attr in ("module", "nerf_net"):
        while hasattr(net, attr):
            net = getattr(net, attr)
    return net


class PPTrainStep:
    def __init__(self, camera_model, nets, N_rays, cascade_samples, camera_idx=0, perturb=True, min_depth=1e-4,
                 precision=None, seed=0, keep_rgb=False):
        self.lib = _lib.load()
        self.cam = camera_model
        self.nets = [_nerf_net(n) for n in nets]
        self.N, self.L = int(N_rays), len(self.nets)
        assert self.L in (1, 2) and len(cascade_samples) >= self.L
        dev = self.cam.intrinsics_initial.device
        self.dev = dev
        cfg = _lib.PPStepCfg()
        cfg.cascade_level = self.L
        cfg.cascade_samples[0] = int(cascade_samples[0])
        cfg.cascade_samples[1] = int(cascade_samples[1]) if self.L > 1 else 0
        cfg.precision = _lib.PRECISION[precision or self.nets[0].precision]
        cfg.perturb, cfg.min_depth, cfg.seed = int(bool(perturb)), float(min_depth), int(seed)
        self.cfg = cfg
        self.camera_idx = int(camera_idx)
        named = []
        for m, net in enumerate(self.nets):
            named += [(f"net{m}.fg.{i}", t) for i, t in enumerate(net.fg_net.field_tensors())]
            named += [(f"net{m}.bg.{i}", t) for i, t in enumerate(net.bg_net.field_tensors())]
        self.has_dist = hasattr(self.cam, "distortion_noise")
        cam_names = CAM_NAMES + (("distortion_noise",) if self.has_dist else ())
        named += [("camera." + n, getattr(self.cam, n)) for n in cam_names]
        self.grads = FlatGrads(named, dev)
        self._g_structs = []
        self.g_nets = _lib.PPNets()
        for m, net in enumerate(self.nets):
            nf = len(net.fg_net.field_tensors())
            gf = net.fg_net.c_struct([self.grads.views[f"net{m}.fg.{i}"] for i in range(nf)], pts_dim=3)
            gb = net.bg_net.c_struct([self.grads.views[f"net{m}.bg.{i}"] for i in range(nf)], pts_dim=4)
            self._g_structs += [gf, gb]
            self.g_nets.fg[m], self.g_nets.bg[m] = C.pointer(gf), C.pointer(gb)
        self.g_cam = _lib.CameraGrads()
        for n in CAM_NAMES:
            setattr(self.g_cam, n, _lib.ptr(self.grads.views["camera." + n]))
        self.g_dist = self.grads.views["camera.distortion_noise"] if self.has_dist else None
        m_fg, m_bg = self.nets[0].fg_net.c_struct(pts_dim=3), self.nets[0].bg_net.c_struct(pts_dim=4)
        self.ws_bytes = self.lib.scnerf_pp_train_step_workspace_bytes(C.byref(cfg), C.byref(m_fg), C.byref(m_bg), self.N)
        if self.ws_bytes == 0:
            raise RuntimeError("scnerf_pp_train_step: unsupported configuration")
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        N = self.N
        self.sel_dev = torch.empty(N, dtype=torch.int64, device=dev)
        self.target_dev = torch.empty(N, 3, dtype=torch.float32, device=dev)
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.miss_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.rgb_dev = torch.empty(self.L, N, 3, dtype=torch.float32, device=dev) if keep_rgb else None
        self.sel_host = torch.empty(N, dtype=torch.int64).pin_memory()
        self.target_host = torch.empty(N, 3, dtype=torch.float32).pin_memory()
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.min_depth_dev = None            # optional [N] per-ray near depth (ray_batch['min_depth'])
        self.h2d_bytes, self.d2h_bytes = N * (8 + 12), 4

    def _io(self):
        io = _lib.PPStepIO()
        io.select_inds_host, io.target_host, io.loss_host = (self.sel_host.data_ptr(), self.target_host.data_ptr(),
                                                              self.loss_host.data_ptr())
        io.select_inds_dev, io.target_dev, io.loss_dev = (self.sel_dev.data_ptr(), self.target_dev.data_ptr(),
                                                           self.loss_dev.data_ptr())
        io.min_depth_dev = self.min_depth_dev.data_ptr() if self.min_depth_dev is not None else None
        io.miss_dev = self.miss_dev.data_ptr()
        io.rgb_dev = self.rgb_dev.data_ptr() if self.rgb_dev is not None else None
        return io

    def _call(self, on_host, rand=None):
        keep = []
        cam = self.cam.c_struct()
        a = _lib.PPRaygenArgs()
        a.cam = C.pointer(cam)
        if self.has_dist:
            a.distortion_initial = _lib.ptr(self.cam.distortion_initial.detach())
            a.distortion_noise = _lib.ptr(self.cam.distortion_noise.detach())
            a.distortion_noise_scale = float(self.cam.args.distortion_noise_scale)
        a.camera_idx = self.camera_idx
        nets = _lib.PPNets()
        for m, net in enumerate(self.nets):
            f, b = net.fg_net.c_struct(pts_dim=3), net.bg_net.c_struct(pts_dim=4)
            keep += [f, b]
            nets.fg[m], nets.bg[m] = C.pointer(f), C.pointer(b)
        rnd = None
        if rand is not None:
            rnd = _lib.PPStepRand()
            for k in ("t_fg", "t_bg", "u_fg", "u_bg"):
                t = rand.get(k)
                if t is not None:
                    t = _lib.f32(t)
                    keep.append(t)
                    setattr(rnd, k, _lib.ptr(t))
        self.cfg.seed = (self.cfg.seed + 1) & 0xFFFFFFFFFFFFFFFF
        self.grads.zero_()
        io = self._io()
        _lib.check(self.lib.scnerf_pp_train_step(
            C.byref(a), C.byref(self.g_cam), _lib.ptr(self.g_dist), C.byref(self.cfg), C.byref(nets), C.byref(self.g_nets),
            C.byref(rnd) if rnd is not None else None, C.byref(io), int(on_host), self.N, _lib.ptr(self.ws),
            self.ws_bytes, _lib.stream()), "pp_train_step")

    def step_device(self, select_inds=None, target=None, rand=None):
        """Inputs already in HBM.  Returns the device loss tensor; gradients are in ``self.grads``."""
        if select_inds is not None:
            self.sel_dev.copy_(select_inds); self.target_dev.copy_(target)
        self._call(False, rand)
        return self.loss_dev

    def step_host(self, select_inds=None, target=None, rand=None):
        """Inputs in pinned host memory; H2D copies and the D2H loss read are inside the call."""
        if select_inds is not None:
            self.sel_host.copy_(select_inds); self.target_host.copy_(target)
        self._call(True, rand)
        return self.loss_host

    def check_sphere(self):
        """The reference raises when a ray never enters the unit sphere (ddp_train_nerf.py:61-65); the fused step
        counts such rays instead of synchronising.  Call this where a device sync is acceptable."""
        if int(self.miss_dev.item()) > 0:
            raise Exception("Not all your cameras are bounded by the unit sphere; please make sure "
                            "the cameras are normalized properly!")

    def assign_grads(self):
        """Point every parameter's ``.grad`` at its view of the flat buffer (for torch optimisers)."""
        for m, net in enumerate(self.nets):
            for i, p in enumerate(net.fg_net.field_tensors()):
                p.grad = self.grads.views[f"net{m}.fg.{i}"]
            for i, p in enumerate(net.bg_net.field_tensors()):
                p.grad = self.grads.views[f"net{m}.bg.{i}"]
        for n in CAM_NAMES + (("distortion_noise",) if self.has_dist else ()):
            getattr(self.cam, n).grad = self.grads.views["camera." + n]
