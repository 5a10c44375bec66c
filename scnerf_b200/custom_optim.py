"""The reference's positional-weight-decay Adam (NeRF/create_nerf.py:199-336; nerfplusplus/custom_optim.py:11-147)
as one fused multi-tensor CUDA launch per step (`scnerf_adam_step`), plus its learning-rate schedule
(NeRF/run_nerf.py:617-621; nerfplusplus/ddp_train_nerf.py:386-392).  SURVEY.md §8 row f2."""
import torch
from torch import optim

from . import _lib


def decay_index_from(n_params, args):
    """First parameter POSITION that receives weight decay (create_nerf.py:222-230): the trailing ray_o /
    ray_d / distortion tensors of ``grad_vars`` — identified by the camera-model NAME, not by the tensors."""
    k = n_params
    if args.camera_model != "none":
        k -= "rayo" in args.camera_model
        k -= "rayd" in args.camera_model
        k -= "dist" in args.camera_model
    return k


class CustomAdamOptimizer(optim.Optimizer):
    def __init__(self, params, lr, args, H, W, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)
        self.args = args
        self.H, self.W = H, W

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            with_grad = [p for p in group["params"] if p.grad is not None]
            k0 = decay_index_from(len(with_grad), self.args)
            for p in with_grad:
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group["amsgrad"]:
                        state["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
            # The pointer table is rebuilt only when a tensor moved (new .grad storage, reloaded state): with gradients
            # living in a persistent flat buffer (engine.TrainStep.assign_grads) it is built once — ~50 ctypes records
            # per step were 0.5 ms of host time on the critical path between two render steps.
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in with_grad]
            key = tuple((p.data_ptr(), g.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for p, g in zip(with_grad, grads))
            cache = self.__dict__.setdefault("_tables", {})
            ent = cache.get(gi)
            if ent is None or ent[0] != key or ent[1] != bool(group["amsgrad"]):
                tab = (_lib.AdamTensor * max(len(with_grad), 1))()
                for i, (p, g) in enumerate(zip(with_grad, grads)):
                    state = self.state[p]
                    t = tab[i]
                    t.param, t.grad = _lib.ptr(p.data), _lib.ptr(g)
                    t.exp_avg, t.exp_avg_sq = _lib.ptr(state["exp_avg"]), _lib.ptr(state["exp_avg_sq"])
                    t.max_exp_avg_sq = _lib.ptr(state["max_exp_avg_sq"]) if group["amsgrad"] else None
                    t.numel, t.decay = p.numel(), int(i >= k0)
                ent = cache[gi] = (key, bool(group["amsgrad"]), tab)
            tab = ent[2]
            for i, p in enumerate(with_grad):
                tab[i].step = self.state[p]["step"]
            beta1, beta2 = group["betas"]
            _lib.check(lib.scnerf_adam_step(tab, len(with_grad), float(group["lr"]), float(beta1), float(beta2),
                                            float(group["eps"]), float(group["weight_decay"]), _lib.stream()),
                       "adam_step")
            del grads
        return loss


def decayed_lrate(lrate, lrate_decay, global_step, decay_rate=0.1):
    """NeRF/run_nerf.py:617-621: lrate * 0.1 ** (global_step / (lrate_decay * 1000))."""
    return lrate * (decay_rate ** (global_step / (lrate_decay * 1000)))


def update_lrate(optimizer, lrate, lrate_decay, global_step):
    new_lrate = decayed_lrate(lrate, lrate_decay, global_step)
    for group in optimizer.param_groups:
        group["lr"] = new_lrate
    return new_lrate
