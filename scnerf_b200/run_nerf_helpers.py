"""MLP / positional-encoding objects with the reference's names (NeRF/run_nerf_helpers.py).

``NeRF`` owns parameters with the reference's state_dict keys and init; the rendering path reads
them through ``c_struct()`` inside the CUDA field kernels (PE is fused there, so ``Embedder`` is
only an API-parity object that records the number of frequencies).
"""
import random

import numpy as np
import torch
import torch.nn as nn

from . import _lib

# NOTE: the reference turns on torch.autograd.set_detect_anomaly(True) at import
# (run_nerf_helpers.py:7).  Deliberately not reproduced (debug aid, large slowdown).

img2mse = lambda x, y: torch.mean((x - y) ** 2)                                   # noqa: E731
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))  # noqa: E731


class DenseLayer(nn.Linear):
    """nn.Linear with xavier-uniform(gain(activation)) weights and zero bias
    (NeRF/run_nerf_helpers.py:13-21)."""

    def __init__(self, in_dim, out_dim, activation="relu", *args, **kwargs):
        self.activation = activation
        super().__init__(in_dim, out_dim, *args, **kwargs)

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight, gain=nn.init.calculate_gain(self.activation))
        if self.bias is not None:
            nn.init.zeros_(self.bias)


class Embedder:
    """NeRF/run_nerf_helpers.py:24-54.  ``embed`` runs the CUDA PE kernel (scnerf_posenc_fwd)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d = kwargs["input_dims"]
        assert d == 3 and kwargs["include_input"] and kwargs["log_sampling"], \
            "only the configuration the trainers use (3-D input, include_input, log sampling)"
        assert kwargs["max_freq_log2"] == kwargs["num_freqs"] - 1
        self.num_freqs = int(kwargs["num_freqs"])
        self.out_dim = d * (1 + 2 * self.num_freqs)

    def embed(self, inputs):
        lib = _lib.load()
        x = _lib.f32(inputs).reshape(-1, 3)
        out = torch.empty(x.shape[0], self.out_dim, device=x.device, dtype=torch.float32)
        _lib.check(lib.scnerf_posenc_fwd(_lib.ptr(x), x.shape[0], self.num_freqs, _lib.ptr(out),
                                         _lib.stream()), "posenc_fwd")
        return out.reshape(*inputs.shape[:-1], self.out_dim)


def get_embedder(multires, i=0):
    """NeRF/run_nerf_helpers.py:57-72.  Returns (embed_fn, out_dim); embed_fn.num_freqs is what
    the fused kernels consume."""
    if i == -1:
        ident = nn.Identity()
        ident.num_freqs = 0
        return ident, 3
    eo = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos])

    def embed(x, eo=eo):
        return eo.embed(x)
    embed.num_freqs = eo.num_freqs
    return embed, eo.out_dim


class NeRF(nn.Module):
    """NeRF/run_nerf_helpers.py:76-128: same constructor, parameters, state_dict keys."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4],
                 use_viewdirs=False):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips, self.use_viewdirs, self.output_ch = skips, use_viewdirs, output_ch
        assert len(skips) <= 1, "one skip connection (the reference always uses skips=[4])"
        self.pts_linears = nn.ModuleList(
            [DenseLayer(input_ch, W, activation="relu")]
            + [DenseLayer(W + input_ch if i in skips else W, W, activation="relu") for i in range(D - 1)])
        self.views_linears = nn.ModuleList([DenseLayer(input_ch_views + W, W // 2, activation="relu")])
        if use_viewdirs:
            self.feature_linear = DenseLayer(W, W, activation="linear")
            self.alpha_linear = DenseLayer(W, 1, activation="linear")
            self.rgb_linear = DenseLayer(W // 2, 3, activation="linear")
        else:
            self.output_linear = DenseLayer(W, output_ch, activation="linear")

    # ---- C-ABI view ------------------------------------------------------------------------------
    def field_tensors(self):
        """Parameters in the fixed order the autograd wrapper uses (weights then biases per layer)."""
        ts = []
        for l in self.pts_linears:
            ts += [l.weight, l.bias]
        if self.use_viewdirs:
            for l in (self.views_linears[0], self.feature_linear, self.alpha_linear, self.rgb_linear):
                ts += [l.weight, l.bias]
        else:
            ts += [self.output_linear.weight, self.output_linear.bias]
        return ts

    def c_struct(self, tensors=None):
        """ctypes scnerf_mlp over ``tensors`` (defaults to the parameters; pass gradient buffers
        in ``field_tensors()`` order to describe a gradient struct)."""
        ts = [t.detach() for t in (tensors if tensors is not None else self.field_tensors())]
        m = _lib.Mlp()
        m.D, m.W, m.input_ch, m.input_ch_views = self.D, self.W, self.input_ch, self.input_ch_views
        m.skip = self.skips[0] if len(self.skips) else -1
        m.use_viewdirs, m.output_ch = int(self.use_viewdirs), self.output_ch
        m.L_pos = (self.input_ch - 3) // 6
        m.L_dir = (self.input_ch_views - 3) // 6 if self.use_viewdirs else 0
        for i in range(self.D):
            m.pts_w[i] = _lib.ptr(ts[2 * i]).value
            m.pts_b[i] = _lib.ptr(ts[2 * i + 1]).value
        rest = ts[2 * self.D:]
        if self.use_viewdirs:
            (m.views_w, m.views_b, m.feature_w, m.feature_b, m.alpha_w, m.alpha_b, m.rgb_w,
             m.rgb_b) = [_lib.ptr(t) for t in rest]
        else:
            m.output_w, m.output_b = _lib.ptr(rest[0]), _lib.ptr(rest[1])
        m._keep = ts
        return m

    def tensor_core_shape(self):
        """True for the one network shape the tcgen05 kernels are specialised for (the reference's default: 8 x 256, skip at
        layer 4, view directions, multires 10 / 4).  Other shapes run on the fp32 CUDA-core kernels."""
        return (self.D == 8 and self.W == 256 and list(self.skips) == [4] and bool(self.use_viewdirs)
                and self.input_ch == 63 and self.input_ch_views == 27)

    def forward(self, x):
        """The reference's ``NeRF.forward`` (run_nerf_helpers.py:105-128) consumes pre-embedded features and
        is only ever reached through ``run_network`` (create_nerf.py:18-32).  Here the positional encoding
        is fused into the CUDA field kernels, which start from points, so there is no embedded-feature
        entry; evaluating this module with eager PyTorch ops would be a silent library fallback."""
        raise NotImplementedError(
            "scnerf_b200.NeRF is evaluated by the CUDA field kernels from points: call "
            "scnerf_b200.create_nerf.run_network(pts, viewdirs, net, ...) / network_query_fn instead of net(x)")


class SingleDeviceParallel(nn.Module):
    """Stand-in for the reference's nn.DataParallel wrapper (NeRF/create_nerf.py:56,64): keeps the
    ``module.`` prefix in state_dict keys so reference checkpoints load, without replicating the
    network (one process drives one GPU here)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def unwrap(net):
    return net.module if hasattr(net, "module") and isinstance(net.module, NeRF) else net


def fix_seeds(random_seed):
    """NeRF/run_nerf_helpers.py:160-169."""
    np.random.seed(random_seed)
    torch.manual_seed(random_seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(random_seed)
    torch.backends.cudnn.benchmark = False
    random.seed(random_seed)
