"""nerfplusplus/nerf_network.py: Embedder (:11-60) and MLPNet (:68-142) as parameter owners.

The positional encoding is fused into the CUDA field kernels, so ``Embedder`` only carries the shape
bookkeeping (``out_dim``) its callers read; ``MLPNet`` owns ``nn.Linear`` parameters under the
reference's state-dict keys and hands them to the kernels through ``c_struct()``."""
import torch
import torch.nn as nn

from .. import _lib


class Embedder(nn.Module):
    def __init__(self, input_dim, max_freq_log2, N_freqs, log_sampling=True, include_input=True,
                 periodic_fns=(torch.sin, torch.cos)):
        super().__init__()
        if not (log_sampling and include_input and len(periodic_fns) == 2 and max_freq_log2 == N_freqs - 1):
            raise NotImplementedError("Embedder: the CUDA field kernels implement the log-sampled sin/cos "
                                      "encoding with the raw input included (the reference's only configuration)")
        self.input_dim, self.N_freqs, self.include_input = input_dim, N_freqs, include_input
        self.out_dim = input_dim * (1 + 2 * N_freqs)
        self.freq_bands = [2.0 ** i for i in range(N_freqs)]

    def forward(self, input):
        raise NotImplementedError("Embedder is fused into the CUDA field kernels (NerfNet.forward); "
                                  "it is not evaluated on its own")


class MLPNet(nn.Module):
    """nerf_network.py:68-118: same layers, same state-dict keys, PyTorch default init."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_viewdirs=3, skips=[4], use_viewdirs=False):
        super().__init__()
        if not use_viewdirs:
            raise NotImplementedError("MLPNet: the reference's MLPNet always consumes view directions")
        self.D, self.W = D, W
        self.input_ch, self.input_ch_viewdirs = input_ch, input_ch_viewdirs
        self.use_viewdirs, self.skips = use_viewdirs, list(skips)
        layers, dim = [], input_ch
        for i in range(D):
            layers.append(nn.Sequential(nn.Linear(dim, W), nn.ReLU()))
            dim = W
            if i in self.skips and i != D - 1:
                dim += input_ch
        self.base_layers = nn.ModuleList(layers)
        self.sigma_layers = nn.Sequential(nn.Linear(dim, 1))
        self.base_remap_layers = nn.Sequential(nn.Linear(dim, 256))
        self.rgb_layers = nn.Sequential(nn.Linear(256 + input_ch_viewdirs, W // 2), nn.ReLU(),
                                        nn.Linear(W // 2, 3), nn.Sigmoid())

    def tensor_core_shape(self):
        """True for the shape the tcgen05 kernels are specialised for: 8 x 256, skip 4, view directions, 10 / 4 frequencies
        (63 channels for the foreground's 3-D points, 84 for the background's 4-D ones)."""
        return (self.D == 8 and self.W == 256 and self.skips == [4] and bool(self.use_viewdirs)
                and self.input_ch in (63, 84) and self.input_ch_viewdirs == 27)

    def field_tensors(self):
        """Parameters in the order of scnerf_mlp: trunk, views (rgb_layers.0), feature (base_remap),
        alpha (sigma), rgb (rgb_layers.2)."""
        ts = []
        for l in self.base_layers:
            ts += [l[0].weight, l[0].bias]
        for l in (self.rgb_layers[0], self.base_remap_layers[0], self.sigma_layers[0], self.rgb_layers[2]):
            ts += [l.weight, l.bias]
        return ts

    def c_struct(self, tensors=None, pts_dim=3):
        ts = [t.detach() for t in (tensors if tensors is not None else self.field_tensors())]
        m = _lib.Mlp()
        m.D, m.W, m.input_ch, m.input_ch_views = self.D, self.W, self.input_ch, self.input_ch_viewdirs
        m.skip = self.skips[0] if len(self.skips) else -1
        m.use_viewdirs, m.output_ch, m.pts_dim = 1, 4, pts_dim
        m.L_pos = (self.input_ch // pts_dim - 1) // 2
        m.L_dir = (self.input_ch_viewdirs - 3) // 6
        for i in range(self.D):
            m.pts_w[i] = _lib.ptr(ts[2 * i]).value
            m.pts_b[i] = _lib.ptr(ts[2 * i + 1]).value
        (m.views_w, m.views_b, m.feature_w, m.feature_b, m.alpha_w, m.alpha_b, m.rgb_w,
         m.rgb_b) = [_lib.ptr(t) for t in ts[2 * self.D:]]
        m._keep = ts
        return m

    def forward(self, input):
        raise NotImplementedError("MLPNet is evaluated by the CUDA field kernels from points "
                                  "(NerfNet.forward); it has no embedded-feature entry")
