"""Factory with the reference's signature and return tuple (NeRF/create_nerf.py:34-184), wiring
this package's CUDA-backed modules: ``render_kwargs_train`` / ``render_kwargs_test`` feed
``render.render`` exactly as NeRF/run_nerf.py:225-239,482-492 expects.
"""
import ctypes as C
import os

import torch

from . import _lib
from .camera_dict import camera_dict
from .run_nerf_helpers import NeRF, SingleDeviceParallel, get_embedder, unwrap


def batchify(fn, chunk):
    """NeRF/create_nerf.py:187-196."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64, precision=None):
    """NeRF/create_nerf.py:18-32: pts [N,S,3] (+ viewdirs [N,3]) -> raw [N,S,4|5].

    One C-ABI call (scnerf_field_fwd): positional encoding of points and directions is fused with
    the MLP, so ``embed_fn`` / ``embeddirs_fn`` / ``netchunk`` only document the configuration
    (their frequency counts must match the network).  Forward only; gradients flow through
    ``render_rays``' fused node, which is what the trainers differentiate."""
    lib = _lib.load()
    net = unwrap(fn)
    if not isinstance(net, NeRF):
        raise NotImplementedError("run_network: fn must be scnerf_b200's NeRF module")
    if getattr(embed_fn, "num_freqs", (net.input_ch - 3) // 6) != (net.input_ch - 3) // 6:
        raise ValueError("embed_fn frequencies do not match the network's input_ch")
    pts = _lib.f32(inputs)
    N, S = pts.shape[0], pts.shape[1]
    vd = _lib.f32(viewdirs) if viewdirs is not None else None
    m = net.c_struct()
    rc = 4 if net.use_viewdirs else net.output_ch
    raw = torch.empty(N, S, rc, device=pts.device, dtype=torch.float32)
    prec = _lib.PRECISION[_lib.resolve_precision(precision, net)]
    nbytes = lib.scnerf_field_infer_workspace_bytes(C.byref(m), N * S, prec)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=pts.device)
    _lib.check(lib.scnerf_field_fwd(C.byref(m), _lib.ptr(pts), _lib.ptr(vd), N, S, _lib.ptr(raw), prec, _lib.ptr(ws), nbytes,
                                    _lib.stream()),
               "field_fwd")
    return raw


def create_nerf(args, noisy_focal, noisy_poses, H, W, mode="train", device="cuda"):
    """NeRF/create_nerf.py:34-184 -> (render_kwargs_train, render_kwargs_test, start, grad_vars,
    optimizer, camera_model).  grad_vars order: coarse MLP, fine MLP, camera (:57,65,123)."""
    camera_model = None
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]

    def make(D, Wd):
        net = NeRF(D=D, W=Wd, input_ch=input_ch, output_ch=output_ch, skips=skips,
                   input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs)
        return SingleDeviceParallel(net).to(device)

    model = make(args.netdepth, args.netwidth)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = make(args.netdepth_fine, args.netwidth_fine)
        grad_vars += list(model_fine.parameters())

    n_gpus = getattr(args, "n_gpus", 1) or 1
    netchunk = args.netchunk_per_gpu * n_gpus if hasattr(args, "netchunk_per_gpu") else 1024 * 64
    network_query_fn = lambda inputs, viewdirs, network_fn: run_network(   # noqa: E731
        inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=netchunk)

    render_kwargs_train = {
        "network_query_fn": network_query_fn, "perturb": args.perturb,
        "N_importance": args.N_importance, "network_fine": model_fine, "N_samples": args.N_samples,
        "network_fn": model, "use_viewdirs": args.use_viewdirs, "white_bkgd": args.white_bkgd,
        "raw_noise_std": args.raw_noise_std,
    }
    if args.dataset_type != "llff" or args.no_ndc:   # NDC only for forward-facing data (:83-87)
        print("Not ndc!")
        render_kwargs_train["ndc"] = False
        render_kwargs_train["lindisp"] = args.lindisp
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test["perturb"] = False
    render_kwargs_test["raw_noise_std"] = 0.

    if args.camera_model != "none":
        colmap_free = getattr(args, "run_without_colmap", "none") != "none"
        fx_init = W if colmap_free else noisy_focal
        fy_init = H if colmap_free else noisy_focal
        intrinsic_init = torch.tensor([[fx_init, 0, W / 2, 0], [0, fy_init, H / 2, 0],
                                       [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32)
        with torch.no_grad():
            camera_model = camera_dict[args.camera_model](
                intrinsics=intrinsic_init, extrinsics=noisy_poses, args=args, H=H, W=W).to(device)
        grad_vars += list(camera_model.parameters())

    if getattr(args, "use_custom_optim", False):        # create_nerf.py:126-130
        from .custom_optim import CustomAdamOptimizer
        optimizer = CustomAdamOptimizer(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999),
                                        weight_decay=args.non_linear_weight_decay, H=H, W=W, args=args)
    else:
        optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start = 0
    ckpts = []
    if getattr(args, "ft_path", None) not in (None, "None"):
        ckpts = [args.ft_path]
    else:
        d = os.path.join(args.basedir, args.expname)
        if os.path.isdir(d):
            ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if "tar" in f]
    print("Found ckpts", ckpts)
    if ckpts and not args.no_reload:
        print("Reloading from", ckpts[-1])
        ckpt = torch.load(ckpts[-1], map_location=device)
        start = ckpt["global_step"]
        optim_dict = optimizer.state_dict()
        optim_dict["state"].update(ckpt["optimizer_state_dict"]["state"])
        optimizer.load_state_dict(optim_dict)
        model.load_state_dict(ckpt["network_fn_state_dict"])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt["network_fine_state_dict"])
        if camera_model is not None and "camera_model" in ckpt:
            camera_model.load_state_dict(ckpt["camera_model"])

    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer, camera_model
