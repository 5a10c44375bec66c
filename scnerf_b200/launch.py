"""Run the UNMODIFIED reference trainers on the B200 hot path.

    python -m scnerf_b200.launch nerf   /path/to/SCNeRF -- --config configs/llff_data/fern.txt ...
    python -m scnerf_b200.launch nerfpp /path/to/SCNeRF -- --config configs/tanks_and_temples/tat_training_Truck.txt ...
    python -m scnerf_b200.launch nerf   /path/to/SCNeRF --check      # resolve names only, print a JSON report

What it does — and all it does:
  * puts the shim directories (``dropin/`` [+ ``dropin/nerfplusplus/``]) ahead of the trainer's own directory on
    ``sys.path`` (a plain ``python run_nerf.py`` puts the script's directory first, so ``PYTHONPATH`` alone cannot
    shadow ``render.py`` & co; ``PYTHONSAFEPATH=1`` + ``PYTHONPATH`` is the launcher-free alternative);
  * ``chdir`` into the trainer's directory (the trainers use ``".."``-relative paths) and imports the trainer module
    by name, so that ``torch.multiprocessing.spawn`` children (nerfplusplus/ddp_train_nerf.py:631) re-import the same
    reference file through the same path;
  * NeRF++: rebinds the samplers the trainer defines in its own module (ddp_train_nerf.py:50-132,135-256) —
    ``scnerf_b200.nerfplusplus.create_nerf.patch_trainer``; spawned children are patched when they call
    ``create_nerf`` (ddp_train_nerf.py:359);
  * calls the trainer's own ``train()``.
"""
import ast
import importlib
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(REPO, "dropin")

TRAINERS = {
    # kind: (sub-directory, module, shim dirs ahead of it, hot-path names that must resolve to this repo)
    "nerf": ("NeRF", "run_nerf", [DROPIN],
             ["render", "render_path", "create_nerf", "get_rays_full_image_no_camera",
              "get_rays_full_image_use_camera", "get_rays_kps_no_camera", "get_rays_kps_use_camera", "get_rays_np",
              "img2mse", "mse2psnr", "fix_seeds", "preprocess_match", "proj_ray_dist_loss_single",
              "PinholeModelRotNoiseLearning10kRayoRayd", "PinholeModelRotNoiseLearning10kRayoRaydDistortion"]),
    "nerfpp": ("nerfplusplus", "ddp_train_nerf", [os.path.join(DROPIN, "nerfplusplus"), DROPIN],
               ["create_nerf", "render_ray_from_camera", "intersect_sphere", "perturb_samples", "sample_pdf",
                "render_single_image", "preprocess_match", "proj_ray_dist_loss_single"]),
}


def setup_path(kind, ref_root):
    sub, mod, shims, _ = TRAINERS[kind]
    tdir = os.path.join(os.path.abspath(ref_root), sub)
    if not os.path.isfile(os.path.join(tdir, mod + ".py")):
        raise SystemExit(f"scnerf_b200.launch: {tdir}/{mod}.py not found")
    keep = [p for p in sys.path if os.path.abspath(p or ".") not in
            [os.path.abspath(x) for x in shims + [tdir]]]
    sys.path[:] = shims + [tdir, os.path.join(os.path.abspath(ref_root), "model"), REPO] + keep
    os.chdir(tdir)
    return tdir, mod


def import_trainer(kind, ref_root):
    tdir, mod = setup_path(kind, ref_root)
    sys.argv[0] = os.path.join(tdir, mod + ".py")
    trainer = importlib.import_module(mod)
    if kind == "nerfpp":
        from .nerfplusplus.create_nerf import patch_trainer
        patch_trainer(trainer)
    return trainer


def _owner(obj):
    return getattr(obj, "__module__", None) or type(obj).__module__


def _star_import_usage(trainer, path):
    """Names the trainer's code reads that only a ``from X import *`` can have provided: each must exist."""
    tree = ast.parse(open(path).read())
    stars = [n.module for n in ast.walk(tree) if isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names)]
    bound = set(dir(__builtins__)) if not isinstance(__builtins__, dict) else set(__builtins__)
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            bound.add(n.name)
            if isinstance(n, ast.FunctionDef):
                a = n.args
                bound.update(x.arg for x in a.args + a.kwonlyargs + a.posonlyargs)
                bound.update(x.arg for x in (a.vararg, a.kwarg) if x)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            bound.update((a.asname or a.name).split(".")[0] for a in n.names if a.name != "*")
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            bound.add(n.id)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            bound.add(n.name)
        elif isinstance(n, ast.arg):
            bound.add(n.arg)
    used = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
    need = sorted(used - bound)
    return stars, {name: (hasattr(trainer, name), _owner(getattr(trainer, name, None))) for name in need}


def check(kind, ref_root):
    """Import the reference trainer behind the shims and report where every hot-path name comes from."""
    trainer = import_trainer(kind, ref_root)
    names = TRAINERS[kind][3]
    report = {"trainer": trainer.__file__, "names": {}, "modules": {}}
    for n in names:
        obj = getattr(trainer, n, None)
        report["names"][n] = _owner(obj) if obj is not None else None
    for m in ("render", "get_rays", "create_nerf", "run_nerf_helpers", "camera_dict", "camera_model",
              "model.camera_model", "model.ray_dist_loss", "model.camera_dict", "ddp_model", "nerf_network",
              "custom_optim", "nerf_sample_ray_split", "model.reprojection", "data_loader_split"):
        if m in sys.modules:
            mod = sys.modules[m]
            report["modules"][m] = {"file": getattr(mod, "__file__", None),
                                    "impl": getattr(mod, "__scnerf_impl__", None),
                                    "wraps": getattr(mod, "__scnerf_wraps__", None)}
    stars, usage = _star_import_usage(trainer, trainer.__file__)
    report["star_imports"] = stars
    report["star_names"] = usage
    if kind == "nerfpp":        # the dataset sampler must reach the CUDA ray generator through its own globals
        nsr = sys.modules["nerf_sample_ray_split"]
        fn = nsr.RaySamplerSingleImage.random_sample.__globals__["render_ray_from_camera"]
        report["names"]["RaySamplerSingleImage.random_sample -> render_ray_from_camera"] = _owner(fn)
    return report


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2 or argv[0] not in TRAINERS:
        raise SystemExit(__doc__)
    kind, ref_root, rest = argv[0], argv[1], argv[2:]
    if rest and rest[0] == "--check":
        print(json.dumps(check(kind, ref_root), indent=1))
        return
    if rest and rest[0] == "--":
        rest = rest[1:]
    sys.argv = [sys.argv[0]] + rest
    trainer = import_trainer(kind, ref_root)
    sys.argv = [trainer.__file__] + rest
    if kind == "nerf":
        import torch
        torch.set_default_tensor_type('torch.cuda.FloatTensor')      # run_nerf.py:1126
        trainer.train()
    else:
        trainer.setup_logger()                                       # ddp_train_nerf.py:637-638
        trainer.train()


if __name__ == "__main__":
    main()
