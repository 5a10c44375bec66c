"""Helpers shared by the drop-in shims.

``reexport(globals, "scnerf_b200.x")``   — whole-module replacement: the shim's namespace becomes the
                                           implementation module's namespace.
``load_next(name, globals, __file__)``    — partial replacement: execute the REFERENCE's own module of the same
                                           name (the next one on sys.path after this shim) inside the shim's
                                           namespace, so everything that is not on the hot path (dataset
                                           samplers, loggers …) keeps working; the shim then rebinds the
                                           hot-path functions on top.  Functions of the reference module that
                                           call a rebound name (e.g. ``RaySamplerSingleImage.random_sample`` →
                                           ``render_ray_from_camera``) pick up the replacement, because they
                                           look the name up in this very namespace.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _is_shim_dir(p):
    p = os.path.abspath(p or ".")
    return p == HERE or os.path.dirname(p) == HERE


def reexport(ns, impl_name, leak=()):
    """``leak``: module names the reference's module imports at top level and thereby re-exports through
    ``from X import *`` — the trainers rely on some of them (run_nerf.py gets ``wandb`` and ``np`` that way)."""
    impl = importlib.import_module(impl_name)
    ns.update({k: v for k, v in vars(impl).items() if not k.startswith("__")})
    for spec in leak:
        name, _, alias = spec.partition(" as ")
        try:
            mod = importlib.import_module(name)
        except Exception:
            continue
        ns.setdefault(alias or name.split(".")[0], mod if alias or "." not in name else sys.modules[name.split(".")[0]])
    ns["__scnerf_impl__"] = impl_name
    return impl


def find_next(name, subdir=None):
    """Path of the reference's ``<name>.py`` (or ``<subdir>/<name>.py``): first hit on sys.path (plus the
    parent of the running script's directory, where the trainers themselves look: ``sys.path.insert(0, "..")``)
    that is not one of the shim directories."""
    rel = os.path.join(subdir, name + ".py") if subdir else name + ".py"
    cands = list(sys.path)
    main = getattr(sys.modules.get("__main__"), "__file__", None)
    for base in [os.path.dirname(os.path.abspath(main))] if main else []:
        cands += [base, os.path.dirname(base)]
    cands += [os.getcwd(), os.path.dirname(os.getcwd())]
    for p in cands:
        if _is_shim_dir(p):
            continue
        f = os.path.join(os.path.abspath(p or "."), rel)
        if os.path.isfile(f):
            return f
    return None


def load_next(name, ns, subdir=None):
    f = find_next(name, subdir)
    if f is None:
        raise ImportError(f"scnerf_b200 drop-in: cannot find the reference's {name}.py behind the shim "
                          "(put the reference trainer's directory on sys.path after dropin/)")
    with open(f) as fh:
        code = compile(fh.read(), f, "exec", dont_inherit=True)
    ns["__scnerf_wraps__"] = f
    exec(code, ns)
    return f


def _pin_model_package():
    """The trainers do ``sys.path.insert(0, "..")`` and then ``from model.camera_model import *`` /
    ``from model.ray_dist_loss import ...`` (NeRF/run_nerf.py:50,55-62): the reference root then precedes
    every other entry, so the only way the shim package ``dropin/model`` can answer is to be imported — and
    therefore cached in ``sys.modules`` — before that line runs.  Every shim imports this helper first."""
    if HERE not in [os.path.abspath(p or ".") for p in sys.path]:
        return
    # ``model`` (package) and the modules the loaders reach through ``sys.path.insert(0, "../model")``
    # (NeRF/load_llff.py:7, load_blender.py:11: ``from camera_model import ...``)
    for name in ("model", "camera_model", "camera_dict", "ray_dist_loss"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:  # no reference checkout around (unit tests of a single shim): nothing to pin
                pass


_pin_model_package()
