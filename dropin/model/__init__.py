"""Drop-in shim for the reference's ``model`` package (``from model.camera_model import *``,
``from model.ray_dist_loss import ...`` in NeRF/run_nerf.py:55-62 and nerfplusplus/ddp_train_nerf.py:21-25).

``camera_model``, ``camera_dict``, ``camera_utils`` and ``ray_dist_loss`` come from scnerf_b200; every other
submodule (``reprojection``, ``lookup``, ``prd_evaluation`` … — matching / evaluation code, out of scope) is
still found in the reference's own ``model/`` directory, which is appended to this package's search path."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_main = getattr(_sys.modules.get("__main__"), "__file__", None)
_cands = list(_sys.path) + [_os.getcwd(), _os.path.dirname(_os.getcwd())]
if _main:
    _d = _os.path.dirname(_os.path.abspath(_main))
    _cands += [_d, _os.path.dirname(_d)]
for _p in _cands:
    _m = _os.path.join(_os.path.abspath(_p or "."), "model")
    if _m != _here and _os.path.isfile(_os.path.join(_m, "reprojection.py")) and _m not in __path__:
        __path__.append(_m)
