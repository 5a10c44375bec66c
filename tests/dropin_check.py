"""Subprocess body of tests/test_dropin.py: import the UNMODIFIED reference trainer behind the shims and print
where every hot-path name resolves (scnerf_b200.launch.check).  Third-party packages the reference imports but
this image lacks (imageio, configargparse, piqa, …) are stubbed; nothing of the hot path is."""
import json
import os
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in ("imageio", "configargparse", "matplotlib", "matplotlib.backends", "matplotlib.backends.backend_agg",
             "matplotlib.figure", "matplotlib.cm", "matplotlib.pyplot", "piqa", "piqa.ssim", "piqa.lpips",
             "tensorboardX", "lpips", "torchvision", "torchvision.transforms", "kornia", "wandb", "cv2", "tqdm"):
    try:
        __import__(name)
    except Exception:
        sys.modules[name] = mock.MagicMock()
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
from scnerf_b200 import launch  # noqa: E402

print("REPORT" + json.dumps(launch.check(sys.argv[1], sys.argv[2])))
