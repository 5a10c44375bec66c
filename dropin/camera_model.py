"""Drop-in shim for model/camera_model.py (imported as ``camera_model`` via model/ on sys.path)."""
from scnerf_b200.camera_model import *  # noqa: F401,F403
from scnerf_b200.camera_model import (CameraModel, PinholeModelRotNoiseLearning10kRayoRayd,  # noqa: F401
                                      PinholeModelRotNoiseLearning10kRayoRaydDistortion)
