"""cylindertag_b200 -- B200-native (sm_100a) implementation of CylinderTag's per-frame detection front end.

The package holds the CUDA kernels + C ABI (csrc/, include/ctag.h) and the host-side mirror of the reference's
`CylinderTag` interface.  There is no CPU implementation in here: importing works anywhere, but every detection call
needs the built library and a B200.
"""
from . import _capi
from ._capi import CtagError
from .detector import Detector, detect_batch_multi
from .api import CamInfo, CylinderTag, MarkerInfo, ModelInfo, PoseInfo

__all__ = ["CylinderTag", "Detector", "detect_batch_multi", "MarkerInfo", "ModelInfo", "CamInfo", "PoseInfo", "CtagError", "_capi"]
__version__ = "0.1"
