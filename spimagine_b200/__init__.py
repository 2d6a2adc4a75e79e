"""spimagine_b200 -- the volume-raycasting hot path of spimagine (max_project / iso_surface behind
VolumeRenderer), rebuilt as hand-written CUDA for NVIDIA B200 (sm_100a) behind a C ABI.

    from spimagine_b200 import VolumeRenderer
    from spimagine_b200.utils.transform_matrices import mat4_perspective, mat4_translate, mat4_rotation
"""
from .volumerender import VolumeRenderer  # noqa: F401
from ._lib import pinned_empty  # noqa: F401
from .utils.transform_matrices import *  # noqa: F401,F403
# the names spimagine/__init__.py:20-27 exports that have a counterpart here (the GUI entry points do not)
from .frames import (DataModel, DemoData, SpimData, TiffData, TiffFolderData, NumpyData, RawData,  # noqa: F401
                     RawMultipleFiles, XwingData, GenericData)
from .frames import OverlayData, CZIData  # noqa: F401  (spimagine/__init__.py:20-21)
from .keyframes import TransformData  # noqa: F401
from .transform_model import TransformModel  # noqa: F401
from .utils.quaternion import Quaternion  # noqa: F401
from .utils.tiffio import read3dTiff, write3dTiff  # noqa: F401

__version__ = "0.1.0"
