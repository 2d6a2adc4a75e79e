"""spimagine_b200 -- the volume-raycasting hot path of spimagine (max_project / iso_surface behind
VolumeRenderer), rebuilt as hand-written CUDA for NVIDIA B200 (sm_100a) behind a C ABI.

    from spimagine_b200 import VolumeRenderer
    from spimagine_b200.utils.transform_matrices import mat4_perspective, mat4_translate, mat4_rotation
"""
from .volumerender import VolumeRenderer  # noqa: F401
from ._lib import pinned_empty  # noqa: F401
from .utils.transform_matrices import *  # noqa: F401,F403

__version__ = "0.1.0"
