"""equilib_b200 -- B200-native (sm_100a) implementation of equilib's remap hot path.

Same class / func API as ``equilib`` (equilib/__init__.py:3-21) for CUDA tensors:
per-pixel ray generation, ray-to-spherical mapping and nearest / bilinear / bicubic
sampling are fused into one hand-written CUDA kernel behind a C ABI
(include/equilib_b200.h).  No sampling grid is materialised, ``F.grid_sample`` is
never called and there is no CPU fallback.
"""

from .cube2equi import Cube2Equi, cube2equi
from .equi2cube import Equi2Cube, equi2cube
from .equi2cube2equi import Equi2Cube2Equi, equi2cube2equi
from .equi2equi import Equi2Equi, equi2equi
from .equi2pers import Equi2Pers, equi2pers, get_bounding_fov
from .grid_sample import grid_sample
from .pers2equi import Pers2Equi, pers2equi
from .static import StaticRemap

__version__ = "0.1.0"

__all__ = [
    "Cube2Equi",
    "Equi2Cube",
    "Equi2Equi",
    "Equi2Pers",
    "Pers2Equi",
    "cube2equi",
    "equi2cube",
    "equi2equi",
    "equi2pers",
    "pers2equi",
    "grid_sample",
    "get_bounding_fov",
    "StaticRemap",
    "Equi2Cube2Equi",
    "equi2cube2equi",
]
