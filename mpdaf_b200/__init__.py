"""mpdaf_b200 -- B200-native replacement for MPDAF's cube-combination hot path.

`CubeList.combine` / `CubeList.median` (reference lib/mpdaf/obj/cubelist.py)
behind the same ctypes C ABI as the reference's `_ctools` (src/merging.c),
implemented as hand-written sm_100a CUDA kernels.  See DESIGN.md.
"""
from .cubelist import Cube, CubeList, CubeMosaic, DeviceStack, StatTable  # noqa: F401

__all__ = ['CubeList', 'CubeMosaic', 'Cube', 'StatTable', 'DeviceStack']
