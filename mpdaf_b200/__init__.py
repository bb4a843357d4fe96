"""mpdaf_b200 -- B200-native replacement for MPDAF's cube-combination hot path.

`CubeList.combine` / `CubeList.median` (reference lib/mpdaf/obj/cubelist.py)
behind the same ctypes C ABI as the reference's `_ctools` (src/merging.c),
implemented as hand-written sm_100a CUDA kernels.  See DESIGN.md.

The classes are imported on first use, so that pure-Python helpers (`mpdaf_b200.fits`)
can be imported without loading the CUDA library.
"""
__all__ = ['CubeList', 'CubeMosaic', 'Cube', 'StatTable', 'DeviceStack']


def __getattr__(name):
    if name in __all__:
        from . import cubelist
        return getattr(cubelist, name)
    raise AttributeError(f"module 'mpdaf_b200' has no attribute {name!r}")
