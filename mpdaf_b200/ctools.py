"""ctypes loader for the B200 `_ctools` library.

Mirrors the reference loader lib/mpdaf/tools/ctools.py:52-92: the shared object
`_ctools` is loaded from this package's directory with
numpy.ctypeslib.load_library and the two merging symbols get exactly the
reference's argtypes, so `ctools.mpdaf_merging_median(...)` /
`ctools.mpdaf_merging_sigma_clipping(...)` are called the same way as in
lib/mpdaf/obj/cubelist.py:456 and :512-515.  As in the reference, a failed load
logs an error and leaves the module without a `ctools` attribute
(ctools.py:57-60).  The array entry points and helpers of
include/mpdaf_b200.h are declared as well.
"""
import ctypes
import logging
import os

from numpy.ctypeslib import load_library, ndpointer

LIBRARY_PATH = os.path.dirname(os.path.abspath(__file__))

# element types / locations / return codes (include/mpdaf_b200.h)
F32, F64 = 4, 8
HOST, DEVICE = 0, 1
ALL_DEVICES = -1   # `device` of a HOST-location call: spread the wavelength slabs over the configured devices
OK, ERR_ARG, ERR_CUDA, ERR_IO, ERR_SHAPE = 0, 1, 2, 3, 4

try:
    # MPDAF_B200_LIBNAME selects an alternative build of the library (tuning experiments)
    ctools = load_library(os.environ.get("MPDAF_B200_LIBNAME", "_ctools"), LIBRARY_PATH)
except OSError:  # pragma: no cover
    logging.getLogger(__name__).error(
        "mpdaf_b200's CUDA extension `_ctools` is missing; build it with "
        "`python -m mpdaf_b200.build` (needs nvcc). There is no CPU fallback.")
else:
    charptr = ctypes.POINTER(ctypes.c_char)
    array_1d_double = ndpointer(dtype=ctypes.c_double, ndim=1, flags='C_CONTIGUOUS')
    array_1d_int = ndpointer(dtype=ctypes.c_int, ndim=1, flags='C_CONTIGUOUS')
    c_void_pp = ctypes.POINTER(ctypes.c_void_p)

    # ---- the two symbols the reference binds (ctools.py:69-92) ----
    ctools.mpdaf_merging_median.restype = ctypes.c_int
    ctools.mpdaf_merging_median.argtypes = [
        charptr,          # char* input
        array_1d_double,  # double* data
        array_1d_int,     # int* expmap
        array_1d_int      # int* valid_pix
    ]
    ctools.mpdaf_merging_sigma_clipping.restype = ctypes.c_int
    ctools.mpdaf_merging_sigma_clipping.argtypes = [
        charptr,          # char* input
        array_1d_double,  # double* data
        array_1d_double,  # double* var
        array_1d_int,     # int* expmap
        array_1d_double,  # double* scale
        array_1d_double,  # double* offset
        array_1d_int,     # int* selected_pix
        array_1d_int,     # int* valid_pix
        ctypes.c_int,     # int nmax
        ctypes.c_double,  # double nclip_low
        ctypes.c_double,  # double nclip_up
        ctypes.c_int,     # int nstop
        ctypes.c_int,     # int typ_var
        ctypes.c_int      # int mad
    ]

    # ---- array entry points: raw addresses so host and device buffers both fit ----
    ctools.mpdaf_b200_sigma_clipping_arrays.restype = ctypes.c_int
    ctools.mpdaf_b200_sigma_clipping_arrays.argtypes = [
        ctypes.c_int, c_void_pp, c_void_pp, ctypes.c_int, ctypes.c_longlong,
        ctypes.c_int, ctypes.c_int,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,   # out, outvar, expmap
        ctypes.c_void_p, ctypes.c_void_p,                    # scale, offset (host)
        array_1d_int, array_1d_int,                          # selected_pix, valid_pix
        ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
    ]
    ctools.mpdaf_b200_median_arrays.restype = ctypes.c_int
    ctools.mpdaf_b200_median_arrays.argtypes = [
        ctypes.c_int, c_void_pp, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
        ctypes.c_void_p, ctypes.c_void_p, array_1d_int, ctypes.c_void_p,
    ]
    ctools.mpdaf_b200_last_kernel_ms.restype = ctypes.c_double
    ctools.mpdaf_b200_last_kernel_ms.argtypes = []
    ctools.mpdaf_b200_last_launch_count.restype = ctypes.c_int
    ctools.mpdaf_b200_last_launch_count.argtypes = []
    ctools.mpdaf_b200_last_kernel_name.restype = ctypes.c_char_p
    ctools.mpdaf_b200_last_kernel_name.argtypes = []
    ctools.mpdaf_b200_last_error.restype = ctypes.c_char_p
    ctools.mpdaf_b200_last_error.argtypes = []
    ctools.mpdaf_b200_last_expmap_histogram.restype = ctypes.c_int
    ctools.mpdaf_b200_last_expmap_histogram.argtypes = [ndpointer(dtype=ctypes.c_longlong, ndim=1, flags='C_CONTIGUOUS'),
                                                        ctypes.c_int]
    ctools.mpdaf_b200_slow_voxels.restype = ctypes.c_longlong
    ctools.mpdaf_b200_slow_voxels.argtypes = [ctypes.c_int, ctypes.c_int]
    ctools.mpdaf_b200_device_count.restype = ctypes.c_int
    ctools.mpdaf_b200_device_count.argtypes = []
    ctools.mpdaf_b200_set_devices.restype = ctypes.c_int
    ctools.mpdaf_b200_set_devices.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    ctools.mpdaf_b200_release.restype = ctypes.c_int
    ctools.mpdaf_b200_release.argtypes = []
    ctools.mpdaf_b200_finite_mask.restype = ctypes.c_int
    ctools.mpdaf_b200_finite_mask.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.POINTER(ctypes.c_longlong), ctypes.c_void_p]
    ctools.mpdaf_b200_fits_image_info.restype = ctypes.c_int
    ctools.mpdaf_b200_fits_image_info.argtypes = [
        ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_long * 3),
        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong),
    ]
    ctools.mpdaf_b200_synth_fill.restype = ctypes.c_int
    ctools.mpdaf_b200_synth_fill.argtypes = [
        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong,
        ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
        ctypes.c_void_p,
    ]


class CtoolsError(RuntimeError):
    """A merging entry point returned a non-zero MPDAF_B200_ERR_* code."""


def check(rc, what):
    if rc != OK:
        msg = ctools.mpdaf_b200_last_error().decode("utf8", "replace")
        raise CtoolsError(f"{what} failed (code {rc}): {msg}")
