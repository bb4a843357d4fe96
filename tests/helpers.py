"""Shared test utilities (test infrastructure)."""
import os

import numpy as np

from mpdaf_b200 import fits as mfits

RTOL = 1e-6  # BASELINE.json: sigma-clipped means and variances within 1e-6 relative


def write_exposures(tmpdir, data, var=None, prefix="exp", dtype=np.float32, primary=None):
    """Write one FITS cube per exposure (DATA [+ STAT], float32 BE, NaN = masked)."""
    files = []
    for i, d in enumerate(data):
        path = os.path.join(str(tmpdir), f"{prefix}-{i:03d}.fits")
        hdr = dict(primary[i]) if primary else {}
        mfits.write_cube(path, d, None if var is None else var[i], primary_header=hdr, dtype=dtype)
        files.append(path)
    return files


def random_stack(rng, n, shape, nan_frac=0.1, outlier_frac=0.03, ties=False, with_var=True,
                 dtype=np.float32):
    """Random exposures with NaNs, outliers and (optionally) heavy ties."""
    data, var = [], []
    for i in range(n):
        if ties:
            d = rng.integers(0, 6, size=shape).astype(dtype)
        else:
            d = (10 + 0.1 * i + rng.standard_normal(shape)).astype(dtype)
        out = rng.random(shape) < outlier_frac
        d[out] += dtype(40.0) * rng.choice([-1, 1], size=out.sum()).astype(dtype)
        d[rng.random(shape) < nan_frac] = np.nan
        data.append(d)
        var.append(rng.uniform(0.5, 2.0, size=shape).astype(dtype))
    return data, (var if with_var else None)


def assert_same_result(got, want, median=False, rtol=RTOL, exact_data=False):
    """Parity bar of BASELINE.json: bit-exact expmap / counters / medians, NaN positions
    identical, <= 1e-6 relative on sigma-clipped means and variances."""
    np.testing.assert_array_equal(got["expmap"], want["expmap"])
    np.testing.assert_array_equal(got["valid_pix"], want["valid_pix"])
    if not median:
        np.testing.assert_array_equal(got["select_pix"], want["select_pix"])
    np.testing.assert_array_equal(np.isnan(got["data"]), np.isnan(want["data"]))
    if median or exact_data:
        np.testing.assert_array_equal(got["data"], want["data"])
    else:
        np.testing.assert_allclose(got["data"], want["data"], rtol=rtol, atol=0, equal_nan=True)
    if not median:
        np.testing.assert_array_equal(np.isnan(got["var"]), np.isnan(want["var"]))
        np.testing.assert_allclose(got["var"], want["var"], rtol=rtol, atol=0, equal_nan=True)


# ---- NumPy restatement of csrc/synth_core.h (pins the generator): oracle/synth_np.py ----
from oracle.synth_np import synth_data_numpy, synth_stat_numpy  # noqa: E402,F401
