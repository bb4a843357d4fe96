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


# ---- NumPy restatement of csrc/synth_core.h (pins the generator) ----------------
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix64(z):
    z = z.copy()
    z ^= z >> np.uint64(30)
    z *= _M1
    z ^= z >> np.uint64(27)
    z *= _M2
    z ^= z >> np.uint64(31)
    return z


def _key(seed, exposure, voxels):
    with np.errstate(over="ignore"):
        k = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(exposure + 1) * np.uint64(0xD1B54A32D192ED03)
        return k + voxels.astype(np.uint64)


def synth_data_numpy(seed, exposure, v0, count, nx, ny):
    with np.errstate(over="ignore"):
        vox = np.arange(v0, v0 + count, dtype=np.int64)
        key = _key(seed, exposure, vox)
        z1 = _mix64(key)
        z2 = _mix64(key ^ np.uint64(0xA0761D6478BD642F))
        x = vox % nx
        y = (vox // nx) % ny
        edge = (x < 2) | (x >= nx - 2) | (y < (exposure % 3))
        hole = ((z2 >> np.uint64(32)) & np.uint64(0xFFFF)) < 3277
        m = np.uint64(0xFFFF)
        s = ((z1 & m) + ((z1 >> np.uint64(16)) & m) + ((z1 >> np.uint64(32)) & m) + ((z1 >> np.uint64(48)) & m)
             + (z2 & m) + ((z2 >> np.uint64(16)) & m)).astype(np.int64)
        g = (s - 196605).astype(np.float32) * np.float32(2.1579186e-5)
        base = np.float32(10.0) + np.float32(0.02) * np.float32((exposure % 7) - 3)
        v = (np.float32(base) + g).astype(np.float32)
        outl = ((z2 >> np.uint64(48)) & m) < 655
        v = np.where(outl, (v + np.float32(50.0)).astype(np.float32), v)
        v = np.where(edge | hole, np.float32(np.nan), v)
    return v.astype(np.float32)


def synth_stat_numpy(seed, exposure, v0, count):
    with np.errstate(over="ignore"):
        vox = np.arange(v0, v0 + count, dtype=np.int64)
        z3 = _mix64(_key(seed, exposure, vox) ^ np.uint64(0xE7037ED1A0B428DB))
        u = (z3 >> np.uint64(40)).astype(np.float32)
        return (np.float32(0.5) + (u * np.float32(8.9406967e-8)).astype(np.float32)).astype(np.float32)
