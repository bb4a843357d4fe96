"""NumPy restatement of mpdaf_b200/csrc/synth_core.h (test infrastructure).

Pins the counter-based synthetic exposure generator (tests/test_abi.py) and feeds the reference arm of
bench.py, which must not load the CUDA library: a sample is a pure function of (seed, exposure, voxel).
"""
import numpy as np

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


def default_scales(n):
    return np.linspace(0.9, 1.1, n)


def default_offsets(n):
    return np.linspace(-0.1, 0.1, n)
