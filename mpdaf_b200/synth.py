"""Synthetic MUSE-like exposures, host and device, bit-identical.

Thin wrappers over `mpdaf_b200_synth_fill` (csrc/synth_core.h): a sample is a pure
function of (seed, exposure, global voxel index), so the oracle can regenerate on
the host exactly the slab the GPU combined.
"""
import numpy as np

from .ctools import DEVICE, HOST, check, ctools

DATA, STAT = 0, 1

# cfg3 of BASELINE.json: per-exposure scales / offsets
def default_scales(n):
    return np.linspace(0.9, 1.1, n)


def default_offsets(n):
    return np.linspace(-0.1, 0.1, n)


def fill_host(kind, seed, exposure, v0, count, nx, ny):
    """float32 array of `count` samples of `exposure` starting at global voxel v0."""
    out = np.empty(count, dtype=np.float32)
    rc = ctools.mpdaf_b200_synth_fill(out.ctypes.data, HOST, 0, kind, seed, exposure, v0, count, nx, ny, None)
    check(rc, 'mpdaf_b200_synth_fill')
    return out


def host_stack(n, shape, seed=1, with_var=True, z0=0):
    """n exposures of `shape` (nz, ny, nx) whose first plane is global plane z0."""
    nz, ny, nx = shape
    count = nz * ny * nx
    v0 = z0 * ny * nx
    data = [fill_host(DATA, seed, i, v0, count, nx, ny).reshape(shape) for i in range(n)]
    var = [fill_host(STAT, seed, i, v0, count, nx, ny).reshape(shape) for i in range(n)] if with_var else None
    return data, var


def fill_device(ptr, device, kind, seed, exposure, v0, count, nx, ny, stream=None):
    rc = ctools.mpdaf_b200_synth_fill(ptr, DEVICE, device, kind, seed, exposure, v0, count, nx, ny, stream)
    check(rc, 'mpdaf_b200_synth_fill')


def device_stack(n, shape, seed=1, with_var=True, z0=0, device=0):
    """Generate the stack directly in HBM; returns (DeviceStack, data tensor, var tensor)."""
    import torch

    from .cubelist import DeviceStack
    nz, ny, nx = shape
    count = nz * ny * nx
    v0 = z0 * ny * nx
    dev = torch.device('cuda', device)
    # HBM layout: one [n][pitch] allocation per HDU, rows padded to a multiple of 32 samples so that every row
    # starts on a 128-byte line (TMA needs 16-byte aligned rows; 3681 x 326 x 331 is not a multiple of 4)
    pitch = (count + 31) // 32 * 32
    data = torch.empty((n, pitch), dtype=torch.float32, device=dev)[:, :count]
    var = torch.empty((n, pitch), dtype=torch.float32, device=dev)[:, :count] if with_var else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    for i in range(n):
        fill_device(data[i].data_ptr(), device, DATA, seed, i, v0, count, nx, ny, stream)
        if with_var:
            fill_device(var[i].data_ptr(), device, STAT, seed, i, v0, count, nx, ny, stream)
    torch.cuda.synchronize(dev)
    stack = DeviceStack([data[i].data_ptr() for i in range(n)],
                        None if var is None else [var[i].data_ptr() for i in range(n)],
                        shape, 4, device, keepalive=(data, var))
    return stack, data, var
