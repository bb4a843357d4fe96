"""Calls into the CUDA library through its C ABI (test infrastructure).

`combine_cuda` / `median_cuda` take host NumPy exposures and run them either as a
HOST-location call (the library stages slabs itself) or as a DEVICE-location call
(inputs copied to HBM with torch first, outputs read back), returning the same
dict layout as oracle.combine_port / median_port.
"""
import ctypes

import numpy as np

from mpdaf_b200.ctools import DEVICE, HOST, check, ctools


def _table(addresses):
    arr = (ctypes.c_void_p * len(addresses))(*addresses)
    return arr, ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))


def _typ(vartype):
    return {"propagate": 0, "stat_mean": 1, "stat_one": 2}.get(vartype, vartype)


def combine_cuda(data, var=None, scales=None, offsets=None, nmax=2, nclip=5.0, nstop=2,
                 vartype="propagate", mad=False, location="device", device=0):
    n = len(data)
    data = [np.ascontiguousarray(a) for a in data]
    dt = data[0].dtype
    nvox = data[0].size
    var = None if var is None else [np.ascontiguousarray(a, dtype=dt) for a in var]
    lo, up = (nclip, nclip) if np.isscalar(nclip) else nclip
    scale = None if scales is None else np.ascontiguousarray(scales, dtype=np.float64)
    offset = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64)
    sel = np.zeros(n, dtype=np.intc)
    val = np.zeros(n, dtype=np.intc)
    sp = None if scale is None else scale.ctypes.data
    op = None if offset is None else offset.ctypes.data
    params = (int(nmax), float(lo), float(up), int(nstop), int(_typ(vartype)), int(mad))
    if location in ("host", "host_stacked"):
        out = np.empty(nvox)
        outvar = np.empty(nvox)
        expmap = np.empty(nvox, dtype=np.intc)
        if location == "host_stacked":     # one [N][V] array per HDU: equally spaced rows -> one 2-D copy per chunk
            sd = np.stack([a.reshape(-1) for a in data])
            sv = None if var is None else np.stack([a.reshape(-1) for a in var])
            data = list(sd)
            var = None if sv is None else list(sv)
        kd, dptr = _table([a.ctypes.data for a in data])
        kv, vptr = (None, None) if var is None else _table([a.ctypes.data for a in var])
        rc = ctools.mpdaf_b200_sigma_clipping_arrays(
            n, dptr, vptr, dt.itemsize, nvox, HOST, device, out.ctypes.data, outvar.ctypes.data,
            expmap.ctypes.data, sp, op, sel, val, *params, None)
        check(rc, "mpdaf_b200_sigma_clipping_arrays[host]")
    else:
        import torch
        dev = torch.device("cuda", device)
        tdt = torch.float32 if dt == np.float32 else torch.float64
        if location == "device_stacked":      # one [N][V] allocation: equally spaced rows
            sd = torch.from_numpy(np.stack([a.reshape(-1) for a in data])).to(dev)
            td = list(sd)
            tv = None if var is None else list(torch.from_numpy(np.stack([a.reshape(-1) for a in var])).to(dev))
        elif location == "device_pitched":    # one [N][pitch] allocation, rows padded to 128-byte lines (the pair kernel's layout)
            pitch = (nvox + 31) // 32 * 32
            sd = torch.full((n, pitch), float("nan"), dtype=tdt, device=dev)
            sd[:, :nvox] = torch.from_numpy(np.stack([a.reshape(-1) for a in data])).to(dev)
            td = [sd[i, :nvox] for i in range(n)]
            tv = None
            if var is not None:
                sv = torch.full((n, pitch), float("nan"), dtype=tdt, device=dev)
                sv[:, :nvox] = torch.from_numpy(np.stack([a.reshape(-1) for a in var])).to(dev)
                tv = [sv[i, :nvox] for i in range(n)]
        elif location == "device_scattered":  # rows at irregular distances inside one pool
            gaps = [0] + [32 * (1 + (7 * i) % 5) for i in range(n)]
            starts = [int(sum(gaps[: i + 1]) + i * nvox) for i in range(n)]
            pool_d = torch.full((starts[-1] + nvox + 32,), float("nan"), dtype=tdt, device=dev)
            pool_v = torch.full((starts[-1] + nvox + 32,), float("nan"), dtype=tdt, device=dev)
            td, tv = [], (None if var is None else [])
            for i in range(n):
                td.append(pool_d[starts[i]: starts[i] + nvox])
                td[-1].copy_(torch.from_numpy(data[i].reshape(-1)))
                if var is not None:
                    tv.append(pool_v[starts[i]: starts[i] + nvox])
                    tv[-1].copy_(torch.from_numpy(var[i].reshape(-1)))
        else:
            td = [torch.from_numpy(a.reshape(-1)).to(dev) for a in data]
            tv = None if var is None else [torch.from_numpy(a.reshape(-1)).to(dev) for a in var]
        tout = torch.empty(max(nvox, 1), dtype=torch.float64, device=dev)
        tvar = torch.empty(max(nvox, 1), dtype=torch.float64, device=dev)
        texp = torch.empty(max(nvox, 1), dtype=torch.int32, device=dev)
        kd, dptr = _table([t.data_ptr() if nvox else 1 for t in td])
        kv, vptr = (None, None) if tv is None else _table([t.data_ptr() if nvox else 1 for t in tv])
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = ctools.mpdaf_b200_sigma_clipping_arrays(
            n, dptr, vptr, dt.itemsize, nvox, DEVICE, device, tout.data_ptr(), tvar.data_ptr(),
            texp.data_ptr(), sp, op, sel, val, *params, stream)
        check(rc, "mpdaf_b200_sigma_clipping_arrays[device]")
        assert tdt is not None
        out = tout[:nvox].cpu().numpy()
        outvar = tvar[:nvox].cpu().numpy()
        expmap = texp[:nvox].cpu().numpy()
    return dict(data=out, var=outvar, expmap=expmap, select_pix=sel, valid_pix=val)


def median_cuda(data, location="device", device=0):
    n = len(data)
    data = [np.ascontiguousarray(a) for a in data]
    dt = data[0].dtype
    nvox = data[0].size
    val = np.zeros(n, dtype=np.intc)
    if location == "host":
        out = np.empty(nvox)
        expmap = np.empty(nvox, dtype=np.intc)
        kd, dptr = _table([a.ctypes.data for a in data])
        rc = ctools.mpdaf_b200_median_arrays(n, dptr, dt.itemsize, nvox, HOST, device, out.ctypes.data,
                                             expmap.ctypes.data, val, None)
        check(rc, "mpdaf_b200_median_arrays[host]")
    else:
        import torch
        dev = torch.device("cuda", device)
        td = [torch.from_numpy(a.reshape(-1)).to(dev) for a in data]
        tout = torch.empty(max(nvox, 1), dtype=torch.float64, device=dev)
        texp = torch.empty(max(nvox, 1), dtype=torch.int32, device=dev)
        kd, dptr = _table([t.data_ptr() if nvox else 1 for t in td])
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = ctools.mpdaf_b200_median_arrays(n, dptr, dt.itemsize, nvox, DEVICE, device, tout.data_ptr(),
                                             texp.data_ptr(), val, stream)
        check(rc, "mpdaf_b200_median_arrays[device]")
        out = tout[:nvox].cpu().numpy()
        expmap = texp[:nvox].cpu().numpy()
    return dict(data=out, expmap=expmap, valid_pix=val)


def kernel_name():
    return ctools.mpdaf_b200_last_kernel_name().decode()


def slow_voxels(reset=True, device=0):
    """Voxels the certified fast path handed to the exact fallback since the last reset."""
    return int(ctools.mpdaf_b200_slow_voxels(device, 1 if reset else 0))
