#!/usr/bin/env python3
"""Host-buffer (e2e) call of the combine path at several staging-chunk sizes, next to the raw H2D bandwidth
of this box.  python tools/e2e_probe.py [--nz 32]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NY, NX = 326, 331


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nz", type=int, default=32)
    ap.add_argument("--n", type=int, default=48)
    args = ap.parse_args()
    import torch
    from mpdaf_b200 import synth
    from mpdaf_b200.ctools import HOST, check, ctools
    from mpdaf_b200.cubelist import _ptr_table
    n, nvox = args.n, args.nz * NY * NX
    dev = torch.device("cuda", 0)
    stack, tdata, tvar = synth.device_stack(n, (args.nz, NY, NX), seed=1, with_var=True)
    hd = tdata.cpu().pin_memory()
    hv = tvar.cpu().pin_memory()
    # raw H2D: one big copy
    dst = torch.empty_like(tdata)
    for _ in range(2):
        dst.copy_(hd, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dst.copy_(hd, non_blocking=True)
    torch.cuda.synchronize()
    raw = 5 * hd.numel() * 4 / (time.perf_counter() - t0) / 1e9
    print(f"raw H2D (one {hd.numel() * 4 / 1e6:.0f} MB pinned copy): {raw:.1f} GB/s", flush=True)
    ho = torch.empty(nvox, dtype=torch.float64).pin_memory()
    hvv = torch.empty(nvox, dtype=torch.float64).pin_memory()
    he = torch.empty(nvox, dtype=torch.int32).pin_memory()
    k1, hdptr = _ptr_table([hd[i].data_ptr() for i in range(n)])
    k2, hvptr = _ptr_table([hv[i].data_ptr() for i in range(n)])
    scales, offsets = synth.default_scales(n), synth.default_offsets(n)
    for mib in (16, 32, 64, 128, 256, 512):
        os.environ["MPDAF_B200_SLAB_BYTES"] = str(mib << 20)

        def step():
            valid = np.zeros(n, dtype=np.intc)
            select = np.zeros(n, dtype=np.intc)
            check(ctools.mpdaf_b200_sigma_clipping_arrays(n, hdptr, hvptr, 4, nvox, HOST, 0, ho.data_ptr(), hvv.data_ptr(),
                                                          he.data_ptr(), scales.ctypes.data, offsets.ctypes.data, select, valid,
                                                          2, 5.0, 5.0, 2, 0, 0, None), "e2e")
        for _ in range(2):
            step()
        t0 = time.perf_counter()
        for _ in range(5):
            step()
        sec = (time.perf_counter() - t0) / 5
        print(f"slab {mib:4d} MiB: {sec * 1e3:7.2f} ms  {n * nvox / sec / 1e9:6.2f} Gvoxel.exp/s  H2D {2 * n * nvox * 4 / sec / 1e9:5.1f} GB/s", flush=True)


    # the same call on pageable NumPy buffers (what CubeList passes): inputs and results go through pinned staging
    os.environ["MPDAF_B200_SLAB_BYTES"] = str(128 << 20)
    pd, pv = hd.numpy().copy(), hv.numpy().copy()
    po, pvv, pe = np.empty(nvox), np.empty(nvox), np.empty(nvox, dtype=np.intc)
    k3, pdptr = _ptr_table([pd[i].ctypes.data for i in range(n)])
    k4, pvptr = _ptr_table([pv[i].ctypes.data for i in range(n)])

    def pstep():
        valid = np.zeros(n, dtype=np.intc)
        select = np.zeros(n, dtype=np.intc)
        check(ctools.mpdaf_b200_sigma_clipping_arrays(n, pdptr, pvptr, 4, nvox, HOST, 0, po.ctypes.data, pvv.ctypes.data,
                                                      pe.ctypes.data, scales.ctypes.data, offsets.ctypes.data, select, valid,
                                                      2, 5.0, 5.0, 2, 0, 0, None), "e2e pageable")
    for _ in range(2):
        pstep()
    t0 = time.perf_counter()
    for _ in range(5):
        pstep()
    sec = (time.perf_counter() - t0) / 5
    same = np.array_equal(pe, he.numpy()) and np.array_equal(np.nan_to_num(po), np.nan_to_num(ho.numpy()))
    print(f"pageable buffers: {sec * 1e3:7.2f} ms  {n * nvox / sec / 1e9:6.2f} Gvoxel.exp/s  same result as pinned: {same}", flush=True)


if __name__ == "__main__":
    main()
