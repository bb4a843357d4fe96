#!/usr/bin/env python3
"""Wall-clock time of the drop-in file-name symbol and of the HOST array entry point (pageable / pinned) on one
bounded sample; used to tune the host side of the slab pipeline.  python tools/dropin_probe.py [planes]"""
import ctypes
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    nz = int(sys.argv[1]) if len(sys.argv) > 1 else 39
    cfg = bench.CONFIGS["cfg3"]
    tmp = tempfile.mkdtemp(prefix="probe_", dir="/dev/shm")
    try:
        files = bench.cpu_sample_setup(cfg, nz, tmp)
        from mpdaf_b200.ctools import ctools
        n, nvox = cfg["n"], nz * bench.PLANE
        scales, offsets = bench.scales_offsets(cfg)
        out, outvar, expmap = np.empty(nvox), np.empty(nvox), np.empty(nvox, dtype=np.intc)
        names = "\n".join(files).encode()
        for rep in range(4):
            sel, val = np.zeros(n, dtype=np.intc), np.zeros(n, dtype=np.intc)
            t0 = time.perf_counter()
            rc = ctools.mpdaf_merging_sigma_clipping(ctypes.c_char_p(names), out, outvar, expmap, scales, offsets, sel, val, 2, 5.0, 5.0, 2, 0, 0)
            dt = time.perf_counter() - t0
            print(f"files rep {rep}: rc={rc} {dt * 1e3:.1f} ms  {n * nvox / dt / 1e9:.2f} Gvoxel*exp/s  kernel {ctools.mpdaf_b200_last_kernel_ms():.2f} ms", flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
