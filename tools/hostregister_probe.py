#!/usr/bin/env python3
"""How fast can pageable host memory be pinned in place (cudaHostRegister)?  The alternative to gathering pageable inputs
into the library's pinned staging slots would be registering the caller's arrays (or the mapped FITS pages) for every
call; this probe times registration + unregistration of a touched 1 GiB NumPy array against the staging copy it would
replace.  python tools/hostregister_probe.py"""
import json
import time

import numpy as np
import torch


def main():
    torch.cuda.init()
    rt = torch.cuda.cudart()
    nbytes = 1 << 30
    res = {}
    for trial in range(3):
        a = np.ones(nbytes // 4, dtype=np.float32)      # touched: pages exist
        t0 = time.perf_counter()
        rc = rt.cudaHostRegister(a.ctypes.data, nbytes, 0)
        t1 = time.perf_counter()
        rt.cudaHostUnregister(a.ctypes.data)
        t2 = time.perf_counter()
        res[f"register_GBps_{trial}"] = nbytes / (t1 - t0) / 1e9
        res[f"unregister_GBps_{trial}"] = nbytes / (t2 - t1) / 1e9
        res["rc"] = int(rc) if not hasattr(rc, "value") else int(rc.value)
    pinned = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    src = np.ones(nbytes // 4, dtype=np.float32)
    dst = pinned.numpy()
    t0 = time.perf_counter()
    np.copyto(dst, src)
    t1 = time.perf_counter()
    res["single_thread_memcpy_GBps"] = nbytes / (t1 - t0) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
