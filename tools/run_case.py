#!/usr/bin/env python3
"""Run one device-resident combine case a few times and print kernel time / GB/s.

Used for profiling (`ncu ... python tools/run_case.py ...`) and for the stack-depth sweep.
  python tools/run_case.py --n 48 --nz 32 --scaled --var propagate --steps 3
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NY, NX = 326, 331


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=48)
    ap.add_argument("--nz", type=int, default=32)
    ap.add_argument("--scaled", action="store_true")
    ap.add_argument("--median", action="store_true")
    ap.add_argument("--mad", action="store_true")
    ap.add_argument("--var", default="propagate")
    ap.add_argument("--nclip", type=float, default=5.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--f64", action="store_true", help="float64 samples (BITPIX -64)")
    ap.add_argument("--check", action="store_true", help="compare the first planes with the oracle port")
    args = ap.parse_args()

    import torch
    from mpdaf_b200 import synth
    from mpdaf_b200.ctools import DEVICE, check, ctools
    from mpdaf_b200.cubelist import _ptr_table

    n, nz = args.n, args.nz
    nvox = nz * NY * NX
    typ = {"propagate": 0, "stat_mean": 1, "stat_one": 2}[args.var]
    stack, tdata, tvar = synth.device_stack(n, (nz, NY, NX), seed=1, with_var=(typ == 0 and not args.median))
    dev = torch.device("cuda", 0)
    out = torch.empty(nvox, dtype=torch.float64, device=dev)
    outvar = torch.empty(nvox, dtype=torch.float64, device=dev)
    expmap = torch.empty(nvox, dtype=torch.int32, device=dev)
    scales = synth.default_scales(n) if args.scaled else np.ones(n)
    offsets = synth.default_offsets(n) if args.scaled else np.zeros(n)
    esz = 4
    dptrs, vptrs = stack.data_ptrs, stack.var_ptrs
    if args.f64:    # the same samples as float64, one [n][pitch] allocation per HDU
        esz = 8
        pitch = (nvox + 31) // 32 * 32
        d64 = torch.empty((n, pitch), dtype=torch.float64, device=dev)
        d64[:, :nvox] = tdata
        dptrs = [d64[i].data_ptr() for i in range(n)]
        if tvar is not None:
            v64 = torch.empty((n, pitch), dtype=torch.float64, device=dev)
            v64[:, :nvox] = tvar
            vptrs = [v64[i].data_ptr() for i in range(n)]
        del stack, tdata, tvar
    kd, dptr = _ptr_table(dptrs)
    kv, vptr = (None, None) if vptrs is None else _ptr_table(vptrs)
    stream = torch.cuda.current_stream(dev).cuda_stream
    times = []
    for it in range(args.warmup + args.steps):
        valid = np.zeros(n, dtype=np.intc)
        select = np.zeros(n, dtype=np.intc)
        if args.median:
            rc = ctools.mpdaf_b200_median_arrays(n, dptr, esz, nvox, DEVICE, 0, out.data_ptr(), expmap.data_ptr(),
                                                 valid, stream)
        else:
            rc = ctools.mpdaf_b200_sigma_clipping_arrays(
                n, dptr, vptr, esz, nvox, DEVICE, 0, out.data_ptr(), outvar.data_ptr(), expmap.data_ptr(),
                scales.ctypes.data, offsets.ctypes.data, select, valid, 2, args.nclip, args.nclip, 2, typ,
                int(args.mad), stream)
        check(rc, "combine")
        if it >= args.warmup:
            times.append(ctools.mpdaf_b200_last_kernel_ms())
    ms = float(np.mean(times))
    a = 2 if (typ == 0 and not args.median) else 1
    bpv = n * esz * a + (12 if args.median else 20)
    res = dict(kernel=ctools.mpdaf_b200_last_kernel_name().decode(), n=n, nz=nz, ms=ms,
               gvox_exp_s=n * nvox / ms / 1e6, gbs=bpv * nvox / ms / 1e6, bytes_per_voxel=bpv,
               rejected=int(valid.sum() - select.sum()) if not args.median else None,
               slow_voxels=int(ctools.mpdaf_b200_slow_voxels(0, 1)))
    if args.check:
        import oracle
        cz = min(nz, 2)
        data, var = synth.host_stack(n, (cz, NY, NX), seed=1, with_var=(typ == 0 and not args.median))
        cv = cz * NY * NX
        if args.median:
            want = oracle.median_port(data)
            ok = np.array_equal(np.nan_to_num(out[:cv].cpu().numpy()), np.nan_to_num(want["data"]))
        else:
            want = oracle.combine_port(data, var, scales=scales, offsets=offsets, nclip=args.nclip, vartype=args.var,
                                       mad=args.mad)
            g = out[:cv].cpu().numpy()
            ok = bool(np.array_equal(expmap[:cv].cpu().numpy(), want["expmap"])
                      and np.allclose(g, want["data"], rtol=1e-6, atol=0, equal_nan=True)
                      and np.allclose(outvar[:cv].cpu().numpy(), want["var"], rtol=1e-6, atol=0, equal_nan=True))
        res["parity_first_planes"] = ok
    print(json.dumps(res))


if __name__ == "__main__":
    main()
