#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference C code.

Run in the build container (needs /root/reference to compile oracle/_ref):

    OMP_NUM_THREADS=2 python tests/golden/make_golden.py

Every fixture holds the inputs (or the generator parameters), the call parameters
and the outputs of the reference's own OpenMP src/merging.c (through
oracle/standin/fitsio.h) for one case.  OMP_NUM_THREADS=2 keeps threads < planes
(the reference double-counts its counters otherwise).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
os.environ["OMP_NUM_THREADS"] = "2"

import oracle  # noqa: E402
from helpers import random_stack, write_exposures  # noqa: E402
from mpdaf_b200 import build, synth  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def run_ref(tmp, data, var, prefix, **kw):
    files = write_exposures(tmp, data, var, prefix=prefix)
    return oracle.combine_ref(files, data[0].shape, **kw)


def main():
    build.build_library()
    oracle.build()
    with tempfile.TemporaryDirectory() as tmp:
        # cfg1 of BASELINE.json: 5 synthetic 40x30x30 cubes with var, sigma-clip nclip=3 + median
        shape = (40, 30, 30)
        data, var = synth.host_stack(5, shape, seed=5)
        r = run_ref(tmp, data, var, "cfg1", nmax=2, nclip=3.0, nstop=2, vartype="propagate")
        files = write_exposures(tmp, data, var, prefix="cfg1m")
        m = oracle.median_ref(files, shape)
        save("cfg1", n=5, shape=shape, seed=5, nclip=3.0,
             data=r["data"], var=r["var"], expmap=r["expmap"].astype(np.int8),
             select_pix=r["select_pix"], valid_pix=r["valid_pix"],
             med_data=m["data"], med_expmap=m["expmap"].astype(np.int8),
             med_valid_pix=m["valid_pix"])

        # every var x mad mode, asymmetric clip, scales/offsets, edge voxels (inputs stored)
        rng = np.random.default_rng(2024)
        shape = (12, 7, 9)
        n = 9
        data, var = random_stack(rng, n, shape, nan_frac=0.15, outlier_frac=0.05)
        data[0][:, 0, :] = np.nan
        for d in data:
            d[3, 2, 1] = np.nan
        for d in data[1:]:
            d[4, 4, 4] = np.nan
        data[0][4, 4, 4] = 3.25
        scales = np.linspace(0.8, 1.3, n)
        offsets = np.linspace(-0.5, 0.5, n)
        out = dict(in_data=np.stack(data), in_var=np.stack(var), scales=scales, offsets=offsets,
                   nclip=np.array([2.0, 2.5]), nmax=2, nstop=2)
        for vt in ("propagate", "stat_mean", "stat_one"):
            for mad in (0, 1):
                r = run_ref(tmp, data, var, f"modes-{vt}-{mad}", scales=scales, offsets=offsets, nmax=2,
                            nclip=(2.0, 2.5), nstop=2, vartype=vt, mad=bool(mad))
                for k, v in r.items():
                    out[f"{vt}_{mad}_{k}"] = v
        save("modes", **out)

        # deep stack (cfg3-like): 48 exposures with per-exposure scales/offsets, nclip 5 and 3
        shape = (10, 6, 8)
        data, var = synth.host_stack(48, shape, seed=48)
        out = dict(n=48, shape=shape, seed=48)
        for nclip in (5.0, 3.0):
            r = run_ref(tmp, data, var, f"deep-{nclip}", scales=synth.default_scales(48),
                        offsets=synth.default_offsets(48), nmax=2, nclip=nclip, nstop=2, vartype="propagate")
            for k, v in r.items():
                out[f"nclip{int(nclip)}_{k}"] = v
        save("deep48", **out)

        # beyond the register kernels: 70 exposures, stat_mean
        shape = (10, 4, 5)
        data, var = synth.host_stack(70, shape, seed=70)
        r = run_ref(tmp, data, None, "n70", nmax=3, nclip=2.5, nstop=3, vartype="stat_mean")
        save("n70", n=70, shape=shape, seed=70, **r)

        # parameter sweep on one tie-heavy stack: nmax, nstop, asymmetric nclip
        rng = np.random.default_rng(99)
        shape = (10, 5, 6)
        data, var = random_stack(rng, 12, shape, nan_frac=0.2, outlier_frac=0.1, ties=True)
        out = dict(in_data=np.stack(data), in_var=np.stack(var))
        combos = [(0, 2.0, 2.0, 2), (1, 1.5, 3.0, 1), (3, 1.0, 1.0, 3), (2, 3.0, 1.5, 5), (5, 1.2, 1.2, 2)]
        out["combos"] = np.array(combos)
        for c, (nmax, lo, up, nstop) in enumerate(combos):
            r = run_ref(tmp, data, var, f"sweep-{c}", nmax=nmax, nclip=(lo, up), nstop=nstop, vartype="stat_one")
            for k, v in r.items():
                out[f"c{c}_{k}"] = v
        save("sweep_ties", **out)


if __name__ == "__main__":
    main()
