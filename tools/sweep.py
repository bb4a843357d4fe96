#!/usr/bin/env python3
"""Stack-depth sweep (BASELINE.json configs[4]): N = 4..256, all combine modes, one GPU.

For every depth the slab is sized to ~6 GB of float32 samples (throughput per voxel does not
depend on nz); prints one JSON line per (N, mode) with kernel time, Gvoxel.exposures/s and the
achieved fraction of the measured HBM peak.  Usage: python tools/sweep.py > profiles/r02_sweep.jsonl
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLANE = 326 * 331
DEPTHS = (4, 8, 12, 16, 24, 32, 48, 64, 96, 128, 256)
MODES = (("median", ["--median"]), ("propagate", ["--var", "propagate"]), ("propagate+scales", ["--var", "propagate", "--scaled"]),
         ("stat_mean", ["--var", "stat_mean"]), ("stat_one+scales", ["--var", "stat_one", "--scaled"]),
         ("mad+propagate", ["--var", "propagate", "--mad"]),
         ("f64 propagate+scales", ["--var", "propagate", "--scaled", "--f64"]))


def main():
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    quick = "--quick" in sys.argv
    for n in DEPTHS:
        for name, flags in MODES:
            if "f64" in name and n > 128:
                continue
            planes_per_voxel = n * (2 if "propagate" in name else 1) * (2 if "f64" in name else 1)
            # ~6 GB of samples per case and never fewer than 32 planes: 3.4 Mvoxel fill the 148 SMs many times over
            # (round 1 ran the deep rows on 2-7 planes: 0.15 ms kernels, not throughput measurements)
            nz = max(32, min(512, int(6e9 / (planes_per_voxel * 4 * PLANE))))
            if quick and n not in (12, 48, 96):
                continue
            cmd = [sys.executable, os.path.join(ROOT, "tools", "run_case.py"), "--n", str(n), "--nz", str(nz),
                   "--steps", "2", "--warmup", "1", "--check"] + flags
            res = subprocess.run(cmd, capture_output=True, text=True)
            line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
            try:
                r = json.loads(line)
            except Exception:
                r = {"error": (res.stderr or line)[-300:]}
            r.update(mode=name, n=n, nz=nz, frac_of_measured_peak=(r.get("gbs", 0) / peak))
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
