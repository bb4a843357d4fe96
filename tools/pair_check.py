#!/usr/bin/env python3
"""Development check of the pair kernel (sigclip_pair.cuh) against the oracle port on random stacks.

  python tools/pair_check.py            # parity sweep, prints one line per case and a summary
Not a test (tests/test_gpu_parity.py holds those); used on the GPU box while iterating on the kernel.
"""
import itertools
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle
    import cuda_runner as cr
    from helpers import assert_same_result, random_stack
    oracle.build()
    bad = 0
    cases = 0
    for n, ties in itertools.product([2, 3, 4, 5, 12, 24, 33, 48, 64], [False, True]):
        rng = np.random.default_rng(7 * n + ties)
        shape = (5, 13, 17) if n > 24 else (9, 13, 17)
        data, var = random_stack(rng, n, shape, nan_frac=0.12, outlier_frac=0.04, ties=ties)
        scales = np.linspace(0.9, 1.1, n)
        offsets = np.linspace(-0.1, 0.1, n)
        for kw in (dict(), dict(scales=scales, offsets=offsets)):
            for vartype, nclip, nmax in (("propagate", 3.0, 2), ("stat_mean", (2.0, 3.0), 2), ("stat_one", 5.0, 2),
                                         ("propagate", 1.5, 5), ("stat_mean", 5.0, 0)):
                cases += 1
                try:
                    want = oracle.combine_port(data, var, vartype=vartype, nclip=nclip, nmax=nmax, **kw)
                    got = cr.combine_cuda(data, var, vartype=vartype, nclip=nclip, nmax=nmax, location="device_pitched", **kw)
                    name = cr.kernel_name()
                    slow = cr.slow_voxels()
                    assert "pair" in name, name
                    assert_same_result(got, want)
                    print(f"ok   n={n:3d} ties={int(ties)} {'scaled' if kw else 'unit  '} {vartype:9s} nclip={nclip} nmax={nmax} "
                          f"slow={slow}/{data[0].size} rejected={int(want['valid_pix'].sum() - want['select_pix'].sum())}")
                except Exception as e:   # noqa: BLE001
                    bad += 1
                    msg = str(e).strip().splitlines()
                    print(f"FAIL n={n:3d} ties={int(ties)} {'scaled' if kw else 'unit  '} {vartype:9s} nclip={nclip} nmax={nmax}: "
                          f"{type(e).__name__} {' | '.join(msg[:6])}")
                    if not isinstance(e, AssertionError):
                        traceback.print_exc()
    print(f"{cases - bad}/{cases} cases ok")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
