"""Small workload that touches every kernel family once (compute-sanitizer target, see profiles/r02_sanitizer_*.log):
pair-packed kernel (pitched stacks, with the exact kernel for queued voxels), warp kernel (separate allocations, MAD,
partial depth buckets), sorted-runs kernel (N > 128), generic kernels (float64), median kernels, host slab pipeline."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, oracle, cuda_runner
from helpers import random_stack, assert_same_result
rng = np.random.default_rng(0)
for n in (5, 12, 48, 70, 100, 130):
    data, var = random_stack(rng, n, (3, 7, 13), nan_frac=0.2, outlier_frac=0.1, ties=(n == 12))
    for kw in (dict(), dict(scales=np.linspace(.9, 1.1, n), offsets=np.linspace(-.1, .1, n))):
        for vt in ('propagate', 'stat_mean'):
            want = oracle.combine_port(data, var, nclip=2.0, vartype=vt, **kw)
            for loc in ('device', 'device_pitched'):
                assert_same_result(cuda_runner.combine_cuda(data, var, nclip=2.0, vartype=vt, location=loc, **kw), want)
    assert_same_result(cuda_runner.combine_cuda(data, var, nclip=2.0, mad=True), oracle.combine_port(data, var, nclip=2.0, mad=True))
    assert_same_result(cuda_runner.median_cuda(data), oracle.median_port(data), median=True)
    assert_same_result(cuda_runner.combine_cuda(data, var, nclip=2.0, location='host'), oracle.combine_port(data, var, nclip=2.0))
data, var = random_stack(rng, 7, (3, 7, 13), dtype=np.float64)
assert_same_result(cuda_runner.combine_cuda(data, var, nclip=2.0), oracle.combine_port(data, var, nclip=2.0))
print('sanitizer workload ok')
