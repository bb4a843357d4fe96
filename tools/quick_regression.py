import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, oracle, cuda_runner
from helpers import random_stack, assert_same_result
rng=np.random.default_rng(0)
for n in (5,12,48,70,100):
    data,var=random_stack(rng,n,(3,7,13),nan_frac=0.2,outlier_frac=0.1)
    for kw in (dict(), dict(scales=np.linspace(.9,1.1,n), offsets=np.linspace(-.1,.1,n))):
        for vt in ('propagate','stat_mean'):
            got=cuda_runner.combine_cuda(data,var,nclip=2.0,vartype=vt,**kw)
            assert_same_result(got, oracle.combine_port(data,var,nclip=2.0,vartype=vt,**kw))
    assert_same_result(cuda_runner.median_cuda(data), oracle.median_port(data), median=True)
    got=cuda_runner.combine_cuda(data,var,nclip=2.0,location='host')
print('sanitizer workload ok')
