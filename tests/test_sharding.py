"""Multi-rank host logic on CPU: world_size-2 gloo processes (no GPU involved).

Each rank combines its own wavelength slab with the CPU oracle standing in for the
kernel; counters are all-reduced and slabs gathered exactly as bench.py does over NCCL.
The sharded result must equal the unsharded one.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from mpdaf_b200 import sharding, synth


def test_slab_bounds_cover_the_cube():
    assert sharding.all_slabs(3681, 8) == [(0, 461)] + [(461 + 460 * i, 461 + 460 * (i + 1)) for i in range(7)]
    assert sharding.all_slabs(3681, 2) == [(0, 1841), (1841, 3681)]
    assert sharding.all_slabs(3681, 4)[0] == (0, 921)
    for nz in (1, 5, 40, 3681):
        for world in (1, 2, 3, 4, 8):
            slabs = sharding.all_slabs(nz, world)
            assert slabs[0][0] == 0 and slabs[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
            sizes = [b - a for a, b in slabs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.slab_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


SHAPE = (9, 6, 7)
N = 6


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nz, ny, nx = SHAPE
    z0, z1 = sharding.slab_bounds(nz, world, rank)
    data, var = synth.host_stack(N, (z1 - z0, ny, nx), seed=4, z0=z0)
    r = oracle.combine_port(data, var, nclip=2.0)
    valid, select = sharding.allreduce_counters(r["valid_pix"], r["select_pix"])
    full = sharding.gather_slabs(torch.from_numpy(r["data"]), nz, ny * nx)
    fullexp = sharding.gather_slabs(torch.from_numpy(r["expmap"]), nz, ny * nx)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), valid=valid, select=select, data=full.numpy(),
             expmap=fullexp.numpy())
    dist.destroy_process_group()


def test_two_rank_sharded_combine_equals_unsharded(tmp_path):
    oracle.build(ref=False)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    data, var = synth.host_stack(N, SHAPE, seed=4)
    want = oracle.combine_port(data, var, nclip=2.0)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        np.testing.assert_array_equal(got["valid"], want["valid_pix"])
        np.testing.assert_array_equal(got["select"], want["select_pix"])
        np.testing.assert_array_equal(got["expmap"], want["expmap"])
        np.testing.assert_array_equal(got["data"], want["data"])
