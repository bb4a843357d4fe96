"""Wavelength-slab sharding of the combine path across the GPUs of one box.

Every voxel is independent and the reference itself parallelises over contiguous
wavelength-plane ranges (src/merging.c:112-137, compute_loop_limits), so each rank
(one process per GPU) combines its own contiguous slab of planes with no data-path
collective.  The only exchange is the sum of the per-exposure counters
(valid_pix / selected_pix, 2*N integers) and, when a caller wants the whole cube on
every rank, a gather of the output slabs.  Works with any torch.distributed backend
(NCCL on the GPUs, gloo in the CPU tests).
"""
import numpy as np


def slab_bounds(nz, world_size, rank):
    """Contiguous plane range [z0, z1) of `rank`: the first nz % world ranks get one more
    plane (3681 planes on 8 ranks -> 461 + 7 x 460)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    base, extra = divmod(nz, world_size)
    z0 = rank * base + min(rank, extra)
    return z0, z0 + base + (1 if rank < extra else 0)


def all_slabs(nz, world_size):
    return [slab_bounds(nz, world_size, r) for r in range(world_size)]


def allreduce_counters(valid_pix, select_pix=None, group=None, device=None):
    """Sum the per-exposure counters over all ranks (64-bit on the wire).

    Returns new int64 NumPy arrays; a no-op when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist

    valid = np.asarray(valid_pix, dtype=np.int64)
    select = None if select_pix is None else np.asarray(select_pix, dtype=np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return valid.copy(), None if select is None else select.copy()
    packed = np.concatenate([valid, select]) if select is not None else valid
    t = torch.from_numpy(packed.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    packed = t.cpu().numpy()
    n = valid.size
    return packed[:n].copy(), None if select is None else packed[n:].copy()


def gather_slabs(local, nz, plane_voxels, group=None):
    """All-gather flat per-rank output slabs (torch tensors, uneven last slabs allowed)
    into one tensor of nz * plane_voxels elements on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [(z1 - z0) * plane_voxels for z0, z1 in all_slabs(nz, world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[:local.numel()] = local
    parts = [torch.empty(pad, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])
