#!/usr/bin/env python3
"""Aggregate pinned-host -> device copy rate of the box for 1, 2, 4, 8 GPUs at once (the ceiling of every host-fed
e2e number: the combine kernels need 8.4 bytes of input per voxel.exposure).  One thread per device, each copying its
own pinned 1 GiB buffer on its own stream; prints GB/s per device set.  python tools/h2d_ceiling.py"""
import json
import threading
import time

import torch


def run(devs, nbytes=1 << 30, reps=8):
    bufs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in devs]
    dst = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}") for d in devs]
    streams = [torch.cuda.Stream(device=d) for d in devs]
    barrier = threading.Barrier(len(devs) + 1)
    done = [0.0] * len(devs)

    def work(k):
        torch.cuda.set_device(devs[k])
        with torch.cuda.stream(streams[k]):
            dst[k].copy_(bufs[k], non_blocking=True)
        streams[k].synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        with torch.cuda.stream(streams[k]):
            for _ in range(reps):
                dst[k].copy_(bufs[k], non_blocking=True)
        streams[k].synchronize()
        done[k] = time.perf_counter() - t0

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(devs))]
    for t in th:
        t.start()
    barrier.wait()
    for t in th:
        t.join()
    per = [nbytes * reps / s / 1e9 for s in done]
    return {"devices": devs, "aggregate_gbs": nbytes * reps * len(devs) / max(done) / 1e9, "per_device_gbs": [round(x, 1) for x in per]}


def main():
    n = torch.cuda.device_count()
    sets = [[0]] + [list(range(k)) for k in (2, 4, 8) if k <= n]
    if n >= 4:
        sets += [[0, 2], [0, 4] if n >= 8 else [0, 3], [0, 2, 4, 6] if n >= 8 else [0, 1, 2, 3]]
    for s in sets:
        print(json.dumps(run(s)), flush=True)


if __name__ == "__main__":
    main()
