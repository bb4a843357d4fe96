#!/usr/bin/env python3
"""bench.py -- Gvoxel.exposures/s of CubeList.combine (sigma clip) on B200.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
48 synthetic MUSE-WFM exposures of nz x 326 x 331 voxels (nz = 3681 when the stack
fits the GPU, else the largest nz that does), float32 DATA + STAT resident in HBM,
sigma clip nmax=2 nclip=5 nstop=2, var='propagate', per-exposure scales/offsets.

  python bench.py --gpus N --steps K --warmup W          (under torchrun for N > 1)
  python bench.py --impl reference ...                   (the reference's CPU path)

One step = one pass of the hot path over the rank's whole slab.  `value` times the
device-resident call with CUDA events; `e2e` times the same C-ABI call with HOST
(pinned) buffers, host<->device copies inside the timed region.  Multi-GPU: one
process per GPU, each rank combines its own wavelength slab (weak scaling, no
data-path collective); the per-exposure counters are summed over NCCL every step.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NY, NX = 326, 331
PLANE = NY * NX
NZ_FULL = 3681
NFILES = 48
CLIP = dict(nmax=2, nclip=5.0, nstop=2)


def algorithmic_bytes_per_voxel(n, propagate=True):
    """B = N * 4 * a + 20 (float32 samples, f64 data + f64 var + i32 expmap)."""
    return n * 4 * (2 if propagate else 1) + 20


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.lines = []
        exe = shutil.which("nvidia-smi")
        if exe:
            self.proc = subprocess.Popen(
                [exe, "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[2:6]):
                if val == "Active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------
# CPU side: the reference's own OpenMP merging.c (oracle/_ref) or the port
# ---------------------------------------------------------------------------------------
def cpu_sample_setup(nz, tmpdir):
    """Write the bounded sample as FITS files (float32 BE, DATA + STAT) for the reference."""
    from mpdaf_b200 import fits as mfits
    from mpdaf_b200 import synth
    files = []
    for i in range(NFILES):
        d = synth.fill_host(synth.DATA, 1, i, 0, nz * PLANE, NX, NY).reshape(nz, NY, NX)
        v = synth.fill_host(synth.STAT, 1, i, 0, nz * PLANE, NX, NY).reshape(nz, NY, NX)
        path = os.path.join(tmpdir, f"exp-{i:03d}.fits")
        mfits.write_cube(path, d, v)
        files.append(path)
    return files


def dropin_files_run(files, nz, scales, offsets, ref_result):
    """The drop-in symbol itself (`mpdaf_merging_sigma_clipping`, FITS file names in, NumPy buffers
    out -- exactly the call lib/mpdaf/obj/cubelist.py:512-515 makes) on the same FITS files the
    reference just combined: wall-clock time including FITS reading, H2D, kernel and D2H."""
    import ctypes
    from mpdaf_b200.ctools import ctools
    nvox = nz * PLANE
    out, outvar = np.empty(nvox), np.empty(nvox)
    expmap = np.empty(nvox, dtype=np.intc)
    names = "\n".join(files).encode("utf8")
    best = None
    for _ in range(3):
        sel, val = np.zeros(NFILES, dtype=np.intc), np.zeros(NFILES, dtype=np.intc)
        t0 = time.perf_counter()
        rc = ctools.mpdaf_merging_sigma_clipping(ctypes.c_char_p(names), out, outvar, expmap, scales, offsets, sel, val,
                                                 CLIP["nmax"], CLIP["nclip"], CLIP["nclip"], CLIP["nstop"], 0, 0)
        dt = time.perf_counter() - t0
        if rc != 0:
            return {"error": f"rc={rc}"}
        best = dt if best is None else min(best, dt)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = float(np.nanmax(np.abs(out - ref_result["data"]) / np.abs(ref_result["data"])))
    ok = bool(np.array_equal(expmap, ref_result["expmap"]) and np.array_equal(sel, ref_result["select_pix"])
              and np.array_equal(val, ref_result["valid_pix"]) and rel <= 1e-6)
    return {"seconds": best, "value": NFILES * nvox / best / 1e9, "unit": "Gvoxel*exposures/s",
            "parity_with_reference_ok": ok, "max_rel_err_data": rel,
            "what": "mpdaf_merging_sigma_clipping(file names) on the reference's own FITS inputs, FITS decode included"}


def cpu_reference_run(nz, steps, warmup, keep=False, dropin=False):
    """Times mpdaf_merging_sigma_clipping of the compiled reference on nz planes of the workload."""
    import oracle
    from mpdaf_b200 import synth
    oracle.build()
    scales, offsets = synth.default_scales(NFILES), synth.default_offsets(NFILES)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmpdir = tempfile.mkdtemp(prefix="mpdaf_b200_cpu_", dir=base)
    try:
        if oracle.have_ref():
            kind = "reference"
            files = cpu_sample_setup(nz, tmpdir)
            cores = oracle.ref_threads(NFILES, 0)

            def run():
                return oracle.combine_ref(files, (nz, NY, NX), scales=scales, offsets=offsets, vartype="propagate", **CLIP)
            sample = (f"unmodified src/merging.c (oracle/_ref, -O2 -fopenmp) on {NFILES} x ({nz},{NY},{NX}) float32 "
                      f"FITS cubes in {base or 'tmp'} via the stand-in fitsio.h (page-cache reads), "
                      f"effective OpenMP team {cores} of {os.cpu_count()} cores (merging.c:58-90)")
        else:
            kind = "port"
            data, var = synth.host_stack(NFILES, (nz, NY, NX), seed=1)
            cores = oracle.port().oracle_num_threads()

            def run():
                return oracle.combine_port(data, var, scales=scales, offsets=offsets, vartype="propagate", **CLIP)
            sample = f"oracle port (merge_oracle.c, OpenMP {cores} threads) on {NFILES} x ({nz},{NY},{NX}) float32 arrays"
        for _ in range(warmup):
            run()
        times, result = [], None
        for _ in range(steps):
            t0 = time.perf_counter()
            result = run()
            times.append(time.perf_counter() - t0)
        extra = None
        if dropin and kind == "reference":
            extra = dropin_files_run(files, nz, scales, offsets, result)
    finally:
        if not keep:
            shutil.rmtree(tmpdir, ignore_errors=True)
    sec = float(np.mean(times))
    value = NFILES * nz * PLANE / sec / 1e9
    return dict(value=value, unit="Gvoxel*exposures/s", cores=cores, kind=kind, sample=sample,
                seconds_per_pass=sec, dropin_files=extra), result


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    base, _ = cpu_reference_run(args.cpu_nz, args.steps, args.warmup)
    line = {
        "impl": "reference",
        "metric": "Gvoxel*exposures/s for CubeList.combine sigma-clip",
        "value": base["value"], "unit": "Gvoxel*exposures/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": base["seconds_per_pass"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.cpu_nz, 1, bounded_sample=True),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "Gvoxel*exposures/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(nz, world, bounded_sample=False):
    cfg = {
        "workload": f"MUSE WFM exposures: {NFILES} synthetic {nz}x{NY}x{NX} cubes per GPU, sigma-clip "
                    f"var='propagate' with per-exposure scales/offsets (BASELINE.json configs[2])",
        "exposures": NFILES, "shape": [nz, NY, NX], "nmax": CLIP["nmax"], "nclip": CLIP["nclip"],
        "nstop": CLIP["nstop"], "var": "propagate", "input_dtype": "float32",
        "sharding": f"wavelength slabs, {world} rank(s), {nz} planes each",
        "l2": "inputs larger than L2 (no flush needed)",
    }
    if bounded_sample:
        cfg["note"] = "bounded CPU sample of the same workload (throughput per voxel is nz-independent)"
    return cfg


# ---------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------
def pick_nz(torch, device, requested):
    if requested:
        return requested
    free, _ = torch.cuda.mem_get_info(device)
    per_plane = PLANE * (NFILES * 2 * 4 + 20)
    fit = int((free - (6 << 30)) // per_plane)
    return max(8, min(NZ_FULL, fit))


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    entry.build(quiet=True)
    from mpdaf_b200 import sharding, synth
    from mpdaf_b200.ctools import DEVICE, HOST, check, ctools
    from mpdaf_b200.cubelist import _ptr_table

    rank, world, local = dist_env()
    if args.gpus > 1 and world == 1:
        # not under torchrun: relaunch ourselves the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # (with NCCL_DEBUG=VERSION in the environment, as on the GPU boxes, NCCL prints its version banner on stdout
        # before the JSON line; the environment is left as the harness set it)
        dist.init_process_group("nccl", device_id=dev)

    nz = pick_nz(torch, dev, args.nz)
    if world > 1:
        t = torch.tensor([nz], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        nz = int(t.item())
    nvox = nz * PLANE
    z0 = rank * nz                      # weak scaling: rank r owns planes [r*nz, (r+1)*nz) of a deeper cube
    stack, tdata, tvar = synth.device_stack(NFILES, (nz, NY, NX), seed=1, with_var=True, z0=z0, device=local)
    out = torch.empty(nvox, dtype=torch.float64, device=dev)
    outvar = torch.empty(nvox, dtype=torch.float64, device=dev)
    expmap = torch.empty(nvox, dtype=torch.int32, device=dev)
    scales, offsets = synth.default_scales(NFILES), synth.default_offsets(NFILES)
    keep_d, dptr = _ptr_table(stack.data_ptrs)
    keep_v, vptr = _ptr_table(stack.var_ptrs)
    stream = torch.cuda.current_stream(dev)
    counters = torch.zeros(2 * NFILES, dtype=torch.int64, device=dev)
    kernel_ms, launches = [], [0]

    def step(timed=True):
        valid = np.zeros(NFILES, dtype=np.intc)
        select = np.zeros(NFILES, dtype=np.intc)
        rc = ctools.mpdaf_b200_sigma_clipping_arrays(
            NFILES, dptr, vptr, 4, nvox, DEVICE, local, out.data_ptr(), outvar.data_ptr(), expmap.data_ptr(),
            scales.ctypes.data, offsets.ctypes.data, select, valid, CLIP["nmax"], CLIP["nclip"], CLIP["nclip"],
            CLIP["nstop"], 0, 0, stream.cuda_stream)
        check(rc, "mpdaf_b200_sigma_clipping_arrays")
        if timed:
            kernel_ms.append(ctools.mpdaf_b200_last_kernel_ms())
            launches[0] += ctools.mpdaf_b200_last_launch_count()
        if world > 1:   # the path's only exchange: sum of the per-exposure counters over NVLink
            counters.copy_(torch.from_numpy(np.concatenate([valid, select]).astype(np.int64)))
            dist.all_reduce(counters)
        return valid, select

    for _ in range(args.warmup):
        step(timed=False)
    torch.cuda.synchronize(dev)
    ctools.mpdaf_b200_slow_voxels(local, 1)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        valid, select = step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    elapsed = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    kmean = torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        dist.all_reduce(kmean, op=dist.ReduceOp.MAX)
    ms_per_step = float(elapsed.item()) / args.steps
    value = NFILES * nvox * world / (ms_per_step * 1e-3) / 1e9

    # size-independent sanity of the full-size result (cheap, outside the timed region)
    n_sel = int(select.astype(np.int64).sum())
    assert n_sel == int(expmap.sum(dtype=torch.int64).item()), "sum(expmap) != sum(selected_pix)"
    for i in (0, NFILES // 2, NFILES - 1):
        assert int(valid[i]) == nvox - int(torch.isnan(tdata[i]).sum().item()), "valid_pix != #non-NaN samples"

    # ---- e2e: same C-ABI call with HOST (pinned) buffers, copies inside the timed region
    nz_e = min(nz, args.e2e_nz)
    nvox_e = nz_e * PLANE
    hd = tdata[:, :nvox_e].cpu().pin_memory()
    hv = tvar[:, :nvox_e].cpu().pin_memory()
    ho = torch.empty(nvox_e, dtype=torch.float64).pin_memory()
    hvv = torch.empty(nvox_e, dtype=torch.float64).pin_memory()
    he = torch.empty(nvox_e, dtype=torch.int32).pin_memory()
    keep_hd, hdptr = _ptr_table([hd[i].data_ptr() for i in range(NFILES)])
    keep_hv, hvptr = _ptr_table([hv[i].data_ptr() for i in range(NFILES)])

    def e2e_step():
        valid_e = np.zeros(NFILES, dtype=np.intc)
        select_e = np.zeros(NFILES, dtype=np.intc)
        rc = ctools.mpdaf_b200_sigma_clipping_arrays(
            NFILES, hdptr, hvptr, 4, nvox_e, HOST, local, ho.data_ptr(), hvv.data_ptr(), he.data_ptr(),
            scales.ctypes.data, offsets.ctypes.data, select_e, valid_e, CLIP["nmax"], CLIP["nclip"], CLIP["nclip"],
            CLIP["nstop"], 0, 0, None)
        check(rc, "mpdaf_b200_sigma_clipping_arrays[host]")
        return valid_e, select_e

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_sec = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_sec, op=dist.ReduceOp.MAX)
    e2e_value = NFILES * nvox_e * world / float(e2e_sec.item()) / 1e9
    # e2e result must equal the device-resident result on the same planes
    assert torch.equal(he, expmap[:nvox_e].cpu()), "e2e expmap differs from the device-resident run"
    assert torch.equal(torch.nan_to_num(ho), torch.nan_to_num(out[:nvox_e].cpu())), "e2e data differs"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N = 1 only) + parity of the same planes
    cpu_baseline, parity, dropin_files = None, None, None
    if world == 1 and not args.no_cpu:
        base, ref = cpu_reference_run(args.cpu_nz, 1, 0, dropin=True)
        cpu_baseline = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        dropin_files = base.get("dropin_files")
        nv = args.cpu_nz * PLANE
        g_exp = expmap[:nv].cpu().numpy()
        g_dat = out[:nv].cpu().numpy()
        g_var = outvar[:nv].cpu().numpy()
        ok_exp = bool(np.array_equal(g_exp, ref["expmap"]))
        ok_nan = bool(np.array_equal(np.isnan(g_dat), np.isnan(ref["data"])))
        with np.errstate(invalid="ignore", divide="ignore"):
            rel_d = np.nanmax(np.abs(g_dat - ref["data"]) / np.abs(ref["data"]))
            rel_v = np.nanmax(np.abs(g_var - ref["var"]) / np.abs(ref["var"]))
        parity = {"planes": args.cpu_nz, "expmap_bit_exact": ok_exp, "nan_positions_equal": ok_nan,
                  "max_rel_err_data": float(rel_d), "max_rel_err_var": float(rel_v),
                  "ok": bool(ok_exp and ok_nan and rel_d <= 1e-6 and rel_v <= 1e-6)}

    peak, peak_src = measured_peak()
    bpv = algorithmic_bytes_per_voxel(NFILES, True)
    kms = float(kmean.item())
    achieved = bpv * nvox / (kms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture
    # (dram__bytes_read.sum + dram__bytes_write.sum per voxel of the profiled launch, scaled
    # to this launch's voxel count); null until a capture exists.
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = float(json.load(f)["dram_bytes_per_voxel"]) * nvox
        except Exception:
            traffic = None
    line = {
        "metric": "Gvoxel*exposures/s for CubeList.combine sigma-clip",
        "value": value, "unit": "Gvoxel*exposures/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(nz, world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": ctools.mpdaf_b200_last_kernel_name().decode(),
                     "kernel_ms": kms, "algorithmic_bytes_per_voxel": bpv, "voxels_per_launch": nvox,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": "Gvoxel*exposures/s",
                "h2d_bytes_per_step": NFILES * 2 * 4 * nvox_e, "d2h_bytes_per_step": 20 * nvox_e,
                "planes": nz_e, "seconds_per_step": float(e2e_sec.item()),
                "api": "mpdaf_b200_sigma_clipping_arrays(location=HOST), pinned host buffers"},
        "gpu_launches": launches[0],
        "clocks": clocks,
        "parity": parity,
        "dropin_files": dropin_files,
        "rejected_fraction": float(1.0 - n_sel / max(1, int(valid.astype(np.int64).sum()))),
        "exact_fallback_voxels": int(ctools.mpdaf_b200_slow_voxels(local, 0)),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nz", type=int, default=0, help="planes per GPU (default: 3681 or what fits)")
    ap.add_argument("--e2e-nz", type=int, default=32, help="planes of the host-buffer e2e call")
    ap.add_argument("--cpu-nz", type=int, default=24, help="planes of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
