#!/usr/bin/env python3
"""bench.py -- Gvoxel.exposures/s of CubeList.combine / CubeList.median on B200.

Workloads (BASELINE.json `configs`; --config picks one, default cfg3 = the configuration the metric is quoted on):
  cfg2  12 synthetic 3681x326x331 exposures, CubeList.median
  cfg3  48 synthetic 3681x326x331 exposures, sigma clip nmax=2 nclip=5 nstop=2, var='propagate', scales/offsets
  cfg4  96 synthetic 3681x326x331 exposures, sigma clip, var='stat_mean'
float32 DATA (+ STAT) resident in HBM as one [N][pitch] allocation per HDU.

  python bench.py --gpus N --steps K --warmup W          (under torchrun for N > 1)
  python bench.py --impl reference ...                   (the reference's own CPU path, all host threads it can use)

One step = one pass of the hot path over the whole cube.  With N ranks the ONE cube is sharded along wavelength
(461 + 7 x 460 planes at N = 8, mpdaf_b200/sharding.py): strong scaling, no data-path collective; the per-exposure
counters are summed over NCCL every step.  `value` times the device-resident call with CUDA events (max over ranks);
`e2e` times the same C-ABI call with pinned HOST buffers, host<->device copies inside the timed region, and carries
the pageable-NumPy and file-name (drop-in symbol, FITS decode included) variants beside it.
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
CLIP = dict(nmax=2, nclip=5.0, nstop=2)
CONFIGS = {
    "cfg2": dict(n=12, median=True, var=None, scaled=False, index=1,
                 what="12 synthetic {nz}x326x331 cubes, CubeList.median"),
    "cfg3": dict(n=48, median=False, var="propagate", scaled=True, index=2,
                 what="48 synthetic {nz}x326x331 cubes, sigma-clip var='propagate' with per-exposure scales/offsets"),
    "cfg4": dict(n=96, median=False, var="stat_mean", scaled=False, index=3,
                 what="96 synthetic {nz}x326x331 cubes with var, sigma-clip var='stat_mean'"),
}
TYP = {"propagate": 0, "stat_mean": 1, "stat_one": 2, None: 1}
METRIC = {False: "Gvoxel*exposures/s for CubeList.combine sigma-clip", True: "Gvoxel*exposures/s for CubeList.median"}
DTYPE = "f64 statistics (float32 samples; bfloat16/float32 sort keys with certified decisions; float32 variance chains)"


def algorithmic_bytes_per_voxel(cfg):
    """B = N * 4 * a + o (float32 samples; f64 data [+ f64 var] + i32 expmap)."""
    a = 2 if cfg["var"] == "propagate" else 1
    return cfg["n"] * 4 * a + (12 if cfg["median"] else 20)


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


def workload_config(name, world):
    """Identical in both arms: it names the workload, not how much of it a bounded CPU sample covers."""
    cfg = CONFIGS[name]
    out = {
        "workload": "MUSE WFM exposures: " + cfg["what"].format(nz=NZ_FULL) + f" (BASELINE.json configs[{cfg['index']}])",
        "exposures": cfg["n"], "shape": [NZ_FULL, NY, NX], "input_dtype": "float32",
        "sharding": f"one cube, wavelength slabs over {world} rank(s)",
        "l2": "inputs larger than L2 (no flush needed)",
    }
    if not cfg["median"]:
        out.update(nmax=CLIP["nmax"], nclip=CLIP["nclip"], nstop=CLIP["nstop"], var=cfg["var"],
                   scales_offsets=bool(cfg["scaled"]))
    return out


def scales_offsets(cfg):
    n = cfg["n"]
    if cfg["scaled"]:
        return np.linspace(0.9, 1.1, n), np.linspace(-0.1, 0.1, n)
    return np.ones(n), np.zeros(n)


# ---------------------------------------------------------------------------------------
# CPU side: the reference's own OpenMP merging.c (oracle/_ref) or the port.  Nothing here loads the CUDA library.
# ---------------------------------------------------------------------------------------
def cpu_sample_setup(cfg, nz, tmpdir):
    """The bounded sample as FITS files (float32 BE, DATA [+ STAT]) from the NumPy generator."""
    from mpdaf_b200 import fits as mfits          # pure Python (the package imports its classes lazily)
    from oracle import synth_np
    files = []
    for i in range(cfg["n"]):
        d = synth_np.synth_data_numpy(1, i, 0, nz * PLANE, NX, NY).reshape(nz, NY, NX)
        v = synth_np.synth_stat_numpy(1, i, 0, nz * PLANE).reshape(nz, NY, NX) if cfg["var"] else None
        path = os.path.join(tmpdir, f"exp-{i:03d}.fits")
        mfits.write_cube(path, d, v)
        files.append(path)
    return files


def cpu_team(cfg):
    """(team, planes): the reference's effective OpenMP team on this host (merging.c:58-90) when given every core,
    and a plane count that loads every one of those threads (merging.c:122-126: nloops = nz / team + 1)."""
    import oracle
    cores = os.cpu_count() or 1
    team = oracle.ref_threads(cfg["n"], TYP[cfg["var"]] if not cfg["median"] else 1, omp_threads=cores)
    return team, max(team * 4 - 1, team + 1)


def cpu_reference_run(name, nz, steps, warmup, keep=False):
    """Times the compiled reference (or the port) on nz planes of the workload.  Returns (summary, result, files, tmpdir)."""
    import oracle
    cfg = CONFIGS[name]
    oracle.build()
    n = cfg["n"]
    scales, offsets = scales_offsets(cfg)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmpdir = tempfile.mkdtemp(prefix="mpdaf_b200_cpu_", dir=base)
    files = None
    try:
        team, _ = cpu_team(cfg)
        if oracle.have_ref():
            kind = "reference"
            # launchers such as torchrun export OMP_NUM_THREADS=1: give the reference every core explicitly
            oracle.set_ref_threads(os.cpu_count() or 1)
            files = cpu_sample_setup(cfg, nz, tmpdir)
            nloops = nz // team + 1 if team < nz else 1
            busy = min(team, (nz + nloops - 1) // nloops)

            def run():
                if cfg["median"]:
                    return oracle.median_ref(files, (nz, NY, NX))
                return oracle.combine_ref(files, (nz, NY, NX), scales=scales, offsets=offsets, vartype=cfg["var"], **CLIP)
            sample = (f"unmodified src/merging.c (oracle/_ref, -O2 -fopenmp) on {n} x ({nz},{NY},{NX}) float32 FITS cubes in "
                      f"{base or 'tmp'} via the stand-in fitsio.h (page-cache reads); OpenMP team {team} of {os.cpu_count()} "
                      f"cores after the reference's own cap (merging.c:58-90), {busy} of them with planes (merging.c:122-126)")
            cores = team
        else:
            kind = "port"
            from oracle import synth_np
            data = [synth_np.synth_data_numpy(1, i, 0, nz * PLANE, NX, NY).reshape(nz, NY, NX) for i in range(n)]
            var = [synth_np.synth_stat_numpy(1, i, 0, nz * PLANE).reshape(nz, NY, NX) for i in range(n)] if cfg["var"] else None
            cores = oracle.port().oracle_num_threads()

            def run():
                if cfg["median"]:
                    return oracle.median_port(data)
                return oracle.combine_port(data, var, scales=scales, offsets=offsets, vartype=cfg["var"], **CLIP)
            sample = f"oracle port (merge_oracle.c, OpenMP {cores} threads) on {n} x ({nz},{NY},{NX}) float32 arrays"
        for _ in range(warmup):
            run()
        times, result = [], None
        for _ in range(steps):
            t0 = time.perf_counter()
            result = run()
            times.append(time.perf_counter() - t0)
    except Exception:
        shutil.rmtree(tmpdir, ignore_errors=True)
        raise
    if not keep:
        shutil.rmtree(tmpdir, ignore_errors=True)
        files, tmpdir = None, None
    sec = float(np.mean(times))
    value = n * nz * PLANE / sec / 1e9
    return dict(value=value, unit="Gvoxel*exposures/s", cores=cores, kind=kind, sample=sample, seconds_per_pass=sec), result, files, tmpdir


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    nz = args.cpu_nz or cpu_team(cfg)[1]
    base, _, _, _ = cpu_reference_run(args.config, nz, args.steps, args.warmup)
    line = {
        "impl": "reference",
        "metric": METRIC[cfg["median"]],
        "value": base["value"], "unit": "Gvoxel*exposures/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": base["seconds_per_pass"] * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.config, args.gpus),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "Gvoxel*exposures/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "step": f"one pass of the reference over a bounded sample of {nz} planes (throughput per voxel is nz-independent)",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------
def pick_nz(torch, device, requested, cfg):
    if requested:
        return requested
    free, _ = torch.cuda.mem_get_info(device)
    planes = cfg["n"] * (2 if cfg["var"] == "propagate" else 1)
    per_plane = PLANE * (planes * 4 + 20)
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

    cfg = CONFIGS[args.config]
    n = cfg["n"]
    median = cfg["median"]
    typ = TYP[cfg["var"]]
    with_var = cfg["var"] == "propagate"
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
        dist.init_process_group("nccl", device_id=dev)

    # ONE cube of nz_total planes, sharded along wavelength over the ranks (strong scaling)
    nz_total = args.nz or NZ_FULL
    z0, z1 = sharding.slab_bounds(nz_total, world, rank)
    fit = pick_nz(torch, dev, 0, cfg)
    if z1 - z0 > fit:                                   # does not fit this GPU: shrink the cube for every rank alike
        nz_total = fit * world
        z0, z1 = sharding.slab_bounds(nz_total, world, rank)
    if world > 1:
        t = torch.tensor([nz_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) != nz_total:
            nz_total = int(t.item())
            z0, z1 = sharding.slab_bounds(nz_total, world, rank)
    nz = z1 - z0
    nvox = nz * PLANE
    stack, tdata, tvar = synth.device_stack(n, (nz, NY, NX), seed=1, with_var=with_var, z0=z0, device=local)
    out = torch.empty(nvox, dtype=torch.float64, device=dev)
    outvar = torch.empty(nvox, dtype=torch.float64, device=dev)
    expmap = torch.empty(nvox, dtype=torch.int32, device=dev)
    scales, offsets = scales_offsets(cfg)
    keep_d, dptr = _ptr_table(stack.data_ptrs)
    keep_v, vptr = _ptr_table(stack.var_ptrs) if with_var else (None, None)
    stream = torch.cuda.current_stream(dev)
    counters = torch.zeros(2 * n, dtype=torch.int64, device=dev)
    kernel_ms, launches = [], [0]

    def call(location, dp, vp, nv, o, ov, e, strm):
        valid = np.zeros(n, dtype=np.intc)
        select = np.zeros(n, dtype=np.intc)
        if median:
            rc = ctools.mpdaf_b200_median_arrays(n, dp, 4, nv, location, local, o, e, valid, strm)
        else:
            rc = ctools.mpdaf_b200_sigma_clipping_arrays(
                n, dp, vp, 4, nv, location, local, o, ov, e, scales.ctypes.data, offsets.ctypes.data, select, valid,
                CLIP["nmax"], CLIP["nclip"], CLIP["nclip"], CLIP["nstop"], typ, 0, strm)
        check(rc, "merging call")
        return valid, select

    def step(timed=True):
        valid, select = call(DEVICE, dptr, vptr, nvox, out.data_ptr(), outvar.data_ptr(), expmap.data_ptr(), stream.cuda_stream)
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
    total_vox = nz_total * PLANE
    value = n * total_vox / (ms_per_step * 1e-3) / 1e9

    # size-independent sanity of the full-size result (cheap, outside the timed region)
    if not median:
        n_sel = int(select.astype(np.int64).sum())
        assert n_sel == int(expmap.sum(dtype=torch.int64).item()), "sum(expmap) != sum(selected_pix)"
    else:
        n_sel = int(valid.astype(np.int64).sum())
        assert n_sel == int(expmap.sum(dtype=torch.int64).item()), "sum(expmap) != sum(valid_pix)"
    for i in (0, n // 2, n - 1):
        assert int(valid[i]) == nvox - int(torch.isnan(tdata[i]).sum().item()), "valid_pix != #non-NaN samples"

    # ---- e2e: same C-ABI call with HOST (pinned) buffers, copies inside the timed region
    nz_e = min(nz, args.e2e_nz)
    nvox_e = nz_e * PLANE
    hd = tdata[:, :nvox_e].cpu().pin_memory()
    hv = tvar[:, :nvox_e].cpu().pin_memory() if with_var else None
    ho = torch.empty(nvox_e, dtype=torch.float64).pin_memory()
    hvv = torch.empty(nvox_e, dtype=torch.float64).pin_memory()
    he = torch.empty(nvox_e, dtype=torch.int32).pin_memory()
    keep_hd, hdptr = _ptr_table([hd[i].data_ptr() for i in range(n)])
    keep_hv, hvptr = _ptr_table([hv[i].data_ptr() for i in range(n)]) if with_var else (None, None)

    def timed_host_calls(fn, steps):
        for _ in range(max(1, min(args.warmup, 3))):
            fn()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize(dev)
        sec = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return float(sec.item())

    e2e_steps = max(3, min(args.steps, 10))
    e2e_sec = timed_host_calls(lambda: call(HOST, hdptr, hvptr, nvox_e, ho.data_ptr(), hvv.data_ptr(), he.data_ptr(), None), e2e_steps)
    e2e_value = n * nvox_e * world / e2e_sec / 1e9
    # the e2e result must equal the device-resident result on the same planes
    assert torch.equal(he, expmap[:nvox_e].cpu()), "e2e expmap differs from the device-resident run"
    assert torch.equal(torch.nan_to_num(ho), torch.nan_to_num(out[:nvox_e].cpu())), "e2e data differs"
    h2d = n * (2 if with_var else 1) * 4 * nvox_e
    d2h = (12 if median else 20) * nvox_e
    e2e = {"value": e2e_value, "unit": "Gvoxel*exposures/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "planes_per_rank": nz_e, "seconds_per_step": e2e_sec,
           "api": "mpdaf_b200_sigma_clipping_arrays / _median_arrays (location=HOST), pinned host buffers, one call per rank"}

    # pageable NumPy buffers: what a caller of the array entry point holds without pinning anything (N = 1 only)
    if world == 1:
        pd = [hd[i].numpy().copy() for i in range(n)]
        pv = [hv[i].numpy().copy() for i in range(n)] if with_var else None
        po, pvv, pe = np.empty(nvox_e), np.empty(nvox_e), np.empty(nvox_e, dtype=np.intc)
        kpd, pdptr = _ptr_table([a.ctypes.data for a in pd])
        kpv, pvptr = _ptr_table([a.ctypes.data for a in pv]) if with_var else (None, None)
        sec_p = timed_host_calls(lambda: call(HOST, pdptr, pvptr, nvox_e, po.ctypes.data, pvv.ctypes.data, pe.ctypes.data, None), 3)
        e2e["pageable"] = {"value": n * nvox_e / sec_p / 1e9, "seconds_per_step": sec_p,
                           "api": "same call, pageable NumPy inputs and outputs (staged through pinned slots by the library)"}
        del pd, pv

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N = 1 only) + parity of the same planes + the drop-in symbol on its files
    cpu_baseline, parity, dropin = None, None, None
    if world == 1 and not args.no_cpu:
        cpu_nz = min(nz, args.cpu_nz or cpu_team(cfg)[1])
        base, ref, files, tmpdir = cpu_reference_run(args.config, cpu_nz, 1, 0, keep=True)
        try:
            cpu_baseline = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            nv = cpu_nz * PLANE
            g_exp = expmap[:nv].cpu().numpy()
            g_dat = out[:nv].cpu().numpy()
            ok_exp = bool(np.array_equal(g_exp, ref["expmap"]))
            ok_nan = bool(np.array_equal(np.isnan(g_dat), np.isnan(ref["data"])))
            with np.errstate(invalid="ignore", divide="ignore"):
                rel_d = float(np.nanmax(np.abs(g_dat - ref["data"]) / np.abs(ref["data"])))
            parity = {"planes": cpu_nz, "expmap_bit_exact": ok_exp, "nan_positions_equal": ok_nan, "max_rel_err_data": rel_d}
            ok = ok_exp and ok_nan and (rel_d == 0.0 if median else rel_d <= 1e-6)
            if not median:
                g_var = outvar[:nv].cpu().numpy()
                with np.errstate(invalid="ignore", divide="ignore"):
                    rel_v = float(np.nanmax(np.abs(g_var - ref["var"]) / np.abs(ref["var"])))
                parity["max_rel_err_var"] = rel_v
                ok = ok and rel_v <= 1e-6
            parity["ok"] = bool(ok)
            if files:
                dropin = dropin_files_run(cfg, files, cpu_nz, scales, offsets, ref)
        finally:
            if tmpdir:
                shutil.rmtree(tmpdir, ignore_errors=True)
    if dropin:
        e2e["dropin_files"] = dropin

    peak, peak_src = measured_peak()
    bpv = algorithmic_bytes_per_voxel(cfg)
    kms = float(kmean.item())
    achieved = bpv * nvox / (kms * 1e-3) / 1e9          # per GPU: the slowest rank's kernel time over its own slab
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                t = json.load(f)
            if t.get("config", "cfg3") == args.config:
                traffic = float(t["dram_bytes_per_voxel"]) * nvox
        except Exception:
            traffic = None
    line = {
        "metric": METRIC[median],
        "value": value, "unit": "Gvoxel*exposures/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args.config, world),
        "planes_on_device": {"total": nz_total, "this_rank": nz},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": ctools.mpdaf_b200_last_kernel_name().decode(),
                     "kernel_ms": kms, "algorithmic_bytes_per_voxel": bpv, "voxels_per_launch": nvox,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "cpu_baseline": cpu_baseline,
        "e2e": e2e,
        "gpu_launches": launches[0],
        "clocks": clocks,
        "parity": parity,
        "rejected_fraction": None if median else float(1.0 - n_sel / max(1, int(valid.astype(np.int64).sum()))),
        "exact_fallback_voxels": int(ctools.mpdaf_b200_slow_voxels(local, 0)),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dropin_files_run(cfg, files, nz, scales, offsets, ref_result):
    """The drop-in symbols themselves (`mpdaf_merging_sigma_clipping` / `mpdaf_merging_median`: FITS file names in, NumPy
    buffers out -- exactly the calls lib/mpdaf/obj/cubelist.py:456,512-515 make) on the FITS files the reference just
    combined: wall-clock time including FITS reading, H2D, kernel and D2H."""
    import ctypes
    from mpdaf_b200.ctools import ctools
    n = cfg["n"]
    nvox = nz * PLANE
    out, outvar = np.empty(nvox), np.empty(nvox)
    expmap = np.empty(nvox, dtype=np.intc)
    names = "\n".join(files).encode("utf8")
    best = None
    for _ in range(3):
        sel, val = np.zeros(n, dtype=np.intc), np.zeros(n, dtype=np.intc)
        t0 = time.perf_counter()
        if cfg["median"]:
            rc = ctools.mpdaf_merging_median(ctypes.c_char_p(names), out, expmap, val)
        else:
            rc = ctools.mpdaf_merging_sigma_clipping(ctypes.c_char_p(names), out, outvar, expmap, scales, offsets, sel, val,
                                                     CLIP["nmax"], CLIP["nclip"], CLIP["nclip"], CLIP["nstop"], TYP[cfg["var"]], 0)
        dt = time.perf_counter() - t0
        if rc != 0:
            return {"error": f"rc={rc}"}
        best = dt if best is None else min(best, dt)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = float(np.nanmax(np.abs(out - ref_result["data"]) / np.abs(ref_result["data"])))
    ok = bool(np.array_equal(expmap, ref_result["expmap"]) and np.array_equal(val, ref_result["valid_pix"]) and rel <= 1e-6)
    if not cfg["median"]:
        ok = ok and bool(np.array_equal(sel, ref_result["select_pix"]))
    return {"seconds": best, "value": n * nvox / best / 1e9, "unit": "Gvoxel*exposures/s", "planes": nz,
            "parity_with_reference_ok": ok, "max_rel_err_data": rel,
            "api": "the drop-in file-name symbol on the reference's own FITS inputs, FITS decode included"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--nz", type=int, default=0, help="planes of the whole cube (default: 3681 or what fits)")
    ap.add_argument("--e2e-nz", type=int, default=32, help="planes per rank of the host-buffer e2e call")
    ap.add_argument("--cpu-nz", type=int, default=0, help="planes of the bounded CPU sample (default: 4 x team - 1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
