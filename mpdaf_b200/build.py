"""Builds mpdaf_b200/_ctools.so (the CUDA replacement for MPDAF's `_ctools`).

In-tree build with nvcc for sm_100a only; the .so is git-ignored but travels to
the GPU box with the repository snapshot.  `python -m mpdaf_b200.build` or
`__graft_entry__.build()`.  The kernels are split into one translation unit per
stack-depth bucket (warp_<NB>.cu) so that they compile in parallel.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "_ctools.so")
DEPTHS = (4, 8, 12, 16, 24, 32, 48, 64, 96, 128)
SOURCES = ("ctools_abi.cu", "fits_reader.cpp", "host_tools.cpp") + tuple(f"warp_{n}.cu" for n in DEPTHS) + ("warp_runs.cu", "pair64_12.cu", "pair64_24.cu", "pair64_48.cu", "pair64_64.cu")
HEADERS = ("combine_kernels.cuh", "stack_common.cuh", "sigclip_warp.cuh", "sigclip_pair.cuh", "sigclip_duo.cuh", "median_pipe.cuh", "sigclip_runs.cuh", "warp_inst.cuh", "launch.h",
           "sort_networks.cuh", "synth_core.h", "fits_reader.h", os.path.join("..", "..", "include", "mpdaf_b200.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                       # no FMA contraction in the float64 statistics
    "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _mtime(path):
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _newest_header():
    return max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)


def is_stale():
    t = _mtime(LIB)
    return t == 0.0 or any(_mtime(os.path.join(CSRC, s)) > t for s in SOURCES) or _newest_header() > t


def build_library(force=False, verbose=False, extra_flags=(), lib=LIB):
    """Compile what is out of date and link the library.  Returns its path."""
    if not force and not extra_flags and lib == LIB and not is_stale():
        return lib
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = _newest_header()
    tag = "".join(c if c.isalnum() else "_" for c in "".join(extra_flags))
    jobs = []
    for src in SOURCES:
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + tag + ".o")
        if force or _mtime(obj) < max(_mtime(os.path.join(CSRC, src)), hdr_t):
            cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        return src, res

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as pool:
        for src, res in pool.map(run, jobs):
            if verbose or res.returncode != 0:
                sys.stderr.write(f"--- {src}\n{res.stdout}{res.stderr}")
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ, os.path.splitext(s)[0] + tag + ".o") for s in SOURCES]
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs,
                         cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("linking _ctools.so failed")
    return lib


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
