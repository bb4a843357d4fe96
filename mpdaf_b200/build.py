"""Builds mpdaf_b200/_ctools.so (the CUDA replacement for MPDAF's `_ctools`).

In-tree build with nvcc for sm_100a only; the .so is git-ignored but travels to
the GPU box with the repository snapshot.  `python -m mpdaf_b200.build` or
`__graft_entry__.build()`.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "_ctools.so")
SOURCES = ("ctools_abi.cu", "fits_reader.cpp", "host_tools.cpp")
DEPS = SOURCES + ("combine_kernels.cuh", "sigclip_tile.cuh", "sigclip_warp.cuh", "sort_networks.cuh", "synth_core.h", "fits_reader.h",
                  os.path.join("..", "..", "include", "mpdaf_b200.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                       # no FMA contraction in the float64 statistics
    "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force=False, verbose=False):
    """Compile the library if a source is newer than the .so.  Returns its path."""
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + list(SOURCES)
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    env = dict(os.environ)
    env.setdefault("PATH", "")
    res = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building _ctools.so")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
