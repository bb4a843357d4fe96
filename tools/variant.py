#!/usr/bin/env python3
"""Tuning builds: recompile ONE depth bucket with extra -D flags and link it with the stock objects into
mpdaf_b200/_ctools_<tag>.so (select it with MPDAF_B200_LIBNAME=_ctools_<tag>).
  python tools/variant.py late0b4 warp_48.cu -DMB_LATE_STAT=0 -DMB_WARP_BLOCKS=4
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpdaf_b200 import build as B


def main():
    tag, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
    if not os.environ.get("VARIANT_NO_STOCK"):
        B.build_library()   # stock objects up to date
    obj = os.path.join(B.OBJ, os.path.splitext(src)[0] + "_" + tag + ".o")
    cmd = [B._nvcc()] + B.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-c", "-o", obj, src]
    res = subprocess.run(cmd, cwd=B.CSRC, capture_output=True, text=True)
    if res.returncode:
        sys.exit((res.stdout + res.stderr)[:4000])
    for ln in (res.stdout + res.stderr).splitlines():
        if "registers" in ln or "spill" in ln:
            pass
    objs = [obj if s == src else os.path.join(B.OBJ, os.path.splitext(s)[0] + ".o") for s in B.SOURCES]
    lib = os.path.join(B.HERE, f"_ctools_{tag}.so")
    subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs, check=True)
    out = (res.stdout + res.stderr).splitlines()
    for k, ln in enumerate(out):
        if "Compiling entry function" in ln and "ELb0ELb1ELb0ELb1" in ln and "sigclip_warp" in ln:
            print(tag, out[k + 2].strip(), "|", out[k + 3].strip())
    print(lib)


if __name__ == "__main__":
    main()
