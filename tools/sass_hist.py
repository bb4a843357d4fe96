#!/usr/bin/env python3
"""Static SASS opcode histogram of one kernel (no GPU needed): compile a translation unit of
mpdaf_b200/csrc with extra -D flags and count the instructions of the functions matching a regex.
  python tools/sass_hist.py warp_48.cu 'sigclip_warp_kernelILi48ELb0ELb1ELb0ELb1' -DMB_EXACT_SUM=0
"""
import collections
import os
import re
import subprocess
import sys

CSRC = os.environ.get("CSRC") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mpdaf_b200", "csrc")


def main():
    src, pat, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
    obj = "/tmp/sass/" + os.path.splitext(src)[0] + "".join(c if c.isalnum() else "_" for c in "".join(extra)) + ".o"
    os.makedirs("/tmp/sass", exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
           "-Xptxas", "-v"] + extra + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode:
        sys.exit((res.stdout + res.stderr)[:3000])
    lines = (res.stdout + res.stderr).splitlines()
    for k, ln in enumerate(lines):
        if "Compiling entry function" in ln and re.search(pat, ln):
            print("\n".join(x.strip() for x in lines[k:k + 4]))
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur = None
    hist = collections.defaultdict(collections.Counter)
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1) if re.search(pat, m.group(1)) else None
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            base = op.split(".")[0]
            hist[cur][op if base in ("F2F", "LDG", "LDCU", "LDC") else base] += 1
    for fn, h in hist.items():
        print(fn, "total", sum(h.values()))
        print("  " + "  ".join(f"{k}:{v}" for k, v in h.most_common(40)))


if __name__ == "__main__":
    main()
