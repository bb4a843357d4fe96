"""oracle -- TEST INFRASTRUCTURE (never imported by the mpdaf_b200 product path).

Two CPU checkers for the combine path:

* `port`  : oracle/merge_oracle.c, a restatement of src/merging.c + src/tools.c on
            in-memory arrays (`combine_port`, `median_port`, `clip_port`);
* `ref`   : oracle/_ref/libmpdaf_ref.so, the UNMODIFIED reference sources compiled
            against the stand-in fitsio.h (`combine_ref`, `median_ref`, `clip_ref`).
            It reads FITS files, so callers write their exposures with
            `mpdaf_b200.fits.write_cube` first.  It only exists where it was built
            (this container has /root/reference; the GPU box receives the built .so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "_build", "libmerge_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libmpdaf_ref.so")
REFERENCE_ROOT = "/root/reference"

_port = None
_ref = None


def build(ref=True):
    """Compile the port and, where the reference sources exist, the reference."""
    targets = ["port"]
    if ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def have_ref():
    return os.path.exists(REF_LIB)


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_LIB):
            build(ref=False)
        _port = ctypes.CDLL(PORT_LIB)
        _port.oracle_num_threads.restype = ctypes.c_int
    return _port


def ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_LIB):
            build(ref=True)
        _ref = ctypes.CDLL(REF_LIB)
    return _ref


def _ptrs(arrays):
    tbl = (ctypes.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    return tbl


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _params(nmax, nclip, nstop, var, mad):
    lo, up = (nclip, nclip) if np.isscalar(nclip) else nclip
    typ = {"propagate": 0, "stat_mean": 1, "stat_one": 2}.get(var, var)
    return (ctypes.c_int(nmax), ctypes.c_double(lo), ctypes.c_double(up), ctypes.c_int(nstop),
            ctypes.c_int(int(typ)), ctypes.c_int(int(mad)))


def combine_port(data, var=None, scales=None, offsets=None, nmax=2, nclip=5.0, nstop=2,
                 vartype="propagate", mad=False):
    """Array restatement of mpdaf_merging_sigma_clipping.  Returns a dict of flat arrays."""
    lib = port()
    n = len(data)
    data = [np.ascontiguousarray(a) for a in data]
    esz = data[0].dtype.itemsize
    nvox = data[0].size
    var = None if var is None else [np.ascontiguousarray(a, dtype=data[0].dtype) for a in var]
    scale = np.ones(n) if scales is None else np.ascontiguousarray(scales, dtype=np.float64)
    offset = np.zeros(n) if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64)
    out = np.empty(nvox)
    outvar = np.empty(nvox)
    expmap = np.empty(nvox, dtype=np.intc)
    sel = np.zeros(n, dtype=np.intc)
    val = np.zeros(n, dtype=np.intc)
    dt = _ptrs(data)
    vt = _ptrs(var) if var is not None else None
    rc = lib.oracle_sigma_clipping(
        ctypes.c_int(n), dt, vt, ctypes.c_int(esz), ctypes.c_longlong(nvox), _dp(out), _dp(outvar),
        _ip(expmap), _dp(scale), _dp(offset), _ip(sel), _ip(val), *_params(nmax, nclip, nstop, vartype, mad))
    assert rc == 0
    return dict(data=out, var=outvar, expmap=expmap, select_pix=sel, valid_pix=val)


def median_port(data):
    lib = port()
    n = len(data)
    data = [np.ascontiguousarray(a) for a in data]
    nvox = data[0].size
    out = np.empty(nvox)
    expmap = np.empty(nvox, dtype=np.intc)
    val = np.zeros(n, dtype=np.intc)
    rc = lib.oracle_median(ctypes.c_int(n), _ptrs(data), ctypes.c_int(data[0].dtype.itemsize),
                           ctypes.c_longlong(nvox), _dp(out), _ip(expmap), _ip(val))
    assert rc == 0
    return dict(data=out, expmap=expmap, valid_pix=val)


def _clip(lib_fn, w, nmax, lo, up, nstop):
    w = np.ascontiguousarray(w, dtype=np.float64)
    n = w.size
    idx = np.arange(n, dtype=np.intc)
    x = np.zeros(3)
    lib_fn(_dp(w), ctypes.c_int(n), _dp(x), ctypes.c_int(nmax), ctypes.c_double(lo), ctypes.c_double(up),
           ctypes.c_int(nstop), _ip(idx))
    kept = int(x[2])
    return x[0], x[1], kept, sorted(int(i) for i in idx[:kept])


def clip_ref(w, nmax=2, lo=5.0, up=5.0, nstop=2, mad=False):
    """One voxel through the reference's own tools.c clip routine."""
    lib = ref()
    fn = lib.mpdaf_mean_madsigma_clip if mad else lib.mpdaf_mean_sigma_clip
    fn.restype = None
    return _clip(fn, w, nmax, lo, up, nstop)


def clip_port(w, nmax=2, lo=5.0, up=5.0, nstop=2, mad=False):
    lib = port()
    w = np.ascontiguousarray(w, dtype=np.float64)
    n = w.size
    idx = np.arange(n, dtype=np.intc)
    x = np.zeros(3)
    lib.oracle_clip.restype = None
    lib.oracle_clip(_dp(w), ctypes.c_int(n), _dp(x), ctypes.c_int(nmax), ctypes.c_double(lo),
                    ctypes.c_double(up), ctypes.c_int(nstop), ctypes.c_int(int(mad)), _ip(idx))
    kept = int(x[2])
    return x[0], x[1], kept, sorted(int(i) for i in idx[:kept])


def _silenced(fn, *args):
    """The reference prints file lists and progress to stdout (merging.c:50-53,312-316)."""
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        return fn(*args)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


_omp_team = None


def set_ref_threads(n):
    """Default OpenMP team size for the reference's next parallel regions (`omp_set_num_threads` of the libgomp the
    reference build links): launchers such as torchrun export OMP_NUM_THREADS=1, and tests on a few planes need fewer
    threads than planes (the reference double-counts its counters otherwise, merging.c:129-132)."""
    global _omp_team
    import ctypes
    ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    _omp_team = int(n)


def ref_threads(nfiles, typ_var, omp_threads=None):
    """Effective team size of the reference (merging.c:58-90)."""
    import resource
    soft = resource.getrlimit(resource.RLIMIT_NOFILE)[0]
    t = int(soft / nfiles * 0.9)
    t = min(t, 1000 // nfiles)
    if typ_var == 0:
        t //= 2
    omp = omp_threads or _omp_team or int(os.environ.get("OMP_NUM_THREADS", os.cpu_count()))
    return max(1, min(t, omp))


def combine_ref(files, shape, scales=None, offsets=None, nmax=2, nclip=5.0, nstop=2,
                vartype="propagate", mad=False, quiet=True):
    """The reference's mpdaf_merging_sigma_clipping on FITS files (fresh name buffer per call,
    SURVEY.md 8b; keep nz > threads)."""
    lib = ref()
    n = len(files)
    nvox = int(np.prod(shape))
    scale = np.ones(n) if scales is None else np.ascontiguousarray(scales, dtype=np.float64)
    offset = np.zeros(n) if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64)
    out = np.empty(nvox)
    outvar = np.empty(nvox)
    expmap = np.empty(nvox, dtype=np.intc)
    sel = np.zeros(n, dtype=np.intc)
    val = np.zeros(n, dtype=np.intc)
    buf = ctypes.create_string_buffer("\n".join(files).encode("utf8"))
    args = (buf, _dp(out), _dp(outvar), _ip(expmap), _dp(scale), _dp(offset), _ip(sel), _ip(val),
            *_params(nmax, nclip, nstop, vartype, mad))
    rc = _silenced(lib.mpdaf_merging_sigma_clipping, *args) if quiet else lib.mpdaf_merging_sigma_clipping(*args)
    assert rc == 0
    return dict(data=out, var=outvar, expmap=expmap, select_pix=sel, valid_pix=val)


def median_ref(files, shape, quiet=True):
    lib = ref()
    n = len(files)
    nvox = int(np.prod(shape))
    out = np.empty(nvox)
    expmap = np.empty(nvox, dtype=np.intc)
    val = np.zeros(n, dtype=np.intc)
    buf = ctypes.create_string_buffer("\n".join(files).encode("utf8"))
    args = (buf, _dp(out), _ip(expmap), _ip(val))
    rc = _silenced(lib.mpdaf_merging_median, *args) if quiet else lib.mpdaf_merging_median(*args)
    assert rc == 0
    return dict(data=out, expmap=expmap, valid_pix=val)
