"""C-ABI surface and host-side pieces (CPU only; no GPU compute is attempted)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from helpers import synth_data_numpy, synth_stat_numpy, write_exposures
from mpdaf_b200 import fits as mfits
from mpdaf_b200 import synth
from mpdaf_b200.ctools import ERR_CUDA, ERR_IO, ERR_SHAPE, ctools

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpdaf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:mpdaf_\w+)|indexx)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert "mpdaf_merging_median" in names and "mpdaf_merging_sigma_clipping" in names
    assert len(names) >= 24
    for name in names:
        assert hasattr(ctools, name), f"_ctools.so does not export {name}"


def test_reference_symbol_set_is_covered():
    """Every function the reference `_ctools` exports from tools.c / merging.c that Python or
    Cython can bind (src/tools.h:12-30, ctools.py:69-92) exists in the replacement."""
    for name in ("mpdaf_mean", "mpdaf_sum", "mpdaf_median", "mpdaf_mean_mad", "indexx",
                 "mpdaf_mean_sigma_clip", "mpdaf_mean_madsigma_clip", "mpdaf_median_sigma_clip",
                 "mpdaf_locate", "mpdaf_linear_interpolation", "mpdaf_minmax", "mpdaf_minmax_int",
                 "mpdaf_merging_median", "mpdaf_merging_sigma_clipping"):
        assert hasattr(ctools, name)


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


@pytest.mark.skipif(not oracle.have_ref(), reason="reference build not available")
def test_host_helper_symbols_match_reference_tools():
    ref = oracle.ref()
    rng = np.random.default_rng(3)
    for fn in ("mpdaf_sum", "mpdaf_median", "mpdaf_linear_interpolation"):
        getattr(ref, fn).restype = ctypes.c_double
        getattr(ctools, fn).restype = ctypes.c_double
    for trial in range(400):
        n = int(rng.integers(2, 40))
        w = rng.integers(0, 9, n).astype(float) if trial % 3 == 0 else rng.standard_normal(n) * 5
        w[rng.random(n) < 0.08] += 60
        perm = rng.permutation(n).astype(np.intc)
        for name in ("mpdaf_mean", "mpdaf_mean_mad"):
            xa, xb = np.zeros(3), np.zeros(3)
            ia, ib = perm.copy(), perm.copy()
            getattr(ctools, name)(_dp(w), n, _dp(xa), _ip(ia))
            getattr(ref, name)(_dp(w), n, _dp(xb), _ip(ib))
            np.testing.assert_array_equal(xa[:2], xb[:2])
        ia, ib = perm.copy(), perm.copy()
        assert ctools.mpdaf_sum(_dp(w), n, _ip(ia)) == ref.mpdaf_sum(_dp(w), n, _ip(ib))
        assert ctools.mpdaf_median(_dp(w), n, _ip(ia)) == ref.mpdaf_median(_dp(w), n, _ip(ib))
        np.testing.assert_array_equal(w[ia], w[ib])                  # both argsorted in place
        ia, ib = perm.copy(), perm.copy()
        ctools.indexx(n, _dp(w), _ip(ia))
        ref.indexx(n, _dp(w), _ip(ib))
        np.testing.assert_array_equal(w[ia], np.sort(w))
        np.testing.assert_array_equal(w[ia], w[ib])
        nmax, nstop = int(rng.integers(0, 4)), int(rng.integers(1, 4))
        lo, up = float(rng.choice([1.5, 2, 3, 5])), float(rng.choice([1.5, 2, 3, 5]))
        for name in ("mpdaf_mean_sigma_clip", "mpdaf_mean_madsigma_clip", "mpdaf_median_sigma_clip"):
            xa, xb = np.zeros(3), np.zeros(3)
            ia, ib = np.arange(n, dtype=np.intc), np.arange(n, dtype=np.intc)
            getattr(ctools, name)(_dp(w), n, _dp(xa), nmax, ctypes.c_double(lo), ctypes.c_double(up), nstop, _ip(ia))
            getattr(ref, name)(_dp(w), n, _dp(xb), nmax, ctypes.c_double(lo), ctypes.c_double(up), nstop, _ip(ib))
            np.testing.assert_array_equal(xa, xb)
            k = int(xa[2])
            np.testing.assert_array_equal(np.sort(ia[:k]), np.sort(ib[:k]))
        ra, rb = np.zeros(2), np.zeros(2)
        ctools.mpdaf_minmax(_dp(w), n, _ip(perm), _dp(ra))
        ref.mpdaf_minmax(_dp(w), n, _ip(perm), _dp(rb))
        np.testing.assert_array_equal(ra, rb)
        wi = rng.integers(-50, 50, n).astype(np.intc)
        qa, qb = np.zeros(2, dtype=np.intc), np.zeros(2, dtype=np.intc)
        ctools.mpdaf_minmax_int(_ip(wi), n, _ip(perm), _ip(qa))
        ref.mpdaf_minmax_int(_ip(wi), n, _ip(perm), _ip(qb))
        np.testing.assert_array_equal(qa, qb)
        xs = np.sort(rng.standard_normal(n)).cumsum() + np.arange(n)
        ys = rng.standard_normal(n)
        x = float(rng.uniform(xs[0] - 1, xs[-1] + 1))
        assert ctools.mpdaf_locate(_dp(xs), n, ctypes.c_double(x)) == ref.mpdaf_locate(_dp(xs), n, ctypes.c_double(x))
        assert (ctools.mpdaf_linear_interpolation(_dp(xs), _dp(ys), n, ctypes.c_double(x))
                == ref.mpdaf_linear_interpolation(_dp(xs), _dp(ys), n, ctypes.c_double(x)))


def test_synth_generator_is_pinned_by_numpy_restatement():
    nx, ny = 13, 7
    for exposure in (0, 1, 5, 47):
        for v0 in (0, 1000, 123456789):
            got = synth.fill_host(synth.DATA, 9, exposure, v0, 4000, nx, ny)
            want = synth_data_numpy(9, exposure, v0, 4000, nx, ny)
            np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
            got = synth.fill_host(synth.STAT, 9, exposure, v0, 4000, nx, ny)
            want = synth_stat_numpy(9, exposure, v0, 4000)
            np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    d = synth.fill_host(synth.DATA, 1, 3, 0, 200000, 331, 326)
    assert 0.04 < np.isnan(d).mean() < 0.08
    good = d[~np.isnan(d)]
    assert 0.005 < (good > 40).mean() < 0.015
    assert abs(np.median(good) - 10.0) < 0.05 and 0.9 < good[good < 40].std() < 1.1
    s = synth.fill_host(synth.STAT, 1, 3, 0, 200000, 331, 326)
    assert s.min() >= 0.5 and s.max() < 2.0


def test_synth_slabs_tile_the_full_cube():
    """A slab regenerated on its own equals the same planes of the full cube."""
    shape, nx, ny = (6, 5, 4), 4, 5
    full, fvar = synth.host_stack(3, shape, seed=2)
    part, pvar = synth.host_stack(3, (2, 5, 4), seed=2, z0=3)
    for i in range(3):
        np.testing.assert_array_equal(full[i][3:5].view(np.uint32), part[i].view(np.uint32))
        np.testing.assert_array_equal(fvar[i][3:5], pvar[i])


def test_fits_roundtrip_and_probe(tmp_path):
    rng = np.random.default_rng(0)
    data = rng.standard_normal((4, 3, 5)).astype(np.float32)
    data[1, 2, 3] = np.nan
    var = rng.random((4, 3, 5)).astype(np.float32)
    path = str(tmp_path / "c.fits")
    mfits.write_cube(path, data, var, primary_header={"OBJECT": "TEST", "EXPTIME": 100,
                                                      "HIERARCH ESO OBS ID": 42})
    np.testing.assert_array_equal(np.asarray(mfits.read_image(path, "DATA")), data)
    np.testing.assert_array_equal(np.asarray(mfits.read_image(path, "stat")), var)
    hdus = mfits.read_hdus(path)
    assert hdus[0][0]["OBJECT"] == "TEST" and hdus[0][0]["EXPTIME"] == 100
    assert hdus[0][0]["HIERARCH ESO OBS ID"] == 42
    naxes = (ctypes.c_long * 3)()
    bitpix = ctypes.c_int()
    off = ctypes.c_longlong()
    rc = ctools.mpdaf_b200_fits_image_info(path.encode(), b"STAT", ctypes.byref(naxes), ctypes.byref(bitpix),
                                           ctypes.byref(off))
    assert rc == 0 and list(naxes) == [5, 3, 4] and bitpix.value == -32
    assert off.value == mfits.find_image(path, "STAT")[1]
    rc = ctools.mpdaf_b200_fits_image_info(path.encode(), b"NOPE", ctypes.byref(naxes), None, None)
    assert rc == ERR_IO
    rc = ctools.mpdaf_b200_fits_image_info(b"/nonexistent.fits", b"DATA", ctypes.byref(naxes), None, None)
    assert rc == ERR_IO


@pytest.mark.skipif(__import__("conftest").HAS_GPU, reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu(tmp_path):
    """The product path fails loudly when it cannot run on a B200: no silent CPU route."""
    assert ctools.mpdaf_b200_device_count() == 0
    data = [np.ones((3, 2, 2), dtype=np.float32)] * 2
    files = write_exposures(tmp_path, data, data)
    out = np.full(12, -1.0)
    expmap = np.full(12, -1, dtype=np.intc)
    valid = np.zeros(2, dtype=np.intc)
    buf = ctypes.create_string_buffer("\n".join(files).encode())
    rc = ctools.mpdaf_merging_median(buf, out, expmap, valid)
    assert rc == ERR_CUDA
    assert buf.value == "\n".join(files).encode()          # the name list is never mutated
    # a failed call never leaves caller-allocated buffers half written (the reference's caller ignores the return
    # value, cubelist.py:456): NaN cube, empty exposure map, untouched counters
    assert np.isnan(out).all() and (expmap == 0).all() and (valid == 0).all()
    from mpdaf_b200 import CubeList
    from mpdaf_b200.ctools import CtoolsError
    with pytest.raises(CtoolsError):
        CubeList(files).combine()


def test_shape_mismatch_is_an_error(tmp_path):
    a = write_exposures(tmp_path, [np.ones((3, 2, 2), dtype=np.float32)], prefix="a")
    b = write_exposures(tmp_path, [np.ones((3, 2, 3), dtype=np.float32)], prefix="b")
    out = np.zeros(12)
    expmap = np.zeros(12, dtype=np.intc)
    valid = np.zeros(2, dtype=np.intc)
    rc = ctools.mpdaf_merging_median(ctypes.create_string_buffer((a[0] + "\n" + b[0]).encode()), out, expmap, valid)
    assert rc == ERR_SHAPE                                 # merging.c:195-199
    assert np.isnan(out).all() and (expmap == 0).all()     # outputs of the failed call are filled, not left as they were


def _patch_card(path, key, card_text):
    """Overwrite (or append before END) one 80-character card of the DATA extension's header."""
    raw = bytearray(open(path, "rb").read())
    pos = raw.find(b"EXTNAME = 'DATA")
    assert pos > 0
    start = pos - pos % 2880
    end = raw.find(b"END" + b" " * 77, start)
    card = card_text.ljust(80).encode()
    k = raw.find(key.ljust(8).encode() + b"=", start, end)
    if k < 0:
        # replace the END card and put END in the next slot (the header block has room in these small files)
        raw[end:end + 80] = card
        raw[end + 80:end + 160] = b"END".ljust(80)
    else:
        raw[k:k + 80] = card
    open(path, "wb").write(bytes(raw))


def test_fits_header_hardening(tmp_path):
    """Malformed or unsupported headers are refused with an error code, never indexed out of bounds
    (fits_reader.cpp: NAXIS0 / NAXIS9 cards, BSCALE / BZERO, which cfitsio would apply while reading)."""
    cube = [np.arange(12, dtype=np.float32).reshape(3, 2, 2)]
    naxes = (ctypes.c_long * 3)()
    bitpix = ctypes.c_int()
    off = ctypes.c_longlong()
    for key, text, ok in (("NAXIS9", "NAXIS9  =                   77", True), ("NAXIS0", "NAXIS0  =                   77", True),
                          ("BSCALE", "BSCALE  =                  2.0", False), ("BZERO", "BZERO   =                 10.0", False),
                          ("BSCALE", "BSCALE  =                  1.0", True)):
        f = write_exposures(tmp_path, cube, prefix="h" + key + str(ok))[0]
        _patch_card(f, key, text)
        rc = ctools.mpdaf_b200_fits_image_info(f.encode(), b"DATA", ctypes.byref(naxes), ctypes.byref(bitpix), ctypes.byref(off))
        assert (rc == 0) == ok, (key, text, rc)
        if ok:
            assert tuple(naxes) == (2, 2, 3) and bitpix.value == -32


REF_LOADER = "/root/reference/lib/mpdaf/tools/ctools.py"


@pytest.mark.skipif(not os.path.exists(REF_LOADER), reason="the reference tree is not present on this machine")
def test_library_binds_through_the_reference_loader(tmp_path):
    """Drop-in boundary: the reference's UNMODIFIED lib/mpdaf/tools/ctools.py (loaded from where it lies, never
    copied into the repository) finds `_ctools.so` beside itself, binds both merging symbols with its own argtypes
    and gets a clean error code and filled outputs on a machine without a B200 (ctools.py:52-92, cubelist.py:456)."""
    import importlib.util
    import shutil
    import subprocess
    import sys
    work = tmp_path / "tools"
    work.mkdir()
    shutil.copy(REF_LOADER, work / "ctools.py")            # temp dir only: the loader looks for _ctools next to itself
    os.symlink(os.path.join(ROOT, "mpdaf_b200", "_ctools.so"), work / "_ctools.so")
    data = [np.full((3, 2, 2), v, dtype=np.float32) for v in (1.0, 2.0, 6.0)]
    files = write_exposures(tmp_path, data, data)
    code = f"""
import sys, importlib.util, numpy as np, ctypes
spec = importlib.util.spec_from_file_location('refctools', r'{work / "ctools.py"}')
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
lib = m.ctools
n, nv = 3, 12
files = {files!r}
data, var = np.empty(nv), np.empty(nv)
expmap = np.empty(nv, dtype=np.int32)
sel, val = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
rc = lib.mpdaf_merging_sigma_clipping(ctypes.c_char_p('\\n'.join(files).encode('utf8')), data, var, expmap,
                                      np.ones(n), np.zeros(n), sel, val, 2, 5.0, 5.0, 2, 0, 0)
rc2 = lib.mpdaf_merging_median(ctypes.c_char_p('\\n'.join(files).encode('utf8')), data, expmap, val)
print('RC', rc, rc2, int(np.isnan(data).all()), int((expmap == 0).all()), int(np.isnan(var).all()))
"""
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RC")][0].split()
    if ctools.mpdaf_b200_device_count() == 0:
        assert line[1:] == [str(ERR_CUDA), str(ERR_CUDA), "1", "1", "1"]
    else:
        assert line[1:3] == ["0", "0"]
