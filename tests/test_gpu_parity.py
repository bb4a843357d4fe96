"""Parity of the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.

Sizes are chosen so the oracle finishes in seconds.  Bit-exact: expmap, per-exposure
counters, medians, NaN positions.  <= 1e-6 relative: sigma-clipped means and variances.
"""
import ctypes
import itertools

import numpy as np
import pytest

import oracle
from helpers import assert_same_result, random_stack, write_exposures

pytestmark = pytest.mark.gpu


def cuda():
    import cuda_runner
    return cuda_runner


def is_main_kernel(name):
    """The two sigma-clip kernel families for float32 stacks up to 128 deep: the pair-packed TMA-pipelined kernel
    (sigclip_pair.cuh: equally pitched, 16-byte aligned rows, <= 64 exposures, no MAD) and the warp-tiled one."""
    return "sigclip_pair<" in name or "sigclip_warp<" in name or "sigclip_duo<" in name


@pytest.fixture(params=["pair", "warp"])
def family(request, monkeypatch):
    """Runs a test once per kernel family (MPDAF_B200_NO_PAIR is read per call)."""
    if request.param == "warp":
        monkeypatch.setenv("MPDAF_B200_NO_PAIR", "1")
    else:
        monkeypatch.delenv("MPDAF_B200_NO_PAIR", raising=False)
    return request.param


@pytest.mark.parametrize("location", ["device", "host"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 9, 12, 13, 16, 24, 25, 32, 33, 48, 49, 64, 65, 96, 97, 128])
def test_register_kernels_every_bucket(n, location, family):
    rng = np.random.default_rng(100 + n)
    shape = (6, 11, 13)          # 858 voxels: not a multiple of the block size
    data, var = random_stack(rng, n, shape, nan_frac=0.12, outlier_frac=0.04)
    scales = np.linspace(0.9, 1.1, n)
    offsets = np.linspace(-0.1, 0.1, n)
    for kw in (dict(), dict(scales=scales, offsets=offsets)):
        for vartype, nclip in (("propagate", 3.0), ("stat_mean", (2.0, 3.0)), ("stat_one", 5.0)):
            want = oracle.combine_port(data, var, vartype=vartype, nclip=nclip, **kw)
            got = cuda().combine_cuda(data, var, vartype=vartype, nclip=nclip, location=location, **kw)
            assert_same_result(got, want)
            assert is_main_kernel(cuda().kernel_name())
            if family == "warp":
                assert "sigclip_warp<" in cuda().kernel_name()
            elif n > 64:
                assert "sigclip_warp<" in cuda().kernel_name() or "sigclip_duo<" in cuda().kernel_name()
    want = oracle.median_port(data)
    got = cuda().median_cuda(data, location=location)
    assert_same_result(got, want, median=True)


def test_unit_scale_path_is_bit_exact(monkeypatch):
    """Warp kernel: voxels without a rejection are identical to the bit (first mean in exposure order);
    after a rejection the mean comes from shifted moments: <= 1e-12 relative."""
    monkeypatch.setenv("MPDAF_B200_NO_PAIR", "1")
    rng = np.random.default_rng(5)
    data, var = random_stack(rng, 48, (5, 9, 10), nan_frac=0.05, outlier_frac=0.03)
    for nclip in (5.0, 3.0, 2.0):
        want = oracle.combine_port(data, var, nclip=nclip, vartype="stat_one")
        got = cuda().combine_cuda(data, var, nclip=nclip, vartype="stat_one")
        assert "unit" in cuda().kernel_name()
        assert_same_result(got, want, rtol=1e-12)
        untouched = want["expmap"] == sum(~np.isnan(d.reshape(-1)) for d in data)
        np.testing.assert_array_equal(got["data"][untouched], want["data"][untouched])
        assert untouched.any() and (~untouched).any()
        assert (want["select_pix"] < want["valid_pix"]).any()


def test_scaled_path_first_mean_is_bit_exact(monkeypatch):
    monkeypatch.setenv("MPDAF_B200_NO_PAIR", "1")
    rng = np.random.default_rng(6)
    n = 24
    data, var = random_stack(rng, n, (5, 9, 10), nan_frac=0.05, outlier_frac=0.05)
    want = oracle.combine_port(data, var, scales=np.linspace(0.5, 2, n), offsets=np.linspace(-3, 3, n),
                               nclip=2.5, vartype="stat_mean")
    got = cuda().combine_cuda(data, var, scales=np.linspace(0.5, 2, n), offsets=np.linspace(-3, 3, n),
                              nclip=2.5, vartype="stat_mean")
    assert "scaled" in cuda().kernel_name()
    assert_same_result(got, want, rtol=1e-12)
    untouched = want["expmap"] == sum(~np.isnan(d.reshape(-1)) for d in data)
    np.testing.assert_array_equal(got["data"][untouched], want["data"][untouched])


@pytest.mark.parametrize("scaled", [False, True])
def test_pair_kernel_means_within_1e_12(scaled):
    """Pair kernel: every mean is pivot + sum(w - pivot) / n in float64 (no exposure-order sum of w): within a few ulp
    of the reference whether or not something was rejected; decisions (expmap, counters) identical."""
    rng = np.random.default_rng(15)
    n = 48
    data, var = random_stack(rng, n, (5, 9, 10), nan_frac=0.05, outlier_frac=0.03)
    kw = dict(scales=np.linspace(0.5, 2, n), offsets=np.linspace(-3, 3, n)) if scaled else {}
    for nclip, vartype in ((5.0, "stat_one"), (3.0, "stat_mean"), (2.0, "propagate")):
        want = oracle.combine_port(data, var, nclip=nclip, vartype=vartype, **kw)
        got = cuda().combine_cuda(data, var, nclip=nclip, vartype=vartype, location="device_pitched", **kw)
        assert "sigclip_pair<48," + ("scaled" if scaled else "unit") in cuda().kernel_name()
        np.testing.assert_array_equal(got["expmap"], want["expmap"])
        np.testing.assert_array_equal(got["select_pix"], want["select_pix"])
        np.testing.assert_allclose(got["data"], want["data"], rtol=1e-12, atol=0, equal_nan=True)
        if vartype != "propagate":
            np.testing.assert_allclose(got["var"], want["var"], rtol=1e-12, atol=0, equal_nan=True)
        else:
            np.testing.assert_allclose(got["var"], want["var"], rtol=3e-7, atol=0, equal_nan=True)   # float32 chains of 3


def test_pair_kernel_deep_bucket_with_stat():
    """49 .. 64 exposures with var = 'propagate': the pair kernel (eight warps, each refilling its own stage) is the
    dispatch's choice since the producer warp went (3591 vs 2602 GB/s of the warp kernel at N = 64)."""
    rng = np.random.default_rng(64)
    for n in (49, 64):
        data, var = random_stack(rng, n, (4, 9, 10), nan_frac=0.1, outlier_frac=0.05)
        kw = dict(scales=np.linspace(0.8, 1.3, n), offsets=np.linspace(-1, 1, n), nclip=2.5, vartype="propagate")
        got = cuda().combine_cuda(data, var, location="device_pitched", **kw)
        assert "sigclip_pair<64" in cuda().kernel_name()
        assert_same_result(got, oracle.combine_port(data, var, **kw))


@pytest.mark.parametrize("nmax,lo,up,nstop", [(0, 2.0, 2.0, 2), (1, 1.5, 3.0, 1), (3, 1.0, 1.0, 3),
                                              (2, 3.0, 1.5, 5), (5, 1.2, 1.2, 2), (2, 0.5, 0.5, 1),
                                              (2, 5.0, 5.0, 0), (10, 1.0, 2.0, 2)])
@pytest.mark.parametrize("ties", [False, True])
def test_clip_parameters(nmax, lo, up, nstop, ties, family):
    rng = np.random.default_rng(nmax * 31 + nstop)
    data, var = random_stack(rng, 16, (4, 8, 9), nan_frac=0.2, outlier_frac=0.1, ties=ties)
    if nstop < 1:
        # nstop >= 1 is a precondition of the reference (ni == 0 would recurse on an empty
        # set); the replacement treats nstop < 1 as 1.
        want = oracle.combine_port(data, var, nmax=nmax, nclip=(lo, up), nstop=1, vartype="stat_one")
    else:
        want = oracle.combine_port(data, var, nmax=nmax, nclip=(lo, up), nstop=nstop, vartype="stat_one")
    got = cuda().combine_cuda(data, var, nmax=nmax, nclip=(lo, up), nstop=nstop, vartype="stat_one")
    assert_same_result(got, want)


@pytest.mark.parametrize("n", [5, 48, 70, 96, 128, 130, 200])
@pytest.mark.parametrize("vartype", ["propagate", "stat_mean"])
def test_mad_and_deep_stacks(n, vartype):
    rng = np.random.default_rng(n)
    data, var = random_stack(rng, n, (3, 5, 7), nan_frac=0.1, outlier_frac=0.05)
    kw = dict(scales=np.linspace(0.9, 1.1, n), offsets=np.linspace(-0.1, 0.1, n), nclip=3.0, vartype=vartype)
    for mad in (True, False):
        want = oracle.combine_port(data, var, mad=mad, **kw)
        got = cuda().combine_cuda(data, var, mad=mad, **kw)
        assert_same_result(got, want)
        if n > 128:
            assert ("generic" if mad else "sigclip_runs") in cuda().kernel_name()
        else:
            assert is_main_kernel(cuda().kernel_name()) and ("mad" in cuda().kernel_name()) == mad
    if n > 64:
        assert_same_result(cuda().median_cuda(data), oracle.median_port(data), median=True)


@pytest.mark.parametrize("n", [65, 77, 96, 97, 111, 128])
def test_duo_kernel_65_to_128_exposures(n):
    """65 .. 128 exposures on pitched rows: the pair kernel with the stack split between two lanes (sigclip_duo.cuh):
    every variance mode, scales, ties, heavy clipping, no clip at all, ragged last tile, host route."""
    rng = np.random.default_rng(2000 + n)
    shape = (3, 7, 23)      # 483 voxels: the last tile of 32 is ragged
    for ties in (False, True):
        data, var = random_stack(rng, n, shape, nan_frac=0.12, outlier_frac=0.04, ties=ties)
        scales = np.linspace(0.9, 1.1, n)
        offsets = np.linspace(-0.1, 0.1, n)
        for kw in (dict(nclip=3.0, vartype="propagate"),
                   dict(scales=scales, offsets=offsets, nclip=(2.0, 3.0), vartype="stat_mean"),
                   dict(scales=scales, offsets=offsets, nclip=5.0, vartype="stat_one"),
                   dict(scales=scales, offsets=offsets, nclip=1.5, nmax=5, vartype="propagate"),
                   dict(nclip=5.0, nmax=0, vartype="stat_mean"),
                   dict(nclip=1.0, nmax=3, nstop=40, vartype="propagate")):
            want = oracle.combine_port(data, var, **kw)
            for location in ("device_pitched", "host"):
                got = cuda().combine_cuda(data, var, location=location, **kw)
                assert "sigclip_duo<" in cuda().kernel_name(), cuda().kernel_name()
                assert_same_result(got, want)
        # MAD with var = 'propagate': sigma from the four sorted runs of absolute key deviations of the two columns
        for kw in (dict(nclip=3.0, mad=True), dict(scales=scales, offsets=offsets, nclip=(2.0, 2.5), nmax=4, mad=True)):
            want = oracle.combine_port(data, var, **kw)
            got = cuda().combine_cuda(data, var, location="device_pitched", **kw)
            assert "sigclip_duo<" in cuda().kernel_name() and "mad" in cuda().kernel_name(), cuda().kernel_name()
            assert_same_result(got, want)
    # non-finite samples and a tiny queue (voxels recomputed in place)
    data[3][0, 0, :6] = np.inf
    data[70 % n][0, 1, :6] = -np.inf
    want = oracle.combine_port(data, var, nclip=2.0)
    assert_same_result(cuda().combine_cuda(data, var, nclip=2.0, location="device_pitched"), want)
    # the median of the same stacks: two sorted half-stacks per voxel, the middle of their union (median_duo_kernel)
    data[5][1, :, :] = np.nan
    wm = oracle.median_port(data)
    gm = cuda().median_cuda(data, location="host")
    assert ("median_duo<" in cuda().kernel_name()) == (n <= 96), cuda().kernel_name()
    assert_same_result(gm, wm, median=True)


@pytest.mark.parametrize("n", [129, 192, 256, 257, 400, 500])
def test_runs_kernel_deep_stacks(n):
    """129 .. 512 exposures (sorted runs of 64, median by elimination over the runs, tiles handed out dynamically):
    separate allocations (row table) and one [N][V] allocation (strided rows), ties, every variance mode."""
    rng = np.random.default_rng(1000 + n)
    shape = (2, 9, 31)     # 558 voxels: 18 tiles, the last one ragged
    for ties in (False, True):
        data, var = random_stack(rng, n, shape, nan_frac=0.1, outlier_frac=0.03, ties=ties)
        for kw in (dict(nclip=3.0, vartype="propagate"),
                   dict(scales=np.linspace(0.9, 1.1, n), offsets=np.linspace(-0.1, 0.1, n), nclip=(2.0, 2.5), vartype="stat_mean"),
                   dict(scales=np.linspace(0.5, 2.0, n), nclip=1.5, nmax=4, vartype="propagate")):
            want = oracle.combine_port(data, var, **kw)
            for location in ("device_scattered", "device_stacked", "host"):
                got = cuda().combine_cuda(data, var, location=location, **kw)
                name = cuda().kernel_name()
                assert "sigclip_runs<" in name and (("strided" in name) == (location != "device_scattered"))
                assert_same_result(got, want)
        # the median of deep stacks runs on the same sorted runs (bit-identical to the reference)
        data[1][0, 0, :5] = np.inf
        data[2][0, 1, :5] = -np.inf
        want = oracle.median_port(data)
        for location in ("device", "host"):
            got = cuda().median_cuda(data, location=location)
            assert "median_runs<" in cuda().kernel_name()
            assert_same_result(got, want, median=True)


@pytest.mark.parametrize("location", ["device", "host"])
def test_float64_exposures(location):
    rng = np.random.default_rng(8)
    data, var = random_stack(rng, 7, (4, 5, 6), dtype=np.float64)
    want = oracle.combine_port(data, var, nclip=2.0)
    got = cuda().combine_cuda(data, var, nclip=2.0, location=location)
    assert_same_result(got, want)
    assert_same_result(cuda().median_cuda(data, location=location), oracle.median_port(data), median=True)


@pytest.mark.parametrize("n", [2, 5, 12, 13, 24, 37, 48, 49, 64])
def test_float64_exposures_on_the_pair_kernel(n):
    """BITPIX -64 stacks up to 64 deep on equally pitched rows: the pair kernel with float64 stages (float64 samples in
    the moments and in every exact decision, float64 STAT products); every mode, host route included."""
    rng = np.random.default_rng(300 + n)
    data, var = random_stack(rng, n, (5, 9, 11), nan_frac=0.1, outlier_frac=0.04, dtype=np.float64)
    # samples float32 cannot tell apart: the decisions must come from the float64 values
    data[0] += 1e-9 * rng.standard_normal(data[0].shape)
    scales = np.linspace(0.9, 1.1, n)
    offsets = np.linspace(-0.1, 0.1, n)
    for kw in (dict(), dict(scales=scales, offsets=offsets)):
        for vartype, nclip, mad in (("propagate", 3.0, False), ("stat_mean", (2.0, 3.0), False), ("stat_one", 5.0, False),
                                    ("propagate", 2.5, True)):
            want = oracle.combine_port(data, var, vartype=vartype, nclip=nclip, mad=mad, **kw)
            got = cuda().combine_cuda(data, var, vartype=vartype, nclip=nclip, mad=mad, location="device_pitched", **kw)
            assert "sigclip_pair<" in cuda().kernel_name() and "f64" in cuda().kernel_name()
            assert_same_result(got, want)
    want = oracle.combine_port(data, var, nclip=3.0)
    assert_same_result(cuda().combine_cuda(data, var, nclip=3.0, location="host"), want)
    assert_same_result(cuda().combine_cuda(data, var, nclip=3.0, location="device"), want)     # separate allocations: generic kernel


def test_float64_ties_and_non_finite_samples():
    rng = np.random.default_rng(41)
    data, var = random_stack(rng, 24, (4, 7, 9), ties=True, dtype=np.float64)
    data[3][0, 0, :4] = np.inf
    data[5][0, 1, :4] = -np.inf
    data[7][0, 2, :4] = 1e300
    for vartype in ("propagate", "stat_mean"):
        want = oracle.combine_port(data, var, vartype=vartype, nclip=2.0)
        got = cuda().combine_cuda(data, var, vartype=vartype, nclip=2.0, location="device_pitched")
        assert "f64" in cuda().kernel_name()
        assert_same_result(got, want)


@pytest.mark.parametrize("dtype,n", [(np.float32, 5), (np.float32, 41), (np.float64, 8), (np.float64, 31)])
def test_partial_bucket_reads_no_row_beyond_the_stack(dtype, n):
    """A stack that does not fill its depth bucket, placed at the very END of its own device allocation: the STAT loads of
    the pair kernel must not address the bucket's empty slots (an out-of-bounds read there is an illegal-address fault)."""
    import torch
    rng = np.random.default_rng(77 + n)
    shape = (3, 16, 32)                      # 1536 voxels: rows are multiples of 128 bytes
    nvox = int(np.prod(shape))
    data, var = random_stack(rng, n, shape, nan_frac=0.1, outlier_frac=0.04, dtype=dtype)
    want = oracle.combine_port(data, var, nclip=3.0)
    dev = torch.device("cuda", 0)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    seg = 64 << 20                            # blocks of this size come straight from cudaMalloc: nothing mapped behind them
    bufs = []
    ptrs = []
    for arrs in (data, var):
        buf = torch.empty(seg // np.dtype(dtype).itemsize, dtype=tdt, device=dev)
        tail = buf[buf.numel() - n * nvox:].view(n, nvox)
        tail.copy_(torch.from_numpy(np.stack([a.reshape(-1) for a in arrs])))
        bufs.append(buf)
        ptrs.append([tail[i].data_ptr() for i in range(n)])
    out = torch.empty(nvox, dtype=torch.float64, device=dev)
    outvar = torch.empty(nvox, dtype=torch.float64, device=dev)
    expmap = torch.empty(nvox, dtype=torch.int32, device=dev)
    sel = np.zeros(n, dtype=np.intc)
    val = np.zeros(n, dtype=np.intc)
    from mpdaf_b200.ctools import DEVICE, check, ctools
    kd = (ctypes.c_void_p * n)(*ptrs[0])
    kv = (ctypes.c_void_p * n)(*ptrs[1])
    rc = ctools.mpdaf_b200_sigma_clipping_arrays(
        n, ctypes.cast(kd, ctypes.POINTER(ctypes.c_void_p)), ctypes.cast(kv, ctypes.POINTER(ctypes.c_void_p)),
        np.dtype(dtype).itemsize, nvox, DEVICE, 0, out.data_ptr(), outvar.data_ptr(), expmap.data_ptr(), None, None, sel, val,
        2, 3.0, 3.0, 2, 0, 0, torch.cuda.current_stream(dev).cuda_stream)
    check(rc, "combine")
    torch.cuda.synchronize(dev)               # a fault would surface here
    assert "sigclip_pair<" in cuda().kernel_name()
    got = dict(data=out.cpu().numpy(), var=outvar.cpu().numpy(), expmap=expmap.cpu().numpy(), select_pix=sel, valid_pix=val)
    assert_same_result(got, want)


def test_edge_cases():
    c = cuda()
    # empty stack
    empty = [np.zeros((0, 3, 3), dtype=np.float32)] * 3
    got = c.combine_cuda(empty, empty)
    assert got["data"].size == 0 and (got["valid_pix"] == 0).all()
    # every sample NaN / exactly one valid exposure / constant data (sigma == 0)
    shape = (2, 3, 4)
    nan = np.full(shape, np.nan, dtype=np.float32)
    one = np.full(shape, 7.5, dtype=np.float32)
    for stack in ([nan, nan, nan], [nan, one, nan], [one, one, one, one], [one]):
        var = [np.full(shape, 2.0, dtype=np.float32) for _ in stack]
        for vartype in ("propagate", "stat_mean", "stat_one"):
            want = oracle.combine_port(stack, var, vartype=vartype, scales=[2.0] * len(stack))
            got = c.combine_cuda(stack, var, vartype=vartype, scales=[2.0] * len(stack))
            assert_same_result(got, want, exact_data=True)
        assert_same_result(c.median_cuda(stack), oracle.median_port(stack), median=True)
    # ragged NaN pattern: exposure i misses the first i planes entirely
    rng = np.random.default_rng(1)
    data, var = random_stack(rng, 6, (8, 4, 4), nan_frac=0.0)
    for i, d in enumerate(data):
        d[:i] = np.nan
    assert_same_result(c.combine_cuda(data, var, nclip=2.0), oracle.combine_port(data, var, nclip=2.0))
    # counters are ADDED to caller values (merging.c:486-491): run twice through one buffer
    r1 = c.combine_cuda(data, var, nclip=2.0)
    assert (r1["valid_pix"] == oracle.combine_port(data, var, nclip=2.0)["valid_pix"]).all()


def test_large_slab_multiple_chunks(monkeypatch):
    """Host-location call with a tiny staging budget: many pipelined chunks, same result."""
    monkeypatch.setenv("MPDAF_B200_SLAB_BYTES", str(1 << 20))
    rng = np.random.default_rng(77)
    data, var = random_stack(rng, 12, (40, 30, 31))
    want = oracle.combine_port(data, var, nclip=3.0)
    got = cuda().combine_cuda(data, var, nclip=3.0, location="host")
    assert_same_result(got, want)
    assert_same_result(cuda().median_cuda(data, location="host"), oracle.median_port(data), median=True)


@pytest.mark.parametrize("vartype,mad", list(itertools.product(["propagate", "stat_mean", "stat_one"], [0, 1])))
def test_dropin_file_symbols_against_reference(tmp_path, vartype, mad):
    """The two symbols ctools.py binds, on FITS files, against the reference's own build."""
    if not oracle.have_ref():
        pytest.skip("reference build (oracle/_ref) not shipped")
    from mpdaf_b200.ctools import ctools
    rng = np.random.default_rng(21)
    shape, n = (12, 9, 10), 6
    data, var = random_stack(rng, n, shape, nan_frac=0.1, outlier_frac=0.06)
    files = write_exposures(tmp_path, data, var)
    scales, offsets = np.linspace(0.7, 1.2, n), np.linspace(-1, 1, n)
    want = oracle.combine_ref(files, shape, scales=scales, offsets=offsets, nclip=(2.0, 2.5), vartype=vartype,
                              mad=bool(mad))
    nvox = int(np.prod(shape))
    out, outvar = np.empty(nvox), np.empty(nvox)
    expmap = np.empty(nvox, dtype=np.intc)
    sel, val = np.zeros(n, dtype=np.intc), np.zeros(n, dtype=np.intc)
    names = "\n".join(files).encode()
    typ = {"propagate": 0, "stat_mean": 1, "stat_one": 2}[vartype]
    rc = ctools.mpdaf_merging_sigma_clipping(ctypes.c_char_p(names), out, outvar, expmap, scales, offsets, sel, val,
                                             2, 2.0, 2.5, 2, typ, mad)
    assert rc == 0
    assert names == "\n".join(files).encode()
    assert_same_result(dict(data=out, var=outvar, expmap=expmap, select_pix=sel, valid_pix=val), want)
    wantm = oracle.median_ref(files, shape)
    val = np.zeros(n, dtype=np.intc)
    rc = ctools.mpdaf_merging_median(ctypes.c_char_p(names), out, expmap, val)
    assert rc == 0
    assert_same_result(dict(data=out, expmap=expmap, valid_pix=val), wantm, median=True)


# ---- certified decisions: the exact fallback must catch everything the bounds cannot prove -------

def test_bench_like_data_rarely_needs_the_fallback():
    from mpdaf_b200 import synth
    data, var = synth.host_stack(48, (3, 40, 50), seed=1)
    c = cuda()
    c.slow_voxels()
    kw = dict(scales=synth.default_scales(48), offsets=synth.default_offsets(48))
    got = c.combine_cuda(data, var, **kw)
    assert_same_result(got, oracle.combine_port(data, var, **kw))
    assert c.slow_voxels() <= 6000 * 1e-3      # measured ~1e-4 of the voxels (two-outlier stacks near a bound)
    assert (got["select_pix"] < got["valid_pix"]).any()


@pytest.mark.parametrize("n", [6, 16, 48])
def test_quantised_data_with_bounds_on_samples(n):
    """Small integers, heavy ties, clip bounds that land exactly on sample values: the band test
    fires and the exact path must reproduce the reference's strict inequalities."""
    rng = np.random.default_rng(n)
    shape = (4, 16, 16)
    data = [rng.integers(0, 4, size=shape).astype(np.float32) for _ in range(n)]
    for d in data:
        d[rng.random(shape) < 0.05] = np.nan
        d[rng.random(shape) < 0.03] += 8
    var = [np.ones(shape, dtype=np.float32) for _ in range(n)]
    c = cuda()
    c.slow_voxels()
    total = 0
    for nclip in (1.0, 2.0, (1.0, 2.0), 0.5):
        for vartype in ("propagate", "stat_one"):
            want = oracle.combine_port(data, var, nclip=nclip, vartype=vartype, nmax=3)
            got = c.combine_cuda(data, var, nclip=nclip, vartype=vartype, nmax=3)
            assert_same_result(got, want)
        total += c.slow_voxels()
    assert total > 0


def test_ill_conditioned_and_extreme_values():
    rng = np.random.default_rng(3)
    shape = (2, 12, 12)
    n = 12
    c = cuda()
    # mean >> sigma: relative spread 1e-6 (the moment variance cannot be trusted)
    data = [(1.0e6 + rng.standard_normal(shape)).astype(np.float32) for _ in range(n)]
    var = [np.ones(shape, dtype=np.float32)] * n
    assert_same_result(c.combine_cuda(data, var, nclip=2.0), oracle.combine_port(data, var, nclip=2.0))
    # the pivot (first exposure) is a gross outlier
    data = [(10 + rng.standard_normal(shape)).astype(np.float32) for _ in range(n)]
    data[0] = data[0] + np.float32(1.0e7)
    assert_same_result(c.combine_cuda(data, var, nclip=3.0, vartype="stat_mean"),
                       oracle.combine_port(data, var, nclip=3.0, vartype="stat_mean"))
    # scaled values beyond the float32 range (keys saturate)
    scales = np.full(n, 1.0e35)
    data = [(1.0e4 * (1 + 0.1 * rng.standard_normal(shape))).astype(np.float32) for _ in range(n)]
    data[3][0, 0, :] = -3.0e4
    assert_same_result(c.combine_cuda(data, var, scales=scales, nclip=2.0, vartype="stat_one"),
                       oracle.combine_port(data, var, scales=scales, nclip=2.0, vartype="stat_one"))
    # denormal-range values and exact zeros
    data = [(1.0e-42 * rng.integers(-5, 6, size=shape)).astype(np.float32) for _ in range(n)]
    assert_same_result(c.combine_cuda(data, var, nclip=1.5, vartype="stat_one"),
                       oracle.combine_port(data, var, nclip=1.5, vartype="stat_one"))
    # every sample identical except one: sigma of the survivors is exactly 0 after the first clip
    data = [np.full(shape, 5.0, dtype=np.float32) for _ in range(n)]
    data[7][:] = 100.0
    assert_same_result(c.combine_cuda(data, var, nclip=2.0, vartype="stat_one"),
                       oracle.combine_port(data, var, nclip=2.0, vartype="stat_one"), exact_data=True)


def test_mad_is_bit_exact(monkeypatch):
    """The MAD variant of the warp kernel runs the reference algorithm on exact sorted values: everything identical."""
    monkeypatch.setenv("MPDAF_B200_NO_PAIR", "1")
    rng = np.random.default_rng(77)
    c = cuda()
    for n, ties in ((7, False), (16, True), (48, False), (64, True), (100, False)):
        data, var = random_stack(rng, n, (3, 9, 11), nan_frac=0.15, outlier_frac=0.08, ties=ties)
        for kw in (dict(), dict(scales=np.linspace(0.7, 1.4, n), offsets=np.linspace(-1, 1, n))):
            for vartype in ("propagate", "stat_one"):
                for nclip in (2.0, (1.0, 3.0)):
                    want = oracle.combine_port(data, var, mad=True, nclip=nclip, vartype=vartype, nmax=3, **kw)
                    got = c.combine_cuda(data, var, mad=True, nclip=nclip, vartype=vartype, nmax=3, **kw)
                    assert "mad" in c.kernel_name()
                    assert_same_result(got, want, exact_data=True)
                    if vartype == "stat_one":
                        np.testing.assert_array_equal(got["var"], want["var"])


def test_mad_in_the_pair_kernel():
    """mad=True with var='propagate' on pitched stacks up to 48 deep runs in the pair kernel: sigma = 1.4826 * MAD is
    taken from the sorted keys with its uncertainty folded into the clip bounds' margin, decisions (expmap, counters)
    stay identical to the reference, means within a few ulp, propagated variances within the float32-chain bound."""
    rng = np.random.default_rng(78)
    c = cuda()
    for n, ties in ((7, False), (12, True), (24, False), (48, False), (48, True)):
        data, var = random_stack(rng, n, (5, 13, 17), nan_frac=0.12, outlier_frac=0.06, ties=ties)
        # (offsets not symmetric about 0: integer data with symmetric offsets has voxels whose mean cancels to rounding
        # noise, where a relative tolerance means nothing)
        for kw in (dict(), dict(scales=np.linspace(0.7, 1.4, n), offsets=np.linspace(-1, 1.3, n))):
            for nclip, nmax in ((2.0, 3), ((1.0, 3.0), 2), (5.0, 2), (3.0, 0)):
                want = oracle.combine_port(data, var, mad=True, nclip=nclip, vartype="propagate", nmax=nmax, **kw)
                got = c.combine_cuda(data, var, mad=True, nclip=nclip, vartype="propagate", nmax=nmax,
                                     location="device_pitched", **kw)
                assert "sigclip_pair<" in c.kernel_name() and "mad" in c.kernel_name()
                assert_same_result(got, want)
                np.testing.assert_allclose(got["data"], want["data"], rtol=1e-12, atol=0, equal_nan=True)
    # the statistical variances need the exact MAD: they stay on the warp kernel
    got = c.combine_cuda(data, var, mad=True, nclip=2.0, vartype="stat_one", location="device_pitched")
    assert "sigclip_warp<" in c.kernel_name()
    assert_same_result(got, oracle.combine_port(data, var, mad=True, nclip=2.0, vartype="stat_one"), exact_data=True)


def test_random_fuzz_against_oracle():
    """Many small random configurations (depth, NaN rate, outliers, parameters, scales)."""
    rng = np.random.default_rng(2025)
    c = cuda()
    for trial in range(60):
        n = int(rng.integers(2, 129))
        shape = (2, int(rng.integers(3, 9)), int(rng.integers(3, 40)))
        data, var = random_stack(rng, n, shape, nan_frac=float(rng.uniform(0, 0.4)),
                                 outlier_frac=float(rng.uniform(0, 0.15)), ties=bool(trial % 5 == 0))
        kw = dict(nmax=int(rng.integers(0, 5)), nclip=(float(rng.uniform(0.8, 5)), float(rng.uniform(0.8, 5))),
                  nstop=int(rng.integers(1, 5)), vartype=["propagate", "stat_mean", "stat_one"][trial % 3],
                  mad=bool(trial % 4 == 3))
        if trial % 2:
            kw.update(scales=rng.uniform(0.5, 2.0, n), offsets=rng.uniform(-2, 2, n))
        assert_same_result(c.combine_cuda(data, var, **kw), oracle.combine_port(data, var, **kw))


# ---- size-independent properties at MUSE plane size (too large for the oracle) --------------------

def _device_run(stack, n, nvox, scales=None, offsets=None, nclip=5.0, vartype=0, median=False):
    import torch
    from mpdaf_b200.ctools import DEVICE, check, ctools
    from mpdaf_b200.cubelist import _ptr_table
    dev = torch.device("cuda", 0)
    out = torch.empty(nvox, dtype=torch.float64, device=dev)
    outvar = torch.empty(nvox, dtype=torch.float64, device=dev)
    expmap = torch.empty(nvox, dtype=torch.int32, device=dev)
    kd, dptr = _ptr_table(stack.data_ptrs)
    kv, vptr = _ptr_table(stack.var_ptrs)
    valid, select = np.zeros(n, dtype=np.intc), np.zeros(n, dtype=np.intc)
    stream = torch.cuda.current_stream(dev).cuda_stream
    if median:
        rc = ctools.mpdaf_b200_median_arrays(n, dptr, 4, nvox, DEVICE, 0, out.data_ptr(), expmap.data_ptr(), valid, stream)
    else:
        sp = None if scales is None else scales.ctypes.data
        op = None if offsets is None else offsets.ctypes.data
        rc = ctools.mpdaf_b200_sigma_clipping_arrays(n, dptr, vptr, 4, nvox, DEVICE, 0, out.data_ptr(), outvar.data_ptr(),
                                                     expmap.data_ptr(), sp, op, select, valid, 2, nclip, nclip, 2, vartype, 0, stream)
    check(rc, "device run")
    return out, outvar, expmap, valid, select


def test_properties_at_muse_plane_size():
    """48 exposures x 64 MUSE planes (6.9 Mvoxel): checksums, exact power-of-two scaling,
    permutation of the exposure order."""
    import torch
    from mpdaf_b200 import synth
    from mpdaf_b200.cubelist import DeviceStack
    n, shape = 48, (64, 326, 331)
    nvox = int(np.prod(shape))
    stack, tdata, tvar = synth.device_stack(n, shape, seed=3)
    scales, offsets = synth.default_scales(n), synth.default_offsets(n)
    out, outvar, expmap, valid, select = _device_run(stack, n, nvox, scales, offsets)
    # checksums: every kept sample is counted once in expmap and once in selected_pix
    assert int(expmap.sum(dtype=torch.int64)) == int(select.astype(np.int64).sum())
    for i in (0, 17, 47):
        assert int(valid[i]) == nvox - int(torch.isnan(tdata[i]).sum())
    assert int((expmap == 0).sum()) == int(torch.isnan(out).sum()) == int(torch.isnan(outvar).sum())
    assert 0 < int(valid.astype(np.int64).sum() - select.astype(np.int64).sum()) < 0.02 * n * nvox
    # scaling every exposure by 4 (exact in binary floating point): data x4, var x16, same decisions
    out4, outvar4, expmap4, valid4, select4 = _device_run(stack, n, nvox, scales * 4.0, offsets)
    assert torch.equal(expmap4, expmap) and np.array_equal(select4, select)
    assert torch.equal(torch.nan_to_num(out4), torch.nan_to_num(out * 4.0))
    assert torch.equal(torch.nan_to_num(outvar4), torch.nan_to_num(outvar * 16.0))
    # exposure order reversed: same kept sets (counters permuted), means equal to rounding
    rev = DeviceStack(stack.data_ptrs[::-1], stack.var_ptrs[::-1], shape, 4, 0, keepalive=stack.keepalive)
    outr, outvarr, expmapr, validr, selectr = _device_run(rev, n, nvox, scales[::-1].copy(), offsets[::-1].copy())
    assert torch.equal(expmapr, expmap)
    assert np.array_equal(validr[::-1], valid) and np.array_equal(selectr[::-1], select)
    good = ~torch.isnan(out)
    assert float(((outr[good] - out[good]).abs() / out[good].abs()).max()) < 1e-12
    # median: independent of the exposure order, bit for bit
    m1 = _device_run(stack, n, nvox, median=True)
    m2 = _device_run(rev, n, nvox, median=True)
    assert torch.equal(torch.nan_to_num(m1[0]), torch.nan_to_num(m2[0])) and torch.equal(m1[2], m2[2])
    assert np.array_equal(m1[3], m2[3][::-1])


@pytest.mark.parametrize("n", [129, 150, 192, 256, 257, 300, 448, 500])
def test_multi_run_kernel_for_very_deep_stacks(n):
    """sigclip_runs (R sorted runs of 64, k-th of R sorted columns by bisection) for 129..512 exposures."""
    rng = np.random.default_rng(n)
    c = cuda()
    data, var = random_stack(rng, n, (2, 5, 23), nan_frac=0.25, outlier_frac=0.08, ties=(n % 2 == 0))
    for i, d in enumerate(data):
        d[0, 0, :3] = np.nan                         # empty voxels
        if i > 0:
            d[0, 1, 0] = np.nan                      # a single valid sample
        d[0, 2, 0] = 7.25                            # zero spread
        if i < n // 2:
            d[1, 0, 0] = np.nan                      # a whole run (and more) invalid
    data[0][0, 1, 0] = 3.5
    for kw in (dict(), dict(scales=np.linspace(0.6, 1.7, n), offsets=np.linspace(-2, 2, n))):
        for vartype, nclip, nmax, nstop in (("propagate", 2.0, 2, 2), ("stat_mean", (1.5, 3.0), 6, 2), ("stat_one", 1.0, 30, n // 2)):
            want = oracle.combine_port(data, var, vartype=vartype, nclip=nclip, nmax=nmax, nstop=nstop, **kw)
            got = c.combine_cuda(data, var, vartype=vartype, nclip=nclip, nmax=nmax, nstop=nstop, **kw)
            assert "sigclip_runs" in c.kernel_name()
            assert_same_result(got, want)


@pytest.mark.parametrize("n", [4, 16, 24, 48, 64, 96, 128])
def test_equally_spaced_and_scattered_rows_agree(n):
    """The STRIDED kernels (rows of one [N][V] allocation) and the pointer-table kernels are the same arithmetic."""
    rng = np.random.default_rng(900 + n)
    c = cuda()
    data, var = random_stack(rng, n, (3, 11, 37), nan_frac=0.15, outlier_frac=0.06)
    for kw in (dict(), dict(scales=np.linspace(0.8, 1.3, n), offsets=np.linspace(-1, 1, n))):
        for vartype in ("propagate", "stat_mean"):
            want = oracle.combine_port(data, var, vartype=vartype, nclip=2.5, **kw)
            a = c.combine_cuda(data, var, vartype=vartype, nclip=2.5, location="device_stacked", **kw)
            assert "strided" in c.kernel_name() or "sigclip_warp<" not in c.kernel_name()
            b = c.combine_cuda(data, var, vartype=vartype, nclip=2.5, location="device_scattered", **kw)
            assert "strided" not in c.kernel_name()
            assert_same_result(a, want)
            assert_same_result(b, want)
            for key in ("data", "var", "expmap", "select_pix", "valid_pix"):
                np.testing.assert_array_equal(a[key], b[key])


@pytest.mark.parametrize("n", [3, 12, 48])
def test_host_rows_in_one_array_use_2d_copies(n, monkeypatch):
    """HOST location with equally spaced rows (one [N][V] array: a single 2-D copy per chunk) gives the same
    result as separate arrays (one copy per exposure), also when the stack takes several chunks."""
    rng = np.random.default_rng(200 + n)
    data, var = random_stack(rng, n, (9, 17, 23), nan_frac=0.1, outlier_frac=0.05)   # 3519 voxels
    kw = dict(scales=np.linspace(0.9, 1.1, n), offsets=np.linspace(-0.1, 0.1, n), nclip=3.0)
    want = oracle.combine_port(data, var, **kw)
    for budget in (None, str(1024 * 2 * n * 4)):           # default chunking, then 1024-voxel chunks
        if budget is not None:
            monkeypatch.setenv("MPDAF_B200_SLAB_BYTES", budget)
        a = cuda().combine_cuda(data, var, location="host", **kw)
        b = cuda().combine_cuda(data, var, location="host_stacked", **kw)
        assert_same_result(a, want)
        assert_same_result(b, want)
        for key in ("data", "var", "expmap", "select_pix", "valid_pix"):
            np.testing.assert_array_equal(a[key], b[key])


# ---------------------------------------------------------------------------------------
# parity margins (round-1 review): non-finite samples, adversarial STAT, MUSE-plane-size oracle parity
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [12, 48, 96])
@pytest.mark.parametrize("location", ["device_scattered", "device_pitched"])
def test_infinite_samples(n, location, family):
    """+/-inf in two or three exposures (never exposure 0) of a fifth of the voxels.  The reference gives such a voxel
    a non-finite mean, sigma = NaN and keeps every sample (([1, inf, 2]) -> inf, nan, 3); the keys of
    non-finite samples must never address memory (scattered allocations: a stray index leaves the allocation)."""
    rng = np.random.default_rng(900 + n)
    shape = (3, 11, 13)
    data, var = random_stack(rng, n, shape, nan_frac=0.08, outlier_frac=0.03)
    hit = rng.random(shape) < 0.2
    for k in range(3):
        e = rng.integers(1, n, size=shape)
        sign = rng.choice(np.array([-np.inf, np.inf], dtype=np.float32), size=shape)
        for i in range(1, n):
            m = hit & (e == i) & ((k < 2) | (rng.random(shape) < 0.5))
            data[i][m] = sign[m]
    finite = np.ones(shape, dtype=bool)
    for d in data:
        finite &= ~np.isinf(d)
    finite = finite.reshape(-1)
    assert (~finite).sum() > 20
    kw = dict(scales=np.linspace(0.9, 1.1, n), offsets=np.linspace(-0.1, 0.1, n))
    for vartype in ("stat_mean", "propagate"):
        want = oracle.combine_port(data, var, vartype=vartype, nclip=3.0, **kw)
        got = cuda().combine_cuda(data, var, vartype=vartype, nclip=3.0, location=location, **kw)
        np.testing.assert_array_equal(got["expmap"], want["expmap"])
        np.testing.assert_array_equal(got["valid_pix"], want["valid_pix"])
        np.testing.assert_array_equal(got["select_pix"], want["select_pix"])
        np.testing.assert_allclose(got["data"][finite], want["data"][finite], rtol=1e-6, atol=0, equal_nan=True)
        np.testing.assert_allclose(got["var"][finite], want["var"][finite], rtol=1e-6, atol=0, equal_nan=True)
        # non-finite voxels: the same non-finite values (inf of the same sign, or NaN)
        np.testing.assert_array_equal(got["data"][~finite], want["data"][~finite])
        # MAD: the reference itself is unpinned on non-finite input (NaN sort keys); the finite voxels and the
        # valid counters must still be right and nothing may fault
        wantm = oracle.combine_port(data, var, vartype=vartype, nclip=3.0, mad=True, **kw)
        gotm = cuda().combine_cuda(data, var, vartype=vartype, nclip=3.0, mad=True, location=location, **kw)
        np.testing.assert_array_equal(gotm["valid_pix"], wantm["valid_pix"])
        np.testing.assert_array_equal(gotm["expmap"][finite], wantm["expmap"][finite])
        np.testing.assert_allclose(gotm["data"][finite], wantm["data"][finite], rtol=1e-12, atol=0, equal_nan=True)


@pytest.mark.parametrize("n", [48, 96, 128])
@pytest.mark.parametrize("kind", ["wide", "negative", "nan"])
def test_propagated_variance_margins(n, kind, family):
    """var = 'propagate' is accumulated in float32 chains (<= 4 * 2^-24 relative in the pair kernel, <= (N / NACC + 2) *
    2^-24 in the warp kernel) -- a bound that needs non-negative terms.  STAT spanning twelve decades keeps it; a
    negative STAT sample sends the voxel to float64 products; NaN STAT gives NaN exactly where the reference does."""
    rng = np.random.default_rng(3000 + n)
    shape = (3, 12, 16)
    data, _ = random_stack(rng, n, shape, nan_frac=0.06, outlier_frac=0.04)
    var = [np.exp(rng.uniform(np.log(1e-6), np.log(1e6), size=shape)).astype(np.float32) for _ in range(n)]
    if kind == "negative":
        for v in var:
            v[rng.random(shape) < 0.05] *= np.float32(-1.0)
    if kind == "nan":
        for v in var:
            v[rng.random(shape) < 0.01] = np.nan
        for d, v in zip(data, var):
            v[np.isnan(d)] = np.float32(-np.nan)      # what a masked STAT usually holds; the reference never reads it
    kw = dict(scales=np.linspace(0.5, 2.0, n), offsets=np.linspace(-1, 1, n), nclip=3.0, vartype="propagate")
    want = oracle.combine_port(data, var, **kw)
    for location in ("device_pitched", "device"):
        got = cuda().combine_cuda(data, var, location=location, **kw)
        assert_same_result(got, want)


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4"])
def test_muse_planes_against_the_reference(cfg, tmp_path):
    """Two full MUSE planes (2 x 326 x 331) of BASELINE.json's configurations 2-4 on the synthetic exposures bench.py
    uses, against the reference's own OpenMP loop (oracle/_ref on FITS files; the port when it was not built)."""
    from mpdaf_b200 import synth
    n = {"cfg2": 12, "cfg3": 48, "cfg4": 96}[cfg]
    shape = (2, 326, 331)
    data, var = synth.host_stack(n, shape, seed=1, with_var=(cfg == "cfg3"))
    scales, offsets = synth.default_scales(n), synth.default_offsets(n)
    if oracle.have_ref():
        oracle.set_ref_threads(1)      # fewer threads than planes: the reference double-counts its counters otherwise
    if cfg == "cfg2":
        want = oracle.median_port(data)
        if oracle.have_ref():
            files = write_exposures(tmp_path, data)
            ref = oracle.median_ref(files, shape)
            np.testing.assert_array_equal(np.nan_to_num(ref["data"]), np.nan_to_num(want["data"]))
            want = ref
        for location in ("device", "host"):
            assert_same_result(cuda().median_cuda(data, location=location), want, median=True)
        return
    kw = dict(scales=scales, offsets=offsets, nclip=5.0, nmax=2, nstop=2, vartype="propagate" if cfg == "cfg3" else "stat_mean")
    if oracle.have_ref():
        files = write_exposures(tmp_path, data, var)
        want = oracle.combine_ref(files, shape, **kw)
    else:
        want = oracle.combine_port(data, var, **kw)
    for location in ("device_pitched", "host"):
        got = cuda().combine_cuda(data, var, location=location, **kw)
        assert_same_result(got, want)
    name = cuda().kernel_name()
    assert ("sigclip_pair<48" in name) if cfg == "cfg3" else ("sigclip_duo<96" in name)


# ---------------------------------------------------------------------------------------
# wavelength slabs over several devices inside one call; resources kept between calls
# ---------------------------------------------------------------------------------------
def test_all_devices_equals_one_device(monkeypatch):
    """A HOST-location call with device = ALL_DEVICES deals its chunks to every usable device (one device: the same
    code path with a single entry) and must equal the single-device call bit for bit, counters included."""
    from mpdaf_b200.ctools import ALL_DEVICES, ctools
    monkeypatch.setenv("MPDAF_B200_SLAB_BYTES", str(1 << 20))      # many chunks, so every device gets several
    rng = np.random.default_rng(31)
    n = 12
    data, var = random_stack(rng, n, (40, 30, 31))
    kw = dict(scales=np.linspace(0.9, 1.1, n), offsets=np.linspace(-0.1, 0.1, n), nclip=3.0)
    one = cuda().combine_cuda(data, var, location="host", device=0, **kw)
    every = cuda().combine_cuda(data, var, location="host", device=ALL_DEVICES, **kw)
    assert_same_result(one, oracle.combine_port(data, var, **kw))
    for key in ("data", "var", "expmap", "select_pix", "valid_pix"):
        np.testing.assert_array_equal(one[key], every[key])
    m1 = cuda().median_cuda(data, location="host", device=0)
    m2 = cuda().median_cuda(data, location="host", device=ALL_DEVICES)
    for key in ("data", "expmap", "valid_pix"):
        np.testing.assert_array_equal(m1[key], m2[key])
    ndev = ctools.mpdaf_b200_device_count()
    if ndev >= 2:
        # an explicit subset, and the caller's current device is left as it was
        import ctypes
        import torch
        ids = (ctypes.c_int * 2)(1, 0)
        assert ctools.mpdaf_b200_set_devices(2, ids) == 0
        torch.cuda.set_device(1)
        two = cuda().combine_cuda(data, var, location="host", device=ALL_DEVICES, **kw)
        assert torch.cuda.current_device() == 1
        torch.cuda.set_device(0)
        assert ctools.mpdaf_b200_set_devices(0, None) == 0
        for key in ("data", "var", "expmap", "select_pix", "valid_pix"):
            np.testing.assert_array_equal(one[key], two[key])


def test_release_frees_and_the_next_call_allocates_again():
    import torch
    from mpdaf_b200.ctools import ctools
    rng = np.random.default_rng(32)
    data, var = random_stack(rng, 8, (10, 30, 31))
    want = oracle.combine_port(data, var, nclip=3.0)
    assert_same_result(cuda().combine_cuda(data, var, nclip=3.0, location="host"), want)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    assert ctools.mpdaf_b200_release() == 0
    free1, _ = torch.cuda.mem_get_info(0)
    assert free1 > free0                                   # the staging slots went back to the driver
    assert_same_result(cuda().combine_cuda(data, var, nclip=3.0, location="host"), want)
    assert_same_result(cuda().combine_cuda(data, var, nclip=3.0, location="device_pitched"), want)


def test_outputs_stay_inside_their_buffers(family):
    """Canary margins around every output of a DEVICE-location call (compute-sanitizer is closed on this pool): ragged
    voxel counts (not a multiple of the 64-voxel tile, odd, tiny) must leave the words before and after untouched."""
    import torch
    from mpdaf_b200.ctools import DEVICE, check, ctools
    from mpdaf_b200.cubelist import _ptr_table
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(99)
    n = 12
    for nvox in (1, 63, 65, 127, 1001, 4097):
        pitch = (nvox + 31) // 32 * 32
        d = torch.full((n, pitch), float("nan"), dtype=torch.float32, device=dev)
        v = torch.full((n, pitch), float("nan"), dtype=torch.float32, device=dev)
        d[:, :nvox] = torch.from_numpy((10 + rng.standard_normal((n, nvox))).astype(np.float32)).to(dev)
        v[:, :nvox] = 1.0
        pad = 64
        out = torch.full((nvox + 2 * pad,), -7.0, dtype=torch.float64, device=dev)
        outvar = torch.full((nvox + 2 * pad,), -7.0, dtype=torch.float64, device=dev)
        expmap = torch.full((nvox + 2 * pad,), -7, dtype=torch.int32, device=dev)
        kd, dptr = _ptr_table([d[i].data_ptr() for i in range(n)])
        kv, vptr = _ptr_table([v[i].data_ptr() for i in range(n)])
        valid, select = np.zeros(n, dtype=np.intc), np.zeros(n, dtype=np.intc)
        rc = ctools.mpdaf_b200_sigma_clipping_arrays(n, dptr, vptr, 4, nvox, DEVICE, 0, out[pad:].data_ptr(), outvar[pad:].data_ptr(),
                                                     expmap[pad:].data_ptr(), None, None, select, valid, 2, 3.0, 3.0, 2, 0, 0,
                                                     torch.cuda.current_stream(dev).cuda_stream)
        check(rc, "canary run")
        torch.cuda.synchronize()
        for t in (out, outvar, expmap):
            assert bool((t[:pad] == -7).all()) and bool((t[pad + nvox:] == -7).all()), (nvox, cuda().kernel_name())
            assert not bool((t[pad:pad + nvox] == -7).any())
        assert (valid == nvox).all()


def test_slow_queue_overflow_is_recomputed_in_place(monkeypatch):
    """Voxels the pair kernel cannot certify go to a queue for the exact kernel; when the queue is full the main kernel
    recomputes them itself.  Heavy ties (integer data) queue most voxels; a 5-entry queue overflows at once."""
    rng = np.random.default_rng(77)
    n = 24
    data, var = random_stack(rng, n, (6, 13, 17), nan_frac=0.1, outlier_frac=0.05, ties=True)
    kw = dict(scales=np.linspace(0.9, 1.1, n), offsets=np.linspace(-0.1, 0.1, n), nclip=1.5, nmax=5)
    want = oracle.combine_port(data, var, **kw)
    for cap in ("5", "0", None):
        if cap is None:
            monkeypatch.delenv("MPDAF_B200_SLOW_CAP", raising=False)
        else:
            monkeypatch.setenv("MPDAF_B200_SLOW_CAP", cap)
        cuda().slow_voxels()
        got = cuda().combine_cuda(data, var, location="device_pitched", **kw)
        assert "sigclip_pair<24" in cuda().kernel_name()
        assert_same_result(got, want)
        assert cuda().slow_voxels() > 100
    # the same routes of the two-lane kernel (65 .. 128 exposures): queue, and in-place recomputation when it is full
    for n in (96, 128):
        data, var = random_stack(rng, n, (3, 9, 17), nan_frac=0.1, outlier_frac=0.05, ties=True)
        kw = dict(scales=np.linspace(0.9, 1.1, n), offsets=np.linspace(-0.1, 0.1, n), nclip=1.5, nmax=5)
        want = oracle.combine_port(data, var, **kw)
        for cap in ("3", "0", None):
            if cap is None:
                monkeypatch.delenv("MPDAF_B200_SLOW_CAP", raising=False)
            else:
                monkeypatch.setenv("MPDAF_B200_SLOW_CAP", cap)
            cuda().slow_voxels()
            got = cuda().combine_cuda(data, var, location="device_pitched", **kw)
            assert "sigclip_duo<" in cuda().kernel_name()
            assert_same_result(got, want)
            assert cuda().slow_voxels() > 20
