"""Golden vectors produced by the UNMODIFIED reference C code (tests/golden/make_golden.py).

Each fixture is checked twice: against the CPU port (runs everywhere) and against the
CUDA library through its C ABI (`-m gpu`).  Bar: bit-exact expmap / counters / medians and
NaN positions, <= 1e-6 relative on sigma-clipped means and variances (BASELINE.json).
"""
import os

import numpy as np
import pytest

import oracle
from helpers import assert_same_result
from mpdaf_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

IMPLS = [pytest.param("port", id="port"),
         pytest.param("cuda-device", marks=pytest.mark.gpu, id="cuda-device"),
         pytest.param("cuda-host", marks=pytest.mark.gpu, id="cuda-host")]


def run_combine(impl, data, var, **kw):
    if impl == "port":
        return oracle.combine_port(data, var, **kw)
    import cuda_runner
    return cuda_runner.combine_cuda(data, var, location=impl.split("-")[1], **kw)


def run_median(impl, data):
    if impl == "port":
        return oracle.median_port(data)
    import cuda_runner
    return cuda_runner.median_cuda(data, location=impl.split("-")[1])


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def want_of(g, prefix, median=False):
    keys = ("data", "expmap", "valid_pix") if median else ("data", "var", "expmap", "select_pix", "valid_pix")
    return {k: g[prefix + k].astype(np.intc) if g[prefix + k].dtype.kind == "i" else g[prefix + k] for k in keys}


@pytest.mark.parametrize("impl", IMPLS)
def test_cfg1_combine_and_median(impl):
    """BASELINE.json configs[0]: 5 synthetic 40x30x30 cubes with var, sigma-clip nclip=3, median."""
    g = load("cfg1")
    shape = tuple(int(s) for s in g["shape"])
    data, var = synth.host_stack(int(g["n"]), shape, seed=int(g["seed"]))
    got = run_combine(impl, data, var, nmax=2, nclip=float(g["nclip"]), nstop=2, vartype="propagate")
    assert_same_result(got, want_of(g, ""))
    assert (g["expmap"] < 5).any()   # NaN holes; 5 samples cannot be clipped at 3 sigma (|x-med| <= 3 sigma)
    got = run_median(impl, data)
    assert_same_result(got, want_of(g, "med_", median=True), median=True)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("vartype", ["propagate", "stat_mean", "stat_one"])
@pytest.mark.parametrize("mad", [0, 1])
def test_all_modes(impl, vartype, mad):
    g = load("modes")
    data = list(g["in_data"])
    var = list(g["in_var"])
    got = run_combine(impl, data, var, scales=g["scales"], offsets=g["offsets"], nmax=int(g["nmax"]),
                      nclip=tuple(g["nclip"]), nstop=int(g["nstop"]), vartype=vartype, mad=bool(mad))
    assert_same_result(got, want_of(g, f"{vartype}_{mad}_"))


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("nclip", [5, 3])
def test_deep48_scaled(impl, nclip):
    """cfg3-like: 48 exposures, per-exposure scales/offsets, var='propagate'."""
    g = load("deep48")
    shape = tuple(int(s) for s in g["shape"])
    data, var = synth.host_stack(48, shape, seed=int(g["seed"]))
    got = run_combine(impl, data, var, scales=synth.default_scales(48), offsets=synth.default_offsets(48),
                      nmax=2, nclip=float(nclip), nstop=2, vartype="propagate")
    assert_same_result(got, want_of(g, f"nclip{nclip}_"))
    assert (g[f"nclip{nclip}_select_pix"] < g[f"nclip{nclip}_valid_pix"]).any()   # rejections happened


@pytest.mark.parametrize("impl", IMPLS)
def test_n70_generic(impl):
    g = load("n70")
    shape = tuple(int(s) for s in g["shape"])
    data, _ = synth.host_stack(70, shape, seed=int(g["seed"]), with_var=False)
    got = run_combine(impl, data, None, nmax=3, nclip=2.5, nstop=3, vartype="stat_mean")
    assert_same_result(got, want_of(g, ""))


@pytest.mark.parametrize("impl", IMPLS)
def test_parameter_sweep_with_ties(impl):
    g = load("sweep_ties")
    data, var = list(g["in_data"]), list(g["in_var"])
    for c, (nmax, lo, up, nstop) in enumerate(g["combos"]):
        got = run_combine(impl, data, var, nmax=int(nmax), nclip=(float(lo), float(up)), nstop=int(nstop),
                          vartype="stat_one")
        assert_same_result(got, want_of(g, f"c{c}_"))
