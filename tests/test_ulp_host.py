"""CPU: the product math headers compiled for the host, swept against a higher-precision truth and
against SLEEF u10 (tests/ulp/ulp_sweep.cpp).  Pass criterion: max error < 1 ULP (faithful) AND
integer ULP distance to SLEEF u10 <= 1 wherever SLEEF is inside its own bound -- the north-star
tolerance.  The full 2^32-input exhaustive runs are recorded in profiles/ulp_report_r01.jsonl
(tools: `ulp_sweep <fn> f32`); here a 2^27-input window per function keeps the suite fast, and the
GPU tests prove the device executes these same expressions bit for bit."""
import json
import subprocess

import pytest

from _helpers import build_host_tool

F32 = ["exp", "log", "sin", "cos", "tanh", "tan", "asin", "acos", "atan", "sinh", "cosh", "asinh", "acosh", "atanh"]
F64 = F32


@pytest.fixture(scope="module")
def sweep():
    return build_host_tool("ulp_sweep.cpp", "ulp_sweep")


@pytest.mark.parametrize("fn", F32)
def test_f32_window(fn, sweep):
    # bit patterns 0x3c000000..0x44000000: |x| in [2^-7, 512) positive -- dense where the functions live
    out = subprocess.check_output([sweep, fn, "f32", "0x3c000000", "0x44000000"]).decode()
    r = json.loads(out.strip().splitlines()[-1])
    assert r["max_ulp"] < 1.0, r
    assert r["max_ulp_dist_vs_sleef_u10"] <= 1, r


@pytest.mark.parametrize("fn", F64)
def test_f64_samples(fn, sweep):
    out = subprocess.check_output([sweep, fn, "f64", "4000000"]).decode()
    r = json.loads(out.strip().splitlines()[-1])
    assert r["max_ulp"] < 1.0, r
    assert r["max_ulp_dist_vs_sleef_u10"] <= 1, r


@pytest.mark.parametrize("fn,typ", [("pow", "f64"), ("atan2", "f64"), ("pow", "f32b"), ("atan2", "f32b")])
def test_binary_samples(fn, typ, sweep):
    out = subprocess.check_output([sweep, fn, typ, "4000000"]).decode()
    r = json.loads(out.strip().splitlines()[-1])
    assert r["max_ulp"] < 1.0, r
    assert r["max_ulp_dist_vs_sleef_u10"] <= 1, r


def test_f64_main_paths_equal_complete_functions(sweep):
    """Every branch-free main path of the double functions (pack split, region split: asinh / acosh polynomial AND
    logarithm paths) returns the complete function's bits wherever it does not raise `slow` -- so which path a tile
    takes on the device (map.cuh:kRegionSplit votes per warp) can never change a result.  The exhaustive single-float
    counterpart (`ulp_sweep all fast32`, all 2^32 inputs per function) is recorded in profiles/fast_vs_full_r02.jsonl."""
    out = subprocess.check_output([sweep, "all", "fast64", "400000"]).decode()
    rows = [json.loads(l) for l in out.strip().splitlines()]
    assert len(rows) >= 13
    for r in rows:
        assert r["fast_differs_from_full"] == 0, r
        assert r["flagged_slow"] < r["inputs"], r
