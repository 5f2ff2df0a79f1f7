"""Pins the oracle: the reference's own known-answer tests for the hot path
(tests/golden/reference_kats.json, transcribed from sum.lisp / maximum.lisp / minimum.lisp /
mean.lisp / variance.lisp / n-arg-fn-tests.lisp) must come out of oracle/ exactly."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))


def to_np(v, dtype):
    def conv(x):
        if isinstance(x, str):
            return float(Fraction(x))
        if isinstance(x, list):
            return [conv(y) for y in x]
        return x
    return np.array(conv(v), dtype=dtype)


@pytest.mark.parametrize("kat", KATS["reductions"], ids=lambda k: "%s-axes%s-%s" % (k["fn"], k["axes"], k["keep_dims"]))
def test_reduction_kats(kat, refnu):
    for et in kat["element_types"]:
        dt = refnu.ELEMENT_TYPES[et][1]
        x = to_np(kat["x"], dt)
        axes = kat["axes"]
        fn = kat["fn"]
        if fn in ("sum", "maximum", "minimum"):
            got = refnu.reduce(fn, x, axes, kat["keep_dims"])
        elif fn == "mean":
            got = refnu.mean(x, axes, kat["keep_dims"])
        else:
            got = refnu.variance(x, axes, kat["keep_dims"])
        want = to_np(kat["expect"], dt)
        got = np.asarray(got)
        assert got.shape == want.shape, (et, got.shape, want.shape)
        assert np.array_equal(got, want), (kat["source"], et, got, want)


@pytest.mark.parametrize("kat", KATS["nary"], ids=lambda k: "%s-%d" % (k["op"], len(k["args"])))
def test_nary_kats(kat, refnu):
    op = {"+": "add", "-": "sub", "*": "mul", "/": "div"}[kat["op"]]
    args = [np.asarray(a, dtype=np.float32) for a in kat["args"]]
    if len(args) == 1:
        ident = np.asarray({"add": 0, "sub": 0, "mul": 1, "div": 1}[op], dtype=np.float32)
        got = refnu.two_arg(op, ident, args[0]) if op in ("sub", "div") else args[0]
    else:
        got = refnu.n_ary(op, *args)
    want = to_np(kat["expect"], np.float32)
    assert np.array_equal(np.asarray(got), want), (kat["source"], got, want)
