"""The reference's elementwise test ladder (define-numericals-one-arg-test /
define-numericals-two-arg-test, src/basic-math/test-definition-macros.lisp:5-105, 276-421) replayed
through the device API: simple vector; (2,N) and (N,2) shapes; strided views incl. NEGATIVE steps;
a strided view broadcast into a (10 45 45) OUT; the 4-D broadcast case; "in-place only touches the
view".  Expectations come from the oracle walking the same views on the host with one BMAS call
per innermost row, which is what the reference does.  Bit-exact (arithmetic, compare) or <= 1 ULP
(transcendentals)."""
import numpy as np
import pytest

from _helpers import ETYPE_OF, NP_OF, rnd, ulp_distance

pytestmark = pytest.mark.gpu

OPS2 = [("add", "add"), ("subtract", "sub"), ("multiply", "mul"), ("divide", "div"), ("two_arg_max", "max"),
        ("two_arg_min", "min")]
CMPS = [("two_arg_lt", "lt"), ("two_arg_le", "le"), ("two_arg_eq", "eq"), ("two_arg_ne", "neq"),
        ("two_arg_ge", "ge"), ("two_arg_gt", "gt")]


def same(a, b):
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        return np.array_equal(na, nb) and np.where(na, 0, a).tobytes() == np.where(nb, 0, b).tobytes()
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


def view_pairs(nu, host, index):
    """(device view, host view) of the same storage through the same subscripts"""
    dev = nu.asarray(host, type=host.dtype)
    return dev[index], host[index]


LADDER = [  # (shape, index for x, index for y)  -- strides, negative steps, offsets
    ((200,), np.s_[:], np.s_[:]),
    ((2, 1000), np.s_[:, :], np.s_[:, :]),
    ((1000, 2), np.s_[:, :], np.s_[:, :]),
    ((100, 100), np.s_[10:50, 20:60], np.s_[40:80, 0:40]),          # offset views, inner contiguous
    ((100, 100), np.s_[0:100:2, 0:100:5], np.s_[0:50, 10:90:4]),     # strided both axes
    ((100, 100), np.s_[::-2, ::-1], np.s_[0:50, :]),                 # negative steps
    ((64, 7, 33), np.s_[::3, :, ::-2], np.s_[0:22, :, 0:17]),
]


@pytest.mark.parametrize("prefix", ["s", "d", "i64", "i32", "i16", "i8", "u64", "u32", "u16", "u8"])
def test_two_arg_ladder(prefix, nu, refnu):
    rng = np.random.default_rng(21)
    dt = NP_OF[prefix]
    is_f = prefix in ("s", "d")
    for shape, ix, iy in LADDER:
        hx, hy = rnd(rng, dt, shape), rnd(rng, dt, shape)
        dx, vx = view_pairs(nu, hx, ix)
        dy, vy = view_pairs(nu, hy, iy)
        for api, op in OPS2:
            if op == "div" and not is_f:
                with pytest.raises(nu.NoCFunction):
                    getattr(nu, api)(dx, dy)
                continue
            got = getattr(nu, api)(dx, dy).to_numpy()
            assert same(got, refnu.two_arg(op, vx, vy)), (prefix, api, shape)
        for api, op in CMPS:
            got = getattr(nu, api)(dx, dy)
            assert got.element_type == "(unsigned-byte 8)"
            assert same(got.to_numpy(), refnu.compare(op, vx, vy)), (prefix, api, shape)


def test_out_strided_only_touches_the_view(nu, refnu):
    """test-definition-macros.lisp:86-99"""
    rng = np.random.default_rng(22)
    hx = rnd(rng, np.float32, (100, 100))
    hy = rnd(rng, np.float32, (100, 100))
    ho = rnd(rng, np.float32, (100, 100))
    before = ho.copy()
    do = nu.asarray(ho)
    out_view = do[10:90:2, 5:95:3]
    x = nu.asarray(hx)[10:90:2, 5:95:3]
    y = nu.asarray(hy)[::-2, 0:30][0:40, :]
    r = nu.multiply(x, y, out=out_view)
    assert r is out_view
    want = before.copy()
    want[10:90:2, 5:95:3] = hx[10:90:2, 5:95:3] * hy[::-2, 0:30][0:40, :]
    assert same(do.to_numpy(), want)
    # in-place on a strided view: (nu:add! view other)
    z = nu.asarray(hx)
    nu.add_(z[::3, ::-1], nu.asarray(hy)[::3, :])
    want = hx.copy()
    want[::3, ::-1] = hx[::3, ::-1] + hy[::3, :]
    assert same(z.to_numpy(), want)


def test_broadcast_cases(nu, refnu):
    rng = np.random.default_rng(23)
    # strided (45 45) view -> (10 45 45) OUT   (test-definition-macros.lisp:76-85)
    h = rnd(rng, np.float32, (100, 100))
    v_dev, v_host = view_pairs(nu, h, np.s_[5:95:2, 10:55])
    out = nu.zeros(10, 45, 45)
    nu.fround(v_dev, out=out)
    assert same(out.to_numpy(), np.broadcast_to(np.rint(v_host), (10, 45, 45)))
    # 4-D: (10 1 10) op (10 45 10) -> (2 10 45 10) OUT   (test-definition-macros.lisp:398-421)
    a, b = rnd(rng, np.float64, (10, 1, 10)), rnd(rng, np.float64, (10, 45, 10))
    out = nu.zeros(2, 10, 45, 10, type="double-float")
    nu.subtract(nu.asarray(a), nu.asarray(b), out=out)
    assert same(out.to_numpy(), np.broadcast_to(a - b, (2, 10, 45, 10)))
    # BASELINE C3 shapes, scaled: column + row, matrix * row vector, with OUT preallocated
    col, row = rnd(rng, np.float32, (1024, 1)), rnd(rng, np.float32, (1, 1024))
    out = nu.empty((1024, 1024), "single-float")
    nu.add(nu.asarray(col), nu.asarray(row), out=out)
    assert same(out.to_numpy(), refnu.two_arg("add", col, row))
    A, v = rnd(rng, np.float32, (512, 768)), rnd(rng, np.float32, (768,))
    out = nu.empty((512, 768), "single-float")
    nu.multiply(nu.asarray(A), nu.asarray(v), out=out)
    assert same(out.to_numpy(), A * v)
    # ragged inner extent (not a multiple of the vector width) and an inner-broadcast operand
    A, c = rnd(rng, np.float32, (37, 1001)), rnd(rng, np.float32, (37, 1))
    assert same(nu.add(nu.asarray(A), nu.asarray(c)).to_numpy(), A + c)
    # incompatible shapes signal the reference's condition
    with pytest.raises(nu.IncompatibleBroadcastDimensions):
        nu.add(nu.zeros(3), nu.zeros(4))
    with pytest.raises(nu.IncompatibleBroadcastDimensions):
        nu.add(nu.zeros(3, 1), nu.zeros(1, 3), broadcast=False)


def test_scalars_types_and_promotion(nu, refnu):
    rng = np.random.default_rng(24)
    h = rnd(rng, np.float32, (33, 65))
    x = nu.asarray(h)
    assert same(nu.multiply(x, 3).to_numpy(), h * np.float32(3))
    assert same(nu.subtract(2.5, x).to_numpy(), np.float32(2.5) - h)
    assert same(nu.two_arg_lt(x, 0.5).to_numpy(), (h < np.float32(0.5)).astype(np.uint8))
    assert nu.add(1, 2) == 3 and nu.two_arg_lt(1, 2) == 1              # number op number: host arithmetic
    # mixed types without OUT -> max-type (two-arg-fn-non-comparison.lisp:72-88)
    hi = rnd(rng, np.int32, (33, 65))
    r = nu.add(x, nu.asarray(hi))
    assert r.element_type == "single-float" and same(r.to_numpy(), h + hi.astype(np.float32))
    u8, s8 = rnd(rng, np.uint8, 100), rnd(rng, np.int8, 100)
    # the reference's table only has casts INTO floats, so u8 + s8 (max-type (signed-byte 16)) has no cast
    with pytest.raises(nu.NukError):
        nu.add(nu.asarray(u8), nu.asarray(s8))
    # with OUT -> OUT's type (two-arg-fn-non-comparison.lisp:90-105)
    out = nu.zeros(33, 65, type="double-float")
    nu.add(x, nu.asarray(hi), out=out)
    assert same(out.to_numpy(), h.astype(np.float64) + hi.astype(np.float64))
    # unary float op on an integer array: cast to *default-float-format* first (one-arg-fn-float.lisp:303-320)
    r = nu.fround(nu.asarray(hi))
    assert r.element_type == "single-float" and same(r.to_numpy(), np.rint(hi.astype(np.float32)))
    # rank-0 and empty arrays
    z = nu.add(nu.asarray(np.float32(2)), nu.asarray(np.float32(3)))
    assert z.dimensions == () and z.item() == 5
    e = nu.add(nu.zeros(0, 5), nu.zeros(0, 5))
    assert e.dimensions == (0, 5)


def test_nary_kats_and_folds(nu):
    import json
    import os
    from fractions import Fraction
    kats = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))
    fns = {"+": nu.plus, "-": nu.minus, "*": nu.times, "/": nu.slash}
    for k in kats["nary"]:
        args = [a if isinstance(a, int) else nu.asarray(a) for a in k["args"]]
        got = fns[k["op"]](*args)
        want = np.array([float(Fraction(str(v))) for v in k["expect"]], dtype=np.float32)
        assert same(got.to_numpy(), want), k
        if len(args) >= 2 and not any(isinstance(a, int) for a in args):
            out = nu.zeros(3)
            assert fns[k["op"]](*args, out=out) is out and same(out.to_numpy(), want), k
    x = nu.asarray([1, 5, 3])
    assert same(nu.lt(0, x, 4).to_numpy(), np.array([1, 0, 1], dtype=np.uint8))      # chained with logand
    assert same(nu.max(x, nu.asarray([2, 2, 9])).to_numpy(), np.array([2, 5, 9], dtype=np.float32))


@pytest.mark.parametrize("fn,lo,hi", [("sin", 0, 1), ("cos", 0, 1), ("tan", 0, 1), ("asin", 0, 1), ("acos", 0, 1),
                                      ("atan", 0, 1), ("sinh", 0, 1), ("cosh", 0, 1), ("tanh", 0, 1),
                                      ("asinh", 0, 1), ("acosh", 1, 2), ("atanh", 0, 1), ("exp", 0, 1),
                                      ("log", 0, 1), ("sqrt", 0, 1)])
def test_one_arg_float_ladder(fn, lo, hi, nu, refnu):
    """the reference's own domains ([0,1), acosh [1,2)) and tolerances (2f-7 / 1d-15 relative), plus
    the north-star's <= 1 ULP against the oracle, on simple / strided / broadcast views"""
    rng = np.random.default_rng(25)
    for dt, tol in ((np.float32, 2e-7), (np.float64, 1e-15)):
        for shape, ix, _ in LADDER[:6]:
            h = rnd(rng, dt, shape, lo, hi)
            if fn == "log":
                h = np.maximum(h, dt(1e-3))
            dv, hv = view_pairs(nu, h, ix)
            got = getattr(nu, fn)(dv).to_numpy()
            want = refnu.one_arg(fn, np.ascontiguousarray(hv))
            assert ulp_distance(got.reshape(-1), want.reshape(-1)).max() <= 1, (fn, dt, shape)
            truth = getattr(np, {"asin": "arcsin", "acos": "arccos", "atan": "arctan", "asinh": "arcsinh",
                                 "acosh": "arccosh", "atanh": "arctanh"}.get(fn, fn))(hv.astype(np.float64))
            ok = np.isfinite(truth) & (np.abs(truth) > 0)
            rel = np.abs(got[ok].astype(np.float64) - truth[ok]) / np.abs(truth[ok])
            assert rel.max() < tol * (4 if fn in ("acos", "log", "acosh") and dt == np.float64 else 1), (fn, dt, rel.max())
        # in place: (nu:sin! x) only touches the view
        h = rnd(rng, dt, (50, 60), lo, hi)
        if fn == "log":
            h = np.maximum(h, dt(1e-3))
        d = nu.asarray(h)
        getattr(nu, fn)(d[::2, ::-3], out=d[::2, ::-3], broadcast=False)
        want = h.copy()
        want[::2, ::-3] = refnu.one_arg(fn, np.ascontiguousarray(h[::2, ::-3]))
        got = d.to_numpy()
        mask = np.zeros_like(h, dtype=bool)
        mask[::2, ::-3] = True
        assert same(got[~mask], h[~mask])
        assert ulp_distance(got[mask], want[mask]).max() <= 1


def test_expt_atan2_ladder(nu, refnu):
    rng = np.random.default_rng(26)
    for dt in (np.float32, np.float64):
        for shape, ix, iy in LADDER[:6]:
            hx, hy = rnd(rng, dt, shape, 0, 10), rnd(rng, dt, shape, 0, 10)     # reference domain [0, 10)
            dx, vx = view_pairs(nu, hx, ix)
            dy, vy = view_pairs(nu, hy, iy)
            for api, op in (("expt", "pow"), ("atan2", "atan2")):
                got = getattr(nu, api)(dx, dy).to_numpy()
                want = refnu.two_arg(op, np.ascontiguousarray(vx), np.ascontiguousarray(vy))
                assert ulp_distance(got.reshape(-1), want.reshape(-1)).max() <= 1, (api, dt, shape)
    x = nu.asarray(rnd(rng, np.float32, 100, 1, 10))
    assert np.allclose(nu.log(x, 2.0).to_numpy(), np.log2(x.to_numpy()), rtol=3e-7)


@pytest.mark.parametrize("prefix,ncols", [("s", 1024), ("s", 4100), ("d", 512), ("d", 2050), ("i16", 2048), ("u8", 4096), ("i64", 1000)])
def test_contiguous_out_with_broadcast_operands_flat2(prefix, ncols, nu, libnuk, refnu):
    """k_map_flat2 (OUT one contiguous block; operands laid out like OUT, row vector, column vector or scalar --
    the BASELINE C3 shapes): every operand-kind pair, full tiles + the ragged last tile, in-place, compare and
    cast outputs; bit-exact against the oracle's per-row calls, and identical to the row-looping kernel."""
    rng = np.random.default_rng(31)
    dt = NP_OF[prefix]
    et = ETYPE_OF[prefix]
    is_f = prefix in ("s", "d")
    nrows = 37
    full = rnd(rng, dt, (nrows, ncols))
    kinds = {"full": full, "row": rnd(rng, dt, (1, ncols)), "vec": rnd(rng, dt, (ncols,)), "col": rnd(rng, dt, (nrows, 1)),
             "one": rnd(rng, dt, (1, 1))}
    dev = {k: nu.asarray(v, type=et) for k, v in kinds.items()}
    for kx in kinds:
        for ky in kinds:
            if kx != "full" and ky != "full" and not ({kx, ky} & {"row", "vec"} and {kx, ky} & {"col"}):
                continue                                           # the result must be (nrows, ncols)
            for api, op in (("multiply", "mul"), ("two_arg_max", "max")) + ((("divide", "div"),) if is_f else ()):
                out = nu.empty((nrows, ncols), et)
                getattr(nu, api)(dev[kx], dev[ky], out=out)
                want = refnu.two_arg(op, kinds[kx], kinds[ky])
                assert same(out.to_numpy(), want), (prefix, ncols, kx, ky, api)
                libnuk.nuk_set_option(b"rows_rpc", 3)              # the same call through k_map_rows
                try:
                    out2 = nu.empty((nrows, ncols), et)
                    getattr(nu, api)(dev[kx], dev[ky], out=out2)
                finally:
                    libnuk.nuk_set_option(b"rows_rpc", 0)
                assert same(out2.to_numpy(), want), (prefix, ncols, kx, ky, api, "rows kernel")
            c = nu.two_arg_lt(dev[kx], dev[ky])
            assert same(c.to_numpy(), refnu.compare("lt", kinds[kx], kinds[ky])), (prefix, ncols, kx, ky)
    # scalar polymorphs (host scalar by value) and in-place on the full operand
    s = kinds["one"][0, 0].item()
    assert same(nu.multiply(dev["full"], s).to_numpy(), refnu.two_arg("mul", full, np.full((1,), s, dtype=dt)))
    z = nu.asarray(full, type=et)
    nu.multiply(z, dev["vec"], out=z)
    assert same(z.to_numpy(), refnu.two_arg("mul", full, kinds["vec"]))
    if not is_f:
        assert same(nu.astype(dev["full"], "double-float").to_numpy(), full.astype(np.float64))
    else:
        nu.sqrt(nu.abs(dev["full"]), out=z)
        assert same(z.to_numpy(), np.sqrt(np.abs(full)))
    # 3-D that coalesces to the same problem: (3, 7, ncols) * (ncols)
    t3 = rnd(rng, dt, (3, 7, ncols))
    got = nu.multiply(nu.asarray(t3, type=et), dev["vec"]).to_numpy()
    assert same(got, refnu.two_arg("mul", t3, kinds["vec"]))


def test_flat2_writes_only_its_block(nu, refnu):
    """OUT is a contiguous block INSIDE a larger array (rows 3..40 of a matrix): k_map_flat2 must leave the rows
    before and after it alone -- tile remainders, the last ragged tile and the column-vector prefetch included
    (compute-sanitizer is closed on this GPU pool; whole-buffer comparisons stand in for memcheck)."""
    rng = np.random.default_rng(33)
    for dt, et, ncols in ((np.float32, "single-float", 1028), (np.float64, "double-float", 1030), (np.uint8, "(unsigned-byte 8)", 4112)):
        big = rnd(rng, dt, (50, ncols))
        m = rnd(rng, dt, (50, ncols))
        v, c = rnd(rng, dt, (ncols,)), rnd(rng, dt, (50, 1))
        dbig, dm = nu.asarray(big, type=et), nu.asarray(m, type=et)
        nu.multiply(dm[3:40], nu.asarray(v, type=et), out=dbig[3:40])
        want = big.copy()
        want[3:40] = refnu.two_arg("mul", m[3:40], v)
        assert same(dbig.to_numpy(), want), (et, "row vector")
        nu.two_arg_max(dm[3:40], nu.asarray(c, type=et)[3:40], out=dbig[3:40])
        want[3:40] = refnu.two_arg("max", np.ascontiguousarray(m[3:40]), np.ascontiguousarray(c[3:40]))
        assert same(dbig.to_numpy(), want), (et, "column vector")
