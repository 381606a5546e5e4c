"""Host-side unit tests of the decimal / fixed-point headers against Python's decimal (libmpdec).

The headers under phanotate_b200/csrc are __host__ __device__; tests/native/dec_harness.cpp compiles
them with g++ so that the arithmetic the CUDA kernels run can be checked digit for digit here, on the
CPU box.  The harness is test infrastructure: the product never loads it.
"""
import ctypes
import os
import random
import subprocess
from decimal import Decimal, getcontext, localcontext

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "dec_harness.cpp")
SO = os.path.join(HERE, "native", "dec_harness.so")

TDEC = np.dtype([("lo", "<u8"), ("hi", "<u8"), ("e", "<i4"), ("neg", "<i4")])


@pytest.fixture(scope="module")
def lib():
    deps = [SRC] + [os.path.join(HERE, "..", "phanotate_b200", "csrc", f)
                    for f in ("wide.cuh", "dec.cuh", "fxpow.cuh", "frepr.cuh", "tables.inc", "hold.cuh", "score.cuh", "pipeline.cuh",
                              "dec2double.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", SO, SRC])
    return ctypes.CDLL(SO)


def pack(vals):
    a = np.zeros(len(vals), dtype=TDEC)
    for i, d in enumerate(vals):
        s, digits, e = Decimal(d).as_tuple()
        c = int("".join(map(str, digits)))
        a[i] = (c & (2 ** 64 - 1), c >> 64, e, s)
    return a


def unpack(a):
    out = []
    for r in a:
        c = int(r["lo"]) | (int(r["hi"]) << 64)
        out.append(Decimal((int(r["neg"]), tuple(map(int, str(c))), int(r["e"]))))
    return out


def same(a: Decimal, b: Decimal):
    return a.as_tuple() == b.as_tuple()


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def rand_dec(rng, maxd=28):
    nd = rng.choice([1, 2, 3, 5, 14, 27, 28, maxd])
    c = rng.randrange(0, 10 ** nd)
    if rng.random() < 0.15:
        c = c - c % (10 ** rng.randrange(0, nd + 1))     # trailing zeros
    if rng.random() < 0.05:
        c = 10 ** rng.randrange(0, nd + 1)               # powers of ten / rounding-carry cases
    if rng.random() < 0.05:
        c = 10 ** nd - 1
    e = rng.choice([0, 0, -1, -5, -27, -28, -29, -30, -40, -56, -84, 3, 12, 25])
    return Decimal((rng.random() < 0.3, tuple(map(int, str(c))), e))


def binop(lib, op, xs, ys, prec=28):
    a, b = pack(xs), pack(ys)
    o = np.zeros(len(xs), dtype=TDEC)
    lib.t_binop(op, len(xs), P(a), P(b), prec, P(o))
    return unpack(o)


@pytest.mark.parametrize("op,fn", [(0, lambda x, y: x + y), (1, lambda x, y: x - y),
                                    (2, lambda x, y: x * y), (3, lambda x, y: x / y)])
def test_binops_match_decimal(lib, op, fn):
    rng = random.Random(1234 + op)
    xs = [rand_dec(rng) for _ in range(40000)]
    ys = [rand_dec(rng) for _ in range(40000)]
    if op == 3:
        ys = [y if y != 0 else Decimal(7) for y in ys]
    got = binop(lib, op, xs, ys)
    getcontext().prec = 28
    for x, y, g in zip(xs, ys, got):
        assert same(fn(x, y), g), (op, x, y, fn(x, y), g)


def test_path_specific_identities(lib):
    # the mixed int/Decimal forms the reference uses (functions.py:26-46,140,174-178; orfs.py:168-173)
    p = Decimal("0.05520932534679849777229484991")
    cases = [(1, Decimal(1), p), (1, Decimal(1), Decimal("0E-84")), (3, Decimal(1), Decimal("0.05")),
             (0, Decimal(0), p), (3, Decimal(3), Decimal(6)), (3, Decimal(24), Decimal(96)),
             (0, Decimal("1.000000000000000000000000000"), Decimal(20)), (3, Decimal(0), Decimal(93)),
             (2, Decimal("0E-28"), p), (0, Decimal("0E-56"), Decimal("0E-57"))]
    for op, x, y in cases:
        want = [lambda: x + y, lambda: x - y, lambda: x * y, lambda: x / y][op]()
        assert same(want, binop(lib, op, [x], [y])[0]), (op, x, y)


def test_integer_power_follows_libmpdec(lib):
    rng = random.Random(99)
    xs, ns = [], []
    for _ in range(20000):
        p = Decimal(rng.randrange(1, 15 * 10 ** 26)) / Decimal(10 ** 28)
        xs.append(1 - p)
        ns.append(rng.choice([1, 2, 3, 4, 7, 33, 99, 100, 101, 255, 256, 333, 499, 502]))
    xs += [Decimal(1), Decimal("1.000"), Decimal("1.000000000000000000000000000"), Decimal("0.953125"), Decimal("0.5")]
    ns += [5, 3, 100, 0, 10]
    a = pack(xs)
    nn = np.array(ns, dtype=np.uint32)
    o = np.zeros(len(xs), dtype=TDEC)
    lib.t_powi(len(xs), P(a), P(nn), 28, P(o))
    for x, n, g in zip(xs, ns, unpack(o)):
        assert same(x ** Decimal(n), g), (x, n, x ** Decimal(n), g)


def test_real_power_decimal_exponent(lib):
    # stage E: ((1-pstop)**pos_max[i])**pos_min[j]  (functions.py:293,298)
    rng = random.Random(5)
    xs, ys = [], []
    for _ in range(30000):
        kind = rng.random()
        if kind < 0.6:
            x = 1 - Decimal(rng.randrange(1, 15 * 10 ** 26)) / Decimal(10 ** 28)
        elif kind < 0.8:
            x = 1 - Decimal(rng.randrange(1, 10 ** 12)) / Decimal(10 ** rng.choice([13, 18, 24, 28]))
        else:
            x = Decimal(rng.randrange(10 ** 27, 10 ** 28)) / Decimal(10 ** 28)   # any value in (0.1, 1)
        cnt = rng.randrange(2, 200000)
        y = Decimal(rng.randrange(1, cnt)) / Decimal(cnt)
        xs.append(+x)
        ys.append(y)
    xs += [Decimal(1), Decimal("1.000000000000000000000000000"), Decimal("0.9999999999999999999999999999")]
    ys += [Decimal("0.5"), Decimal("0.3333333333333333333333333333"), Decimal("0.0001531628120692295910552917752")]
    a, b = pack(xs), pack(ys)
    o = np.zeros(len(xs), dtype=TDEC)
    ok = np.zeros(len(xs), dtype=np.int32)
    lib.t_powr(len(xs), P(a), P(b), 28, P(o), P(ok))
    assert ok.all()
    for x, y, g in zip(xs, ys, unpack(o)):
        assert same(x ** y, g), (x, y, x ** y, g)


def test_real_power_float_exponent(lib):
    # score_gap: g ** Decimal(length/3) with length/3 a Python float (functions.py:41-43)
    rng = random.Random(6)
    xs, ys = [], []
    for _ in range(300):
        g = 1 - Decimal(rng.randrange(10 ** 26, 15 * 10 ** 26)) / Decimal(10 ** 28)
        for length in range(-2, 301):
            if length % 3 == 0:
                continue
            xs.append(g)
            ys.append(length / 3)
    a = pack(xs)
    yv = np.array(ys, dtype=np.float64)
    o = np.zeros(len(xs), dtype=TDEC)
    ok = np.zeros(len(xs), dtype=np.int32)
    lib.t_powd(len(xs), P(a), P(yv), 28, P(o), P(ok))
    assert ok.all()
    for x, y, g in zip(xs, ys, unpack(o)):
        assert same(x ** Decimal(y), g), (x, y)


def test_decimal_of_float_repr(lib):
    # Orf.score: Decimal(str(weight_rbs)) (orfs.py:126)
    rng = random.Random(7)
    vals = [1.0, 100.0, 1e15, 1e16, 1e-4, 1e-5, 0.1, 123456.75, 5e-324 * 0 + 2.5, 1 / 3, 2 / 3, 1e22, 9.999999999999999e22]
    for _ in range(100000):
        n1, n2 = rng.randrange(1, 5000), rng.randrange(28, 10 ** 7)
        m1, m2 = rng.randrange(1, 10 ** 6), rng.randrange(28, 2 * 10 ** 7)
        vals.append((n1 / n2) / (m1 / m2))
    for _ in range(20000):
        vals.append(rng.random() * 10 ** rng.randrange(-9, 12))
    v = np.array(vals, dtype=np.float64)
    o = np.zeros(len(v), dtype=TDEC)
    ok = np.zeros(len(v), dtype=np.int32)
    lib.t_repr(len(v), P(v), P(o), P(ok))
    assert ok.all()
    for x, g in zip(vals, unpack(o)):
        assert same(Decimal(str(x)), g), (x, str(x), g)


def test_milli_integer(lib):
    rng = random.Random(8)
    xs = [rand_dec(rng) for _ in range(5000)] + [Decimal("-6.1E+28"), Decimal("1.2345E+50"), Decimal("-0.0004")]
    a = pack(xs)
    mag = np.zeros((len(xs), 8), dtype=np.uint32)
    ok = np.zeros(len(xs), dtype=np.int32)
    lib.t_milli(len(xs), P(a), P(mag), P(ok))
    for x, m, k in zip(xs, mag, ok):
        want = abs(int((x * 1000).to_integral_value(rounding="ROUND_DOWN")))
        if want < 2 ** 255 and want < 10 ** 72:
            assert k
            got = sum(int(v) << (32 * i) for i, v in enumerate(m))
            with localcontext() as ctx:
                ctx.prec = 100
                want = abs(int((x * 1000).to_integral_value(rounding="ROUND_DOWN")))
            assert got == want, (x, got, want)


def test_fast_hold_multiply_is_exact(lib):
    """hold * factor for 28-digit operands (functions.py:293): the one-product fast path must either
    decide the half-even rounding correctly or hand over to the exact multiply."""
    rng = random.Random(11)
    xs, ys, plain = [], [], []
    for i in range(200000):
        a = rng.randrange(10 ** 27, 10 ** 28)
        b = rng.randrange(10 ** 27, 10 ** 28)
        plain.append(all(i % m for m in (7, 11, 13, 17)))
        if i % 7 == 0:      # products just around 10^55 (27 vs 28 dropped digits)
            b = (10 ** 55 // a) + rng.randrange(-3, 4)
            b = min(max(b, 10 ** 27), 10 ** 28 - 1)
        if i % 11 == 0:     # products whose dropped part is (almost) exactly one half
            k = rng.randrange(10 ** 27, 10 ** 28)
            a = 2 * rng.randrange(5 * 10 ** 26, 5 * 10 ** 27) + 1
            b = ((2 * k + 1) * 10 ** 27 // (2 * a)) + rng.randrange(-1, 2)
            b = min(max(b, 10 ** 27), 10 ** 28 - 1)
        if i % 13 == 0:
            a, b = 10 ** 28 - 1 - rng.randrange(0, 3), 10 ** 28 - 1 - rng.randrange(0, 3)
        if i % 17 == 0:
            a = 10 ** 27 + rng.randrange(0, 3)
        xs.append(Decimal((0, tuple(map(int, str(a))), -28)))
        ys.append(Decimal((0, tuple(map(int, str(b))), rng.choice([-28, -29, -31]))))
    A, Bv = pack(xs), pack(ys)
    o = np.zeros(len(xs), dtype=TDEC)
    ok = np.zeros(len(xs), dtype=np.int32)
    lib.t_hold_fast(len(xs), P(A), P(Bv), P(o), P(ok))
    assert ok[np.array(plain)].mean() > 0.9999      # random operands: the fast path decides practically always
    getcontext().prec = 28
    for x, y, g in zip(xs, ys, unpack(o)):
        assert same(x * y, g), (x, y, x * y, g)


def test_dec_to_double_is_float_of_decimal(lib):
    rng = random.Random(12)
    xs = [rand_dec(rng) for _ in range(20000)]
    xs += [Decimal("-4.827980747824565E+2"), Decimal("-6.1E+28"), Decimal("1E+60"), Decimal("4.9406564584124654E-30")]
    xs = [x for x in xs if x == 0 or -70 < x.adjusted() < 100]
    a = pack(xs)
    o = np.zeros(len(xs), dtype=np.float64)
    ok = np.zeros(len(xs), dtype=np.int32)
    lib.t_to_double(len(xs), P(a), P(o), P(ok))
    for x, g, k in zip(xs, o, ok):
        if k:
            assert float(x) == g and str(float(x)) == str(g), (x, float(x), g)
    assert ok.mean() > 0.9


def _limbs(v):
    return [(v >> (32 * j)) & 0xFFFFFFFF for j in range(7)]


def _unlimbs(a):
    return sum(int(x) << (32 * j) for j, x in enumerate(a))


def test_fixed_point_exp_ln_absolute_accuracy(lib):
    """exp and ln in Q32.192 must be good to ~2^-180 absolute: the single rounding to 28 digits that follows
    (csrc/fxpow.cuh) is then wrong only if exp(y ln x) lies within 1e-50 of a rounding boundary."""
    rng = random.Random(21)
    with localcontext() as ctx:
        ctx.prec = 90
        ONE = 1 << 192
        ts, negs = [], []
        for _ in range(3000):
            mag = rng.choice([1e-20, 1e-9, 1e-4, 0.01, 0.2, 0.7, 3.0, 17.0])
            t = int(Decimal(rng.random() * mag) * ONE)
            ts.append(t)
            negs.append(1 if (mag > 1 or rng.random() < 0.8) else 0)
        tl = np.array([_limbs(t) for t in ts], dtype=np.uint32)
        ng = np.array(negs, dtype=np.int32)
        out = np.zeros_like(tl)
        ok = np.zeros(len(ts), dtype=np.int32)
        lib.t_fx_exp(len(ts), P(tl), P(ng), P(out), P(ok))
        assert ok.all()
        worst = 0
        for t, s, o in zip(ts, negs, out):
            x = Decimal(t) / ONE
            want = (-x if s else x).exp() * ONE
            worst = max(worst, abs(Decimal(_unlimbs(o)) - want) / max(Decimal(1), want / ONE))
        print("exp worst", float(worst))
        assert worst < 2 ** 12, worst            # error below 2^-180 relative to max(1, value)
        xs = []
        for _ in range(3000):
            kind = rng.random()
            x = 1 - rng.random() * rng.choice([1e-12, 1e-6, 0.01, 0.15]) if kind < 0.7 else rng.uniform(0.05, 1.9)
            xs.append(int(Decimal(x) * ONE))
        xl = np.array([_limbs(x) for x in xs], dtype=np.uint32)
        out = np.zeros_like(xl)
        neg = np.zeros(len(xs), dtype=np.int32)
        lib.t_fx_ln(len(xs), P(xl), P(out), P(neg), P(ok))
        assert ok.all()
        worst = 0
        for x, s, o in zip(xs, neg, out):
            want = (Decimal(x) / ONE).ln() * ONE
            got = Decimal(_unlimbs(o)) * (-1 if s else 1)
            worst = max(worst, abs(got - want))
        assert worst < 2 ** 12, worst


def test_ln_of_rounded_power_shortcut(lib):
    """ln(round28(exp(T))) = T + (a-V)/V must match the true ln(a) to ~2^-180 (csrc/fxpow.cuh)."""
    rng = random.Random(22)
    with localcontext() as ctx:
        ctx.prec = 90
        ONE = 1 << 192
        ts = [int(Decimal(rng.random() * rng.choice([1e-12, 1e-6, 1e-3, 0.05, 0.17])) * ONE) for _ in range(3000)]
        tl = np.array([_limbs(t) for t in ts], dtype=np.uint32)
        a = np.zeros(len(ts), dtype=TDEC)
        out = np.zeros_like(tl)
        neg = np.zeros(len(ts), dtype=np.int32)
        ok = np.zeros(len(ts), dtype=np.int32)
        lib.t_ln_rounded(len(ts), P(tl), P(a), P(out), P(neg), P(ok))
        assert ok.all()
        worst = 0
        for av, s, o in zip(unpack(a), neg, out):
            want = av.ln() * ONE
            got = Decimal(_unlimbs(o)) * (-1 if s else 1)
            worst = max(worst, abs(got - want))
        print("ln-of-rounded worst", float(worst))
        assert worst < 2 ** 12, worst


def test_small_integer_division_fast_path(lib):
    """count / Decimal(len(seq)) (orfs.py:169-172) through the one-limb division: exact quotients keep the
    ideal exponent, everything else is rounded half-even to 28 digits."""
    rng = random.Random(31)
    A = [rng.randrange(0, 6000) for _ in range(60000)] + [0, 1, 24, 3, 5, 1, 4095, 1000, 999999, 2 ** 31 - 1]
    Bv = [rng.choice([rng.randrange(1, 6000), 96, 100, 125, 128, 300, 1024, 3000]) for _ in range(60000)] + \
         [93, 3, 96, 6, 2, 7, 4096, 8, 1000000, 2 ** 31 - 1]
    a = np.array(A, dtype=np.uint32)
    b = np.array(Bv, dtype=np.uint32)
    o = np.zeros(len(A), dtype=TDEC)
    lib.t_div_u32(len(A), P(a), P(b), P(o))
    getcontext().prec = 28
    for x, y, g in zip(A, Bv, unpack(o)):
        assert same(x / Decimal(y), g), (x, y, x / Decimal(y), g)
