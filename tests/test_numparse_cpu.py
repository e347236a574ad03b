"""CPU: K1's numeric-literal parser (csrc/numparse.h, compiled for the host) equals strtod
(Python float(): correctly rounded) bit for bit -- the value htslib would store, before its
float32 cast -- on fixed corner cases and random literals."""
import ctypes as C

import numpy as np

from variantcalling_b200 import lib

CORNERS = ["0", "0.0", "1", "16777217", "0.1", "1e5", "1E-3", "1.5e+2", "+2.5", "7.", "000123.4500", "33.333333333",
           "1e22", "1e23", "123456789012345", "0.000001234", "9007199254740992", "9007199254740993", "3.4028234e38",
           "1e-5", "4.35", "2.675", "0.30000000000000004", "1.0000001192092896", "8388608.5", "0.1234567890123456789",
           "123456789.123456789e-5", "1e-30", "1e-45", "2.2250738585072014e-308", "4.9e-324", "2.47e-324", "2.5e-324",
           "17976931348623157e292", "1.8e308", "1.17549435e-38", "7.00649232e-46", "5e-324", "-0.0", "-1.5e-7",
           "18446744073709551615", "0.1e1", "1e0", "100e-2", "1e400", "1e-400"]


def parse(L, s):
    f, d, n = C.c_float(), C.c_double(), C.c_int()
    st = L.ugvc_test_parse_float(s.encode(), C.byref(f), C.byref(d), C.byref(n))
    return st, d.value, f.value, n.value


def test_corner_literals():
    L = lib.load_library()
    for s in CORNERS:
        st, v, f32, used = parse(L, s)
        assert st == 0 and used == len(s), s
        assert v == float(s), (s, v, float(s))
        with np.errstate(over="ignore"):
            assert f32 == np.float32(float(s)) or (np.isinf(f32) and np.isinf(np.float32(float(s)))), s
    assert parse(L, ".")[0] == 1 and parse(L, ".\t")[3] == 1          # missing value
    assert parse(L, "abc")[0] == 2 and parse(L, "1e")[0] == 2 and parse(L, "-.")[0] == 2
    st, v, _, used = parse(L, "nan;")
    assert st == 0 and v != v and used == 3
    assert parse(L, "-inf,")[1] == float("-inf") and parse(L, "Infinity\t")[3] == 8
    assert parse(L, "12,5")[3] == 2 and parse(L, "3.5;X")[1] == 3.5   # stops at the delimiter


def test_random_literals_equal_strtod():
    L = lib.load_library()
    rng = np.random.default_rng(2024)
    n_bad = 0
    for _ in range(60000):
        digits = int(rng.integers(1, 22))
        m = "".join(str(d) for d in rng.integers(0, 10, size=digits))
        point = int(rng.integers(0, digits + 1))
        s = (m[:point] or "0") + ("." + m[point:] if point < digits else "")
        r = rng.random()
        if r < 0.35:
            s += "e" + str(int(rng.integers(-340, 320)))
        elif r < 0.6:
            s += "E" + str(int(rng.integers(-30, 30)))
        st, v, _, used = parse(L, s)
        if st == 2:  # > 19 significant digits whose tail decides the rounding: refused, never approximated
            assert digits > 19
            n_bad += 1
            continue
        assert used == len(s) and v == float(s), (s, v, float(s))
    assert n_bad < 200
