"""String.toDouble() (Main.kt:791) as the device computes it — csrc/kq_parse.cuh, integer-only, compiled here with g++ —
against the CPU oracle's javaParseDouble (glibc strtod: correctly rounded) bit for bit, and the oracle against Python's
float()/float.fromhex as a second, independent correctly-rounded implementation."""
import ctypes
import math
import os
import random
import struct
import subprocess
from decimal import Decimal, getcontext
from fractions import Fraction

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def dev_parse(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("parse") / "libparse.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "query-engines_b200", "csrc"), "-o", so,
                    os.path.join(HERE, "parse_harness.cpp")], check=True)
    L = ctypes.CDLL(so)
    L.kq_test_parse.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]

    def parse(s):
        b = s.encode()
        out = ctypes.c_uint64()
        st = L.kq_test_parse(b, len(b), ctypes.byref(out))
        return None if st else out.value
    return parse


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def oracle_bits(oracle, s):
    try:
        return bits(oracle.parse_double(s))
    except oracle.OracleError as e:
        assert e.code == 5          # NumberFormatException
        return None


def test_random_decimals_match_the_oracle_bit_for_bit(dev_parse, oracle):
    rng = random.Random(1)
    n = 0
    for _ in range(120_000):
        nd = rng.randint(1, 40) if rng.random() < 0.9 else rng.randint(40, 900)
        digs = "".join(rng.choice("0123456789") for _ in range(nd))
        if rng.random() < 0.6:
            k = rng.randint(0, nd)
            digs = digs[:k] + "." + digs[k:]
        s = digs
        if rng.random() < 0.7:
            s += rng.choice("eE") + rng.choice(["", "+", "-"]) + str(rng.randint(0, 340))
        if rng.random() < 0.3:
            s = "-" + s
        want = oracle_bits(oracle, s)
        assert dev_parse(s) == want, s
        if want is not None:
            assert want == bits(float(s)), s             # the oracle itself against Python's float()
            n += 1
    assert n > 100_000


def test_halfway_cases_and_their_neighbours(dev_parse, oracle):
    """Exact midpoints between adjacent doubles (up to ~770 significant digits) round to even; one unit in the 800th place
    either side does not — normal, subnormal and near-overflow ranges."""
    getcontext().prec = 1200
    rng = random.Random(2)
    for _ in range(3000):
        e = rng.choice([rng.randint(-1074, -1000), rng.randint(-1022, 1022), rng.randint(-60, 60)])
        x = math.ldexp(1 + rng.getrandbits(52) / 2 ** 52, e) if e > -1023 else math.ldexp(rng.getrandbits(52) + 1, -1074)
        nxt = math.nextafter(x, math.inf)
        if math.isinf(x) or math.isinf(nxt):
            continue
        mid = (Fraction(x) + Fraction(nxt)) / 2
        d = Decimal(mid.numerator) / Decimal(mid.denominator)            # exact: the denominator is a power of two
        for delta in (0, 1, -1):
            s = format(d + Decimal(delta) * Decimal(10) ** (d.adjusted() - 800), "e")
            want = bits(float(s))
            assert dev_parse(s) == want == oracle_bits(oracle, s), (s[:40], len(s))


def test_hex_floats(dev_parse, oracle):
    rng = random.Random(3)
    for _ in range(20_000):
        nh = rng.randint(1, 20)
        h = "".join(rng.choice("0123456789abcdefABCDEF") for _ in range(nh))
        if rng.random() < 0.6:
            k = rng.randint(0, nh)
            h = h[:k] + "." + h[k:]
        s = "0x" + h + rng.choice("pP") + rng.choice(["", "+", "-"]) + str(rng.randint(0, 1100)) + rng.choice(["", "", "d", "F"])
        assert dev_parse(s) == oracle_bits(oracle, s), s


@pytest.mark.parametrize("s,ok", [("1d", 1), ("1f", 1), ("1.5F", 1), ("0x1p3d", 1), ("0x1.8p1", 1), ("0xfp0", 1), ("0x1", 0), ("", 0),
                                  (" 12 ", 1), ("\t-7.25e1\n", 1), ("1e", 0), ("e5", 0), (".", 0), ("1.", 1), (".5", 1), ("NaN", 1),
                                  ("-Infinity", 1), ("infinity", 0), ("nan", 0), ("1_0", 0), ("0x.p1", 0), ("1e400", 1), ("1e-400", 1),
                                  ("4.9e-324", 1), ("2.4703282292062327e-324", 1), ("2.4703282292062328e-324", 1),
                                  ("1.7976931348623159e308", 1), ("+.0e-0D", 1), ("1e23", 1), ("9007199254740993", 1),
                                  ("123456789012345678901234567890", 1), ("0.1e-5000", 1), ("1e5000", 1), ("--1", 0), ("1 2", 0),
                                  ("0x1.fffffffffffff8p1023", 1), ("0x0.0000000000001p-1022", 1), ("0x1p-1075", 1), ("0x1.0000000000001p-1075", 1)])
def test_grammar_and_edges(dev_parse, oracle, s, ok):
    got, want = dev_parse(s), oracle_bits(oracle, s)
    assert (got is not None) == bool(ok)
    assert got == want
