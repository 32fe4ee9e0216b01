"""CPU test of the PLY ingest's literal parser (ndt-net_b200/csrc/ply_decimal.cuh) against CPython's float()/int(),
which is what the reference reader calls per token (/root/reference/ndnet/datasets/CARLA_Seg.py:120-123).
The header is compiled for the host with g++ (no GPU needed)."""
import ctypes as C
import math
import os
import random
import struct
import subprocess
from decimal import Decimal, getcontext

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "ply_decimal_host.cpp")
OUT = os.path.join(HERE, "native", "_build", "libply_decimal_host.so")


@pytest.fixture(scope="module")
def host():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(HERE, "..", "ndt-net_b200", "csrc", "ply_decimal.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT])
    L = C.CDLL(OUT)
    L.ply_host_parse_float.argtypes = [C.c_char_p, C.c_long, C.POINTER(C.c_double)]
    L.ply_host_parse_int.argtypes = [C.c_char_p, C.c_long, C.POINTER(C.c_longlong)]
    L.ply_host_parse_many.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
    L.ply_host_parse_line.argtypes = [C.c_char_p, C.c_long, C.c_void_p, C.POINTER(C.c_longlong)]
    return L


def parse_many(L, toks):
    buf = "".join(toks).encode()
    off = np.zeros(len(toks) + 1, np.int64)
    off[1:] = np.cumsum([len(t) for t in toks])
    out = np.zeros(len(toks), np.float64)
    st = np.zeros(len(toks), np.int32)
    L.ply_host_parse_many(buf, off.ctypes.data, len(toks), out.ctypes.data, st.ctypes.data)
    return out, st


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def check_all(L, toks):
    out, st = parse_many(L, toks)
    for t, v, s in zip(toks, out, st):
        assert s == 0, (t, s)
        assert bits(float(v)) == bits(float(t)), (t, float(v), float(t))


def test_plain_decimals_and_reprs(host):
    rng = random.Random(0)
    toks = ["0", "-0", "0.0", "+1", "1.", ".5", "-.25", "1e0", "1E+2", "1.5e-3", "000123.4500", "0.000", "-0.0e10",
            "12345678901234567", "9007199254740993", "9007199254740992", "0.30000000000000004", "1e22", "1e23", "1e-22",
            "1e-23", "8.41e21", "9999999999999999999", "1.7976931348623157e55", "4.9e-54", "123456789012345678e-37"]
    for _ in range(20000):
        x = rng.uniform(-200, 200)
        toks.append(repr(x))                       # 15-17 significant digits: mostly the exact long-division path
        toks.append(f"{x:.6f}")                    # what LiDAR exporters write: the one-operation fast path
        toks.append(f"{x:.3e}")
        toks.append(repr(np.float32(x).item()))
    for _ in range(5000):
        e = rng.randint(-40, 40)
        toks.append(f"{rng.randint(1, 10**19 - 1)}e{e}")
        toks.append(repr(rng.uniform(1, 10) * 10.0 ** rng.randint(-36, 36)))
    check_all(host, toks)


def test_literals_next_to_rounding_boundaries(host):
    """19-digit literals just below / above the midpoint of two adjacent doubles: the cases a float-arithmetic parser
    gets wrong."""
    getcontext().prec = 60
    rng = random.Random(1)
    toks = []
    for _ in range(4000):
        d = rng.uniform(1, 2) * 2.0 ** rng.randint(-100, 100)
        mid = (Decimal(d) + Decimal(math.nextafter(d, math.inf))) / 2
        for q in ("ROUND_FLOOR", "ROUND_CEILING"):
            getcontext().rounding = q
            t = "{:.18e}".format(mid)                     # 19 significant digits, rounded down / up
            m, e = t.split("e")
            if abs(int(e)) + 18 <= 55:
                toks.append(m.replace(".", "") + "e" + str(int(e) - 18))
    check_all(host, toks)


def test_float32_after_double_matches_numpy(host):
    """The reader keeps np.asarray(points) (float64) -> torch .float(): double rounding, reproduced by rounding the
    correctly rounded double to float32."""
    rng = random.Random(2)
    toks = [f"{rng.uniform(-100, 100):.9f}" for _ in range(20000)]
    out, st = parse_many(host, toks)
    assert not st.any()
    want = np.array([float(t) for t in toks], np.float64).astype(np.float32)
    assert np.array_equal(out.astype(np.float32).view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("tok,code", [("", 2), ("-", 2), (".", 2), ("1e", 2), ("1e+", 2), ("1..2", 2), ("--1", 2), ("1+2", 2),
                                      ("1.2.3", 2), ("e5", 2), ("abc", 4), ("nan", 4), ("inf", 4), ("1_0", 4), ("0x10", 4),
                                      ("12345678901234567891", 4), ("1e56", 4), ("1e-60", 4)])
def test_rejected_float_tokens(host, tok, code):
    v = C.c_double()
    assert host.ply_host_parse_float(tok.encode(), len(tok), C.byref(v)) == code
    if code == 2:
        with pytest.raises(ValueError):
            float(tok)


def test_trailing_zero_digits_beyond_19_are_exact(host):
    check_all(host, ["1234567890123456789000", "0.12345678901234567890000", "100000000000000000000000000000"])


@pytest.mark.parametrize("tok,want", [("0", 0), ("7", 7), ("+12", 12), ("-3", -3), ("0028", 28), ("65536", 65536)])
def test_int_tokens(host, tok, want):
    v = C.c_longlong()
    assert host.ply_host_parse_int(tok.encode(), len(tok), C.byref(v)) == 0
    assert v.value == want == int(tok)


@pytest.mark.parametrize("tok,code", [("", 2), ("1.0", 2), ("1e3", 2), ("-", 2), ("+-1", 2), ("x", 4), ("1_0", 4)])
def test_rejected_int_tokens(host, tok, code):
    v = C.c_longlong()
    assert host.ply_host_parse_int(tok.encode(), len(tok), C.byref(v)) == code


def test_grammar_agrees_with_python_on_random_tokens(host):
    """300 000 random tokens over the alphabet of decimal literals: accepted <=> float() accepts (same bits), rejected with
    the 'bad literal' code <=> float() raises ValueError, 'unsupported' only for literals float() accepts but that lie
    outside what the parser converts exactly (more than 19 significant digits or a decimal exponent beyond +-55)."""
    rng = random.Random(7)
    alphabet = "0123456789" * 3 + "+-..eE"
    toks = ["".join(rng.choice(alphabet) for _ in range(rng.randint(1, 14))) for _ in range(300_000)]
    out, st = parse_many(host, toks)
    n_ok = n_bad = n_unsup = 0
    for t, v, s in zip(toks, out, st):
        try:
            want = float(t)
            ok = True
        except ValueError:
            ok = False
        if s == 0:
            assert ok and bits(float(v)) == bits(want), (t, float(v))
            n_ok += 1
        elif s == 2:
            assert not ok, t
            n_bad += 1
        else:
            assert s == 4 and ok, (t, s)
            n_unsup += 1
    assert n_ok > 50_000 and n_bad > 50_000                      # both branches are exercised
    assert n_unsup < n_ok // 2                                   # refusals are the exception (huge exponents like 9e9999)


def test_int_grammar_agrees_with_python_on_random_tokens(host):
    rng = random.Random(8)
    alphabet = "0123456789" * 4 + "+-.e"
    for _ in range(50_000):
        t = "".join(rng.choice(alphabet) for _ in range(rng.randint(1, 9)))
        v = C.c_longlong()
        s = host.ply_host_parse_int(t.encode(), len(t), C.byref(v))
        try:
            want = int(t)
            assert s == 0 and v.value == want, (t, s, v.value)
        except ValueError:
            assert s == 2, (t, s)


def _reference_line(line: str):
    """CARLA_Seg.py:118-123 on one line -> ("ok", x, y, z, tag) or the exception class name."""
    data = line.strip().split()
    try:
        x = float(data[0])
        y = float(data[1])
        z = float(data[2])
        tag = int(data[-1])
    except IndexError:
        return ("IndexError",)
    except ValueError:
        return ("ValueError",)
    return ("ok", x, y, z, tag)


def test_whole_lines_agree_with_the_reference_reader(host):
    """The shared host/device line parser (tokenising on str.split() whitespace, evaluation order of the four conversions)
    against the reference's statements on 100 000 random lines: every ASCII whitespace character Python splits on, empty and
    short lines, bad literals in any position, trailing junk."""
    rng = random.Random(11)
    spaces = [" ", "  ", "\t", " \t ", "\r", "\x0b", "\x0c", "\x1c", "\x1d", "\x1e", "\x1f", "\n"]

    def number():
        k = rng.random()
        if k < 0.55:
            return f"{rng.uniform(-100, 100):.{rng.randint(0, 8)}f}"
        if k < 0.7:
            return repr(rng.uniform(-1e3, 1e3))
        if k < 0.8:
            return str(rng.randint(-50, 50))
        if k < 0.9:
            return f"{rng.uniform(-9, 9):.3e}"
        return rng.choice(["abc", "1.2.3", "--1", "1e", ".", "+", "1_0", "nan", "0x1", "1,5", "-", "e3"])

    xyz = (C.c_double * 3)()
    tag = C.c_longlong()
    seen = {"ok": 0, "IndexError": 0, "ValueError": 0}
    for _ in range(100_000):
        ntok = rng.choice([0, 1, 2, 3, 3, 4, 5, 6, 6, 6, 7])
        toks = [number() for _ in range(ntok)]
        if ntok and rng.random() < 0.8:
            toks[-1] = str(rng.randint(-3, 40)) if rng.random() < 0.9 else rng.choice(["7.0", "x", "1e1", "+4", "-0"])
        line = rng.choice(["", " ", "\t"]) + "".join(t + rng.choice(spaces[:-1]) for t in toks).rstrip("\n") + rng.choice(["\n", "\r\n", ""])
        raw = line.encode("ascii")
        s = host.ply_host_parse_line(raw, len(raw), xyz, C.byref(tag))
        want = _reference_line(line)
        seen[want[0]] += 1
        if want[0] == "ok":
            if s == 4:            # refused, never approximated: only literals float() takes but the parser does not (nan, 1_0)
                assert any(c in line for c in "na_"), line
                continue
            assert s == 0, (line, s)
            assert [bits(xyz[k]) for k in range(3)] == [bits(v) for v in want[1:4]] and tag.value == want[4], line
        elif want[0] == "IndexError":
            # a literal float() takes but the parser refuses (nan, 1_0) in front of the missing token is reported first
            assert s == 1 or (s == 4 and any(c in line for c in "na_")), (line, s)
        else:
            assert s in (2, 4), (line, s)
    assert min(seen.values()) > 5000, seen
