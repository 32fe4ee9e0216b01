"""CPU property test of the voxel-coordinate shortcuts (ndt-net_b200/csrc/ndt_cell.cuh, compiled for the host): both
must give exactly the reference's `(unsigned)floor((p - off) / vs)` (core_legacy/src/voxel.c:89-91) whenever they decide a
point themselves - the claim voxel ids and per-voxel membership rest on.  Points are placed at random, exactly on cell
boundaries, and a few ulps either side of them, for the voxel sizes the bisection visits (ndt.h:38-43: 0.01 .. 30)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "ndt_cell_host.cpp")
OUT = os.path.join(HERE, "native", "_build", "libndt_cell_host.so")


@pytest.fixture(scope="module")
def host():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(HERE, "..", "ndt-net_b200", "csrc", "ndt_cell.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        # x86-64 baseline has no FMA, -ffp-contract=off states it: the device units are built with -fmad=false
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT])
    L = C.CDLL(OUT)
    L.ndt_host_check_exact64.restype = C.c_long
    L.ndt_host_check_exact64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long)]
    L.ndt_host_check_prefilter32.restype = C.c_long
    L.ndt_host_check_prefilter32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    return L


def _voxel_sizes(rng, n):
    """Sizes the search can produce: lo + (hi - lo)/2 chains inside [0.01, 30] plus arbitrary doubles in that range."""
    vs = np.empty(n)
    lo, hi = np.full(n, 0.01), np.full(n, 30.0)
    g = lo + (hi - lo) / 2
    steps = rng.integers(0, 15, n)
    for k in range(15):
        go_hi = rng.random(n) < 0.5
        active = steps > k
        hi = np.where(active & go_hi, g, hi)
        lo = np.where(active & ~go_hi, g, lo)
        g = np.where(active, lo + (hi - lo) / 2, g)
    vs[:] = g
    arb = rng.random(n) < 0.3
    vs[arb] = rng.uniform(0.01, 30.0, arb.sum())
    return vs


def _points_f64(rng, n):
    vs = _voxel_sizes(rng, n)
    off = rng.uniform(-120, 10, n)
    k = rng.integers(0, 4000, n).astype(np.float64)
    kind = rng.integers(0, 4, n)
    p = off + rng.uniform(0, 250, n)                                     # anywhere
    on = off + k * vs                                                    # (about) on a boundary
    p = np.where(kind == 1, on, p)
    ulps = rng.integers(-6, 7, n)
    near = on.copy()
    for _ in range(6):                                                   # a few ulps either side
        near = np.where(ulps > 0, np.nextafter(near, np.inf), np.where(ulps < 0, np.nextafter(near, -np.inf), near))
        ulps = ulps - np.sign(ulps)
    p = np.where(kind >= 2, near, p)
    p = np.maximum(p, off)                                               # off is the minimum of the cloud
    return p, off, vs


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_exact_fp64_shortcut_equals_the_reference_floor(host, seed):
    rng = np.random.default_rng(seed)
    p, off, vs = _points_f64(rng, 4_000_000)
    first = C.c_long(-1)
    bad = host.ndt_host_check_exact64(p.ctypes.data, off.ctypes.data, vs.ctypes.data, len(p), C.byref(first))
    assert bad == 0, (bad, p[first.value], off[first.value], vs[first.value])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fp32_prefilter_never_decides_wrongly(host, seed):
    rng = np.random.default_rng(100 + seed)
    p64, off64, vs = _points_f64(rng, 4_000_000)
    off = off64.astype(np.float32)
    # fp32 inputs: the offset is the minimum of fp32 values; boundary points rounded to fp32 and their fp32 neighbours
    p = (off.astype(np.float64) + (p64 - off64)).astype(np.float32)
    step = rng.integers(-2, 3, len(p))
    p = np.where(step > 0, np.nextafter(p, np.float32(np.inf)), np.where(step < 0, np.nextafter(p, np.float32(-np.inf)), p)).astype(np.float32)
    p = np.maximum(p, off)
    first, decided = C.c_long(-1), C.c_long(0)
    bad = host.ndt_host_check_prefilter32(p.ctypes.data, off.ctypes.data, vs.ctypes.data, len(p), C.byref(first), C.byref(decided))
    assert bad == 0, (bad, float(p[first.value]), float(off[first.value]), vs[first.value])
    assert decided.value > 0.1 * len(p)                  # three quarters of these inputs sit on or next to a boundary
