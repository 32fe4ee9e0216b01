// ndt_cell.cuh — the voxel coordinate of a point along one axis, floor((p - off) / vs) as core_legacy/src/voxel.c:89-91
// computes it (IEEE fp64 subtraction, division, floor), without the division in the common case.  Plain arithmetic shared
// by the device kernels (ndt_device.cuh) and a host build (tests/native/ndt_cell_host.cpp, tests/test_ndt_cell.py), where the
// "same floor as the reference" claim of both shortcuts is checked on tens of millions of points placed on and next to
// cell boundaries.
#pragma once
#include <cmath>
#include <cstring>

#ifdef __CUDACC__
#define NDT_HD __host__ __device__ __forceinline__
#else
#define NDT_HD inline
#endif

namespace ndt {

NDT_HD int float_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i;
    std::memcpy(&i, &f, 4);
    return i;
#endif
}

// Exact path: with rv = RN(1/vs), q0 = RN(a * rv) is within 2 ulp of the IEEE quotient, so whenever q0 is further than
// 2^-48 (relative) from an integer both have the same floor; only the rare near-integer cases take the real division.
NDT_HD unsigned cell_exact64(double p, double off, double vs, double rv) {
    const double a = p - off;
    const double q0 = a * rv;
    const double f = floor(q0);
    const double t = q0 - f;                                  // exact (Sterbenz / small integers)
    const double eps = q0 * 3.5527136788005009e-15;           // 2^-48 * q0  (>= 16 ulp)
    // a >= 0 (off is the minimum), so the quotient cannot fall below cell 0: no lower-boundary ambiguity there
    if ((t > eps || f == 0.0) && (1.0 - t) > eps) return (unsigned)f;
    return (unsigned)floor(a / vs);
}

// fp32 prefilter for fp32 inputs: the offset is the minimum of fp32 values, hence itself an fp32 value, so
// q32 = RN32(RN32(p - off) * RN32(1/vs)) differs from the reference's fp64 quotient by less than 2e-7 q.  When q32
// is further than 4e-7 q from the nearest integer the two have the same floor (returns true, cell set); otherwise (about
// 1e-5 of the points) the caller takes the exact path.  The integer is taken with the 2^23 magic-number add (FADD/integer
// pipes): the conversion instructions (F2F/FRND/F2I) all issue on the quarter-rate XU pipe, which is what bounded k_count.
// Ten instructions per axis: FADD, FMUL, FADD, IADD, FADD, FADD, FMUL, two FSETP, one predicated IADD.
//   * rn = RN(q) exactly while q < 2^22; beyond 1.25e6 the margin eps = 4e-7 q exceeds 0.5 >= |q - rn|, so large, infinite
//     and NaN quotients never decide (no range test needed);
//   * d >= 0 (off is the minimum), so q >= 0 and a quotient below 1/2 is in cell 0 whatever its distance from 0:
//     flat ground at the minimum z puts most of a scan exactly there.
NDT_HD bool cell_prefilter32(float p, float off32, float rv32, unsigned &cell) {
    const float d = p - off32;
    const float q = d * rv32;
    const float m = q + 8388608.0f;                           // RN(q) in the low mantissa bits
    const int ri = float_bits(m) - 0x4B000000;
    const float t = q - (m - 8388608.0f);                     // signed distance from the nearest integer, |t| <= 1/2
    const float eps = q * 4e-7f;
    cell = (unsigned)(t < 0.0f ? ri - 1 : ri);                // nearest -> floor
    return fabsf(t) > eps || q < 0.5f;
}

}  // namespace ndt
