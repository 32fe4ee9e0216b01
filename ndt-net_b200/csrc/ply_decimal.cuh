// ply_decimal.cuh — decimal literal -> correctly rounded double, as CPython's float()/int() on the tokens of
// /root/reference/ndnet/datasets/CARLA_Seg.py:120-123.  Integer arithmetic only, so that the same code gives the same
// bits on the device and (for the CPU unit test tests/test_ply_decimal.py, which compiles this header with g++) on the host.
#pragma once
#include <cmath>
#include <cstdint>

#ifndef __CUDACC__
#define PLY_HD
#else
#define PLY_HD __host__ __device__
#endif

namespace ply {

typedef unsigned long long u64;
typedef unsigned __int128 u128;

// error codes of one line (low byte of the packed error word); the C ABI returns -300 - code
enum : int {
    kErrShortLine = 1,     // fewer tokens than the reader indexes: IndexError in the reference (:120-123)
    kErrLiteral = 2,       // float()/int() would raise ValueError (:120-123)
    kErrClassBound = 3,    // class_tag > n_classes: the reference's own ValueError (:127-128)
    kErrUnsupported = 4,   // possibly valid for CPython but outside what this parser converts exactly
    kErrNegativeClass = 5, // np.asarray(classes, dtype=np.uint16) on a negative tag (:146)
    kErrEncoding = 6,      // lone '\r' line ends or non-ASCII bytes
};

PLY_HD inline int clz64(u64 x) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

PLY_HD inline double pow10_exact(int m) {       // 10^m, m <= 22: exactly representable
    constexpr double t[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                              1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[m];
}
constexpr int kMaxExp10 = 55;     // 5^55 < 2^128

// (q + f) * 2^E with 0 <= f < 1, sticky = (f != 0); q has at least 55 significant bits whenever sticky is set.
PLY_HD inline double round_pack(u64 q, bool sticky, int E) {
    const int nb = 64 - clz64(q);
    if (nb > 53) {
        const int shift = nb - 53;
        const u64 low = q & ((1ull << shift) - 1), half = 1ull << (shift - 1);
        u64 q53 = q >> shift;
        if (low > half || (low == half && (sticky || (q53 & 1ull)))) q53++;
        q = q53;
        E += shift;
    }
    return scalbn((double)(long long)q, E);       // q <= 2^53: exact, and the exponent stays far inside the normal range
}

// w * 10^e10 for w > 0, |e10| <= kMaxExp10, correctly rounded (round-half-even) — CPython float() semantics.
PLY_HD inline double decimal_to_double(u64 w, int e10) {
    const int m = e10 < 0 ? -e10 : e10;
    if (w <= (1ull << 53) && m <= 22) {           // both operands exact, one IEEE operation
        const double d = (double)(long long)w;
        return e10 < 0 ? d / pow10_exact(m) : d * pow10_exact(m);
    }
    u128 p5 = 1;
    for (int i = 0; i < m; i++) p5 *= 5;
    if (e10 >= 0) {                               // N = w * 5^m (<= 192 bits); value = N * 2^m
        const u128 lo = (u128)(u64)p5 * w;
        const u128 mid = (u128)(u64)(p5 >> 64) * w + (u64)(lo >> 64);
        const u64 n0 = (u64)lo, n1 = (u64)mid, n2 = (u64)(mid >> 64);
        int top;
        if (n2) top = 191 - clz64(n2);
        else if (n1) top = 127 - clz64(n1);
        else top = 63 - clz64(n0);
        if (top <= 63) return round_pack(n0, false, m);
        const int sh = top - 63;                  // keep bits [sh, top]
        u64 q; bool sticky;
        if (sh < 64) {
            q = (n0 >> sh) | (n1 << (64 - sh));
            sticky = (n0 << (64 - sh)) != 0;
        } else if (sh == 64) {
            q = n1; sticky = n0 != 0;
        } else {
            const int s2 = sh - 64;
            q = (n1 >> s2) | (n2 << (64 - s2));
            sticky = n0 != 0 || (n1 << (64 - s2)) != 0;
        }
        return round_pack(q, sticky, m + sh);
    }
    // value = w / (5^m 2^m): long division of (w << lz) * 2^s by 5^m with s = bitlen(5^m) - 1, so 2^62 < q < 2^64
    const int lz = clz64(w);
    const u64 wn = w << lz;
    const u64 dhi = (u64)(p5 >> 64);
    const int s = (dhi ? 127 - clz64(dhi) : 63 - clz64((u64)p5));
    u128 R = 0;
    u64 q = 0;
    for (int i = 63; i >= 0; --i) {
        const bool carry = (u64)(R >> 127) != 0;
        R = (R << 1) | ((wn >> i) & 1ull);
        const bool ge = carry || R >= p5;
        if (ge) R -= p5;
        q = (q << 1) | (u64)ge;
    }
    for (int i = 0; i < s; ++i) {
        const bool carry = (u64)(R >> 127) != 0;
        R <<= 1;
        const bool ge = carry || R >= p5;
        if (ge) R -= p5;
        q = (q << 1) | (u64)ge;
    }
    return round_pack(q, R != 0, -s - m - lz);
}

// One token as CPython's float(): [+-] (digits [. digits*] | . digits) [(e|E) [+-] digits].  0 ok, else an error code.
PLY_HD inline int parse_float(const unsigned char *t, long a, long b, double *out) {
    long i = a;
    bool neg = false;
    bool foreign = false;                          // characters that only other CPython literals use
    for (long k = a; k < b; k++) {
        const unsigned c = t[k];
        if (!((c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E')) foreign = true;
    }
    const int bad = foreign ? kErrUnsupported : kErrLiteral;
    if (i < b && (t[i] == '+' || t[i] == '-')) { neg = t[i] == '-'; i++; }
    u64 w = 0;
    int nd = 0;            // significant digits held in w
    int e10 = 0;
    int ndigits = 0;       // mantissa digits seen at all
    bool dropped_nonzero = false;
    bool frac = false;
    for (; i < b; i++) {
        const unsigned c = t[i];
        if (c == '.') {
            if (frac) return bad;
            frac = true;
            continue;
        }
        if (c < '0' || c > '9') break;
        ndigits++;
        const unsigned d = c - '0';
        if (nd < 19) {
            if (w || d) { w = w * 10 + d; nd++; }
            if (frac) e10--;
        } else {
            dropped_nonzero |= d != 0;
            if (!frac) e10++;
        }
    }
    if (ndigits == 0) return bad;
    if (i < b) {
        if (t[i] != 'e' && t[i] != 'E') return bad;
        i++;
        bool eneg = false;
        if (i < b && (t[i] == '+' || t[i] == '-')) { eneg = t[i] == '-'; i++; }
        if (i >= b) return bad;
        int ex = 0;
        for (; i < b; i++) {
            const unsigned c = t[i];
            if (c < '0' || c > '9') return bad;
            if (ex < 100000) ex = ex * 10 + (int)(c - '0');
        }
        e10 += eneg ? -ex : ex;
    }
    if (dropped_nonzero) return kErrUnsupported;
    double v = 0.0;
    if (w) {
        while (w % 10 == 0 && e10 < 0) { w /= 10; e10++; }       // "1.500" -> 15e-1: keeps common inputs on the fast path
        if (e10 > kMaxExp10 || e10 < -kMaxExp10) return kErrUnsupported;
        v = decimal_to_double(w, e10);
    }
    *out = neg ? -v : v;
    return 0;
}

// One token as CPython's int(): [+-] digits.  Values beyond 2^62 saturate (they are out of bounds for any class set).
PLY_HD inline int parse_int(const unsigned char *t, long a, long b, long long *out) {
    long i = a;
    bool neg = false, foreign = false;
    for (long k = a; k < b; k++) {
        const unsigned c = t[k];
        if (!((c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E')) foreign = true;
    }
    const int bad = foreign ? kErrUnsupported : kErrLiteral;
    if (i < b && (t[i] == '+' || t[i] == '-')) { neg = t[i] == '-'; i++; }
    if (i >= b) return bad;
    long long v = 0;
    for (; i < b; i++) {
        const unsigned c = t[i];
        if (c < '0' || c > '9') return bad;
        if (v < (1ll << 62) / 10) v = v * 10 + (long long)(c - '0');
        else v = 1ll << 62;
    }
    *out = neg ? -v : v;
    return 0;
}

// str.split() / str.strip() whitespace within ASCII: \t \n \v \f \r, FS GS RS US, space
PLY_HD inline bool is_space(unsigned c) { return c == 0x20u || (c >= 0x09u && c <= 0x0du) || (c >= 0x1cu && c <= 0x1fu); }

// One data line text[a, b) as the reference reads it (CARLA_Seg.py:118-123):
//   data = point.strip().split(); x = float(data[0]); y = float(data[1]); z = float(data[2]); class_tag = int(data[-1])
// in that evaluation order.  Returns 0 or the error code of the first failing step.
PLY_HD inline int parse_line(const unsigned char *text, long a, long b, double xyz[3], long long *tag) {
    long ta[3], tb[3], la = -1, lb = -1;
    int ntok = 0;
    long i = a;
    while (i < b) {
        while (i < b && is_space(text[i])) i++;
        if (i >= b) break;
        const long s = i;
        while (i < b && !is_space(text[i])) i++;
        if (ntok < 3) { ta[ntok] = s; tb[ntok] = i; }
        la = s; lb = i;
        ntok++;
    }
    for (int k = 0; k < 3; k++) {
        if (ntok <= k) return kErrShortLine;
        const int r = parse_float(text, ta[k], tb[k], &xyz[k]);
        if (r) return r;
    }
    return parse_int(text, la, lb, tag);
}

}  // namespace ply
