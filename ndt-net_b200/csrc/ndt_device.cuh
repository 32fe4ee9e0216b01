// ndt_device.cuh — device-side building blocks of the NDT hot path (sm_100a).
//
// All arithmetic that decides voxel ids, statistics, divergences and the selection is fp64 and is
// compiled with -fmad=false: the reference is x86-64 -O0 C (no FMA contraction) and
// parity is bit-level.  Reference lines each piece follows are cited inline
// (paths relative to /root/reference/core_legacy/).
#pragma once
#include "ndt_cell.cuh"
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace ndt {

constexpr int kMaxGuessIterations = 15;     // include/ndnet_core/ndt.h:43
constexpr int kSearchPairLaunches = 6;      // (k_count, k_decide) launch pairs of the voxel-size search; k_search_tail runs whatever rounds remain
constexpr double kMinVoxelGuess = 0.01;     // ndt.h:41
constexpr double kMaxVoxelGuess = 30.0;     // ndt.h:42
constexpr double kUpperThreshold = 0.2;     // ndt.h:38
constexpr int kWorkers = 8;                 // normal_distributions.h:39
constexpr int kDirs = 6;                    // voxel.h:35-43  X+,X-,Y+,Y-,Z+,Z-

constexpr unsigned kMaxGridCells = 1u << 25;   // cells per cloud the bitmap workspace can hold
constexpr unsigned kSmemBitmapBits = 1u << 17; // grids up to this many cells are counted in shared memory (16 KB: 8 CTAs of k_count per SM)
constexpr int kCountPointsPerCta = 4096;
constexpr int kCountCtasPerCloud = 8;          // CTAs of k_count per cloud (each walks N / 8 consecutive points)
constexpr int kRankTile = 4096;                // points per rank tile (one warp walks one tile in order)
constexpr unsigned kDropped = 0xFFFFFFFFu;
constexpr int kLimWords = 8;                   // per cloud: encoded max x,y,z, min x,y,z, NaN-seen flag, (pad)
constexpr int kStatusNaNInput = -5;            // a coordinate is NaN: refused (the reference's behaviour is undefined, voxel.c:89-91)
constexpr int kSortedStride = 4;               // voxel-sorted points are {x, y, z, label bits}: one 16-byte store per point
constexpr int kSmemLabelBins = 64;
constexpr unsigned kHeavyVoxel = 128;           // voxels with at least this many points go to k_stats (four lanes each), the rest to k_stats_light

// Per-cloud search/grid state (device resident, one per cloud).
struct CloudState {
    double lim[6];        // max x,y,z, min x,y,z   (pointclouds.c:40-66)
    double guess, lo, hi; // ndt.c:136-138
    double off[3];
    int len[3];
    int risky;            // some axis has (max-min)/vs an exact integer: points may fall outside (A4)
    int status;           // 1 searching, 0 accepted, <0 error code of ndt_downsample
    int iter;             // rejected guesses so far (ndt.c:143,185)
    int evals;
    unsigned G;           // cells of the current grid
    unsigned nwords;      // bitmap words of the current grid
    unsigned V;           // occupied voxels of the accepted pass
    unsigned K;           // divergence-list entries
    unsigned n_valid;     // after pruning
    unsigned walk;        // idx_to_remove at the end of prune_nds' walk
    int prune_ret;
    unsigned n_out;
    unsigned n_survivors;
    unsigned fail[kWorkers]; // first point of each worker chunk that left the grid (A4), else 0xFFFFFFFF
    unsigned n_heavy;     // voxels with >= kHeavyVoxel points (first n_heavy entries of vox_order)
    int passes;           // guesses that needed a pass over the points (the others were decided by skip_small_grids)
};

// Order-preserving map double <-> uint64 for atomicMin/atomicMax.
__device__ __forceinline__ unsigned long long enc_f64(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_f64(unsigned long long u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
    return __longlong_as_double((long long)u);
}

// floor((p - off) / vs) exactly as voxel.c:89-91 computes it: see ndt_cell.cuh (shared with the host-side property test)
__device__ __forceinline__ unsigned axis_cell(double p, double off, double vs, double rv) { return cell_exact64(p, off, vs, rv); }

// voxel.c:83-103 + :177-189.  Returns false when the point is outside the grid.
__device__ __forceinline__ bool voxel_of(double x, double y, double z, const double off[3], double vs, double rv,
                                         const int len[3], unsigned &id) {
    const unsigned vx = axis_cell(x, off[0], vs, rv);
    const unsigned vy = axis_cell(y, off[1], vs, rv);
    const unsigned vz = axis_cell(z, off[2], vs, rv);
    if (vx >= (unsigned)len[0] || vy >= (unsigned)len[1] || vz >= (unsigned)len[2]) return false;
    id = vz * (unsigned)len[0] * (unsigned)len[1] + vy * (unsigned)len[0] + vx;
    return true;
}

// out-of-line copy for the fp32 prefilter's rare fallback, so that the compiler cannot speculate the fp64 path
static __device__ __noinline__ unsigned axis_cell_rare(double p, double off, double vs, double rv) { return axis_cell(p, off, vs, rv); }

// Grid parameters of one pass, with fp32 copies for the prefilter below.
struct GridCtx {
    double off[3], vs, rv;
    int len[3];
    float off32[3], rv32;
};

__device__ __forceinline__ GridCtx make_grid_ctx(const CloudState &s) {
    GridCtx g;
    g.vs = s.guess; g.rv = 1.0 / s.guess; g.rv32 = (float)g.rv;
#pragma unroll
    for (int a = 0; a < 3; a++) { g.off[a] = s.off[a]; g.len[a] = s.len[a]; g.off32[a] = (float)s.off[a]; }
    // keep the fp32 copies in registers: left alone, the compiler re-converts the doubles at every use, and F2F issues on
    // the quarter-rate XU pipe (two per point in k_count's inner loop)
    asm volatile("" : "+f"(g.off32[0]), "+f"(g.off32[1]), "+f"(g.off32[2]), "+f"(g.rv32));
    return g;
}

// fp32 prefilter (ndt_cell.cuh) with the out-of-line exact fallback
__device__ __forceinline__ unsigned axis_cell32(float p, int a, const GridCtx &g) {
    unsigned cell;
    if (cell_prefilter32(p, g.off32[a], g.rv32, cell)) return cell;
    return axis_cell_rare((double)p, g.off[a], g.vs, g.rv);
}

template <typename T>
__device__ __forceinline__ bool point_cell(const T *__restrict__ p, long i, const GridCtx &g, unsigned &id);

template <>
__device__ __forceinline__ bool point_cell<float>(const float *__restrict__ p, long i, const GridCtx &g, unsigned &id) {
    const float x = p[i * 3 + 0], y = p[i * 3 + 1], z = p[i * 3 + 2];
    const unsigned vx = axis_cell32(x, 0, g), vy = axis_cell32(y, 1, g), vz = axis_cell32(z, 2, g);
    if (vx >= (unsigned)g.len[0] || vy >= (unsigned)g.len[1] || vz >= (unsigned)g.len[2]) return false;
    id = vz * (unsigned)g.len[0] * (unsigned)g.len[1] + vy * (unsigned)g.len[0] + vx;
    return true;
}

template <>
__device__ __forceinline__ bool point_cell<double>(const double *__restrict__ p, long i, const GridCtx &g, unsigned &id) {
    return voxel_of(p[i * 3 + 0], p[i * 3 + 1], p[i * 3 + 2], g.off, g.vs, g.rv, g.len, id);
}

// ---------------------------------------------------------------------------------------------
// GSL 2.7.1 3x3 routines, operation order identical to oracle/ndt_oracle.c (which is checked
// against the general-size loops of oracle/gsl_shim/gsl_shim.c).
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ void swap_rows(double a[9], int r0, int r1) {
#pragma unroll
    for (int c = 0; c < 3; c++) { double t = a[r0 * 3 + c]; a[r0 * 3 + c] = a[r1 * 3 + c]; a[r1 * 3 + c] = t; }
}

// gsl_linalg_LU_decomp (level-2 path), in place.  perm packs the permutation (2 bits per entry).
__device__ __forceinline__ void lu3(double a[9], int &perm, int &signum) {
    // column 0
    int p0 = 0; { double mx = 0.0; double v;
        v = fabs(a[0]); if (v > mx) { mx = v; p0 = 0; }
        v = fabs(a[3]); if (v > mx) { mx = v; p0 = 1; }
        v = fabs(a[6]); if (v > mx) { mx = v; p0 = 2; } }
    if (p0 == 1) swap_rows(a, 0, 1); else if (p0 == 2) swap_rows(a, 0, 2);
    {
        const double ajj = a[0];
        if (fabs(ajj) >= DBL_MIN) { const double r = 1.0 / ajj; a[3] *= r; a[6] *= r; }
        else { a[3] /= ajj; a[6] /= ajj; }
        const double t1 = -1.0 * a[3];
        a[4] += a[1] * t1; a[5] += a[2] * t1;
        const double t2 = -1.0 * a[6];
        a[7] += a[1] * t2; a[8] += a[2] * t2;
    }
    // column 1
    int p1 = 1; { double mx = 0.0; double v;
        v = fabs(a[4]); if (v > mx) { mx = v; p1 = 1; }
        v = fabs(a[7]); if (v > mx) { mx = v; p1 = 2; } }
    if (p1 == 2) swap_rows(a, 1, 2);
    {
        const double ajj = a[4];
        if (fabs(ajj) >= DBL_MIN) { const double r = 1.0 / ajj; a[7] *= r; }
        else { a[7] /= ajj; }
        const double t2 = -1.0 * a[7];
        a[8] += a[5] * t2;
    }
    // permutation vector by successive transpositions, signum = parity
    int q0 = 0, q1 = 1, q2 = 2;
    signum = 1;
    if (p0 == 1) { int t = q1; q1 = q0; q0 = t; signum = -signum; }
    else if (p0 == 2) { int t = q2; q2 = q0; q0 = t; signum = -signum; }
    if (p1 == 2) { int t = q2; q2 = q1; q1 = t; signum = -signum; }
    perm = q0 | (q1 << 2) | (q2 << 4);
}

__device__ __forceinline__ double lu3_det(const double lu[9], int signum) {
    double det = (double)signum;
    det *= lu[0]; det *= lu[4]; det *= lu[8];
    return det;
}

// true when gsl_linalg_LU_sgndet would return 0 (some u_ii == 0)
__device__ __forceinline__ bool lu3_sgndet_zero(const double lu[9]) {
    if (lu[0] == 0) return true;
    if (lu[4] == 0) return true;
    if (lu[8] == 0) return true;
    return false;
}

// gsl_linalg_LU_invert
__device__ __forceinline__ void lu3_invert(const double lu[9], int perm, double inv[9]) {
    double t[9];
#pragma unroll
    for (int i = 0; i < 9; i++) t[i] = lu[i];
    t[0] = 1.0 / t[0];
    t[4] = 1.0 / t[4];
    t[1] = (0.0 + t[1] * t[0]) * (-t[4]);
    t[8] = 1.0 / t[8];
    {
        const double x0 = (0.0 + t[5] * t[1]) + t[2] * t[0];
        const double x1 = 0.0 + t[5] * t[4];
        t[2] = x0 * (-t[8]);
        t[5] = x1 * (-t[8]);
    }
    t[7] = (t[7] + 0.0) * -1.0;
    {
        const double x1 = t[6] + (0.0 + t[3] * t[7]);
        const double x0 = t[3] + 0.0;
        t[3] = x0 * -1.0;
        t[6] = x1 * -1.0;
    }
    {
        t[0] += (0.0 + t[3] * t[1]) + t[6] * t[2];
        const double u11 = t[4];
        t[4] += 0.0 + t[7] * t[5];
        if (u11 == 0.0) t[3] = 0.0;
        else if (u11 != 1.0) t[3] *= u11;
        { const double tmp = 1.0 * t[5]; if (tmp != 0.0) t[3] += tmp * t[6]; }
        { const double tmp = 0.0 + t[7] * t[2]; t[1] += 1.0 * tmp; }
        t[6] *= t[8];
        t[7] *= t[8];
    }
    const int q0 = perm & 3, q1 = (perm >> 2) & 3;          // the third entry is whatever is left
#pragma unroll
    for (int i = 0; i < 3; i++) {
        // inv[i][perm[k]] = t[i][k]
        const double v0 = t[i * 3 + 0], v1 = t[i * 3 + 1], v2 = t[i * 3 + 2];
        inv[i * 3 + 0] = (q0 == 0) ? v0 : ((q1 == 0) ? v1 : v2);
        inv[i * 3 + 1] = (q0 == 1) ? v0 : ((q1 == 1) ? v1 : v2);
        inv[i * 3 + 2] = (q0 == 2) ? v0 : ((q1 == 2) ? v1 : v2);
    }
}

// kullback_leibler.c:60-115 after both operands were factorised.  P = packed LU of p (current
// state), Q = packed LU of q with its permutation.  Returns false for the -2 (singular) exits.
__device__ __forceinline__ bool pseudo_kl(const double P[9], int psign, const double Q[9], int qperm, int qsign,
                                          double &div) {
    const double p_det = lu3_det(P, psign);
    const double q_det = lu3_det(Q, qsign);
    if (p_det == 0 || q_det == 0) return false;                 // :66
    if (lu3_sgndet_zero(P) || lu3_sgndet_zero(Q)) return false; // :71-78
    double qinv[9];
    lu3_invert(Q, qperm, qinv);                                 // :92
    // diagonal of cblas_dgemm(qinv, P), beta = 0: k outer, zero entries of qinv skipped (:96-102)
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double t0 = 1.0 * qinv[0 * 3 + k]; if (t0 != 0.0) c0 += t0 * P[k * 3 + 0];
        const double t1 = 1.0 * qinv[1 * 3 + k]; if (t1 != 0.0) c1 += t1 * P[k * 3 + 1];
        const double t2 = 1.0 * qinv[2 * 3 + k]; if (t2 != 0.0) c2 += t2 * P[k * 3 + 2];
    }
    double trace = 0.0;
    trace += c0; trace += c1; trace += c2;
    const double first_part_result = 0.0;                       // aliased dgemm, beta = 0 (:105-112)
    div = 0.5 * (first_part_result + trace - log(q_det / p_det) - 3);   // :115
    return true;
}

// Voxel-sorted point records {x, y, z, label bits} (k_scatter writes them with ONE vector store per point: the scatter is
// bound by L2 store transactions, not bytes; k_stats / k_stats_light / k_votes read them back).
template <typename T> __device__ __forceinline__ void store_sorted(T *q, T x, T y, T z, unsigned label);
template <> __device__ __forceinline__ void store_sorted<float>(float *q, float x, float y, float z, unsigned label) {
    *reinterpret_cast<float4 *>(q) = make_float4(x, y, z, __uint_as_float(label));
}
template <> __device__ __forceinline__ void store_sorted<double>(double *q, double x, double y, double z, unsigned label) {
    *reinterpret_cast<double2 *>(q) = make_double2(x, y);
    *reinterpret_cast<double2 *>(q + 2) = make_double2(z, __longlong_as_double((long long)label));
}
template <typename T> __device__ __forceinline__ void load_sorted(const T *p, T &x, T &y, T &z);
template <> __device__ __forceinline__ void load_sorted<float>(const float *p, float &x, float &y, float &z) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    x = v.x; y = v.y; z = v.z;
}
template <> __device__ __forceinline__ void load_sorted<double>(const double *p, double &x, double &y, double &z) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    x = v.x; y = v.y; z = p[2];
}
// the whole record: coordinates and the label lane
template <typename T> __device__ __forceinline__ void load_sorted_rec(const T *p, T &x, T &y, T &z, unsigned &label);
template <> __device__ __forceinline__ void load_sorted_rec<float>(const float *p, float &x, float &y, float &z, unsigned &label) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    x = v.x; y = v.y; z = v.z; label = __float_as_uint(v.w);
}
template <> __device__ __forceinline__ void load_sorted_rec<double>(const double *p, double &x, double &y, double &z, unsigned &label) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    const double2 w = *reinterpret_cast<const double2 *>(p + 2);
    x = v.x; y = v.y; z = w.x; label = (unsigned)__double_as_longlong(w.y);
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned sorted_label(const float *p) { return __float_as_uint(p[3]); }
__device__ __forceinline__ unsigned sorted_label(const double *p) { return (unsigned)__double_as_longlong(p[3]); }

}  // namespace ndt
