// ndt_kernels.cu — the batched NDT hot path for B200 (sm_100a): bounding box, voxel-size search,
// stable voxel assignment, per-voxel sequential statistics, neighbour pseudo-KL with LU-history
// emulation, descending stable sort with the NaN rule, head-first prune, ascending compaction.
//
// Compiled with -fmad=false (see ndt_device.cuh).  Replaces /root/reference/core_legacy/src/
// {pointclouds,voxel,normal_distributions,kullback_leibler,ndt}.c for a whole batch of clouds per
// launch sequence; nothing here synchronises with the host.
#include "ndt_host.h"
#include "ndt_device.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifndef NDT_STATS_MIN_CTAS
#define NDT_STATS_MIN_CTAS 12                  // one-warp CTAs of k_stats per SM the register allocation must allow (<= 168 registers: below that the
                                               // scheduler stops interleaving the tails of earlier steps with the mean chain)
#endif
constexpr unsigned kMaxSortedHeavy = 1024;       // k_offsets sorts up to this many heavy voxels exactly (one thread each)
constexpr int kSelectThreads = 512;             // k_select: three 512-thread CTAs per SM overlap one another's barriers
#ifndef NDT_RANK_MATCH_ANY
#define NDT_RANK_MATCH_ANY 0
#endif

namespace ndt {

#ifdef NDT_CHECKS
#define CHK(cond, id) do { if (!(cond)) { printf("NDT_CHECK %d failed: block %d thread %d\n", id, blockIdx.x, threadIdx.x); __trap(); } } while (0)
#else
#define CHK(cond, id) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// K1  bounding box (pointclouds.c:40-66).  grid (chunks, B), block 256.
// ------------------------------------------------------------------------------------------------
template <typename T> struct LimAcc {
    T mx[3], mn[3];
    bool nan;
    __device__ __forceinline__ void init() { nan = false; for (int a = 0; a < 3; a++) { mx[a] = -INFINITY; mn[a] = INFINITY; } }
    __device__ __forceinline__ void add(int a, T v) {
        nan = nan || v != v;
        mx[a] = v > mx[a] ? v : mx[a];          // NaN never wins a comparison (a NaN cloud is refused anyway)
        mn[a] = v < mn[a] ? v : mn[a];
    }
};

// whether cloud b of a [B][N][3] fp32 array starts on a 16-byte boundary (then 4 points = 3 aligned float4)
template <typename T> __device__ __forceinline__ bool cloud_vectorizable(const T *) { return false; }
template <> __device__ __forceinline__ bool cloud_vectorizable<float>(const float *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// four consecutive points with three coalesced 16-byte loads (p: the 16-byte aligned first coordinate of the group)
__device__ __forceinline__ void load_points4(const float *p, float x[4], float y[4], float z[4]) {
    const float4 *q = reinterpret_cast<const float4 *>(p);
    const float4 a = q[0], b = q[1], c = q[2];
    x[0] = a.x; y[0] = a.y; z[0] = a.z; x[1] = a.w; y[1] = b.x; z[1] = b.y;
    x[2] = b.z; y[2] = b.w; z[2] = c.x; x[3] = c.y; y[3] = c.z; z[3] = c.w;
}
__device__ __forceinline__ void load_points4(const double *, double *, double *, double *) {}   // never taken (cloud_vectorizable<double> is false)

template <typename T>
__global__ void __launch_bounds__(256) k_limits(const T *__restrict__ pts, long N, unsigned long long *__restrict__ lim_enc) {
    const int b = blockIdx.y;
    const T *p = pts + (size_t)b * N * 3;
    LimAcc<T> acc;
    acc.init();
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long)gridDim.x * blockDim.x;
    if (cloud_vectorizable<T>(p)) {
        const long groups = N / 4;
        for (long g = tid; g < groups; g += nthreads) {
            T x[4], y[4], z[4];
            load_points4(p + g * 12, x, y, z);
#pragma unroll
            for (int k = 0; k < 4; k++) { acc.add(0, x[k]); acc.add(1, y[k]); acc.add(2, z[k]); }
        }
        for (long i = groups * 4 + tid; i < N; i += nthreads) { acc.add(0, p[i * 3 + 0]); acc.add(1, p[i * 3 + 1]); acc.add(2, p[i * 3 + 2]); }
    } else {
        for (long i = tid; i < N; i += nthreads) { acc.add(0, p[i * 3 + 0]); acc.add(1, p[i * 3 + 1]); acc.add(2, p[i * 3 + 2]); }
    }
    // warp reduce through the order-preserving encoding, then one atomic per warp (the +-inf of a thread that saw no
    // point is the identity of the reduction)
    unsigned long long e[6];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        e[a] = enc_f64((double)acc.mx[a]);
        e[3 + a] = enc_f64((double)acc.mn[a]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            unsigned long long v = __shfl_xor_sync(0xffffffffu, e[a], o);
            e[a] = v > e[a] ? v : e[a];
            unsigned long long w = __shfl_xor_sync(0xffffffffu, e[3 + a], o);
            e[3 + a] = w < e[3 + a] ? w : e[3 + a];
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            atomicMax(&lim_enc[b * kLimWords + a], e[a]);
            atomicMin(&lim_enc[b * kLimWords + 3 + a], e[3 + a]);
        }
    }
    // a NaN coordinate is refused (status -5): the reference converts floor(NaN) to unsigned, which is undefined
    // behaviour in C (voxel.c:89-91); silently voxelising such a point anywhere would not be parity with anything
    if (__any_sync(0xffffffffu, acc.nan) && (threadIdx.x & 31) == 0) lim_enc[b * kLimWords + 6] = 1ull;
}

// grid + risk flag for a guess (voxel.c:61-81); returns false when the grid cannot be held
__device__ bool set_grid(CloudState &s) {
    double q[3];
    bool risky = false;
    double cells = 1.0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const double dim = s.lim[a] - s.lim[3 + a];
        q[a] = dim / s.guess;
        const double c = ceil(q[a]);
        if (!(c >= 0.0) || c > 2147483647.0) return false;
        s.len[a] = (int)c;
        s.off[a] = s.lim[3 + a];
        if (c == q[a]) risky = true;   // floor(q) == len is reachable (A4); includes degenerate axes (len 0)
        cells *= c;
    }
    if (cells > (double)kMaxGridCells) return false;
    s.G = (unsigned)s.len[0] * (unsigned)s.len[1] * (unsigned)s.len[2];
    s.nwords = (s.G + 31u) >> 5;
    s.risky = risky ? 1 : 0;
    return true;
}

// A grid with fewer cells than D cannot hold D occupied voxels: whatever the points are, the pass would count
// num_nds <= G < D and take the branch of ndt.c:171 (hi = guess).  Take it without reading the points; the evaluation
// still counts (the reference runs it), so `evaluations` and the 15-guess limit are unchanged.  A degenerate axis
// (G == 0, A3) goes the same way down to -3.
__device__ void skip_small_grids(CloudState &s, unsigned long D) {
    while (s.status == 1 && (unsigned long)s.G < D) {
        s.evals++;
        s.hi = s.guess;                                                                                    // ndt.c:171
        s.guess = s.lo + (s.hi - s.lo) / 2.0;                                                              // ndt.c:183
        s.iter++;
        if (s.iter >= kMaxGuessIterations) { s.status = -3; break; }                                       // ndt.c:187-194
        if (!set_grid(s)) { s.status = -1; break; }
    }
}

// ------------------------------------------------------------------------------------------------
// K2  search bookkeeping (ndt.c:136-194).  grid B, block 256.  phase 0 = initialise from the limits,
// phase 1 = count the bitmap of the pass that just ran and accept / bisect.
// On acceptance it also builds the voxel directory: per word {bits, prefix popcount} and the list of
// occupied cell ids in ascending order (slot -> cell).
// ------------------------------------------------------------------------------------------------
// (a device function: k_decide runs it once per launch, k_search_tail once per round; block of 256 threads.)
// `states` and `bitmap` carry NO __restrict__ here, in k_decide and in k_search_tail: thread 0 writes the new search state
// (grid, word count, status) and the other threads read it behind a barrier, round after round in k_search_tail.  With
// __restrict__ the compiler may keep a value loaded before the barrier (it did, once everything was inlined into the tail's
// loop: the bitmap of the next guess was cleared with the PREVIOUS guess's word count by every thread but thread 0).
__device__ __forceinline__ void decide_body(CloudState *states, const unsigned long long *__restrict__ lim_enc,
                                            uint2 *bitmap, size_t bitmap_stride,
                                            unsigned *__restrict__ vox_cell, unsigned vcap, long num_desired, int phase, int b) {
    CloudState &s = states[b];
    uint2 *bm = bitmap + (size_t)b * bitmap_stride;
    __shared__ unsigned s_part[8];
    __shared__ unsigned s_total;
    __shared__ int s_action;   // 0 nothing, 1 zero bitmap for next pass, 2 accepted -> build directory
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (phase == 0) {
        if (tid == 0) {
            for (int a = 0; a < 3; a++) {
                const double m = dec_f64(lim_enc[b * kLimWords + a]);
                s.lim[a] = m > DBL_MIN ? m : DBL_MIN;              // pointclouds.c:44-46,55-61 (A2)
                const double n = dec_f64(lim_enc[b * kLimWords + 3 + a]);
                s.lim[3 + a] = n < DBL_MAX ? n : DBL_MAX;
            }
            s.guess = (double)(kMaxVoxelGuess - kMinVoxelGuess) / 2.0;   // ndt.c:136
            s.lo = kMinVoxelGuess; s.hi = kMaxVoxelGuess;
            s.iter = 0; s.evals = 0; s.passes = 0; s.V = 0; s.K = 0; s.n_valid = 0; s.walk = 0; s.prune_ret = 0; s.n_out = 0;
            s.n_survivors = 0;
            for (int w = 0; w < kWorkers; w++) s.fail[w] = kDropped;
            s.status = 1;
            if (!set_grid(s)) { s.status = -1; s.G = 0; s.nwords = 0; }
            if (lim_enc[b * kLimWords + 6]) { s.status = kStatusNaNInput; s.G = 0; s.nwords = 0; }
            skip_small_grids(s, (unsigned long)num_desired);
            s_action = s.status == 1 ? 1 : 0;
        }
        __syncthreads();
    } else {
        if (s.status != 1) return;
        // count occupied cells of the pass
        unsigned cnt = 0;
        for (unsigned w = tid; w < s.nwords; w += blockDim.x) cnt += __popc(__ldcg(&bm[w].x));     // L2: set by atomics, possibly in this very kernel (k_search_tail)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) s_part[wid] = cnt;
        __syncthreads();
        if (tid == 0) {
            unsigned total = 0;
            for (int w = 0; w < 8; w++) total += s_part[w];
            s_total = total;
            s.evals++;
            s.passes++;
            const unsigned long num_nds = total;
            const unsigned long D = (unsigned long)num_desired;
            if ((double)num_nds > (double)D * (1 + kUpperThreshold)) { s.lo = s.guess; s_action = 1; }      // ndt.c:169
            else if (num_nds < D) { s.hi = s.guess; s_action = 1; }                                            // ndt.c:171
            else { s_action = 2; s.status = 0; s.V = total; }
            if (s_action == 1) {
                s.guess = s.lo + (s.hi - s.lo) / 2.0;                                                          // ndt.c:183
                s.iter++;
                if (s.iter >= kMaxGuessIterations) { s.status = -3; s_action = 0; }                            // ndt.c:187-194
                else {
                    for (int w = 0; w < kWorkers; w++) s.fail[w] = kDropped;
                    if (!set_grid(s)) { s.status = -1; s_action = 0; }
                    else { skip_small_grids(s, D); if (s.status != 1) s_action = 0; }
                }
            }
        }
        __syncthreads();
    }
    const unsigned nwords_now = *(volatile unsigned *)&s.nwords;      // thread 0 may just have set the next guess's grid
    if (s_action == 1) {
        for (unsigned w = tid; w < nwords_now; w += blockDim.x) bm[w] = make_uint2(0u, 0u);
    } else if (s_action == 2) {
        // exclusive prefix popcount over the words + slot -> cell list
        __shared__ unsigned s_carry;
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (unsigned base = 0; base < nwords_now; base += blockDim.x) {
            const unsigned w = base + tid;
            const unsigned bits = w < nwords_now ? __ldcg(&bm[w].x) : 0u;
            const unsigned c = __popc(bits);
            unsigned inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            if (lane == 31) s_part[wid] = inc;
            __syncthreads();
            unsigned woff = 0;
            for (int k = 0; k < wid; k++) woff += s_part[k];
            unsigned blk_total = 0;
            for (int k = 0; k < 8; k++) blk_total += s_part[k];
            const unsigned excl = s_carry + woff + inc - c;
            if (w < nwords_now) {
                bm[w].y = excl;
                unsigned rest = bits, k = 0;
                while (rest) {
                    const int bit = __ffs(rest) - 1;
                    rest &= rest - 1;
                    if (excl + k < vcap) vox_cell[(size_t)b * vcap + excl + k] = w * 32u + (unsigned)bit;
                    k++;
                }
            }
            __syncthreads();
            if (tid == 0) s_carry += blk_total;
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256) k_decide(CloudState *states, const unsigned long long *__restrict__ lim_enc,
                                                uint2 *bitmap, size_t bitmap_stride,
                                                unsigned *__restrict__ vox_cell, unsigned vcap, long num_desired, int phase) {
    decide_body(states, lim_enc, bitmap, bitmap_stride, vox_cell, vcap, num_desired, phase, (int)blockIdx.x);
}

// ------------------------------------------------------------------------------------------------
// K3  one estimate pass reduced to what the search needs: which cells are occupied
// (normal_distributions.c:28-137 without the statistics).  grid (chunks, B), block 256, each CTA owns
// kCountPointsPerCta consecutive points.  Small grids are marked in a shared-memory bitmap and merged
// with one atomicOr per non-zero word; large grids go straight to the global bitmap.
// A "risky" grid (A4) is handled by CTA 0 alone in two phases so that the worker-chunk early-return
// semantics are exact.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load_point(const T *p, long i, double &x, double &y, double &z) {
    x = (double)p[i * 3 + 0]; y = (double)p[i * 3 + 1]; z = (double)p[i * 3 + 2];
}

// cell id of an fp32 point of a NON-risky grid: the fp32 prefilter per axis, the exact fp64 path (out of line) for the
// ~1e-5 of the points it does not decide.  No bound test: floor((p - off) / vs) < ceil((max - min) / vs) for every
// p <= max when that quotient is not an integer, which is what "not risky" means (set_grid).
static __device__ __noinline__ unsigned cell_id_exact(float x, float y, float z, const GridCtx &g) {
    const unsigned vx = axis_cell((double)x, g.off[0], g.vs, g.rv), vy = axis_cell((double)y, g.off[1], g.vs, g.rv),
                   vz = axis_cell((double)z, g.off[2], g.vs, g.rv);
    return vz * (unsigned)g.len[0] * (unsigned)g.len[1] + vy * (unsigned)g.len[0] + vx;
}
__device__ __forceinline__ unsigned cell_id_fast(float x, float y, float z, const GridCtx &g, unsigned lxly) {
    unsigned vx, vy, vz;
    const bool ok = cell_prefilter32(x, g.off32[0], g.rv32, vx) & cell_prefilter32(y, g.off32[1], g.rv32, vy) &
                    cell_prefilter32(z, g.off32[2], g.rv32, vz);
    if (ok) return vz * lxly + vy * (unsigned)g.len[0] + vx;
    return cell_id_exact(x, y, z, g);
}

// marks the cells of the points [begin, end) of a non-risky grid in the CTA's shared-memory bitmap (kSmem) or straight in
// the global one.  Test before set: only ~V of the G cells are occupied, so after the first few hundred points of the CTA
// nearly every bit is already there and the (serialising) atomic is skipped.
template <typename T, bool kSmem>
__device__ __forceinline__ void count_span(const T *__restrict__ p, long begin, long end, const GridCtx &gc, unsigned *s_bits, uint2 *bm) {
    auto mark = [&](unsigned id) {
        if (kSmem) { if (!((s_bits[id >> 5] >> (id & 31)) & 1u)) atomicOr(&s_bits[id >> 5], 1u << (id & 31)); }
        else if (!((bm[id >> 5].x >> (id & 31)) & 1u)) atomicOr(&bm[id >> 5].x, 1u << (id & 31));
    };
    if (cloud_vectorizable<T>(p)) {
        const unsigned lxly = (unsigned)gc.len[0] * (unsigned)gc.len[1];
        const long vend = begin + (end - begin) / 4 * 4;
        for (long i = begin + threadIdx.x * 4; i < vend; i += blockDim.x * 4) {
            T x[4], y[4], z[4];
            load_points4(p + i * 3, x, y, z);
#pragma unroll
            for (int k = 0; k < 4; k++) mark(cell_id_fast((float)x[k], (float)y[k], (float)z[k], gc, lxly));
        }
        for (long i = vend + threadIdx.x; i < end; i += blockDim.x) {
            unsigned id;
            if (point_cell<T>(p, i, gc, id)) mark(id);
        }
    } else {
        for (long i = begin + threadIdx.x; i < end; i += blockDim.x) {
            unsigned id;
            if (point_cell<T>(p, i, gc, id)) mark(id);
        }
    }
}

// one pass over cloud b as CTA `cta` of `nctas` (k_count: its grid; k_search_tail: the only one)
template <typename T>
__device__ __forceinline__ void count_pass(const T *__restrict__ pts, long N, CloudState *states,
                                           uint2 *bitmap, size_t bitmap_stride, int b, int cta, int nctas,
                                           unsigned *s_bits, unsigned *s_fail) {
    CloudState &s = states[b];
    if (s.status != 1) return;                           // converged or failed clouds cost one exiting CTA per launch slot
    const T *p = pts + (size_t)b * N * 3;
    uint2 *bm = bitmap + (size_t)b * bitmap_stride;
    const unsigned G = s.G, nwords = s.nwords;
    if (G == 0) return;                                  // a degenerate axis: no voxel can be valid (A3)
    const double vs = s.guess;
    const double rv = 1.0 / vs;
    const double off[3] = {s.off[0], s.off[1], s.off[2]};
    const int len[3] = {s.len[0], s.len[1], s.len[2]};
    const long chunk = N / kWorkers;
    const long n_used = chunk * kWorkers;               // the last N % 8 points are never voxelised (A5)
    const bool use_smem = G <= kSmemBitmapBits;

    if (s.risky) {
        if (cta != 0) return;
        if (threadIdx.x < kWorkers) s_fail[threadIdx.x] = kDropped;
        if (use_smem) for (unsigned w = threadIdx.x; w < nwords; w += blockDim.x) s_bits[w] = 0u;
        __syncthreads();
        for (long i = threadIdx.x; i < n_used; i += blockDim.x) {
            double x, y, z; load_point(p, i, x, y, z);
            unsigned id;
            if (!voxel_of(x, y, z, off, vs, rv, len, id)) atomicMin(&s_fail[i / chunk], (unsigned)i);
        }
        __syncthreads();
        for (long i = threadIdx.x; i < n_used; i += blockDim.x) {
            if ((unsigned)i >= s_fail[i / chunk]) continue;
            double x, y, z; load_point(p, i, x, y, z);
            unsigned id;
            voxel_of(x, y, z, off, vs, rv, len, id);
            if (use_smem) atomicOr(&s_bits[id >> 5], 1u << (id & 31));
            else atomicOr(&bm[id >> 5].x, 1u << (id & 31));
        }
        __syncthreads();
        if (threadIdx.x < kWorkers) s.fail[threadIdx.x] = s_fail[threadIdx.x];
        if (use_smem) for (unsigned w = threadIdx.x; w < nwords; w += blockDim.x) if (s_bits[w]) bm[w].x = s_bits[w];
        return;
    }

    // this CTA's contiguous span of the cloud (a multiple of 4 * 256 points, so every thread's groups of four are
    // 16-byte aligned when the cloud is)
    long per = (n_used + nctas - 1) / nctas;
    per = (per + 1023) / 1024 * 1024;
    const long begin = (long)cta * per;
    if (begin >= n_used) return;
    const GridCtx gc = make_grid_ctx(s);
    const long end = begin + per < n_used ? begin + per : n_used;
    if (use_smem) {
        for (unsigned w = threadIdx.x; w < nwords; w += blockDim.x) s_bits[w] = 0u;
        __syncthreads();
        count_span<T, true>(p, begin, end, gc, s_bits, bm);
    } else {
        count_span<T, false>(p, begin, end, gc, s_bits, bm);
    }
    if (use_smem) {
        __syncthreads();
        for (unsigned w = threadIdx.x; w < nwords; w += blockDim.x) {
            const unsigned v = s_bits[w];
            if (v) atomicOr(&bm[w].x, v);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_count(const T *__restrict__ pts, long N, CloudState *__restrict__ states,
                                               uint2 *__restrict__ bitmap, size_t bitmap_stride) {
    __shared__ unsigned s_bits[kSmemBitmapBits / 32];
    __shared__ unsigned s_fail[kWorkers];
    count_pass<T>(pts, N, states, bitmap, bitmap_stride, (int)blockIdx.y, (int)blockIdx.x, (int)gridDim.x, s_bits, s_fail);
}

// The rest of the search in ONE launch: after kSearchPairLaunches (count, decide) pairs nearly every cloud has its voxel size
// (3.2 passes on average); the clouds that are still searching finish here, one CTA per cloud running count and decide rounds
// back to back (a pass by a single CTA takes longer than by eight, but a launch pair for finished clouds - ~10 us, nine of
// them per batch - costs more than the rare straggler).  grid B, block 256.  No __restrict__ on the state and the bitmap
// (see decide_body); the fences do between rounds what the launch boundary does between k_count and k_decide.
template <typename T>
__global__ void __launch_bounds__(256) k_search_tail(const T *__restrict__ pts, long N, CloudState *states,
                                                     const unsigned long long *__restrict__ lim_enc, uint2 *bitmap,
                                                     size_t bitmap_stride, unsigned *__restrict__ vox_cell, unsigned vcap, long num_desired) {
    __shared__ unsigned s_bits[kSmemBitmapBits / 32];
    __shared__ unsigned s_fail[kWorkers];
    const int b = blockIdx.x;
    for (int round = 0; round < kMaxGuessIterations; round++) {
        if (*(volatile int *)&states[b].status != 1) return;      // written by thread 0 in front of a barrier of the previous round
        count_pass<T>(pts, N, states, bitmap, bitmap_stride, b, 0, 1, s_bits, s_fail);
        __threadfence();                                 // the pass's atomics on the bitmap words have reached L2 ...
        __syncthreads();                                 // ... before any thread counts them
        decide_body(states, lim_enc, bitmap, bitmap_stride, vox_cell, vcap, num_desired, 1, b);
        __threadfence();                                 // the cleared words / the new search state are in L2 ...
        __syncthreads();                                 // ... before the next round's atomics and loads touch them
    }
}

// ------------------------------------------------------------------------------------------------
// K4  stable voxel assignment, step 1: every warp walks one tile of kRankTile consecutive points in
// order and gives each point (slot, rank among earlier points of the tile in the same slot).
// grid (ceil(tiles/4), B), block 128 (4 warps = 4 tiles); dynamic smem 4 * vcap u16 counters.
// ------------------------------------------------------------------------------------------------
// slot of a cell in the accepted grid's directory (bitmap word = {occupancy bits, occupied cells before the word})
__device__ __forceinline__ unsigned slot_of_cell(const uint2 *__restrict__ bm, unsigned id) {
    const uint2 w = bm[id >> 5];
    return w.y + __popc(w.x & ((1u << (id & 31)) - 1u));
}

template <typename T>
__global__ void __launch_bounds__(128) k_rank(const T *__restrict__ pts, long N, const CloudState *__restrict__ states,
                                              const uint2 *__restrict__ bitmap, size_t bitmap_stride, unsigned vcap,
                                              int ntiles, unsigned *__restrict__ slot_rank, unsigned *__restrict__ tile_cnt,
                                              int *__restrict__ point_voxel) {
    extern __shared__ unsigned short s_cnt_all[];
    __shared__ __align__(16) unsigned s_slot[4][128], s_cell[4][128];     // per warp: slots / cells of 128 consecutive points
    const int b = blockIdx.y;
    const CloudState &s = states[b];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tile = blockIdx.x * 4 + wid;
    if (tile >= ntiles) return;
    unsigned *sr = slot_rank + (size_t)b * N;
    int *pv = point_voxel ? point_voxel + (size_t)b * N : nullptr;
    const long begin = (long)tile * kRankTile;
    const long tend = begin + kRankTile < N ? begin + kRankTile : N;
    if (s.status != 0) {
        for (long i = begin + lane; i < tend; i += 32) { sr[i] = kDropped; if (pv) pv[i] = -1; }
        return;
    }
    unsigned short *cnt = s_cnt_all + (size_t)wid * vcap;
    const unsigned V = s.V;
    for (unsigned v = lane; v < V; v += 32) cnt[v] = 0;
    __syncwarp();
    const T *p = pts + (size_t)b * N * 3;
    const uint2 *bm = bitmap + (size_t)b * bitmap_stride;
    const long chunk = N / kWorkers;
    const long n_used = chunk * kWorkers;
    const GridCtx gc = make_grid_ctx(s);
    const bool risky = s.risky != 0;
    const int slot_bits = 32 - __clz(V | 1u);            // slots are < V
    // ranks the 32 points [base, base + 32) in order: rank among earlier points of the tile in the same slot
    auto rank32 = [&](long base, unsigned slot, unsigned id) {
        const long i = base + lane;
        // lanes of the group that hold the same slot: one ballot per bit of the slot number (11 for ~1200 voxels), each
        // lane keeping the lanes that agree with it on that bit.  (MATCH.ANY iterates over the distinct values of the warp,
        // ~31 for a scan in random point order: 0.70 ms per 512 scans; 32 shuffle-and-compare steps: 0.52 ms.)
#if NDT_RANK_MATCH_ANY
        const unsigned peers = __match_any_sync(0xffffffffu, slot);
#else
        unsigned peers = __ballot_sync(0xffffffffu, slot != kDropped);
        for (int bit = 0; bit < slot_bits; bit++) {
            const unsigned m = __ballot_sync(0xffffffffu, (slot >> bit) & 1u);
            peers &= ((slot >> bit) & 1u) ? m : ~m;
        }
#endif
        unsigned packed = kDropped;
        unsigned basec = 0;
        if (slot != kDropped) {
            const unsigned before = __popc(peers & ((1u << lane) - 1u));
            basec = cnt[slot];
            packed = slot * (unsigned)kRankTile + basec + before;
        }
        __syncwarp();
        if (slot != kDropped && lane == 31 - __clz(peers)) cnt[slot] = (unsigned short)(basec + __popc(peers));
        __syncwarp();
        if (i < tend) { sr[i] = packed; if (pv) pv[i] = slot != kDropped ? (int)id : -1; }
    };
    if (cloud_vectorizable<T>(p) && !risky) {
        // 128 points per step: every lane takes four consecutive points (three 16-byte loads), the slots go through
        // shared memory so that the ranking still walks the points in index order, 32 at a time
        const unsigned lxly = (unsigned)gc.len[0] * (unsigned)gc.len[1];
        for (long base = begin; base < tend; base += 128) {
            const long i0 = base + 4 * lane;
            unsigned sl[4] = {kDropped, kDropped, kDropped, kDropped}, id[4] = {0u, 0u, 0u, 0u};
            if (i0 + 3 < n_used) {
                T x[4], y[4], z[4];
                load_points4(p + i0 * 3, x, y, z);
#pragma unroll
                for (int k = 0; k < 4; k++) { id[k] = cell_id_fast((float)x[k], (float)y[k], (float)z[k], gc, lxly); sl[k] = slot_of_cell(bm, id[k]); }
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (i0 + k < n_used && point_cell<T>(p, i0 + k, gc, id[k])) sl[k] = slot_of_cell(bm, id[k]);
            }
            *reinterpret_cast<uint4 *>(&s_slot[wid][4 * lane]) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
            if (pv) *reinterpret_cast<uint4 *>(&s_cell[wid][4 * lane]) = make_uint4(id[0], id[1], id[2], id[3]);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (base + 32 * j >= tend) break;
                rank32(base + 32 * j, s_slot[wid][32 * j + lane], pv ? s_cell[wid][32 * j + lane] : 0u);
            }
            __syncwarp();
        }
    } else {
        for (long base = begin; base < tend; base += 32) {
            const long i = base + lane;
            unsigned slot = kDropped;
            unsigned id = 0;
            bool live = i < n_used;
            if (live && risky) live = (unsigned)i < s.fail[(unsigned long)i / (unsigned long)chunk];   // A4: only risky grids drop points
            if (live && point_cell<T>(p, i, gc, id)) slot = slot_of_cell(bm, id);
            rank32(base, slot, id);
        }
    }
    __syncwarp();
    unsigned *tc = tile_cnt + ((size_t)b * ntiles + tile) * vcap;
    for (unsigned v = lane; v < V; v += 32) tc[v] = cnt[v];
}

// ------------------------------------------------------------------------------------------------
// K5  stable voxel assignment, step 2: per slot, turn the per-tile counts into exclusive prefixes (k_tile_prefix, the
// part that moves data: ntiles x V counters per cloud, spread over ceil(V/256) CTAs per cloud) and the slot totals into
// segment starts (k_offsets, one CTA per cloud over V values).
// ------------------------------------------------------------------------------------------------
// step 2a: thread per slot, exclusive prefix of its per-tile counts in place and the slot total.  grid (ceil(vcap/256), B).
__global__ void __launch_bounds__(256) k_tile_prefix(const CloudState *__restrict__ states, unsigned vcap, int ntiles,
                                                     unsigned *__restrict__ tile_cnt, unsigned *__restrict__ vox_n) {
    const int b = blockIdx.y;
    const CloudState &s = states[b];
    if (s.status != 0) return;
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= s.V) return;
    unsigned *tc = tile_cnt + (size_t)b * ntiles * vcap;
    unsigned acc = 0;
    // the kernel is a chain of memory round trips per thread: issue 32 loads before the first dependent store (the running
    // sum is the only dependency), so that a 59-tile scan is two round trips instead of eight
    for (int t0 = 0; t0 < ntiles; t0 += 32) {
        unsigned c[32];
#pragma unroll
        for (int j = 0; j < 32; j++) c[j] = t0 + j < ntiles ? tc[(size_t)(t0 + j) * vcap + v] : 0u;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if (t0 + j < ntiles) tc[(size_t)(t0 + j) * vcap + v] = acc;
            acc += c[j];
        }
    }
    vox_n[(size_t)b * vcap + v] = acc;
}

// step 2b: slot totals -> segment starts, heaviest-first processing order, heavy/light split.  grid B, block 1024.
__global__ void __launch_bounds__(1024) k_offsets(CloudState *__restrict__ states, unsigned vcap, unsigned heavy_min,
                                                  const unsigned *__restrict__ vox_n,
                                                  unsigned *__restrict__ vox_start, unsigned *__restrict__ vox_order) {
    const int b = blockIdx.x;
    CloudState &s = states[b];
    if (s.status != 0) return;
    const unsigned V = s.V;
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < V; base += blockDim.x) {
        const unsigned v = base + tid;
        const unsigned acc = v < V ? vox_n[(size_t)b * vcap + v] : 0u;
        unsigned inc = acc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        unsigned woff = 0, total = 0;
        for (int k = 0; k < 32; k++) { const unsigned x = s_warp[k]; if (k < wid) woff += x; total += x; }
        if (v < V) vox_start[(size_t)b * (vcap + 1) + v] = s_carry + woff + inc - acc;
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (tid == 0) vox_start[(size_t)b * (vcap + 1) + V] = s_carry;
    // processing order for k_stats: voxels bucketed by floor(log2(n)), heaviest bucket first, so the long
    // sequential chains of the few huge voxels (up to ~25 % of a LiDAR scan in one cell) start first
    __shared__ unsigned s_hist[32], s_cursor[32];
    if (tid < 32) { s_hist[tid] = 0; s_cursor[tid] = 0; }
    __syncthreads();
    for (unsigned v = tid; v < V; v += blockDim.x) atomicAdd(&s_hist[__clz(vox_n[(size_t)b * vcap + v] | 1u)], 1u);
    __syncthreads();
    if (tid == 0) {
        unsigned run = 0;
        for (int k = 0; k < 32; k++) { const unsigned c = s_hist[k]; s_hist[k] = run; run += c; }
        s.n_heavy = s_hist[__clz(heavy_min) + 1];        // keys 0..clz(heavy_min) hold the voxels with n >= heavy_min (a power of two <= kHeavyVoxel)
    }
    __syncthreads();
    for (unsigned v = tid; v < V; v += blockDim.x) {
        const int key = __clz(vox_n[(size_t)b * vcap + v] | 1u);
        vox_order[(size_t)b * vcap + s_hist[key] + atomicAdd(&s_cursor[key], 1u)] = v;
    }
    // the heavy voxels in exact descending size (ties by voxel slot): k_stats runs eight consecutive entries per warp for as
    // many steps as the largest of them has points
    __shared__ unsigned s_hv[kMaxSortedHeavy], s_hn[kMaxSortedHeavy];
    __syncthreads();
    const unsigned nh = s.n_heavy;
    if (nh > kMaxSortedHeavy) return;                    // (more than 1024 voxels of >= kHeavyVoxel points: leave them bucketed)
    if ((unsigned)tid < nh) { const unsigned v = vox_order[(size_t)b * vcap + tid]; s_hv[tid] = v; s_hn[tid] = vox_n[(size_t)b * vcap + v]; }
    __syncthreads();
    if ((unsigned)tid < nh) {
        const unsigned v = s_hv[tid], n = s_hn[tid];
        unsigned rank = 0;
        for (unsigned t = 0; t < nh; t++) rank += (s_hn[t] > n || (s_hn[t] == n && s_hv[t] < v)) ? 1u : 0u;
        vox_order[(size_t)b * vcap + rank] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// K6  stable voxel assignment, step 3: scatter the points into voxel-major, ascending-index order;
// vote labels into the per-voxel histogram.  grid (chunks, B), block 256.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_scatter(const T *__restrict__ pts, const uint16_t *__restrict__ labels, int label_bytes, long N,
                                                 const CloudState *__restrict__ states, unsigned vcap, int ntiles,
                                                 const unsigned *__restrict__ slot_rank, const unsigned *__restrict__ tile_cnt,
                                                 const unsigned *__restrict__ vox_start, T *__restrict__ sorted,
                                                 unsigned *__restrict__ hist, int nbins) {
    const int b = blockIdx.y;
    if (states[b].status != 0) return;
    // one point per thread: four points per thread (16-byte loads) measured SLOWER here (0.61 vs 0.49 ms per 512 scans) -
    // the kernel lives on the number of independent gather -> store chains in flight, which is the thread count
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned packed = slot_rank[(size_t)b * N + i];
    if (packed == kDropped) return;
    const unsigned slot = packed / (unsigned)kRankTile, rank = packed % (unsigned)kRankTile;
    const int tile = (int)(i / kRankTile);
    // position in the voxel-major order: the voxel's segment start + its points in earlier tiles + the rank within the tile
    const unsigned pos = vox_start[(size_t)b * (vcap + 1) + slot] + tile_cnt[((size_t)b * ntiles + tile) * vcap + slot] + rank;
    const T *p = pts + ((size_t)b * N + i) * 3;
    // labels: uint16 per point (the reference's dtype) or, with NDNET_B200_LABELS_U8, one byte per point
    const unsigned l = !labels ? 0u : (label_bytes == 1 ? (unsigned)reinterpret_cast<const uint8_t *>(labels)[(size_t)b * N + i]
                                                        : (unsigned)labels[(size_t)b * N + i]);
    store_sorted<T>(sorted + ((size_t)b * N + pos) * kSortedStride, p[0], p[1], p[2], l);
    // wide label sets (> kSmemLabelBins classes) vote through global atomics; small ones are counted by the statistics kernels
    if (labels && hist && l < (unsigned)nbins) atomicAdd(&hist[((size_t)b * vcap + slot) * nbins + l], 1u);
}

// ------------------------------------------------------------------------------------------------
// K7  per-voxel statistics: the sequential recurrence of normal_distributions.c:76-104 over the voxel's
// points in ascending index order (bit-exact; the off-diagonal is order dependent, A6), plus the label
// vote (:107-121).
//
// Heavy voxels (k_stats): FOUR LANES PER VOXEL, eight voxels per warp.  Lane j < 3 of a voxel owns dimension j: its mean
// chain  mu_j += (x_j - mu_j) / i  (the only loop-carried fp64 dependency), its m2_j, and one off-diagonal sum
// (lane 0: c01, lane 1: c12, lane 2: c02), whose second factor comes from a sibling lane by one shuffle per point; lane 3
// reads the label lane of the same 16-byte records and counts the votes.  Every lane of the warp is at the same point count,
// so the reciprocal pair of the count is one uniform load per step.  Nothing is staged in shared memory and no lane waits
// for another voxel: a warp issues ~40 instructions per step for its eight voxels and runs as many steps as its largest
// voxel has points (the voxels are sorted by size, k_offsets, so the eight of a warp are alike).
// Light voxels (k_stats_light): one thread per voxel.
// ------------------------------------------------------------------------------------------------

// RN(d / cnt) for an integer-valued cnt in [1, 2^26) without a division on the dependent chain:
// rh = RN(1/cnt), rl = RN((1 - cnt rh) rh) give rh + rl = (1/cnt)(1 + e), |e| < 2^-104, so the single
// rounding of fma(d, rh, RN(d rl)) rounds a value within 2^-103 (relative) of d/cnt.  d/cnt (53-bit
// numerator over an integer < 2^26) is either exactly representable or at least 2^-80 (relative) away
// from every rounding boundary, hence the result is the correctly rounded quotient.  Zeros, operands
// whose low product could underflow and huge operands take the IEEE division.
__device__ __forceinline__ double div_by_count(double d, double cnt) {
    const double ad = fabs(d);
    if (ad > 1e-250 && ad < 1e290) {
        const double rh = 1.0 / cnt;
        const double rl = fma(-cnt, rh, 1.0) * rh;
        return fma(d, rh, d * rl);
    }
    return d / cnt;
}

// magnitude tests on the high word (exponent and top mantissa bits) of a double: integer pipe, no fp64 compare
__device__ __forceinline__ unsigned hi_abs(double u) { return (unsigned)__double2hiint(u) & 0x7fffffffu; }
__device__ __forceinline__ bool exp_all_ones(double u) { return (hi_abs(u) & 0x7ff00000u) == 0x7ff00000u; }
// 2^-830 <= |u| < 2^963 (inside the range (1e-250, 1e290) for which div_by_count's reciprocal form is proven) or exactly
// zero: fma(+-0, rh, +-0 * rl) is the correctly signed zero quotient, so zeros need no division either (a coordinate that is
// constant over a voxel - flat ground - makes every difference and product of that axis zero)
__device__ __forceinline__ bool recip_ok(double u) {
    const unsigned h = hi_abs(u);
    return h - 0x0C100000u < 0x70100000u || (h | (unsigned)__double2loint(u)) == 0u;
}
// the negation of the same test, OR-ed into an accumulator without a branch (four integer instructions; doubling the high
// word drops the sign bit): acc becomes non-zero when u is neither zero nor inside the proven range
__device__ __forceinline__ void note_recip_unsafe(double u, unsigned &acc) {
    const unsigned hi = (unsigned)__double2hiint(u), lo = (unsigned)__double2loint(u);
    asm("{\n\t.reg .pred p;\n\t.reg .u32 t, z;\n\t"
        "add.u32 t, %1, %1;\n\tsub.u32 t, t, 0x18200000;\n\tsetp.ge.u32 p, t, 0xE0200000;\n\t"
        "and.b32 z, %1, 0x7fffffff;\n\tor.b32 z, z, %2;\n\t@p or.b32 %0, %0, z;\n\t}"
        : "+r"(acc) : "r"(hi), "r"(lo));
}

// {RN(1/c), RN((1 - c RN(1/c)) RN(1/c))} for c = 1 .. n (index c - 1): the reciprocal pairs of div_by_count, tabulated
// once per workspace (every voxel of every cloud divides by the same running counts)
__global__ void k_fill_recip(double2 *__restrict__ tab, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double c = (double)(i + 1);
        const double rh = 1.0 / c;
        tab[i] = make_double2(rh, fma(-c, rh, 1.0) * rh);
    }
}

template <typename T> __device__ __forceinline__ unsigned record_label(T v);
template <> __device__ __forceinline__ unsigned record_label<float>(float v) { return __float_as_uint(v); }
template <> __device__ __forceinline__ unsigned record_label<double>(double v) { return (unsigned)__double_as_longlong(v); }
// the coordinate a lane computes on.  Label lanes compute too (their results are never written): the label bits of a float
// record widen to a harmless normal double, those of a double record are a denormal and would take the rare paths
template <typename T> __device__ __forceinline__ double record_coord(T v, bool label_lane);
template <> __device__ __forceinline__ double record_coord<float>(float v, bool) { return (double)v; }
template <> __device__ __forceinline__ double record_coord<double>(double v, bool label_lane) { return label_lane ? 0.0 : v; }

constexpr int kStatsVoxelsPerWarp = 8;          // four lanes per voxel
constexpr int kStatsUnroll = 16;                // steps per straight-line block
#ifndef NDT_STATS_STAGE
#define NDT_STATS_STAGE 64
#endif
constexpr int kStatsStage = NDT_STATS_STAGE;    // steps per TMA stage: that many records of each of the 8 voxels + as many reciprocal pairs
constexpr int kStatsStages = 2;                 // stages in the ring: the operands are requested one to two stages before their use

// Shared memory of one warp of k_stats: a ring of stages filled by 1-D bulk copies (TMA), the label counters, the barriers.
// The records of a voxel are padded by 16 bytes per stage so that the 4-byte reads of the 32 lanes (8 voxels x {x, y, z,
// label}) fall into 32 different banks.
template <typename T> struct StatsSmem {
    static constexpr int kVoxelElems = kStatsStage * kSortedStride + 16 / (int)sizeof(T);
    static constexpr unsigned kVoxelBytes = kStatsStage * kSortedStride * sizeof(T);
    static constexpr unsigned kRecipBytes = kStatsStage * sizeof(double2);
    alignas(128) T recs[kStatsStages][kStatsVoxelsPerWarp][kVoxelElems];
    alignas(16) double2 rc[kStatsStages][kStatsStage];
    alignas(8) unsigned long long bar[kStatsStages];
    unsigned hist[(kSmemLabelBins + 1) * kStatsVoxelsPerWarp];        // [class][voxel of the warp]; row vote_bins = out of range
    unsigned dump[32];                                                // increments of the lanes that carry no label
};

namespace sptx {
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit; bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
}  // namespace sptx

// m ? a : b on the bit patterns (m = all ones or zero, loop invariant per lane): two LOP3, no predicate to rebuild per step
__device__ __forceinline__ double blend_bits(double a, double b, unsigned m) {
    const unsigned lo = ((unsigned)__double2loint(a) & m) | ((unsigned)__double2loint(b) & ~m);
    const unsigned hi = ((unsigned)__double2hiint(a) & m) | ((unsigned)__double2hiint(b) & ~m);
    return __hiloint2double((int)hi, (int)lo);
}

// One step of the recurrence for the warp's eight voxels:
//   d = x - mu ; mu' = mu + RN(d / i) ; e = x - mu' ; m2 += d e                              (normal_distributions.c:76-89)
//   c_jk += RN((x_j - mu_j')(x_k - mu_k) / i), j < k: mu_k is not yet updated when dimension j runs      (:91-104)
// Lane roles: lane 0 sends e0 and receives d1 (c01 = e0 d1), lane 1 sends d1 and receives d2 (c12 = e1 d2), lane 2 sends d2
// and receives e0 (c02 = d2 e0).
// Fast form: straight-line, the quotients in reciprocal form whatever the operands; it only RECORDS whether an operand left
// the range the reciprocal form is proven for.  The caller then redoes the block with the careful form.
// The step is split in two so that the source order already is the software pipeline: the chain part of step i (the only
// loop-carried latency) is followed by the tail of step i-1, whose operands have arrived by then.
__device__ __forceinline__ void stats_chain_fast(double x, double rh, double rl, int src, unsigned m_j0,
                                                 double &mu, double &d, double &e, double &rcv) {
    d = x - mu;
    mu = mu + fma(d, rh, d * rl);
    e = x - mu;
    rcv = __shfl_sync(0xffffffffu, blend_bits(e, d, m_j0), src);
}
// kTextbook (NDNET_B200_TEXTBOOK_KL): the off-diagonal sum takes the products themselves and is divided by the final count
// once, at the end - the population covariance - instead of the legacy per-step division by the running count.
template <bool kTextbook>
__device__ __forceinline__ void stats_tail_fast(double d, double e, double rcv, double rh, double rl, unsigned m_j2,
                                                double &m2, double &c, unsigned &bad) {
    m2 = m2 + d * e;
    const double pr = blend_bits(d, e, m_j2) * rcv;
    c = c + (kTextbook ? pr : fma(pr, rh, pr * rl));
    note_recip_unsafe(d, bad);
    if (!kTextbook) note_recip_unsafe(pr, bad);
}

// out-of-line IEEE divisions for operands outside the reciprocal form's proven range (tiny, huge, non-finite): rare
static __device__ __noinline__ double quotient_rare(double u, double cnt) { return u / cnt; }

// Careful form: IEEE division wherever the reciprocal form is not proven, NaN -> 0 on the off-diagonal sum
// (normal_distributions.c:98-100; a sum that is not NaN stays not NaN under a finite addend, so the fast form needs no test).
template <bool kTextbook>
__device__ __forceinline__ void stats_step_careful(double x, double rh, double rl, double cnt, int src, unsigned m_j0, unsigned m_j2,
                                                   double &mu, double &m2, double &c) {
    const double d = x - mu;
    mu = mu + (recip_ok(d) ? fma(d, rh, d * rl) : quotient_rare(d, cnt));
    const double e = x - mu;
    m2 = m2 + d * e;
    const double rcv = __shfl_sync(0xffffffffu, blend_bits(e, d, m_j0), src);
    const double pr = blend_bits(d, e, m_j2) * rcv;
    c = c + (kTextbook ? pr : (recip_ok(pr) ? fma(pr, rh, pr * rl) : quotient_rare(pr, cnt)));
    if (exp_all_ones(c) && c != c) c = 0.0;
}

// a voxel's last point was just added: lane j < 3 writes its mean, its variance m2 / n (normal_distributions.c:86-89, NaN -> 0)
// and its off-diagonal sum (mirrored); the label lane writes the vote (most frequent class, lowest index on ties, 0 when
// nothing was counted, :107-121)
template <bool kTextbook>
__device__ __forceinline__ void stats_finish(double *__restrict__ mean, double *__restrict__ cov, uint16_t *__restrict__ cls, size_t slot,
                                             int j, int q, unsigned n, double mu, double m2, double c, const unsigned *s_hist, int vote_bins) {
    if (j < 3) {
        double var = m2 / (double)n;
        if (var != var) var = 0.0;
        if (kTextbook) { c = c / (double)n; if (c != c) c = 0.0; }
        mean[slot * 3 + j] = mu;
        double *co = cov + slot * 9;
        co[j * 4] = var;
        const int a = j == 0 ? 1 : (j == 1 ? 5 : 2);        // c01 -> [0][1], c12 -> [1][2], c02 -> [0][2]
        const int m = j == 0 ? 3 : (j == 1 ? 7 : 6);
        co[a] = c; co[m] = c;
    } else if (vote_bins > 0) {
        unsigned best = 0, bc = 0;
        for (int t = 0; t < vote_bins; t++) { const unsigned x = s_hist[t * kStatsVoxelsPerWarp + q]; if (x > best) { best = x; bc = (unsigned)t; } }
        cls[slot] = (uint16_t)(best > 0 ? bc : 0u);
    }
}

// K7 (heavy voxels).  grid (B, ceil(max heavy voxels / 8)), block 32: warp g of a cloud takes entries 8g .. 8g+7 of vox_order
// (descending size), so the first wave of CTAs holds every cloud's largest voxels.  The operands never pass through
// registers on their way in: the warp asks the TMA unit for the next 32 records of each live voxel and the 32 reciprocal
// pairs of those counts (nine bulk copies per stage, the next stage in flight while this one is consumed), waits on the stage's mbarrier and reads the
// step's operands with two shared-memory loads.  A finished voxel's lanes keep computing on whatever follows in the ring
// (finite numbers; their results are not used).  vote_bins > 0: the label vote is taken here too (shared-memory counters,
// <= kSmemLabelBins classes).
template <typename T, bool kTextbook>
__global__ void __launch_bounds__(32, NDT_STATS_MIN_CTAS) k_stats(const CloudState *__restrict__ states, unsigned vcap, long N,
                                              const T *__restrict__ sorted, const unsigned *__restrict__ vox_start,
                                              const unsigned *__restrict__ vox_order, const double2 *__restrict__ recip,
                                              double *__restrict__ mean, double *__restrict__ cov,
                                              uint16_t *__restrict__ cls, int vote_bins) {
    const int b = blockIdx.x;
    const CloudState &s = states[b];
    if (s.status != 0) return;
    const unsigned n_heavy = s.n_heavy;
    const unsigned first = blockIdx.y * kStatsVoxelsPerWarp;
    if (first >= n_heavy) return;
    const int lane = threadIdx.x, q = lane >> 2, j = lane & 3;
    using Smem = StatsSmem<T>;
    __shared__ Smem sm;
    {
        // the ring starts as zeros: a lane whose voxel has ended (or that has none) must find finite numbers in it
        uint4 *z = reinterpret_cast<uint4 *>(&sm);
        for (int i = lane; i < (int)(sizeof(Smem) / sizeof(uint4)); i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < kStatsStages; t++) sptx::mbar_init(&sm.bar[t], 1);
        sptx::fence_barrier_init();
    }
    sptx::fence_proxy_async_smem();                  // the zero fill (generic proxy) before the bulk copies (async proxy)
    __syncwarp();
    unsigned v = 0, n = 0, st = 0;
    if (first + q < n_heavy) {
        v = vox_order[(size_t)b * vcap + first + q];
        st = vox_start[(size_t)b * (vcap + 1) + v];
        n = vox_start[(size_t)b * (vcap + 1) + v + 1] - st;
    }
    const unsigned nmax = __reduce_max_sync(0xffffffffu, n);           // steps of the warp (uniform)
    const unsigned nstages = (nmax + kStatsStage - 1) / kStatsStage;
    const T *const pv = sorted + ((size_t)b * N + st) * kSortedStride;  // the voxel's records
    // stage t -> buffer t % kStatsStages: 32 records of every voxel that still has points at step 32 t, and the reciprocals
    auto request = [&](unsigned t) {
        const unsigned kb = t * kStatsStage, buf = t % kStatsStages;
        const bool mine = j == 0 && kb < n;
        const unsigned live = __popc(__ballot_sync(0xffffffffu, mine));
        if (lane == 0) sptx::mbar_expect_tx(&sm.bar[buf], live * Smem::kVoxelBytes + Smem::kRecipBytes);
        __syncwarp();
        if (mine) sptx::bulk_load(&sm.recs[buf][q][0], pv + (size_t)kb * kSortedStride, Smem::kVoxelBytes, &sm.bar[buf]);
        if (lane == 0) sptx::bulk_load(&sm.rc[buf][0], recip + kb, Smem::kRecipBytes, &sm.bar[buf]);
    };
    for (unsigned t = 0; t < (unsigned)kStatsStages && t < nstages; t++) request(t);

    const int src = (lane & ~3) | (j == 0 ? 1 : (j == 1 ? 2 : 0));
    const unsigned m_j0 = j == 0 ? 0xffffffffu : 0u, m_j2 = j == 2 ? 0xffffffffu : 0u;
    // the vote is branch-free (a branch per step would cut the unrolled block into pieces the scheduler cannot overlap):
    // label lanes count class min(label, vote_bins) in their voxel's column, the other lanes increment a word of their own
    unsigned *const my_hist = j == 3 ? sm.hist + q : sm.dump + lane;
    const unsigned my_bins = (unsigned)vote_bins, my_stride = j == 3 ? (unsigned)kStatsVoxelsPerWarp : 0u;
    double mu = 0.0, m2 = 0.0, c = 0.0;
    for (unsigned t = 0; t < nstages; t++) {
        const unsigned buf = t % kStatsStages;
        sptx::mbar_wait(&sm.bar[buf], (t / kStatsStages) & 1u);
#pragma unroll 1
        for (int h = 0; h < kStatsStage / kStatsUnroll; h++) {
            const unsigned k0 = t * kStatsStage + h * kStatsUnroll;
            if (k0 >= nmax) break;
            const T *xb = &sm.recs[buf][q][h * kStatsUnroll * kSortedStride + j];
            const double2 *rb = &sm.rc[buf][h * kStatsUnroll];
            if (!__any_sync(0xffffffffu, n - k0 - 1u < (unsigned)kStatsUnroll)) {
                const double mu0 = mu, m20 = m2, c0 = c;
                unsigned bad = 0u;
                double pd = 0.0, pe = 0.0, prcv = 0.0, prh = 0.0, prl = 0.0;     // the previous step's tail operands
#pragma unroll
                for (int i = 0; i < kStatsUnroll; i++) {
                    const T raw = xb[i * kSortedStride];
                    const double2 r = rb[i];
                    double d, e, rcv;
                    stats_chain_fast(record_coord<T>(raw, j == 3), r.x, r.y, src, m_j0, mu, d, e, rcv);
                    if (i > 0) stats_tail_fast<kTextbook>(pd, pe, prcv, prh, prl, m_j2, m2, c, bad);
                    atomicAdd(my_hist + min(record_label<T>(raw), my_bins) * my_stride, 1u);
                    pd = d; pe = e; prcv = rcv; prh = r.x; prl = r.y;
                }
                stats_tail_fast<kTextbook>(pd, pe, prcv, prh, prl, m_j2, m2, c, bad);
                if (__any_sync(0xffffffffu, bad != 0u)) {      // rare: redo the block (its operands are still in the ring)
                    mu = mu0; m2 = m20; c = c0;
                    for (int i = 0; i < kStatsUnroll; i++) {
                        const double2 r = rb[i];
                        stats_step_careful<kTextbook>(record_coord<T>(xb[i * kSortedStride], j == 3), r.x, r.y, (double)(k0 + i + 1), src, m_j0, m_j2, mu, m2, c);
                    }
                }
            } else {
                // some voxel of the warp ends inside this block: one step at a time, each lane looking for its last point
                const double mu0 = mu, m20 = m2, c0 = c;
                unsigned bad = 0u;
#pragma unroll 1
                for (int i = 0; i < kStatsUnroll; i++) {
                    const T raw = xb[i * kSortedStride];
                    const double2 r = rb[i];
                    double d, e, rcv;
                    stats_chain_fast(record_coord<T>(raw, j == 3), r.x, r.y, src, m_j0, mu, d, e, rcv);
                    stats_tail_fast<kTextbook>(d, e, rcv, r.x, r.y, m_j2, m2, c, bad);
                    atomicAdd(my_hist + min(record_label<T>(raw), my_bins) * my_stride, 1u);   // (after its end a voxel's counters are dead)
                    if (k0 + i + 1 == n) stats_finish<kTextbook>(mean, cov, cls, (size_t)b * vcap + v, j, q, n, mu, m2, c, sm.hist, vote_bins);
                }
                if (__any_sync(0xffffffffu, bad != 0u)) {      // rare: again with the careful form (the votes are in already)
                    mu = mu0; m2 = m20; c = c0;
                    for (int i = 0; i < kStatsUnroll; i++) {
                        const double2 r = rb[i];
                        stats_step_careful<kTextbook>(record_coord<T>(xb[i * kSortedStride], j == 3), r.x, r.y, (double)(k0 + i + 1), src, m_j0, m_j2, mu, m2, c);
                        if (k0 + i + 1 == n) stats_finish<kTextbook>(mean, cov, cls, (size_t)b * vcap + v, j, q, n, mu, m2, c, sm.hist, 0);
                    }
                }
            }
        }
        __syncwarp();                                       // every lane is done with the buffer
        if (t + kStatsStages < nstages) request(t + kStatsStages);
    }
}

// Statistics of the light voxels (fewer than kHeavyVoxel points: about 1040 of a scan's 1060 voxels, half of its
// points): one THREAD per voxel runs the literal recurrence of normal_distributions.c:76-104, so a warp advances 32
// voxels per instruction instead of one.  vox_order is bucketed by size, so the voxels of a warp have similar
// lengths.  The divisions by the running count share one reciprocal pair per count, tabulated once per CTA (every thread
// of a warp is at the same count; see div_by_count for why the reciprocal form is the correctly rounded quotient).
// vote_bins > 0: the label vote (normal_distributions.c:107-121) is taken here too, in per-thread 16-bit counters.
// grid (ceil(vcap/128), B), block 128, dynamic smem vote_bins * 128 * 2 bytes.
template <typename T, bool kTextbook>
__global__ void __launch_bounds__(128) k_stats_light(const CloudState *__restrict__ states, unsigned vcap, long N,
                                                     const T *__restrict__ sorted, const unsigned *__restrict__ vox_start,
                                                     const unsigned *__restrict__ vox_order,
                                                     double *__restrict__ mean, double *__restrict__ cov,
                                                     uint16_t *__restrict__ cls, int vote_bins) {
    extern __shared__ unsigned short s_votes[];          // [vote_bins][128]
    __shared__ double2 s_rc[kHeavyVoxel];                // {rh, rl} ~ 1 / (k + 1)
    const int b = blockIdx.y;
    const CloudState &s = states[b];
    if (s.status != 0) return;
    const unsigned first = s.n_heavy + blockIdx.x * blockDim.x;
    if (first >= s.V) return;
    const int tid = threadIdx.x;
    {
        // vox_order lists the voxels heaviest bucket first, so the first voxel of the CTA bounds every count it will see
        const unsigned v0 = vox_order[(size_t)b * vcap + first];
        const unsigned n0 = vox_start[(size_t)b * (vcap + 1) + v0 + 1] - vox_start[(size_t)b * (vcap + 1) + v0];
        unsigned bound = 1u << (32 - __clz(n0 | 1u));                       // counts of this bucket are < bound
        if (bound > kHeavyVoxel) bound = kHeavyVoxel;
        for (unsigned k = tid; k < bound; k += blockDim.x) {
            const double c = (double)(k + 1);
            const double rh = 1.0 / c;
            s_rc[k] = make_double2(rh, fma(-c, rh, 1.0) * rh);
        }
        for (int j = 0; j < vote_bins; j++) s_votes[j * 128 + tid] = 0;
    }
    __syncthreads();
    const unsigned idx = first + tid;
    if (idx >= s.V) return;
    const unsigned v = vox_order[(size_t)b * vcap + idx];
    const unsigned st = vox_start[(size_t)b * (vcap + 1) + v], en = vox_start[(size_t)b * (vcap + 1) + v + 1];
    const T *p = sorted + ((size_t)b * N + st) * kSortedStride;
    double mu0 = 0, mu1 = 0, mu2 = 0, m20 = 0, m21 = 0, m22 = 0, c01 = 0, c02 = 0, c12 = 0;
    const unsigned n = en - st;
    T a0 = 0, a1 = 0, a2 = 0;
    unsigned lab = 0;
    if (n > 0) load_sorted_rec<T>(p, a0, a1, a2, lab);
    for (unsigned k = 0; k < n; k++) {
        const double x0 = (double)a0, x1 = (double)a1, x2 = (double)a2;
        if (vote_bins > 0 && lab < (unsigned)vote_bins) s_votes[lab * 128 + tid]++;
        if ((k & 7u) == 0 && k + 8 < n) prefetch_l2(p + (size_t)(k + 8) * kSortedStride);   // the next 128-byte line of this voxel
        if (k + 1 < n) load_sorted_rec<T>(p + (size_t)(k + 1) * kSortedStride, a0, a1, a2, lab);   // overlaps this point's chain
        const double2 r = s_rc[k];
        const double rh = r.x, rl = r.y;
        // RN(u / c): reciprocal form inside its proven range, IEEE division otherwise
#define QDIV(u) (recip_ok(u) ? fma((u), rh, (u) * rl) : (u) / (double)(k + 1))
        // off-diagonal addend: the legacy quotient by the running count, or (textbook mode) the product itself
#define CDIV(u) (kTextbook ? (u) : QDIV(u))
        // j = 0
        const double d0 = x0 - mu0;
        const double o0 = mu0;
        mu0 = mu0 + QDIV(d0);
        const double e0 = x0 - mu0;
        m20 += (x0 - o0) * e0;
        const double f1 = x1 - mu1, f2 = x2 - mu2;                                                  // mu1, mu2 still old
        { const double u = e0 * f1; c01 += CDIV(u); if (exp_all_ones(c01) && c01 != c01) c01 = 0.0; }
        { const double u = e0 * f2; c02 += CDIV(u); if (exp_all_ones(c02) && c02 != c02) c02 = 0.0; }
        // j = 1
        const double o1 = mu1;
        mu1 = mu1 + QDIV(f1);
        const double e1 = x1 - mu1;
        m21 += (x1 - o1) * e1;
        { const double u = e1 * f2; c12 += CDIV(u); if (exp_all_ones(c12) && c12 != c12) c12 = 0.0; }
        // j = 2
        const double o2 = mu2;
        mu2 = mu2 + QDIV(f2);
        m22 += (x2 - o2) * (x2 - mu2);
#undef CDIV
#undef QDIV
    }
    const double cn = (double)n;
    double v0 = m20 / cn, v1 = m21 / cn, v2 = m22 / cn;
    if (v0 != v0) v0 = 0.0;
    if (v1 != v1) v1 = 0.0;
    if (v2 != v2) v2 = 0.0;
    if (kTextbook) {
        c01 = c01 / cn; c02 = c02 / cn; c12 = c12 / cn;
        if (c01 != c01) c01 = 0.0;
        if (c02 != c02) c02 = 0.0;
        if (c12 != c12) c12 = 0.0;
    }
    double *mo = mean + ((size_t)b * vcap + v) * 3;
    mo[0] = mu0; mo[1] = mu1; mo[2] = mu2;
    double *co = cov + ((size_t)b * vcap + v) * 9;
    co[0] = v0; co[1] = c01; co[2] = c02; co[3] = c01; co[4] = v1; co[5] = c12; co[6] = c02; co[7] = c12; co[8] = v2;
    if (vote_bins > 0) {
        unsigned best = 0, bc = 0;
        for (int j = 0; j < vote_bins; j++) { const unsigned x = s_votes[j * 128 + tid]; if (x > best) { best = x; bc = (unsigned)j; } }
        cls[(size_t)b * vcap + v] = (uint16_t)(best > 0 ? bc : 0u);
    }
}

// Label vote (normal_distributions.c:107-121): most frequent class of the voxel, lowest index on ties.  One warp
// per voxel over the label lane of its sorted point records, counts in shared memory; wide label sets were counted by k_scatter's global
// atomics and only need the arg-max here.  grid (B, ceil(vcap/4)), block 128.
template <typename T>
__global__ void __launch_bounds__(128) k_votes(const CloudState *__restrict__ states, unsigned vcap, long N,
                                               const T *__restrict__ sorted, const unsigned *__restrict__ vox_start,
                                               const unsigned *__restrict__ hist, int nbins, uint16_t *__restrict__ cls) {
    const int b = blockIdx.x;
    const CloudState &s = states[b];
    if (s.status != 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned v = blockIdx.y * 4 + warp;
    if (v >= s.V) return;
    __shared__ unsigned s_h[4][kSmemLabelBins];
    const unsigned *h;
    if (hist) {
        h = hist + ((size_t)b * vcap + v) * nbins;
    } else {
        for (int j = lane; j < kSmemLabelBins; j += 32) s_h[warp][j] = 0;
        __syncwarp();
        const unsigned st = vox_start[(size_t)b * (vcap + 1) + v], en = vox_start[(size_t)b * (vcap + 1) + v + 1];
        const T *sl = sorted + (size_t)b * N * kSortedStride;
        for (unsigned i = st + lane; i < en; i += 32) {
            const unsigned l = sorted_label(sl + (size_t)i * kSortedStride);
            if (l < (unsigned)nbins) atomicAdd(&s_h[warp][l], 1u);
        }
        __syncwarp();
        h = s_h[warp];
    }
    unsigned best = 0; int bc = 0x7fffffff;
    for (int j = lane; j < nbins; j += 32) { const unsigned x = h[j]; if (x > best) { best = x; bc = j; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        if (ob > best || (ob == best && oc < bc)) { best = ob; bc = oc; }
    }
    if (lane == 0) cls[(size_t)b * vcap + v] = (uint16_t)(best > 0 ? bc : 0);
}

// ------------------------------------------------------------------------------------------------
// K8  neighbour pseudo-KL (kullback_leibler.c:28-202) in closed "event index" form.
//
// The reference LU-factorises both covariances IN PLACE on every call where both voxels have more
// than one sample (A10), while traversing voxels in ascending index and directions X+,X-,Y+,Y-,Z+,Z-.
// So the matrix a voxel contributes to a given call is LU^k(Sigma), k = number of qualifying calls
// that touched it earlier.  For voxel A the touch order is: as q of its Z-, Y-, X- neighbours, as p in
// its own six directions, as q of its X+, Y+, Z+ neighbours.  One thread per voxel replays A's own
// chain and, for each of its calls, the neighbour's chain up to the needed index.  It also emits the
// final (mangled) covariance that to_point_cloud exports (ndt.c:107).
// grid (ceil(vcap/64), B), block 64.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int neighbor_slot(const uint2 *bm, unsigned idx, int lx, int ly, int lz, int dir) {
    // voxel.c:116-175 (unsigned wrap at 0 included)
    unsigned z = idx / (unsigned)(lx * ly), y = (idx % (unsigned)(lx * ly)) / (unsigned)lx, x = idx % (unsigned)lx;
    switch (dir) {
        case 0: x += 1u; break;
        case 1: x -= 1u; break;
        case 2: y += 1u; break;
        case 3: y -= 1u; break;
        case 4: z += 1u; break;
        default: z -= 1u; break;
    }
    if (x >= (unsigned)lx || y >= (unsigned)ly || z >= (unsigned)lz) return -1;
    const unsigned n = z * (unsigned)lx * (unsigned)ly + y * (unsigned)lx + x;
    const uint2 w = bm[n >> 5];
    if (!((w.x >> (n & 31)) & 1u)) return -1;
    return (int)(w.y + __popc(w.x & ((1u << (n & 31)) - 1u)));
}

__global__ void __launch_bounds__(64) k_kl(const CloudState *__restrict__ states, unsigned vcap,
                                           const uint2 *__restrict__ bitmap, size_t bitmap_stride,
                                           const unsigned *__restrict__ vox_cell, const unsigned *__restrict__ vox_n,
                                           const double *__restrict__ cov, double *__restrict__ cov_final,
                                           double *__restrict__ kl_div, unsigned char *__restrict__ kl_flag) {
    const int b = blockIdx.y;
    const CloudState &s = states[b];
    if (s.status != 0) return;
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= s.V) return;
    const uint2 *bm = bitmap + (size_t)b * bitmap_stride;
    const unsigned *vn = vox_n + (size_t)b * vcap;
    const unsigned *vc = vox_cell + (size_t)b * vcap;
    const double *cv = cov + (size_t)b * vcap * 9;
    const int lx = s.len[0], ly = s.len[1], lz = s.len[2];
    const unsigned cell = vc[v];
    const unsigned nA = vn[v];

    int nb[kDirs]; unsigned nn[kDirs];
#pragma unroll
    for (int d = 0; d < kDirs; d++) {
        nb[d] = neighbor_slot(bm, cell, lx, ly, lz, d);
        nn[d] = nb[d] >= 0 ? vn[nb[d]] : 0u;
    }
    double A[9];
#pragma unroll
    for (int i = 0; i < 9; i++) A[i] = cv[(size_t)v * 9 + i];
    double *out_div = kl_div + ((size_t)b * vcap + v) * kDirs;
    unsigned char *out_flag = kl_flag + ((size_t)b * vcap + v) * kDirs;

    if (nA > 1) {
        // calls that reached A as q before its own turn: from Z- (dir 5), Y- (dir 3), X- (dir 1) neighbours
        const int pre = (nn[5] > 1) + (nn[3] > 1) + (nn[1] > 1);
        int perm, sgn;
        for (int k = 0; k < pre; k++) lu3(A, perm, sgn);
    }
#pragma unroll 1
    for (int d = 0; d < kDirs; d++) {
        if (nb[d] < 0) { out_flag[d] = 0; continue; }
        if (nA <= 1 || nn[d] <= 1) { out_div[d] = 0.0; out_flag[d] = 1; continue; }   // -1: div 0 still inserted (A11)
        int pperm, psign;
        lu3(A, pperm, psign);
        // neighbour B's chain
        const unsigned bslot = (unsigned)nb[d];
        const unsigned bcell = vc[bslot];
        int qual[kDirs];
#pragma unroll
        for (int e = 0; e < kDirs; e++) {
            const int ns = neighbor_slot(bm, bcell, lx, ly, lz, e);
            qual[e] = (ns >= 0 && vn[ns] > 1) ? 1 : 0;
        }
        const int all = qual[0] + qual[1] + qual[2] + qual[3] + qual[4] + qual[5];
        const int early = qual[5] + qual[3] + qual[1];
        int kB;
        switch (d) {             // A lies in direction (d^1) from B
            case 4: kB = 0; break;                          // A is B's Z- neighbour
            case 2: kB = qual[5]; break;                    // A is B's Y- neighbour
            case 0: kB = qual[5] + qual[3]; break;          // A is B's X- neighbour
            case 1: kB = early + all; break;                // A is B's X+ neighbour
            case 3: kB = early + all + qual[0]; break;      // A is B's Y+ neighbour
            default: kB = early + all + qual[0] + qual[2]; break;   // A is B's Z+ neighbour
        }
        double Q[9];
#pragma unroll
        for (int i = 0; i < 9; i++) Q[i] = cv[(size_t)bslot * 9 + i];
        int qperm, qsign;
        for (int k = 0; k <= kB; k++) lu3(Q, qperm, qsign);
        double div;
        if (pseudo_kl(A, psign, Q, qperm, qsign, div)) { out_div[d] = div; out_flag[d] = 1; }
        else out_flag[d] = 0;                                // -2: pair skipped, mutation kept (A11)
    }
    if (nA > 1) {
        const int post = (nn[0] > 1) + (nn[2] > 1) + (nn[4] > 1);
        int perm, sgn;
        for (int k = 0; k < post; k++) lu3(A, perm, sgn);
    }
    double *cf = cov_final + ((size_t)b * vcap + v) * 9;
#pragma unroll
    for (int i = 0; i < 9; i++) cf[i] = A[i];
}

// K8t  textbook mode (NDNET_B200_TEXTBOOK_KL; the algorithm the reference's README describes, README.md:6): the
// Kullback-Leibler divergence of the two Gaussians themselves,
//   KL(p || q) = 1/2 [ tr(Sq^-1 Sp) + (mq - mp)^T Sq^-1 (mq - mp) - 3 + ln(det Sq / det Sp) ],
// on the population covariances (nothing is factorised in place, no call sees another call's leftovers), for every
// occupied voxel p and each of its occupied 6-neighbours q in enum order.  Pairs with fewer than two points on either
// side, a covariance that is not positive definite (det <= 0 or a non-positive leading minor) or a non-finite result
// get no entry.  cov_final = the covariance itself.  grid (ceil(vcap/64), B), block 64.
__device__ __forceinline__ bool spd_det_inv(const double S[9], double &det, double inv[9]) {
    // symmetric 3x3: cofactors, determinant by the first row, leading minors for positive definiteness
    const double a = S[0], b = S[1], c = S[2], d = S[4], e = S[5], f = S[8];
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    det = a * c00 + b * c01 + c * c02;
    if (!(a > 0.0) || !(a * d - b * b > 0.0) || !(det > 0.0)) return false;
    const double r = 1.0 / det;
    inv[0] = c00 * r; inv[1] = c01 * r; inv[2] = c02 * r;
    inv[3] = inv[1]; inv[4] = (a * f - c * c) * r; inv[5] = (b * c - a * e) * r;
    inv[6] = inv[2]; inv[7] = inv[5]; inv[8] = (a * d - b * b) * r;
    return true;
}

__global__ void __launch_bounds__(64) k_kl_textbook(const CloudState *__restrict__ states, unsigned vcap,
                                                    const uint2 *__restrict__ bitmap, size_t bitmap_stride,
                                                    const unsigned *__restrict__ vox_cell, const unsigned *__restrict__ vox_n,
                                                    const double *__restrict__ mean, const double *__restrict__ cov,
                                                    double *__restrict__ cov_final, double *__restrict__ kl_div,
                                                    unsigned char *__restrict__ kl_flag) {
    const int b = blockIdx.y;
    const CloudState &s = states[b];
    if (s.status != 0) return;
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= s.V) return;
    const uint2 *bm = bitmap + (size_t)b * bitmap_stride;
    const unsigned *vn = vox_n + (size_t)b * vcap;
    const double *cv = cov + (size_t)b * vcap * 9, *mu = mean + (size_t)b * vcap * 3;
    const int lx = s.len[0], ly = s.len[1], lz = s.len[2];
    const unsigned cell = vox_cell[(size_t)b * vcap + v];
    double P[9], Pinv[9], pdet;
#pragma unroll
    for (int i = 0; i < 9; i++) P[i] = cv[(size_t)v * 9 + i];
    double *cf = cov_final + ((size_t)b * vcap + v) * 9;
#pragma unroll
    for (int i = 0; i < 9; i++) cf[i] = P[i];
    const bool p_ok = vn[v] > 1 && spd_det_inv(P, pdet, Pinv);
    double *out_div = kl_div + ((size_t)b * vcap + v) * kDirs;
    unsigned char *out_flag = kl_flag + ((size_t)b * vcap + v) * kDirs;
#pragma unroll 1
    for (int d = 0; d < kDirs; d++) {
        out_flag[d] = 0;
        const int q = neighbor_slot(bm, cell, lx, ly, lz, d);
        if (q < 0 || !p_ok || vn[q] <= 1) continue;
        double Q[9], Qinv[9], qdet;
#pragma unroll
        for (int i = 0; i < 9; i++) Q[i] = cv[(size_t)q * 9 + i];
        if (!spd_det_inv(Q, qdet, Qinv)) continue;
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int k = 0; k < 3; k++) tr += Qinv[i * 3 + k] * P[k * 3 + i];
        const double d0 = mu[(size_t)q * 3 + 0] - mu[(size_t)v * 3 + 0], d1 = mu[(size_t)q * 3 + 1] - mu[(size_t)v * 3 + 1],
                     d2 = mu[(size_t)q * 3 + 2] - mu[(size_t)v * 3 + 2];
        const double maha = d0 * (Qinv[0] * d0 + Qinv[1] * d1 + Qinv[2] * d2) + d1 * (Qinv[3] * d0 + Qinv[4] * d1 + Qinv[5] * d2) +
                            d2 * (Qinv[6] * d0 + Qinv[7] * d1 + Qinv[8] * d2);
        const double div = 0.5 * (tr + maha - 3.0 + log(qdet / pdet));
        if (isfinite(div)) { out_div[d] = div; out_flag[d] = 1; }
    }
}

// ------------------------------------------------------------------------------------------------
// K9  divergence list: compaction in insertion order, NaN rule (A14), stable descending sort
// (kullback_leibler.c:181-195), prune walk (ndt.c:45-72) and ascending compaction (ndt.c:75-117).
// One CTA per cloud, block 1024.  Sort keys live in shared memory when the padded list fits, else in
// the global scratch.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long desc_key(double d) {
    // larger divergence -> smaller key; -0.0 and +0.0 compare equal in the reference
    if (d == 0.0) d = 0.0;
    return ~enc_f64(d);
}

__device__ __forceinline__ unsigned block_scan_step(unsigned val, unsigned *s_warp, unsigned &total) {
    // inclusive scan over the block (up to 1024 threads); returns exclusive prefix of this thread, total of block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned inc = val;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    __syncthreads();
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    unsigned woff = 0; total = 0;
    for (int k = 0; k < nwarps; k++) { const unsigned x = s_warp[k]; if (k < wid) woff += x; total += x; }
    return woff + inc - val;
}

// smallest bin whose inclusive prefix count reaches `target` (>= 1), and the count in front of it.  Every thread calls it.
__device__ __forceinline__ void find_rank_bin(const unsigned *hist, int nbins, unsigned target, unsigned *s_warp, unsigned *s_out,
                                              unsigned &bin, unsigned &before) {
    const int per = (nbins + (int)blockDim.x - 1) / (int)blockDim.x;
    const int b0 = (int)threadIdx.x * per;
    unsigned sum = 0;
    for (int k = 0; k < per; k++) if (b0 + k < nbins) sum += hist[b0 + k];
    unsigned total;
    const unsigned excl = block_scan_step(sum, s_warp, total);
    if (threadIdx.x == 0) { s_out[0] = (unsigned)nbins - 1u; s_out[1] = total - hist[nbins - 1]; }     // target beyond the total: last bin
    __syncthreads();
    if (excl < target && target <= excl + sum) {
        unsigned run = excl;
        for (int k = 0; k < per; k++) {
            const unsigned c = hist[b0 + k];
            if (run + c >= target) { s_out[0] = (unsigned)(b0 + k); s_out[1] = run; break; }
            run += c;
        }
    }
    __syncthreads();
    bin = s_out[0]; before = s_out[1];
    __syncthreads();
}

constexpr int kSelectPartialCap = 2048;         // candidates the partial selection sorts in shared memory (24 KB) at D ~ 1000 ...
constexpr int kSelectPartialCapMax = 16384;     // ... and at most, for large D (196 KB): the host sizes the launch by 6 (Vcap - D)
constexpr int kSelectPartialThreads = 256;

// kPartial: only the head of the sorted list is produced - the walk removes the first V - D distinct p's and every p owns at
// most six entries, so it never looks past entry 6 (V - D); those entries are found by a two-level radix selection on the
// keys (12 + 11 bits) and sorted alone (<= 2048 of the ~6400 entries of a scan).  The full list (legacy handles, inspection)
// is kept by the kPartial = false instantiation.
template <int kThreads, bool kPartial>
__global__ void __launch_bounds__(kThreads, kPartial ? 5 : 3) k_select(CloudState *__restrict__ states, unsigned vcap, long num_desired,
                                                 const unsigned *__restrict__ vox_cell, const unsigned *__restrict__ vox_n,
                                                 const double *__restrict__ mean, const double *__restrict__ cov_final,
                                                 const uint16_t *__restrict__ cls, int has_labels,
                                                 const double *__restrict__ kl_div, const unsigned char *__restrict__ kl_flag,
                                                 unsigned long long *__restrict__ g_key, unsigned *__restrict__ g_seq,
                                                 size_t kcap, int smem_cap, unsigned *__restrict__ g_firstpos,
                                                 unsigned char *__restrict__ g_removed, unsigned flags,
                                                 float *__restrict__ out_feat, double *__restrict__ out_feat64,
                                                 uint16_t *__restrict__ out_labels, int *__restrict__ out_voxel,
                                                 double *__restrict__ list_div, unsigned *__restrict__ list_seq,
                                                 NdtCloudInfo *__restrict__ info) {
    extern __shared__ unsigned long long s_dyn[];
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    __shared__ double s_wmin[32];
    __shared__ double s_cmin;
    const int b = blockIdx.x;
    CloudState &s = states[b];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long D = num_desired;
    float *of = out_feat ? out_feat + (size_t)b * D * 12 : nullptr;
    double *of64 = out_feat64 ? out_feat64 + (size_t)b * D * 12 : nullptr;
    uint16_t *ol = out_labels ? out_labels + (size_t)b * D : nullptr;
    int *ov = out_voxel ? out_voxel + (size_t)b * D : nullptr;

    if (s.status != 0) {
        for (long i = tid; i < D * 12; i += blockDim.x) { if (of) of[i] = 0.f; if (of64) of64[i] = 0.0; }
        for (long i = tid; i < D; i += blockDim.x) { if (ol) ol[i] = 0; if (ov) ov[i] = -1; }
        if (tid == 0 && info) {
            NdtCloudInfo &o = info[b];
            o.status = s.status; o.prune_status = 0; o.evaluations = s.evals;
            for (int a = 0; a < 3; a++) { o.len[a] = (unsigned)s.len[a]; o.offset[a] = s.off[a]; }
            o.num_voxels = 0; o.num_valid = 0; o.num_kl = 0; o.num_kl_after = 0; o.num_out = 0; o.num_survivors = 0;
            o.voxel_size = s.guess;
            for (int a = 0; a < 6; a++) o.limits[a] = s.lim[a];
        }
        return;
    }
    const unsigned V = s.V;
    const unsigned nslots = V * kDirs;
    const double *kd = kl_div + (size_t)b * vcap * kDirs;
    const unsigned char *kf = kl_flag + (size_t)b * vcap * kDirs;
    // the sort scratch is padded to a power of two per cloud (kpad >= kcap): the global-memory fallback
    // pads the list up to the next power of two
    size_t kpad = 1; while (kpad < kcap) kpad <<= 1;
    unsigned long long *gk = g_key + (size_t)b * kpad;
    unsigned *gs = g_seq + (size_t)b * kpad;

    // ---- 1. compaction in insertion order + NaN rule: a NaN takes the exclusive prefix-minimum of
    //         the finite (non-NaN) divergences inserted before it, +inf if none.
    if (tid == 0) { s_carry = 0; s_cmin = __longlong_as_double(0x7FF0000000000000ll); }
    __syncthreads();
    for (unsigned base = 0; base < nslots; base += blockDim.x) {
        const unsigned q = base + tid;
        const bool present = q < nslots && kf[q];
        const double d = present ? kd[q] : 0.0;
        const bool isn = present && isnan(d);
        // exclusive prefix min of non-NaN present values
        double mval = (present && !isn) ? d : __longlong_as_double(0x7FF0000000000000ll);
        double inc = mval;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = u < inc ? u : inc; }
        double excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = __longlong_as_double(0x7FF0000000000000ll);
        unsigned total;
        const unsigned pos_in = block_scan_step(present ? 1u : 0u, s_warp, total);   // contains __syncthreads
        if (lane == 31) s_wmin[wid] = inc;
        __syncthreads();
        double wpre = s_cmin;
        double blockmin = s_cmin;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) { const double x = s_wmin[k]; if (k < wid) wpre = x < wpre ? x : wpre; blockmin = x < blockmin ? x : blockmin; }
        excl = wpre < excl ? wpre : excl;
        if (present) {
            const unsigned pos = s_carry + pos_in;
            CHK(pos < kcap, 1);
            if (pos < kcap) {
                // legacy: descending divergence (the list head, removed first, holds the LARGEST divergences - what the
                // compiled reference does); textbook: ascending (the most redundant distributions go first, README.md:6)
                gk[pos] = (flags & 4u) ? ~desc_key(d) : desc_key(isn ? excl : d);
                gs[pos] = q;
            }
        }
        __syncthreads();
        if (tid == 0) { s_carry += total; s_cmin = blockmin; }
        __syncthreads();
    }
    const unsigned K = s_carry;
    CHK(K <= kcap && V <= vcap, 2);
    // ---- 2. stable sort (key ascending == divergence descending, then insertion sequence ascending) of the first Ks
    //         entries of that order
    unsigned Ks = K;
    bool selected = false;
    if (kPartial) {
        __shared__ unsigned s_sel[4];
        const unsigned long need64 = 6ul * (unsigned long)(V - (unsigned)D);
        const unsigned need = need64 < (unsigned long)K ? (unsigned)need64 : K;
        if (need == 0) Ks = 0;
        else if (need < K) {
            unsigned *hist = (unsigned *)s_dyn;                               // 4096 bins, then 2048 bins
            for (int i = tid; i < 4096; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            for (unsigned i = tid; i < K; i += blockDim.x) atomicAdd(&hist[(unsigned)(gk[i] >> 52)], 1u);
            __syncthreads();
            unsigned binA, beforeA;
            find_rank_bin(hist, 4096, need, s_warp, s_sel, binA, beforeA);
            for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            for (unsigned i = tid; i < K; i += blockDim.x) {
                const unsigned long long k = gk[i];
                if ((unsigned)(k >> 52) == binA) atomicAdd(&hist[(unsigned)(k >> 41) & 0x7ffu], 1u);
            }
            __syncthreads();
            unsigned binB, beforeB;
            find_rank_bin(hist, 2048, need - beforeA, s_warp, s_sel, binB, beforeB);
            const unsigned cand = beforeA + beforeB + hist[binB];
            __syncthreads();
            if ((int)cand <= smem_cap) {
                // the candidates are exactly the first `cand` entries of the sorted order (everything below a key boundary)
                unsigned Pc = 1; while (Pc < cand) Pc <<= 1;
                unsigned long long *ck = s_dyn;
                unsigned *cs = (unsigned *)(s_dyn + Pc);
                if (tid == 0) s_sel[2] = 0u;
                __syncthreads();
                for (unsigned i = tid; i < K; i += blockDim.x) {
                    const unsigned long long k = gk[i];
                    const unsigned a = (unsigned)(k >> 52), bsub = (unsigned)(k >> 41) & 0x7ffu;
                    if (a < binA || (a == binA && bsub <= binB)) { const unsigned at = atomicAdd(&s_sel[2], 1u); ck[at] = k; cs[at] = gs[i]; }
                }
                __syncthreads();
                CHK(s_sel[2] == cand, 6);
                Ks = cand;
                selected = true;
            }
        }
    }
    unsigned P = 1; while (P < Ks) P <<= 1;
    const bool in_smem = (int)P <= smem_cap;
    CHK(in_smem || P <= kpad, 3);
    unsigned long long *key = in_smem ? s_dyn : gk;
    unsigned *seq = in_smem ? (unsigned *)(s_dyn + P) : gs;
    if (in_smem && !selected) { for (unsigned i = tid; i < Ks; i += blockDim.x) { key[i] = gk[i]; seq[i] = gs[i]; } }
    for (unsigned i = Ks + tid; i < P; i += blockDim.x) { key[i] = ~0ull; seq[i] = 0xFFFFFFFFu; }
    __syncthreads();
    for (unsigned k = 2; k <= P; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = tid; i < P; i += blockDim.x) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const unsigned long long ka = key[i], kb = key[l];
                    const unsigned sa = seq[i], sb = seq[l];
                    const bool a_gt_b = ka > kb || (ka == kb && sa > sb);
                    const bool up = (i & k) == 0;
                    if (a_gt_b == up) { key[i] = kb; key[l] = ka; seq[i] = sb; seq[l] = sa; }
                }
            }
            __syncthreads();
        }
    }
    // keep the sorted list (for the legacy handles / inspection)
    if (list_div && !kPartial) {
        double *ld = list_div + (size_t)b * kcap; unsigned *ls = list_seq + (size_t)b * kcap;
        for (unsigned i = tid; i < K; i += blockDim.x) { const unsigned q = seq[i]; CHK(q < nslots, 4); ld[i] = kd[q]; ls[i] = q; }
    }
    // ---- 3. prune walk: first occurrence of each p in list order, the first to_remove of them go,
    //         subject to the shrinking-length stop (ndt.c:53)
    unsigned *firstpos = g_firstpos + (size_t)b * vcap;
    unsigned char *removed = g_removed + (size_t)b * vcap;
    for (unsigned v = tid; v < V; v += blockDim.x) { firstpos[v] = 0xFFFFFFFFu; removed[v] = 0; }
    __syncthreads();
    for (unsigned i = tid; i < Ks; i += blockDim.x) { CHK(seq[i] / kDirs < V, 5); atomicMin(&firstpos[seq[i] / kDirs], i); }
    __syncthreads();
    const unsigned to_remove = (unsigned)((unsigned long)V - (unsigned long)D);   // V >= D on acceptance
    if (tid == 0) s_carry = 0;
    __syncthreads();
    unsigned walk_local = 0;       // max over threads of (pos+1) for removed entries
    for (unsigned base = 0; base < Ks; base += blockDim.x) {
        const unsigned i = base + tid;
        const bool first = i < Ks && firstpos[seq[i] / kDirs] == i;
        unsigned total;
        const unsigned r = s_carry + block_scan_step(first ? 1u : 0u, s_warp, total);
        if (first && r < to_remove && ((flags & 4u) || (unsigned long)i + r < (unsigned long)K)) {
            removed[seq[i] / kDirs] = 1;
            walk_local = i + 1;
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
        if (s_carry >= to_remove) break;
    }
    // number removed and walk end
    {
        unsigned cnt = 0;
        for (unsigned v = tid; v < V; v += blockDim.x) cnt += removed[v];
        unsigned total;
        __syncthreads();
        block_scan_step(cnt, s_warp, total);
        // reduce walk_local (max)
        unsigned w = walk_local;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { unsigned u = __shfl_xor_sync(0xffffffffu, w, o); w = u > w ? u : w; }
        __syncthreads();
        if (lane == 0) s_warp[wid] = w;
        __syncthreads();
        if (tid == 0) {
            unsigned wm = 0;
            for (int k = 0; k < (int)(blockDim.x >> 5); k++) wm = s_warp[k] > wm ? s_warp[k] : wm;
            const unsigned n_removed = total;
            s.K = K;
            s.n_valid = V - n_removed;
            if (n_removed < to_remove) { s.prune_ret = -2; s.walk = K - n_removed; }   // walk stopped at idx == K - removed
            else { s.prune_ret = 0; s.walk = wm; }
        }
        __syncthreads();
    }
    // ---- 4. ascending compaction of the survivors (ndt.c:75-117), clamped to D rows (A15)
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const double *mu = mean + (size_t)b * vcap * 3;
    const double *cf = cov_final + (size_t)b * vcap * 9;
    for (unsigned base = 0; base < V; base += blockDim.x) {
        const unsigned v = base + tid;
        const bool alive = v < V && !removed[v];
        unsigned total;
        const unsigned row = s_carry + block_scan_step(alive ? 1u : 0u, s_warp, total);
        if (alive && (long)row < D) {
            double f[12];
            f[0] = mu[v * 3 + 0]; f[1] = mu[v * 3 + 1]; f[2] = mu[v * 3 + 2];
#pragma unroll
            for (int k = 0; k < 9; k++) f[3 + k] = cf[(size_t)v * 9 + k];
            if (of64) {
#pragma unroll
                for (int k = 0; k < 12; k++) of64[(size_t)row * 12 + k] = f[k];
            }
            if (of) {
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    float x = (float)f[k];
                    if ((flags & 1u) && !isfinite(x)) x = 0.f;
                    of[(size_t)row * 12 + k] = x;
                }
            }
            if (ol) ol[row] = has_labels ? cls[(size_t)b * vcap + v] : (uint16_t)0;
            if (ov) ov[row] = (int)vox_cell[(size_t)b * vcap + v];
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    const unsigned survivors = s_carry;
    const unsigned rows = (long)survivors < D ? survivors : (unsigned)D;
    for (long i = (long)rows * 12 + tid; i < D * 12; i += blockDim.x) { if (of) of[i] = 0.f; if (of64) of64[i] = 0.0; }
    for (long i = (long)rows + tid; i < D; i += blockDim.x) { if (ol) ol[i] = 0; if (ov) ov[i] = -1; }
    if (tid == 0) {
        s.n_out = rows; s.n_survivors = survivors;
        if (info) {
            NdtCloudInfo &o = info[b];
            o.status = 0; o.prune_status = s.prune_ret; o.evaluations = s.evals;
            for (int a = 0; a < 3; a++) { o.len[a] = (unsigned)s.len[a]; o.offset[a] = s.off[a]; }
            o.num_voxels = V; o.num_valid = s.n_valid; o.num_kl = K;
            o.num_kl_after = K - (V - s.n_valid);   // what prune_nds leaves in *num_kl_divergences (ndt.c:65)
            o.num_out = rows; o.num_survivors = survivors;
            o.voxel_size = s.guess;
            for (int a = 0; a < 6; a++) o.limits[a] = s.lim[a];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// self test of div_by_count against the IEEE division, on pseudo-random operands (tests/ call this)
// ------------------------------------------------------------------------------------------------
__global__ void k_selftest_div(long n, unsigned seed, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32; h *= 0x94D049BB133111EBull; h ^= h >> 29;
        const unsigned mode = (unsigned)(i & 7);
        // mantissa from the hash; exponent spread depends on the mode (typical coordinates .. extreme)
        const int espan = mode < 4 ? 20 : (mode < 6 ? 200 : 1000);
        const int e = 1023 + (int)((h >> 52) % (unsigned)(2 * espan + 1)) - espan;
        unsigned long long bits = (h & 0x800FFFFFFFFFFFFFull) | ((unsigned long long)(e < 1 ? 1 : (e > 2046 ? 2046 : e)) << 52);
        if (mode == 3) bits &= 0xFFFFFFFFE0000000ull;            // fp32-representable numerators
        const double d = __longlong_as_double((long long)bits);
        unsigned long long h2 = h * 0xD6E8FEB86659FD93ull; h2 ^= h2 >> 32;
        const unsigned c = (mode & 1) ? (unsigned)(h2 % 120000u) + 1u : (unsigned)(h2 % 67000000u) + 1u;
        const double cnt = (double)c;
        const double want = d / cnt, got = div_by_count(d, cnt);
        if (__double_as_longlong(want) != __double_as_longlong(got)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t fill_recip_table(double2 *tab, long n) {
    if (n <= 0) return cudaSuccess;
    k_fill_recip<<<(unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256>>>(tab, n);
    return cudaDeviceSynchronize();
}

cudaError_t selftest_div(long n, unsigned seed, unsigned long long *mismatches_host) {
    unsigned long long *d = nullptr;
    cudaError_t e = cudaMalloc((void **)&d, 8);
    if (e != cudaSuccess) return e;
    cudaMemset(d, 0, 8);
    k_selftest_div<<<592, 256>>>(n, seed, d);
    e = cudaMemcpy(mismatches_host, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void k_init_limits(unsigned long long *lim_enc, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * kLimWords) { const int w = i % kLimWords; lim_enc[i] = w < 3 ? 0ull : (w < 6 ? ~0ull : 0ull); }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
// NDNET_B200_DEBUG_SYNC=1: synchronise after every launch and name the kernel that faulted
static bool debug_sync(const char *name) {
    const char *e = getenv("NDNET_B200_DEBUG_SYNC");
    return e && (*e == '1' || strcmp(e, name) == 0);
}
#define DBG(name) do { if (debug_sync(name)) { cudaError_t e_ = cudaStreamSynchronize(st); if (e_ == cudaSuccess) e_ = cudaGetLastError(); \
    if (e_ != cudaSuccess) { fprintf(stderr, "ndnet_b200: kernel %s failed: %s\n", name, cudaGetErrorString(e_)); return e_; } } } while (0)

template <typename T>
static cudaError_t run_typed(Workspace &w, const T *pts, const uint16_t *labels, int B, long N, int num_classes,
                             long D, unsigned flags, float *out_feat, double *out_feat64, uint16_t *out_labels,
                             int *out_voxel, NdtCloudInfo *info, cudaStream_t st) {
    const unsigned vcap = w.vcap;
    const int ntiles = (int)((N + kRankTile - 1) / kRankTile);
    const int nbins = num_classes + 1;
    const bool textbook = (flags & 4u) != 0;          // NDNET_B200_TEXTBOOK_KL
    StageTimer &tm = w.timer;
    if (tm.enabled && !tm.created) { for (auto &e : tm.ev) cudaEventCreate(&e); tm.created = true; }
    tm.mark(ST_LIMITS, st);
    k_init_limits<<<(B * kLimWords + 127) / 128, 128, 0, st>>>(w.lim_enc, B);
    {
        int chunks = (int)((N + 256 * 16 - 1) / (256 * 16));
        if (chunks < 1) chunks = 1;
        k_limits<T><<<dim3(chunks, B), 256, 0, st>>>(pts, N, w.lim_enc); DBG("k_limits");
    }
    k_decide<<<B, 256, 0, st>>>(w.states, w.lim_enc, w.bitmap, w.bitmap_stride, w.vox_cell, vcap, D, 0); DBG("k_decide");
    tm.mark(ST_SEARCH, st);
    {
        // a few CTAs per cloud, each walking a long contiguous span: a launch for clouds that are done is one wave of exiting
        // CTAs, and a live CTA merges its shared-memory bitmap into the global one once per ~15 k points
        int chunks = (int)((N + kCountPointsPerCta - 1) / kCountPointsPerCta);
        if (chunks > kCountCtasPerCloud) chunks = kCountCtasPerCloud;
        if (chunks < 1) chunks = 1;
        // scans rarely need more than five passes (3.2 on average): kSearchPairLaunches (count, decide) launch pairs, then one
        // launch in which the clouds that are still searching run their remaining rounds (k_search_tail)
        for (int it = 0; it < kSearchPairLaunches; it++) {
            k_count<T><<<dim3(chunks, B), 256, 0, st>>>(pts, N, w.states, w.bitmap, w.bitmap_stride); DBG("k_count");
            k_decide<<<B, 256, 0, st>>>(w.states, w.lim_enc, w.bitmap, w.bitmap_stride, w.vox_cell, vcap, D, 1); DBG("k_decide");
        }
        k_search_tail<T><<<B, 256, 0, st>>>(pts, N, w.states, w.lim_enc, w.bitmap, w.bitmap_stride, w.vox_cell, vcap, D); DBG("k_search_tail");
    }
    tm.mark(ST_RANK, st);
    if (N > 0) {
        const size_t rank_smem = 4 * (size_t)vcap * sizeof(unsigned short);
        if (rank_smem > 48 * 1024) CK(cudaFuncSetAttribute(k_rank<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rank_smem));
        k_rank<T><<<dim3((ntiles + 3) / 4, B), 128, rank_smem, st>>>(
            pts, N, w.states, w.bitmap, w.bitmap_stride, vcap, ntiles, w.slot_rank, w.tile_cnt, w.keep_point_voxels ? w.point_voxel : nullptr);
        DBG("k_rank");
    }
    tm.mark(ST_OFFSETS, st);
    k_tile_prefix<<<dim3((vcap + 255) / 256, B), 256, 0, st>>>(w.states, vcap, ntiles, w.tile_cnt, w.vox_n); DBG("k_tile_prefix");
    // voxels with at least heavy_min points go to k_stats (four lanes each), the rest to k_stats_light (a thread each)
    static const unsigned heavy_min = [] { const char *e = getenv("NDNET_B200_HEAVY_LOG2"); const int v = e ? atoi(e) : 0; return v >= 3 && (1u << v) <= kHeavyVoxel ? 1u << v : kHeavyVoxel; }();
    k_offsets<<<B, 1024, 0, st>>>(w.states, vcap, heavy_min, w.vox_n, w.vox_start, w.vox_order); DBG("k_offsets");
    tm.mark(ST_SCATTER, st);
    const bool wide_labels = labels && nbins > kSmemLabelBins;
    if (wide_labels) CK(cudaMemsetAsync(w.hist, 0, (size_t)B * vcap * nbins * sizeof(unsigned), st));
    if (N > 0) {
        k_scatter<T><<<dim3((unsigned)((N + 255) / 256), B), 256, 0, st>>>(
            pts, labels, (flags & 2u) ? 1 : 2, N, w.states, vcap, ntiles, w.slot_rank, w.tile_cnt, w.vox_start, (T *)w.sorted,
            wide_labels ? w.hist : nullptr, nbins);
        DBG("k_scatter");
    }
    if (w.mark_front) {
        if (!w.ev_front) CK(cudaEventCreateWithFlags(&w.ev_front, cudaEventDisableTiming));
        CK(cudaEventRecord(w.ev_front, st));
    }
    tm.mark(ST_STATS, st);
    // heavy voxels (a warp each; at most N / kHeavyVoxel of them per cloud), then the light ones (a thread each)
    {
        unsigned max_heavy = (unsigned)(N / heavy_min) + 1;
        if (max_heavy > vcap) max_heavy = vcap;
        // the light kernel and the votes are independent of the heavy one: fork them onto the side stream so the
        // latency-bound tails overlap
        if (!w.side) { CK(cudaStreamCreateWithFlags(&w.side, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&w.ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&w.ev_join, cudaEventDisableTiming)); }
        CK(cudaEventRecord(w.ev_fork, st));
        CK(cudaStreamWaitEvent(w.side, w.ev_fork, 0));
        // label vote: taken inside the statistics kernels from the label lane of the records they read anyway when the
        // class count fits their shared-memory counters; wide label sets were counted by k_scatter's global atomics
        const int vote_bins = labels && !wide_labels ? nbins : 0;
        const dim3 heavy_grid(B, (max_heavy + kStatsVoxelsPerWarp - 1) / kStatsVoxelsPerWarp), light_grid((vcap + 127) / 128, B);
        const size_t light_smem = (size_t)vote_bins * 128 * sizeof(unsigned short);
        if (textbook) {
            k_stats<T, true><<<heavy_grid, 32, 0, st>>>(w.states, vcap, N, (const T *)w.sorted, w.vox_start, w.vox_order, w.recip, w.mean, w.cov, w.cls, vote_bins);
            DBG("k_stats");
            k_stats_light<T, true><<<light_grid, 128, light_smem, w.side>>>(w.states, vcap, N, (const T *)w.sorted, w.vox_start, w.vox_order, w.mean, w.cov, w.cls, vote_bins);
        } else {
            k_stats<T, false><<<heavy_grid, 32, 0, st>>>(w.states, vcap, N, (const T *)w.sorted, w.vox_start, w.vox_order, w.recip, w.mean, w.cov, w.cls, vote_bins);
            DBG("k_stats");
            k_stats_light<T, false><<<light_grid, 128, light_smem, w.side>>>(w.states, vcap, N, (const T *)w.sorted, w.vox_start, w.vox_order, w.mean, w.cov, w.cls, vote_bins);
        }
        if (wide_labels) {
            k_votes<T><<<dim3(B, (vcap + 3) / 4), 128, 0, w.side>>>(w.states, vcap, N, (const T *)w.sorted, w.vox_start, w.hist, nbins, w.cls);
        }
        CK(cudaEventRecord(w.ev_join, w.side));
        CK(cudaStreamWaitEvent(st, w.ev_join, 0));
        DBG("k_stats_light/k_votes");
    }
    tm.mark(ST_KL, st);
    if (textbook)
        k_kl_textbook<<<dim3((vcap + 63) / 64, B), 64, 0, st>>>(w.states, vcap, w.bitmap, w.bitmap_stride, w.vox_cell, w.vox_n, w.mean, w.cov,
                                                               w.cov_final, w.kl_div, w.kl_flag);
    else
        k_kl<<<dim3((vcap + 63) / 64, B), 64, 0, st>>>(w.states, vcap, w.bitmap, w.bitmap_stride, w.vox_cell, w.vox_n, w.cov,
                                                      w.cov_final, w.kl_div, w.kl_flag);
    DBG("k_kl");
    tm.mark(ST_SELECT, st);
    {
        const size_t kcap = (size_t)vcap * kDirs;
        static bool attr_set[64] = {};   // function attributes are per device
        int dev = 0; cudaGetDevice(&dev);
        const size_t max_dyn = 200 * 1024;
        if (!attr_set[dev & 63]) { CK(cudaFuncSetAttribute(k_select<kSelectThreads, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn)); attr_set[dev & 63] = true; }
        if (w.keep_kl_list) {
            // the whole sorted list stays behind (legacy handles: the prune() continuation walks it; inspection)
            size_t P = 1; while (P < kcap) P <<= 1;
            int smem_cap = (int)P;
            size_t bytes = P * 12;
            while (bytes > max_dyn) { smem_cap >>= 1; bytes = (size_t)smem_cap * 12; }
            k_select<kSelectThreads, false><<<B, kSelectThreads, bytes, st>>>(
                w.states, vcap, D, w.vox_cell, w.vox_n, w.mean, w.cov_final, w.cls, labels ? 1 : 0, w.kl_div, w.kl_flag, w.key, w.seq, kcap,
                smem_cap, w.firstpos, w.removed, flags, out_feat, out_feat64, out_labels, out_voxel, w.list_div, w.list_seq, info);
        } else {
            // only the head of the list is sorted, in shared memory sized for what the walk can reach: 6 (V - D) <= 6 (Vcap - D)
            // entries - 24 KB at D = 1000, 96 KB at D = 4096 (BASELINE config 5); a selection that does not fit the largest
            // launch (D beyond ~13 k) sorts in the global scratch
            size_t reach = 6 * ((size_t)vcap - (size_t)(D < (long)vcap ? D : (long)vcap));
            if (reach > kcap) reach = kcap;
            int cap = kSelectPartialCap;
            while ((size_t)cap < reach && cap < kSelectPartialCapMax) cap <<= 1;
            static bool part_attr_set[64] = {};
            if (cap > 4096 && !part_attr_set[dev & 63]) {
                CK(cudaFuncSetAttribute(k_select<kSelectPartialThreads, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelectPartialCapMax * 12));
                part_attr_set[dev & 63] = true;
            }
            k_select<kSelectPartialThreads, true><<<B, kSelectPartialThreads, (size_t)cap * 12, st>>>(
                w.states, vcap, D, w.vox_cell, w.vox_n, w.mean, w.cov_final, w.cls, labels ? 1 : 0, w.kl_div, w.kl_flag, w.key, w.seq, kcap,
                cap, w.firstpos, w.removed, flags, out_feat, out_feat64, out_labels, out_voxel, w.list_div, w.list_seq, info);
        }
        DBG("k_select");
    }
    DBG("end");
    tm.mark(ST_COUNT, st);
    count_launches(3 + 2 * kSearchPairLaunches + 1 + (N > 0 ? 2 : 0) + 6 + (wide_labels ? 1 : 0));
    if (tm.enabled) {
        CK(cudaEventSynchronize(tm.ev[ST_COUNT]));
        for (int i = 0; i < ST_COUNT; i++) { float ms = 0; cudaEventElapsedTime(&ms, tm.ev[i], tm.ev[i + 1]); tm.ms[i] += ms; }
        tm.runs++;
    }
    return cudaGetLastError();
}

cudaError_t run_batch(Workspace &w, const void *pts, int dtype, const uint16_t *labels, int B, long N, int num_classes,
                      long D, unsigned flags, float *out_feat, double *out_feat64, uint16_t *out_labels, int *out_voxel,
                      NdtCloudInfo *info, cudaStream_t st) {
    if (dtype == 0)
        return run_typed<float>(w, (const float *)pts, labels, B, N, num_classes, D, flags, out_feat, out_feat64,
                                out_labels, out_voxel, info, st);
    return run_typed<double>(w, (const double *)pts, labels, B, N, num_classes, D, flags, out_feat, out_feat64, out_labels,
                             out_voxel, info, st);
}

}  // namespace ndt
