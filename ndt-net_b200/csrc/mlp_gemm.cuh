// mlp_gemm.cuh — hand-written sm_100a GEMM for the PointNet / NDT-Net shared MLP:
//   D[m, n] = sum_k A[m, k] * B[n, k]      (both operands K-major bf16, fp32 accumulation)
// TMA (cp.async.bulk.tensor, 128B swizzle) stages the operand tiles in shared memory, one elected
// thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN) into a TMEM accumulator, and four
// epilogue warps read it back with tcgen05.ld and apply the fused epilogue:
//   MODE_ROWS    rows = points of one cloud, columns = output channels:
//                +bias (+per-cloud bias) (+ReLU) -> bf16 row-major activations for the next layer
//   MODE_LOGSM   same orientation, final layer: +bias -> log_softmax over the valid columns -> fp32
//   MODE_CHMAX   rows = output channels, columns = points of one cloud: the global max-pool over the
//                points is a per-thread running max over TMEM columns, then +bias (+ReLU) and one
//                atomicMax per channel: the [B, F, N] tensor of ndtnet.py:161 is never written.
// Tiles never straddle clouds (3-D tensor maps {K, rows, cloud}; TMA zero-fills rows past the cloud).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "tc_common.cuh"

namespace mlp {

enum { MODE_ROWS = 0, MODE_LOGSM = 1, MODE_CHMAX = 2 };

struct GemmArgs {
    int K;              // multiple of 64
    int P;              // points per cloud
    int a_batched;      // third TMA coordinate of the 128-row operand is the cloud index
    int b_batched;      // ... of the BN-row operand
    int mode;
    int relu;
    int n_valid;        // valid output channels
    const float *bias;  // [n_valid] or null
    const float *cbias; // per-cloud bias [B][ldcb] or null
    int ldcb;
    __nv_bfloat16 *out; // MODE_ROWS: [B][P][ldo]
    int ldo;
    float *outf;        // MODE_LOGSM: [B][P][n_valid]
    unsigned *gmax;     // MODE_CHMAX: [B][ldg], order-preserving encoded floats (see enc_f32)
    int ldg;
};

__device__ __forceinline__ unsigned enc_f32(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float dec_f32(unsigned u) {
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}


// operand stages; MODE_ROWS re-uses them (all MMAs have retired) to stage the 128 x BN bf16 output tile for the TMA store
template <int BN, int STAGES>
__host__ __device__ constexpr size_t gemm_stage_bytes() { return (size_t)STAGES * (128 + BN) * 128 > (size_t)BN * 256 ? (size_t)STAGES * (128 + BN) * 128 : (size_t)BN * 256; }
template <int BN, int STAGES>
__host__ __device__ constexpr size_t gemm_smem_bytes() { return gemm_stage_bytes<BN, STAGES>() + 1024 /*align*/ + 256 /*barriers*/ + BN * 4 /*bias tile*/; }

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads)
k_gemm(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapC,
       const GemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                               // STAGES x 16 KB
    uint8_t *sB = smem + (size_t)STAGES * 128 * 128;  // STAGES x BN*128 B
    uint64_t *bars = (uint64_t *)(smem + gemm_stage_bytes<BN, STAGES>());
    uint64_t *full = bars, *empty = bars + STAGES, *tmem_full = bars + 2 * STAGES;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES + 1);
    float *s_bias = (float *)(bars + 2 * STAGES + 2);   // BN floats: bias (+ per-cloud bias) of this tile's columns

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z;
    const int nkb = args.K / 64;
    // operand tile origins
    int a_row0, b_row0;
    if (args.mode == MODE_CHMAX) { a_row0 = blockIdx.x * 128; b_row0 = blockIdx.y * BN; }
    else { a_row0 = blockIdx.y * 128; b_row0 = blockIdx.x * BN; }
    const int a_z = args.a_batched ? b : 0, b_z = args.b_batched ? b : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(tmem_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { ptx::prefetch_tmap(&mapA); ptx::prefetch_tmap(&mapB); if (args.mode == MODE_ROWS) ptx::prefetch_tmap(&mapC); }
    if (args.mode == MODE_ROWS && threadIdx.x < 128) {
        const float *cb = args.cbias ? args.cbias + (size_t)b * args.ldcb : nullptr;
        for (int i = threadIdx.x; i < BN; i += 128) {
            const int n = b_row0 + i;
            float v = 0.f;
            if (n < args.n_valid) { if (args.bias) v += args.bias[n]; if (cb) v += cb[n]; }
            s_bias[i] = v;
        }
    }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                ptx::mbar_expect_tx(&full[s], (128 + BN) * 128);
                ptx::tma_load_3d(&mapA, &full[s], sA + (size_t)s * 128 * 128, kb * 64, a_row0, a_z);
                ptx::tma_load_3d(&mapB, &full[s], sB + (size_t)s * BN * 128, kb * 64, b_row0, b_z);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BN);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                ptx::mbar_wait(&full[s], ph);
                ptx::tc_fence_after();
                const uint64_t da = make_kmajor_sw128_desc(ptx::smem_u32(sA + (size_t)s * 128 * 128));
                const uint64_t db = make_kmajor_sw128_desc(ptx::smem_u32(sB + (size_t)s * BN * 128));
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++)   // UMMA_K = 16 bf16 = 32 B = 2 descriptor units
                    ptx::umma_bf16(tmem_base, da + (uint64_t)(k4 * 2), db + (uint64_t)(k4 * 2), idesc, (kb | k4) ? 1u : 0u);
                ptx::umma_commit(&empty[s]);     // frees the smem stage when these MMAs retire
            }
            ptx::umma_commit(tmem_full);         // accumulator complete
        }
    } else {
        // ---- epilogue: warp w owns TMEM lanes [32w, 32w+32)
        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int m = warp * 32 + lane;          // accumulator row of this thread
        if (args.mode == MODE_CHMAX) {
            const int ch = a_row0 + m;
            int valid = args.P - b_row0; if (valid > BN) valid = BN;
            float best = -INFINITY;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (c0 >= valid) break;
                uint32_t v[32];
                ptx::tmem_ld32(taddr + (uint32_t)c0, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) if (c0 + j < valid) best = fmaxf(best, __uint_as_float(v[j]));
            }
            if (ch < args.n_valid) {
                float r = best + (args.bias ? args.bias[ch] : 0.f);
                if (args.relu) r = fmaxf(r, 0.f);
                atomicMax(&args.gmax[(size_t)b * args.ldg + ch], enc_f32(r));
            }
        } else if (args.mode == MODE_ROWS) {
            // The bf16 tile goes to shared memory in the 128B-swizzled box layout of the output tensor map (one 128 x 64
            // sub-tile per 64 columns; 16-byte chunk c of row r sits at chunk c ^ (r & 7), so the eight rows of a quarter
            // warp hit eight different bank groups), then ONE thread hands it to the TMA: full 128-byte lines leave the SM
            // instead of 128 row-strided 16-byte stores per column chunk.  Rows past the cloud are clipped by the TMA.
            uint8_t *stage = smem;                     // operand stages are dead: every MMA of this tile has retired
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                ptx::tmem_ld32(taddr + (uint32_t)c0, v);
                ptx::tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const int n = b_row0 + c0 + j;
                    float x0 = __uint_as_float(v[j]) + s_bias[c0 + j], x1 = __uint_as_float(v[j + 1]) + s_bias[c0 + j + 1];
                    if (n >= args.n_valid) x0 = 0.f;
                    if (n + 1 >= args.n_valid) x1 = 0.f;
                    if (args.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                    packed[j >> 1] = *reinterpret_cast<uint32_t *>(&h);
                }
                uint8_t *sub = stage + (size_t)(c0 >> 6) * 16384 + (size_t)m * 128;
                const int chunk0 = (c0 & 63) >> 3;       // first of this step's four 16-byte chunks within the 128-byte row
#pragma unroll
                for (int q = 0; q < 4; q++)
                    *reinterpret_cast<uint4 *>(sub + (((chunk0 + q) ^ (m & 7)) << 4)) =
                        make_uint4(packed[q * 4], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
            }
            ptx::fence_proxy_async_smem();
            ptx::named_barrier_sync(1, 128);             // the four epilogue warps only
            if (threadIdx.x == 0) {
#pragma unroll
                for (int st = 0; st < BN / 64; st++) ptx::tma_store_3d(&mapC, stage + (size_t)st * 16384, b_row0 + st * 64, a_row0, b);
                ptx::tma_store_commit();
                ptx::tma_store_wait_read();              // the CTA's shared memory must outlive the TMA's reads
            }
        } else {   // MODE_LOGSM: BN == 32, the whole row is in this thread
            const int row = a_row0 + m;
            uint32_t v[32];
            ptx::tmem_ld32(taddr, v);
            ptx::tmem_ld_wait();
            // The tile's rows are one contiguous block of the [B, P, n_valid] fp32 output (row stride = n_valid floats), so
            // it goes through shared memory (row m at word m * n_valid: an odd stride for 29 classes, conflict-free) and
            // leaves as consecutive words: a row-per-thread store would fill 4 bytes of every 32-byte sector it touches.
            float *s_out = reinterpret_cast<float *>(smem);          // operand stages are dead (every MMA has retired)
            const int nv = args.n_valid;
            {
                float x[32];
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    x[j] = (j < nv) ? __uint_as_float(v[j]) + args.bias[j] : -INFINITY;
                    mx = fmaxf(mx, x[j]);
                }
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j++) if (j < nv) sum += __expf(x[j] - mx);
                const float lse = mx + __logf(sum);
#pragma unroll
                for (int j = 0; j < 32; j++) if (j < nv) s_out[m * nv + j] = x[j] - lse;
            }
            ptx::named_barrier_sync(1, 128);
            int rows_here = args.P - a_row0; if (rows_here > 128) rows_here = 128;
            float *o = args.outf + ((size_t)b * args.P + a_row0) * nv;
            const int total = rows_here * nv;
            for (int i = threadIdx.x; i < total; i += 128) o[i] = s_out[i];
            (void)row;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

// ------------------------------------------------------------------------------------------------
// k_gemm_chmax: the max-pooled layers (conv 128 -> 1024 / F followed by max over the points of a cloud,
// ndtnet.py:47-50,161,186,224).  D[channel, point] = W[channel, :] . act[point, :], K = 128.
//
// One CTA owns one cloud and a group of up to 4 channel tiles (512 channels).  The group's weights (4 x 128 x 128
// bf16 = 128 KB) are loaded ONCE and stay in shared memory; the cloud's activations stream through a 3-stage TMA
// ring in tiles of 64 points (16 KB), so every activation byte is read C/512 times instead of C/128 times and the
// loop is paced by the tensor pipe, not by L2.  The accumulators (4 x 64 columns) are double buffered in TMEM:
// while tcgen05.mma fills one buffer the four epilogue warps drain the other with tcgen05.ld into a per-thread
// running max (thread = channel).  Each (cloud, channel) maximum is produced by exactly one thread: plain store,
// no atomics.  grid (ceil(C/512), B), block 192.
// ------------------------------------------------------------------------------------------------
constexpr int kChmaxStages = 3;
constexpr size_t kChmaxSmemBytes = 4 * 32768 + kChmaxStages * 16384 + 1024 + 256;

__global__ void __launch_bounds__(kGemmThreads)
k_gemm_chmax(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX, const GemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sW = smem;                                   // [tile 4][kb 2][128 rows x 128 B]
    uint8_t *sX = smem + 4 * 32768;                       // [stage][kb 2][64 rows x 128 B]
    uint64_t *bars = (uint64_t *)(sX + kChmaxStages * 16384);
    uint64_t *w_full = bars, *x_full = bars + 1, *x_empty = x_full + kChmaxStages;
    uint64_t *t_full = x_empty + kChmaxStages, *t_empty = t_full + 2;
    uint32_t *tmem_slot = (uint32_t *)(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int ch0 = blockIdx.x * 512;
    int ntile = (args.n_valid - ch0 + 127) / 128; if (ntile > 4) ntile = 4;
    const int npt = (args.P + 63) / 64;

    if (threadIdx.x == 0) {
        ptx::mbar_init(w_full, 1);
        for (int s = 0; s < kChmaxStages; s++) { ptx::mbar_init(&x_full[s], 1); ptx::mbar_init(&x_empty[s], 1); }
        for (int i = 0; i < 2; i++) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { ptx::prefetch_tmap(&mapW); ptx::prefetch_tmap(&mapX); }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            ptx::mbar_expect_tx(w_full, (uint32_t)ntile * 32768u);
            for (int ct = 0; ct < ntile; ct++)
                for (int kb = 0; kb < 2; kb++)
                    ptx::tma_load_3d(&mapW, w_full, sW + (size_t)(ct * 2 + kb) * 16384, kb * 64, ch0 + ct * 128, 0);
            for (int t = 0; t < npt; t++) {
                const int s = t % kChmaxStages;
                const uint32_t ph = (uint32_t)(t / kChmaxStages) & 1u;
                ptx::mbar_wait(&x_empty[s], ph ^ 1u);
                ptx::mbar_expect_tx(&x_full[s], 16384u);
                ptx::tma_load_3d(&mapX, &x_full[s], sX + (size_t)s * 16384, 0, t * 64, b);
                ptx::tma_load_3d(&mapX, &x_full[s], sX + (size_t)s * 16384 + 8192, 64, t * 64, b);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 64);
            ptx::mbar_wait(w_full, 0);
            for (int t = 0; t < npt; t++) {
                const int s = t % kChmaxStages, buf = t & 1;
                ptx::mbar_wait(&x_full[s], (uint32_t)(t / kChmaxStages) & 1u);
                ptx::mbar_wait(&t_empty[buf], ((uint32_t)(t >> 1) & 1u) ^ 1u);     // the epilogue drained this buffer
                ptx::tc_fence_after();
                for (int ct = 0; ct < ntile; ct++) {
                    const uint32_t d = tmem_base + (uint32_t)(buf * 256 + ct * 64);
#pragma unroll
                    for (int kb = 0; kb < 2; kb++) {
                        const uint64_t da = make_kmajor_sw128_desc(ptx::smem_u32(sW + (size_t)(ct * 2 + kb) * 16384));
                        const uint64_t db = make_kmajor_sw128_desc(ptx::smem_u32(sX + (size_t)s * 16384 + (size_t)kb * 8192));
#pragma unroll
                        for (int k4 = 0; k4 < 4; k4++)
                            ptx::umma_bf16(d, da + (uint64_t)(k4 * 2), db + (uint64_t)(k4 * 2), idesc, (kb | k4) ? 1u : 0u);
                    }
                }
                ptx::umma_commit(&x_empty[s]);
                ptx::umma_commit(&t_full[buf]);
            }
        }
    } else {
        // ---- epilogue: thread = channel (TMEM lane); running max over the cloud's points per channel tile
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int t = 0; t < npt; t++) {
            const int buf = t & 1;
            int valid = args.P - t * 64; if (valid > 64) valid = 64;
            ptx::mbar_wait(&t_full[buf], (uint32_t)(t >> 1) & 1u);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 256);
            if (valid == 64) {
                // full tile of points: the 64 accumulator columns of channel tile ct+1 are in flight (tcgen05.ld is
                // asynchronous) while the 64 of tile ct go through the running max, so the TMEM read latency is paid once
                // per point tile instead of eight times
                uint32_t a0[32], a1[32], b0[32], b1[32];
                ptx::tmem_ld32(taddr, a0);
                ptx::tmem_ld32(taddr + 32u, a1);
                ptx::tmem_ld_wait();
                ptx::tmem_ld_fence(a0); ptx::tmem_ld_fence(a1);
#pragma unroll
                for (int ct = 0; ct < 4; ct++) {
                    if (ct < ntile) {
                        const bool more = ct + 1 < ntile;
                        if ((ct & 1) == 0) {
                            if (more) { ptx::tmem_ld32(taddr + (uint32_t)((ct + 1) * 64), b0); ptx::tmem_ld32(taddr + (uint32_t)((ct + 1) * 64 + 32), b1); }
#pragma unroll
                            for (int j = 0; j < 32; j++) best[ct] = fmaxf(best[ct], fmaxf(__uint_as_float(a0[j]), __uint_as_float(a1[j])));
                            if (more) { ptx::tmem_ld_wait(); ptx::tmem_ld_fence(b0); ptx::tmem_ld_fence(b1); }
                        } else {
                            if (more) { ptx::tmem_ld32(taddr + (uint32_t)((ct + 1) * 64), a0); ptx::tmem_ld32(taddr + (uint32_t)((ct + 1) * 64 + 32), a1); }
#pragma unroll
                            for (int j = 0; j < 32; j++) best[ct] = fmaxf(best[ct], fmaxf(__uint_as_float(b0[j]), __uint_as_float(b1[j])));
                            if (more) { ptx::tmem_ld_wait(); ptx::tmem_ld_fence(a0); ptx::tmem_ld_fence(a1); }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int ct = 0; ct < 4; ct++) {
                    if (ct < ntile) {
#pragma unroll
                        for (int c0 = 0; c0 < 64; c0 += 32) {
                            if (c0 < valid) {
                                uint32_t v[32];
                                ptx::tmem_ld32(taddr + (uint32_t)(ct * 64 + c0), v);
                                ptx::tmem_ld_wait();
#pragma unroll
                                for (int j = 0; j < 32; j++) if (c0 + j < valid) best[ct] = fmaxf(best[ct], __uint_as_float(v[j]));
                            }
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(&t_empty[buf])) : "memory");
            }
        }
#pragma unroll
        for (int ct = 0; ct < 4; ct++) {
            const int ch = ch0 + ct * 128 + warp * 32 + lane;
            if (ct < ntile && ch < args.n_valid) {
                float r = best[ct] + (args.bias ? args.bias[ch] : 0.f);
                if (args.relu) r = fmaxf(r, 0.f);
                args.gmax[(size_t)b * args.ldg + ch] = enc_f32(r);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// k_head12: the first two layers of the segmentation head fused (ndtnet.py:231-232):
//   H1 = relu(X . W1a^T + cb[cloud])      X: [P, 64] x_t2 rows of one cloud, W1a: [512, 64], cb = W1g . g + b1 per cloud
//   out = relu(H1 . W2^T + b2)            W2: [256, 512]
// The 512-wide activation never leaves the SM: it is produced 64 columns at a time in TMEM (MMA1, N = 64), read back by the
// four epilogue warps (tcgen05.ld), biased / rectified / rounded to bf16 and written into shared memory in the K-major
// 128B-swizzled operand layout, from where MMA2 (N = 256) accumulates out += H1[:, chunk] . W2[:, chunk]^T in TMEM.
// The W2 chunks (32 KB each) stream through a TMA ring; W1a (64 KB) and the X tile are loaded once.  TMEM: columns
// [0, 256) = out accumulator, [256, 320) and [320, 384) = the two H1 chunk buffers.  MMA1 of chunk c is issued before
// MMA2 of chunk c - 1, so the tensor pipe works on chunk c while the epilogue warps convert chunk c - 1.
// The conversion of a chunk (TMEM -> registers -> bias / ReLU / bf16 -> swizzled shared memory) paces the chain, so eight
// epilogue warps share it: warp w owns TMEM lanes 32 (w % 4) .. +32 (the hardware's rule) and the column half w / 4.
// One CTA per (128-row tile, cloud); grid (ceil(P/128), B), block 320 (warps 0-7 epilogue, 8 TMA, 9 MMA).  The unfused pair
// moved 1 GB per 512 scans through HBM for this activation (write 524 MB + read 524 MB).
// ------------------------------------------------------------------------------------------------
constexpr int kHead12W2Stages = 2;
constexpr int kHead12Threads = 320;
constexpr size_t kHead12SmemBytes = 16384 /*X*/ + 65536 /*W1a, later the output staging*/ + kHead12W2Stages * 32768 /*W2 ring*/ +
                                    2 * 16384 /*H1 bf16 chunk buffers*/ + 1024 /*align*/ + 256 /*barriers*/ + 512 * 4 + 256 * 4;

struct Head12Args {
    int P;
    const float *cbias;     // [B][512]
    const float *bias2;     // [256]
};

__global__ void __launch_bounds__(kHead12Threads)
k_head12(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
         const __grid_constant__ CUtensorMap mapOut, const Head12Args args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sX = smem;                                        // [128 rows x 128 B]
    uint8_t *sW1 = sX + 16384;                                 // 8 chunks of [64 rows x 128 B]
    uint8_t *sW2 = sW1 + 65536;                                // ring of [256 rows x 128 B]
    uint8_t *sH = sW2 + (size_t)kHead12W2Stages * 32768;       // 2 x [128 rows x 128 B]
    uint64_t *bars = (uint64_t *)(sH + 2 * 16384);
    uint64_t *x_full = bars, *acc_full = bars + 1;
    uint64_t *w2_full = bars + 2, *w2_empty = w2_full + kHead12W2Stages;
    uint64_t *ht_full = w2_empty + kHead12W2Stages, *ht_empty = ht_full + 2;     // H1 chunk in TMEM
    uint64_t *hs_full = ht_empty + 2, *hs_empty = hs_full + 2;                   // H1 chunk in shared memory (bf16)
    uint32_t *tmem_slot = (uint32_t *)(hs_empty + 2);
    float *s_cb = (float *)(bars + 32);                        // 512 per-cloud biases of layer 1, then 256 biases of layer 2
    float *s_b2 = s_cb + 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, row0 = blockIdx.x * 128;

    if (threadIdx.x == 0) {
        ptx::mbar_init(x_full, 1); ptx::mbar_init(acc_full, 1);
        for (int s = 0; s < kHead12W2Stages; s++) { ptx::mbar_init(&w2_full[s], 1); ptx::mbar_init(&w2_empty[s], 1); }
        for (int i = 0; i < 2; i++) { ptx::mbar_init(&ht_full[i], 1); ptx::mbar_init(&ht_empty[i], 8); ptx::mbar_init(&hs_full[i], 8); ptx::mbar_init(&hs_empty[i], 1); }
        ptx::fence_barrier_init();
    }
    if (warp == 8 && lane == 0) { ptx::prefetch_tmap(&mapX); ptx::prefetch_tmap(&mapW1); ptx::prefetch_tmap(&mapW2); ptx::prefetch_tmap(&mapOut); }
    if (threadIdx.x < 256) {
        for (int i = threadIdx.x; i < 512; i += 256) s_cb[i] = args.cbias[(size_t)b * 512 + i];
        s_b2[threadIdx.x] = args.bias2[threadIdx.x];
    }
    if (warp == 9) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            ptx::mbar_expect_tx(x_full, 16384u + 65536u);
            ptx::tma_load_3d(&mapX, x_full, sX, 0, row0, b);
            ptx::tma_load_3d(&mapW1, x_full, sW1, 0, 0, 0);                     // rows 0..255 of W1a
            ptx::tma_load_3d(&mapW1, x_full, sW1 + 32768, 0, 256, 0);           // rows 256..511
            for (int c = 0; c < 8; c++) {
                const int s = c % kHead12W2Stages;
                ptx::mbar_wait(&w2_empty[s], ((uint32_t)(c / kHead12W2Stages) & 1u) ^ 1u);
                ptx::mbar_expect_tx(&w2_full[s], 32768u);
                ptx::tma_load_3d(&mapW2, &w2_full[s], sW2 + (size_t)s * 32768, c * 64, 0, 0);     // K columns [64c, 64c + 64) of all 256 rows
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            constexpr uint32_t idesc1 = make_idesc_bf16(128, 64), idesc2 = make_idesc_bf16(128, 256);
            ptx::mbar_wait(x_full, 0);
            ptx::tc_fence_after();
            const uint64_t dx = make_kmajor_sw128_desc(ptx::smem_u32(sX));
            for (int c = 0; c <= 8; c++) {
                if (c < 8) {
                    // MMA1: H1[:, 64c .. 64c+64) into TMEM buffer c & 1 (free once the epilogue has read chunk c - 2 out of it)
                    ptx::mbar_wait(&ht_empty[c & 1], ((uint32_t)(c >> 1) & 1u) ^ 1u);
                    ptx::tc_fence_after();
                    const uint64_t dw = make_kmajor_sw128_desc(ptx::smem_u32(sW1 + (size_t)c * 8192));
                    const uint32_t d1 = tmem_base + 256u + (uint32_t)((c & 1) * 64);
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++) ptx::umma_bf16(d1, dx + (uint64_t)(k4 * 2), dw + (uint64_t)(k4 * 2), idesc1, k4 ? 1u : 0u);
                    ptx::umma_commit(&ht_full[c & 1]);
                }
                if (c >= 1) {
                    // MMA2: out += H1 chunk (c - 1) (bf16 in shared memory) . W2[:, chunk]^T
                    const int cc = c - 1, s = cc % kHead12W2Stages;
                    ptx::mbar_wait(&hs_full[cc & 1], (uint32_t)(cc >> 1) & 1u);
                    ptx::mbar_wait(&w2_full[s], (uint32_t)(cc / kHead12W2Stages) & 1u);
                    ptx::tc_fence_after();
                    const uint64_t da = make_kmajor_sw128_desc(ptx::smem_u32(sH + (size_t)(cc & 1) * 16384));
                    const uint64_t db = make_kmajor_sw128_desc(ptx::smem_u32(sW2 + (size_t)s * 32768));
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++) ptx::umma_bf16(tmem_base, da + (uint64_t)(k4 * 2), db + (uint64_t)(k4 * 2), idesc2, (cc | k4) ? 1u : 0u);
                    ptx::umma_commit(&w2_empty[s]);
                    ptx::umma_commit(&hs_empty[cc & 1]);
                }
            }
            ptx::umma_commit(acc_full);
        }
    } else {
        // ---- epilogue warps: thread = row m of the tile (TMEM lane quadrant warp % 4), column half warp / 4
        const int m = (warp & 3) * 32 + lane, half = warp >> 2;
        const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int c = 0; c < 8; c++) {
            const int buf = c & 1;
            ptx::mbar_wait(&ht_full[buf], (uint32_t)(c >> 1) & 1u);
            ptx::tc_fence_after();
            uint32_t v[32];
            ptx::tmem_ld32(tlane + 256u + (uint32_t)(buf * 64 + half * 32), v);
            ptx::tmem_ld_wait();
            ptx::tmem_ld_fence(v);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(&ht_empty[buf])) : "memory");
            uint32_t packed[16];
            const float *cb = s_cb + c * 64 + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float a0 = fmaxf(__uint_as_float(v[j]) + cb[j], 0.f), a1 = fmaxf(__uint_as_float(v[j + 1]) + cb[j + 1], 0.f);
                __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
                packed[j >> 1] = *reinterpret_cast<uint32_t *>(&h);
            }
            // the MMA2 that read this shared-memory buffer (chunk c - 2) has retired
            ptx::mbar_wait(&hs_empty[buf], ((uint32_t)(c >> 1) & 1u) ^ 1u);
            uint8_t *rowp = sH + (size_t)buf * 16384 + (size_t)m * 128;
#pragma unroll
            for (int q = 0; q < 4; q++)
                *reinterpret_cast<uint4 *>(rowp + (((half * 4 + q) ^ (m & 7)) << 4)) = make_uint4(packed[q * 4], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(&hs_full[buf])) : "memory");
        }
        // ---- layer-2 epilogue: +b2, ReLU, bf16, swizzled staging (W1a's buffer: every MMA1 has retired), TMA store
        ptx::mbar_wait(acc_full, 0);
        ptx::tc_fence_after();
        uint8_t *stage = sW1;
#pragma unroll 1
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
            uint32_t v[32];
            ptx::tmem_ld32(tlane + (uint32_t)c0, v);
            ptx::tmem_ld_wait();
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float x0 = fmaxf(__uint_as_float(v[j]) + s_b2[c0 + j], 0.f), x1 = fmaxf(__uint_as_float(v[j + 1]) + s_b2[c0 + j + 1], 0.f);
                __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                packed[j >> 1] = *reinterpret_cast<uint32_t *>(&h);
            }
            uint8_t *sub = stage + (size_t)(c0 >> 6) * 16384 + (size_t)m * 128;
            const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
            for (int q = 0; q < 4; q++)
                *reinterpret_cast<uint4 *>(sub + (((chunk0 + q) ^ (m & 7)) << 4)) = make_uint4(packed[q * 4], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
        }
        ptx::fence_proxy_async_smem();
        ptx::named_barrier_sync(1, 256);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int st = 0; st < 4; st++) ptx::tma_store_3d(&mapOut, stage + (size_t)st * 16384, st * 64, row0, b);
            ptx::tma_store_commit();
            ptx::tma_store_wait_read();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace mlp
