// train_gemm.cuh — tcgen05 / TMA GEMM for the training step (train.cu), TF32 operands read straight from the fp32
// activations, gradients and parameters (no converted copies), fp32 accumulation in TMEM, fp32 output:
//   C[m, n] (+)= sum_k A[m, k] * B[n, k] (+ bias[n])        A: [M, K] row-major (lda), B: [N, K] row-major (ldb)
// Same pipeline as mlp_gemm.cuh's k_gemm (warp 4 = TMA producer, warp 5 = TMEM allocator + single-thread tcgen05.mma
// issue, warps 0-3 = epilogue through tcgen05.ld), with 128-byte K-blocks of 32 fp32 values, kind::tf32 (UMMA_K = 8),
// a 3-stage ring, and split-K over blockIdx.z with fp32 atomics for the weight-gradient GEMMs whose contraction runs
// over the rows.  TMA zero-fills rows/columns past the matrix, so no operand is padded.
#pragma once
#include "tc_common.cuh"

namespace train {

struct Tf32Args {
    int M, N, K;
    float *C; long ldc;
    const float *bias;
    int accumulate;       // C += result (ignored when splitk > 1: the atomics always add)
    int splitk, kb_per_split;
    // train-mode BatchNorm: per-column sum and sum of squares of the OUTPUT (bias included) over the rows, accumulated here
    // in the epilogue (fp32 over the 32 rows of a warp, then one fp64 atomic per column per warp) instead of a separate
    // column-reduction pass over the activation.  Only with splitk == 1 and accumulate == 0; null = off.
    double *stat_sum, *stat_sq;
};

// kind::tf32 instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, K-major both, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int kTf32Stages = 3;
template <int BN>
constexpr size_t tf32_smem_bytes() { return (size_t)kTf32Stages * (128 + BN) * 128 + 1024 /*align*/ + 256 /*barriers*/; }

template <int BN>
__global__ void __launch_bounds__(mlp::kGemmThreads)
k_gemm_tf32(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Tf32Args args) {
    namespace ptx = mlp::ptx;
    constexpr int STAGES = kTf32Stages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                               // STAGES x 128 rows x 128 B
    uint8_t *sB = smem + (size_t)STAGES * 128 * 128;  // STAGES x BN rows x 128 B
    uint64_t *bars = (uint64_t *)(sB + (size_t)STAGES * BN * 128);
    uint64_t *full = bars, *empty = bars + STAGES, *tmem_full = bars + 2 * STAGES;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_row0 = blockIdx.y * 128, b_row0 = blockIdx.x * BN;
    const int nkb_total = (args.K + 31) / 32;
    const int kb0 = blockIdx.z * args.kb_per_split;
    int nkb = nkb_total - kb0;
    if (nkb > args.kb_per_split) nkb = args.kb_per_split;      // the host guarantees nkb >= 1 for every z

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(tmem_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { ptx::prefetch_tmap(&mapA); ptx::prefetch_tmap(&mapB); }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, BN);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            for (int i = 0; i < nkb; i++) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                ptx::mbar_expect_tx(&full[s], (128 + BN) * 128);
                ptx::tma_load_3d(&mapA, &full[s], sA + (size_t)s * 128 * 128, (kb0 + i) * 32, a_row0, 0);
                ptx::tma_load_3d(&mapB, &full[s], sB + (size_t)s * BN * 128, (kb0 + i) * 32, b_row0, 0);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(128, BN);
            for (int i = 0; i < nkb; i++) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                ptx::mbar_wait(&full[s], ph);
                ptx::tc_fence_after();
                const uint64_t da = mlp::make_kmajor_sw128_desc(ptx::smem_u32(sA + (size_t)s * 128 * 128));
                const uint64_t db = mlp::make_kmajor_sw128_desc(ptx::smem_u32(sB + (size_t)s * BN * 128));
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++)   // UMMA_K = 8 tf32 = 32 B = 2 descriptor units
                    umma_tf32(tmem_base, da + (uint64_t)(k4 * 2), db + (uint64_t)(k4 * 2), idesc, (i | k4) ? 1u : 0u);
                ptx::umma_commit(&empty[s]);
            }
            ptx::umma_commit(tmem_full);
        }
    } else {
        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int row = a_row0 + warp * 32 + lane;
        const bool add_bias = args.bias != nullptr && blockIdx.z == 0;
        const bool vec = args.splitk == 1 && args.ldc % 4 == 0 && ((uintptr_t)args.C & 15) == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (b_row0 + c0 >= args.N) break;                 // warp-uniform
            uint32_t v[32];
            ptx::tmem_ld32(taddr + (uint32_t)c0, v);
            ptx::tmem_ld_wait();
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; j++) {
                y[j] = __uint_as_float(v[j]);
                if (add_bias && b_row0 + c0 + j < args.N) y[j] += args.bias[b_row0 + c0 + j];
            }
            if (row < args.M) {
                float *dst = args.C + (size_t)row * args.ldc + b_row0 + c0;
                if (vec && b_row0 + c0 + 32 <= args.N) {          // full 32-column chunk: eight 16-byte stores per row
                    float4 *d4 = reinterpret_cast<float4 *>(dst);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        float4 r = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
                        if (args.accumulate) { const float4 o = d4[q]; r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
                        d4[q] = r;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int col = b_row0 + c0 + j;
                        if (col < args.N) {
                            if (args.splitk > 1) atomicAdd(dst + j, y[j]);
                            else if (args.accumulate) dst[j] += y[j];
                            else dst[j] = y[j];
                        }
                    }
                }
            }
            if (args.stat_sum) {
                // column sums over the warp's 32 rows by recursive halving: after the step with distance d a lane holds
                // the partial sums of d columns; lane l ends with column l.  62 shuffles per chunk for both statistics.
                float q[32];
#pragma unroll
                for (int j = 0; j < 32; j++) { if (row >= args.M) y[j] = 0.f; q[j] = y[j] * y[j]; }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const bool upper = (lane & d) != 0;
#pragma unroll
                    for (int j = 0; j < d; j++) {
                        const float ys = upper ? y[j] : y[j + d], yk = upper ? y[j + d] : y[j];
                        const float qs = upper ? q[j] : q[j + d], qk = upper ? q[j + d] : q[j];
                        y[j] = yk + __shfl_xor_sync(0xffffffffu, ys, d);
                        q[j] = qk + __shfl_xor_sync(0xffffffffu, qs, d);
                    }
                }
                const int col = b_row0 + c0 + lane;
                if (col < args.N) {
                    atomicAdd(args.stat_sum + col, (double)y[0]);
                    atomicAdd(args.stat_sq + col, (double)q[0]);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, BN);
}

}  // namespace train
