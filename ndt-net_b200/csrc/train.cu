// train.cu — train-mode forward + backward of NDTNetSegmentation on the device (BASELINE config 3).
//
// Restates, as explicit kernels, what torch autograd does for the reference's training step
// (/root/reference/tools/train.py:66-76 -> ndnet/models/ndtnet.py:33-62 TNet, :112-164 NDTNet, :218-243
// NDTNetSegmentation) with BatchNorm in TRAINING mode (batch statistics over all rows, running-stat update with
// momentum 0.1 and the unbiased variance, eps 1e-5): every Conv1d(k=1)/Linear is a GEMM over rows = normal
// distributions, followed by column statistics, a normalise(+ReLU) pass, and in the backward pass the matching
// column reductions, the BN/ReLU gradient, dgrad and wgrad GEMMs.  The loss stays outside (torch, [B,N,C+1] tensor).
//
// Round-1 state of this row: everything is fp32 on the CUDA cores (one generic strided 64x64x16 SIMT GEMM with
// split-K, double-precision column sums), chosen so that parity against torch's fp32 autograd is tight (1e-4) before the
// GEMMs move to tcgen05.  At config 3's per-GPU size (2 clouds x 1000 rows) the step is launch-latency bound, not
// FLOP bound.  Rows are [B*N, C] row-major fp32 throughout.
#include "../../include/ndnet_b200.h"
#include "train_gemm.cuh"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

namespace ndt { void count_launches(long n); long launches(); }


namespace train {

constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;

// ------------------------------------------------------------------------------------------------ GEMM
// C[i, j] (+)= sum_k A(i, k) * B(j, k) (+ bias[j]);  A(i,k) = A[i*sai + k*sak], B(j,k) = B[j*sbj + k*sbk]; batched by z.
struct GemmP {
    const float *A; long sai, sak, sab;
    const float *B; long sbj, sbk, sbb;
    float *C; long ldc, scb;
    const float *bias;
    int M, N, K;
    int accumulate, splitk, kchunk;
};

constexpr int kTile = 64, kTileK = 16;

__global__ void __launch_bounds__(256) k_gemm32(const GemmP p) {
    __shared__ float As[kTileK][kTile + 4];
    __shared__ float Bs[kTileK][kTile + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int batch = blockIdx.z / p.splitk, split = blockIdx.z % p.splitk;
    const int i0 = blockIdx.y * kTile, j0 = blockIdx.x * kTile;
    const float *A = p.A + batch * p.sab, *B = p.B + batch * p.sbb;
    float *C = p.C + batch * p.scb;
    const int k_begin = split * p.kchunk;
    const int k_end = min(p.K, k_begin + p.kchunk);
    const bool a_kfast = p.sak == 1, b_kfast = p.sbk == 1;
    float acc[4][4] = {};
    for (int k0 = k_begin; k0 < k_end; k0 += kTileK) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int e = tid + r * 256;
            int i, k;
            if (a_kfast) { i = e >> 4; k = e & 15; } else { i = e & 63; k = e >> 6; }
            const int gi = i0 + i, gk = k0 + k;
            As[k][i] = (gi < p.M && gk < k_end) ? A[gi * p.sai + gk * p.sak] : 0.f;
            if (b_kfast) { i = e >> 4; k = e & 15; } else { i = e & 63; k = e >> 6; }
            const int gj = j0 + i, gk2 = k0 + k;
            Bs[k][i] = (gj < p.N && gk2 < k_end) ? B[gj * p.sbj + gk2 * p.sbk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTileK; k++) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int v = 0; v < 4; v++) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int gi = i0 + ty * 4 + u;
        if (gi >= p.M) continue;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const int gj = j0 + tx * 4 + v;
            if (gj >= p.N) continue;
            float r = acc[u][v];
            if (p.bias && split == 0) r += p.bias[gj];
            float *dst = C + (long)gi * p.ldc + gj;
            if (p.splitk > 1) atomicAdd(dst, r);
            else if (p.accumulate) *dst += r;
            else *dst = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------ column reductions
enum { RED_STATS = 0, RED_BNBWD = 1, RED_SUM = 2 };

struct RedP {
    const float *P; long ldp;        // STATS: Y      BNBWD: dA      SUM: X
    const float *Y; long ldy;        //               BNBWD: Y (pre-BN)
    const float *Aact; long lda;     //               BNBWD: post-activation output (ReLU mask), null when no ReLU
    const float *mean, *rstd;        //               BNBWD
    int rows, cols;
    long batch_stride_rows;          // rows of P per z
    double *s0, *s1;                 // [z*cols + c]
};

template <int OP>
__global__ void __launch_bounds__(256) k_colred(const RedP p) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;                    // 0..7
    const long row_base = (long)blockIdx.z * p.batch_stride_rows;
    double a0 = 0.0, a1 = 0.0;
    if (c < p.cols) {
        float mu = 0.f, rs = 0.f;
        if (OP == RED_BNBWD) { mu = p.mean[c]; rs = p.rstd[c]; }
        for (int r = blockIdx.y * 8 + ry; r < p.rows; r += gridDim.y * 8) {
            const long row = row_base + r;
            const float v = p.P[row * p.ldp + c];
            if (OP == RED_STATS) { a0 += (double)v; a1 += (double)v * (double)v; }
            else if (OP == RED_SUM) { a0 += (double)v; }
            else {
                const float m = (p.Aact == nullptr || p.Aact[row * p.lda + c] > 0.f) ? v : 0.f;
                const float xhat = (p.Y[row * p.ldy + c] - mu) * rs;
                a0 += (double)m; a1 += (double)m * (double)xhat;
            }
        }
    }
    __shared__ double s[2][8][32];
    s[0][ry][threadIdx.x & 31] = a0; s[1][ry][threadIdx.x & 31] = a1;
    __syncthreads();
    if (ry == 0 && c < p.cols) {
        for (int k = 1; k < 8; k++) { a0 += s[0][k][threadIdx.x]; a1 += s[1][k][threadIdx.x]; }
        atomicAdd(&p.s0[(long)blockIdx.z * p.cols + c], a0);
        if (OP != RED_SUM) atomicAdd(&p.s1[(long)blockIdx.z * p.cols + c], a1);
    }
}

// BatchNorm1d training statistics (ndtnet.py:28-32 modules in train(); torch semantics)
__global__ void k_bn_finalize(const double *__restrict__ s0, const double *__restrict__ s1, int rows, int C, float *__restrict__ mean,
                              float *__restrict__ rstd, float *__restrict__ run_mean, float *__restrict__ run_var,
                              long long *__restrict__ nbt, int update) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && update && nbt) *nbt += 1;
    if (c >= C) return;
    const double mu = s0[c] / rows;
    double var = s1[c] / rows - mu * mu;
    if (var < 0) var = 0;
    mean[c] = (float)mu;
    rstd[c] = (float)(1.0 / sqrt(var + (double)kBnEps));
    if (update) {
        const double unbiased = rows > 1 ? var * rows / (rows - 1) : var;
        run_mean[c] = (1.f - kBnMomentum) * run_mean[c] + kBnMomentum * (float)mu;
        run_var[c] = (1.f - kBnMomentum) * run_var[c] + kBnMomentum * (float)unbiased;
    }
}

__global__ void k_bn_apply(const float *__restrict__ Y, long rows, int C, const float *__restrict__ mean, const float *__restrict__ rstd,
                           const float *__restrict__ g, const float *__restrict__ be, int relu, float *__restrict__ A) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * C) return;
    const int c = (int)(t % C);
    float v = (Y[t] - mean[c]) * rstd[c] * g[c] + be[c];
    if (relu && v < 0.f) v = 0.f;
    A[t] = v;
}

__global__ void k_bnbwd_finalize(const double *__restrict__ s0, const double *__restrict__ s1, int C, float *__restrict__ dbeta,
                                 float *__restrict__ dgamma) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)s0[c];
    dgamma[c] = (float)s1[c];
}

// in place: dA -> dY
__global__ void k_bnbwd_apply(float *__restrict__ dA, const float *__restrict__ Y, const float *__restrict__ Aact, long rows, int C,
                              const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ g,
                              const double *__restrict__ s0, const double *__restrict__ s1) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * C) return;
    const int c = (int)(t % C);
    const float m = (Aact == nullptr || Aact[t] > 0.f) ? dA[t] : 0.f;
    const float xhat = (Y[t] - mean[c]) * rstd[c];
    const float inv = 1.f / (float)rows;
    dA[t] = g[c] * rstd[c] * (m - (float)s0[c] * inv - xhat * (float)s1[c] * inv);
}

__global__ void k_sum_finalize(const double *__restrict__ s0, long n, float *__restrict__ out) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = (float)s0[t];
}

// ------------------------------------------------------------------------------------------------ pooling, transforms, head
// max over the N rows of each cloud (torch.max(x, 2), ndtnet.py:50,224) with the first arg-max
__global__ void __launch_bounds__(256) k_maxpool_fwd(const float *__restrict__ A, int N, int C, float *__restrict__ G, int *__restrict__ idx) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5, b = blockIdx.y;
    float best = -INFINITY;
    int bi = 0x7fffffff;                                 // "nothing seen yet"
    if (c < C)
        for (int n = ry; n < N; n += 8) {                // ascending n: a strict '>' keeps the first maximum
            const float v = A[((long)b * N + n) * C + c];
            if (bi == 0x7fffffff || v > best) { best = v; bi = n; }
        }
    __shared__ float sv[8][32];
    __shared__ int si[8][32];
    sv[ry][threadIdx.x & 31] = best; si[ry][threadIdx.x & 31] = bi;
    __syncthreads();
    if (ry == 0 && c < C) {
        for (int k = 1; k < 8; k++) {
            const float v = sv[k][threadIdx.x];
            const int i = si[k][threadIdx.x];
            if (i != 0x7fffffff && (v > best || (v == best && i < bi) || bi == 0x7fffffff)) { best = v; bi = i; }
        }
        G[(long)b * C + c] = best;
        idx[(long)b * C + c] = bi;
    }
}

__global__ void k_maxpool_bwd(const float *__restrict__ dG, const int *__restrict__ idx, int N, int C, long total, float *__restrict__ dA) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C);
    const long row = t / C;
    const int b = (int)(row / N), n = (int)(row % N);
    dA[t] = idx[(long)b * C + c] == n ? dG[(long)b * C + c] : 0.f;
}

// p' = T p, cov' = T cov (3x3, row-major)  (ndtnet.py:131-146); feat row = [p(3) | cov(9)]
__global__ void k_apply_t_fwd(const float *__restrict__ feat, const float *__restrict__ T, long rows, int N, float *__restrict__ X12) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *t = T + (r / N) * 9, *f = feat + r * 12;
    float *o = X12 + r * 12;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        o[i] = t[i * 3 + 0] * f[0] + t[i * 3 + 1] * f[1] + t[i * 3 + 2] * f[2];
#pragma unroll
        for (int j = 0; j < 3; j++)
            o[3 + i * 3 + j] = t[i * 3 + 0] * f[3 + j] + t[i * 3 + 1] * f[6 + j] + t[i * 3 + 2] * f[9 + j];
    }
}

// dT[b][i][k] = sum_n ( dp'[i] p[k] + sum_j dcov'[i][j] cov[k][j] )
__global__ void __launch_bounds__(256) k_apply_t_bwd(const float *__restrict__ feat, const float *__restrict__ dX12, int N, float *__restrict__ dT) {
    const int b = blockIdx.x;
    float acc[9] = {};
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float *f = feat + ((long)b * N + n) * 12, *d = dX12 + ((long)b * N + n) * 12;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float v = d[i] * f[k];
#pragma unroll
                for (int j = 0; j < 3; j++) v += d[3 + i * 3 + j] * f[3 + k * 3 + j];
                acc[i * 3 + k] += v;
            }
    }
    __shared__ float s[9][256];
    for (int q = 0; q < 9; q++) s[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) for (int q = 0; q < 9; q++) s[q][threadIdx.x] += s[q][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 9) dT[b * 9 + threadIdx.x] = s[threadIdx.x][0];
}

__global__ void k_add_identity(float *__restrict__ T, int B, int d) {       // ndtnet.py:59
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B * d) T[(long)(t / d) * d * d + (t % d) * (d + 1)] += 1.f;
}

// H0 = [x_t2 | global feature of the cloud]  (ndtnet.py:227-230)
__global__ void k_concat_fwd(const float *__restrict__ X2, const float *__restrict__ G, long rows, int N, int F, float *__restrict__ H0) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int W = 64 + F;
    if (t >= rows * W) return;
    const long r = t / W;
    const int c = (int)(t % W);
    H0[t] = c < 64 ? X2[r * 64 + c] : G[(r / N) * F + (c - 64)];
}

__global__ void k_logsm_fwd(const float *__restrict__ Z, long rows, int C, float *__restrict__ out) {       // ndtnet.py:239
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *z = Z + r * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; c++) mx = fmaxf(mx, z[c]);
    float s = 0.f;
    for (int c = 0; c < C; c++) s += expf(z[c] - mx);
    const float l = mx + logf(s);
    for (int c = 0; c < C; c++) out[r * C + c] = z[c] - l;
}

__global__ void k_logsm_bwd(const float *__restrict__ dlogp, const float *__restrict__ logp, long rows, int C, float *__restrict__ dZ) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float s = 0.f;
    for (int c = 0; c < C; c++) s += dlogp[r * C + c];
    for (int c = 0; c < C; c++) dZ[r * C + c] = dlogp[r * C + c] - expf(logp[r * C + c]) * s;
}

// softmax over the C outputs of a row (ndtnet.py:194, pointnet.py:161) and its backward: dZ = p (dP - sum_c dP p)
__global__ void k_softmax_fwd(const float *__restrict__ Z, long rows, int C, float *__restrict__ out) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *z = Z + r * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; c++) mx = fmaxf(mx, z[c]);
    float s = 0.f;
    for (int c = 0; c < C; c++) s += expf(z[c] - mx);
    const float inv = 1.f / s;
    for (int c = 0; c < C; c++) out[r * C + c] = expf(z[c] - mx) * inv;
}

__global__ void k_softmax_bwd(const float *__restrict__ dP, const float *__restrict__ P, long rows, int C, float *__restrict__ dZ) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float s = 0.f;
    for (int c = 0; c < C; c++) s += dP[r * C + c] * P[r * C + c];
    for (int c = 0; c < C; c++) dZ[r * C + c] = P[r * C + c] * (dP[r * C + c] - s);
}

// ReLU of a layer without BatchNorm (the classification heads, ndtnet.py:189-190) and its backward (in place on dA)
__global__ void k_relu_fwd(const float *__restrict__ Y, long n, float *__restrict__ A) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) A[t] = fmaxf(Y[t], 0.f);
}
__global__ void k_relu_bwd(float *__restrict__ dA, const float *__restrict__ A, long n) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && !(A[t] > 0.f)) dA[t] = 0.f;
}

// torch.nan_to_num(x, nan=0.0) (pointnet.py:117) in place: NaN -> 0, +-inf -> +-FLT_MAX; `finite` remembers which entries
// were left alone (the gradient of the others is zero)
__global__ void k_nan_to_num_fwd(float *__restrict__ X, long n, unsigned char *__restrict__ finite) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float v = X[t];
    const bool ok = isfinite(v);
    finite[t] = ok ? 1 : 0;
    if (!ok) X[t] = v != v ? 0.f : (v > 0.f ? 3.402823466e+38f : -3.402823466e+38f);
}
__global__ void k_nan_to_num_bwd(float *__restrict__ dX, long n, const unsigned char *__restrict__ finite) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && !finite[t]) dX[t] = 0.f;
}

// ------------------------------------------------------------------------------------------------ host side
struct Lin { int w = -1, b = -1, in = 0, out = 0; };
struct Bn { int g = -1, be = -1, rm = -1, rv = -1, nbt = -1; };

struct Block {            // Linear (+ BatchNorm (+ ReLU)) over `rows` rows
    Lin lin; Bn bn; bool has_bn = false, relu = false;
    long rows = 0;
    float *Y = nullptr, *A = nullptr, *dA = nullptr, *mean = nullptr, *rstd = nullptr;
    double *red = nullptr;                // 2*out doubles for the forward stats, reused by the backward reduction
    double *red_bias = nullptr;           // out doubles for the bias gradient
};

// every kernel launch of this file goes through the library-wide launch counter (bench.py's gpu_launches); under a CUDA
// graph the capture pass counts once and each replay adds the number counted then (run_graphed)
static inline void train_count() { ndt::count_launches(1); }

struct TNet {
    int d = 0;
    Block c1, c2, c3, f1, f2, f3;
    float *G = nullptr, *dG = nullptr; int *idx = nullptr;
    float *T = nullptr, *dT = nullptr;    // [B, d, d]
};

struct Trainer {
    int device = 0;
    std::map<std::string, int> index;
    std::vector<std::vector<int64_t>> shapes;
    int n_tensors = 0;
    int F = 0, C = 0;                     // feature_dim, network outputs per row (segmentation: num_classes + 1; classification: num_classes)
    int kind = 1;                         // 0 NDTNetClassification, 1 NDTNetSegmentation, 2 PointNetClassification, 3 PointNetSegmentation
    int D = 12;                           // input width: 12 = [mean | covariance] for the NDT networks, point_dim for PointNet
    bool is_seg() const { return kind == 1 || kind == 3; }
    bool is_ndt() const { return kind <= 1; }
    long out_elems() const { return is_seg() ? (long)B * N * C : (long)B * C; }
    TNet t1, t2;
    Block c1, c2, c3, h1, h2, h3, h4;     // h1..h4: segmentation head; classification head: h1..h3 over one row per cloud
    unsigned char *finite = nullptr;      // PointNet: nan_to_num mask of the transformed input
    int B = 0, N = 0;
    char *arena = nullptr; size_t arena_bytes = 0; size_t zero_begin = 0, zero_end = 0;
    float *X12 = nullptr, *dX12 = nullptr, *X2 = nullptr, *H0 = nullptr, *dH0 = nullptr, *Gf = nullptr, *dGf = nullptr, *dX4 = nullptr;
    double *red_gf = nullptr;
    int *idxf = nullptr;
    float *logp = nullptr, *dZ = nullptr;
    const float *feat = nullptr;
    // CUDA graphs: the ~100 (forward) / ~250 (backward) launches of one pass are captured once per (shape, pointer set) and
    // replayed; the first pass of a configuration runs eagerly (one-time attribute calls stay out of the capture)
    int use_graph = 0;
    cudaStream_t cap = nullptr;
    struct GraphSlot { cudaGraphExec_t exec = nullptr; std::vector<const void *> key; int seen = 0; long kernels = 0; } gf, gb;
    float *feat_static = nullptr, *dlogp_static = nullptr, *flat_grad = nullptr;
    std::vector<long> grad_off;           // per tensor: offset (floats) of its gradient in the flat buffer, -1 for buffers
    long flat_elems = 0;
    std::vector<float *> static_grads;
    // gradient buckets: the flat layout follows the order in which backward finishes the parameters (head, trunk + feature
    // T-Net, first layer + input T-Net); an event is recorded as each bucket completes, so that its all-reduce can start
    // while the rest of the backward still runs
    static constexpr int kBuckets = 3;
    long bucket_end[kBuckets] = {0, 0, 0};
    cudaEvent_t bucket_ev[kBuckets] = {nullptr, nullptr, nullptr};
    int defer_copy = 0;                   // graph mode: the caller fetches flat_grad bucket by bucket (bucket_ready) instead of one copy at the end
    int tf32 = 0;                         // 1: eligible GEMMs run on the tensor cores (kind::tf32), 0: fp32 FMA everywhere
    float *sT1 = nullptr, *sT2 = nullptr, *sW = nullptr;      // transposed dY / X / W for the tensor-core wgrad and dgrad
    long ldT = 0;
    std::string err;
};

static bool find(Trainer &t, const std::string &name, int &out, std::initializer_list<int64_t> shape, bool required = true) {
    auto it = t.index.find(name);
    if (it == t.index.end()) {
        if (required) t.err = "missing tensor " + name;
        return !required;
    }
    const auto &s = t.shapes[it->second];
    int64_t want = 1, have = 1;
    for (auto v : shape) want *= v;
    for (auto v : s) have *= v;
    if (want != have) { t.err = "unexpected shape of " + name; return false; }
    out = it->second;
    return true;
}

static bool bind_lin(Trainer &t, const std::string &name, int in, int out, Lin &l) {
    l.in = in; l.out = out;
    return find(t, name + ".weight", l.w, {out, in}) && find(t, name + ".bias", l.b, {out});
}

static bool bind_bn(Trainer &t, const std::string &name, int C, Bn &b) {
    return find(t, name + ".weight", b.g, {C}) && find(t, name + ".bias", b.be, {C}) && find(t, name + ".running_mean", b.rm, {C}) &&
           find(t, name + ".running_var", b.rv, {C}) && find(t, name + ".num_batches_tracked", b.nbt, {1}, false);
}

static bool bind_block(Trainer &t, Block &blk, const std::string &lin, const std::string &bn, int in, int out, bool relu) {
    blk.relu = relu;
    blk.has_bn = !bn.empty();
    if (!bind_lin(t, lin, in, out, blk.lin)) return false;
    return bn.empty() || bind_bn(t, bn, out, blk.bn);
}

static bool bind_tnet(Trainer &t, TNet &n, const std::string &p, int d) {
    n.d = d;
    return bind_block(t, n.c1, p + ".conv1", p + ".bn1", d, 64, true) && bind_block(t, n.c2, p + ".conv2", p + ".bn2", 64, 128, true) &&
           bind_block(t, n.c3, p + ".conv3", p + ".bn3", 128, 1024, true) && bind_block(t, n.f1, p + ".fc1", p + ".bn4", 1024, 512, true) &&
           bind_block(t, n.f2, p + ".fc2", p + ".bn5", 512, 256, true) && bind_block(t, n.f3, p + ".fc3", "", 256, d * d, false);
}

struct Bump {
    size_t off = 0;
    template <typename T> size_t take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        const size_t at = off;
        off += count * sizeof(T);
        return at;
    }
};

// two passes over the same layout: first to size the arena, then to hand out pointers
static void layout_block(Bump &b, char *base, Block &k, long rows, bool zero_region) {
    k.rows = rows;
    const size_t n = (size_t)rows * k.lin.out;
    if (!zero_region) {
        size_t o;
        o = b.take<float>(n); if (base) k.Y = (float *)(base + o);
        if (k.has_bn || k.relu) { o = b.take<float>(n); if (base) k.A = (float *)(base + o); } else if (base) k.A = k.Y;
        o = b.take<float>(n); if (base) k.dA = (float *)(base + o);
        o = b.take<float>(k.lin.out); if (base) k.mean = (float *)(base + o);
        o = b.take<float>(k.lin.out); if (base) k.rstd = (float *)(base + o);
    } else {
        size_t o;
        o = b.take<double>(4 * (size_t)k.lin.out); if (base) k.red = (double *)(base + o);
        o = b.take<double>(k.lin.out); if (base) k.red_bias = (double *)(base + o);
    }
}

static void layout(Trainer &t, Bump &b, char *base, int B, int N) {
    const long M = (long)B * N;
    std::vector<Block *> big = {&t.t1.c1, &t.t1.c2, &t.t1.c3, &t.t2.c1, &t.t2.c2, &t.t2.c3, &t.c1, &t.c2, &t.c3};
    std::vector<Block *> small = {&t.t1.f1, &t.t1.f2, &t.t1.f3, &t.t2.f1, &t.t2.f2, &t.t2.f3};
    for (Block *k : {&t.h1, &t.h2, &t.h3}) (t.is_seg() ? big : small).push_back(k);
    if (t.is_seg()) big.push_back(&t.h4);
    for (Block *k : big) layout_block(b, base, *k, M, false);
    for (Block *k : small) layout_block(b, base, *k, B, false);
    size_t o;
#define TAKE(T, ptr, count) o = b.take<T>(count); if (base) ptr = (T *)(base + o)
    for (TNet *n : {&t.t1, &t.t2}) {
        TAKE(float, n->G, (size_t)B * 1024); TAKE(float, n->dG, (size_t)B * 1024); TAKE(int, n->idx, (size_t)B * 1024);
        TAKE(float, n->T, (size_t)B * n->d * n->d); TAKE(float, n->dT, (size_t)B * n->d * n->d);
    }
    TAKE(float, t.X12, (size_t)M * t.D); TAKE(float, t.dX12, (size_t)M * t.D); TAKE(float, t.X2, (size_t)M * 64);
    if (!t.is_ndt()) { TAKE(unsigned char, t.finite, (size_t)M * t.D); }
    // (the classification heads keep dH0 too: its first 64 columns carry dX2 through the backward of the trunk)
    TAKE(float, t.H0, (size_t)M * (64 + t.F)); TAKE(float, t.dH0, (size_t)M * (64 + t.F));
    TAKE(float, t.Gf, (size_t)B * t.F); TAKE(float, t.dGf, (size_t)B * t.F); TAKE(float, t.dX4, (size_t)M * t.F);
    TAKE(int, t.idxf, (size_t)B * t.F);
    const size_t outs = t.is_seg() ? (size_t)M * t.C : (size_t)B * t.C;
    TAKE(float, t.logp, outs); TAKE(float, t.dZ, outs);
    TAKE(float, t.feat_static, (size_t)M * t.D); TAKE(float, t.dlogp_static, outs); TAKE(float, t.flat_grad, (size_t)t.flat_elems);
    {
        const long ldT = (M + 3) / 4 * 4;
        const size_t wide = (size_t)(64 + t.F > 1024 ? 64 + t.F : 1024);
        if (base) t.ldT = ldT;
        TAKE(float, t.sT1, wide * ldT); TAKE(float, t.sT2, wide * ldT);
        TAKE(float, t.sW, (size_t)1024 * (64 + t.F > 1024 ? 64 + t.F : 1024));
    }
    // the region zeroed at the start of every pass: reduction accumulators
    b.off = (b.off + 255) & ~(size_t)255;
    if (base) t.zero_begin = b.off;
    for (Block *k : big) layout_block(b, base, *k, M, true);
    for (Block *k : small) layout_block(b, base, *k, B, true);
    TAKE(double, t.red_gf, (size_t)B * t.F);
    b.off = (b.off + 255) & ~(size_t)255;
    if (base) t.zero_end = b.off;
#undef TAKE
}

static cudaError_t reserve(Trainer &t, int B, int N) {
    if (B == t.B && N == t.N && t.arena) return cudaSuccess;
    Bump size;
    layout(t, size, nullptr, B, N);
    if (size.off > t.arena_bytes) {
        if (t.arena) cudaFree(t.arena);
        t.arena = nullptr; t.arena_bytes = 0;
        cudaError_t e = cudaMalloc((void **)&t.arena, size.off);
        if (e != cudaSuccess) return e;
        t.arena_bytes = size.off;
    }
    Bump place;
    layout(t, place, t.arena, B, N);
    t.B = B; t.N = N;
    return cudaSuccess;
}

static inline unsigned cdiv(long a, long b) { return (unsigned)((a + b - 1) / b); }

static void gemm(cudaStream_t st, const float *A, long sai, long sak, const float *Bm, long sbj, long sbk, float *C, long ldc, int M, int N,
                 int K, const float *bias, bool accumulate, int batch = 1, long sab = 0, long sbb = 0, long scb = 0) {
    GemmP p;
    p.A = A; p.sai = sai; p.sak = sak; p.sab = sab;
    p.B = Bm; p.sbj = sbj; p.sbk = sbk; p.sbb = sbb;
    p.C = C; p.ldc = ldc; p.scb = scb; p.bias = bias; p.M = M; p.N = N; p.K = K; p.accumulate = accumulate ? 1 : 0;
    const long tiles = (long)cdiv(M, kTile) * cdiv(N, kTile) * batch;
    int splitk = 1;
    if (tiles < 148 && K >= 256) {
        splitk = (int)((296 + tiles - 1) / tiles);
        const int max_split = (K + 127) / 128;
        if (splitk > max_split) splitk = max_split;
        if (splitk < 1) splitk = 1;
    }
    p.splitk = splitk;
    p.kchunk = (int)(((long)cdiv(K, kTileK) + splitk - 1) / splitk) * kTileK;
    if (splitk > 1 && !accumulate) {
        // split-K adds into the output: clear it first (contiguous outputs only: wgrad and the per-cloud transforms)
        if (batch > 1 || ldc == N) cudaMemsetAsync(C, 0, sizeof(float) * (batch > 1 ? (size_t)scb * batch : (size_t)M * N), st);
        else cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st);
    }
    dim3 grid(cdiv(N, kTile), cdiv(M, kTile), batch * splitk);
    train_count(), k_gemm32<<<grid, 256, 0, st>>>(p);
}

// ---- tcgen05 / TF32 path (train_gemm.cuh) --------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp32 [rows, K] row-major (ld floats between rows), read as TF32 (TMA rounds to nearest on load), 128-byte K-blocks
static bool make_map_tf32(CUtensorMap *m, const float *ptr, long ld, int K, int rows, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 1};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * 4 * (cuuint64_t)rows};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float *>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool tf32_eligible(const float *A, long lda, const float *Bm, long ldb, int M, int N, int K) {
    return M >= 64 && N >= 32 && K >= 32 && lda % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)Bm & 15) == 0;
}

template <int BN>
static bool launch_tf32(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, const Tf32Args &a, dim3 grid) {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = tf32_smem_bytes<BN>();
    if (dev < 64 && !attr_done[dev]) {
        if (cudaFuncSetAttribute(k_gemm_tf32<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        attr_done[dev] = true;
    }
    train_count(), k_gemm_tf32<BN><<<grid, mlp::kGemmThreads, smem, st>>>(ma, mb, a);
    return true;
}

// C[M, N] (+)= A[M, K] . B[N, K]^T (+ bias) on the tensor cores; false when the shape / alignment is not eligible
static bool gemm_tf32(cudaStream_t st, const float *A, long lda, const float *Bm, long ldb, float *C, long ldc, int M, int N, int K,
                      const float *bias, bool accumulate, double *stat_sum = nullptr, double *stat_sq = nullptr, bool *stats_fused = nullptr) {
    if (stats_fused) *stats_fused = false;
    if (!tf32_eligible(A, lda, Bm, ldb, M, N, K)) return false;
    const int BN = N > 64 ? 128 : 64;
    CUtensorMap ma, mb;
    if (!make_map_tf32(&ma, A, lda, K, M, 128) || !make_map_tf32(&mb, Bm, ldb, K, N, BN)) return false;
    Tf32Args a;
    a.M = M; a.N = N; a.K = K; a.C = C; a.ldc = ldc; a.bias = bias; a.accumulate = accumulate ? 1 : 0;
    const int nkb = (K + 31) / 32;
    const long tiles = (long)cdiv(M, 128) * cdiv(N, BN);
    int splitk = 1;
    if (tiles < 74 && nkb >= 16) {
        splitk = (int)((148 + tiles - 1) / tiles);
        if (splitk > nkb / 4) splitk = nkb / 4;
        if (splitk > 32) splitk = 32;          // more splits only multiply the atomics into the same small output
        if (splitk < 1) splitk = 1;
    }
    a.kb_per_split = (nkb + splitk - 1) / splitk;
    a.splitk = (nkb + a.kb_per_split - 1) / a.kb_per_split;          // every z gets at least one K-block
    const bool fuse = stat_sum && stat_sq && a.splitk == 1 && !accumulate;      // BatchNorm statistics in the epilogue
    a.stat_sum = fuse ? stat_sum : nullptr; a.stat_sq = fuse ? stat_sq : nullptr;
    if (stats_fused) *stats_fused = fuse;
    if (a.splitk > 1 && !accumulate) {
        if (ldc == N) cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st);
        else cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st);
    }
    dim3 grid(cdiv(N, BN), cdiv(M, 128), a.splitk);
    return BN == 128 ? launch_tf32<128>(st, ma, mb, a, grid) : launch_tf32<64>(st, ma, mb, a, grid);
}

// dst[c, r] = src[r, c]  (dst row stride ldd >= rows, a multiple of 4 floats for the TMA)
__global__ void __launch_bounds__(256) k_transpose(const float *__restrict__ src, long lds, int rows, int cols, float *__restrict__ dst, long ldd) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int r = r0 + k, c = c0 + tx;
        tile[k][tx] = (r < rows && c < cols) ? src[(long)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, r = r0 + tx;
        if (c < cols && r < rows) dst[(long)c * ldd + r] = tile[tx][k];
    }
}

static void transpose(cudaStream_t st, const float *src, long lds, int rows, int cols, float *dst, long ldd) {
    train_count(), k_transpose<<<dim3(cdiv(cols, 32), cdiv(rows, 32)), 256, 0, st>>>(src, lds, rows, cols, dst, ldd);
}

template <int OP>
static void colred(cudaStream_t st, RedP p, int batch) {
    const unsigned ysplit = p.rows >= 4096 ? 32 : p.rows >= 512 ? 8 : 1;
    dim3 grid(cdiv(p.cols, 32), ysplit, batch);
    train_count(), k_colred<OP><<<grid, 256, 0, st>>>(p);
}

struct Pass {
    Trainer &t;
    float *const *tensors;
    float *const *grads;
    cudaStream_t st;
    int update_running;
    float *P(int i) const { return i >= 0 ? tensors[i] : nullptr; }
    float *Gr(int i) const { return (i >= 0 && grads) ? grads[i] : nullptr; }
};

static void block_fwd(const Pass &ps, Block &k, const float *X, long ldx) {
    const int out = k.lin.out, in = k.lin.in;
    // train-mode BatchNorm statistics (sum y, sum y^2 per channel): accumulated in the tensor-core GEMM's epilogue when it
    // runs unsplit, else by a column-reduction pass over Y (fp32 parity mode, split-K shapes, the small FC layers)
    bool stats_fused = false;
    if (!(ps.t.tf32 && gemm_tf32(ps.st, X, ldx, ps.P(k.lin.w), in, k.Y, out, (int)k.rows, out, in, ps.P(k.lin.b), false,
                                 k.has_bn ? k.red : nullptr, k.has_bn ? k.red + out : nullptr, &stats_fused)))
        gemm(ps.st, X, ldx, 1, ps.P(k.lin.w), in, 1, k.Y, out, (int)k.rows, out, in, ps.P(k.lin.b), false);
    if (!k.has_bn) {
        if (k.relu) train_count(), k_relu_fwd<<<cdiv(k.rows * out, 256), 256, 0, ps.st>>>(k.Y, k.rows * out, k.A);
        return;
    }
    if (!stats_fused) {
        RedP r{};
        r.P = k.Y; r.ldp = out; r.rows = (int)k.rows; r.cols = out; r.s0 = k.red; r.s1 = k.red + out;
        colred<RED_STATS>(ps.st, r, 1);
    }
    train_count(), k_bn_finalize<<<cdiv(out, 128), 128, 0, ps.st>>>(k.red, k.red + out, (int)k.rows, out, k.mean, k.rstd, ps.P(k.bn.rm), ps.P(k.bn.rv),
                                                     (long long *)ps.P(k.bn.nbt), ps.update_running);
    train_count(), k_bn_apply<<<cdiv(k.rows * out, 256), 256, 0, ps.st>>>(k.Y, k.rows, out, k.mean, k.rstd, ps.P(k.bn.g), ps.P(k.bn.be), k.relu ? 1 : 0, k.A);
}

// k.dA holds dL/dA on entry; on return it holds dL/dY.  dX (optional) receives or accumulates dL/dX.
static void block_bwd(const Pass &ps, Block &k, const float *X, long ldx, float *dX, long lddx, bool accumulate_dx) {
    const int out = k.lin.out, in = k.lin.in;
    if (k.has_bn) {
        double *s0 = k.red + 2 * out, *s1 = k.red + 3 * out;
        RedP r{};
        r.P = k.dA; r.ldp = out; r.Y = k.Y; r.ldy = out; r.Aact = k.relu ? k.A : nullptr; r.lda = out; r.mean = k.mean; r.rstd = k.rstd;
        r.rows = (int)k.rows; r.cols = out; r.s0 = s0; r.s1 = s1;
        colred<RED_BNBWD>(ps.st, r, 1);
        if (ps.Gr(k.bn.be) && ps.Gr(k.bn.g)) train_count(), k_bnbwd_finalize<<<cdiv(out, 128), 128, 0, ps.st>>>(s0, s1, out, ps.Gr(k.bn.be), ps.Gr(k.bn.g));
        train_count(), k_bnbwd_apply<<<cdiv(k.rows * out, 256), 256, 0, ps.st>>>(k.dA, k.Y, k.relu ? k.A : nullptr, k.rows, out, k.mean, k.rstd, ps.P(k.bn.g), s0, s1);
    } else if (k.relu) {
        train_count(), k_relu_bwd<<<cdiv(k.rows * out, 256), 256, 0, ps.st>>>(k.dA, k.A, k.rows * out);
    }
    float *dY = k.dA;
    if (ps.Gr(k.lin.b) && k.has_bn) {
        // a bias in front of a train-mode BatchNorm has no effect on the output: its gradient, the column sums of the
        // BatchNorm input gradient, is identically zero (autograd returns the rounding noise of that sum)
        cudaMemsetAsync(ps.Gr(k.lin.b), 0, sizeof(float) * out, ps.st);
    } else if (ps.Gr(k.lin.b)) {
        RedP r{};
        r.P = dY; r.ldp = out; r.rows = (int)k.rows; r.cols = out; r.s0 = k.red_bias;
        colred<RED_SUM>(ps.st, r, 1);
        train_count(), k_sum_finalize<<<cdiv(out, 128), 128, 0, ps.st>>>(k.red_bias, out, ps.Gr(k.lin.b));
    }
    Trainer &t = ps.t;
    const int rows = (int)k.rows;
    if (ps.Gr(k.lin.w)) {    // dW[o, i] = sum_rows dY[r, o] X[r, i]
        bool done = false;
        if (t.tf32 && rows >= 256 && out >= 64 && in >= 32) {      // contraction over the rows: both operands transposed to K-major
            transpose(ps.st, dY, out, rows, out, t.sT1, t.ldT);
            transpose(ps.st, X, ldx, rows, in, t.sT2, t.ldT);
            done = gemm_tf32(ps.st, t.sT1, t.ldT, t.sT2, t.ldT, ps.Gr(k.lin.w), in, out, in, rows, nullptr, false);
        }
        if (!done) gemm(ps.st, dY, 1, out, X, 1, ldx, ps.Gr(k.lin.w), in, out, in, rows, nullptr, false);
    }
    if (dX) {                // dX[r, i] = sum_o dY[r, o] W[o, i]
        bool done = false;
        if (t.tf32 && rows >= 256 && in >= 32 && out >= 32 && out % 4 == 0) {
            transpose(ps.st, ps.P(k.lin.w), in, out, in, t.sW, out);          // W^T [in, out]
            done = gemm_tf32(ps.st, dY, out, t.sW, out, dX, lddx, rows, in, out, nullptr, accumulate_dx);
        }
        if (!done) gemm(ps.st, dY, out, 1, ps.P(k.lin.w), 1, in, dX, lddx, rows, in, out, nullptr, accumulate_dx);
    }
}

static void tnet_fwd(const Pass &ps, TNet &n, const float *X, long ldx) {
    Trainer &t = ps.t;
    block_fwd(ps, n.c1, X, ldx);
    block_fwd(ps, n.c2, n.c1.A, 64);
    block_fwd(ps, n.c3, n.c2.A, 128);
    train_count(), k_maxpool_fwd<<<dim3(cdiv(1024, 32), t.B), 256, 0, ps.st>>>(n.c3.A, t.N, 1024, n.G, n.idx);
    block_fwd(ps, n.f1, n.G, 1024);
    block_fwd(ps, n.f2, n.f1.A, 512);
    block_fwd(ps, n.f3, n.f2.A, 256);
    cudaMemcpyAsync(n.T, n.f3.Y, sizeof(float) * t.B * n.d * n.d, cudaMemcpyDeviceToDevice, ps.st);
    train_count(), k_add_identity<<<cdiv((long)t.B * n.d, 128), 128, 0, ps.st>>>(n.T, t.B, n.d);
}

// n.dT holds dL/dT on entry
static void tnet_bwd(const Pass &ps, TNet &n, const float *X, long ldx, float *dX, long lddx, bool accumulate_dx) {
    Trainer &t = ps.t;
    const long M = (long)t.B * t.N;
    cudaMemcpyAsync(n.f3.dA, n.dT, sizeof(float) * t.B * n.d * n.d, cudaMemcpyDeviceToDevice, ps.st);
    block_bwd(ps, n.f3, n.f2.A, 256, n.f2.dA, 256, false);
    block_bwd(ps, n.f2, n.f1.A, 512, n.f1.dA, 512, false);
    block_bwd(ps, n.f1, n.G, 1024, n.dG, 1024, false);
    train_count(), k_maxpool_bwd<<<cdiv(M * 1024, 256), 256, 0, ps.st>>>(n.dG, n.idx, t.N, 1024, M * 1024, n.c3.dA);
    block_bwd(ps, n.c3, n.c2.A, 128, n.c2.dA, 128, false);
    block_bwd(ps, n.c2, n.c1.A, 64, n.c1.dA, 64, false);
    block_bwd(ps, n.c1, X, ldx, dX, lddx, accumulate_dx);
}

static int forward(Trainer &t, const float *feat, int B, int N, float *const *tensors, float *out_logp, int update_running, cudaStream_t st) {
    cudaError_t e = reserve(t, B, N);
    if (e != cudaSuccess) { t.err = std::string("workspace: ") + cudaGetErrorString(e); return -100 - (int)e; }
    const long M = (long)B * N;
    Pass ps{t, tensors, nullptr, st, update_running};
    t.feat = feat;
    const int D = t.D;
    cudaMemsetAsync(t.arena + t.zero_begin, 0, t.zero_end - t.zero_begin, st);
    tnet_fwd(ps, t.t1, feat, D);                                                          // ndtnet.py:131-132, pointnet.py:113
    if (t.is_ndt()) {
        train_count(), k_apply_t_fwd<<<cdiv(M, 128), 128, 0, st>>>(feat, t.t1.T, M, N, t.X12);          // ndtnet.py:134-146
    } else {
        // x' = T1 x per cloud, rows: x'[n, i] = sum_k x[n, k] T1[i, k]; then nan_to_num (pointnet.py:115-117)
        gemm(st, feat, D, 1, t.t1.T, D, 1, t.X12, D, N, D, D, nullptr, false, B, (long)N * D, (long)D * D, (long)N * D);
        train_count(), k_nan_to_num_fwd<<<cdiv(M * D, 256), 256, 0, st>>>(t.X12, M * D, t.finite);
    }
    block_fwd(ps, t.c1, t.X12, D);                                                        // :149
    tnet_fwd(ps, t.t2, t.c1.A, 64);                                                       // :152
    gemm(st, t.c1.A, 64, 1, t.t2.T, 1, 64, t.X2, 64, N, 64, 64, nullptr, false, B, (long)N * 64, 4096, (long)N * 64);   // :153-155
    block_fwd(ps, t.c2, t.X2, 64);                                                        // :160
    block_fwd(ps, t.c3, t.c2.A, 128);                                                     // :161
    train_count(), k_maxpool_fwd<<<dim3(cdiv(t.F, 32), B), 256, 0, st>>>(t.c3.A, N, t.F, t.Gf, t.idxf); // :186, :224
    if (t.is_seg()) {
        train_count(), k_concat_fwd<<<cdiv(M * (64 + t.F), 256), 256, 0, st>>>(t.X2, t.Gf, M, N, t.F, t.H0);   // :227-230
        block_fwd(ps, t.h1, t.H0, 64 + t.F);                                              // :233
        block_fwd(ps, t.h2, t.h1.A, 512);
        block_fwd(ps, t.h3, t.h2.A, 256);
        block_fwd(ps, t.h4, t.h3.A, 128);                                                 // :236
        train_count(), k_logsm_fwd<<<cdiv(M, 128), 128, 0, st>>>(t.h4.Y, M, t.C, t.logp);                // :239
    } else {
        block_fwd(ps, t.h1, t.Gf, t.F);                                                   // ndtnet.py:189-191: one row per cloud
        block_fwd(ps, t.h2, t.h1.A, 512);
        block_fwd(ps, t.h3, t.h2.A, 256);
        train_count(), k_softmax_fwd<<<cdiv(B, 128), 128, 0, st>>>(t.h3.Y, B, t.C, t.logp);              // :194 (probabilities, not logs)
    }
    if (out_logp) cudaMemcpyAsync(out_logp, t.logp, sizeof(float) * t.out_elems(), cudaMemcpyDeviceToDevice, st);
    e = cudaGetLastError();
    if (e != cudaSuccess) { t.err = std::string("forward: ") + cudaGetErrorString(e); return -100 - (int)e; }
    return 0;
}

// the gradients of bucket i are final: record its event (an external-event node when the pass is being captured)
static void record_bucket(Trainer &t, int i, cudaStream_t st) {
    if (!t.bucket_ev[i] && cudaEventCreateWithFlags(&t.bucket_ev[i], cudaEventDisableTiming) != cudaSuccess) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    cudaEventRecordWithFlags(t.bucket_ev[i], st, cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
}

static int backward(Trainer &t, const float *dlogp, float *const *tensors, float *const *grads, cudaStream_t st) {
    if (!t.arena || !t.feat) { t.err = "backward without forward"; return -204; }
    const int B = t.B, N = t.N;
    const long M = (long)B * N;
    const int W = 64 + t.F;
    Pass ps{t, tensors, grads, st, 0};
    // the forward's column sums are already folded into mean/rstd: clear every accumulator for this pass
    cudaMemsetAsync(t.arena + t.zero_begin, 0, t.zero_end - t.zero_begin, st);
    const int D = t.D;
    if (t.is_seg()) {
        train_count(), k_logsm_bwd<<<cdiv(M, 128), 128, 0, st>>>(dlogp, t.logp, M, t.C, t.h4.dA);
        block_bwd(ps, t.h4, t.h3.A, 128, t.h3.dA, 128, false);
        block_bwd(ps, t.h3, t.h2.A, 256, t.h2.dA, 256, false);
        block_bwd(ps, t.h2, t.h1.A, 512, t.h1.dA, 512, false);
        block_bwd(ps, t.h1, t.H0, W, t.dH0, W, false);
        record_bucket(t, 0, st);
        // dH0 = [dX2 | per-row gradient of the broadcast global feature]
        RedP r{};
        r.P = t.dH0 + 64; r.ldp = W; r.rows = N; r.cols = t.F; r.batch_stride_rows = N; r.s0 = t.red_gf;
        colred<RED_SUM>(st, r, B);
        train_count(), k_sum_finalize<<<cdiv((long)B * t.F, 128), 128, 0, st>>>(t.red_gf, (long)B * t.F, t.dGf);
    } else {
        train_count(), k_softmax_bwd<<<cdiv(B, 128), 128, 0, st>>>(dlogp, t.logp, B, t.C, t.h3.dA);
        block_bwd(ps, t.h3, t.h2.A, 256, t.h2.dA, 256, false);
        block_bwd(ps, t.h2, t.h1.A, 512, t.h1.dA, 512, false);
        block_bwd(ps, t.h1, t.Gf, t.F, t.dGf, t.F, false);
        record_bucket(t, 0, st);
        // x_t2 only feeds the trunk here: its gradient (first 64 columns of dH0) starts at zero
        cudaMemset2DAsync(t.dH0, sizeof(float) * W, 0, sizeof(float) * 64, M, st);
    }
    train_count(), k_maxpool_bwd<<<cdiv(M * t.F, 256), 256, 0, st>>>(t.dGf, t.idxf, N, t.F, M * t.F, t.c3.dA);
    block_bwd(ps, t.c3, t.c2.A, 128, t.c2.dA, 128, false);
    block_bwd(ps, t.c2, t.X2, 64, t.dH0, W, true);                     // dX2 lives in the first 64 columns of dH0
    // x2 = x1 . T2 per cloud:  dX1 = dX2 . T2^T ;  dT2 = X1^T . dX2
    gemm(st, t.dH0, W, 1, t.t2.T, 64, 1, t.c1.dA, 64, N, 64, 64, nullptr, false, B, (long)N * W, 4096, (long)N * 64);
    gemm(st, t.c1.A, 1, 64, t.dH0, 1, W, t.t2.dT, 64, 64, 64, N, nullptr, false, B, (long)N * 64, (long)N * W, 4096);
    tnet_bwd(ps, t.t2, t.c1.A, 64, t.c1.dA, 64, true);
    record_bucket(t, 1, st);
    block_bwd(ps, t.c1, t.X12, D, t.dX12, D, false);
    if (t.is_ndt()) {
        train_count(), k_apply_t_bwd<<<B, 256, 0, st>>>(t.feat, t.dX12, N, t.t1.dT);
    } else {
        // x' = nan_to_num(T1 x): dT1[i, k] = sum_n dx'[n, i] x[n, k] over the entries nan_to_num left alone
        train_count(), k_nan_to_num_bwd<<<cdiv(M * D, 256), 256, 0, st>>>(t.dX12, M * D, t.finite);
        gemm(st, t.dX12, 1, D, t.feat, 1, D, t.t1.dT, D, D, D, N, nullptr, false, B, (long)N * D, (long)N * D, (long)D * D);
    }
    tnet_bwd(ps, t.t1, t.feat, D, nullptr, 0, false);
    record_bucket(t, 2, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { t.err = std::string("backward: ") + cudaGetErrorString(e); return -100 - (int)e; }
    return 0;
}

// ---- CUDA-graph replay of a pass ------------------------------------------------------------------------------------
template <typename Body>
static int run_graphed(Trainer &t, Trainer::GraphSlot &g, const std::vector<const void *> &key, cudaStream_t st, Body body) {
    if (g.key != key) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g.exec = nullptr; g.key = key; g.seen = 0;
    }
    if (g.seen == 0) { g.seen = 1; return body(st); }              // first pass of this configuration: eager
    if (g.seen == 1) {
        if (!t.cap && cudaStreamCreateWithFlags(&t.cap, cudaStreamNonBlocking) != cudaSuccess) return body(st);
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(t.cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return body(st); }
        const long before = ndt::launches();
        const int rc = body(t.cap);
        g.kernels = ndt::launches() - before;
        cudaError_t e = cudaStreamEndCapture(t.cap, &graph);
        if (rc != 0 || e != cudaSuccess || !graph) { cudaGetLastError(); if (graph) cudaGraphDestroy(graph); g.seen = 0; return rc != 0 ? rc : body(st); }
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; g.seen = 0; return body(st); }
        g.seen = 2;
    } else {
        ndt::count_launches(g.kernels);
    }
    const cudaError_t e = cudaGraphLaunch(g.exec, st);
    if (e != cudaSuccess) { t.err = std::string("graph launch: ") + cudaGetErrorString(e); return -100 - (int)e; }
    return 0;
}

static std::vector<const void *> pass_key(const Trainer &t, float *const *tensors, int extra) {
    std::vector<const void *> key;
    key.reserve(t.n_tensors + 4);
    key.push_back((const void *)t.arena);
    key.push_back((const void *)(uintptr_t)(((uint64_t)t.B << 32) | (uint32_t)t.N));
    key.push_back((const void *)(uintptr_t)((t.tf32 << 8) | extra));
    for (int i = 0; i < t.n_tensors; i++) key.push_back(tensors[i]);
    return key;
}

static int forward_entry(Trainer &t, const float *feat, int B, int N, float *const *tensors, float *out_logp, int update_running, cudaStream_t st) {
    if (!t.use_graph) return forward(t, feat, B, N, tensors, out_logp, update_running, st);
    cudaError_t e = reserve(t, B, N);
    if (e != cudaSuccess) { t.err = std::string("workspace: ") + cudaGetErrorString(e); return -100 - (int)e; }
    const long M = (long)B * N;
    cudaMemcpyAsync(t.feat_static, feat, sizeof(float) * M * t.D, cudaMemcpyDeviceToDevice, st);
    const int rc = run_graphed(t, t.gf, pass_key(t, tensors, update_running ? 1 : 0), st, [&](cudaStream_t s) {
        return forward(t, t.feat_static, B, N, tensors, nullptr, update_running, s);
    });
    if (rc != 0) return rc;
    cudaMemcpyAsync(out_logp, t.logp, sizeof(float) * t.out_elems(), cudaMemcpyDeviceToDevice, st);
    return 0;
}

// gradients of all parameters into one flat buffer (layout: Trainer::grad_off)
static int backward_flat(Trainer &t, const float *dlogp, float *const *tensors, float *flat_out, cudaStream_t st) {
    if (!t.arena || !t.feat) { t.err = "backward without forward"; return -204; }
    std::vector<float *> grads(t.n_tensors, nullptr);
    if (!t.use_graph) {
        for (int i = 0; i < t.n_tensors; i++) if (t.grad_off[i] >= 0) grads[i] = flat_out + t.grad_off[i];
        return backward(t, dlogp, tensors, grads.data(), st);
    }
    cudaMemcpyAsync(t.dlogp_static, dlogp, sizeof(float) * t.out_elems(), cudaMemcpyDeviceToDevice, st);
    t.static_grads.assign(t.n_tensors, nullptr);
    for (int i = 0; i < t.n_tensors; i++) if (t.grad_off[i] >= 0) t.static_grads[i] = t.flat_grad + t.grad_off[i];
    const int rc = run_graphed(t, t.gb, pass_key(t, tensors, 2), st, [&](cudaStream_t s) {
        return backward(t, t.dlogp_static, tensors, t.static_grads.data(), s);
    });
    if (rc != 0) return rc;
    if (!t.defer_copy) cudaMemcpyAsync(flat_out, t.flat_grad, sizeof(float) * t.flat_elems, cudaMemcpyDeviceToDevice, st);
    return 0;
}

// makes `side` wait until bucket i of the last backward is final and (graph mode with deferred copy) copies it from the
// library's static buffer into the caller's flat buffer on `side`
static int bucket_ready(Trainer &t, int i, float *flat_out, cudaStream_t side) {
    if (i < 0 || i >= Trainer::kBuckets || !t.bucket_ev[i]) { t.err = "bucket_ready: no backward pass recorded"; return -204; }
    cudaError_t e = cudaStreamWaitEvent(side, t.bucket_ev[i], 0);
    if (e != cudaSuccess) { t.err = std::string("bucket wait: ") + cudaGetErrorString(e); return -100 - (int)e; }
    if (t.use_graph && t.defer_copy) {
        const long b0 = i == 0 ? 0 : t.bucket_end[i - 1], b1 = t.bucket_end[i];
        e = cudaMemcpyAsync(flat_out + b0, t.flat_grad + b0, sizeof(float) * (size_t)(b1 - b0), cudaMemcpyDeviceToDevice, side);
        if (e != cudaSuccess) { t.err = std::string("bucket copy: ") + cudaGetErrorString(e); return -100 - (int)e; }
    }
    return 0;
}

}  // namespace train

// ------------------------------------------------------------------------------------------------ C ABI
struct ndnet_b200_trainer { train::Trainer t; };

extern "C" int ndnet_b200_trainer_create(int device, int n_tensors, const char *const *names, const int64_t *const *shapes,
                                         const int *ndims, ndnet_b200_trainer **out) {
    if (!out || n_tensors <= 0 || !names || !shapes || !ndims) return -200;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        fprintf(stderr, "ndnet_b200: trainer: no CUDA device %d; this library has no CPU path\n", device);
        return -201;
    }
    ndnet_b200_trainer *h = new (std::nothrow) ndnet_b200_trainer();
    if (!h) return -202;
    train::Trainer &t = h->t;
    t.device = device; t.n_tensors = n_tensors;
    for (int i = 0; i < n_tensors; i++) {
        t.index[names[i]] = i;
        t.shapes.emplace_back(shapes[i], shapes[i] + ndims[i]);
    }
    auto fail = [&](const char *why) {
        fprintf(stderr, "ndnet_b200_trainer_create: %s\n", t.err.empty() ? why : t.err.c_str());
        delete h;
        return -205;
    };
    // which of the reference's four networks this state_dict is (ndtnet.py:166-243, pointnet.py:137-214): a segmentation
    // head has conv4 and BatchNorm layers of its own; the NDT trunk takes 12-wide rows behind a 3x3 input transform, the
    // PointNet trunk point_dim-wide rows behind a point_dim x point_dim one
    auto it = t.index.find("feature_extractor.conv3.weight");
    auto i1 = t.index.find("feature_extractor.conv1.weight");
    auto it1 = t.index.find("feature_extractor.t1.conv1.weight");
    if (it == t.index.end() || i1 == t.index.end() || it1 == t.index.end()) return fail("not an NDTNet / PointNet state_dict");
    const bool seg = t.index.count("conv4.weight") != 0;
    auto ic = t.index.find(seg ? "conv4.weight" : "conv3.weight");
    if (ic == t.index.end()) return fail("no output layer (conv3 / conv4) in the state_dict");
    t.F = (int)t.shapes[it->second][0];
    t.C = (int)t.shapes[ic->second][0];
    t.D = (int)t.shapes[i1->second][1];
    const int d1 = (int)t.shapes[it1->second][1];
    const bool ndt = t.D == 12 && d1 == 3;
    if (!ndt && d1 != t.D) return fail("input transform and first layer disagree on the point width");
    if (t.D < 1 || t.D > 64) return fail("unsupported point width");
    t.kind = (ndt ? 0 : 2) + (seg ? 1 : 0);
    using namespace train;
    const std::string fe = "feature_extractor";
    bool ok = bind_tnet(t, t.t1, fe + ".t1", d1) && bind_tnet(t, t.t2, fe + ".t2", 64) &&
              bind_block(t, t.c1, fe + ".conv1", fe + ".bn1", t.D, 64, false) && bind_block(t, t.c2, fe + ".conv2", fe + ".bn2", 64, 128, false) &&
              bind_block(t, t.c3, fe + ".conv3", fe + ".bn3", 128, t.F, false);
    if (seg)
        ok = ok && bind_block(t, t.h1, "conv1", "bn1", 64 + t.F, 512, true) && bind_block(t, t.h2, "conv2", "bn2", 512, 256, true) &&
             bind_block(t, t.h3, "conv3", "bn3", 256, 128, true) && bind_block(t, t.h4, "conv4", "", 128, t.C, false);
    else
        ok = ok && bind_block(t, t.h1, "conv1", "", t.F, 512, true) && bind_block(t, t.h2, "conv2", "", 512, 256, true) &&
             bind_block(t, t.h3, "conv3", "", 256, t.C, false);
    if (!ok) return fail("tensor binding failed");
    {   // flat gradient layout: parameters (weights, biases, BatchNorm scale/shift) in the order backward finishes them,
        // 16-byte aligned; three buckets (head | trunk conv2-3 + feature T-Net | conv1 + input T-Net)
        t.grad_off.assign(n_tensors, -1);
        std::vector<Block *> order;
        if (seg) order.push_back(&t.h4);
        for (Block *b : {&t.h3, &t.h2, &t.h1}) order.push_back(b);
        const int head_last = (int)order.size() - 1;
        for (Block *b : {&t.c3, &t.c2, &t.t2.f3, &t.t2.f2, &t.t2.f1, &t.t2.c3, &t.t2.c2, &t.t2.c1,
                         &t.c1, &t.t1.f3, &t.t1.f2, &t.t1.f1, &t.t1.c3, &t.t1.c2, &t.t1.c1}) order.push_back(b);
        const int bucket_last[Trainer::kBuckets] = {head_last, head_last + 8, head_last + 15};
        long off = 0;
        int bucket = 0;
        auto place = [&](int i) {
            if (i < 0 || t.grad_off[i] >= 0) return;
            long numel = 1;
            for (auto v : t.shapes[i]) numel *= v;
            t.grad_off[i] = off;
            off += (numel + 3) / 4 * 4;
        };
        for (int k = 0; k < (int)order.size(); k++) {
            Block *b = order[k];
            place(b->lin.w); place(b->lin.b);
            if (b->has_bn) { place(b->bn.g); place(b->bn.be); }
            if (bucket < Trainer::kBuckets && k == bucket_last[bucket]) t.bucket_end[bucket++] = off;
        }
        t.flat_elems = off;
    }
    *out = h;
    return 0;
}

extern "C" int ndnet_b200_trainer_info(const ndnet_b200_trainer *h, int *kind, int *point_width, int *outputs) {
    if (!h) return -200;
    if (kind) *kind = h->t.kind;
    if (point_width) *point_width = h->t.D;
    if (outputs) *outputs = h->t.C;
    return 0;
}

extern "C" int ndnet_b200_trainer_backward_flat(ndnet_b200_trainer *h, const float *dlogp, float *const *tensors, float *flat_grads,
                                                void *stream) {
    if (!h || !dlogp || !tensors || !flat_grads) return -200;
    cudaError_t e = cudaSetDevice(h->t.device);
    if (e != cudaSuccess) return -100 - (int)e;
    return train::backward_flat(h->t, dlogp, tensors, flat_grads, (cudaStream_t)stream);
}

extern "C" int ndnet_b200_trainer_num_buckets(const ndnet_b200_trainer *h) { return h ? train::Trainer::kBuckets : -200; }

extern "C" int ndnet_b200_trainer_bucket_range(const ndnet_b200_trainer *h, int i, long *begin, long *end) {
    if (!h || i < 0 || i >= train::Trainer::kBuckets || !begin || !end) return -200;
    *begin = i == 0 ? 0 : h->t.bucket_end[i - 1];
    *end = h->t.bucket_end[i];
    return 0;
}

extern "C" int ndnet_b200_trainer_set_deferred_copy(ndnet_b200_trainer *h, int enable) {
    if (!h || (enable != 0 && enable != 1)) return -200;
    h->t.defer_copy = enable;
    return 0;
}

extern "C" int ndnet_b200_trainer_bucket_ready(ndnet_b200_trainer *h, int i, float *flat_grads, void *side_stream) {
    if (!h || !flat_grads) return -200;
    cudaError_t e = cudaSetDevice(h->t.device);
    if (e != cudaSuccess) return -100 - (int)e;
    return train::bucket_ready(h->t, i, flat_grads, (cudaStream_t)side_stream);
}

extern "C" long ndnet_b200_trainer_grad_layout(const ndnet_b200_trainer *h, long *offsets, int n) {
    if (!h) return -200;
    for (int i = 0; offsets && i < n && i < h->t.n_tensors; i++) offsets[i] = h->t.grad_off[i];
    return h->t.flat_elems;
}

extern "C" int ndnet_b200_trainer_set_graph(ndnet_b200_trainer *h, int enable) {
    if (!h || (enable != 0 && enable != 1)) return -200;
    h->t.use_graph = enable;
    return 0;
}

extern "C" int ndnet_b200_trainer_set_precision(ndnet_b200_trainer *h, int tf32) {
    if (!h || (tf32 != 0 && tf32 != 1)) return -200;
    h->t.tf32 = tf32;
    return 0;
}

// Test hook: C[M,N] (+)= A[M,K] . B[N,K]^T (+ bias) through the training GEMMs.  mode 0 = fp32 FMA kernel, 1 = tcgen05 TF32
// kernel (-206 when the shape is not eligible for it).
extern "C" int ndnet_b200_debug_train_gemm(int mode, const float *A, long lda, const float *B, long ldb, float *C, long ldc, int M,
                                           int N, int K, const float *bias, int accumulate, void *stream) {
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return -200;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1) {
        if (!train::gemm_tf32(st, A, lda, B, ldb, C, ldc, M, N, K, bias, accumulate != 0)) return -206;
    } else {
        train::gemm(st, A, lda, 1, B, ldb, 1, C, ldc, M, N, K, bias, accumulate != 0);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -100 - (int)e;
}

extern "C" const char *ndnet_b200_trainer_last_error(const ndnet_b200_trainer *h) { return h ? h->t.err.c_str() : "null trainer"; }

extern "C" int ndnet_b200_trainer_forward(ndnet_b200_trainer *h, const float *feat, int B, int N, float *const *tensors, float *out_logp,
                                          int update_running_stats, void *stream) {
    if (!h || !feat || !tensors || !out_logp || B < 2 || N < 1) return -200;      // BatchNorm needs more than one row per channel
    cudaError_t e = cudaSetDevice(h->t.device);
    if (e != cudaSuccess) return -100 - (int)e;
    return train::forward_entry(h->t, feat, B, N, tensors, out_logp, update_running_stats, (cudaStream_t)stream);
}

extern "C" int ndnet_b200_trainer_backward(ndnet_b200_trainer *h, const float *dlogp, float *const *tensors, float *const *grads,
                                           void *stream) {
    if (!h || !dlogp || !tensors || !grads) return -200;
    cudaError_t e = cudaSetDevice(h->t.device);
    if (e != cudaSuccess) return -100 - (int)e;
    return train::backward(h->t, dlogp, tensors, grads, (cudaStream_t)stream);
}

// Test hook: copies one internal activation / gradient buffer ("h3.dA", "t2.c1.Y", "c1.A", "t1.T", ...) to `out` (device).
// Returns the element count, or -200 for an unknown name.
extern "C" long ndnet_b200_trainer_debug_buffer(ndnet_b200_trainer *h, const char *name, float *out, void *stream) {
    if (!h || !name || !h->t.arena) return -200;
    train::Trainer &t = h->t;
    std::map<std::string, train::Block *> blocks = {
        {"t1.c1", &t.t1.c1}, {"t1.c2", &t.t1.c2}, {"t1.c3", &t.t1.c3}, {"t1.f1", &t.t1.f1}, {"t1.f2", &t.t1.f2}, {"t1.f3", &t.t1.f3},
        {"t2.c1", &t.t2.c1}, {"t2.c2", &t.t2.c2}, {"t2.c3", &t.t2.c3}, {"t2.f1", &t.t2.f1}, {"t2.f2", &t.t2.f2}, {"t2.f3", &t.t2.f3},
        {"c1", &t.c1}, {"c2", &t.c2}, {"c3", &t.c3}, {"h1", &t.h1}, {"h2", &t.h2}, {"h3", &t.h3}, {"h4", &t.h4}};
    const std::string n(name);
    const float *src = nullptr;
    long count = 0;
    const long M = (long)t.B * t.N;
    if (n == "t1.T") { src = t.t1.T; count = (long)t.B * t.t1.d * t.t1.d; }
    else if (n == "t2.T") { src = t.t2.T; count = (long)t.B * 4096; }
    else if (n == "t1.dT") { src = t.t1.dT; count = (long)t.B * t.t1.d * t.t1.d; }
    else if (n == "t2.dT") { src = t.t2.dT; count = (long)t.B * 4096; }
    else if (n == "dH0") { src = t.dH0; count = M * (64 + t.F); }
    else if (n == "dX12") { src = t.dX12; count = M * t.D; }
    else if (n == "X12") { src = t.X12; count = M * t.D; }
    else if (n == "X2") { src = t.X2; count = M * 64; }
    else if (n == "H0") { src = t.H0; count = M * (64 + t.F); }
    else if (n == "Gf") { src = t.Gf; count = (long)t.B * t.F; }
    else if (n == "dGf") { src = t.dGf; count = (long)t.B * t.F; }
    else if (n == "logp") { src = t.logp; count = t.out_elems(); }
    else if (n == "t1.G") { src = t.t1.G; count = (long)t.B * 1024; }
    else if (n == "t2.G") { src = t.t2.G; count = (long)t.B * 1024; }
    else if (n == "t1.dG") { src = t.t1.dG; count = (long)t.B * 1024; }
    else if (n == "t2.dG") { src = t.t2.dG; count = (long)t.B * 1024; }
    else {
        const size_t dot = n.rfind('.');
        if (dot == std::string::npos) return -200;
        auto it = blocks.find(n.substr(0, dot));
        if (it == blocks.end()) return -200;
        train::Block &k = *it->second;
        const std::string f = n.substr(dot + 1);
        count = k.rows * k.lin.out;
        if (f == "Y") src = k.Y; else if (f == "A") src = k.A; else if (f == "dA") src = k.dA;
        else if (f == "mean") { src = k.mean; count = k.lin.out; } else if (f == "rstd") { src = k.rstd; count = k.lin.out; }
        else return -200;
    }
    if (out) cudaMemcpyAsync(out, src, sizeof(float) * count, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return count;
}

extern "C" void ndnet_b200_trainer_destroy(ndnet_b200_trainer *h) {
    if (!h) return;
    cudaSetDevice(h->t.device);
    if (h->t.gf.exec) cudaGraphExecDestroy(h->t.gf.exec);
    if (h->t.gb.exec) cudaGraphExecDestroy(h->t.gb.exec);
    if (h->t.cap) cudaStreamDestroy(h->t.cap);
    for (cudaEvent_t e : h->t.bucket_ev) if (e) cudaEventDestroy(e);
    if (h->t.arena) cudaFree(h->t.arena);
    delete h;
}
