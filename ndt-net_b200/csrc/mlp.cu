// mlp.cu — NDT-Net / PointNet forward on B200 (eval mode, BatchNorm folded), reference:
// /root/reference/ndnet/models/ndtnet.py:33-62 (TNet), :112-164 (NDTNet), :181-196 (classification head),
// :218-243 (segmentation head).
//
// Dense contractions with K >= 64 run as tcgen05/TMA bf16 GEMMs (mlp_gemm.cuh) with fused
// bias/BN/ReLU epilogues; the three 128 -> 1024/F layers fold the global max-pool into the epilogue, so
// the [B, F, N] activation of ndtnet.py:47,161 is never materialised.  The segmentation head's first
// layer is split algebraically: conv1([x_t2 ; g]) = W[:, :64] x_t2 + (W[:, 64:] g) where g is the
// per-cloud max-pooled vector (ndtnet.py:224-230), so the 1024-wide part is one GEMV per cloud.
// K = 3 / 12 layers and the per-cloud FC stacks (M = B rows) run on CUDA cores in fp32.
#include "mlp_host.h"
#include "mlp_gemm.cuh"

namespace ndt { void count_launches(long n); }

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace mlp {

// ------------------------------------------------------------------------------------------------
// small CUDA-core kernels
// ------------------------------------------------------------------------------------------------

// T-Net(3) first layer: relu(bn1(conv1(p))) on the 3 mean coordinates (ndtnet.py:45).  Thread per point.
__global__ void __launch_bounds__(128) k_tnet3_l1(const float *__restrict__ feat, int total, const float *__restrict__ W /*[64][3]*/,
                                                  const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out /*[total][64]*/) {
    __shared__ float sW[64 * 3], sb[64];
    for (int i = threadIdx.x; i < 192; i += blockDim.x) sW[i] = W[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const float x = feat[(size_t)p * 12 + 0], y = feat[(size_t)p * 12 + 1], z = feat[(size_t)p * 12 + 2];
    uint4 *o = reinterpret_cast<uint4 *>(out + (size_t)p * 64);
#pragma unroll
    for (int c8 = 0; c8 < 8; c8++) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c8 * 8 + j * 2;
            const float a = fmaxf(sW[c * 3] * x + sW[c * 3 + 1] * y + sW[c * 3 + 2] * z + sb[c], 0.f);
            const float b = fmaxf(sW[c * 3 + 3] * x + sW[c * 3 + 4] * y + sW[c * 3 + 5] * z + sb[c + 1], 0.f);
            __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            pk[j] = *reinterpret_cast<uint32_t *>(&h);
        }
        o[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// T-Net(d) first layer for a generic input width d <= 16 (PointNet with point_dim != 3, pointnet.py:45).
__global__ void __launch_bounds__(128) k_tnet_l1(const float *__restrict__ x, int d, int total, const float *__restrict__ W /*[64][d]*/,
                                                 const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out /*[total][64]*/) {
    __shared__ float sW[64 * 16], sb[64];
    for (int i = threadIdx.x; i < 64 * d; i += blockDim.x) sW[i] = W[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = k < d ? x[(size_t)p * d + k] : 0.f;
    uint4 *o = reinterpret_cast<uint4 *>(out + (size_t)p * 64);
#pragma unroll 1
    for (int c8 = 0; c8 < 8; c8++) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c8 * 8 + j * 2;
            float a = sb[c], b = sb[c + 1];
#pragma unroll
            for (int k = 0; k < 16; k++) if (k < d) { a += sW[c * d + k] * v[k]; b += sW[(c + 1) * d + k] * v[k]; }
            __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(a, 0.f), fmaxf(b, 0.f));
            pk[j] = *reinterpret_cast<uint32_t *>(&h);
        }
        o[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// PointNet input transform + first trunk layer (pointnet.py:111-120): x' = T x, nan_to_num(nan=0), bn1(conv1(x')).
__global__ void __launch_bounds__(128) k_trunk_l1_pn(const float *__restrict__ x, int d, int P, int total, const float *__restrict__ T1 /*[B][d*d]*/,
                                                     const float *__restrict__ W /*[64][d]*/, const float *__restrict__ bias,
                                                     __nv_bfloat16 *__restrict__ out /*[total][64]*/) {
    __shared__ float sW[64 * 16], sb[64];
    for (int i = threadIdx.x; i < 64 * d; i += blockDim.x) sW[i] = W[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const float *t = T1 + (size_t)(p / P) * d * d;
    float v[16], y[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = k < d ? x[(size_t)p * d + k] : 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float a = 0.f;
        if (i < d) {
#pragma unroll
            for (int k = 0; k < 16; k++) if (k < d) a += t[i * d + k] * v[k];
            if (isnan(a)) a = 0.f;                                   // torch.nan_to_num(x, nan=0.0): inf -> +-FLT_MAX
            else if (isinf(a)) a = a > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
        }
        y[i] = a;
    }
    uint4 *o = reinterpret_cast<uint4 *>(out + (size_t)p * 64);
#pragma unroll 1
    for (int c8 = 0; c8 < 8; c8++) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c8 * 8 + j * 2;
            float a = sb[c], b = sb[c + 1];
#pragma unroll
            for (int k = 0; k < 16; k++) if (k < d) { a += sW[c * d + k] * y[k]; b += sW[(c + 1) * d + k] * y[k]; }
            __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            pk[j] = *reinterpret_cast<uint32_t *>(&h);
        }
        o[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// Input transform + first trunk layer (ndtnet.py:132-149): p' = T1 p, Sigma' = T1 Sigma, bn1(conv1([p';Sigma'])).
__global__ void __launch_bounds__(128) k_trunk_l1(const float *__restrict__ feat, int P, int total, const float *__restrict__ T1 /*[B][9]*/,
                                                  const float *__restrict__ W /*[64][12]*/, const float *__restrict__ bias,
                                                  __nv_bfloat16 *__restrict__ out /*[total][64]*/) {
    __shared__ float sW[64 * 12], sb[64];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) sW[i] = W[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const float *t = T1 + (size_t)(p / P) * 9;
    float f[12], x[12];
#pragma unroll
    for (int i = 0; i < 12; i++) f[i] = feat[(size_t)p * 12 + i];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        x[i] = t[i * 3] * f[0] + t[i * 3 + 1] * f[1] + t[i * 3 + 2] * f[2];
#pragma unroll
        for (int k = 0; k < 3; k++) x[3 + i * 3 + k] = t[i * 3] * f[3 + k] + t[i * 3 + 1] * f[6 + k] + t[i * 3 + 2] * f[9 + k];
    }
    uint4 *o = reinterpret_cast<uint4 *>(out + (size_t)p * 64);
#pragma unroll 1
    for (int c8 = 0; c8 < 8; c8++) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float a = sb[c8 * 8 + j * 2], b = sb[c8 * 8 + j * 2 + 1];
#pragma unroll
            for (int k = 0; k < 12; k++) { a += sW[(c8 * 8 + j * 2) * 12 + k] * x[k]; b += sW[(c8 * 8 + j * 2 + 1) * 12 + k] * x[k]; }
            __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            pk[j] = *reinterpret_cast<uint32_t *>(&h);
        }
        o[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// y[b][o] = act(W[o] . x[b] + bias[o]) for B rows (the per-cloud FC stacks, ndtnet.py:54-60,189-191): an fp32 SIMT GEMM.
// CTA tile = 32 rows (clouds) x 64 outputs, k-steps of 32, operands of the next step prefetched into registers.  The weights
// are stored transposed ([in][out_pad]), so a warp's 16-byte loads cover whole 128-byte lines of four input channels and go
// to shared memory as they are; the x tile is transposed on the way in (padded rows: conflict-free).  A thread multiplies its
// 2 rows by its 4 outputs: 11 instructions per 8 FMAs.  grid (ceil(out/64), ceil(B/32)), block 256 = 16 x 16.
constexpr int kFcTM = 32, kFcTN = 64, kFcBK = 32;

__global__ void __launch_bounds__(256) k_fc(const float *__restrict__ Wt /*[in][ldw]*/, int ldw, const float *__restrict__ bias,
                                            const void *__restrict__ xin, int ldx, int decode, float *__restrict__ y, int ldy, int B, int in,
                                            int out, int relu, int identity_dim, __nv_bfloat16 *__restrict__ y_t /*[B][dim][dim] transposed bf16*/) {
    __shared__ float sx[kFcBK][kFcTM + 1];
    __shared__ __align__(16) float sw[kFcBK][kFcTN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int o0 = blockIdx.x * kFcTN, r0 = blockIdx.y * kFcTM;
    float acc[2][4] = {};
    const int xr = tid >> 3, xk = (tid & 7) * 4;             // x tile: row tid / 8, four consecutive k (lanes run along k: coalesced)
    const int wk = tid >> 4, wo = (tid & 15) * 4;            // w tile: input channel tid / 16 (+16), four consecutive outputs
    const bool vec = (in & 3) == 0 && (ldx & 3) == 0;
    float xv[4], wv[2][4];
    auto fetch = [&](int k0) {
        {
            const int r = r0 + xr, k = k0 + xk;
#pragma unroll
            for (int j = 0; j < 4; j++) xv[j] = 0.f;
            if (r < B) {
                if (vec && k + 3 < in) {
                    const uint4 q = *reinterpret_cast<const uint4 *>((const unsigned *)xin + (size_t)r * ldx + k);
                    xv[0] = decode ? dec_f32(q.x) : __uint_as_float(q.x); xv[1] = decode ? dec_f32(q.y) : __uint_as_float(q.y);
                    xv[2] = decode ? dec_f32(q.z) : __uint_as_float(q.z); xv[3] = decode ? dec_f32(q.w) : __uint_as_float(q.w);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (k + j < in) xv[j] = decode ? dec_f32(((const unsigned *)xin)[(size_t)r * ldx + k + j]) : ((const float *)xin)[(size_t)r * ldx + k + j];
                }
            }
        }
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const int k = k0 + wk + pass * 16, o = o0 + wo;
#pragma unroll
            for (int j = 0; j < 4; j++) wv[pass][j] = 0.f;
            if (k < in && o < ldw) {                          // ldw is a multiple of 4 and the padding columns hold zeros
                const float4 q = *reinterpret_cast<const float4 *>(Wt + (size_t)k * ldw + o);
                wv[pass][0] = q.x; wv[pass][1] = q.y; wv[pass][2] = q.z; wv[pass][3] = q.w;
            }
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < in; k0 += kFcBK) {
#pragma unroll
        for (int j = 0; j < 4; j++) sx[xk + j][xr] = xv[j];
#pragma unroll
        for (int pass = 0; pass < 2; pass++)
            *reinterpret_cast<float4 *>(&sw[wk + pass * 16][wo]) = make_float4(wv[pass][0], wv[pass][1], wv[pass][2], wv[pass][3]);
        __syncthreads();
        if (k0 + kFcBK < in) fetch(k0 + kFcBK);          // the next step's operands travel while this one is multiplied
#pragma unroll
        for (int k = 0; k < kFcBK; k++) {
            const float a0 = sx[k][2 * ty], a1 = sx[k][2 * ty + 1];
            const float4 w = *reinterpret_cast<const float4 *>(&sw[k][4 * tx]);
            acc[0][0] += a0 * w.x; acc[0][1] += a0 * w.y; acc[0][2] += a0 * w.z; acc[0][3] += a0 * w.w;
            acc[1][0] += a1 * w.x; acc[1][1] += a1 * w.y; acc[1][2] += a1 * w.z; acc[1][3] += a1 * w.w;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int row = r0 + 2 * ty + i;
        if (row >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int o = o0 + 4 * tx + j;
            if (o >= out) continue;
            float v = acc[i][j] + (bias ? bias[o] : 0.f);
            if (identity_dim > 0 && (o / identity_dim) == (o % identity_dim)) v += 1.f;   // + eye (ndtnet.py:59)
            if (relu) v = fmaxf(v, 0.f);
            y[(size_t)row * ldy + o] = v;
            if (y_t) {
                const int ii = o / identity_dim, jj = o % identity_dim;
                y_t[(size_t)row * identity_dim * identity_dim + (size_t)jj * identity_dim + ii] = __float2bfloat16(v);
            }
        }
    }
}

// softmax over the class dimension (ndtnet.py:194).  One block per cloud.
__global__ void __launch_bounds__(256) k_softmax(const float *__restrict__ x, float *__restrict__ y, int n) {
    __shared__ float red[8];
    const float *xi = x + (size_t)blockIdx.x * n;
    float *yo = y + (size_t)blockIdx.x * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, xi[i]);
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int k = 1; k < 8; k++) mx = fmaxf(mx, red[k]);
    __syncthreads();
    float sum = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sum += expf(xi[i] - mx);
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int k = 0; k < 8; k++) sum += red[k];
    for (int i = threadIdx.x; i < n; i += blockDim.x) yo[i] = expf(xi[i] - mx) / sum;
}

// Inspection taps (Model::tap): an internal activation of the last forward as fp32.  src_kind 0 = fp32, 1 = bf16,
// 2 = order-preserving encoded fp32 (the max-pool accumulators).
__global__ void __launch_bounds__(256) k_tap(const void *__restrict__ src, int src_kind, long n, float *__restrict__ dst) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float v;
        if (src_kind == 0) v = ((const float *)src)[i];
        else if (src_kind == 1) v = __bfloat162float(((const __nv_bfloat16 *)src)[i]);
        else v = dec_f32(((const unsigned *)src)[i]);
        dst[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
cudaError_t Scratch::reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (buf) cudaFree(buf);
    buf = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&buf, need);
    if (e == cudaSuccess) bytes = need;
    return e;
}
void Scratch::release() { if (buf) cudaFree(buf); buf = nullptr; bytes = 0; }

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

// bf16 [batch][rows][K] row-major, box = 64 (K) x box_rows x 1, 128B swizzle
static bool make_map(CUtensorMap *m, const void *ptr, int K, int rows, int batch, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows * K * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct DevBuf {
    void *p = nullptr;
    template <typename T> T *as() const { return (T *)p; }
};

struct TNetW {
    int in = 0;
    float *l1_w = nullptr, *l1_b = nullptr;          // in == 3: fp32 [64][3]
    __nv_bfloat16 *l1_wb = nullptr;                  // in == 64: bf16 [64][64]
    __nv_bfloat16 *l2_w = nullptr, *l3_w = nullptr;  // [128][64], [1024][128]
    float *l2_b = nullptr, *l3_b = nullptr;
    float *fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr, *fc3_w = nullptr, *fc3_b = nullptr;
};

struct ModelImpl {
    int kind = 0, F = 0, ncls = 0;   // ncls = number of outputs of the last layer
    bool pointnet = false;           // kinds 2/3: ndnet/models/pointnet.py (no covariance branch, generic point_dim)
    int in_dim = 12;                 // floats per input row
    TNetW t1, t2;
    float *c1_w = nullptr, *c1_b = nullptr;                       // [64][12]
    __nv_bfloat16 *c2_w = nullptr, *c3_w = nullptr; float *c2_b = nullptr, *c3_b = nullptr;
    // segmentation head
    __nv_bfloat16 *h1a_w = nullptr, *h2_w = nullptr, *h3_w = nullptr, *h4_w = nullptr;
    float *h1g_w = nullptr, *h1_b = nullptr, *h2_b = nullptr, *h3_b = nullptr, *h4_b = nullptr;
    // classification head
    float *k1_w = nullptr, *k1_b = nullptr, *k2_w = nullptr, *k2_b = nullptr, *k3_w = nullptr, *k3_b = nullptr;
    std::vector<void *> allocs;
    bool fuse_head = true;           // segmentation head layers 1 + 2 in one kernel (k_head12); off: two GEMMs, "head.l1" can be tapped
    // where the activations of the last forward live (Model::tap)
    struct Tap { const void *ptr; int kind; long count; };
    std::map<std::string, Tap> taps;
};

namespace {

struct HostTensor { const float *data; std::vector<int64_t> shape; size_t numel() const { size_t n = 1; for (auto s : shape) n *= (size_t)s; return n; } };
typedef std::map<std::string, HostTensor> TensorMapHost;

struct Folded { std::vector<float> w, b; int out = 0, in = 0; };

// conv/linear `name` (+ optional batch-norm `bn`) -> folded fp32 weight [out][in] and bias [out]
bool fold(const TensorMapHost &t, const std::string &name, const std::string &bn, Folded &f, std::string &err) {
    auto w = t.find(name + ".weight"), b = t.find(name + ".bias");
    if (w == t.end() || b == t.end()) { err = "missing tensor " + name + ".weight/.bias"; return false; }
    f.out = (int)w->second.shape[0];
    f.in = (int)(w->second.numel() / (size_t)f.out);
    f.w.assign(w->second.data, w->second.data + w->second.numel());
    f.b.assign(b->second.data, b->second.data + f.out);
    if (!bn.empty()) {
        auto g = t.find(bn + ".weight"), be = t.find(bn + ".bias"), mu = t.find(bn + ".running_mean"), var = t.find(bn + ".running_var");
        if (g == t.end() || be == t.end() || mu == t.end() || var == t.end()) { err = "missing batch-norm tensors for " + bn; return false; }
        for (int o = 0; o < f.out; o++) {
            const float s = g->second.data[o] / std::sqrt(var->second.data[o] + 1e-5f);
            for (int k = 0; k < f.in; k++) f.w[(size_t)o * f.in + k] *= s;
            f.b[o] = (f.b[o] - mu->second.data[o]) * s + be->second.data[o];
        }
    }
    return true;
}

uint16_t f2bf(float f) {   // round-to-nearest-even
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

bool up_f32(ModelImpl &m, const std::vector<float> &h, float *&d) {
    if (cudaMalloc((void **)&d, h.size() * 4 + 16) != cudaSuccess) return false;
    m.allocs.push_back(d);
    return cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
}
// fp32 weight [out][in] uploaded TRANSPOSED ([in][out_pad], out_pad = out rounded up to 4): k_fc reads four consecutive
// outputs of one input channel with a single 16-byte load, coalesced over the outputs
bool up_f32_t(ModelImpl &m, const Folded &f, float *&d) {
    const int op = (f.out + 3) / 4 * 4;
    std::vector<float> t((size_t)f.in * op, 0.f);
    for (int o = 0; o < f.out; o++)
        for (int k = 0; k < f.in; k++) t[(size_t)k * op + o] = f.w[(size_t)o * f.in + k];
    return up_f32(m, t, d);
}
bool up_bf16(ModelImpl &m, const float *h, size_t n, __nv_bfloat16 *&d) {
    std::vector<uint16_t> tmp(n);
    for (size_t i = 0; i < n; i++) tmp[i] = f2bf(h[i]);
    if (cudaMalloc((void **)&d, n * 2 + 16) != cudaSuccess) return false;
    m.allocs.push_back(d);
    return cudaMemcpy(d, tmp.data(), n * 2, cudaMemcpyHostToDevice) == cudaSuccess;
}

bool build_tnet(ModelImpl &m, const TensorMapHost &t, const std::string &p, int in, TNetW &w, std::string &err) {
    Folded a, b, c, f1, f2, f3;
    if (!fold(t, p + ".conv1", p + ".bn1", a, err) || !fold(t, p + ".conv2", p + ".bn2", b, err) || !fold(t, p + ".conv3", p + ".bn3", c, err) ||
        !fold(t, p + ".fc1", p + ".bn4", f1, err) || !fold(t, p + ".fc2", p + ".bn5", f2, err) || !fold(t, p + ".fc3", "", f3, err)) return false;
    if (a.in != in || a.out != 64 || b.out != 128 || b.in != 64 || c.out != 1024 || c.in != 128 || f1.in != 1024 || f1.out != 512 ||
        f2.out != 256 || f3.out != in * in) { err = "unexpected T-Net shapes under " + p; return false; }
    w.in = in;
    bool ok = true;
    if (in != 64) ok = ok && up_f32(m, a.w, w.l1_w); else ok = ok && up_bf16(m, a.w.data(), a.w.size(), w.l1_wb);
    ok = ok && up_f32(m, a.b, w.l1_b) && up_bf16(m, b.w.data(), b.w.size(), w.l2_w) && up_f32(m, b.b, w.l2_b) &&
         up_bf16(m, c.w.data(), c.w.size(), w.l3_w) && up_f32(m, c.b, w.l3_b) && up_f32_t(m, f1, w.fc1_w) && up_f32(m, f1.b, w.fc1_b) &&
         up_f32_t(m, f2, w.fc2_w) && up_f32(m, f2.b, w.fc2_b) && up_f32_t(m, f3, w.fc3_w) && up_f32(m, f3.b, w.fc3_b);
    if (!ok) err = "device upload failed";
    return ok;
}

}  // namespace

int Model::build(int kind, int n_tensors, const char *const *names, const float *const *data, const int64_t *const *shapes,
                 const int *ndims, std::string &err) {
    TensorMapHost t;
    for (int i = 0; i < n_tensors; i++) {
        HostTensor h; h.data = data[i];
        for (int d = 0; d < ndims[i]; d++) h.shape.push_back(shapes[i][d]);
        t[names[i]] = h;
    }
    impl = new ModelImpl();
    ModelImpl &m = *impl;
    m.pointnet = kind >= 2;
    m.kind = kind & 1;               // 0 classification head, 1 segmentation head
    const std::string fe = "feature_extractor";
    Folded c1, c2, c3;
    if (!fold(t, fe + ".conv1", fe + ".bn1", c1, err) || !fold(t, fe + ".conv2", fe + ".bn2", c2, err) || !fold(t, fe + ".conv3", fe + ".bn3", c3, err)) return -301;
    m.in_dim = c1.in;
    const int t1_dim = m.pointnet ? c1.in : 3;
    if (m.pointnet ? (c1.in < 1 || c1.in > 16) : (c1.in != 12)) { err = "unsupported input width (NDT-Net: 3 + 9 covariances; PointNet: point_dim <= 16)"; return -302; }
    if (c1.out != 64 || c2.in != 64 || c2.out != 128 || c3.in != 128) { err = "unexpected trunk shapes"; return -302; }
    if (!build_tnet(m, t, fe + ".t1", t1_dim, m.t1, err) || !build_tnet(m, t, fe + ".t2", 64, m.t2, err)) return -301;
    m.F = c3.out;
    bool ok = up_f32(m, c1.w, m.c1_w) && up_f32(m, c1.b, m.c1_b) && up_bf16(m, c2.w.data(), c2.w.size(), m.c2_w) && up_f32(m, c2.b, m.c2_b) &&
              up_bf16(m, c3.w.data(), c3.w.size(), m.c3_w) && up_f32(m, c3.b, m.c3_b);
    if (m.kind == 1) {
        Folded h1, h2, h3, h4;
        if (!fold(t, "conv1", "bn1", h1, err) || !fold(t, "conv2", "bn2", h2, err) || !fold(t, "conv3", "bn3", h3, err) || !fold(t, "conv4", "", h4, err)) return -301;
        if (h1.in != m.F + 64 || h1.out != 512 || h2.in != 512 || h2.out != 256 || h3.in != 256 || h3.out != 128 || h4.in != 128 || h4.out > 32) {
            err = "unexpected segmentation head shapes (at most 32 output classes supported)"; return -302;
        }
        m.ncls = h4.out;
        std::vector<float> wa((size_t)512 * 64), wg((size_t)512 * m.F);
        for (int o = 0; o < 512; o++) {
            for (int k = 0; k < 64; k++) wa[(size_t)o * 64 + k] = h1.w[(size_t)o * h1.in + k];              // x_t2 part (cat order ndtnet.py:230)
            for (int k = 0; k < m.F; k++) wg[(size_t)o * m.F + k] = h1.w[(size_t)o * h1.in + 64 + k];       // global-feature part
        }
        Folded fg; fg.out = 512; fg.in = m.F; fg.w = wg;
        ok = ok && up_bf16(m, wa.data(), wa.size(), m.h1a_w) && up_f32_t(m, fg, m.h1g_w) && up_f32(m, h1.b, m.h1_b) &&
             up_bf16(m, h2.w.data(), h2.w.size(), m.h2_w) && up_f32(m, h2.b, m.h2_b) && up_bf16(m, h3.w.data(), h3.w.size(), m.h3_w) &&
             up_f32(m, h3.b, m.h3_b) && up_bf16(m, h4.w.data(), h4.w.size(), m.h4_w) && up_f32(m, h4.b, m.h4_b);
    } else {
        Folded k1, k2, k3;
        if (!fold(t, "conv1", "", k1, err) || !fold(t, "conv2", "", k2, err) || !fold(t, "conv3", "", k3, err)) return -301;
        if (k1.in != m.F || k1.out != 512 || k2.in != 512 || k2.out != 256 || k3.in != 256) { err = "unexpected classification head shapes"; return -302; }
        m.ncls = k3.out;
        ok = ok && up_f32_t(m, k1, m.k1_w) && up_f32(m, k1.b, m.k1_b) && up_f32_t(m, k2, m.k2_w) && up_f32(m, k2.b, m.k2_b) &&
             up_f32_t(m, k3, m.k3_w) && up_f32(m, k3.b, m.k3_b);
    }
    if (!ok) { err = "device upload failed"; return -303; }
    return 0;
}

int Model::input_dim() const { return impl ? impl->in_dim : 0; }
void Model::set_fused_head(bool on) { if (impl) impl->fuse_head = on; }

long Model::tap(const char *name, float *out, long cap, cudaStream_t st) {
    if (!impl || !name) return -300;
    auto it = impl->taps.find(name);
    if (it == impl->taps.end()) return -307;
    const ModelImpl::Tap &t = it->second;
    if (!out) return t.count;
    if (cap < t.count) return -308;
    long blocks = (t.count + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    k_tap<<<(unsigned)blocks, 256, 0, st>>>(t.ptr, t.kind, t.count, out);
    ndt::count_launches(1);
    return cudaGetLastError() == cudaSuccess ? t.count : -306;
}

void Model::release() {
    if (!impl) return;
    for (void *p : impl->allocs) cudaFree(p);
    delete impl;
    impl = nullptr;
}

namespace {

template <int BN, int STAGES>
bool launch_gemm(const CUtensorMap &a, const CUtensorMap &b, const CUtensorMap &c, const GemmArgs &g, dim3 grid, cudaStream_t st) {
    static bool attr[64] = {};          // function attributes are per device
    constexpr size_t smem = gemm_smem_bytes<BN, STAGES>();
    int dev = 0; cudaGetDevice(&dev);
    if (!attr[dev & 63]) {
        if (cudaFuncSetAttribute(k_gemm<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        attr[dev & 63] = true;
    }
    k_gemm<BN, STAGES><<<grid, kGemmThreads, smem, st>>>(a, b, c, g);
    return cudaGetLastError() == cudaSuccess;
}

struct Fwd {
    int B, P; cudaStream_t st; std::string *err;
    bool ok = true;
    int gemm_launches = 0;

    // activations [B][P][K] x weights [N][K] -> bf16 [B][P][N]  (rows mode)
    void rows(const __nv_bfloat16 *act, int K, const __nv_bfloat16 *w, int N, bool w_batched, const float *bias, const float *cbias,
              int ldcb, bool relu, __nv_bfloat16 *out) {
        if (!ok) return;
        const int BN = N >= 256 ? 256 : (N >= 128 ? 128 : 64);
        CUtensorMap ma, mb, mc;
        if (N % 64 != 0) { ok = false; *err = "rows(): the output width must be a multiple of 64 (TMA store boxes)"; return; }
        if (!make_map(&ma, act, K, P, B, 128) || !make_map(&mb, w, K, N, w_batched ? B : 1, BN) || !make_map(&mc, out, N, P, B, 128)) {
            ok = false; *err = "cuTensorMapEncodeTiled failed"; return;
        }
        GemmArgs g{}; g.K = K; g.P = P; g.a_batched = 1; g.b_batched = w_batched ? 1 : 0; g.mode = MODE_ROWS; g.relu = relu ? 1 : 0; g.n_valid = N;
        g.bias = bias; g.cbias = cbias; g.ldcb = ldcb; g.out = out; g.ldo = N;
        dim3 grid((N + BN - 1) / BN, (P + 127) / 128, B);
        // one k-block (K = 64): a single stage keeps the CTA at 24-48 KB of shared memory so several CTAs share an SM
        // and one CTA's epilogue overlaps another's load/MMA; deeper K: two stages (still 2 CTAs per SM at BN = 256)
        if (K <= 64) {
            if (BN == 256) ok = launch_gemm<256, 1>(ma, mb, mc, g, grid, st);
            else if (BN == 128) ok = launch_gemm<128, 1>(ma, mb, mc, g, grid, st);
            else ok = launch_gemm<64, 1>(ma, mb, mc, g, grid, st);
        } else {
            if (BN == 256) ok = launch_gemm<256, 2>(ma, mb, mc, g, grid, st);
            else if (BN == 128) ok = launch_gemm<128, 2>(ma, mb, mc, g, grid, st);
            else ok = launch_gemm<64, 2>(ma, mb, mc, g, grid, st);
        }
        if (!ok) *err = "gemm launch failed";
        gemm_launches++;
    }
    // weights [C][K=128] x activations [B][P][128] -> max over points per cloud, encoded [B][ldg]
    void chmax(const __nv_bfloat16 *w, int C, const __nv_bfloat16 *act, int K, const float *bias, bool relu, unsigned *gmax, int ldg) {
        if (!ok) return;
        if (K != 128) { ok = false; *err = "k_gemm_chmax expects K = 128"; return; }
        CUtensorMap mw, mx;
        if (!make_map(&mw, w, K, C, 1, 128) || !make_map(&mx, act, K, P, B, 64)) { ok = false; *err = "cuTensorMapEncodeTiled failed"; return; }
        GemmArgs g{}; g.K = K; g.P = P; g.mode = MODE_CHMAX; g.relu = relu ? 1 : 0; g.n_valid = C; g.bias = bias; g.gmax = gmax; g.ldg = ldg;
        static bool attr[64] = {};      // function attributes are per device
        int dev = 0; cudaGetDevice(&dev);
        if (!attr[dev & 63]) {
            if (cudaFuncSetAttribute(k_gemm_chmax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChmaxSmemBytes) != cudaSuccess) { ok = false; *err = "smem attribute"; return; }
            attr[dev & 63] = true;
        }
        dim3 grid((C + 511) / 512, B);
        k_gemm_chmax<<<grid, kGemmThreads, kChmaxSmemBytes, st>>>(mw, mx, g);
        ok = cudaGetLastError() == cudaSuccess;
        if (!ok) *err = "gemm launch failed";
        gemm_launches++;
    }
    // fused head layers 1 + 2 (k_head12): x_t2 [B][P][64] -> relu(relu(x W1a^T + cb[cloud]) W2^T + b2) [B][P][256]
    void head12(const __nv_bfloat16 *xt2, const __nv_bfloat16 *w1a, const float *cbias, const __nv_bfloat16 *w2, const float *b2, __nv_bfloat16 *out) {
        if (!ok) return;
        CUtensorMap mx, mw1, mw2, mo;
        if (!make_map(&mx, xt2, 64, P, B, 128) || !make_map(&mw1, w1a, 64, 512, 1, 256) || !make_map(&mw2, w2, 512, 256, 1, 256) ||
            !make_map(&mo, out, 256, P, B, 128)) { ok = false; *err = "cuTensorMapEncodeTiled failed"; return; }
        static bool attr[64] = {};      // function attributes are per device
        int dev = 0; cudaGetDevice(&dev);
        if (!attr[dev & 63]) {
            if (cudaFuncSetAttribute(k_head12, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHead12SmemBytes) != cudaSuccess) { ok = false; *err = "smem attribute"; return; }
            attr[dev & 63] = true;
        }
        Head12Args a{P, cbias, b2};
        k_head12<<<dim3((P + 127) / 128, B), kHead12Threads, kHead12SmemBytes, st>>>(mx, mw1, mw2, mo, a);
        ok = cudaGetLastError() == cudaSuccess;
        if (!ok) *err = "k_head12 launch failed";
        gemm_launches++;
    }
    void logsm(const __nv_bfloat16 *act, int K, const __nv_bfloat16 *w, int N, const float *bias, float *out) {
        if (!ok) return;
        CUtensorMap ma, mb;
        if (!make_map(&ma, act, K, P, B, 128) || !make_map(&mb, w, K, N, 1, 32)) { ok = false; *err = "cuTensorMapEncodeTiled failed"; return; }
        GemmArgs g{}; g.K = K; g.P = P; g.a_batched = 1; g.b_batched = 0; g.mode = MODE_LOGSM; g.n_valid = N; g.bias = bias; g.outf = out;
        dim3 grid(1, (P + 127) / 128, B);
        ok = launch_gemm<32, 2>(ma, mb, ma, g, grid, st);
        if (!ok) *err = "gemm launch failed";
        gemm_launches++;
    }
    void fc(const float *W, const float *bias, const void *x, int ldx, bool decode, float *y, int ldy, int in, int out, bool relu,
            int identity_dim = 0, __nv_bfloat16 *y_t = nullptr) {
        if (!ok) return;
        dim3 grid((out + kFcTN - 1) / kFcTN, (B + kFcTM - 1) / kFcTM);
        k_fc<<<grid, 256, 0, st>>>(W, (out + 3) / 4 * 4, bias, x, ldx, decode ? 1 : 0, y, ldy, B, in, out, relu ? 1 : 0, identity_dim, y_t);
    }
};

}  // namespace

int Model::forward(Scratch &scratch, const float *feat, int B, int D, float *out, cudaStream_t st, std::string &err) {
    if (!impl) { err = "model not built"; return -300; }
    ModelImpl &m = *impl;
    const int P = D;
    const size_t M = (size_t)B * P;
    // scratch layout (bump allocation, 256 B aligned)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_h64 = take(M * 64 * 2), o_h128 = take(M * 128 * 2), o_a1 = take(M * 64 * 2), o_xt2 = take(M * 64 * 2);
    const size_t o_gmax = take((size_t)B * (1024 + 1024 + m.F) * 4);
    const size_t o_f512 = take((size_t)B * 512 * 4), o_f256 = take((size_t)B * 256 * 4), o_T1 = take((size_t)B * 256 * 4);
    const size_t o_T2 = take((size_t)B * 4096 * 4), o_T2t = take((size_t)B * 4096 * 2);
    const size_t o_cb = take((size_t)B * 512 * 4), o_logit = take((size_t)B * (m.ncls > 0 ? m.ncls : 1) * 4);
    const size_t o_s1 = take(m.kind == 1 ? M * 512 * 2 : 0), o_s2 = take(m.kind == 1 ? M * 256 * 2 : 0);
    if (scratch.reserve(off) != cudaSuccess) { err = "scratch allocation failed"; return -304; }
    uint8_t *base = (uint8_t *)scratch.buf;
    __nv_bfloat16 *h64 = (__nv_bfloat16 *)(base + o_h64), *h128 = (__nv_bfloat16 *)(base + o_h128), *a1 = (__nv_bfloat16 *)(base + o_a1),
                  *xt2 = (__nv_bfloat16 *)(base + o_xt2), *T2t = (__nv_bfloat16 *)(base + o_T2t), *s1 = (__nv_bfloat16 *)(base + o_s1),
                  *s2 = (__nv_bfloat16 *)(base + o_s2);
    unsigned *g1 = (unsigned *)(base + o_gmax), *g2 = g1 + (size_t)B * 1024, *g3 = g2 + (size_t)B * 1024;
    float *f512 = (float *)(base + o_f512), *f256 = (float *)(base + o_f256), *T1 = (float *)(base + o_T1), *T2 = (float *)(base + o_T2),
          *cb = (float *)(base + o_cb), *logit = (float *)(base + o_logit);

    Fwd f{B, P, st, &err};
    if (cudaMemsetAsync(g1, 0, (size_t)B * (2048 + m.F) * 4, st) != cudaSuccess) { err = "memset failed"; return -305; }
    const int total = (int)M;
    // ---- input transform T-Net (ndtnet.py:132-133)
    if (m.pointnet) k_tnet_l1<<<(total + 127) / 128, 128, 0, st>>>(feat, m.in_dim, total, m.t1.l1_w, m.t1.l1_b, h64);
    else k_tnet3_l1<<<(total + 127) / 128, 128, 0, st>>>(feat, total, m.t1.l1_w, m.t1.l1_b, h64);
    f.rows(h64, 64, m.t1.l2_w, 128, false, m.t1.l2_b, nullptr, 0, true, h128);
    f.chmax(m.t1.l3_w, 1024, h128, 128, m.t1.l3_b, true, g1, 1024);
    f.fc(m.t1.fc1_w, m.t1.fc1_b, g1, 1024, true, f512, 512, 1024, 512, true);
    f.fc(m.t1.fc2_w, m.t1.fc2_b, f512, 512, false, f256, 256, 512, 256, true);
    const int td = m.t1.in;
    f.fc(m.t1.fc3_w, m.t1.fc3_b, f256, 256, false, T1, td * td, 256, td * td, false, td);
    // ---- transform + conv1/bn1 (ndtnet.py:135-149)
    if (m.pointnet) k_trunk_l1_pn<<<(total + 127) / 128, 128, 0, st>>>(feat, m.in_dim, P, total, T1, m.c1_w, m.c1_b, a1);
    else k_trunk_l1<<<(total + 127) / 128, 128, 0, st>>>(feat, P, total, T1, m.c1_w, m.c1_b, a1);
    // ---- feature transform T-Net (ndtnet.py:152)
    f.rows(a1, 64, m.t2.l1_wb, 64, false, m.t2.l1_b, nullptr, 0, true, h64);
    f.rows(h64, 64, m.t2.l2_w, 128, false, m.t2.l2_b, nullptr, 0, true, h128);
    f.chmax(m.t2.l3_w, 1024, h128, 128, m.t2.l3_b, true, g2, 1024);
    f.fc(m.t2.fc1_w, m.t2.fc1_b, g2, 1024, true, f512, 512, 1024, 512, true);
    f.fc(m.t2.fc2_w, m.t2.fc2_b, f512, 512, false, f256, 256, 512, 256, true);
    f.fc(m.t2.fc3_w, m.t2.fc3_b, f256, 256, false, T2, 4096, 256, 4096, false, 64, T2t);
    // ---- x . T2 (ndtnet.py:153-155), conv2/bn2, conv3/bn3 + max over the points (ndtnet.py:160-161,186,224)
    f.rows(a1, 64, T2t, 64, true, nullptr, nullptr, 0, false, xt2);
    f.rows(xt2, 64, m.c2_w, 128, false, m.c2_b, nullptr, 0, false, h128);
    f.chmax(m.c3_w, m.F, h128, 128, m.c3_b, false, g3, m.F);
    if (m.kind == 1) {
        // ---- segmentation head (ndtnet.py:224-241)
        f.fc(m.h1g_w, m.h1_b, g3, m.F, true, cb, 512, m.F, 512, false);
        if (m.fuse_head) {
            f.head12(xt2, m.h1a_w, cb, m.h2_w, m.h2_b, s2);           // the 512-wide activation stays in TMEM / shared memory
        } else {
            f.rows(xt2, 64, m.h1a_w, 512, false, nullptr, cb, 512, true, s1);
            f.rows(s1, 512, m.h2_w, 256, false, m.h2_b, nullptr, 0, true, s2);
        }
        f.rows(s2, 256, m.h3_w, 128, false, m.h3_b, nullptr, 0, true, h128);
        f.logsm(h128, 128, m.h4_w, m.ncls, m.h4_b, out);
    } else {
        // ---- classification head (ndtnet.py:186-194)
        f.fc(m.k1_w, m.k1_b, g3, m.F, true, f512, 512, m.F, 512, true);
        f.fc(m.k2_w, m.k2_b, f512, 512, false, f256, 256, 512, 256, true);
        f.fc(m.k3_w, m.k3_b, f256, 256, false, logit, m.ncls, 256, m.ncls, false);
        if (f.ok) k_softmax<<<B, 256, 0, st>>>(logit, out, m.ncls);
    }
    if (!f.ok) return -306;
    {
        const long Bl = B, Ml = (long)M;
        m.taps.clear();
        m.taps["t1"] = {T1, 0, Bl * td * td};                  // input transform  (ndtnet.py:132-133)   [B, d, d]
        m.taps["t2"] = {T2, 0, Bl * 4096};                     // feature transform (ndtnet.py:152)        [B, 64, 64]
        m.taps["t1.pool"] = {g1, 2, Bl * 1024};                // T-Net max-pools (ndtnet.py:50)           [B, 1024]
        m.taps["t2.pool"] = {g2, 2, Bl * 1024};
        m.taps["trunk.l1"] = {a1, 1, Ml * 64};                 // bn1(conv1([T p ; T Sigma]))  (ndtnet.py:149)  [B, N, 64]
        m.taps["trunk.xt2"] = {xt2, 1, Ml * 64};               // x . T2  (ndtnet.py:153-155)              [B, N, 64]
        m.taps["trunk.pool"] = {g3, 2, Bl * m.F};              // max over points of bn3(conv3(.))         [B, F]
        if (m.kind == 1) {
            if (!m.fuse_head) m.taps["head.l1"] = {s1, 1, Ml * 512};   // segmentation head layers (ndtnet.py:231-233); fused: never stored
            m.taps["head.l2"] = {s2, 1, Ml * 256};
            m.taps["head.l3"] = {h128, 1, Ml * 128};
        } else {
            m.taps["trunk.l2"] = {h128, 1, Ml * 128};          // bn2(conv2(.)) survives when no head reuses the buffer
        }
    }
    ndt::count_launches(f.gemm_launches + 2 + (m.kind == 1 ? 7 : 10));   // + the small CUDA-core kernels
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return -100 - (int)e; }
    return 0;
}

}  // namespace mlp
