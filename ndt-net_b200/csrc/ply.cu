// ply.cu — ASCII-PLY ingest on the device.
//
// Replaces the per-line Python loop of the reference's dataset reader
// (/root/reference/ndnet/datasets/CARLA_Seg.py:96-183, `CARLA_Seg.get_data_pcl`): skip `num_header_lines`
// lines, then per line `x y z ... class_tag` -> float(data[0..2]) and int(data[-1]), reject a tag above n_classes
// (:127-128), keep the points as float32 (`.float()`, :173) and the tags as uint16 (:146), then gather the caller's
// random subsample (:141-147) and build the one-hot ground truth (:176-179).
//
// Byte work, HBM-bound: the file is read three times by coalesced 16-byte loads (newline count, line starts) and by
// a thread-per-line parser whose byte loads stay in L1.  Decimal literals are converted to the CORRECTLY ROUNDED
// double (what CPython's float() returns) with integer arithmetic only: Clinger's exact fast path when the digits
// fit 2^53 and |exponent| <= 22, otherwise a 128-bit long division / 192-bit product by 5^|e| followed by
// round-half-even.  Literals outside the supported grammar (inf/nan/underscores, more than 19 significant digits,
// |decimal exponent| > 55) are refused with an error code — never approximated.
#include "../../include/ndnet_b200.h"
#include "ply_decimal.cuh"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <new>

namespace ply {

constexpr int kScanThreads = 256;
constexpr int kBytesPerThread = 16;
constexpr int kBytesPerCta = kScanThreads * kBytesPerThread;   // 4 KB of text per CTA

__device__ __forceinline__ unsigned newline_mask(uint4 v, long base, long nbytes, const unsigned char *text, unsigned *bad) {
    // bit i set when byte i of the 16 is '\n'; flags non-ASCII bytes and '\r' not followed by '\n'
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    unsigned m = 0, b = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const unsigned c = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
        if (base + i < nbytes) {
            m |= (c == '\n') << i;
            b |= c >= 0x80u;
            if (c == '\r') {
                unsigned nxt;
                if (i < 15) nxt = (w[(i + 1) >> 2] >> (((i + 1) & 3) * 8)) & 0xffu;
                else nxt = base + 16 < nbytes ? text[base + 16] : 0u;
                if (base + i + 1 >= nbytes) nxt = '\n';       // a trailing '\r' is stripped as whitespace
                b |= nxt != '\n';
            }
        }
    }
    *bad |= b;
    return m;
}

// Pass 1: newlines per 4 KB of text.
__global__ void __launch_bounds__(kScanThreads) k_ply_count(const unsigned char *__restrict__ text, long nbytes,
                                                            unsigned *__restrict__ cta_count, unsigned *__restrict__ enc_bad) {
    const long base = ((long)blockIdx.x * kScanThreads + threadIdx.x) * kBytesPerThread;
    unsigned bad = 0, n = 0;
    if (base < nbytes) {
        const uint4 v = *reinterpret_cast<const uint4 *>(text + base);        // the buffer is padded to 16 bytes
        n = __popc(newline_mask(v, base, nbytes, text, &bad));
    }
    __shared__ unsigned s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    n = __reduce_add_sync(0xffffffffu, n);
    bad = __any_sync(0xffffffffu, bad != 0);
    if ((threadIdx.x & 31) == 0) {
        if (n) atomicAdd(&s_n, n);
        if (bad) atomicOr(enc_bad, 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) cta_count[blockIdx.x] = s_n;
}

// Exclusive scan of the per-CTA counts (one CTA; the count array is nbytes / 4096 long).
__global__ void __launch_bounds__(1024) k_ply_scan(unsigned *__restrict__ cta_count, int n, u64 *__restrict__ total) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < n ? cta_count[i] : 0u;
        unsigned x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31) >= d) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned wsum = s_warp[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, wsum, d);
                if (threadIdx.x >= d) wsum += y;
            }
            s_warp[threadIdx.x] = wsum;       // inclusive over warps
        }
        __syncthreads();
        const unsigned before = s_carry + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0u) + x - v;
        if (i < n) cta_count[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

// Pass 2: line_start[k + 1] = byte after the k-th newline (line_start[0] = 0 is written by the host side).
__global__ void __launch_bounds__(kScanThreads) k_ply_starts(const unsigned char *__restrict__ text, long nbytes,
                                                             const unsigned *__restrict__ cta_offset, u64 *__restrict__ line_start) {
    const long base = ((long)blockIdx.x * kScanThreads + threadIdx.x) * kBytesPerThread;
    unsigned m = 0, dummy = 0;
    if (base < nbytes) m = newline_mask(*reinterpret_cast<const uint4 *>(text + base), base, nbytes, text, &dummy);
    const unsigned n = __popc(m);
    unsigned x = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, d);
        if ((threadIdx.x & 31) >= d) x += y;
    }
    __shared__ unsigned s_warp[kScanThreads / 32];
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    unsigned before = cta_offset[blockIdx.x] + x - n;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += s_warp[w];
    while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        line_start[++before] = (u64)(base + i + 1);
    }
}

__device__ __forceinline__ void report(u64 *err, long line, int code) {
    atomicMin(err, ((u64)line << 8) | (u64)code);               // the reference raises at the FIRST offending line
}

// [a, b) of the last whitespace-separated token of a line (data[-1]); a = -1 when the line is blank
__device__ __forceinline__ void last_token(const unsigned char *__restrict__ text, long a, long b, long *la, long *lb) {
    *la = -1; *lb = -1;
    long i = a;
    while (i < b) {
        while (i < b && is_space(text[i])) i++;
        if (i >= b) break;
        *la = i;
        while (i < b && !is_space(text[i])) i++;
        *lb = i;
    }
}

// Thread per data line.
__global__ void __launch_bounds__(128) k_ply_parse(const unsigned char *__restrict__ text, long nbytes,
                                                   const u64 *__restrict__ line_start, long n_lines, long header, int n_classes,
                                                   float *__restrict__ pts, uint16_t *__restrict__ lab, u64 *__restrict__ err,
                                                   u64 *__restrict__ err_negative) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long line = header + j;
    if (line >= n_lines) return;
    long a = (long)line_start[line];
    long b = line + 1 < n_lines ? (long)line_start[line + 1] : nbytes;    // one past the line's last byte (incl. its '\n')
    double xyz[3];
    long long tag = 0;
    const int r = parse_line(text, a, b, xyz, &tag);          // tokens, float(data[0..2]), int(data[-1]) in the reference's order
    if (r) { report(err, line, r); return; }
    if (tag > (long long)n_classes) { report(err, line, kErrClassBound); return; }
    if (tag < 0) { report(err_negative, line, kErrNegativeClass); return; }
    pts[j * 3 + 0] = __double2float_rn(xyz[0]);
    pts[j * 3 + 1] = __double2float_rn(xyz[1]);
    pts[j * 3 + 2] = __double2float_rn(xyz[2]);
    lab[j] = (uint16_t)tag;
}

// The class tag of the first offending line, for the reference's "Class tag {tag} out of bounds" message (:128).
__global__ void k_ply_error_value(const unsigned char *__restrict__ text, long nbytes, const u64 *__restrict__ line_start,
                                  long n_lines, const u64 *__restrict__ err, const u64 *__restrict__ err_negative,
                                  long long *__restrict__ value) {
    u64 e = *err != ~0ull ? *err : *err_negative;
    if (e == ~0ull) return;
    const long line = (long)(e >> 8);
    const long a = (long)line_start[line], b = line + 1 < n_lines ? (long)line_start[line + 1] : nbytes;
    long la, lb;
    last_token(text, a, b, &la, &lb);
    long long tag = 0;
    if (la >= 0 && parse_int(text, la, lb, &tag) == 0) *value = tag;
}

// Subsample gather + one-hot rows (CARLA_Seg.py:141-147,173-179).
__global__ void k_ply_sample(const float *__restrict__ pts, const uint16_t *__restrict__ lab, const long long *__restrict__ idx,
                             long n, long n_points, int bins, float *__restrict__ out_pts, uint16_t *__restrict__ out_lab,
                             float *__restrict__ out_onehot, unsigned *__restrict__ oob) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = n * (long)(bins > 3 ? bins : 3);
    if (t >= total) return;
    const int width = bins > 3 ? bins : 3;
    const long row = t / width;
    const int col = (int)(t - row * width);
    long long src = idx ? idx[row] : row;
    if (src < 0) src += n_points;                        // numpy-style negative indexes
    if (src < 0 || src >= n_points) { atomicOr(oob, 1u); return; }
    if (col < 3 && out_pts) out_pts[row * 3 + col] = pts[src * 3 + col];
    const unsigned c = lab[src];
    if (col == 0 && out_lab) out_lab[row] = (uint16_t)c;
    if (col < bins && out_onehot) out_onehot[row * bins + col] = c == (unsigned)col ? 1.0f : 0.0f;
}

}  // namespace ply

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct ndnet_b200_ply {
    int device = 0;
    cudaStream_t stream = nullptr;      // stream-ordered allocations are returned on the stream that made them
    int n_classes = 0;
    long n_points = 0;
    unsigned char *d_text = nullptr;
    unsigned *d_cta = nullptr;
    ply::u64 *d_line_start = nullptr;
    ply::u64 *d_words = nullptr;        // [0] total newlines, [1] first error, [2] first negative tag, [3] error value
    unsigned *d_flags = nullptr;        // [0] encoding, [1] sample index out of range
    float *d_pts = nullptr;
    uint16_t *d_lab = nullptr;
    long long *d_idx = nullptr; size_t idx_cap = 0;
};

namespace {

int ply_fail(cudaError_t e, const char *where) {
    fprintf(stderr, "ndnet_b200: ply: %s: %s\n", where, cudaGetErrorString(e));
    return -100 - (int)e;
}

}  // namespace

extern "C" void ndnet_b200_ply_free(ndnet_b200_ply *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    void *ptrs[] = {p->d_text, p->d_cta, p->d_line_start, p->d_words, p->d_flags, p->d_pts, p->d_lab, p->d_idx};
    for (void *q : ptrs) if (q) cudaFreeAsync(q, p->stream);
    delete p;
}

extern "C" int ndnet_b200_ply_load(int device, const char *text, size_t nbytes, int text_on_device, int num_header_lines,
                                   int n_classes, void *stream, ndnet_b200_ply **out, unsigned long *num_points,
                                   long *bad_line, long *bad_value) {
    using namespace ply;
    if (!out || (!text && nbytes) || num_header_lines < 0 || n_classes < 0) return -200;
    *out = nullptr;
    if (bad_line) *bad_line = -1;
    if (bad_value) *bad_value = 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return ply_fail(e, "cudaGetDeviceCount");
    if (count == 0 || device < 0 || device >= count) {
        fprintf(stderr, "ndnet_b200: ply: no CUDA device %d (found %d); this library has no CPU path\n", device, count);
        return -201;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) return ply_fail(e, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    ndnet_b200_ply *p = new (std::nothrow) ndnet_b200_ply();
    if (!p) return -202;
    p->device = device; p->n_classes = n_classes; p->stream = st;
    {   // buffers come from the device's stream-ordered pool and stay cached in it between files (no cudaMalloc per scan)
        static bool pool_ready[64] = {};
        if (device < 64 && !pool_ready[device]) {
            cudaMemPool_t pool;
            unsigned long long keep = ~0ull;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            pool_ready[device] = true;
        }
    }
    const long nb = (long)nbytes;
    const int n_cta = (int)((nb + kBytesPerCta - 1) / kBytesPerCta);
    const size_t padded = ((size_t)nb + 15) / 16 * 16 + 32;
#define PLY_TRY(call, where) if ((e = (call)) != cudaSuccess) { ndnet_b200_ply_free(p); return ply_fail(e, where); }
    PLY_TRY(cudaMallocAsync((void **)&p->d_text, padded, st), "text allocation");
    PLY_TRY(cudaMallocAsync((void **)&p->d_cta, sizeof(unsigned) * (size_t)(n_cta + 1), st), "scan allocation");
    PLY_TRY(cudaMallocAsync((void **)&p->d_words, sizeof(u64) * 4, st), "state allocation");
    PLY_TRY(cudaMallocAsync((void **)&p->d_flags, sizeof(unsigned) * 2, st), "state allocation");
    PLY_TRY(cudaMemsetAsync(p->d_flags, 0, sizeof(unsigned) * 2, st), "memset");
    PLY_TRY(cudaMemsetAsync(p->d_words, 0xff, sizeof(u64) * 3, st), "memset");
    PLY_TRY(cudaMemsetAsync(p->d_words + 3, 0, sizeof(u64), st), "memset");
    PLY_TRY(cudaMemsetAsync(p->d_text + padded - 32, ' ', 32, st), "memset");
    if (nb) PLY_TRY(cudaMemcpyAsync(p->d_text, text, nbytes, text_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st), "text copy");
    u64 newlines = 0;
    unsigned char last = '\n';
    if (nb) {
        k_ply_count<<<n_cta, kScanThreads, 0, st>>>(p->d_text, nb, p->d_cta, p->d_flags);
        k_ply_scan<<<1, 1024, 0, st>>>(p->d_cta, n_cta, p->d_words);
        PLY_TRY(cudaMemcpyAsync(&newlines, p->d_words, sizeof(u64), cudaMemcpyDeviceToHost, st), "D2H");
        PLY_TRY(cudaMemcpyAsync(&last, p->d_text + nb - 1, 1, cudaMemcpyDeviceToHost, st), "D2H");
        PLY_TRY(cudaStreamSynchronize(st), "line count");
    }
    const long n_lines = (long)newlines + (nb && last != '\n' ? 1 : 0);       // readlines() (:112)
    const long n_points = n_lines > num_header_lines ? n_lines - num_header_lines : 0;
    p->n_points = n_points;
    PLY_TRY(cudaMallocAsync((void **)&p->d_line_start, sizeof(u64) * (size_t)(newlines + 2), st), "line index allocation");
    PLY_TRY(cudaMallocAsync((void **)&p->d_pts, sizeof(float) * 3 * (size_t)(n_points ? n_points : 1), st), "points allocation");
    PLY_TRY(cudaMallocAsync((void **)&p->d_lab, sizeof(uint16_t) * (size_t)(n_points ? n_points : 1), st), "labels allocation");
    u64 words[4] = {0, ~0ull, ~0ull, 0};
    unsigned flags[2] = {0, 0};
    if (nb) {
        PLY_TRY(cudaMemsetAsync(p->d_line_start, 0, sizeof(u64), st), "memset");
        k_ply_starts<<<n_cta, kScanThreads, 0, st>>>(p->d_text, nb, p->d_cta, p->d_line_start);
        if (n_points)
            k_ply_parse<<<(unsigned)((n_points + 127) / 128), 128, 0, st>>>(p->d_text, nb, p->d_line_start, n_lines, num_header_lines,
                                                                             n_classes, p->d_pts, p->d_lab, p->d_words + 1,
                                                                             p->d_words + 2);
        if (n_points)
            k_ply_error_value<<<1, 1, 0, st>>>(p->d_text, nb, p->d_line_start, n_lines, p->d_words + 1, p->d_words + 2,
                                               (long long *)(p->d_words + 3));
        PLY_TRY(cudaGetLastError(), "kernel launch");
        PLY_TRY(cudaMemcpyAsync(words, p->d_words, sizeof(words), cudaMemcpyDeviceToHost, st), "D2H");
        PLY_TRY(cudaMemcpyAsync(flags, p->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st), "D2H");
        PLY_TRY(cudaStreamSynchronize(st), "parse");
    }
#undef PLY_TRY
    int code = 0;
    long line = -1;
    if (flags[0]) code = kErrEncoding;
    else if (words[1] != ~0ull) { code = (int)(words[1] & 0xff); line = (long)(words[1] >> 8); }
    else if (words[2] != ~0ull) { code = (int)(words[2] & 0xff); line = (long)(words[2] >> 8); }
    if (code) {
        if (bad_line) *bad_line = line;
        if (bad_value) *bad_value = (long)(long long)words[3];
        ndnet_b200_ply_free(p);
        return -300 - code;
    }
    // the text and the line index are no longer needed
    cudaFreeAsync(p->d_text, st); p->d_text = nullptr;
    cudaFreeAsync(p->d_line_start, st); p->d_line_start = nullptr;
    cudaFreeAsync(p->d_cta, st); p->d_cta = nullptr;
    if (num_points) *num_points = (unsigned long)n_points;
    *out = p;
    return 0;
}

extern "C" long ndnet_b200_ply_num_points(const ndnet_b200_ply *p) { return p ? p->n_points : -200; }

extern "C" int ndnet_b200_ply_sample(ndnet_b200_ply *p, const int64_t *indexes, size_t n, int indexes_on_device,
                                     float *out_points, uint16_t *out_labels, float *out_onehot, void *stream) {
    using namespace ply;
    if (!p) return -200;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSetDevice(p->device);
    if (e != cudaSuccess) return ply_fail(e, "cudaSetDevice");
    const long rows = indexes ? (long)n : p->n_points;
    if (rows == 0) return 0;
    const long long *d_idx = nullptr;
    if (indexes && indexes_on_device) d_idx = (const long long *)indexes;
    else if (indexes) {
        if (n > p->idx_cap) {
            if (p->d_idx) cudaFreeAsync(p->d_idx, st);
            p->d_idx = nullptr; p->idx_cap = 0;
            if ((e = cudaMallocAsync((void **)&p->d_idx, sizeof(long long) * n, st)) != cudaSuccess) return ply_fail(e, "index allocation");
            p->idx_cap = n;
        }
        if ((e = cudaMemcpyAsync(p->d_idx, indexes, sizeof(long long) * n, cudaMemcpyHostToDevice, st)) != cudaSuccess)
            return ply_fail(e, "H2D indexes");
        d_idx = p->d_idx;
    }
    const int bins = out_onehot ? p->n_classes + 1 : 0;
    const long total = rows * (long)(bins > 3 ? bins : 3);
    if ((e = cudaMemsetAsync(p->d_flags + 1, 0, sizeof(unsigned), st)) != cudaSuccess) return ply_fail(e, "memset");
    k_ply_sample<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p->d_pts, p->d_lab, d_idx, rows, p->n_points, bins, out_points,
                                                                  out_labels, out_onehot, p->d_flags + 1);
    if ((e = cudaGetLastError()) != cudaSuccess) return ply_fail(e, "k_ply_sample");
    unsigned oob = 0;
    if ((e = cudaMemcpyAsync(&oob, p->d_flags + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return ply_fail(e, "D2H");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return ply_fail(e, "sample");
    return oob ? -307 : 0;
}
