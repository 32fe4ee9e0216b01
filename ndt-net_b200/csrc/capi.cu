// capi.cu — the C ABI of libndnet_b200.so (include/ndnet_b200.h): contexts, workspaces, the batched
// device/host entry points and the legacy host-pointer symbols of the reference's libndnet.so.
#include "../../include/ndnet_b200.h"
#include "ndt_host.h"
#include "mlp_host.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using ndt::NdtCloudInfo;
using ndt::Workspace;

static_assert(sizeof(NdtCloudInfo) == sizeof(ndnet_b200_cloud_info), "info layout mismatch");

struct ndnet_b200_ctx {
    int device = 0;
    Workspace ws;
    std::string err;
    // staging for the *_host entry point
    void *d_points = nullptr; size_t d_points_bytes = 0;
    uint16_t *d_labels = nullptr; size_t d_labels_bytes = 0;
    float *d_feat = nullptr; size_t d_feat_bytes = 0;
    double *d_feat64 = nullptr; size_t d_feat64_bytes = 0;
    uint16_t *d_olab = nullptr; size_t d_olab_bytes = 0;
    int32_t *d_ovox = nullptr; size_t d_ovox_bytes = 0;
    NdtCloudInfo *d_info = nullptr; size_t d_info_bytes = 0;
    float *d_logits = nullptr; size_t d_logits_bytes = 0;
    mlp::Scratch mlp_scratch;
    // pipeline lanes of ndnet_b200_infer_{host,device}: chunks of the batch run on separate streams so that
    // copies overlap kernels and the latency-bound tail of one chunk overlaps the bulk work of another
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        cudaEvent_t copied = nullptr;     // this lane's input chunk has arrived (recorded on the copy stream)
        cudaEvent_t consumed = nullptr;   // the kernels that read this lane's input buffers have been enqueued up to here
        bool used = false;
        Workspace ws;
        mlp::Scratch scratch;
        void *d_points = nullptr; size_t d_points_bytes = 0;
        uint16_t *d_labels = nullptr; size_t d_labels_bytes = 0;
        float *d_feat = nullptr; size_t d_feat_bytes = 0;
        float *d_logits = nullptr; size_t d_logits_bytes = 0;
    };
    std::vector<Lane> lanes;
    int n_lanes = 2;
    int chunk = 64;          // scans per chunk of ndnet_b200_infer_host (copies overlap kernels: more, smaller chunks)
    bool stagger = false;    // lanes start their fronts one behind the other (infer_pipelined)
    int chunk_device = 128;  // scans per chunk of ndnet_b200_infer_device (no copies to hide: fewer, larger chunks)
    cudaEvent_t start_ev = nullptr;
    // CUDA-graph replay of the NDT chain for small batches (ndnet_b200_downsample_batch): ~55 dependent launches, most of
    // them a few microseconds long, are launch-latency bound below a few dozen scans
    struct NdtGraph {
        int dtype = 0, B = 0, num_classes = 0; long N = 0, D = 0; unsigned flags = 0, wants = 0, ws_flags = 0;
        unsigned long ws_generation = 0, stamp = 0;
        cudaGraphExec_t exec = nullptr;
        int state = 0;                 // 0: free slot; 1: shape seen once (launched directly); 2: graph ready; 3: capture failed (launch directly)
        void *in_pts = nullptr; uint16_t *in_lab = nullptr;
        float *o_feat = nullptr; double *o_feat64 = nullptr; uint16_t *o_lab = nullptr; int32_t *o_vox = nullptr; NdtCloudInfo *o_info = nullptr;
        long launches = 0;
        void destroy() {
            if (exec) cudaGraphExecDestroy(exec);
            void *p[] = {in_pts, in_lab, o_feat, o_feat64, o_lab, o_vox, o_info};
            for (void *q : p) if (q) cudaFree(q);
            *this = NdtGraph();
        }
    };
    std::vector<NdtGraph> graphs;
    int graph_mode = -1;               // -1: batches of at most kGraphAutoPoints points; 0: never; 1: every batch
    unsigned long graph_clock = 0;
    cudaStream_t graph_stream = nullptr;
    // host -> device copies of ndnet_b200_infer_host all go through ONE stream, in chunk order: copies issued on the lanes'
    // own streams are served concurrently by the copy engine, so every chunk's data would arrive near the end of the
    // step and the kernels could not overlap the transfer
    cudaStream_t copy_stream = nullptr;
};

namespace {

int fail(ndnet_b200_ctx *ctx, cudaError_t e, const char *where) {
    if (ctx) ctx->err = std::string(where) + ": " + cudaGetErrorString(e);
    fprintf(stderr, "ndnet_b200: %s: %s\n", where, cudaGetErrorString(e));
    return -100 - (int)e;
}

template <typename T>
cudaError_t grow(T *&p, size_t &have, size_t need) {
    if (need <= have) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; have = 0;
    cudaError_t e = cudaMalloc((void **)&p, need ? need : 1);
    if (e == cudaSuccess) have = need;
    return e;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
namespace ndt {

template <typename T>
static cudaError_t alloc(T *&p, size_t count) {
    return cudaMalloc((void **)&p, (count ? count : 1) * sizeof(T));
}

// Frees the device arrays and resets every capacity to zero (a later reserve() starts from scratch); the side
// stream and its events survive so that a workspace can grow without re-creating them (destroy() ends them).
void Workspace::release() {
    const bool keep = keep_point_voxels, keep_list = keep_kl_list;
    cudaStream_t keep_side = side; cudaEvent_t keep_fork = ev_fork, keep_join = ev_join, keep_front = ev_front;
    const bool keep_mark = mark_front;
    const unsigned long keep_gen = generation + 1;
    const StageTimer keep_timer = timer;
    void *ptrs[] = {states, lim_enc, bitmap, vox_cell, vox_n, vox_start, vox_order, slot_rank, tile_cnt, hist, point_voxel, sorted,
                    mean, cov, cov_final, cls, kl_div, kl_flag, key, seq, firstpos, removed, list_div, list_seq, recip};
    for (void *p : ptrs) if (p) cudaFree(p);
    *this = Workspace();
    keep_point_voxels = keep; keep_kl_list = keep_list;
    side = keep_side; ev_fork = keep_fork; ev_join = keep_join; ev_front = keep_front; mark_front = keep_mark;
    timer = keep_timer;
    generation = keep_gen;
}

void Workspace::destroy() {
    release();
    if (timer.created) { for (auto &e : timer.ev) cudaEventDestroy(e); timer = StageTimer(); }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_front) cudaEventDestroy(ev_front);
    if (side) cudaStreamDestroy(side);
    side = nullptr; ev_fork = ev_join = ev_front = nullptr;
}

cudaError_t Workspace::reserve(int B, long N, long D, int bins) {
    const unsigned need_vcap = (unsigned)((double)D * 1.2) + 2;
    if (B <= B_cap && N <= N_cap && need_vcap <= vcap && bins <= bins_cap) return cudaSuccess;
    const int nB = B > B_cap ? B : B_cap;
    const long nN = N > N_cap ? N : N_cap;
    const long nD = D > D_cap ? D : D_cap;
    const int nbins = bins > bins_cap ? bins : bins_cap;
    const bool inject_failure = fail_next_reserve;
    release();
    B_cap = nB; N_cap = nN; D_cap = nD; bins_cap = nbins;
    vcap = (unsigned)((double)nD * 1.2) + 2;
    // Bitmap words per cloud: enough for kMaxGridCells when the batch is small, never less than 64K
    // cells; the search answers -1 (as the reference does when malloc fails) beyond what is held.
    bitmap_stride = (size_t)(1u << 25) / 32;
    ntiles_cap = (int)((nN + 2047) / 2048);   // >= ceil(N / kRankTile)
    const size_t kcap = (size_t)vcap * 6;
    cudaError_t e;
    // a failed allocation leaves NO capacity behind: release() frees what was obtained and zeroes the caps, so the next
    // call re-allocates instead of passing the early-out above with null or partial buffers
#define A(ptr, count) if ((e = alloc(ptr, (size_t)(count))) != cudaSuccess) { release(); return e; }
    if (inject_failure) { e = cudaErrorMemoryAllocation; release(); return e; }
    if ((e = cudaMalloc((void **)&states, cloud_state_size() * nB)) != cudaSuccess) { release(); return e; }
    A(lim_enc, (size_t)nB * 8);
    A(bitmap, (size_t)nB * bitmap_stride);
    A(vox_cell, (size_t)nB * vcap);
    A(vox_n, (size_t)nB * vcap);
    A(vox_start, (size_t)nB * (vcap + 1));
    A(vox_order, (size_t)nB * vcap);
    A(slot_rank, (size_t)nB * nN);
    A(tile_cnt, (size_t)nB * ntiles_cap * vcap);
    A(hist, (size_t)nB * vcap * (nbins > 0 ? nbins : 1));
    A(point_voxel, (size_t)nB * nN);
    // {x, y, z, label} records; k_stats' bulk copies run up to one 64-record stage past a voxel's end, so the buffer is
    // padded by a stage and starts as zeros (what is read there only has to be finite)
    {
        const size_t sorted_bytes = ((size_t)nB * nN + 128) * 4 * sizeof(double);
        if ((e = cudaMalloc(&sorted, sorted_bytes)) != cudaSuccess) { release(); return e; }
        if ((e = cudaMemset(sorted, 0, sorted_bytes)) != cudaSuccess) { release(); return e; }
    }
    A(mean, (size_t)nB * vcap * 3);
    A(cov, (size_t)nB * vcap * 9);
    A(cov_final, (size_t)nB * vcap * 9);
    A(cls, (size_t)nB * vcap);
    A(kl_div, (size_t)nB * kcap);
    A(kl_flag, (size_t)nB * kcap);
    size_t kpad = 1; while (kpad < kcap) kpad <<= 1;    // k_select pads its sort to a power of two
    A(key, (size_t)nB * kpad);
    A(seq, (size_t)nB * kpad);
    A(firstpos, (size_t)nB * vcap);
    A(removed, (size_t)nB * vcap);
    A(list_div, (size_t)nB * kcap);
    A(list_seq, (size_t)nB * kcap);
    A(recip, (size_t)nN + 128);       // k_stats fetches the reciprocal pairs in stages of 64, up to a stage past the largest count
    if ((e = fill_recip_table(recip, nN + 128)) != cudaSuccess) { release(); return e; }
#undef A
    return cudaSuccess;
}

}  // namespace ndt

// ------------------------------------------------------------------------------------------------
// contexts
// ------------------------------------------------------------------------------------------------
extern "C" const char *ndnet_b200_version(void) { return "ndnet_b200 0.1 (sm_100a)"; }

extern "C" int ndnet_b200_create(ndnet_b200_ctx **out, int device) {
    if (!out) return -200;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return fail(nullptr, e, "cudaGetDeviceCount");
    if (count == 0 || device < 0 || device >= count) {
        fprintf(stderr, "ndnet_b200: no CUDA device %d (found %d); this library has no CPU path\n", device, count);
        return -201;
    }
    ndnet_b200_ctx *c = new (std::nothrow) ndnet_b200_ctx();
    if (!c) return -202;
    c->device = device;
    *out = c;
    return 0;
}

extern "C" void ndnet_b200_destroy(ndnet_b200_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    c->ws.destroy();
    c->mlp_scratch.release();
    void *ptrs[] = {c->d_points, c->d_labels, c->d_feat, c->d_feat64, c->d_olab, c->d_ovox, c->d_info, c->d_logits};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (auto &l : c->lanes) {
        l.ws.destroy(); l.scratch.release();
        void *lp[] = {l.d_points, l.d_labels, l.d_feat, l.d_logits};
        for (void *p : lp) if (p) cudaFree(p);
        if (l.done) cudaEventDestroy(l.done);
        if (l.copied) cudaEventDestroy(l.copied);
        if (l.consumed) cudaEventDestroy(l.consumed);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    if (c->start_ev) cudaEventDestroy(c->start_ev);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto &g : c->graphs) g.destroy();
    if (c->graph_stream) cudaStreamDestroy(c->graph_stream);
    delete c;
}

extern "C" long ndnet_b200_selftest_div(long n, unsigned seed) {
    unsigned long long bad = 0;
    cudaError_t e = ndt::selftest_div(n, seed, &bad);
    if (e != cudaSuccess) return fail(nullptr, e, "selftest_div");
    return (long)bad;
}

extern "C" long ndnet_b200_launch_count(void) { return ndt::launches(); }

extern "C" int ndnet_b200_test_fail_next_reserve(ndnet_b200_ctx *c) {
    if (!c) return -200;
    c->ws.fail_next_reserve = true;
    return 0;
}

extern "C" int ndnet_b200_stage_timing(ndnet_b200_ctx *c, int enable) {
    if (!c) return -200;
    c->ws.timer.enabled = enable != 0;
    for (double &m : c->ws.timer.ms) m = 0;
    c->ws.timer.runs = 0;
    return 0;
}

extern "C" int ndnet_b200_stage_times(ndnet_b200_ctx *c, double *ms, int cap, long *runs) {
    if (!c || !ms) return -200;
    for (int i = 0; i < cap && i < (int)ndt::ST_COUNT; i++) ms[i] = c->ws.timer.ms[i];
    if (runs) *runs = c->ws.timer.runs;
    return (int)ndt::ST_COUNT;
}

extern "C" int ndnet_b200_last_search_passes(ndnet_b200_ctx *c, double *mean_passes, double *mean_evaluations) {
    if (!c || c->ws.last_B <= 0) return -200;
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(c, e, "synchronise");
    double p = 0, v = 0;
    for (int b = 0; b < c->ws.last_B; b++) {
        ndt::CloudSummary cs;
        if ((e = ndt::read_cloud_summary(c->ws, b, &cs)) != cudaSuccess) return fail(c, e, "read state");
        p += cs.passes; v += cs.evals;
    }
    if (mean_passes) *mean_passes = p / c->ws.last_B;
    if (mean_evaluations) *mean_evaluations = v / c->ws.last_B;
    return 0;
}

extern "C" const char *ndnet_b200_last_error(const ndnet_b200_ctx *c) { return c ? c->err.c_str() : "null context"; }

// ------------------------------------------------------------------------------------------------
// batched entry points
// ------------------------------------------------------------------------------------------------
struct ndnet_b200_model { mlp::Model m; };
// ---- CUDA-graph replay of the NDT chain ---------------------------------------------------------
// A captured graph holds addresses, so it runs on buffers the context owns: the caller's scans are copied in (14 bytes per
// point, device to device) and the results copied out around one cudaGraphLaunch.  Worth it only where the chain is bound
// by launch latency: 120 k-point scans, 2 per call: ~55 launches of 2-10 us each.
constexpr long kGraphAutoPoints = 4L << 20;        // automatic mode: batches of at most ~4 M points (32 scans of 120 k)
constexpr size_t kGraphCacheEntries = 4;

static int downsample_graphed(ndnet_b200_ctx *c, const void *points, int dtype, const uint16_t *labels, int B, long N, int num_classes,
                              long D, unsigned flags, float *out_feat, double *out_feat64, uint16_t *out_labels, int32_t *out_voxel,
                              NdtCloudInfo *info, cudaStream_t st) {
    using G = ndnet_b200_ctx::NdtGraph;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return 1; }   // the caller is capturing: its graph takes our kernels
    const unsigned wants = (labels ? 1u : 0u) | (out_feat ? 2u : 0u) | (out_feat64 ? 4u : 0u) | (out_labels ? 8u : 0u) | (out_voxel ? 16u : 0u) | (info ? 32u : 0u);
    const unsigned ws_flags = (c->ws.keep_point_voxels ? 1u : 0u) | (c->ws.keep_kl_list ? 2u : 0u);
    const size_t esz = dtype == 0 ? 4 : 8, lsz = (flags & NDNET_B200_LABELS_U8) ? 1 : 2;
    G *g = nullptr;
    for (auto &e : c->graphs)
        if (e.state != 0 && e.dtype == dtype && e.B == B && e.N == N && e.num_classes == num_classes && e.D == D && e.flags == flags && e.wants == wants &&
            e.ws_flags == ws_flags && e.ws_generation == c->ws.generation) { g = &e; break; }
    cudaError_t e;
    if (!g) {
        // a shape seen for the first time is launched directly and remembered; it gets its graph when it comes back.
        // Graphs of a workspace that no longer exists go first, then the least recently used entry.
        for (auto &x : c->graphs) if (x.state != 0 && x.ws_generation != c->ws.generation) x.destroy();
        G *slot = nullptr;
        for (auto &x : c->graphs) if (x.state == 0) { slot = &x; break; }
        if (!slot && c->graphs.size() < kGraphCacheEntries) { c->graphs.emplace_back(); slot = &c->graphs.back(); }
        if (!slot) { slot = &c->graphs[0]; for (auto &x : c->graphs) if (x.stamp < slot->stamp) slot = &x; slot->destroy(); }
        slot->dtype = dtype; slot->B = B; slot->N = N; slot->num_classes = num_classes; slot->D = D; slot->flags = flags; slot->wants = wants;
        slot->ws_flags = ws_flags; slot->ws_generation = c->ws.generation; slot->state = 1; slot->stamp = ++c->graph_clock;
        return 1;
    }
    g->stamp = ++c->graph_clock;
    if (g->state == 3) return 1;
    if (g->state == 1) {
        G &n = *g;
#define GA(ptr, bytes) if ((e = cudaMalloc((void **)&(ptr), (bytes))) != cudaSuccess) { n.destroy(); cudaGetLastError(); return 1; }
        GA(n.in_pts, (size_t)B * N * 3 * esz);
        if (labels) GA(n.in_lab, (size_t)B * N * lsz);
        if (out_feat) GA(n.o_feat, (size_t)B * D * 12 * 4);
        if (out_feat64) GA(n.o_feat64, (size_t)B * D * 12 * 8);
        if (out_labels) GA(n.o_lab, (size_t)B * D * 2);
        if (out_voxel) GA(n.o_vox, (size_t)B * D * 4);
        if (info) GA(n.o_info, (size_t)B * sizeof(NdtCloudInfo));
#undef GA
        if (!c->graph_stream && (e = cudaStreamCreateWithFlags(&c->graph_stream, cudaStreamNonBlocking)) != cudaSuccess) { n.destroy(); return fail(c, e, "stream create"); }
        // capture on the context's own stream (recording only: nothing executes until the graph is launched).  The chain
        // has run directly at least once on this context, so whatever it creates lazily (side stream, events, function
        // attributes) exists.
        const long before = ndt::launches();
        cudaGraph_t graph = nullptr;
        if ((e = cudaStreamBeginCapture(c->graph_stream, cudaStreamCaptureModeThreadLocal)) == cudaSuccess) {
            const cudaError_t er = ndt::run_batch(c->ws, n.in_pts, dtype, n.in_lab, B, N, num_classes, D, flags, n.o_feat, n.o_feat64, n.o_lab, n.o_vox,
                                                  n.o_info, c->graph_stream);
            e = cudaStreamEndCapture(c->graph_stream, &graph);
            if (er != cudaSuccess) e = er;
        }
        n.launches = ndt::launches() - before;
        ndt::count_launches(-n.launches);              // recorded, not launched
        if (e == cudaSuccess) e = cudaGraphInstantiate(&n.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) {                        // no graph for this shape: this and later calls launch directly
            cudaGetLastError();
            const G key = n;
            n.destroy();
            n.dtype = key.dtype; n.B = key.B; n.N = key.N; n.num_classes = key.num_classes; n.D = key.D; n.flags = key.flags; n.wants = key.wants;
            n.ws_flags = key.ws_flags; n.ws_generation = key.ws_generation; n.stamp = key.stamp; n.state = 3;
            return 1;
        }
        n.state = 2;
    }
    if ((e = cudaMemcpyAsync(g->in_pts, points, (size_t)B * N * 3 * esz, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return fail(c, e, "D2D points");
    if (labels && (e = cudaMemcpyAsync(g->in_lab, labels, (size_t)B * N * lsz, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return fail(c, e, "D2D labels");
    if ((e = cudaGraphLaunch(g->exec, st)) != cudaSuccess) return fail(c, e, "cudaGraphLaunch");
    ndt::count_launches(g->launches);
#define GO(dst, src, bytes) if (dst && (e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return fail(c, e, "D2D results")
    GO(out_feat, g->o_feat, (size_t)B * D * 12 * 4);
    GO(out_feat64, g->o_feat64, (size_t)B * D * 12 * 8);
    GO(out_labels, g->o_lab, (size_t)B * D * 2);
    GO(out_voxel, g->o_vox, (size_t)B * D * 4);
    GO(info, g->o_info, (size_t)B * sizeof(NdtCloudInfo));
#undef GO
    return 0;
}

extern "C" int ndnet_b200_downsample_batch(ndnet_b200_ctx *c, const void *points, int dtype, const uint16_t *labels,
                                           int B, long N, int num_classes, long D, unsigned flags, float *out_feat,
                                           double *out_feat64, uint16_t *out_labels, int32_t *out_voxel,
                                           ndnet_b200_cloud_info *info, void *stream) {
    if (!c || !points || B <= 0 || N < 0 || D <= 0 || (dtype != 0 && dtype != 1) || num_classes < 0) return -200;
    if ((flags & NDNET_B200_LABELS_U8) && num_classes > 255) return -200;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, e, "cudaSetDevice");
    e = c->ws.reserve(B, N, D, num_classes + 1);
    if (e != cudaSuccess) return fail(c, e, "workspace allocation");
    c->ws.last_B = B; c->ws.last_N = N; c->ws.last_D = D;
    c->ws.mark_front = false;
    static const int graph_env = [] { const char *v = getenv("NDNET_B200_NDT_GRAPH"); return v ? atoi(v) : -2; }();
    const int gmode = graph_env >= -1 ? graph_env : c->graph_mode;
    if (N > 0 && !c->ws.timer.enabled && (gmode == 1 || (gmode == -1 && (long)B * N <= kGraphAutoPoints))) {
        const int r = downsample_graphed(c, points, dtype, labels, B, N, num_classes, D, flags, out_feat, out_feat64, out_labels, out_voxel,
                                         (NdtCloudInfo *)info, (cudaStream_t)stream);
        if (r <= 0) return r;          // 0: replayed; < 0: a CUDA error.  1: no graph for this call - launch the kernels directly
    }
    e = ndt::run_batch(c->ws, points, dtype, labels, B, N, num_classes, D, flags, out_feat, out_feat64, out_labels,
                       out_voxel, (NdtCloudInfo *)info, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(c, e, "ndt::run_batch");
    return 0;
}

extern "C" int ndnet_b200_set_ndt_graph(ndnet_b200_ctx *c, int mode) {
    if (!c || mode < -1 || mode > 1) return -200;
    c->graph_mode = mode;
    return 0;
}

// ---- one-hot <-> class tag (ndtnet_preprocessing.py:34,55-57): the reference-facing call hands labels over as one-hot
// rows; one thread per row, neighbouring threads read neighbouring rows (every 128-byte line is used in full from L1)
__global__ void k_onehot_to_label(const float *__restrict__ onehot, long rows, int width, uint16_t *__restrict__ label) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *o = onehot + r * width;
    float best = o[0];
    int bi = 0;
    for (int c = 1; c < width; c++) { const float v = o[c]; if (v > best || (v != v && best == best)) { best = v; bi = c; } }   // first maximum, NaN counts as the largest (torch.argmax)
    label[r] = (uint16_t)bi;
}
__global__ void k_label_to_onehot(const uint16_t *__restrict__ label, long rows, int width, float *__restrict__ onehot) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * width) return;
    onehot[t] = (int)label[t / width] == (int)(t % width) ? 1.f : 0.f;
}

extern "C" int ndnet_b200_onehot_to_labels(const float *onehot, long rows, int width, uint16_t *labels, void *stream) {
    if (!onehot || !labels || rows < 0 || width < 1 || width > 65536) return -200;
    if (rows == 0) return 0;
    ndt::count_launches(1);
    k_onehot_to_label<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(onehot, rows, width, labels);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -100 - (int)e;
}

extern "C" int ndnet_b200_labels_to_onehot(const uint16_t *labels, long rows, int width, float *onehot, void *stream) {
    if (!onehot || !labels || rows < 0 || width < 1) return -200;
    if (rows == 0) return 0;
    ndt::count_launches(1);
    k_label_to_onehot<<<(unsigned)((rows * width + 255) / 256), 256, 0, (cudaStream_t)stream>>>(labels, rows, width, onehot);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -100 - (int)e;
}

extern "C" int ndnet_b200_downsample_batch_host(ndnet_b200_ctx *c, const void *points, int dtype, const uint16_t *labels,
                                                int B, long N, int num_classes, long D, unsigned flags, float *out_feat,
                                                double *out_feat64, uint16_t *out_labels, int32_t *out_voxel,
                                                ndnet_b200_cloud_info *info, void *stream) {
    if (!c || !points || B <= 0 || N < 0 || D <= 0 || (dtype != 0 && dtype != 1)) return -200;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, e, "cudaSetDevice");
    const size_t esz = dtype == 0 ? 4 : 8;
    const size_t pbytes = (size_t)B * N * 3 * esz;
#define G(ptr, have, need) if ((e = grow(ptr, have, need)) != cudaSuccess) return fail(c, e, "staging allocation")
    G(c->d_points, c->d_points_bytes, pbytes);
    const size_t lsz = (flags & NDNET_B200_LABELS_U8) ? 1 : 2;
    if (labels) G(c->d_labels, c->d_labels_bytes, (size_t)B * N * lsz);
    if (out_feat) G(c->d_feat, c->d_feat_bytes, (size_t)B * D * 12 * 4);
    if (out_feat64) G(c->d_feat64, c->d_feat64_bytes, (size_t)B * D * 12 * 8);
    if (out_labels) G(c->d_olab, c->d_olab_bytes, (size_t)B * D * 2);
    if (out_voxel) G(c->d_ovox, c->d_ovox_bytes, (size_t)B * D * 4);
    if (info) G(c->d_info, c->d_info_bytes, (size_t)B * sizeof(NdtCloudInfo));
#undef G
    if ((e = cudaMemcpyAsync(c->d_points, points, pbytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(c, e, "H2D points");
    if (labels && (e = cudaMemcpyAsync(c->d_labels, labels, (size_t)B * N * lsz, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return fail(c, e, "H2D labels");
    int r = ndnet_b200_downsample_batch(c, c->d_points, dtype, labels ? c->d_labels : nullptr, B, N, num_classes, D, flags,
                                        out_feat ? c->d_feat : nullptr, out_feat64 ? c->d_feat64 : nullptr,
                                        out_labels ? c->d_olab : nullptr, out_voxel ? c->d_ovox : nullptr,
                                        info ? (ndnet_b200_cloud_info *)c->d_info : nullptr, stream);
    if (r != 0) return r;
#define D2H(dst, src, bytes) if (dst && (e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(c, e, "D2H")
    D2H(out_feat, c->d_feat, (size_t)B * D * 12 * 4);
    D2H(out_feat64, c->d_feat64, (size_t)B * D * 12 * 8);
    D2H(out_labels, c->d_olab, (size_t)B * D * 2);
    D2H(out_voxel, c->d_ovox, (size_t)B * D * 4);
    D2H(info, c->d_info, (size_t)B * sizeof(NdtCloudInfo));
#undef D2H
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(c, e, "stream synchronise");
    return 0;
}

// ---- pipelined whole-path inference ------------------------------------------------------------
extern "C" int ndnet_b200_set_pipeline(ndnet_b200_ctx *c, int lanes, int chunk) {
    if (!c || lanes < 1 || lanes > 32 || chunk < 1) return -200;
    if (!c->lanes.empty() && (int)c->lanes.size() != lanes) return -203;   // lanes are fixed once created
    c->n_lanes = lanes; c->chunk = chunk;
    return 0;
}

extern "C" int ndnet_b200_set_stagger(ndnet_b200_ctx *c, int on) {
    if (!c) return -200;
    c->stagger = on != 0;
    return 0;
}

extern "C" int ndnet_b200_set_device_chunk(ndnet_b200_ctx *c, int chunk) {
    if (!c || chunk < 1) return -200;
    c->chunk_device = chunk;
    return 0;
}

static int infer_pipelined(ndnet_b200_ctx *c, ndnet_b200_model *model, const void *points, int dtype, const uint16_t *labels,
                           int B, long N, int num_classes, long D, float *out, long out_elems_per_cloud, cudaStream_t user,
                           bool host_io, unsigned label_flags = 0, bool synchronise = true) {
    const size_t lsz = (label_flags & NDNET_B200_LABELS_U8) ? 1 : 2;      // bytes per point label
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, e, "cudaSetDevice");
    if (c->lanes.empty()) {
        c->lanes.resize(c->n_lanes);
        for (auto &l : c->lanes) {
            if ((e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(c, e, "stream create");
            if ((e = cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming)) != cudaSuccess) return fail(c, e, "event create");
            if ((e = cudaEventCreateWithFlags(&l.copied, cudaEventDisableTiming)) != cudaSuccess) return fail(c, e, "event create");
            if ((e = cudaEventCreateWithFlags(&l.consumed, cudaEventDisableTiming)) != cudaSuccess) return fail(c, e, "event create");
        }
        if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(c, e, "stream create");
        if ((e = cudaEventCreateWithFlags(&c->start_ev, cudaEventDisableTiming)) != cudaSuccess) return fail(c, e, "event create");
    }
    const size_t esz = dtype == 0 ? 4 : 8;
    // everything already enqueued on the caller's stream happens before the lanes start
    // Device buffers: what the caller enqueued on its stream (the producer of the scans) comes first.  Host buffers are
    // ready when the call is made, and the library's own buffers are protected by the order of the lanes' streams and the
    // `consumed` events, so a host call does not wait for the caller's stream: consecutive asynchronous calls overlap
    // (the copies of one batch travel while the kernels of the previous one drain).
    if (!host_io && (e = cudaEventRecord(c->start_ev, user)) != cudaSuccess) return fail(c, e, "event record");
    const int L = (int)c->lanes.size();
    // chunk size: host buffers - at most c->chunk, and a small batch is spread over all lanes so that every copy overlaps
    // kernels; device buffers - c->chunk_device (measured on B200, 2048 scans per call: chunks of 512 to 2048 scans run within
    // 1 % of one another, 132-135 k clouds/s; 256: 127 k, 128: 108 k, 64: 92 k - small chunks under-fill the grids)
    int chunk = host_io ? c->chunk : c->chunk_device;
    if (host_io && (B + L - 1) / L < chunk) chunk = (B + L - 1) / L;
    // staggered lanes: a chunk's front (limits, search, voxel assignment - the kernels that stream every point from HBM)
    // starts behind the front of the chunk before it, so that the lanes do not walk through the same stage side by side
    // (four lanes in lock-step share the HBM in the front and leave it idle in the back): the front of chunk k runs beside
    // the statistics / divergences / selection / network of chunk k-1
    static const int stagger_env = [] { const char *v = getenv("NDNET_B200_STAGGER"); return v ? atoi(v) : -1; }();
    const bool stagger = stagger_env >= 0 ? stagger_env != 0 : c->stagger;
    cudaEvent_t prev_front = nullptr;
    int lane_i = 0;
    for (int b0 = 0; b0 < B; b0 += chunk, lane_i = (lane_i + 1) % L) {
        const int nb = B - b0 < chunk ? B - b0 : chunk;
        ndnet_b200_ctx::Lane &l = c->lanes[lane_i];
        if (!host_io && (e = cudaStreamWaitEvent(l.stream, c->start_ev, 0)) != cudaSuccess) return fail(c, e, "stream wait");
        l.ws.mark_front = stagger;
        if (stagger && prev_front && (e = cudaStreamWaitEvent(l.stream, prev_front, 0)) != cudaSuccess) return fail(c, e, "stream wait");
        const char *psrc = (const char *)points + (size_t)b0 * N * 3 * esz;
        const uint16_t *lsrc = labels ? (const uint16_t *)((const char *)labels + (size_t)b0 * N * lsz) : nullptr;
        float *odst = out + (size_t)b0 * out_elems_per_cloud;
        const void *dp = psrc; const uint16_t *dl = lsrc; float *dout = odst;
        if ((e = grow(l.d_feat, l.d_feat_bytes, (size_t)nb * D * 12 * 4)) != cudaSuccess) return fail(c, e, "lane allocation");
        if (host_io) {
            if ((e = grow(l.d_points, l.d_points_bytes, (size_t)nb * N * 3 * esz)) != cudaSuccess) return fail(c, e, "lane allocation");
            if (labels && (e = grow(l.d_labels, l.d_labels_bytes, (size_t)nb * N * lsz)) != cudaSuccess) return fail(c, e, "lane allocation");
            if ((e = grow(l.d_logits, l.d_logits_bytes, (size_t)nb * out_elems_per_cloud * 4)) != cudaSuccess) return fail(c, e, "lane allocation");
            // in chunk order on the copy stream; the lane's buffers are free once its previous kernels have read them
            if (l.used && (e = cudaStreamWaitEvent(c->copy_stream, l.consumed, 0)) != cudaSuccess) return fail(c, e, "stream wait");
            if ((e = cudaMemcpyAsync(l.d_points, psrc, (size_t)nb * N * 3 * esz, cudaMemcpyHostToDevice, c->copy_stream)) != cudaSuccess) return fail(c, e, "H2D points");
            if (labels && (e = cudaMemcpyAsync(l.d_labels, lsrc, (size_t)nb * N * lsz, cudaMemcpyHostToDevice, c->copy_stream)) != cudaSuccess) return fail(c, e, "H2D labels");
            if ((e = cudaEventRecord(l.copied, c->copy_stream)) != cudaSuccess) return fail(c, e, "event record");
            if ((e = cudaStreamWaitEvent(l.stream, l.copied, 0)) != cudaSuccess) return fail(c, e, "stream wait");
            dp = l.d_points; dl = labels ? l.d_labels : nullptr; dout = l.d_logits;
        }
        if ((e = l.ws.reserve(nb, N, D, num_classes + 1)) != cudaSuccess) return fail(c, e, "lane workspace allocation");
        l.ws.last_B = nb; l.ws.last_N = N; l.ws.last_D = D;
        if ((e = ndt::run_batch(l.ws, dp, dtype, dl, nb, N, num_classes, D, NDNET_B200_NAN_TO_NUM | label_flags, l.d_feat, nullptr, nullptr, nullptr,
                                nullptr, l.stream)) != cudaSuccess) return fail(c, e, "ndt::run_batch");
        prev_front = stagger ? l.ws.ev_front : nullptr;
        if (host_io) {       // the NDT kernels are the only readers of the lane's input buffers
            if ((e = cudaEventRecord(l.consumed, l.stream)) != cudaSuccess) return fail(c, e, "event record");
            l.used = true;
        }
        std::string err;
        int r = model->m.forward(l.scratch, l.d_feat, nb, (int)D, dout, l.stream, err);
        if (r != 0) { c->err = err; fprintf(stderr, "ndnet_b200_infer: %s\n", err.c_str()); return r; }
        if (host_io && (e = cudaMemcpyAsync(odst, l.d_logits, (size_t)nb * out_elems_per_cloud * 4, cudaMemcpyDeviceToHost, l.stream)) != cudaSuccess)
            return fail(c, e, "D2H");
    }
    for (auto &l : c->lanes) {
        if ((e = cudaEventRecord(l.done, l.stream)) != cudaSuccess) return fail(c, e, "event record");
        if ((e = cudaStreamWaitEvent(user, l.done, 0)) != cudaSuccess) return fail(c, e, "stream wait");
    }
    if (host_io && synchronise && (e = cudaStreamSynchronize(user)) != cudaSuccess) return fail(c, e, "stream synchronise");
    return 0;
}

// One call for the whole hot path from HOST buffers: H2D of the scans, NDT, network forward, D2H of the
// per-distribution log-probabilities (segmentation) or class probabilities (classification), synchronise.
extern "C" int ndnet_b200_infer_host(ndnet_b200_ctx *c, ndnet_b200_model *model, const void *points, int dtype,
                                     const uint16_t *labels, int B, long N, int num_classes, long D, float *out_host,
                                     long out_elems_per_cloud, void *stream) {
    if (!c || !model || !points || !out_host || B <= 0 || N < 0 || D <= 0 || (dtype != 0 && dtype != 1)) return -200;
    return infer_pipelined(c, model, points, dtype, labels, B, N, num_classes, D, out_host, out_elems_per_cloud, (cudaStream_t)stream, true);
}

// The same call with one BYTE per point label (num_classes <= 255): 13 instead of 14 bytes per point cross PCIe.
extern "C" int ndnet_b200_infer_host_u8(ndnet_b200_ctx *c, ndnet_b200_model *model, const void *points, int dtype,
                                        const uint8_t *labels, int B, long N, int num_classes, long D, float *out_host,
                                        long out_elems_per_cloud, void *stream) {
    if (!c || !model || !points || !out_host || B <= 0 || N < 0 || D <= 0 || (dtype != 0 && dtype != 1) || num_classes > 255) return -200;
    return infer_pipelined(c, model, points, dtype, (const uint16_t *)labels, B, N, num_classes, D, out_host, out_elems_per_cloud,
                           (cudaStream_t)stream, true, NDNET_B200_LABELS_U8);
}

// Asynchronous form of the host-buffer call: returns once everything is enqueued; the results are in `out_host` after the
// caller's stream has been synchronised (cudaStreamSynchronize / ndnet_b200_infer_wait).  labels_u8 != 0: one byte per label.
extern "C" int ndnet_b200_infer_host_async(ndnet_b200_ctx *c, ndnet_b200_model *model, const void *points, int dtype,
                                           const void *labels, int labels_u8, int B, long N, int num_classes, long D,
                                           float *out_host, long out_elems_per_cloud, void *stream) {
    if (!c || !model || !points || !out_host || B <= 0 || N < 0 || D <= 0 || (dtype != 0 && dtype != 1)) return -200;
    if (labels_u8 && num_classes > 255) return -200;
    return infer_pipelined(c, model, points, dtype, (const uint16_t *)labels, B, N, num_classes, D, out_host, out_elems_per_cloud,
                           (cudaStream_t)stream, true, labels_u8 ? NDNET_B200_LABELS_U8 : 0u, false);
}

extern "C" int ndnet_b200_infer_wait(ndnet_b200_ctx *c, void *stream) {
    if (!c) return -200;
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(c, e, "stream synchronise");
}

// Same from/to DEVICE buffers; asynchronous: on return the caller's stream waits for the result.
extern "C" int ndnet_b200_infer_device(ndnet_b200_ctx *c, ndnet_b200_model *model, const void *points, int dtype,
                                       const uint16_t *labels, int B, long N, int num_classes, long D, float *out_dev,
                                       long out_elems_per_cloud, void *stream) {
    if (!c || !model || !points || !out_dev || B <= 0 || N < 0 || D <= 0 || (dtype != 0 && dtype != 1)) return -200;
    return infer_pipelined(c, model, points, dtype, labels, B, N, num_classes, D, out_dev, out_elems_per_cloud, (cudaStream_t)stream, false);
}

extern "C" int ndnet_b200_keep_point_voxels(ndnet_b200_ctx *c, int enable) {
    if (!c) return -200;
    c->ws.keep_point_voxels = enable != 0;
    return 0;
}

extern "C" int ndnet_b200_last_point_voxels(ndnet_b200_ctx *c, int32_t *out_dev, void *stream) {
    if (!c || !out_dev || c->ws.last_B == 0) return -200;
    if (!c->ws.keep_point_voxels) { c->err = "enable ndnet_b200_keep_point_voxels before the batch"; return -204; }
    cudaError_t e = cudaMemcpyAsync(out_dev, c->ws.point_voxel, (size_t)c->ws.last_B * c->ws.last_N * 4,
                                    cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail(c, e, "last_point_voxels");
}

extern "C" int ndnet_b200_keep_kl_list(ndnet_b200_ctx *c, int enable) {
    if (!c) return -200;
    c->ws.keep_kl_list = enable != 0;
    return 0;
}

extern "C" long ndnet_b200_last_kl_list(ndnet_b200_ctx *c, int b, double *div, int32_t *p_voxel, int32_t *q_voxel, long cap) {
    if (!c || b < 0 || b >= c->ws.last_B) return -200;
    if (!c->ws.keep_kl_list) { c->err = "enable ndnet_b200_keep_kl_list before the batch"; return -204; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(c, e, "synchronise");
    ndt::CloudSummary cs;
    if ((e = ndt::read_cloud_summary(c->ws, b, &cs)) != cudaSuccess) return fail(c, e, "read state");
    if (cs.status != 0) return 0;
    const unsigned K = cs.K, V = cs.V;
    const int *len = cs.len;
    const size_t kcap = (size_t)c->ws.vcap * 6;
    std::vector<double> d(K);
    std::vector<unsigned> s(K), cell(V);
    if (K) {
        if ((e = cudaMemcpy(d.data(), c->ws.list_div + (size_t)b * kcap, K * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return fail(c, e, "D2H list");
        if ((e = cudaMemcpy(s.data(), c->ws.list_seq + (size_t)b * kcap, K * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return fail(c, e, "D2H list");
    }
    if (V && (e = cudaMemcpy(cell.data(), c->ws.vox_cell + (size_t)b * c->ws.vcap, V * 4, cudaMemcpyDeviceToHost)) != cudaSuccess)
        return fail(c, e, "D2H cells");
    static const int dx[6] = {1, -1, 0, 0, 0, 0}, dy[6] = {0, 0, 1, -1, 0, 0}, dz[6] = {0, 0, 0, 0, 1, -1};
    for (unsigned i = 0; i < K && (long)i < cap; i++) {
        const unsigned slot = s[i] / 6, dir = s[i] % 6;
        const unsigned pc = cell[slot];
        const int lx = len[0], ly = len[1];
        const int z = pc / (lx * ly), y = (pc % (lx * ly)) / lx, x = pc % lx;
        if (div) div[i] = d[i];
        if (p_voxel) p_voxel[i] = (int32_t)pc;
        if (q_voxel) q_voxel[i] = (int32_t)((z + dz[dir]) * lx * ly + (y + dy[dir]) * lx + (x + dx[dir]));
    }
    return (long)K;
}

// ------------------------------------------------------------------------------------------------
// legacy ABI (host pointers; signatures of core_legacy/include/ndnet_core/ndt.h)
// ------------------------------------------------------------------------------------------------
namespace {

constexpr uint32_t kMagicNd = 0x4e444e44u, kMagicKl = 0x4b4c4b4cu;

struct Session;
struct NdToken { uint32_t magic; Session *owner; };
struct KlToken { uint32_t magic; Session *owner; };

// Retained state of one ndt_downsample call (what the reference keeps in nd_array / kl_divergences),
// device resident so that prune_nds / to_point_cloud continue on the GPU.
struct Session {
    NdToken nd{kMagicNd, nullptr};
    KlToken kl{kMagicKl, nullptr};
    bool nd_freed = false, kl_freed = false;
    int device = 0;
    ndt::SessionState st;
};

std::mutex g_mutex;
ndnet_b200_ctx *g_ctx = nullptr;

ndnet_b200_ctx *default_ctx() {
    if (!g_ctx) {
        int dev = 0;
        if (const char *e = getenv("NDNET_B200_DEVICE")) dev = atoi(e);
        if (ndnet_b200_create(&g_ctx, dev) != 0) g_ctx = nullptr;
        else g_ctx->ws.keep_kl_list = true;           // the legacy handles carry the whole sorted list (prune() walks on from it)
    }
    return g_ctx;
}

void maybe_delete(Session *s) {
    if (s->nd_freed && s->kl_freed) { cudaSetDevice(s->device); s->st.release(); delete s; }
}

}  // namespace

extern "C" int ndt_downsample(double *point_cloud, unsigned short point_dim, unsigned long num_points,
                              unsigned int *len_x, unsigned int *len_y, unsigned int *len_z, double *offset_x,
                              double *offset_y, double *offset_z, double *voxel_size, unsigned short *classes,
                              unsigned short num_classes, unsigned long num_desired_points,
                              double *downsampled_point_cloud, unsigned long *num_downsampled_points,
                              double *covariances, unsigned short *downsampled_classes,
                              struct normal_distribution_t **nd_array, unsigned long *num_valid_nds,
                              struct kl_divergence_t **kl_divergences, unsigned long *num_kl_divergences) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (nd_array) *nd_array = nullptr;
    ndnet_b200_ctx *c = default_ctx();
    if (!c) { fprintf(stderr, "ndt_downsample: no CUDA device available; libndnet_b200 has no CPU path\n"); return -201; }
    if (!point_cloud || num_desired_points == 0) return -200;
    const long N = (long)num_points, D = (long)num_desired_points;
    // the reference reads 3 doubles per point with a hard-coded stride of 3 (normal_distributions.c:47)
    // but takes the limits with stride point_dim (pointclouds.c:55-61); only point_dim == 3 is consistent.
    if (point_dim != 3) { fprintf(stderr, "ndt_downsample: point_dim must be 3\n"); return -200; }
    std::vector<double> feat((size_t)D * 12);
    std::vector<uint16_t> lab((size_t)D);
    ndnet_b200_cloud_info info;
    int r = ndnet_b200_downsample_batch_host(c, point_cloud, NDNET_B200_F64, classes, 1, N, num_classes, D, 0, nullptr,
                                             feat.data(), classes ? lab.data() : nullptr, nullptr, &info, nullptr);
    if (r != 0) return r;
    if (len_x) *len_x = info.len[0];
    if (len_y) *len_y = info.len[1];
    if (len_z) *len_z = info.len[2];
    if (offset_x) *offset_x = info.offset[0];
    if (offset_y) *offset_y = info.offset[1];
    if (offset_z) *offset_z = info.offset[2];
    if (voxel_size) *voxel_size = info.voxel_size;
    if (info.status != 0) {
        if (info.status == -3) fprintf(stderr, "Reached maximum number of iterations!\n");          // ndt.c:192
        else if (info.status == -5) fprintf(stderr, "ndt_downsample: the point cloud holds NaN coordinates; refused\n");
        else fprintf(stderr, "Error allocating memory for normal distributions: grid too large\n");   // ndt.c:153
        return info.status;
    }
    for (uint32_t i = 0; i < info.num_out; i++) {
        if (downsampled_point_cloud) memcpy(downsampled_point_cloud + (size_t)i * 3, feat.data() + (size_t)i * 12, 3 * sizeof(double));
        if (covariances) memcpy(covariances + (size_t)i * 9, feat.data() + (size_t)i * 12 + 3, 9 * sizeof(double));
        if (downsampled_classes && classes) downsampled_classes[i] = lab[i];
    }
    if (num_downsampled_points) *num_downsampled_points = info.num_out;
    if (num_valid_nds) *num_valid_nds = info.num_valid;
    if (num_kl_divergences) *num_kl_divergences = info.num_kl_after;
    if (info.prune_status == -2) fprintf(stderr, "Reached the end of the divergences array!\n");     // ndt.c:54
    // retain the state for prune_nds / to_point_cloud
    Session *s = new (std::nothrow) Session();
    if (!s) return -202;
    s->nd.owner = s; s->kl.owner = s; s->device = c->device;
    cudaError_t e = s->st.capture(c->ws, 0, classes != nullptr);
    if (e != cudaSuccess) { s->st.release(); delete s; return fail(c, e, "session capture"); }
    if (nd_array) *nd_array = (struct normal_distribution_t *)&s->nd; else s->nd_freed = true;
    if (kl_divergences) *kl_divergences = (struct kl_divergence_t *)&s->kl; else s->kl_freed = true;
    if (s->nd_freed && s->kl_freed) maybe_delete(s);
    return 0;
}

extern "C" int prune_nds(struct normal_distribution_t *nd_array, unsigned int len_x, unsigned int len_y, unsigned int len_z,
                         unsigned long num_desired_nds, unsigned long *num_valid_nds,
                         struct kl_divergence_t *kl_divergences, unsigned long *num_kl_divergences) {
    (void)len_x; (void)len_y; (void)len_z;
    std::lock_guard<std::mutex> lock(g_mutex);
    NdToken *t = (NdToken *)nd_array;
    KlToken *k = (KlToken *)kl_divergences;
    if (!t || t->magic != kMagicNd || !k || k->magic != kMagicKl || k->owner != t->owner) {
        fprintf(stderr, "prune_nds: handles were not produced by this library's ndt_downsample\n");
        return -200;
    }
    Session *s = t->owner;
    // the guard of ndt.c:36-39 on the library's own count: the caller's pointer may be NULL or stale, and an unsigned
    // to_remove = valid - desired must never wrap
    if (num_desired_nds > (unsigned long)s->st.n_valid || (num_valid_nds && num_desired_nds > *num_valid_nds)) {
        fprintf(stderr, "Number of desired normal distributions is greater than the number valid distributions!\n");  // ndt.c:37
        return -1;
    }
    unsigned valid = 0, nkl = 0; int ret = 0;
    cudaError_t e = cudaSetDevice(s->device);      // the session lives on the device that ran its ndt_downsample
    if (e != cudaSuccess) return fail(g_ctx, e, "cudaSetDevice");
    e = s->st.prune((unsigned long)num_desired_nds, &valid, &nkl, &ret);
    if (e != cudaSuccess) return fail(g_ctx, e, "session prune");
    if (num_valid_nds) *num_valid_nds = valid;
    if (num_kl_divergences) *num_kl_divergences = nkl;
    if (ret == -2) fprintf(stderr, "Reached the end of the divergences array!\n");
    return ret;
}

extern "C" int to_point_cloud(struct normal_distribution_t *nd_array, unsigned int len_x, unsigned int len_y, unsigned int len_z,
                              double offset_x, double offset_y, double offset_z, double voxel_size, double *point_cloud,
                              unsigned long *num_points, double *covariances, unsigned short *classes) {
    (void)len_x; (void)len_y; (void)len_z; (void)offset_x; (void)offset_y; (void)offset_z; (void)voxel_size;
    std::lock_guard<std::mutex> lock(g_mutex);
    NdToken *t = (NdToken *)nd_array;
    if (!t || t->magic != kMagicNd) {
        fprintf(stderr, "to_point_cloud: handle was not produced by this library's ndt_downsample\n");
        return -200;
    }
    Session *s = t->owner;
    std::vector<double> feat;
    std::vector<uint16_t> lab;
    unsigned rows = 0;
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return fail(g_ctx, e, "cudaSetDevice");
    e = s->st.output(feat, lab, &rows);
    if (e != cudaSuccess) return fail(g_ctx, e, "session output");
    // The reference writes every surviving row; callers size buffers for the number they asked for,
    // which equals the survivor count unless a walk stopped early (A15).
    for (unsigned i = 0; i < rows; i++) {
        if (point_cloud) memcpy(point_cloud + (size_t)i * 3, feat.data() + (size_t)i * 12, 3 * sizeof(double));
        if (covariances) memcpy(covariances + (size_t)i * 9, feat.data() + (size_t)i * 12 + 3, 9 * sizeof(double));
        if (classes && s->st.has_labels) classes[i] = lab[i];
    }
    if (num_points) *num_points = rows;
    return 0;
}

extern "C" void free_nds(struct normal_distribution_t *nd_array, unsigned long num_nds) {
    (void)num_nds;
    std::lock_guard<std::mutex> lock(g_mutex);
    NdToken *t = (NdToken *)nd_array;
    if (!t || t->magic != kMagicNd) return;     // NULL after a failed downsample (A16): never crash
    Session *s = t->owner;
    if (s->nd_freed) return;
    s->nd_freed = true;
    t->magic = 0;
    maybe_delete(s);
}

extern "C" void free_kl_divergences(struct kl_divergence_t *kl_divergences) {
    std::lock_guard<std::mutex> lock(g_mutex);
    KlToken *t = (KlToken *)kl_divergences;
    if (!t || t->magic != kMagicKl) return;
    Session *s = t->owner;
    if (s->kl_freed) return;
    s->kl_freed = true;
    t->magic = 0;
    maybe_delete(s);
}

extern "C" void print_matrix(double *matrix, int rows, int cols) {
    // core_legacy/src/matrix.c:28-35
    for (int i = 0; i < rows; i++) {
        for (int j = 0; j < cols; j++) printf("%f ", matrix[i * cols + j]);
        printf("\n");
    }
}

// ------------------------------------------------------------------------------------------------
// model entry points: thin shims over mlp.cu
// ------------------------------------------------------------------------------------------------
extern "C" int ndnet_b200_model_create(ndnet_b200_ctx *c, ndnet_b200_model **model, int kind, int n_tensors,
                                       const char *const *names, const float *const *data,
                                       const int64_t *const *shapes, const int *ndims) {
    if (!c || !model || n_tensors <= 0 || !names || !data || !shapes || !ndims) return -200;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, e, "cudaSetDevice");
    ndnet_b200_model *m = new (std::nothrow) ndnet_b200_model();
    if (!m) return -202;
    std::string err;
    int r = m->m.build(kind, n_tensors, names, data, shapes, ndims, err);
    if (r != 0) { c->err = err; fprintf(stderr, "ndnet_b200_model_create: %s\n", err.c_str()); m->m.release(); delete m; return r; }
    *model = m;
    return 0;
}

extern "C" int ndnet_b200_model_input_dim(const ndnet_b200_model *m) { return m ? m->m.input_dim() : -200; }

extern "C" int ndnet_b200_model_set_fused_head(ndnet_b200_model *m, int enable) {
    if (!m || (enable != 0 && enable != 1)) return -200;
    m->m.set_fused_head(enable != 0);
    return 0;
}

extern "C" void ndnet_b200_model_destroy(ndnet_b200_model *m) {
    if (!m) return;
    m->m.release();
    delete m;
}

extern "C" long ndnet_b200_model_tap(ndnet_b200_ctx *c, ndnet_b200_model *m, const char *name, float *out, long cap, void *stream) {
    if (!c || !m) return -200;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, e, "cudaSetDevice");
    return m->m.tap(name, out, cap, (cudaStream_t)stream);
}

extern "C" int ndnet_b200_model_forward(ndnet_b200_ctx *c, ndnet_b200_model *m, const float *feat, int B, int D, float *out,
                                        void *stream) {
    if (!c || !m || !feat || !out || B <= 0 || D <= 0) return -200;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, e, "cudaSetDevice");
    std::string err;
    int r = m->m.forward(c->mlp_scratch, feat, B, D, out, (cudaStream_t)stream, err);
    if (r != 0) { c->err = err; fprintf(stderr, "ndnet_b200_model_forward: %s\n", err.c_str()); }
    return r;
}
