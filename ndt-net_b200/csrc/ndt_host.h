// ndt_host.h — host-side declarations shared by the kernels and the C ABI.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

namespace ndt {

struct CloudState;

// number of kernels this library launched since load (all contexts); bench.py reports the delta
void count_launches(long n);
long launches();

// optional per-stage CUDA-event timing of run_batch (bench.py's roofline leg)
enum Stage { ST_LIMITS = 0, ST_SEARCH, ST_RANK, ST_OFFSETS, ST_SCATTER, ST_STATS, ST_KL, ST_SELECT, ST_COUNT };
struct StageTimer {
    bool enabled = false;
    cudaEvent_t ev[ST_COUNT + 1] = {};
    bool created = false;
    double ms[ST_COUNT] = {};
    long runs = 0;
    int search_launches = 0;
    void mark(int i, cudaStream_t st) { if (enabled) cudaEventRecord(ev[i], st); }
};

// must match ndnet_b200_cloud_info in include/ndnet_b200.h
struct NdtCloudInfo {
    int32_t status, prune_status, evaluations;
    uint32_t len[3];
    uint32_t num_voxels, num_valid, num_kl, num_kl_after, num_out, num_survivors;
    double voxel_size;
    double offset[3];
    double limits[6];
};

// Device workspace for a batch of up to B_cap clouds x N_cap points, D_cap desired distributions.
struct Workspace {
    int B_cap = 0; long N_cap = 0; long D_cap = 0; int bins_cap = 0;
    unsigned long generation = 0;  // bumped whenever the arrays are freed (captured graphs hold their addresses)
    unsigned vcap = 0;             // max occupied voxels per cloud = floor(1.2*D)+2
    size_t bitmap_stride = 0;      // uint2 {bits, prefix} words per cloud
    int ntiles_cap = 0;
    CloudState *states = nullptr;
    unsigned long long *lim_enc = nullptr;
    uint2 *bitmap = nullptr;
    unsigned *vox_cell = nullptr, *vox_n = nullptr, *vox_start = nullptr, *vox_order = nullptr;
    unsigned *slot_rank = nullptr, *tile_cnt = nullptr, *hist = nullptr;
    int *point_voxel = nullptr;
    void *sorted = nullptr;
    bool keep_point_voxels = false;   // inspection only: k_rank also records every point's voxel id
    bool keep_kl_list = false;        // the whole sorted divergence list of every cloud stays behind (legacy handles, inspection)
    double *mean = nullptr, *cov = nullptr, *cov_final = nullptr;
    double2 *recip = nullptr;         // {RN(1/c), residual} for c = 1..N_cap (k_stats' divisions by the running count)
    uint16_t *cls = nullptr;
    double *kl_div = nullptr; unsigned char *kl_flag = nullptr;
    unsigned long long *key = nullptr; unsigned *seq = nullptr;
    unsigned *firstpos = nullptr; unsigned char *removed = nullptr;
    double *list_div = nullptr; unsigned *list_seq = nullptr;
    StageTimer timer;
    cudaStream_t side = nullptr;       // side stream for kernels that are independent of the main chain
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // recorded behind k_scatter when mark_front is set: the bandwidth-bound half of the chain (limits, search, voxel
    // assignment) is enqueued up to here; the pipelined calls start the next chunk's front behind it (capi.cu)
    cudaEvent_t ev_front = nullptr;
    bool mark_front = false;
    // of the last run
    int last_B = 0; long last_N = 0; long last_D = 0;

    bool fail_next_reserve = false;    // test hook (ndnet_b200_test_fail_next_reserve): the next growing reserve() fails

    cudaError_t reserve(int B, long N, long D, int bins);
    void release();                    // frees the arrays, zeroes the capacities
    void destroy();                    // release() + the side stream and its events
};

cudaError_t run_batch(Workspace &w, const void *pts, int dtype, const uint16_t *labels, int B, long N, int num_classes,
                      long D, unsigned flags, float *out_feat, double *out_feat64, uint16_t *out_labels, int *out_voxel,
                      NdtCloudInfo *info, cudaStream_t st);

size_t cloud_state_size();
cudaError_t selftest_div(long n, unsigned seed, unsigned long long *mismatches_host);
cudaError_t fill_recip_table(double2 *tab, long n);

// host copy of the scalar results of one cloud of the last batch
struct CloudSummary { int status; int len[3]; unsigned V, K, n_valid, walk; int prune_ret; int passes, evals; };
cudaError_t read_cloud_summary(const Workspace &w, int b, CloudSummary *out);

// Device-resident state one legacy ndt_downsample call retains for prune_nds / to_point_cloud
// (the reference keeps it in the nd_array / kl_divergences allocations, ndt.c:151,198).
struct SessionState {
    unsigned V = 0, K = 0, start = 0;      // start = first list entry still in the list
    unsigned n_valid = 0, num_kl = 0;      // counters as the reference reports them
    bool has_labels = false;
    unsigned *vox_cell = nullptr; unsigned char *removed = nullptr;
    double *mean = nullptr, *cov_final = nullptr; uint16_t *cls = nullptr;
    unsigned *list_seq = nullptr; unsigned *firstpos = nullptr;
    unsigned *d_result = nullptr;          // 4 words: n_removed, new start, ret flag, rows
    double *d_feat = nullptr; uint16_t *d_lab = nullptr;

    cudaError_t capture(const Workspace &w, int b, bool labels);
    cudaError_t prune(unsigned long desired, unsigned *valid, unsigned *nkl, int *ret);
    cudaError_t output(std::vector<double> &feat, std::vector<uint16_t> &lab, unsigned *rows);
    void release();
};

}  // namespace ndt
