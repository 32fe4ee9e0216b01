// ndt_session.cu — state retained by the legacy ABI between ndt_downsample, prune_nds and
// to_point_cloud (ndnet/preprocessing/ndt_legacy.py:173-240 -> core_legacy/src/ndt.c:28-117), kept on
// the device.  The continuation is defined on the WELL-FORMED list: the reference shifts its list by
// the walk length but keeps a larger count, so after any skipped entry it reads undefined memory
// (SURVEY.md A15); here the list after a walk is exactly the entries behind the walk.
#include "ndt_host.h"
#include "ndt_device.cuh"

namespace ndt {

size_t cloud_state_size() { return sizeof(CloudState); }

static long g_launches = 0;
void count_launches(long n) { __atomic_fetch_add(&g_launches, n, __ATOMIC_RELAXED); }
long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

cudaError_t read_cloud_summary(const Workspace &w, int b, CloudSummary *out) {
    CloudState s;
    cudaError_t e = cudaMemcpy(&s, (const char *)w.states + (size_t)b * sizeof(CloudState), sizeof(s), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    out->status = s.status;
    for (int a = 0; a < 3; a++) out->len[a] = s.len[a];
    out->V = s.V; out->K = s.K; out->n_valid = s.n_valid; out->walk = s.walk; out->prune_ret = s.prune_ret;
    out->passes = s.passes; out->evals = s.evals;
    return cudaSuccess;
}

__device__ __forceinline__ unsigned scan1024(unsigned val, unsigned *s_warp, unsigned &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = val;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    __syncthreads();
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    unsigned woff = 0; total = 0;
    for (int k = 0; k < 32; k++) { const unsigned x = s_warp[k]; if (k < wid) woff += x; total += x; }
    return woff + inc - val;
}

// prune_nds (ndt.c:45-67) on the retained list.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) k_session_prune(unsigned V, unsigned K, unsigned start, unsigned to_remove,
                                                        const unsigned *__restrict__ list_seq, unsigned *__restrict__ firstpos,
                                                        unsigned char *__restrict__ removed, unsigned *__restrict__ result) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (unsigned v = tid; v < V; v += blockDim.x) firstpos[v] = 0xFFFFFFFFu;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (unsigned i = start + tid; i < K; i += blockDim.x) {
        const unsigned p = list_seq[i] / kDirs;
        if (!removed[p]) atomicMin(&firstpos[p], i);
    }
    __syncthreads();
    const unsigned L = K - start;
    unsigned walk_local = 0, n_local = 0;
    for (unsigned base = start; base < K && s_carry < to_remove; base += blockDim.x) {
        const unsigned i = base + tid;
        const bool first = i < K && firstpos[list_seq[i] / kDirs] == i;
        unsigned total;
        const unsigned r = s_carry + scan1024(first ? 1u : 0u, s_warp, total);
        if (first && r < to_remove && (unsigned long)(i - start) + r < (unsigned long)L) {
            removed[list_seq[i] / kDirs] = 1;
            walk_local = i + 1;
            n_local++;
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    unsigned total;
    __syncthreads();
    scan1024(n_local, s_warp, total);
    unsigned wm = walk_local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned u = __shfl_xor_sync(0xffffffffu, wm, o); wm = u > wm ? u : wm; }
    __syncthreads();
    if (lane == 0) s_warp[wid] = wm;
    __syncthreads();
    if (tid == 0) {
        unsigned m = start;
        for (int k = 0; k < 32; k++) m = s_warp[k] > m ? s_warp[k] : m;
        result[0] = total;                         // removed in this call
        result[1] = total < to_remove ? K : m;     // list start afterwards
        result[2] = total < to_remove ? 1u : 0u;   // 1 -> the reference's -2
    }
}

// to_point_cloud (ndt.c:75-117) on the retained state.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) k_session_output(unsigned V, const unsigned char *__restrict__ removed,
                                                         const double *__restrict__ mean, const double *__restrict__ cov_final,
                                                         const uint16_t *__restrict__ cls, double *__restrict__ feat,
                                                         uint16_t *__restrict__ lab, unsigned *__restrict__ result) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const int tid = threadIdx.x;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < V; base += blockDim.x) {
        const unsigned v = base + tid;
        const bool alive = v < V && !removed[v];
        unsigned total;
        const unsigned row = s_carry + scan1024(alive ? 1u : 0u, s_warp, total);
        if (alive) {
            for (int k = 0; k < 3; k++) feat[(size_t)row * 12 + k] = mean[(size_t)v * 3 + k];
            for (int k = 0; k < 9; k++) feat[(size_t)row * 12 + 3 + k] = cov_final[(size_t)v * 9 + k];
            lab[row] = cls ? cls[v] : (uint16_t)0;
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (tid == 0) result[3] = s_carry;
}

void SessionState::release() {
    void *ptrs[] = {vox_cell, removed, mean, cov_final, cls, list_seq, firstpos, d_result, d_feat, d_lab};
    for (void *p : ptrs) if (p) cudaFree(p);
    *this = SessionState();
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t SessionState::capture(const Workspace &w, int b, bool labels) {
    CloudSummary cs;
    CK(read_cloud_summary(w, b, &cs));
    V = cs.V; K = cs.K; n_valid = cs.n_valid; has_labels = labels;
    num_kl = K - (V - n_valid);
    start = cs.prune_ret == 0 ? cs.walk : K;
    const size_t kcap = (size_t)w.vcap * kDirs;
    const size_t v1 = V ? V : 1, k1 = K ? K : 1;
    CK(cudaMalloc((void **)&vox_cell, v1 * 4)); CK(cudaMalloc((void **)&removed, v1)); CK(cudaMalloc((void **)&mean, v1 * 24));
    CK(cudaMalloc((void **)&cov_final, v1 * 72)); CK(cudaMalloc((void **)&cls, v1 * 2)); CK(cudaMalloc((void **)&list_seq, k1 * 4));
    CK(cudaMalloc((void **)&firstpos, v1 * 4)); CK(cudaMalloc((void **)&d_result, 16));
    CK(cudaMalloc((void **)&d_feat, v1 * 96)); CK(cudaMalloc((void **)&d_lab, v1 * 2));
    CK(cudaMemcpy(vox_cell, w.vox_cell + (size_t)b * w.vcap, (size_t)V * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(removed, w.removed + (size_t)b * w.vcap, (size_t)V, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(mean, w.mean + (size_t)b * w.vcap * 3, (size_t)V * 24, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(cov_final, w.cov_final + (size_t)b * w.vcap * 9, (size_t)V * 72, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(cls, w.cls + (size_t)b * w.vcap, (size_t)V * 2, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(list_seq, w.list_seq + (size_t)b * kcap, (size_t)K * 4, cudaMemcpyDeviceToDevice));
    return cudaSuccess;
}

cudaError_t SessionState::prune(unsigned long desired, unsigned *valid, unsigned *nkl, int *ret) {
    const unsigned to_remove = (unsigned)((unsigned long)n_valid - desired);
    k_session_prune<<<1, 1024>>>(V, K, start, to_remove, list_seq, firstpos, removed, d_result);
    unsigned res[4];
    CK(cudaMemcpy(res, d_result, sizeof(res), cudaMemcpyDeviceToHost));
    n_valid -= res[0];
    num_kl -= res[0];
    start = res[1];
    *valid = n_valid; *nkl = num_kl; *ret = res[2] ? -2 : 0;
    return cudaGetLastError();
}

cudaError_t SessionState::output(std::vector<double> &feat, std::vector<uint16_t> &lab, unsigned *rows) {
    k_session_output<<<1, 1024>>>(V, removed, mean, cov_final, has_labels ? cls : nullptr, d_feat, d_lab, d_result);
    unsigned res[4];
    CK(cudaMemcpy(res, d_result, sizeof(res), cudaMemcpyDeviceToHost));
    *rows = res[3];
    feat.resize((size_t)res[3] * 12); lab.resize(res[3]);
    if (res[3]) {
        CK(cudaMemcpy(feat.data(), d_feat, (size_t)res[3] * 96, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(lab.data(), d_lab, (size_t)res[3] * 2, cudaMemcpyDeviceToHost));
    }
    return cudaGetLastError();
}

}  // namespace ndt
