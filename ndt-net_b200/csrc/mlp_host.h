// mlp_host.h — host-side interface of the PointNet / NDT-Net forward (mlp.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <string>

namespace mlp {

struct Scratch {
    void *buf = nullptr; size_t bytes = 0;
    cudaError_t reserve(size_t need);
    void release();
};

struct ModelImpl;

struct Model {
    ModelImpl *impl = nullptr;
    int build(int kind, int n_tensors, const char *const *names, const float *const *data, const int64_t *const *shapes,
              const int *ndims, std::string &err);
    int forward(Scratch &scratch, const float *feat, int B, int D, float *out, cudaStream_t st, std::string &err);
    void release();
    int input_dim() const;
    void set_fused_head(bool on);    // segmentation head layers 1 + 2 as one kernel (default) or as two GEMMs
    // inspection: copy one internal activation of the LAST forward (still in its scratch buffer) to `out` as fp32;
    // returns the element count (out == nullptr: query only) or a negative error
    long tap(const char *name, float *out, long cap, cudaStream_t st);
};

}  // namespace mlp
