// mlp.cu — placeholder until the tcgen05 forward lands (next commit): fails loudly.
#include "mlp_host.h"
namespace mlp {
cudaError_t Scratch::reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (buf) cudaFree(buf);
    buf = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&buf, need);
    if (e == cudaSuccess) bytes = need;
    return e;
}
void Scratch::release() { if (buf) cudaFree(buf); buf = nullptr; bytes = 0; }
int Model::build(int, int, const char *const *, const float *const *, const int64_t *const *, const int *, std::string &err) {
    err = "model forward not built yet"; return -300;
}
int Model::forward(Scratch &, const float *, int, int, float *, cudaStream_t, std::string &err) {
    err = "model forward not built yet"; return -300;
}
void Model::release() {}
}  // namespace mlp
