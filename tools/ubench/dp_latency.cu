// Microbenchmark: dependent-issue latency of fp64 ops on this GPU (one warp, one CTA).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double *out, long long *cyc, double a, double b, int iters) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (OP == 0) x = x + b;
            else if (OP == 1) x = x * b;
            else if (OP == 2) x = fma(x, b, a);
            else if (OP == 3) x = x / b;
            else if (OP == 4) x = (double)((float)x * 1.0001f);
            else if (OP == 5) { float f = (float)x; f = f * 1.0001f + 0.5f; x = (double)f; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; cyc[0] = t1 - t0; }
}
int main() {
    double *o; long long *c; cudaMalloc(&o, 8); cudaMalloc(&c, 8);
    const char *names[] = {"DADD", "DMUL", "DFMA", "DDIV", "F2F+FMUL+F2F", "F2F FFMA F2F"};
    for (int op = 0; op < 6; op++) {
        for (int threads = 32; threads <= 32; threads *= 4) {
            long long h = 0; int iters = 2000;
            for (int rep = 0; rep < 2; rep++) {
                switch (op) {
                    case 0: k<0><<<1, threads>>>(o, c, 1.0, 1e-9, iters); break;
                    case 1: k<1><<<1, threads>>>(o, c, 1.0, 1.0000001, iters); break;
                    case 2: k<2><<<1, threads>>>(o, c, 1.0, 0.999, iters); break;
                    case 3: k<3><<<1, threads>>>(o, c, 1.0, 1.0000001, iters); break;
                    case 4: k<4><<<1, threads>>>(o, c, 1.0, 1.0, iters); break;
                    case 5: k<5><<<1, threads>>>(o, c, 1.0, 1.0, iters); break;
                }
                cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            }
            printf("%-14s threads %3d: %.1f cycles per dependent op\n", names[op], threads, (double)h / (iters * 16.0));
        }
    }
    return 0;
}
