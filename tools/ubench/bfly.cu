// Arithmetic-only throughput of the Fp252 DIF butterfly / multiplication (no memory, no barriers):
// separates the arithmetic pipeline bound from the NTT kernel's memory + sync overheads.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sandstorm_b200/csrc/fp252.cuh"
using namespace ss;
#define ITERS 512
template <int MODE, int ILP>
__global__ void k(Fp *out, long long *cycles) {
    Fp a[ILP], b[ILP], w;
    for (int j = 0; j < ILP; ++j) for (int i = 0; i < 8; ++i) { a[j].l[i] = threadIdx.x * 77 + i + j; b[j].l[i] = blockIdx.x * 13 + i * 3 + j; }
    for (int i = 0; i < 8; ++i) w.l[i] = 0x1234567u * (i + 1);
    a[0].l[7] &= 0x07ffffff; w.l[7] &= 0x07ffffff;
    for (int j = 0; j < ILP; ++j) { a[j].l[7] &= 0x07ffffff; b[j].l[7] &= 0x07ffffff; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (MODE == 0) {            // DIF butterfly
                const Fp d = fp::sub4p(a[j], b[j]);
                a[j] = fp::add_fast(a[j], b[j]);
                b[j] = fp::mul(d, w);
            } else if (MODE == 1) {     // multiplication only
                b[j] = fp::mul(b[j], w);
            } else {                    // add/sub only
                const Fp d = fp::sub4p(a[j], b[j]);
                a[j] = fp::add_fast(a[j], b[j]);
                b[j] = d; fp::cond_sub_4p(b[j]); fp::cond_sub_2p(b[j]);
            }
        }
    }
    long long t1 = clock64();
    Fp acc = a[0];
    for (int j = 0; j < ILP; ++j) acc = fp::add(acc, fp::add(a[j], b[j]));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP>
void run(const char *name) {
    Fp *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(Fp)); cudaMalloc(&cyc, 148 * 8);
    for (int threads : {128, 256, 512, 768, 1024}) {
        k<MODE, ILP><<<148, threads>>>(out, cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s threads=%d launch failed (registers)\n", name, threads); cudaGetLastError(); continue; }
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        double warp_ops = (double)ITERS * ILP * (threads / 32);
        printf("%-28s ILP=%d threads=%4d  cycles/warp-op/SMSP=%7.1f\n", name, ILP, threads, c / (warp_ops / 4));
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 1>("DIF butterfly"); run<0, 2>("DIF butterfly"); run<0, 4>("DIF butterfly");
    run<1, 1>("mul only"); run<1, 2>("mul only"); run<1, 4>("mul only");
    run<2, 2>("add/sub only");
    return 0;
}
