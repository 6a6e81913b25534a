// fp64_mix.cu — does an FP64 instruction hold the SMSP issue port for 2 cycles?  (development aid)
// Each thread runs 4 independent DFMA chains interleaved with K independent integer ops per DFMA.
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int MODE>
__global__ void k(int iters, double *sink, long long *cyc)
{
    double a[4];
    unsigned x[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-7;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; f[i] = 1.0f + i; }
    const double m = 0.9999999, d = 1e-7;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = __fma_rn(a[i], m, d);
#pragma unroll
                for (int q = 0; q < K; ++q) {
                    const int j = (i * K + q) & 7;
                    if (MODE == 0) x[j] = x[j] * 3u + 1u;          // IMAD
                    if (MODE == 1) f[j] = __fmaf_rn(f[j], 0.999f, 0.5f);  // FFMA
                    if (MODE == 2) x[j] = (x[j] << 1) ^ 0x5u;      // LOP3/SHF
                }
            }
    }
    long long t1 = clock64();
    double s = a[0] + a[1] + a[2] + a[3];
    unsigned xs = 0; float fs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { xs += x[i]; fs += f[i]; }
    if (s == 123.456 || xs == 0xdeadbeef || fs == 1.2345f) sink[0] = s + xs + fs;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int K, int MODE>
void run(int warps_per_sm, int n_sm)
{
    double *sink; long long *cyc, h;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    k<K, MODE><<<n_sm, 32 * warps_per_sm>>>(iters, sink, cyc);
    k<K, MODE><<<n_sm, 32 * warps_per_sm>>>(iters, sink, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double nD = (double)iters * 32 * (warps_per_sm / 4.0);  // DFMA warp-instructions per SMSP
    printf("mode %d K %d warps/SM %2d : %.2f cycles per DFMA per SMSP (%.2f with %d other ops each)\n", MODE, K, warps_per_sm,
           (double)h / nD, (double)h / nD, K);
    cudaFree(sink); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int n = p.multiProcessorCount;
    for (int w : {8, 16}) {
        run<0, 0>(w, n); run<1, 0>(w, n); run<2, 0>(w, n); run<3, 0>(w, n); run<4, 0>(w, n);
        run<1, 1>(w, n); run<2, 1>(w, n); run<1, 2>(w, n); run<2, 2>(w, n);
    }
    return 0;
}
