// fp64_lat.cu — FP64 pipe micro-benchmark (development aid): DFMA / DADD / DMUL rate per SM as a function
// of resident warps per SM sub-partition and independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void k(int iters, double *sink, long long *cyc)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-7;
    const double m = 0.9999999, d = 1e-7;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) a[i] = __fma_rn(a[i], m, d);
                if (OP == 1) a[i] = __dadd_rn(a[i], d);
                if (OP == 2) a[i] = __dmul_rn(a[i], m);
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP, int OP>
void run(int warps_per_sm, int n_sm)
{
    double *sink;
    long long *cyc, h;
    cudaMalloc(&sink, 8);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<ILP, OP><<<n_sm, 32 * warps_per_sm>>>(iters, sink, cyc);
    k<ILP, OP><<<n_sm, 32 * warps_per_sm>>>(iters, sink, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops_per_warp = (double)iters * 16 * ILP;
    // warp-instructions per cycle per SM
    printf("op %d ilp %d warps/SM %2d : %.3f warp-inst/clk/SM, %.2f cycles per dependent op\n", OP, ILP, warps_per_sm,
           ops_per_warp * warps_per_sm / (double)h, (double)h / (iters * 16.0));
    cudaFree(sink);
    cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int n = p.multiProcessorCount;
    for (int w : {4, 8, 12, 16, 24, 32}) {
        run<1, 0>(w, n); run<2, 0>(w, n); run<4, 0>(w, n); run<8, 0>(w, n);
    }
    for (int w : {4, 12, 16}) { run<1, 1>(w, n); run<4, 1>(w, n); run<1, 2>(w, n); run<4, 2>(w, n); }
    return 0;
}
