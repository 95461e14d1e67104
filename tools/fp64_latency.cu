// Dependent-issue latency of the FP64 pipe on one warp (evidence for the Cahn-Hilliard solve's latency roofline:
// its recurrence is a chain of 4 dependent FP64 operations per row forward and 2 backward, one warp per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency.bin tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void chain(double* out, long long* cycles, double a, double b, int iters)
{
    double x = a + threadIdx.x, y = b, z = a * 0.5, w = b * 0.25;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            if (MODE == 0) x = fma(x, a, b);                      // one dependent DFMA chain
            if (MODE == 1) x = x * a;                             // dependent DMUL
            if (MODE == 2) { x = fma(x, a, b); y = fma(y, a, b); }  // two independent chains
            if (MODE == 3) { x = fma(x, a, b); y = fma(y, a, b); z = fma(z, a, b); w = fma(w, a, b); }
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = x + y + z + w;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

int main()
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(double));
    cudaMallocManaged(&cyc, sizeof(long long));
    const int iters = 4096;
    const char* names[4] = {"dependent DFMA", "dependent DMUL", "2 independent DFMA chains", "4 independent DFMA chains"};
    for (int threads = 32; threads <= 128; threads *= 4)
        for (int m = 0; m < 4; ++m)
        {
            for (int rep = 0; rep < 2; ++rep)
            {
                if (m == 0) chain<0><<<1, threads>>>(out, cyc, 0.999, 1e-3, iters);
                if (m == 1) chain<1><<<1, threads>>>(out, cyc, 0.999, 1e-3, iters);
                if (m == 2) chain<2><<<1, threads>>>(out, cyc, 0.999, 1e-3, iters);
                if (m == 3) chain<3><<<1, threads>>>(out, cyc, 0.999, 1e-3, iters);
                cudaDeviceSynchronize();
            }
            const int per = m < 2 ? 1 : (m == 2 ? 2 : 4);
            printf("%d threads, %-28s: %.2f cycles per chain step (%.2f per instruction)\n", threads, names[m],
                   (double)cyc[0] / (iters * 16.0), (double)cyc[0] / (iters * 16.0 * per));
        }
    return 0;
}
