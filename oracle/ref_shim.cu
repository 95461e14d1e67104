// GPU oracle shim — TEST INFRASTRUCTURE.
//
// Drives the UNMODIFIED reference library (compiled from the sources where they lie under
// /root/reference/cuSten by oracle/Makefile, only --gpu-architecture changed to sm_100) through a small
// extern "C" surface so that tests can run the reference's own CUDA kernels on the same inputs as the new
// engine.  No reference source is copied: this file only includes the reference's public header and calls
// its API the way examples/src/*.cu do (managed buffers, Create -> Compute -> cudaDeviceSynchronize -> Destroy).
#include "cuSten/cuSten.h"  // resolved with -I/root/reference

#include <cstdio>
#include <cstring>

#include "../custen_b200/csrc/builtin_funs.cuh"

#define EXPORT extern "C" __attribute__((visibility("default")))

typedef double (*FX)(double*, double*, int);
typedef double (*FY)(double*, double*, int, int);
typedef double (*FXY)(double*, double*, int, int, int, int);

__device__ FX rfp_second_diff_x = custen_funs::second_diff_x;
__device__ FX rfp_weighted9_x = custen_funs::weighted9_x;
__device__ FY rfp_weighted9_y = custen_funs::weighted9_y;
__device__ FY rfp_weighted3_y = custen_funs::weighted3_y;
__device__ FXY rfp_weighted_xy = custen_funs::weighted_xy;
__device__ FXY rfp_cubic_xy = custen_funs::cubic_xy;

static double* fun_ptr(const char* name)
{
    void* fp = nullptr;
    if (!name) return nullptr;
#define LOOKUP(N) if (!strcmp(name, #N)) { cudaMemcpyFromSymbol(&fp, rfp_##N, sizeof(void*)); return (double*)fp; }
    LOOKUP(second_diff_x) LOOKUP(weighted9_x) LOOKUP(weighted9_y) LOOKUP(weighted3_y) LOOKUP(weighted_xy) LOOKUP(cubic_xy)
#undef LOOKUP
    return nullptr;
}

struct RefRun
{
    cuSten_t h;
    double *in, *out, *coef;
    int variant_id;
};

// variant ids: 0 Xp 1 Xnp 2 XpFun 3 XnpFun 4 Yp 5 Ynp 6 YpFun 7 YnpFun 8 XYp 9 XYnp 10 XYpFun 11 XYnpFun
static int create(RefRun& r, int v, int tiles, int nx, int ny, int bx, int by, int ncoef, int H, int L, int R, int V,
                  int T, int B, double* fn)
{
    cuSten_t* h = &r.h;
    switch (v)
    {
        case 0: cuStenCreate2DXp(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, H, L, R); break;
        case 1: cuStenCreate2DXnp(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, H, L, R); break;
        case 2: return -2;  // XpFun: no working reference (2d_x_p_fun_kernel.cu:172, custenCreateDestroy2DXpFun.cu:153-157)
        case 3: cuStenCreate2DXnpFun(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, H, L, R, ncoef, fn); break;
        case 4: cuStenCreate2DYp(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, V, T, B); break;
        case 5: cuStenCreate2DYnp(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, V, T, B); break;
        case 6: cuStenCreate2DYpFun(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, V, T, B, ncoef, fn); break;
        case 7: cuStenCreate2DYnpFun(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, V, T, B, fn); break;
        case 8: cuStenCreate2DXYp(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, H, L, R, V, T, B); break;
        case 9: cuStenCreate2DXYnp(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, H, L, R, V, T, B); break;
        case 10: cuStenCreate2DXYpFun(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, H, L, R, V, T, B, fn); break;
        case 11: cuStenCreate2DXYnpFun(h, 0, tiles, nx, ny, bx, by, r.out, r.in, r.coef, H, L, R, V, T, B, fn); break;
        default: return -1;
    }
    r.variant_id = v;
    return 0;
}

static void compute(RefRun& r, bool offload)
{
    cuSten_t* h = &r.h;
    switch (r.variant_id)
    {
        case 0: cuStenCompute2DXp(h, offload); break;
        case 1: cuStenCompute2DXnp(h, offload); break;
        case 3: cuStenCompute2DXnpFun(h, offload); break;
        case 4: cuStenCompute2DYp(h, offload); break;
        case 5: cuStenCompute2DYnp(h, offload); break;
        case 6: cuStenCompute2DYpFun(h, offload); break;
        case 7: cuStenCompute2DYnpFun(h, offload); break;
        case 8: cuStenCompute2DXYp(h, offload); break;
        case 9: cuStenCompute2DXYnp(h, offload); break;
        case 10: cuStenCompute2DXYpFun(h, offload); break;
        case 11: cuStenCompute2DXYnpFun(h, offload); break;
    }
}

static void destroy(RefRun& r)
{
    cuSten_t* h = &r.h;
    switch (r.variant_id)
    {
        case 0: cuStenDestroy2DXp(h); break;
        case 1: cuStenDestroy2DXnp(h); break;
        case 3: cuStenDestroy2DXnpFun(h); break;
        case 4: cuStenDestroy2DYp(h); break;
        case 5: cuStenDestroy2DYnp(h); break;
        case 6: cuStenDestroy2DYpFun(h); break;
        case 7: cuStenDestroy2DYnpFun(h); break;
        case 8: cuStenDestroy2DXYp(h); break;
        case 9: cuStenDestroy2DXYnp(h); break;
        case 10: cuStenDestroy2DXYpFun(h); break;
        case 11: cuStenDestroy2DXYnpFun(h); break;
    }
}

// Run one sweep of the reference on host arrays.  `out_host` is read first (pre-fill / sentinel) and
// overwritten with the reference's result.  Returns 0, or <0 when the reference has no such variant.
EXPORT int ref_sweep(int variant, const double* in_host, double* out_host, int nx, int ny, int tiles, int bx, int by,
                     const double* coef_host, int ncoef, int H, int L, int R, int V, int T, int B, const char* fun,
                     int offload)
{
    RefRun r;
    memset(&r, 0, sizeof r);
    const size_t n = (size_t)nx * ny;
    cudaMallocManaged(&r.in, n * sizeof(double));
    cudaMallocManaged(&r.out, n * sizeof(double));
    cudaMallocManaged(&r.coef, (size_t)(ncoef > 0 ? ncoef : 1) * sizeof(double));
    memcpy(r.in, in_host, n * sizeof(double));
    memcpy(r.out, out_host, n * sizeof(double));
    memcpy(r.coef, coef_host, (size_t)ncoef * sizeof(double));
    int rc = create(r, variant, tiles, nx, ny, bx, by, ncoef, H, L, R, V, T, B, fun_ptr(fun));
    if (rc == 0)
    {
        compute(r, offload != 0);
        cudaDeviceSynchronize();
        checkError("reference sweep");
        memcpy(out_host, r.out, n * sizeof(double));
        destroy(r);
    }
    cudaFree(r.in);
    cudaFree(r.out);
    cudaFree(r.coef);
    return rc;
}

// Time `iters` device-resident sweeps of the reference (managed buffers prefetched to the GPU, offload = DEVICE)
// on a synthetic field; returns milliseconds per sweep, <0 on error.  Used only to fill BASELINE.md's R-GPU rows.
__global__ void ref_fill(double* p, size_t n, unsigned long long seed)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
    {
        unsigned long long z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        p[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
}

EXPORT double ref_time(int variant, int nx, int ny, int tiles, int bx, int by, const double* coef_host, int ncoef, int H,
                       int L, int R, int V, int T, int B, const char* fun, int warmup, int iters)
{
    RefRun r;
    memset(&r, 0, sizeof r);
    const size_t n = (size_t)nx * ny;
    cudaMallocManaged(&r.in, n * sizeof(double));
    cudaMallocManaged(&r.out, n * sizeof(double));
    cudaMallocManaged(&r.coef, (size_t)(ncoef > 0 ? ncoef : 1) * sizeof(double));
    memcpy(r.coef, coef_host, (size_t)ncoef * sizeof(double));
    cudaMemPrefetchAsync(r.in, n * sizeof(double), 0, 0);
    cudaMemPrefetchAsync(r.out, n * sizeof(double), 0, 0);
    ref_fill<<<1024, 256>>>(r.in, n, 0x5EEDull);
    cudaMemset(r.out, 0, n * sizeof(double));
    cudaDeviceSynchronize();
    double ms_per = -1.0;
    if (create(r, variant, tiles, nx, ny, bx, by, ncoef, H, L, R, V, T, B, fun_ptr(fun)) == 0)
    {
        for (int i = 0; i < warmup; ++i) compute(r, false);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, 0);  // legacy stream: ordered against the library's blocking streams
        for (int i = 0; i < iters; ++i) compute(r, false);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        checkError("reference timing");
        ms_per = ms / iters;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        destroy(r);
    }
    cudaFree(r.in);
    cudaFree(r.out);
    cudaFree(r.coef);
    return ms_per;
}


// 13th variant: the reference's WENO advection kernel on host arrays (managed buffers inside, like its example
// examples/src/2d_xyWENOADV_p.cu).
EXPORT int ref_weno(const double* in_host, const double* u_host, const double* v_host, double* out_host, int nx, int ny,
                    int tiles, int bx, int by, double dx, double dy)
{
    const size_t n = (size_t)nx * ny;
    double *in, *out, *u, *v;
    cudaMallocManaged(&in, n * sizeof(double));
    cudaMallocManaged(&out, n * sizeof(double));
    cudaMallocManaged(&u, n * sizeof(double));
    cudaMallocManaged(&v, n * sizeof(double));
    memcpy(in, in_host, n * sizeof(double));
    memcpy(u, u_host, n * sizeof(double));
    memcpy(v, v_host, n * sizeof(double));
    memcpy(out, out_host, n * sizeof(double));
    cuSten_t h;
    cuStenCreate2DXYWENOADVp(&h, 0, tiles, nx, ny, bx, by, dx, dy, u, v, out, in);
    cuStenCompute2DXYWENOADVp(&h, false);
    cudaDeviceSynchronize();
    checkError("reference WENO sweep");
    memcpy(out_host, out, n * sizeof(double));
    cuStenDestroy2DXYWENOADVp(&h);
    cudaFree(in); cudaFree(out); cudaFree(u); cudaFree(v);
    return 0;
}

EXPORT double ref_weno_time(int nx, int ny, int bx, int by, int warmup, int iters)
{
    const size_t n = (size_t)nx * ny;
    double *in, *out, *u, *v;
    cudaMallocManaged(&in, n * sizeof(double));
    cudaMallocManaged(&out, n * sizeof(double));
    cudaMallocManaged(&u, n * sizeof(double));
    cudaMallocManaged(&v, n * sizeof(double));
    for (double* p : {in, out, u, v}) cudaMemPrefetchAsync(p, n * sizeof(double), 0, 0);
    ref_fill<<<1024, 256>>>(in, n, 1);
    ref_fill<<<1024, 256>>>(u, n, 2);
    ref_fill<<<1024, 256>>>(v, n, 3);
    cudaMemset(out, 0, n * sizeof(double));
    cudaDeviceSynchronize();
    cuSten_t h;
    cuStenCreate2DXYWENOADVp(&h, 0, 1, nx, ny, bx, by, 1.0 / nx, 1.0 / ny, u, v, out, in);
    for (int i = 0; i < warmup; ++i) cuStenCompute2DXYWENOADVp(&h, false);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters; ++i) cuStenCompute2DXYWENOADVp(&h, false);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    checkError("reference WENO timing");
    cuStenDestroy2DXYWENOADVp(&h);
    cudaFree(in); cudaFree(out); cudaFree(u); cudaFree(v);
    return ms / iters;
}
