// A user program in the reference's style (cf. examples/src/2d_xy_p_fun.cu): unified-memory buffers, a __device__
// function handed to cuStenCreate2DXYpFun as a pointer read back with cudaMemcpyFromSymbol.  It runs the sweep twice —
// first with the function known to the library only as an opaque pointer, then again after the ONE additional line
// CUSTEN_REGISTER_FUN_XY(...) has let the library inline it — and checks that both give the same bits.
//
//   nvcc -rdc=true -gencode arch=compute_100a,code=sm_100a examples/registered_fun.cu custen_b200/lib/libcuSten.a
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../include/cuSten.h"
#include "../include/cuSten_fun.h"

typedef double (*devArg1XY)(double*, double*, int, int, int, int);

// two user functions with the same body: one registered, one not
__device__ double laplaceOfCube(double* data, double* coe, int loc, int jump, int nx, int ny)
{
    double result = 0.0;
    int count = 0;
    for (int j = 0; j < ny; j++)
    {
        const int temp = loc + j * jump;
        for (int i = 0; i < nx; i++)
        {
            const double current = data[temp + i];
            result += coe[count] * ((current * current * current) - current);
            count++;
        }
    }
    return result;
}
__device__ double laplaceOfCubeOpaque(double* data, double* coe, int loc, int jump, int nx, int ny)
{
    return laplaceOfCube(data, coe, loc, jump, nx, ny);
}

__device__ devArg1XY devFunc = laplaceOfCube;
__device__ devArg1XY devFuncOpaque = laplaceOfCubeOpaque;
CUSTEN_REGISTER_FUN_XY(laplaceOfCube)   // <- the only line a cuSten user adds

static double sweep(double* func, double* out, double* in, double* coe, int nx, int ny, float* ms)
{
    cuSten_t h;
    cuStenCreate2DXYpFun(&h, 0, 1, nx, ny, 16, 32, out, in, coe, 3, 1, 1, 3, 1, 1, func);
    cuStenCompute2DXYpFun(&h, DEVICE);  // warm-up (also migrates the unified-memory pages)
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    cuStenCompute2DXYpFun(&h, DEVICE);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(ms, e0, e1);
    cuStenDestroy2DXYpFun(&h);
    double s = 0.0;
    for (size_t i = 0; i < (size_t)nx * ny; i += 4097) s += out[i];
    return s;
}

int main(int argc, char** argv)
{
    const int nx = argc > 1 ? atoi(argv[1]) : 4096, ny = nx;
    const size_t n = (size_t)nx * ny;
    double *in, *outA, *outB, *coe;
    cudaMallocManaged(&in, n * sizeof(double));
    cudaMallocManaged(&outA, n * sizeof(double));
    cudaMallocManaged(&outB, n * sizeof(double));
    cudaMallocManaged(&coe, 9 * sizeof(double));
    const double dx = 2.0 * M_PI / nx;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) in[(size_t)j * nx + i] = 0.1 * sin(i * dx) * cos(j * dx);
    const double sg = 0.25;
    const double c9[9] = {0.0, sg, 0.0, sg, -4.0 * sg, sg, 0.0, sg, 0.0};
    memcpy(coe, c9, sizeof c9);
    memset(outA, 0, n * sizeof(double));
    memset(outB, 0, n * sizeof(double));

    double *fReg, *fOpq;
    cudaMemcpyFromSymbol(&fReg, devFunc, sizeof(devArg1XY));
    cudaMemcpyFromSymbol(&fOpq, devFuncOpaque, sizeof(devArg1XY));
    float msOpq = 0.f, msReg = 0.f;
    sweep(fOpq, outA, in, coe, nx, ny, &msOpq);
    sweep(fReg, outB, in, coe, nx, ny, &msReg);
    cudaDeviceSynchronize();
    checkError("registered_fun example");
    const bool same = memcmp(outA, outB, n * sizeof(double)) == 0;
    printf("%d x %d: opaque pointer %.3f ms (%.1f Gpoints/s), registered %.3f ms (%.1f Gpoints/s), results %s\n", nx, ny,
           msOpq, n / msOpq / 1e6, msReg, n / msReg / 1e6, same ? "bit-identical" : "DIFFER");
    cudaFree(in); cudaFree(outA); cudaFree(outB); cudaFree(coe);
    return same ? 0 : 1;
}
