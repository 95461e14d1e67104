// The reference's C++ entry points (52 symbols, same mangled names) on top of the generic plan.
// Replaces cuSten/src/struct/custenCreateDestroy2D*.cu and the host halves of cuSten/src/kernels/2d_*_kernel.cu.
#include "plan.h"

#include <cstdio>
#include <cstdlib>

using namespace custen;

// cuSten/src/util/error.cu:43-53 — same message, same exit behaviour.
void checkError(const char* action)
{
    cudaError_t error = cudaGetLastError();
    if (error != cudaSuccess)
    {
        printf("\nError while '%s': %s\nprogram terminated ...\n\n", action, cudaGetErrorString(error));
        exit(EXIT_FAILURE);
    }
}

#define CUSTEN_COMMON(V)                                                                   \
    void cuStenSwap2D##V(cuSten_t* pt_cuSten, double* dataInput) { plan_swap(pt_cuSten, dataInput); } \
    void cuStenDestroy2D##V(cuSten_t* pt_cuSten) { plan_destroy(pt_cuSten); }               \
    void cuStenCompute2D##V(cuSten_t* pt_cuSten, bool offload) { plan_compute(pt_cuSten, offload); }

// ---- X ------------------------------------------------------------------------------------------
void cuStenCreate2DXnp(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                       double* weights, int numSten, int L, int R)
{
    plan_create(h, Spec{DIR_X, 0, 0, 0}, dev, tiles, nx, ny, bx, by, out, in, weights, numSten, L, R, 1, 0, 0, 0, nullptr);
}
CUSTEN_COMMON(Xnp)

void cuStenCreate2DXp(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                      double* weights, int numSten, int L, int R)
{
    plan_create(h, Spec{DIR_X, 1, 0, 0}, dev, tiles, nx, ny, bx, by, out, in, weights, numSten, L, R, 1, 0, 0, 0, nullptr);
}
CUSTEN_COMMON(Xp)

void cuStenCreate2DXnpFun(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                          double* coe, int numSten, int L, int R, int numCoe, double* func)
{
    plan_create(h, Spec{DIR_X, 0, 1, 0}, dev, tiles, nx, ny, bx, by, out, in, coe, numSten, L, R, 1, 0, 0, numCoe, func);
}
CUSTEN_COMMON(XnpFun)

void cuStenCreate2DXpFun(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                         double* coe, int numSten, int L, int R, int numCoe, double* func)
{
    plan_create(h, Spec{DIR_X, 1, 1, 0}, dev, tiles, nx, ny, bx, by, out, in, coe, numSten, L, R, 1, 0, 0, numCoe, func);
}
CUSTEN_COMMON(XpFun)
void cuSenCompute2DXpFun(cuSten_t* pt_cuSten, bool offload) { plan_compute(pt_cuSten, offload); }

// ---- Y ------------------------------------------------------------------------------------------
void cuStenCreate2DYnp(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                       double* weights, int numSten, int T, int B)
{
    plan_create(h, Spec{DIR_Y, 0, 0, 0}, dev, tiles, nx, ny, bx, by, out, in, weights, 1, 0, 0, numSten, T, B, 0, nullptr);
}
CUSTEN_COMMON(Ynp)

void cuStenCreate2DYp(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                      double* weights, int numSten, int T, int B)
{
    plan_create(h, Spec{DIR_Y, 1, 0, 0}, dev, tiles, nx, ny, bx, by, out, in, weights, 1, 0, 0, numSten, T, B, 0, nullptr);
}
CUSTEN_COMMON(Yp)

void cuStenCreate2DYnpFun(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                          double* coe, int numSten, int T, int B, double* func)
{
    // no numCoe in this signature: the kernel takes numSten coefficients (2d_y_np_fun_kernel.cu:102)
    plan_create(h, Spec{DIR_Y, 0, 1, 0}, dev, tiles, nx, ny, bx, by, out, in, coe, 1, 0, 0, numSten, T, B, numSten, func);
}
CUSTEN_COMMON(YnpFun)

void cuStenCreate2DYpFun(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                         double* coe, int numSten, int T, int B, int numCoe, double* func)
{
    plan_create(h, Spec{DIR_Y, 1, 1, 0}, dev, tiles, nx, ny, bx, by, out, in, coe, 1, 0, 0, numSten, T, B, numCoe, func);
}
CUSTEN_COMMON(YpFun)

// ---- XY -----------------------------------------------------------------------------------------
void cuStenCreate2DXYnp(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                        double* weights, int H, int L, int R, int V, int T, int B)
{
    plan_create(h, Spec{DIR_XY, 0, 0, 0}, dev, tiles, nx, ny, bx, by, out, in, weights, H, L, R, V, T, B, 0, nullptr);
}
CUSTEN_COMMON(XYnp)

void cuStenCreate2DXYp(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                       double* weights, int H, int L, int R, int V, int T, int B)
{
    plan_create(h, Spec{DIR_XY, 1, 0, 0}, dev, tiles, nx, ny, bx, by, out, in, weights, H, L, R, V, T, B, 0, nullptr);
}
CUSTEN_COMMON(XYp)

void cuStenCreate2DXYnpFun(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                           double* coe, int H, int L, int R, int V, int T, int B, double* func)
{
    // H*V coefficients (2d_xy_np_fun_kernel.cu:117-119)
    plan_create(h, Spec{DIR_XY, 0, 1, 0}, dev, tiles, nx, ny, bx, by, out, in, coe, H, L, R, V, T, B, H * V, func);
}
CUSTEN_COMMON(XYnpFun)

void cuStenCreate2DXYpFun(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double* out, double* in,
                          double* coe, int H, int L, int R, int V, int T, int B, double* func)
{
    plan_create(h, Spec{DIR_XY, 1, 1, 0}, dev, tiles, nx, ny, bx, by, out, in, coe, H, L, R, V, T, B, H * V, func);
}
CUSTEN_COMMON(XYpFun)

// ---- XY WENO advection (13th variant) ------------------------------------------------------------------------------
void cuStenCreate2DXYWENOADVp(cuSten_t* h, int dev, int tiles, int nx, int ny, int bx, int by, double dx, double dy,
                              double* u, double* v, double* out, double* in)
{
    plan_create_weno(h, dev, tiles, nx, ny, bx, by, dx, dy, u, v, out, in);
}
CUSTEN_COMMON(XYWENOADVp)
