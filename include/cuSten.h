// cuSten-B200: drop-in C++ API of the cuSten 2D stencil engine, re-implemented for sm_100a.
//
// This header replaces, for the 2D X / Y / XY path, the reference header set
//   cuSten/cuSten.h:30-41                         (DEVICE / HOST macros, umbrella include)
//   cuSten/src/struct/cuSten_struct_type.h:84-122 (cuSten_t)
//   cuSten/src/struct/cuSten_struct_functions.h   (Create / Swap / Destroy, 13 variants)
//   cuSten/src/kernels/stencil_kernels.h:52-189   (Compute, 13 variants)
//   cuSten/src/util/util.h:43                     (checkError)
// Signatures and C++ linkage are kept identical (the reference has no extern "C"), so a
// program written against the reference relinks against libcuSten.a from this repo without
// source changes.  The plain-C boundary for FFI users lives in include/custen_c.h.
#ifndef CUSTEN_B200_CUSTEN_H
#define CUSTEN_B200_CUSTEN_H

#include <cuda_runtime.h>

#define DEVICE 0   // offload argument: leave tiles resident on the GPU   (reference cuSten.h:30)
#define HOST 1     // offload argument: return each finished tile to host (reference cuSten.h:31)

// Handle. Field order, types and therefore offsets / sizeof match the reference struct
// (cuSten_struct_type.h:84-122) so user code that pokes at public fields keeps working.
// Storage is owned by the caller; Create fills it, Destroy releases what Create allocated.
typedef struct
{
    int deviceNum;            // CUDA device the handle computes on
    int numStreams;           // public stream count (3, as in the reference)
    int numTiles;             // y-tiles the domain is split into for out-of-core runs
    int nx;                   // points per row
    int ny;                   // rows
    int nyTile;               // rows per tile = ny / numTiles
    int numSten;              // taps (H*V for XY variants)
    int numStenLeft;
    int numStenRight;
    int numStenTop;
    int numStenBottom;
    int numStenHoriz;
    int numStenVert;
    int BLOCK_X;              // reference launch geometry: kept as a hint only
    int BLOCK_Y;
    int xGrid;
    int yGrid;
    int mem_shared;
    double** dataInput;       // per-tile aliases into the caller's input array
    double** dataOutput;      // per-tile aliases into the caller's output array
    double** uVel;            // (WENO variant only)
    double** vVel;
    double* weights;          // caller's weights (weights variants)
    double* coe;              // caller's coefficients (Fun variants)
    double coeDx;
    double coeDy;
    int numCoe;
    int nxLocal;
    int nyLocal;
    double** boundaryTop;     // per-tile pointer to the T rows above the tile
    double** boundaryBottom;  // per-tile pointer to the B rows below the tile
    int numBoundaryTop;
    int numBoundaryBottom;
    cudaStream_t* streams;
    cudaEvent_t* events;
    double* devFunc;          // type-punned __device__ function pointer (Fun variants)
} cuSten_t;

// User function contracts of the Fun variants (reference 2d_x_np_fun_kernel.cu:47,
// 2d_y_p_fun_kernel.cu:49, 2d_xy_p_fun_kernel.cu:51).  `data` is a generic pointer to a
// staged tile, `loc` the centre point (X, Y) or the TOP-LEFT corner of the window (XY),
// `jump` the row pitch of the tile in elements.
typedef double (*cuStenFunX)(double* data, double* coe, int loc);
typedef double (*cuStenFunY)(double* data, double* coe, int loc, int jump);
typedef double (*cuStenFunXY)(double* data, double* coe, int loc, int jump, int nx, int ny);

// ---- error convention (reference util/error.cu:43-53): print and exit(EXIT_FAILURE) ------------
void checkError(const char* action);

// ---- X direction -------------------------------------------------------------------------------
#define CUSTEN_DECL_COMMON(V)                                        \
    void cuStenSwap2D##V(cuSten_t* pt_cuSten, double* dataInput);    \
    void cuStenDestroy2D##V(cuSten_t* pt_cuSten);                    \
    void cuStenCompute2D##V(cuSten_t* pt_cuSten, bool offload);

void cuStenCreate2DXnp(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                       double* dataOutput, double* dataInput, double* weights,
                       int numSten, int numStenLeft, int numStenRight);
CUSTEN_DECL_COMMON(Xnp)

void cuStenCreate2DXp(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                      double* dataOutput, double* dataInput, double* weights,
                      int numSten, int numStenLeft, int numStenRight);
CUSTEN_DECL_COMMON(Xp)

void cuStenCreate2DXnpFun(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                          double* dataOutput, double* dataInput, double* coe,
                          int numSten, int numStenLeft, int numStenRight, int numCoe, double* func);
CUSTEN_DECL_COMMON(XnpFun)

void cuStenCreate2DXpFun(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                         double* dataOutput, double* dataInput, double* coe,
                         int numSten, int numStenLeft, int numStenRight, int numCoe, double* func);
CUSTEN_DECL_COMMON(XpFun)
// The reference defines this entry point under a misspelt name (2d_x_p_fun_kernel.cu:172);
// both spellings are exported.
void cuSenCompute2DXpFun(cuSten_t* pt_cuSten, bool offload);

// ---- Y direction -------------------------------------------------------------------------------
void cuStenCreate2DYnp(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                       double* dataOutput, double* dataInput, double* weights,
                       int numSten, int numStenTop, int numStenBottom);
CUSTEN_DECL_COMMON(Ynp)

void cuStenCreate2DYp(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                      double* dataOutput, double* dataInput, double* weights,
                      int numSten, int numStenTop, int numStenBottom);
CUSTEN_DECL_COMMON(Yp)

// NB: no numCoe here (reference cuSten_struct_functions.h:684-699): numSten coefficients are used.
void cuStenCreate2DYnpFun(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                          double* dataOutput, double* dataInput, double* coe,
                          int numSten, int numStenTop, int numStenBottom, double* func);
CUSTEN_DECL_COMMON(YnpFun)

void cuStenCreate2DYpFun(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                         double* dataOutput, double* dataInput, double* coe,
                         int numSten, int numStenTop, int numStenBottom, int numCoe, double* func);
CUSTEN_DECL_COMMON(YpFun)

// ---- XY direction ------------------------------------------------------------------------------
void cuStenCreate2DXYnp(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                        double* dataOutput, double* dataInput, double* weights,
                        int numStenHoriz, int numStenLeft, int numStenRight,
                        int numStenVert, int numStenTop, int numStenBottom);
CUSTEN_DECL_COMMON(XYnp)

void cuStenCreate2DXYp(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                       double* dataOutput, double* dataInput, double* weights,
                       int numStenHoriz, int numStenLeft, int numStenRight,
                       int numStenVert, int numStenTop, int numStenBottom);
CUSTEN_DECL_COMMON(XYp)

// NB: no numCoe (H*V coefficients are used; reference 2d_xy_np_fun_kernel.cu:117-119).
void cuStenCreate2DXYnpFun(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                           double* dataOutput, double* dataInput, double* coe,
                           int numStenHoriz, int numStenLeft, int numStenRight,
                           int numStenVert, int numStenTop, int numStenBottom, double* func);
CUSTEN_DECL_COMMON(XYnpFun)

void cuStenCreate2DXYpFun(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                          double* dataOutput, double* dataInput, double* coe,
                          int numStenHoriz, int numStenLeft, int numStenRight,
                          int numStenVert, int numStenTop, int numStenBottom, double* func);
CUSTEN_DECL_COMMON(XYpFun)

// ---- XY WENO advection: u dphi/dx + v dphi/dy, periodic, fifth-order upwinded (reference cuSten_struct_functions.h:309) -
void cuStenCreate2DXYWENOADVp(cuSten_t* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                              double dx, double dy, double* u, double* v, double* dataOutput, double* dataInput);
CUSTEN_DECL_COMMON(XYWENOADVp)

#undef CUSTEN_DECL_COMMON

#endif  // CUSTEN_B200_CUSTEN_H
