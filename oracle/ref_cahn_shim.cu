// GPU oracle shim for BASELINE.json config 5 — TEST INFRASTRUCTURE.
//
// Pulls in the reference's GPU timing twin (cuPentSpeedUp/cuPentCahnADITiming/src/cuPentCahnADI.cu) where it lies,
// with its main() renamed, so that its file-local kernels (findCBar, findRHS, findNew), its user function
// (nonLinRHS) and its solver pieces (BatchHyper.cu, cuPentBatch.cu, compiled next to this file by oracle/Makefile)
// can be driven from a test: same calls, same order, same launch shapes as the reference's time loop (:500-566), but
// from a caller-supplied initial field, for a caller-chosen number of steps, returning the final field.
#define main ref_cahn_timing_main
#include "cuPentSpeedUp/cuPentCahnADITiming/src/cuPentCahnADI.cu"  // resolved with -I/root/reference -Ioracle/stub
#undef main

#define EXPORT extern "C" __attribute__((visibility("default")))

// `warm` untimed steps run first (unified-memory pages migrate to the GPU on first touch); then `nsteps` timed steps.
EXPORT int ref_cahn_run2(int nx, int warm, int nsteps, double lx, const double* c0, double* c_out, double* ms_per_step)
{
    const double D = 1.0, gamma = 0.01;
    const int size = nx - 2;
    const double dx = lx / nx, dt = 0.1 * dx;
    const size_t N = (size_t)nx * nx;

    int gridInv = (nx % BLOCK_INV == 0) ? (nx / BLOCK_INV) : (nx / BLOCK_INV + 1);
    dim3 blockDimInv(BLOCK_INV), gridDimInv(gridInv);
    int xGrid = (nx % BLOCK_X == 0) ? (nx / BLOCK_X) : (nx / BLOCK_X + 1);
    int yGrid = (nx % BLOCK_Y == 0) ? (nx / BLOCK_Y) : (nx / BLOCK_Y + 1);
    dim3 blockDim(BLOCK_X, BLOCK_Y), gridDim(xGrid, yGrid);

    double *cOld, *cCurr, *cNonLinRHS, *cBar, *cHalf, *ds, *dl, *diag, *du, *dw, *inv1Multi, *inv2Multi;
    cudaMallocManaged(&cOld, N * sizeof(double));
    cudaMallocManaged(&cCurr, N * sizeof(double));
    cudaMallocManaged(&cNonLinRHS, N * sizeof(double));
    cudaMallocManaged(&cBar, N * sizeof(double));
    cudaMallocManaged(&cHalf, N * sizeof(double));
    for (size_t i = 0; i < N; ++i) cOld[i] = cCurr[i] = c0[i];
    cudaMallocManaged(&ds, (size_t)size * nx * sizeof(double));
    cudaMallocManaged(&dl, (size_t)size * nx * sizeof(double));
    cudaMallocManaged(&diag, (size_t)size * nx * sizeof(double));
    cudaMallocManaged(&du, (size_t)size * nx * sizeof(double));
    cudaMallocManaged(&dw, (size_t)size * nx * sizeof(double));

    cublasHandle_t handleBLAS;
    cublasCreate(&handleBLAS);
    const double alpha = 1.0, beta = 0.0;

    double simgaLin = 2.0 * dt * D * gamma / (3.0 * (pow(dx, 4.0)));
    double a = simgaLin, b = -4 * simgaLin, c = 1 + 6 * simgaLin, d = -4 * simgaLin, e = simgaLin;
    setMultiLHS<<<gridDim, blockDim>>>(ds, dl, diag, du, dw, a, b, c, d, e, size, nx);
    cudaDeviceSynchronize();
    pentFactorBatch<<<gridDimInv, blockDimInv>>>(ds, dl, diag, du, dw, size, nx);
    cudaDeviceSynchronize();

    double omega[4];
    double* inv1Single = (double*)malloc(size * sizeof(double));
    double* inv2Single = (double*)malloc(size * sizeof(double));
    cudaMallocManaged(&inv1Multi, (size_t)nx * size * sizeof(double));
    cudaMallocManaged(&inv2Multi, (size_t)nx * size * sizeof(double));
    findOmega(omega, inv1Single, inv2Single, a, b, c, d, e, nx);
    for (int j = 0; j < size; j++)
        for (int i = 0; i < nx; i++)
        {
            inv1Multi[(size_t)j * nx + i] = inv1Single[j];
            inv2Multi[(size_t)j * nx + i] = inv2Single[j];
        }

    double* weightsLinRHS;
    cudaMallocManaged(&weightsLinRHS, 25 * sizeof(double));
    const double wl[25] = {0.0, 0.0, -1.0 * simgaLin, 0.0, 0.0,
                           0.0, -2.0 * simgaLin, 8.0 * simgaLin, -2.0 * simgaLin, 0.0,
                           -1.0 * simgaLin, 8.0 * simgaLin, -20.0 * simgaLin, 8.0 * simgaLin, -1.0 * simgaLin,
                           0.0, -2.0 * simgaLin, 8.0 * simgaLin, -2.0 * simgaLin, 0.0,
                           0.0, 0.0, -1.0 * simgaLin, 0.0, 0.0};
    for (int i = 0; i < 25; ++i) weightsLinRHS[i] = wl[i];
    cuSten_t linRHS;
    cuStenCreate2DXYp(&linRHS, 0, 1, nx, nx, BLOCK_X, BLOCK_Y, cHalf, cBar, weightsLinRHS, 5, 2, 2, 5, 2, 2);
    cudaDeviceSynchronize();

    cuSten_t nonLinCompute;
    double* func;
    cudaMemcpyFromSymbol(&func, devFunc, sizeof(devArg1XY));
    double sigmaNonLin = (dt / 3.0) * D * (2.0 / pow(dx, 2.0));
    double* coe;
    cudaMallocManaged(&coe, 9 * sizeof(double));
    const double cn[9] = {0.0, 1.0 * sigmaNonLin, 0.0, 1.0 * sigmaNonLin, -4.0 * sigmaNonLin, 1.0 * sigmaNonLin,
                          0.0, 1.0 * sigmaNonLin, 0.0};
    for (int i = 0; i < 9; ++i) coe[i] = cn[i];
    cuStenCreate2DXYpFun(&nonLinCompute, 0, 1, nx, nx, BLOCK_X_FUN, BLOCK_Y_FUN, cNonLinRHS, cCurr, coe, 3, 1, 1, 3, 1, 1, func);
    cudaDeviceSynchronize();

    cudaEvent_t start, stop;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    auto one_step = [&]() {
        findCBar<<<gridDim, blockDim>>>(cOld, cCurr, cBar, nx);
        cudaDeviceSynchronize();
        cuStenCompute2DXYpFun(&nonLinCompute, 0);
        cuStenCompute2DXYp(&linRHS, 0);
        cudaDeviceSynchronize();
        findRHS<<<gridDim, blockDim>>>(cOld, cCurr, cHalf, cNonLinRHS, nx);
        cudaDeviceSynchronize();
        cublasDgeam(handleBLAS, CUBLAS_OP_T, CUBLAS_OP_T, nx, nx, &alpha, cHalf, nx, &beta, NULL, nx, cCurr, nx);
        cudaDeviceSynchronize();
        cyclicInv(ds, dl, diag, du, dw, inv1Multi, inv2Multi, omega, cCurr, a, b, d, e, BLOCK_INV, BLOCK_X, BLOCK_Y, size, nx);
        cublasDgeam(handleBLAS, CUBLAS_OP_T, CUBLAS_OP_T, nx, nx, &alpha, cCurr, nx, &beta, NULL, nx, cHalf, nx);
        cudaDeviceSynchronize();
        cyclicInv(ds, dl, diag, du, dw, inv1Multi, inv2Multi, omega, cHalf, a, b, d, e, BLOCK_INV, BLOCK_X, BLOCK_Y, size, nx);
        cudaDeviceSynchronize();
        findNew<<<gridDim, blockDim>>>(cCurr, cBar, cHalf, nx);
        cudaDeviceSynchronize();
        };
    for (int it = 0; it < warm; ++it) one_step();
    cudaDeviceSynchronize();
    cudaEventRecord(start, 0);
    for (int it = 0; it < nsteps; ++it) one_step();
    cudaDeviceSynchronize();
    cudaEventRecord(stop, 0);
    cudaEventSynchronize(stop);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, start, stop);
    if (ms_per_step) *ms_per_step = nsteps > 0 ? ms / nsteps : 0.0;
    checkError("reference Cahn-Hilliard run");
    for (size_t i = 0; i < N; ++i) c_out[i] = cCurr[i];

    cuStenDestroy2DXYp(&linRHS);
    cuStenDestroy2DXYpFun(&nonLinCompute);
    cublasDestroy(handleBLAS);
    free(inv1Single);
    free(inv2Single);
    for (double* p : {cOld, cCurr, cNonLinRHS, cBar, cHalf, ds, dl, diag, du, dw, inv1Multi, inv2Multi, weightsLinRHS, coe})
        cudaFree(p);
    cudaEventDestroy(start);
    cudaEventDestroy(stop);
    return 0;
}

EXPORT int ref_cahn_run(int nx, int nsteps, double lx, const double* c0, double* c_out, double* ms_per_step)
{
    return ref_cahn_run2(nx, 0, nsteps, lx, c0, c_out, ms_per_step);
}
