// An existing single-GPU cuSten time stepper and the same stepper on several GPUs (new: the reference is single-GPU).
//
// The first half is a program in the reference's style (cf. examples/src/2d_xy_p_fun.cu): unified-memory buffers, a
// __device__ function handed to cuStenCreate2DXYpFun as a pointer, then Compute + Swap per time step.  The second half
// runs the SAME steps on y-slabs of the grid through the additive C entry points custen_mg_* (include/custen_c.h;
// one process drives all GPUs, the halo rows are read from the neighbour GPUs' memory by the sweep itself) and checks
// that the result has the same bits.
//
//   nvcc -rdc=true -gencode arch=compute_100a,code=sm_100a examples/multi_gpu_stencil.cu custen_b200/lib/libcuSten.a
//   ./multi_gpu_stencil [n] [steps] [slabs]     slabs > GPUs: several slabs share a GPU (how a one-GPU box tests it)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/cuSten.h"
#include "../include/cuSten_fun.h"
#include "../include/custen_c.h"

typedef double (*devArg1XY)(double*, double*, int, int, int, int);

// damped c^3 - c through a 3 x 3 Laplacian: c <- c + coe-weighted sum, bounded for |c| < 1
__device__ double reaction(double* data, double* coe, int loc, int jump, int nx, int ny)
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
    return result + data[loc + jump + 1];   // + the centre value (loc is the window's top-left corner)
}
__device__ devArg1XY devFunc = reaction;
CUSTEN_REGISTER_FUN_XY(reaction)   // optional: lets the library inline the function (same bits either way)

int main(int argc, char* argv[])
{
    const int n = argc > 1 ? atoi(argv[1]) : 1024;
    const int steps = argc > 2 ? atoi(argv[2]) : 10;
    int ngpu = 0;
    cudaGetDeviceCount(&ngpu);
    if (ngpu < 1)
    {
        printf("no CUDA device\n");
        return 1;
    }
    const int slabs = argc > 3 ? atoi(argv[3]) : (ngpu > 1 ? ngpu : 2);
    if (n % (32 * slabs))
    {
        printf("n must be a multiple of 32 x slabs\n");
        return 1;
    }
    const size_t count = (size_t)n * n;
    const double s = 0.05;
    const double coe_host[9] = {0.0, s, 0.0, s, -4.0 * s, s, 0.0, s, 0.0};

    std::vector<double> init(count);
    unsigned long long state = 12345;
    for (size_t i = 0; i < count; ++i)
    {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        init[i] = ((double)(state >> 11) / 9007199254740992.0) * 0.2 - 0.1;
    }

    // ---- the single-GPU stepper, as an existing cuSten program has it ---------------------------------------------
    cudaSetDevice(0);
    double *a, *b, *coe;
    cudaMallocManaged(&a, count * sizeof(double));
    cudaMallocManaged(&b, count * sizeof(double));
    cudaMallocManaged(&coe, 9 * sizeof(double));
    memcpy(a, init.data(), count * sizeof(double));
    memset(b, 0, count * sizeof(double));
    memcpy(coe, coe_host, sizeof coe_host);
    double* func;
    cudaMemcpyFromSymbol(&func, devFunc, sizeof(double*));
    cuSten_t h;
    cuStenCreate2DXYpFun(&h, 0, 1, n, n, 16, 32, b, a, coe, 3, 1, 1, 3, 1, 1, func);
    double *in = a, *out = b;
    for (int k = 0; k < steps; ++k)
    {
        cuStenCompute2DXYpFun(&h, DEVICE);
        cuStenSwap2DXYpFun(&h, out);   // the output of this step is the input of the next
        double* t = in;
        in = out;
        out = t;
    }
    cudaDeviceSynchronize();
    checkError("single-GPU stepper");
    std::vector<double> single(in, in + count);   // `in` holds the newest field
    cuStenDestroy2DXYpFun(&h);

    // ---- the same steps on y-slabs over the GPUs of the box ---------------------------------------------------------
    std::vector<int> devices(slabs);
    std::vector<double*> funcs(slabs);
    for (int i = 0; i < slabs; ++i)
    {
        devices[i] = i % ngpu;
        cudaSetDevice(devices[i]);
        cudaMemcpyFromSymbol(&funcs[i], devFunc, sizeof(double*));   // a device function's address is per device
    }
    const int XYpFun = 10;   // index into Xp Xnp XpFun XnpFun Yp Ynp YpFun YnpFun XYp XYnp XYpFun XYnpFun
    void* mg = custen_mg_create(slabs, devices.data(), XYpFun, n, n, coe_host, 9, 3, 1, 1, 3, 1, 1, nullptr, funcs.data());
    custen_mg_scatter(mg, init.data());
    custen_mg_run(mg, steps);                      // steps x (Compute + Swap) on every slab, asynchronous
    std::vector<double> multi(count);
    custen_mg_gather(mg, multi.data(), 0);         // synchronises; 0 = the newest field
    const int timeouts = custen_mg_error(mg);

    // a second batch for a timing
    cudaEvent_t e0, e1;
    cudaSetDevice(devices[0]);
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    custen_mg_synchronize(mg);
    cudaEventRecord(e0, 0);
    custen_mg_run(mg, 20);
    custen_mg_synchronize(mg);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    custen_mg_destroy(mg);

    size_t differing = 0;
    for (size_t i = 0; i < count; ++i) differing += memcmp(&single[i], &multi[i], sizeof(double)) != 0;
    double sum = 0.0;
    for (size_t i = 0; i < count; i += 4099) sum += multi[i];
    printf("n = %d, %d steps, %d slabs on %d GPU(s): %zu of %zu doubles differ from the single-GPU stepper, "
           "neighbour-wait time-outs %d, sample sum %.12e\n",
           n, steps, slabs, ngpu < slabs ? ngpu : slabs, differing, count, timeouts, sum);
    printf("20 more steps: %.3f ms per step (CUDA events around run + synchronise, %d slab(s))\n", ms / 20.0, slabs);
    cudaFree(a);
    cudaFree(b);
    cudaFree(coe);
    return differing == 0 && timeouts == 0 ? 0 : 2;
}
