// Geometry sweep for stream_acc_kernel on a B200 (development tool, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -rdc=true tools/tune_stream.cu custen_b200/lib/libcuSten.a -o build/tune_stream
#include "../custen_b200/csrc/stream_kernels.cuh"
#include <cstdio>
#include <vector>
#include <algorithm>
using namespace custen;

__global__ void fill(double* p, size_t n)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        unsigned long long z = (i + 12345) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        p[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
}

struct Res { char name[96]; double gpts; };
static std::vector<Res> results;
static int g_n = 16384;
static double *g_in, *g_out, *g_coef;

template <int NT, int SR, int NS, int H, int V, int LODD, int CPT>
void run(int cps, int chunk_target)
{
    Band b{};
    b.in = g_in; b.out = g_out; b.coef = g_coef; b.nx = g_n; b.rows = g_n;
    b.L = (H - 1) / 2; b.R = (H - 1) / 2; b.T = (V - 1) / 2; b.B = (V - 1) / 2; b.H = H; b.V = V; b.ncoef = H * V;
    b.dir = V == 1 ? DIR_X : (H == 1 ? DIR_Y : DIR_XY);
    b.wrap_x = 1; b.have_top = V > 1; b.have_bottom = V > 1;
    b.top = g_in + (size_t)(g_n - b.T) * g_n; b.bottom = g_in;
    b.xlo = 0; b.xhi = g_n; b.ylo = 0; b.yhi = g_n;
    StreamArgs a{};
    a.b = b;
    a.TW = CPT * NT; a.Lp = (b.L + 1) & ~1; a.Rp = (b.R + 1) & ~1; a.PW = a.Lp + a.TW + a.Rp; a.Beff = b.B; a.PFX = 0;
    a.nstrips = (g_n + a.TW - 1) / a.TW;
    a.stage_doubles = SR * a.PW;
    auto kernel = stream_acc_kernel<NT, SR, NS, H, V, LODD, CPT, 1>;
    const size_t smem = SMEM_STAGE_OFF + (size_t)NS * a.stage_doubles * 8;
    if (smem > 227 * 1024) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int maxb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, kernel, NT + 32, smem);
    if (cps > maxb) return;
    const int ncta = 148 * cps;
    long items_t = (long)a.nstrips * ((b.rows + chunk_target - 1) / chunk_target);
    long waves = std::max(1L, (items_t + ncta / 2) / ncta);
    long nch = std::max(1L, (waves * ncta) / a.nstrips);
    int ch = (int)((b.rows + nch - 1) / nch);
    a.chunk_rows = ch; a.nchunks = (b.rows + ch - 1) / ch; a.nitems = a.nstrips * a.nchunks;
    const int grid = std::min(a.nitems, ncta);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 4; ++i) kernel<<<grid, NT + 32, smem>>>(a);
    cudaEventRecord(e0);
    const int iters = 25;
    for (int i = 0; i < iters; ++i) kernel<<<grid, NT + 32, smem>>>(a);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    Res r;
    snprintf(r.name, sizeof r.name, "H%dV%d NT%d CPT%d SR%d NS%d cps%d chunk%d(ch=%d) smem%zuK%s", H, V, NT, CPT, SR, NS, cps,
             chunk_target, ch, smem / 1024, err ? " ERR" : "");
    r.gpts = (double)g_n * g_n / (ms / iters) / 1e6;
    results.push_back(r);
    printf("%-70s %8.1f Gpt/s  %5.1f%% of 6548 GB/s\n", r.name, r.gpts, r.gpts * 16 / 6548.5 * 100);
    fflush(stdout);
}

template <int H, int V, int LODD, int SRA, int SRB>
void sweep()
{
    for (int rep = 0; rep < 2; ++rep)
    for (int cps = 1; cps <= 2; ++cps)
    {
        run<512, 8, 3, H, V, LODD, 1>(cps, 192);
        run<512, SRA, 3, H, V, LODD, 1>(cps, 192);
        run<512, SRB, 3, H, V, LODD, 1>(cps, 192);
        run<512, SRA, 4, H, V, LODD, 1>(cps, 192);
        run<256, SRA, 3, H, V, LODD, 2>(cps, 192);
        run<256, SRB, 3, H, V, LODD, 2>(cps, 192);
        run<256, SRA, 4, H, V, LODD, 2>(cps, 192);
        run<128, SRA, 3, H, V, LODD, 2>(cps, 192);
        run<128, SRB, 3, H, V, LODD, 2>(cps, 192);
        run<128, SRA, 4, H, V, LODD, 2>(cps, 192);
        run<256, SRA, 3, H, V, LODD, 1>(cps, 192);
        run<256, SRB, 3, H, V, LODD, 1>(cps, 192);
        run<768, SRA, 3, H, V, LODD, 1>(cps, 192);
    }
}

// ---- tile family (inlined user function) -------------------------------------------------------------------
__device__ inline double cubic_xy(double* data, double* coe, int loc, int jump, int nx, int ny)
{
    double acc = 0.0;
    int c = 0;
    for (int j = 0; j < ny; ++j)
    {
        const int row = loc + j * jump;
        for (int i = 0; i < nx; ++i)
        {
            const double v = data[row + i];
            acc += coe[c++] * ((v * v * v) - v);
        }
    }
    return acc;
}
__device__ inline double weighted9_y(double* data, double* coe, int loc, int jump)
{
    double acc = 0.0;
    for (int k = 0; k < 9; ++k) acc += coe[k] * data[loc + (k - 4) * jump];
    return acc;
}

template <int NT, int SR, int NS, int MINB, class Op>
void run_tile(const char* opname, int H, int V, int cps, int chunk_target)
{
    Band b{};
    b.in = g_in; b.out = g_out; b.coef = g_coef; b.nx = g_n; b.rows = g_n;
    b.L = (H - 1) / 2; b.R = (H - 1) / 2; b.T = (V - 1) / 2; b.B = (V - 1) / 2; b.H = H; b.V = V; b.ncoef = H * V;
    b.dir = V == 1 ? DIR_X : (H == 1 ? DIR_Y : DIR_XY);
    b.wrap_x = 1; b.have_top = V > 1; b.have_bottom = V > 1;
    b.top = g_in + (size_t)(g_n - b.T) * g_n; b.bottom = g_in;
    b.xlo = 0; b.xhi = g_n; b.ylo = 0; b.yhi = g_n;
    StreamArgs a{};
    a.b = b;
    a.TW = NT; a.Lp = (b.L + 1) & ~1; a.Rp = (b.R + 1) & ~1; a.PW = a.Lp + a.TW + a.Rp; a.Beff = b.B; a.PFX = V - 1;
    a.nstrips = (g_n + a.TW - 1) / a.TW;
    a.stage_doubles = (a.PFX + SR) * a.PW;
    auto kernel = stream_tile_kernel<NT, SR, NS, MINB, Op>;
    const size_t smem = SMEM_STAGE_OFF + (size_t)NS * a.stage_doubles * 8;
    if (smem > 227 * 1024) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int maxb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, kernel, NT + 32, smem);
    if (cps > maxb) return;
    const int ncta = 148 * cps;
    long items_t = (long)a.nstrips * ((b.rows + chunk_target - 1) / chunk_target);
    long waves = std::max(1L, (items_t + ncta / 2) / ncta);
    long nch = std::max(1L, (waves * ncta) / a.nstrips);
    int ch = (int)((b.rows + nch - 1) / nch);
    a.chunk_rows = ch; a.nchunks = (b.rows + ch - 1) / ch; a.nitems = a.nstrips * a.nchunks;
    const int grid = std::min(a.nitems, ncta);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 4; ++i) kernel<<<grid, NT + 32, smem>>>(a);
    cudaEventRecord(e0);
    const int iters = 25;
    for (int i = 0; i < iters; ++i) kernel<<<grid, NT + 32, smem>>>(a);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    Res r;
    snprintf(r.name, sizeof r.name, "TILE %s NT%d SR%d NS%d cps%d chunk%d smem%zuK%s", opname, NT, SR, NS, cps, chunk_target,
             smem / 1024, err ? " ERR" : "");
    r.gpts = (double)g_n * g_n / (ms / iters) / 1e6;
    results.push_back(r);
    printf("%-70s %8.1f Gpt/s  %5.1f%% of 6548 GB/s\n", r.name, r.gpts, r.gpts * 16 / 6548.5 * 100);
    fflush(stdout);
}

template <class Op>
void sweep_tile(const char* opname, int H, int V)
{
    for (int cps = 1; cps <= 3; ++cps)
    {
        run_tile<256, 8, 3, 1, Op>(opname, H, V, cps, 192);
        run_tile<256, 8, 4, 1, Op>(opname, H, V, cps, 192);
        run_tile<256, 16, 3, 1, Op>(opname, H, V, cps, 192);
        run_tile<512, 8, 3, 1, Op>(opname, H, V, cps, 192);
        run_tile<512, 8, 4, 1, Op>(opname, H, V, cps, 192);
        run_tile<512, 4, 4, 1, Op>(opname, H, V, cps, 192);
        run_tile<512, 16, 3, 1, Op>(opname, H, V, cps, 192);
        run_tile<768, 8, 3, 1, Op>(opname, H, V, cps, 192);
        run_tile<992, 8, 3, 1, Op>(opname, H, V, cps, 192);
    }
}

int main(int argc, char** argv)
{
    if (argc > 1) g_n = atoi(argv[1]);
    size_t n = (size_t)g_n * g_n;
    cudaMalloc(&g_in, n * 8); cudaMalloc(&g_out, n * 8); cudaMalloc(&g_coef, 128 * 8);
    fill<<<2048, 256>>>(g_in, n);
    fill<<<1, 128>>>(g_coef, 128);
    cudaDeviceSynchronize();
    printf("== XY 5x5\n");
    for (int rep = 0; rep < 2; ++rep)
    for (int cps = 1; cps <= 2; ++cps)
    {
        run<512, 8, 3, 5, 5, 0, 1>(cps, 192);
        run<512, 4, 4, 5, 5, 0, 1>(cps, 192);
        run<512, 4, 6, 5, 5, 0, 1>(cps, 192);
        run<768, 8, 3, 5, 5, 0, 1>(cps, 192);
        run<768, 4, 4, 5, 5, 0, 1>(cps, 192);
        run<992, 8, 3, 5, 5, 0, 1>(cps, 192);
        run<992, 4, 4, 5, 5, 0, 1>(cps, 192);
        run<992, 4, 6, 5, 5, 0, 1>(cps, 192);
        run<384, 8, 3, 5, 5, 0, 1>(cps, 192);
        run<384, 8, 4, 5, 5, 0, 1>(cps, 192);
        run<256, 8, 3, 5, 5, 0, 2>(cps, 192);
        run<384, 8, 3, 5, 5, 0, 2>(cps, 192);
        run<512, 4, 4, 5, 5, 0, 2>(cps, 192);
        run<128, 8, 3, 5, 5, 0, 2>(cps, 192);
        run<128, 8, 4, 5, 5, 0, 2>(cps, 192);
        run<192, 8, 3, 5, 5, 0, 2>(cps, 192);
    }
    for (int cps = 3; cps <= 4; ++cps)
    {
        run<128, 8, 3, 5, 5, 0, 2>(cps, 192);
        run<128, 4, 4, 5, 5, 0, 2>(cps, 192);
        run<192, 8, 3, 5, 5, 0, 2>(cps, 192);
        run<256, 8, 3, 5, 5, 0, 1>(cps, 192);
    }
    std::sort(results.begin(), results.end(), [](const Res& x, const Res& y) { return x.gpts > y.gpts; });
    printf("\n== top 25\n");
    for (size_t i = 0; i < results.size() && i < 25; ++i) printf("%-70s %8.1f\n", results[i].name, results[i].gpts);
    return 0;
}
