// Cahn-Hilliard ADI, tolerance-mode road: the time step built on the partitioned pentadiagonal solve (pent_part.cu), on one
// GPU or on y-slabs over several (BASELINE.json config 5, 1-8 GPUs).
//
// Reference step (cuPentCahnADI.cu:540-590, BatchHyper.cu:520-590): right-hand side, transpose, x-direction cyclic solve,
// transpose back, y-direction cyclic solve, findNew.  Here, per step and per GPU (grid rows [g rows, (g+1) rows), layout
// c[y][x] throughout - nothing is transposed):
//
//   k_rhs_fused<false>   rhs = L(2c - cOld) - (2/3)(c - cOld) + N(c) in one pass; the 2 halo rows either side come
//                        straight from the neighbouring slabs' field buffers (peer memory)                  24 B/point
//   k_part_cols          x-direction: every 32-row x np-column tile solves its partition on its own          16 B/point
//   k_spike_reduce       the x-direction interface unknowns (4 per partition and row; a banded sum)
//   k_part_rows<CORR>    y-direction partition-local solves; the x-direction correction is applied while a tile is
//                        staged, so the corrected x-solution is never written                                16 B/point
//   k_spike_reduce       y-direction interface unknowns: the only inter-GPU exchange of the solve - each GPU reads the
//                        4 interface values of up to `reach` partitions from either neighbour (3 x 4 x n doubles per
//                        side at n = 4096, np = 128), not an n^2 / G all-to-all
//   k_new                y-correction + findNew: c(t+dt) = (2c - cOld) + w over cOld                         32 B/point
//
// Algorithmic traffic 88 B/point/step.  Ordering between GPUs is two counters per neighbour kept in the slab's own
// block and written by the neighbours (st.release.sys after the producing kernel): "my field of step s is final" (the
// neighbour's next right-hand side waits for it) and "my y-direction interface values of step s are final" (the
// neighbour's k_spike_reduce waits for it).  The write-after-read hazards are implied by the same two signals (see
// enqueue_step).  Everything lives on the device, so pairs of steps replay from a CUDA graph.
#include "cahn_part.h"

#include "../../include/cuSten.h"
#include "../../include/custen_c.h"
#include "cahn_rhs_stream.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace custen_cahn {

static void ck(const char* what) { checkError(what); }

static int g_rhs_stream = 1;   // default for slabs created from now on (custen_cahn_set_rhs_stream)

// ---- kernels -------------------------------------------------------------------------------------------------------------

// y-direction correction + findNew (cuPentCahnADI.cu:89-100): cNew = (2 c - cOld) + (g - (W0 q0 + W1 q1 + V0 q2 + V1 q3)),
// written over cOld.  data: [rows][n] (row = unknown, column = system); a CTA covers 128 systems x 32 rows of one partition.
__global__ void __launch_bounds__(128) k_new(const double* __restrict__ data, const double* __restrict__ q,
                                             const double* __restrict__ wv, const double* __restrict__ cCurr, double* cOldNew,
                                             int rows, int n, int np)
{
    const int gx = blockIdx.x * 128 + threadIdx.x;
    const int gy0 = blockIdx.y * 32;
    if (gx >= n) return;
    const int p = gy0 / np, rbase = gy0 - p * np;
    const double* qp = q + ((size_t)p * 4) * n + gx;
    const double q0 = qp[0], q1 = qp[(size_t)n], q2 = qp[(size_t)2 * n], q3 = qp[(size_t)3 * n];
    const int kmax = min(32, rows - gy0);
#pragma unroll 8
    for (int k = 0; k < kmax; ++k)
    {
        const size_t index = (size_t)(gy0 + k) * n + gx;
        const double2* wp = reinterpret_cast<const double2*>(wv + 4 * (rbase + k));
        const double2 w01 = wp[0], w23 = wp[1];
        double corr = w01.x * q0;
        corr = fma(w01.y, q1, corr);
        corr = fma(w23.x, q2, corr);
        corr = fma(w23.y, q3, corr);
        const double w = data[index] - corr;
        const double cBar = 2.0 * cCurr[index] - cOldNew[index];
        cOldNew[index] = cBar + w;
    }
}

// flag words of a slab (in its block, so that the neighbours can write them)
enum
{
    F_FIELD_UP = 0,    // steps whose new field the slab above has finished
    F_FIELD_DOWN = 1,  // ... the slab below
    F_G_UP = 2,        // steps whose y-direction interface values the slab above has finished
    F_G_DOWN = 3,
    F_STEPS = 4,       // my own finished steps
    F_ERROR = 5,       // neighbour waits that timed out
    F_TIMEOUT_NS = 6,
    F_WORDS = 8
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// wait until both neighbours have published (my finished steps + ahead) in the two words flags[first], flags[first + 1]
__global__ void k_wait(unsigned long long* flags, int first, int ahead)
{
    if (threadIdx.x >= 2) return;
    const unsigned long long want = flags[F_STEPS] + (unsigned long long)ahead;
    const unsigned long long t0 = global_ns(), limit = flags[F_TIMEOUT_NS];
    while (ld_acquire_sys(flags + first + threadIdx.x) < want)
    {
        __nanosleep(200);
        if (global_ns() - t0 > limit)
        {
            atomicAdd(flags + F_ERROR, 1ull);
            break;
        }
    }
}

// publish (my finished steps + ahead) into the neighbours' words; bump = 1: the step is over, count it first
__global__ void k_signal(unsigned long long* flags, unsigned long long* up_word, unsigned long long* down_word, int ahead, int bump)
{
    if (threadIdx.x) return;
    if (bump) flags[F_STEPS] += 1ull;
    const unsigned long long v = flags[F_STEPS] + (unsigned long long)ahead;
    __threadfence_system();
    if (up_word) st_release_sys(up_word, v);
    if (down_word) st_release_sys(down_word, v);
}

// ---- one slab ---------------------------------------------------------------------------------------------------------------

struct PartSlab
{
    int n, rows, rank, world, device;
    int np, P, P_loc, reach;
    double D, gamma, lx, dx, dt, sigL, sigN;
    PartPlan* plan;
    RhsCoef rc;
    char* block;               // one allocation (one IPC handle): field 0 | field 1 | Gy | flag words
    size_t field_bytes, g_bytes;
    double* field[2];
    double* Gy;
    unsigned long long* flags;
    double *work, *Gx, *qx, *qy;
    const double** gptr_x;     // device: {Gx}
    const double** gptr_y;     // device: the ranks' Gy arrays (only the neighbours' and mine are filled in)
    char *up_block, *down_block;
    void* ipc_mapped[2];
    int n_ipc;
    int cur;                   // field[cur] = c(t), field[cur ^ 1] = c(t - dt)
    long steps;
    cudaStream_t stream;
    cudaGraphExec_t gexec;
    int gexec_cur, use_graph;
    int rhs_stream;            // 1: row-streaming right-hand side (k_rhs_stream), 0: the tile kernel (k_rhs_fused)
    cudaEvent_t ev0, ev1;
};

static bool multi(const PartSlab* s) { return s->world > 1; }

static void publish_pointers(PartSlab* s)
{
    std::vector<const double*> g(s->world, nullptr);
    const int up = (s->rank + s->world - 1) % s->world, down = (s->rank + 1) % s->world;
    g[s->rank] = s->Gy;
    if (multi(s))
    {
        g[up] = (const double*)(s->up_block + 2 * s->field_bytes);
        g[down] = (const double*)(s->down_block + 2 * s->field_bytes);
    }
    cudaMemcpy(s->gptr_y, g.data(), g.size() * sizeof(double*), cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();   // the slab's stream does not order against the legacy stream
    ck("cahn slab: pointer table");
}

PartSlab* part_slab_create(int n, int rank, int world, double D, double gamma, double lx, double dt_over_dx, int device,
                           int np_wanted)
{
    if (world < 1 || rank < 0 || rank >= world || n % world) return nullptr;
    const int rows = n / world;
    const int np = part_choose_np(n, np_wanted);
    if (!np || rows % np || !part_solve_supported(n, rows, np) || !part_solve_supported(rows, n, np)) return nullptr;
    PartSlab* s = new PartSlab();
    memset(s, 0, sizeof *s);
    s->n = n;
    s->rows = rows;
    s->rank = rank;
    s->world = world;
    s->device = device;
    s->np = np;
    s->P = n / np;
    s->P_loc = rows / np;
    s->D = D;
    s->gamma = gamma;
    s->lx = lx;
    s->dx = lx / n;
    s->dt = dt_over_dx * s->dx;
    // coefficients as the reference computes them (cuPentCahnADI.cu:389-395, :463-467, :511-513)
    s->sigL = 2.0 * s->dt * D * gamma / (3.0 * (pow(s->dx, 4.0)));
    s->sigN = 2.0 * s->dt * D / (3.0 * (pow(s->dx, 2.0)));
    cudaSetDevice(device);
    ck("cahn slab: set device");
    const double co5[5] = {s->sigL, -4 * s->sigL, 1 + 6 * s->sigL, -4 * s->sigL, s->sigL};
    s->plan = part_plan_create(n, np, co5, true);
    if (!s->plan)
    {
        delete s;
        return nullptr;
    }
    s->reach = part_plan_reach(s->plan);
    if (world > 1 && s->reach > s->P_loc)   // the interface coupling would reach past the nearest neighbour
    {
        part_plan_destroy(s->plan);
        delete s;
        return nullptr;
    }
    {
        const double L = s->sigL, Nn = s->sigN;
        const double wl[25] = {0.0, 0.0, -1.0 * L, 0.0, 0.0,
                               0.0, -2.0 * L, 8.0 * L, -2.0 * L, 0.0,
                               -1.0 * L, 8.0 * L, -20.0 * L, 8.0 * L, -1.0 * L,
                               0.0, -2.0 * L, 8.0 * L, -2.0 * L, 0.0,
                               0.0, 0.0, -1.0 * L, 0.0, 0.0};
        const double cn[9] = {0.0, 1.0 * Nn, 0.0, 1.0 * Nn, -4.0 * Nn, 1.0 * Nn, 0.0, 1.0 * Nn, 0.0};
        for (int i = 0; i < 25; ++i) s->rc.wl[i] = wl[i];
        for (int i = 0; i < 9; ++i) s->rc.cn[i] = cn[i];
    }
    const size_t N = (size_t)n * rows;
    s->field_bytes = ((N * sizeof(double)) + 255) & ~(size_t)255;
    s->g_bytes = (((size_t)4 * s->P_loc * n * sizeof(double)) + 255) & ~(size_t)255;
    const size_t total = 2 * s->field_bytes + s->g_bytes + F_WORDS * sizeof(unsigned long long);
    cudaMalloc(&s->block, total);
    ck("cahn slab: allocate fields");
    cudaMemset(s->block, 0, total);
    s->field[0] = (double*)s->block;
    s->field[1] = (double*)(s->block + s->field_bytes);
    s->Gy = (double*)(s->block + 2 * s->field_bytes);
    s->flags = (unsigned long long*)(s->block + 2 * s->field_bytes + s->g_bytes);
    const unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
    cudaMemcpy(s->flags + F_TIMEOUT_NS, &timeout_ns, sizeof timeout_ns, cudaMemcpyHostToDevice);
    cudaMalloc(&s->work, N * sizeof(double));
    cudaMalloc(&s->Gx, (size_t)4 * s->P * rows * sizeof(double));
    cudaMalloc(&s->qx, (size_t)4 * s->P * rows * sizeof(double));
    cudaMalloc(&s->qy, (size_t)4 * s->P_loc * n * sizeof(double));
    cudaMalloc(&s->gptr_x, sizeof(double*));
    cudaMalloc(&s->gptr_y, (size_t)world * sizeof(double*));
    ck("cahn slab: allocate work arrays");
    cudaMemcpy(s->gptr_x, &s->Gx, sizeof(double*), cudaMemcpyHostToDevice);
    s->up_block = s->down_block = s->block;
    publish_pointers(s);
    cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    cudaEventCreate(&s->ev0);
    cudaEventCreate(&s->ev1);
    s->use_graph = 1;
    s->rhs_stream = g_rhs_stream;
    cudaFuncSetAttribute(k_rhs_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM);
    cudaDeviceSynchronize();
    ck("cahn slab: create");
    return s;
}

void part_slab_connect(PartSlab* s, char* up, char* down)
{
    cudaSetDevice(s->device);
    s->up_block = up ? up : s->block;
    s->down_block = down ? down : s->block;
    publish_pointers(s);
    if (s->gexec)   // captured with the old neighbours
    {
        cudaGraphExecDestroy(s->gexec);
        s->gexec = nullptr;
    }
}

static unsigned long long* word_of(const PartSlab* s, char* blk, int which)
{
    return (unsigned long long*)(blk + 2 * s->field_bytes + s->g_bytes) + which;
}

// One time step on s->stream.  Orderings between neighbouring slabs (all through k_wait / k_signal):
//   read-after-write   rhs(s) reads the neighbours' c(t) halo rows            <- their "field final" signal of step s-1
//                      reduce_y(s) reads the neighbours' Gy of step s         <- their "interface values final" signal
//   write-after-read   k_new(s) overwrites the rows their rhs(s) reads: it runs after my reduce_y(s), which waited for
//                      their Gy signal of step s, issued after their rhs(s);
//                      part_rows(s+1) overwrites the Gy their reduce_y(s) reads: it runs after my rhs(s+1), which waited
//                      for their field signal of step s, issued after their reduce_y(s).
static void enqueue_step(PartSlab* s)
{
    cudaStream_t st = s->stream;
    const int n = s->n, rows = s->rows, np = s->np;
    double* c = s->field[s->cur];
    double* cOld = s->field[s->cur ^ 1];
    RhsHalo halo = {nullptr, nullptr, nullptr, nullptr};
    if (multi(s))
    {
        const size_t cur_off = (size_t)s->cur * s->field_bytes, old_off = (size_t)(s->cur ^ 1) * s->field_bytes;
        halo.c_up = (const double*)(s->up_block + cur_off) + (size_t)(rows - 2) * n;
        halo.o_up = (const double*)(s->up_block + old_off) + (size_t)(rows - 2) * n;
        halo.c_down = (const double*)(s->down_block + cur_off);
        halo.o_down = (const double*)(s->down_block + old_off);
        k_wait<<<1, 32, 0, st>>>(s->flags, F_FIELD_UP, 0);
    }
    if (s->rhs_stream)
    {
        dim3 sg((n + RS_W - 1) / RS_W, rows / RS_BR);
        k_rhs_stream<<<sg, RS_NT, RS_SMEM, st>>>(cOld, c, halo, s->work, n, rows, s->rc);
    }
    else
    {
        dim3 tg(n / 32, rows / 32);
        k_rhs_fused<false><<<tg, 128, 0, st>>>(cOld, c, halo, s->work, n, rows, s->rc);
    }
    part_solve_cols(s->plan, s->work, rows, n, s->Gx, st);
    part_reduce(s->plan, s->gptr_x, 1, 0, s->P, rows, s->qx, st);
    part_solve_rows(s->plan, s->work, n, rows, s->Gy, s->qx, rows, st);
    if (multi(s))
    {
        // I am the slab below my upper neighbour and the slab above my lower one
        k_signal<<<1, 32, 0, st>>>(s->flags, word_of(s, s->up_block, F_G_DOWN), word_of(s, s->down_block, F_G_UP), 1, 0);
        k_wait<<<1, 32, 0, st>>>(s->flags, F_G_UP, 1);
    }
    part_reduce(s->plan, s->gptr_y, s->world, s->rank, s->P_loc, n, s->qy, st);
    dim3 ng(n / 128 + (n % 128 != 0), rows / 32);
    k_new<<<ng, 128, 0, st>>>(s->work, s->qy, part_plan_wv(s->plan), c, cOld, rows, n, np);
    if (multi(s))
        k_signal<<<1, 32, 0, st>>>(s->flags, word_of(s, s->up_block, F_FIELD_DOWN), word_of(s, s->down_block, F_FIELD_UP), 0, 1);
    s->cur ^= 1;
    s->steps++;
}

// two steps (after which the field buffers are back in their roles) as an executable graph
static bool step_graph(PartSlab* s)
{
    if (s->gexec) return true;
    const int cur0 = s->cur;
    const long steps0 = s->steps;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    enqueue_step(s);
    enqueue_step(s);
    const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    s->cur = cur0;
    s->steps = steps0;
    if (e != cudaSuccess || !graph)
    {
        cudaGetLastError();
        return false;
    }
    if (cudaGraphInstantiate(&s->gexec, graph, 0) != cudaSuccess)
    {
        cudaGetLastError();
        s->gexec = nullptr;
        cudaGraphDestroy(graph);
        return false;
    }
    cudaGraphDestroy(graph);
    s->gexec_cur = cur0;
    return true;
}

void part_slab_step(PartSlab* s, int nsteps)
{
    cudaSetDevice(s->device);
    for (int it = 0; it < nsteps;)
    {
        // the first step runs kernel by kernel (the shared-memory opt-ins are not capturable); pairs of steps replay
        // from the graph when the field buffers are in the roles it was captured with
        if (s->use_graph && s->steps > 0 && it + 1 < nsteps && (s->gexec || step_graph(s)) && s->cur == s->gexec_cur)
        {
            cudaGraphLaunch(s->gexec, s->stream);
            s->steps += 2;
            it += 2;
            continue;
        }
        enqueue_step(s);
        ++it;
    }
    ck("cahn slab: step");
}

void part_slab_set_graph(PartSlab* s, int on) { s->use_graph = on; }

// All copies go through the slab's own stream (a non-blocking stream does not order against the legacy stream, and a
// plain cudaMemcpy may return before a device-to-device or pageable-host copy has landed).
void part_slab_set_fields(PartSlab* s, const double* c_rows_host, const double* cold_rows_host)
{
    cudaSetDevice(s->device);
    const size_t bytes = (size_t)s->n * s->rows * sizeof(double);
    cudaMemcpyAsync(s->field[0], c_rows_host, bytes, cudaMemcpyHostToDevice, s->stream);
    cudaMemcpyAsync(s->field[1], cold_rows_host ? cold_rows_host : c_rows_host, bytes, cudaMemcpyHostToDevice, s->stream);
    cudaStreamSynchronize(s->stream);
    s->cur = 0;
    ck("cahn slab: set field");
}

void part_slab_load_device(PartSlab* s, const double* c_dev, const double* cold_dev)
{
    cudaSetDevice(s->device);
    const size_t bytes = (size_t)s->n * s->rows * sizeof(double);
    cudaMemcpyAsync(s->field[0], c_dev, bytes, cudaMemcpyDeviceToDevice, s->stream);
    cudaMemcpyAsync(s->field[1], cold_dev, bytes, cudaMemcpyDeviceToDevice, s->stream);
    cudaStreamSynchronize(s->stream);
    s->cur = 0;
    ck("cahn slab: load fields");
}

void part_slab_store_device(PartSlab* s, double* c_dev, double* cold_dev)
{
    cudaSetDevice(s->device);
    const size_t bytes = (size_t)s->n * s->rows * sizeof(double);
    cudaMemcpyAsync(c_dev, s->field[s->cur], bytes, cudaMemcpyDeviceToDevice, s->stream);
    cudaMemcpyAsync(cold_dev, s->field[s->cur ^ 1], bytes, cudaMemcpyDeviceToDevice, s->stream);
    cudaStreamSynchronize(s->stream);
    ck("cahn slab: store fields");
}

void part_slab_get_field(PartSlab* s, double* rows_host)
{
    cudaSetDevice(s->device);
    cudaMemcpyAsync(rows_host, s->field[s->cur], (size_t)s->n * s->rows * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    cudaStreamSynchronize(s->stream);
    ck("cahn slab: get field");
}

void part_slab_synchronize(PartSlab* s)
{
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    ck("cahn slab: synchronize");
}

cudaStream_t part_slab_stream(PartSlab* s) { return s->stream; }
int part_slab_np(const PartSlab* s) { return s->np; }
long part_slab_steps(const PartSlab* s) { return s->steps; }

int part_slab_error(PartSlab* s)
{
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    unsigned long long e = 0;
    cudaMemcpy(&e, s->flags + F_ERROR, sizeof e, cudaMemcpyDeviceToHost);
    return (int)e;
}

void part_slab_set_timeout(PartSlab* s, double seconds)
{
    cudaSetDevice(s->device);
    const unsigned long long ns = (unsigned long long)(seconds * 1e9);
    cudaStreamSynchronize(s->stream);
    cudaMemcpy(s->flags + F_TIMEOUT_NS, &ns, sizeof ns, cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
}

float part_slab_time_steps(PartSlab* s, int nsteps)
{
    cudaSetDevice(s->device);
    cudaEventRecord(s->ev0, s->stream);
    part_slab_step(s, nsteps);
    cudaEventRecord(s->ev1, s->stream);
    cudaEventSynchronize(s->ev1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    ck("cahn slab: time steps");
    return ms;
}

char* part_slab_block(PartSlab* s) { return s->block; }
int part_slab_device(const PartSlab* s) { return s->device; }

void part_slab_destroy(PartSlab* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    if (s->gexec) cudaGraphExecDestroy(s->gexec);
    for (int k = 0; k < s->n_ipc; ++k) cudaIpcCloseMemHandle(s->ipc_mapped[k]);
    for (void* p : {(void*)s->block, (void*)s->work, (void*)s->Gx, (void*)s->qx, (void*)s->qy, (void*)s->gptr_x, (void*)s->gptr_y})
        if (p) cudaFree(p);
    cudaEventDestroy(s->ev0);
    cudaEventDestroy(s->ev1);
    cudaStreamDestroy(s->stream);
    part_plan_destroy(s->plan);
    cudaGetLastError();
    delete s;
}

// neighbours in other processes: map their blocks through CUDA IPC (up == down when world == 2: mapped once)
void part_slab_connect_ipc(PartSlab* s, const void* up_handle64, const void* down_handle64)
{
    cudaSetDevice(s->device);
    char* mapped[2] = {nullptr, nullptr};
    const void* hs[2] = {up_handle64, down_handle64};
    for (int k = 0; k < 2; ++k)
    {
        if (!hs[k]) continue;
        if (k == 1 && hs[0] && memcmp(hs[0], hs[1], sizeof(cudaIpcMemHandle_t)) == 0)
        {
            mapped[1] = mapped[0];
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[k], sizeof h);
        void* p = nullptr;
        cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        ck("cahn slab: open the neighbour's memory (CUDA IPC)");
        mapped[k] = (char*)p;
        s->ipc_mapped[s->n_ipc++] = p;
    }
    part_slab_connect(s, mapped[0], mapped[1]);
}

}  // namespace custen_cahn

// ---- C ABI ---------------------------------------------------------------------------------------------------------------------
using namespace custen_cahn;

namespace {

int g_np_default = 128;

struct MultiGpu
{
    int n, world;
    std::vector<PartSlab*> slab;
};

}  // namespace

namespace custen_cahn {
void part_set_default_np(int np) { g_np_default = np > 0 ? np : 128; }
void part_set_rhs_stream(int on) { g_rhs_stream = on; }
int part_default_np() { return g_np_default; }
}  // namespace custen_cahn

extern "C" {

// tuning / tests: 1 (default) the row-streaming right-hand-side kernel, 0 the tile kernel; for solvers created afterwards
void custen_cahn_set_rhs_stream(int on) { part_set_rhs_stream(on); }

// One y-slab of the tolerance-mode solver per process / GPU.  Returns NULL when the partitioned layout cannot take the
// grid (n / world not a multiple of the partition height 32 / 64 / 128 / 256, or the interface coupling would reach
// past the nearest neighbour).
void* custen_cahn_slab_create(int nx, int rank, int world, double D, double gamma, double lx, double dt_over_dx, int device)
{
    return part_slab_create(nx, rank, world, D, gamma, lx, dt_over_dx, device, g_np_default);
}

void custen_cahn_slab_export(void* h, void* handle64)
{
    PartSlab* s = (PartSlab*)h;
    cudaSetDevice(part_slab_device(s));
    cudaIpcMemHandle_t ipc;
    cudaIpcGetMemHandle(&ipc, part_slab_block(s));
    checkError("cahn slab: export (CUDA IPC)");
    memcpy(handle64, &ipc, sizeof ipc);
}

void custen_cahn_slab_connect(void* h, const void* up_handle64, const void* down_handle64)
{
    part_slab_connect_ipc((PartSlab*)h, up_handle64, down_handle64);
}

// neighbours in the same process: peer access from this slab's GPU to theirs
void custen_cahn_slab_connect_local(void* h, void* up, void* down)
{
    PartSlab* s = (PartSlab*)h;
    cudaSetDevice(part_slab_device(s));
    for (void* o : {up, down})
    {
        if (!o || o == h) continue;
        const int od = part_slab_device((PartSlab*)o);
        if (od == part_slab_device(s)) continue;
        const cudaError_t e = cudaDeviceEnablePeerAccess(od, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) checkError("cahn slab: enable peer access");
        cudaGetLastError();
    }
    part_slab_connect(s, up ? part_slab_block((PartSlab*)up) : nullptr, down ? part_slab_block((PartSlab*)down) : nullptr);
}

void custen_cahn_slab_set_field(void* h, const double* rows_host) { part_slab_set_fields((PartSlab*)h, rows_host, nullptr); }
void custen_cahn_slab_set_fields(void* h, const double* c_rows_host, const double* cold_rows_host)
{
    part_slab_set_fields((PartSlab*)h, c_rows_host, cold_rows_host);
}
void custen_cahn_slab_get_field(void* h, double* rows_host) { part_slab_get_field((PartSlab*)h, rows_host); }
void custen_cahn_slab_step(void* h, int nsteps) { part_slab_step((PartSlab*)h, nsteps); }
float custen_cahn_slab_time_steps(void* h, int nsteps) { return part_slab_time_steps((PartSlab*)h, nsteps); }
void custen_cahn_slab_synchronize(void* h) { part_slab_synchronize((PartSlab*)h); }
int custen_cahn_slab_error(void* h) { return part_slab_error((PartSlab*)h); }
void custen_cahn_slab_set_timeout(void* h, double seconds) { part_slab_set_timeout((PartSlab*)h, seconds); }
void custen_cahn_slab_set_graph(void* h, int on) { part_slab_set_graph((PartSlab*)h, on); }
int custen_cahn_slab_partition_rows(void* h) { return part_slab_np((PartSlab*)h); }
void custen_cahn_slab_destroy(void* h) { part_slab_destroy((PartSlab*)h); }

// ---- one process driving several GPUs ---------------------------------------------------------------------------------
void* custen_cahn_mg_create(int nx, int ngpus, const int* devices, double D, double gamma, double lx, double dt_over_dx)
{
    MultiGpu* m = new MultiGpu();
    m->n = nx;
    m->world = ngpus;
    for (int g = 0; g < ngpus; ++g)
    {
        PartSlab* s = part_slab_create(nx, g, ngpus, D, gamma, lx, dt_over_dx, devices ? devices[g] : g, g_np_default);
        if (!s)
        {
            for (PartSlab* o : m->slab) part_slab_destroy(o);
            delete m;
            return nullptr;
        }
        m->slab.push_back(s);
    }
    if (ngpus > 1)
        for (int g = 0; g < ngpus; ++g)
            custen_cahn_slab_connect_local(m->slab[g], m->slab[(g + ngpus - 1) % ngpus], m->slab[(g + 1) % ngpus]);
    return m;
}

void custen_cahn_mg_set_field(void* h, const double* c0_host)
{
    MultiGpu* m = (MultiGpu*)h;
    const size_t per = (size_t)m->n * (m->n / m->world);
    for (int g = 0; g < m->world; ++g) part_slab_set_fields(m->slab[g], c0_host + g * per, nullptr);
}

void custen_cahn_mg_get_field(void* h, double* out_host)
{
    MultiGpu* m = (MultiGpu*)h;
    const size_t per = (size_t)m->n * (m->n / m->world);
    for (int g = 0; g < m->world; ++g) part_slab_get_field(m->slab[g], out_host + g * per);
}

// asynchronous: every GPU's steps are enqueued on its own stream; the slabs order themselves on the device
void custen_cahn_mg_step(void* h, int nsteps)
{
    MultiGpu* m = (MultiGpu*)h;
    for (PartSlab* s : m->slab) part_slab_step(s, nsteps);
}

// milliseconds for nsteps steps: the slowest GPU's CUDA-event time
float custen_cahn_mg_time_steps(void* h, int nsteps)
{
    MultiGpu* m = (MultiGpu*)h;
    for (PartSlab* s : m->slab) part_slab_synchronize(s);
    for (PartSlab* s : m->slab)
    {
        cudaSetDevice(part_slab_device(s));
        cudaEventRecord(s->ev0, s->stream);
        part_slab_step(s, nsteps);
        cudaEventRecord(s->ev1, s->stream);
    }
    float worst = 0.0f;
    for (PartSlab* s : m->slab)
    {
        cudaSetDevice(part_slab_device(s));
        cudaEventSynchronize(s->ev1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, s->ev0, s->ev1);
        worst = ms > worst ? ms : worst;
    }
    checkError("cahn mg: time steps");
    return worst;
}

int custen_cahn_mg_error(void* h)
{
    MultiGpu* m = (MultiGpu*)h;
    int e = 0;
    for (PartSlab* s : m->slab) e += part_slab_error(s);
    return e;
}

void custen_cahn_mg_set_graph(void* h, int on)
{
    for (PartSlab* s : ((MultiGpu*)h)->slab) part_slab_set_graph(s, on);
}

void custen_cahn_mg_destroy(void* h)
{
    MultiGpu* m = (MultiGpu*)h;
    for (PartSlab* s : m->slab) part_slab_destroy(s);
    delete m;
}

}  // extern "C"
