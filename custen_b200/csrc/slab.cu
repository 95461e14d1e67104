// Multi-GPU y-slab layer in C (new; the reference is single-GPU).
//
// The global ny x nx grid is cut into `world` contiguous y-slabs; slab g owns rows [g ny/world, (g+1) ny/world) of BOTH
// field buffers (input and output, which trade roles at every Swap, like the reference's cuStenSwap2D*,
// cuSten/src/struct/custenCreateDestroy2DXYp.cu:253-310).  The T rows above and B rows below a slab are the reference's
// boundaryTop / boundaryBottom kernel arguments (cuSten/src/kernels/2d_xy_p_kernel.cu:67-68) pointed at the NEIGHBOUR
// GPU'S MEMORY: the sweep's own TMA producer pulls them over NVLink, there is no exchange step and no halo buffer.
//
// Time stepping (Compute, Swap, Compute, ...) needs two orderings per neighbour and sweep:
//   * my sweep s reads the neighbour's edge rows, which are the output of ITS sweep s-1;
//   * my sweep s overwrites the buffer whose edge rows the neighbour read during ITS sweep s-1.
// Both are the same condition - "the neighbour has finished s sweeps" - so every slab keeps one counter, published into
// its neighbours' memory by the last CTA of each sweep (st.release.sys), and the sweep kernel's producer warp checks the
// neighbour's counter only in front of the work items that touch those rows, which it walks last
// (stream_kernels.cuh: slab_wait / slab_epilogue / chunk_of).  Interior rows never wait; a step is ONE kernel launch,
// and because the counters live on the device, steps replay from a CUDA graph.
//
// Two ways to connect slabs:
//   one process per GPU   custen_slab_export -> (the caller moves 64-byte IPC handles: MPI, torch.distributed, a file)
//                         -> custen_slab_connect;
//   one process, G GPUs   custen_mg_create / run / scatter / gather: the C entry point for existing cuSten programs
//                         (SURVEY.md section 7 step 5), peers reached through cudaDeviceEnablePeerAccess.
#include "../../include/custen_c.h"
#include "plan.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace custen;

namespace {

const Spec kSpecs[12] = {{DIR_X, 1, 0, 0}, {DIR_X, 0, 0, 0}, {DIR_X, 1, 1, 0}, {DIR_X, 0, 1, 0},
                         {DIR_Y, 1, 0, 0}, {DIR_Y, 0, 0, 0}, {DIR_Y, 1, 1, 0}, {DIR_Y, 0, 1, 0},
                         {DIR_XY, 1, 0, 0}, {DIR_XY, 0, 0, 0}, {DIR_XY, 1, 1, 0}, {DIR_XY, 0, 1, 0}};

constexpr int kFlagWords = 8;   // [0] written by the slab above, [1] by the slab below, [2..5] = Band::sync_local

struct Slab
{
    Spec spec;
    int device, rank, world;
    int nx, rows, ny;
    int T, B;
    char* block;              // one allocation (one IPC handle): field 0 | field 1 | counters
    size_t field_bytes;
    double* field[2];
    unsigned long long* flags;
    double* d_coef;
    int cur;                  // field[cur] is the input of the next sweep
    long sweeps;
    cuSten_t handle;
    char* up_block;           // the neighbours' blocks as this process sees them (nullptr: physical edge)
    char* down_block;
    void* ipc_mapped[2];
    int n_ipc;
    cudaStream_t stream;      // launches go here (the handle's compute stream unless the owner provides one)
    cudaGraphExec_t gexec;
    int gexec_cur;
    int graph_ok;
    cudaEvent_t ev0, ev1;
};

void ck(const char* what) { checkError(what); }

bool needs_halo(const Slab* s) { return s->spec.dir != DIR_X && (s->T || s->B); }

void refresh_pointers(Slab* s)
{
    Plan* p = plan_of(&s->handle);
    if (!needs_halo(s)) return;
    const size_t row = (size_t)s->nx;
    const double* up_in = s->up_block ? (const double*)(s->up_block + (size_t)s->cur * s->field_bytes) : nullptr;
    const double* down_in = s->down_block ? (const double*)(s->down_block + (size_t)s->cur * s->field_bytes) : nullptr;
    p->slab_enabled = 1;
    p->slab_top = up_in ? up_in + (size_t)(s->rows - s->T) * row : nullptr;
    p->slab_bottom = down_in;
    p->slab_first = s->up_block == nullptr;
    p->slab_last = s->down_block == nullptr;
}

void connect(Slab* s, char* up, char* down)
{
    s->up_block = up;
    s->down_block = down;
    Plan* p = plan_of(&s->handle);
    const bool self_only = (up == nullptr || up == s->block) && (down == nullptr || down == s->block);
    if (needs_halo(s) && !self_only)
    {
        p->sync_local = s->flags + 2;
        p->sync_wait_up = up ? s->flags + 0 : nullptr;
        p->sync_wait_down = down ? s->flags + 1 : nullptr;
        // I am the slab below my upper neighbour (its word 1) and the slab above my lower neighbour (its word 0)
        p->sync_signal_up = up ? (unsigned long long*)(up + 2 * s->field_bytes) + 1 : nullptr;
        p->sync_signal_down = down ? (unsigned long long*)(down + 2 * s->field_bytes) + 0 : nullptr;
    }
    else
        p->sync_local = nullptr;
    refresh_pointers(s);
}

Slab* slab_new(int variant, int device, int rank, int world, int nx, int ny_global, const double* coef_host, int ncoef,
               int H, int L, int R, int V, int T, int B, double* func, cudaStream_t shared_stream)
{
    if (variant < 0 || variant >= 12 || world < 1 || ny_global % world)
    {
        printf("\ncuSten slab: bad variant, or ny not divisible by the number of slabs\nprogram terminated ...\n\n");
        exit(EXIT_FAILURE);
    }
    Slab* s = new Slab();
    memset(s, 0, sizeof *s);
    s->spec = kSpecs[variant];
    s->device = device;
    s->rank = rank;
    s->world = world;
    s->nx = nx;
    s->ny = ny_global;
    s->rows = ny_global / world;
    if (s->spec.dir == DIR_X) { V = 1; T = B = 0; }
    if (s->spec.dir == DIR_Y) { H = 1; L = R = 0; }
    s->T = T;
    s->B = B;
    cudaSetDevice(device);
    ck("slab: set device");
    s->field_bytes = (((size_t)nx * s->rows * sizeof(double)) + 255) & ~(size_t)255;
    cudaMalloc(&s->block, 2 * s->field_bytes + kFlagWords * sizeof(unsigned long long));
    ck("slab: allocate fields");
    cudaMemset(s->block, 0, 2 * s->field_bytes + kFlagWords * sizeof(unsigned long long));
    s->field[0] = (double*)s->block;
    s->field[1] = (double*)(s->block + s->field_bytes);
    s->flags = (unsigned long long*)(s->block + 2 * s->field_bytes);
    const unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
    cudaMemcpy(s->flags + 5, &timeout_ns, sizeof timeout_ns, cudaMemcpyHostToDevice);
    const int nc = ncoef > 0 ? ncoef : H * V;
    cudaMalloc(&s->d_coef, (size_t)nc * sizeof(double));
    cudaMemcpy(s->d_coef, coef_host, (size_t)nc * sizeof(double), cudaMemcpyHostToDevice);
    ck("slab: upload coefficients");
    plan_create(&s->handle, s->spec, device, 1, nx, s->rows, 32, 8, s->field[1], s->field[0], s->d_coef, H, L, R, V, T, B,
                nc, s->spec.fun ? func : nullptr);
    s->stream = shared_stream ? shared_stream : s->handle.streams[0];
    s->graph_ok = shared_stream ? 0 : 1;
    cudaEventCreate(&s->ev0);
    cudaEventCreate(&s->ev1);
    cudaDeviceSynchronize();
    ck("slab: create");
    // until connected: a periodic slab alone in the world wraps onto itself, a non-periodic one has no neighbours
    if (world == 1) connect(s, s->spec.periodic ? s->block : nullptr, s->spec.periodic ? s->block : nullptr);
    return s;
}

void slab_compute(Slab* s)
{
    cudaSetDevice(s->device);
    refresh_pointers(s);
    plan_launch_slab(&s->handle, s->stream);
    s->sweeps++;
}

void slab_swap(Slab* s)
{
    plan_swap(&s->handle, s->field[s->cur ^ 1]);
    s->cur ^= 1;
}

// two steps (after which the buffers are back in their roles) as an executable graph
bool slab_graph(Slab* s)
{
    if (!s->graph_ok) return false;
    if (s->gexec) return true;
    cudaGraph_t graph = nullptr;
    const int cur0 = s->cur;
    const long sweeps0 = s->sweeps;
    if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess)
    {
        cudaGetLastError();
        s->graph_ok = 0;
        return false;
    }
    for (int k = 0; k < 2; ++k)
    {
        slab_compute(s);
        slab_swap(s);
    }
    const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    launches_add((uint64_t)-2);   // nothing ran: capture only records
    s->sweeps = sweeps0;
    if (s->cur != cur0) slab_swap(s);
    if (e != cudaSuccess || !graph || cudaGraphInstantiate(&s->gexec, graph, 0) != cudaSuccess)
    {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        s->gexec = nullptr;
        s->graph_ok = 0;
        return false;
    }
    cudaGraphDestroy(graph);
    s->gexec_cur = cur0;
    return true;
}

void slab_run(Slab* s, int nsteps, bool use_graph)
{
    cudaSetDevice(s->device);
    for (int it = 0; it < nsteps; ++it)
    {
        // pairs of steps replay from the graph once one step has run plainly (first-use set-up is not capturable)
        if (use_graph && s->sweeps > 0 && it + 1 < nsteps && slab_graph(s) && s->cur == s->gexec_cur)
        {
            cudaGraphLaunch(s->gexec, s->stream);
            launches_add(2);
            s->sweeps += 2;
            ++it;
            continue;
        }
        slab_compute(s);
        slab_swap(s);
    }
    ck("slab: run");
}

void slab_free(Slab* s)
{
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    if (s->gexec) cudaGraphExecDestroy(s->gexec);
    for (int k = 0; k < s->n_ipc; ++k) cudaIpcCloseMemHandle(s->ipc_mapped[k]);
    cudaGetLastError();
    cudaEventDestroy(s->ev0);
    cudaEventDestroy(s->ev1);
    plan_destroy(&s->handle);
    cudaFree(s->d_coef);
    cudaFree(s->block);
    delete s;
}

struct Multi
{
    std::vector<Slab*> slabs;
    std::vector<cudaStream_t> shared;   // one stream per device that carries more than one slab
    int nx, ny;
    bool distinct;
};

}  // namespace

extern "C" {

void* custen_slab_create(int variant, int device, int rank, int world, int nx, int ny_global, const double* coef_host,
                         int ncoef, int H, int L, int R, int V, int T, int B, double* func)
{
    return slab_new(variant, device, rank, world, nx, ny_global, coef_host, ncoef, H, L, R, V, T, B, func, nullptr);
}

void custen_slab_export(void* slab, void* handle64, size_t* offset_out)
{
    Slab* s = (Slab*)slab;
    cudaSetDevice(s->device);
    custen_ipc_export(s->block, handle64, offset_out);
}

void custen_slab_connect(void* slab, const void* up_handle64, size_t up_offset, const void* down_handle64, size_t down_offset)
{
    Slab* s = (Slab*)slab;
    cudaSetDevice(s->device);
    char* up = nullptr;
    char* down = nullptr;
    if (up_handle64)
    {
        s->ipc_mapped[s->n_ipc] = custen_ipc_open(up_handle64);
        up = (char*)s->ipc_mapped[s->n_ipc++] + up_offset;
    }
    if (down_handle64)
    {
        if (up_handle64 && !memcmp(up_handle64, down_handle64, 64)) down = (char*)s->ipc_mapped[0] + down_offset;  // two slabs: one peer
        else
        {
            s->ipc_mapped[s->n_ipc] = custen_ipc_open(down_handle64);
            down = (char*)s->ipc_mapped[s->n_ipc++] + down_offset;
        }
    }
    connect(s, up, down);
}

double* custen_slab_field(void* slab, int which)
{
    Slab* s = (Slab*)slab;
    return s->field[which ? s->cur ^ 1 : s->cur];
}

int custen_slab_rows(void* slab) { return ((Slab*)slab)->rows; }

void custen_slab_compute(void* slab)
{
    slab_compute((Slab*)slab);
    ck("slab: compute");
}
void custen_slab_swap(void* slab) { slab_swap((Slab*)slab); }
void custen_slab_run(void* slab, int nsteps) { slab_run((Slab*)slab, nsteps, true); }
void custen_slab_run_plain(void* slab, int nsteps) { slab_run((Slab*)slab, nsteps, false); }

float custen_slab_time_run(void* slab, int nsteps)
{
    Slab* s = (Slab*)slab;
    cudaSetDevice(s->device);
    cudaEventRecord(s->ev0, s->stream);
    slab_run(s, nsteps, true);
    cudaEventRecord(s->ev1, s->stream);
    cudaEventSynchronize(s->ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    ck("slab: timing");
    return ms;
}

void custen_slab_synchronize(void* slab)
{
    Slab* s = (Slab*)slab;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    ck("slab: synchronize");
}

// 0: fine; 1: a wait for a neighbour timed out (results are not to be trusted)
int custen_slab_error(void* slab)
{
    Slab* s = (Slab*)slab;
    cudaSetDevice(s->device);
    unsigned long long v = 0;
    cudaMemcpy(&v, s->flags + 4, sizeof v, cudaMemcpyDeviceToHost);
    ck("slab: read status");
    return v != 0;
}

void custen_slab_set_timeout(void* slab, double seconds)
{
    Slab* s = (Slab*)slab;
    cudaSetDevice(s->device);
    const unsigned long long ns = (unsigned long long)(seconds * 1e9);
    cudaMemcpy(s->flags + 5, &ns, sizeof ns, cudaMemcpyHostToDevice);
}

int custen_slab_last_path(void* slab) { return plan_of(&((Slab*)slab)->handle)->last_path; }

void custen_slab_destroy(void* slab) { slab_free((Slab*)slab); }

// ---- one process, several GPUs ------------------------------------------------------------------------------------

void* custen_mg_create(int ndev, const int* devices, int variant, int nx, int ny, const double* coef_host, int ncoef, int H,
                       int L, int R, int V, int T, int B, const char* builtin_fun, double* const* funcs)
{
    Multi* m = new Multi();
    m->nx = nx;
    m->ny = ny;
    m->distinct = true;
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j) m->distinct = m->distinct && devices[i] != devices[j];
    // peers see each other's memory directly
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < ndev; ++j)
            if (devices[i] != devices[j])
            {
                cudaSetDevice(devices[i]);
                cudaDeviceEnablePeerAccess(devices[j], 0);
                cudaGetLastError();   // already enabled is fine
            }
    for (int i = 0; i < ndev; ++i)
    {
        cudaSetDevice(devices[i]);
        cudaStream_t shared = nullptr;
        if (!m->distinct)
        {
            // slabs that share a GPU share a stream: their sweeps then run in submission order and never spin on each other
            for (int j = 0; j < i; ++j)
                if (devices[j] == devices[i]) shared = m->slabs[j]->stream;
            if (!shared)
            {
                cudaStreamCreate(&shared);
                m->shared.push_back(shared);
            }
        }
        double* func = nullptr;
        if (kSpecs[variant].fun) func = builtin_fun ? custen_builtin_fun(builtin_fun) : (funcs ? funcs[i] : nullptr);
        m->slabs.push_back(slab_new(variant, devices[i], i, ndev, nx, ny, coef_host, ncoef, H, L, R, V, T, B, func, shared));
    }
    const bool periodic = kSpecs[variant].periodic != 0;
    for (int i = 0; i < ndev; ++i)
    {
        Slab* up = i > 0 ? m->slabs[i - 1] : (periodic ? m->slabs[ndev - 1] : nullptr);
        Slab* down = i < ndev - 1 ? m->slabs[i + 1] : (periodic ? m->slabs[0] : nullptr);
        connect(m->slabs[i], up ? up->block : nullptr, down ? down->block : nullptr);
    }
    return m;
}

void custen_mg_scatter(void* mg, const double* host_field)
{
    Multi* m = (Multi*)mg;
    for (Slab* s : m->slabs)
    {
        cudaSetDevice(s->device);
        cudaMemcpy(s->field[s->cur], host_field + (size_t)s->rank * s->rows * s->nx, (size_t)s->rows * s->nx * sizeof(double),
                   cudaMemcpyHostToDevice);
    }
    ck("mg: scatter");
}

// which = 0: the current input field (after run: the latest result), 1: the current output buffer
void custen_mg_gather(void* mg, double* host_field, int which)
{
    Multi* m = (Multi*)mg;
    for (Slab* s : m->slabs)
    {
        cudaSetDevice(s->device);
        cudaDeviceSynchronize();
    }
    for (Slab* s : m->slabs)
    {
        cudaSetDevice(s->device);
        cudaMemcpy(host_field + (size_t)s->rank * s->rows * s->nx, s->field[which ? s->cur ^ 1 : s->cur],
                   (size_t)s->rows * s->nx * sizeof(double), cudaMemcpyDeviceToHost);
    }
    ck("mg: gather");
}

void custen_mg_fill_output(void* mg, double value_bits_as_double)
{
    Multi* m = (Multi*)mg;
    std::vector<double> row;
    for (Slab* s : m->slabs)
    {
        cudaSetDevice(s->device);
        row.assign((size_t)s->rows * s->nx, value_bits_as_double);
        cudaMemcpy(s->field[s->cur ^ 1], row.data(), row.size() * sizeof(double), cudaMemcpyHostToDevice);
    }
    ck("mg: fill");
}

void custen_mg_compute(void* mg)
{
    Multi* m = (Multi*)mg;
    for (Slab* s : m->slabs) slab_compute(s);
    ck("mg: compute");
}
void custen_mg_swap(void* mg)
{
    Multi* m = (Multi*)mg;
    for (Slab* s : m->slabs) slab_swap(s);
}

// nsteps x (Compute + Swap) on every slab.  With one slab per GPU each slab replays its own two-step graph; slabs that
// share a GPU go step by step in rank order.
void custen_mg_run(void* mg, int nsteps)
{
    Multi* m = (Multi*)mg;
    if (m->distinct)
    {
        int it = 0;
        if (m->slabs[0]->sweeps == 0 && nsteps > 0)
        {
            for (Slab* s : m->slabs) slab_run(s, 1, false);
            it = 1;
        }
        for (; it + 1 < nsteps; it += 2)
            for (Slab* s : m->slabs) slab_run(s, 2, true);
        if (it < nsteps)
            for (Slab* s : m->slabs) slab_run(s, 1, false);
        return;
    }
    for (int it = 0; it < nsteps; ++it)
        for (Slab* s : m->slabs) slab_run(s, 1, false);
}

void custen_mg_synchronize(void* mg)
{
    Multi* m = (Multi*)mg;
    for (Slab* s : m->slabs)
    {
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->stream);
    }
    ck("mg: synchronize");
}

int custen_mg_error(void* mg)
{
    Multi* m = (Multi*)mg;
    int e = 0;
    for (Slab* s : m->slabs) e |= custen_slab_error(s);
    return e;
}

void* custen_mg_slab(void* mg, int i) { return ((Multi*)mg)->slabs[i]; }

void custen_mg_destroy(void* mg)
{
    Multi* m = (Multi*)mg;
    for (Slab* s : m->slabs)
    {
        cudaSetDevice(s->device);
        cudaDeviceSynchronize();
    }
    for (Slab* s : m->slabs) slab_free(s);
    for (cudaStream_t st : m->shared) cudaStreamDestroy(st);
    delete m;
}

}  // extern "C"
