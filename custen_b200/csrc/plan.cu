// Plan builder + tile scheduler (host side).  See plan.h for the reference locations re-created here.
#include "plan.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

namespace custen {

static const uintptr_t kMagic = 0xC0573B200C0573ull;

static void check(const char* what, int device)
{
    char msg[256];
    snprintf(msg, sizeof msg, "%s on GPU %d", what, device);
    checkError(msg);
}

Plan* plan_of(cuSten_t* h)
{
    if (!h || !h->streams) return nullptr;
    if (h->numStreams < 3 || h->numStreams > 16) return nullptr;
    uintptr_t* tail = reinterpret_cast<uintptr_t*>(h->streams + h->numStreams);
    if (tail[0] != kMagic) return nullptr;
    return reinterpret_cast<Plan*>(tail[1]);
}

MemKind classify(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
    {
        cudaGetLastError();
        return MK_HOST;
    }
    if (at.type == cudaMemoryTypeManaged) return MK_MANAGED;
    if (at.type == cudaMemoryTypeDevice) return MK_DEVICE;
    return MK_HOST;  // registered / pinned or plain pageable host memory
}

// Tile seam pointers.  Same values the reference computes (custenCreateDestroy2DXYp.cu:194-228):
// the T rows above tile t and the B rows below it, taken from the same array, wrapped at the ends.
static void set_boundaries(cuSten_t* h, double* base)
{
    const int n = h->numTiles;
    const ptrdiff_t nx = h->nx;
    for (int t = 0; t < n; ++t)
    {
        h->boundaryTop[t] = (t == 0) ? base + (ptrdiff_t)(h->ny - h->numStenTop) * nx
                                     : base + ((ptrdiff_t)h->nyTile * t - h->numStenTop) * nx;
        h->boundaryBottom[t] = (t == n - 1) ? base : base + (ptrdiff_t)h->nyTile * (t + 1) * nx;
    }
}

// `dry` builds the plan without touching CUDA (no device, streams or events): used by custen_debug_bands so that
// the tiling / seam / mask logic can be tested on a machine without a GPU.
static void plan_create_impl(cuSten_t* h, Spec spec, int nstreams, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X,
                             int BLOCK_Y, double* dataOutput, double* dataInput, double* coef, int H, int L, int R, int V,
                             int T, int B, int numCoe, double* func, bool dry = false)
{
    memset(h, 0, sizeof *h);
    h->deviceNum = deviceNum;
    h->numStreams = nstreams;
    h->numTiles = numTiles < 1 ? 1 : numTiles;
    h->nx = nx;
    h->ny = ny;
    h->BLOCK_X = BLOCK_X;
    h->BLOCK_Y = BLOCK_Y;

    if (!dry)
    {
        cudaSetDevice(deviceNum);
        check("Setting current device", deviceNum);
    }

    // the public streams (blocking, like the reference's cudaStreamCreate; 3, or 6 for WENO) + hidden tail
    h->streams = (cudaStream_t*)calloc(nstreams + 2, sizeof(cudaStream_t));
    for (int s = 0; s < nstreams && !dry; ++s)
    {
        cudaStreamCreate(&h->streams[s]);
        check("Creating stream", deviceNum);
    }
    h->events = (cudaEvent_t*)calloc(2, sizeof(cudaEvent_t));
    for (int e = 0; e < 2 && !dry; ++e)
    {
        cudaEventCreateWithFlags(&h->events[e], cudaEventDisableTiming);
        check("Creating event", deviceNum);
    }

    Plan* p = (Plan*)calloc(1, sizeof(Plan));
    p->spec = spec;
    p->managed_policy = -1;
    uintptr_t* tail = reinterpret_cast<uintptr_t*>(h->streams + nstreams);
    tail[0] = kMagic;
    tail[1] = reinterpret_cast<uintptr_t>(p);

    // public geometry fields, as the reference fills them
    h->numStenLeft = L;
    h->numStenRight = R;
    h->numStenTop = T;
    h->numStenBottom = B;
    h->numStenHoriz = H;
    h->numStenVert = V;
    h->numSten = H * V;
    h->nxLocal = BLOCK_X + L + R;
    h->nyLocal = BLOCK_Y + T + B;
    if (spec.fun)
    {
        h->coe = coef;
        h->numCoe = numCoe;
        h->devFunc = func;
        p->ncoef = numCoe;
    }
    else
    {
        h->weights = coef;
        p->ncoef = H * V;
    }
    h->mem_shared = (int)(((size_t)h->nxLocal * h->nyLocal + p->ncoef) * sizeof(double));
    h->nyTile = ny / h->numTiles;
    h->xGrid = BLOCK_X > 0 ? (nx + BLOCK_X - 1) / BLOCK_X : 0;
    h->yGrid = BLOCK_Y > 0 ? (h->nyTile + BLOCK_Y - 1) / BLOCK_Y : 0;

    h->dataInput = (double**)calloc(h->numTiles, sizeof(double*));
    h->dataOutput = (double**)calloc(h->numTiles, sizeof(double*));
    const ptrdiff_t off = (ptrdiff_t)nx * h->nyTile;
    for (int t = 0; t < h->numTiles; ++t)
    {
        h->dataInput[t] = dataInput + t * off;
        h->dataOutput[t] = dataOutput + t * off;
    }
    if (spec.dir != DIR_X)
    {
        h->boundaryTop = (double**)calloc(h->numTiles, sizeof(double*));
        h->boundaryBottom = (double**)calloc(h->numTiles, sizeof(double*));
        set_boundaries(h, dataInput);
        h->numBoundaryTop = T * nx;
        h->numBoundaryBottom = B * nx;
    }
}

void plan_create(cuSten_t* h, Spec spec, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                 double* dataOutput, double* dataInput, double* coef, int H, int L, int R, int V, int T, int B,
                 int numCoe, double* func)
{
    plan_create_impl(h, spec, 3, deviceNum, numTiles, nx, ny, BLOCK_X, BLOCK_Y, dataOutput, dataInput, coef, H, L, R, V, T, B,
                     numCoe, func);
}

// 13th variant: periodic WENO5 advection, fixed 7 x 7 cross (custenCreateDestroy2DXYADVWENOp.cu:57-264): six public
// streams, 1/dx and 1/dy in coeDx / coeDy, per-tile aliases into the two velocity arrays.
void plan_create_weno(cuSten_t* h, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y, double dx,
                      double dy, double* u, double* v, double* dataOutput, double* dataInput)
{
    plan_create_impl(h, Spec{DIR_XY, 1, 0, 1}, 6, deviceNum, numTiles, nx, ny, BLOCK_X, BLOCK_Y, dataOutput, dataInput,
                     nullptr, 7, 3, 3, 7, 3, 3, 0, nullptr);
    h->coeDx = 1.0 / dx;
    h->coeDy = 1.0 / dy;
    h->uVel = (double**)calloc(h->numTiles, sizeof(double*));
    h->vVel = (double**)calloc(h->numTiles, sizeof(double*));
    const ptrdiff_t off = (ptrdiff_t)nx * h->nyTile;
    for (int t = 0; t < h->numTiles; ++t)
    {
        h->uVel[t] = u + t * off;
        h->vVel[t] = v + t * off;
    }
    Plan* p = plan_of(h);
    p->ncoef = 0;
    h->mem_shared = (int)(((size_t)h->nxLocal * h->nyLocal + h->numSten) * sizeof(double));
}

void plan_swap(cuSten_t* h, double* dataInput)
{
    for (int t = 0; t < h->numTiles; ++t) std::swap(h->dataInput[t], h->dataOutput[t]);
    // Y / XY variants re-derive the seams from the array that becomes the next input
    // (custenCreateDestroy2DXYp.cu:253-310); X variants ignore the argument (…2DXp.cu:179-190).
    if (h->boundaryTop && dataInput) set_boundaries(h, dataInput);
}

static void unadvise(Plan* p, int dev);
static void join_all(cuSten_t* h, Plan* p);

static void release_staging(Plan* p)
{
    for (int s = 0; s < kSlots; ++s)
    {
        if (p->d_in[s]) cudaFree(p->d_in[s]);
        if (p->d_out[s]) cudaFree(p->d_out[s]);
        if (p->d_aux[s]) cudaFree(p->d_aux[s]);
        p->d_in[s] = p->d_out[s] = p->d_aux[s] = nullptr;
        if (p->events_ready)
        {
            cudaEventDestroy(p->ev_loaded[s]);
            cudaEventDestroy(p->ev_done[s]);
            cudaEventDestroy(p->ev_unloaded[s]);
        }
    }
    if (p->d_coef) cudaFree(p->d_coef);
    p->d_coef = nullptr;
    p->coef_cap = 0;
    p->events_ready = 0;
    p->stage_rows = 0;
}

void plan_destroy(cuSten_t* h)
{
    cudaSetDevice(h->deviceNum);
    check("Setting current device", h->deviceNum);
    Plan* p = plan_of(h);
    if (p)
    {
        // staging buffers may still be in flight: drain this handle's streams before freeing them
        if (p->stage_rows || p->d_coef || p->spread)
            for (int s = 0; s < 3; ++s) cudaStreamSynchronize(h->streams[s]);
        release_staging(p);
        if (p->join_ready)
            for (int k = 0; k < 3; ++k) cudaEventDestroy(p->ev_join[k]);
        if (p->zc_n) unadvise(p, h->deviceNum);
        free(p);
    }
    for (int s = 0; s < h->numStreams; ++s)
    {
        cudaStreamDestroy(h->streams[s]);
        check("Destroying stream", h->deviceNum);
    }
    free(h->streams);
    for (int e = 0; e < 2; ++e)
    {
        cudaEventDestroy(h->events[e]);
        check("Destroying event", h->deviceNum);
    }
    free(h->events);
    free(h->dataInput);
    free(h->dataOutput);
    free(h->boundaryTop);
    free(h->boundaryBottom);
    free(h->uVel);
    free(h->vVel);
    h->uVel = h->vVel = nullptr;
    h->streams = nullptr;
    h->events = nullptr;
    h->dataInput = h->dataOutput = h->boundaryTop = h->boundaryBottom = nullptr;
}

// ------------------------------------------------------------------------------------------------
// band construction
// ------------------------------------------------------------------------------------------------

static Band base_band(const cuSten_t* h, const Plan* p, const double* coef)
{
    Band b{};
    const Spec& s = p->spec;
    b.nx = h->nx;
    b.dir = s.dir;
    b.L = s.dir == DIR_Y ? 0 : h->numStenLeft;
    b.R = s.dir == DIR_Y ? 0 : h->numStenRight;
    b.T = s.dir == DIR_X ? 0 : h->numStenTop;
    b.B = s.dir == DIR_X ? 0 : h->numStenBottom;
    b.H = s.dir == DIR_Y ? 1 : h->numStenHoriz;
    b.V = s.dir == DIR_X ? 1 : h->numStenVert;
    b.coef = coef;
    b.ncoef = p->ncoef;
    b.func = s.fun ? (const void*)h->devFunc : nullptr;
    b.wrap_x = s.periodic && s.dir != DIR_Y;
    b.xlo = 0;
    b.xhi = h->nx;
    if (!s.periodic && s.dir != DIR_Y)
    {
        b.xlo = b.L;
        b.xhi = h->nx - b.R;
        b.zero_right = (s.dir == DIR_X && !s.fun);  // 2d_x_np_kernel.cu:164-176
    }
    return b;
}

// rows [first_tile, last_tile] as one band
static Band make_band(const cuSten_t* h, const Plan* p, const double* coef, int first_tile, int last_tile)
{
    Band b = base_band(h, p, coef);
    const Spec& s = p->spec;
    const int n = h->numTiles;
    b.in = h->dataInput[first_tile];
    b.out = h->dataOutput[first_tile];
    if (s.weno)
    {
        b.weno = 1;
        b.aux0 = h->uVel[first_tile];
        b.aux1 = h->vVel[first_tile];
        b.p0 = h->coeDx;
        b.p1 = h->coeDy;
    }
    b.rows = h->nyTile * (last_tile - first_tile + 1);
    b.ylo = 0;
    b.yhi = b.rows;
    if (s.dir != DIR_X)
    {
        const bool at_top = first_tile == 0, at_bottom = last_tile == n - 1;
        b.top = h->boundaryTop[first_tile];
        b.bottom = h->boundaryBottom[last_tile];
        bool phys_top = at_top, phys_bottom = at_bottom;  // band touches the physical edge of the global grid
        if (p->slab_enabled)
        {
            if (at_top) b.top = p->slab_top;
            if (at_bottom) b.bottom = p->slab_bottom;
            phys_top = at_top && p->slab_first;
            phys_bottom = at_bottom && p->slab_last;
        }
        b.have_top = s.periodic || !phys_top;
        b.have_bottom = s.periodic || !phys_bottom;
        if (!s.periodic)
        {
            if (phys_top) b.ylo = b.T;
            if (phys_bottom) b.yhi = b.rows - b.B;
        }
        if (b.T == 0) b.have_top = 0;
        if (b.B == 0) b.have_bottom = 0;
        if (p->slab_enabled && p->sync_local && at_top && at_bottom)
        {
            // the slab above reads my first B rows as its bottom halo, the slab below my last T rows as its top halo
            b.sync_local = p->sync_local;
            b.wait_up = p->sync_wait_up;
            b.wait_down = p->sync_wait_down;
            b.signal_up = p->sync_signal_up;
            b.signal_down = p->sync_signal_down;
            b.guard_top = b.B;
            b.guard_bottom = b.T;
        }
    }
    return b;
}

static bool tiles_contiguous(const cuSten_t* h)
{
    const ptrdiff_t off = (ptrdiff_t)h->nx * h->nyTile;
    for (int t = 1; t < h->numTiles; ++t)
    {
        if (h->dataInput[t] != h->dataInput[0] + t * off) return false;
        if (h->dataOutput[t] != h->dataOutput[0] + t * off) return false;
        if (h->boundaryTop)
        {
            if (h->boundaryTop[t] != h->dataInput[t] - (ptrdiff_t)h->numStenTop * h->nx) return false;
            if (h->boundaryBottom[t - 1] != h->dataInput[t]) return false;
        }
        if (h->uVel && (h->uVel[t] != h->uVel[0] + t * off || h->vVel[t] != h->vVel[0] + t * off)) return false;
    }
    return true;
}

static void rotate(cuSten_t* h)
{
    cudaStream_t s0 = h->streams[0];
    h->streams[0] = h->streams[1];
    h->streams[1] = h->streams[2];
    h->streams[2] = s0;
    std::swap(h->events[0], h->events[1]);
}

// ------------------------------------------------------------------------------------------------
// Compute: three residency modes
// ------------------------------------------------------------------------------------------------

// device memory (or anything the GPU can address that needs no migration)
static void compute_resident(cuSten_t* h, Plan* p, const double* coef)
{
    if (tiles_contiguous(h))
    {
        const Band b = make_band(h, p, coef, 0, h->numTiles - 1);
        p->last_path = launch_band(b, h->streams[0]);
        p->last_mode = 0;
        check("Error computing grid", h->deviceNum);
        return;
    }
    p->last_mode = 1;
    join_all(h, p);
    p->spread = 1;
    for (int t = 0; t < h->numTiles; ++t)
    {
        const Band b = make_band(h, p, coef, t, t);
        p->last_path = launch_band(b, h->streams[0]);
        check("Error computing tile", h->deviceNum);
        rotate(h);
    }
}

static void prefetch(const void* ptr, size_t bytes, int dst, cudaStream_t st)
{
    if (ptr && bytes) cudaMemPrefetchAsync(ptr, bytes, dst, st);
}

static void prefetch_tile(cuSten_t* h, int t, int dst, cudaStream_t st)
{
    const size_t tile_bytes = (size_t)h->nx * h->nyTile * sizeof(double);
    prefetch(h->dataInput[t], tile_bytes, dst, st);
    prefetch(h->dataOutput[t], tile_bytes, dst, st);
    if (h->uVel)
    {
        prefetch(h->uVel[t], tile_bytes, dst, st);
        prefetch(h->vVel[t], tile_bytes, dst, st);
    }
    if (h->boundaryTop)
    {
        prefetch(h->boundaryTop[t], (size_t)h->numBoundaryTop * sizeof(double), dst, st);
        prefetch(h->boundaryBottom[t], (size_t)h->numBoundaryBottom * sizeof(double), dst, st);
    }
}

// ---- unified memory ------------------------------------------------------------------------------------------------
// Three roads, chosen per call (custen_set_managed_policy(1) pins the first one):
//   pipeline  the reference's load / compute / unload rotation over three streams and two events
//             (2d_xy_p_kernel.cu:561-655), ordered on the device with cudaStreamWaitEvent instead of host-side
//             cudaEventSynchronize / cudaStreamSynchronize;
//   resident  offload == DEVICE and every range was last prefetched to this GPU (what the previous DEVICE call
//             left behind): nothing has to move, so the prefetch calls - which cost more than the sweep itself on
//             multi-GiB ranges even when they move nothing - are skipped and the grid is swept like device memory.
//             Pages the CPU touched in between come back through ordinary GPU page faults;
//   zero-copy offload == HOST: the grid is meant to live on the CPU between sweeps.  Instead of migrating every
//             tile to the GPU and back (2 x 16 B per point over the host link at page-migration speed), the ranges
//             are advised "preferred location CPU, accessed by this GPU" and the kernel's own TMA producer reads
//             the rows over the host link while stores go straight back (8 + 8 B per point at link speed).  Pages
//             that happen to be on the GPU are read there and sent home by a prefetch behind the kernel.

struct Span
{
    const void* p;
    size_t bytes;
};

// the unified-memory arrays a sweep touches, as whole ranges (tiles are carved from one array each)
static int managed_spans(const cuSten_t* h, Span* sp)
{
    int n = 0;
    const size_t bytes = (size_t)h->nx * h->nyTile * h->numTiles * sizeof(double);
    const void* c[4] = {h->dataInput[0], h->dataOutput[0], h->uVel ? h->uVel[0] : nullptr, h->vVel ? h->vVel[0] : nullptr};
    for (int k = 0; k < 4; ++k)
        if (c[k] && classify(c[k]) == MK_MANAGED) sp[n++] = Span{c[k], bytes};
    return n;
}

// Where this handle's previous unified-memory call left the arrays (1 = on the GPU, 2 = at home on the CPU under the
// zero-copy advice).  Kept per handle because the driver offers no cheap residency query; a Swap exchanges in and out,
// so the comparison ignores order.
static bool left_at(const Plan* p, const Span* sp, int nsp, int where)
{
    if (p->res_where != where || p->res_n != nsp) return false;
    for (int k = 0; k < nsp; ++k)
    {
        bool found = false;
        for (int j = 0; j < p->res_n; ++j) found |= (p->res_ptr[j] == sp[k].p && p->res_bytes[j] == sp[k].bytes);
        if (!found) return false;
    }
    return true;
}
static void note_left_at(Plan* p, const Span* sp, int nsp, int where)
{
    p->res_where = where;
    p->res_n = nsp;
    for (int k = 0; k < nsp; ++k)
    {
        p->res_ptr[k] = sp[k].p;
        p->res_bytes[k] = sp[k].bytes;
    }
}

static void unadvise(Plan* p, int dev)
{
    for (int k = 0; k < p->zc_n; ++k)
    {
        cudaMemAdvise(p->zc_ptr[k], p->zc_bytes[k], cudaMemAdviseUnsetAccessedBy, dev);
        cudaMemAdvise(p->zc_ptr[k], p->zc_bytes[k], cudaMemAdviseUnsetPreferredLocation, dev);
    }
    cudaGetLastError();  // the caller may already have freed the arrays
    p->zc_n = 0;
}

static bool advised(const Plan* p, const Span& s)
{
    for (int k = 0; k < p->zc_n; ++k)
        if (p->zc_ptr[k] == s.p && p->zc_bytes[k] == s.bytes) return true;
    return false;
}

static int managed_policy_default = 0;   // what a handle without its own setting follows
void set_managed_policy(int policy) { managed_policy_default = policy; }
void set_handle_managed_policy(cuSten_t* h, int policy)
{
    if (Plan* p = plan_of(h)) p->managed_policy = policy;
}
static int managed_policy_of(const Plan* p) { return p->managed_policy >= 0 ? p->managed_policy : managed_policy_default; }

// Ordering between consecutive calls on one handle.  The single-launch roads use streams[0] only; the tile pipelines
// (per-tile resident launches, the unified-memory prefetch pipeline, the staged pipeline for host grids) spread a call
// over all three streams.  Whenever the previous call or the coming one is of the second kind, the call starts with a
// three-way join: every stream waits for whatever the previous call left running on the other two.  Without it the
// next call's uploads could overwrite a staging slot a kernel is still reading, or read a host array the previous
// call's downloads are still writing (Compute, Swap, Compute with no device sync in between).
static void join_all(cuSten_t* h, Plan* p)
{
    if (p->joined_now) return;  // this call has already joined (plan_compute does it first when the previous call was spread)
    p->joined_now = 1;
    if (!p->join_ready)
    {
        for (int k = 0; k < 3; ++k) cudaEventCreateWithFlags(&p->ev_join[k], cudaEventDisableTiming);
        p->join_ready = 1;
    }
    for (int k = 0; k < 3; ++k) cudaEventRecord(p->ev_join[k], h->streams[k]);
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j)
            if (j != k) cudaStreamWaitEvent(h->streams[k], p->ev_join[j], 0);
    p->spread = 0;
}

static void compute_managed(cuSten_t* h, Plan* p, const double* coef, MemKind kcoef, bool offload)
{
    const int dev = h->deviceNum;
    Span sp[kMaxSpans];
    const int nsp = (managed_policy_of(p) == 0 && tiles_contiguous(h)) ? managed_spans(h, sp) : 0;
    const bool spans_complete = nsp > 0;

    if (spans_complete && !offload)
    {
        if (left_at(p, sp, nsp, 1))
        {
            // the coefficients came over with the call that brought the grid; a prefetch that moves nothing still costs
            // a few hundred microseconds of stream time, so it is not repeated (pages the CPU rewrote fault back in)
            if (kcoef == MK_MANAGED && p->coef_on_gpu != coef) prefetch(coef, (size_t)p->ncoef * sizeof(double), dev, h->streams[0]);
            p->coef_on_gpu = coef;
            compute_resident(h, p, coef);
            p->last_mode = 4;
            return;
        }
    }
    if (spans_complete && offload)
    {
        int concurrent = 0;
        cudaDeviceGetAttribute(&concurrent, cudaDevAttrConcurrentManagedAccess, dev);
        if (concurrent)
        {
            bool ok = true;
            for (int k = 0; k < nsp && ok; ++k)
            {
                if (advised(p, sp[k])) continue;
                // a range the caller has already given a preferred location keeps it: the zero-copy road only takes
                // ranges nobody has advised (this handle's own advice is withdrawn again by unadvise())
                int pref = cudaInvalidDeviceId;
                if (cudaMemRangeGetAttribute(&pref, sizeof pref, cudaMemRangeAttributePreferredLocation, sp[k].p, sp[k].bytes) !=
                        cudaSuccess ||
                    pref != cudaInvalidDeviceId)
                {
                    cudaGetLastError();
                    ok = false;
                    break;
                }
                ok = cudaMemAdvise(sp[k].p, sp[k].bytes, cudaMemAdviseSetPreferredLocation, cudaCpuDeviceId) == cudaSuccess &&
                     cudaMemAdvise(sp[k].p, sp[k].bytes, cudaMemAdviseSetAccessedBy, dev) == cudaSuccess;
                if (ok && p->zc_n < kMaxSpans)
                {
                    p->zc_ptr[p->zc_n] = sp[k].p;
                    p->zc_bytes[p->zc_n] = sp[k].bytes;
                    ++p->zc_n;
                }
            }
            if (ok)
            {
                    if (kcoef == MK_MANAGED && p->coef_on_gpu != coef) prefetch(coef, (size_t)p->ncoef * sizeof(double), dev, h->streams[0]);
                p->coef_on_gpu = coef;
                cudaStream_t st = h->streams[0];
                compute_resident(h, p, coef);
                p->last_mode = 5;
                // anything that was on the GPU goes home behind the kernel (stream order); once there, the advice
                // keeps it there
                if (!left_at(p, sp, nsp, 2))
                {
                    for (int k = 0; k < nsp; ++k) prefetch(sp[k].p, sp[k].bytes, cudaCpuDeviceId, st);
                    note_left_at(p, sp, nsp, 2);
                }
                check("Error in unified-memory zero-copy sweep", dev);
                return;
            }
            cudaGetLastError();
            unadvise(p, dev);
        }
    }

    if (p->zc_n) unadvise(p, dev);
    join_all(h, p);
    p->spread = 1;
    p->last_mode = 2;
    p->res_where = 0;
    if (spans_complete && !offload) note_left_at(p, sp, nsp, 1);  // the pipeline below leaves every tile on the GPU
    if (kcoef == MK_MANAGED) prefetch(coef, (size_t)p->ncoef * sizeof(double), dev, h->streams[1]);
    p->coef_on_gpu = kcoef == MK_MANAGED ? coef : nullptr;
    prefetch_tile(h, 0, dev, h->streams[1]);
    cudaEventRecord(h->events[0], h->streams[1]);
    for (int t = 0; t < h->numTiles; ++t)
    {
        cudaStreamWaitEvent(h->streams[0], h->events[0], 0);
        const Band b = make_band(h, p, coef, t, t);
        p->last_path = launch_band(b, h->streams[0]);
        check("Error computing tile", dev);
        if (offload)
        {
            const size_t tile_bytes = (size_t)h->nx * h->nyTile * sizeof(double);
            prefetch(h->dataOutput[t], tile_bytes, cudaCpuDeviceId, h->streams[0]);
            prefetch(h->dataInput[t], tile_bytes, cudaCpuDeviceId, h->streams[0]);
        }
        if (t + 1 < h->numTiles)
        {
            prefetch_tile(h, t + 1, dev, h->streams[1]);
            cudaEventRecord(h->events[1], h->streams[1]);
        }
        rotate(h);
    }
    check("Error in managed tile pipeline", dev);
}

// host-resident grids (pinned or pageable): a ring of device slots; tile t+1 uploads on the load
// stream while tile t computes and tile t-1 downloads.  Only the written rectangle is copied back,
// so the untouched frame of the non-periodic variants keeps the caller's values.
static void compute_staged(cuSten_t* h, Plan* p, const double* coef)
{
    const int dev = h->deviceNum;
    const Spec& s = p->spec;
    const int T = s.dir == DIR_X ? 0 : h->numStenTop;
    const int B = s.dir == DIR_X ? 0 : h->numStenBottom;
    const size_t nx = h->nx;
    const size_t need_rows = (size_t)h->nyTile + T + B;
    p->last_mode = 3;

    if (p->stage_rows < need_rows)
    {
        for (int k = 0; k < 3; ++k) cudaStreamSynchronize(h->streams[k]);
        const int ready = p->events_ready;
        double* keep_coef = p->d_coef;
        p->d_coef = nullptr;
        p->events_ready = 0;
        release_staging(p);
        p->d_coef = keep_coef;
        p->events_ready = ready;
        const int slots = h->numTiles < kSlots ? h->numTiles : kSlots;
        for (int k = 0; k < slots; ++k)
        {
            cudaMalloc(&p->d_in[k], need_rows * nx * sizeof(double));
            cudaMalloc(&p->d_out[k], (size_t)h->nyTile * nx * sizeof(double));
            if (s.weno) cudaMalloc(&p->d_aux[k], (size_t)2 * h->nyTile * nx * sizeof(double));
            check("Allocating staging slot", dev);
        }
        p->stage_rows = need_rows;
    }
    if (!p->events_ready)
    {
        for (int k = 0; k < kSlots; ++k)
        {
            cudaEventCreateWithFlags(&p->ev_loaded[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&p->ev_done[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&p->ev_unloaded[k], cudaEventDisableTiming);
        }
        p->events_ready = 1;
    }

    // the previous call on this handle may still be computing on / downloading from the slots, or (after a Swap)
    // writing the host array this call uploads from
    join_all(h, p);
    p->spread = 1;
    cudaStream_t s_comp = h->streams[0], s_load = h->streams[1], s_unload = h->streams[2];
    const int slots = h->numTiles < kSlots ? h->numTiles : kSlots;
    const size_t row_bytes = nx * sizeof(double);
    for (int t = 0; t < h->numTiles; ++t)
    {
        const int k = t % slots;
        Band b = make_band(h, p, coef, t, t);
        // slot k is free once its previous occupant has been downloaded
        if (t >= slots) cudaStreamWaitEvent(s_load, p->ev_unloaded[k], 0);
        double* din = p->d_in[k];
        if (b.have_top) cudaMemcpyAsync(din, b.top, (size_t)T * row_bytes, cudaMemcpyDefault, s_load);
        cudaMemcpyAsync(din + (size_t)T * nx, b.in, (size_t)h->nyTile * row_bytes, cudaMemcpyDefault, s_load);
        if (b.have_bottom)
            cudaMemcpyAsync(din + ((size_t)T + h->nyTile) * nx, b.bottom, (size_t)B * row_bytes, cudaMemcpyDefault,
                            s_load);
        if (s.weno)
        {
            // the velocities of the tile's own rows (the kernel reads them per output point: no halo)
            double* daux = p->d_aux[k];
            cudaMemcpyAsync(daux, b.aux0, (size_t)h->nyTile * row_bytes, cudaMemcpyDefault, s_load);
            cudaMemcpyAsync(daux + (size_t)h->nyTile * nx, b.aux1, (size_t)h->nyTile * row_bytes, cudaMemcpyDefault, s_load);
            b.aux0 = daux;
            b.aux1 = daux + (size_t)h->nyTile * nx;
        }
        cudaEventRecord(p->ev_loaded[k], s_load);

        double* host_out = b.out;
        b.top = din;
        b.in = din + (size_t)T * nx;
        b.bottom = din + ((size_t)T + h->nyTile) * nx;
        b.out = p->d_out[k];
        cudaStreamWaitEvent(s_comp, p->ev_loaded[k], 0);
        p->last_path = launch_band(b, s_comp);
        check("Error computing staged tile", dev);
        cudaEventRecord(p->ev_done[k], s_comp);

        cudaStreamWaitEvent(s_unload, p->ev_done[k], 0);
        const int x0 = b.xlo, x1 = b.zero_right ? b.nx : b.xhi;
        if (b.yhi > b.ylo && x1 > x0)
        {
            const size_t o = (size_t)b.ylo * nx + x0;
            if (x0 == 0 && x1 == b.nx)
                cudaMemcpyAsync(host_out + o, b.out + o, (size_t)(b.yhi - b.ylo) * row_bytes, cudaMemcpyDefault,
                                s_unload);
            else
                cudaMemcpy2DAsync(host_out + o, row_bytes, b.out + o, row_bytes, (size_t)(x1 - x0) * sizeof(double),
                                  (size_t)(b.yhi - b.ylo), cudaMemcpyDefault, s_unload);
        }
        cudaEventRecord(p->ev_unloaded[k], s_unload);
    }
    // later work on the compute stream (and the caller's device-wide sync) sees the downloads
    cudaStreamWaitEvent(s_comp, p->ev_unloaded[(h->numTiles - 1) % slots], 0);
    check("Error in staged tile pipeline", dev);
}

int plan_launch_slab(cuSten_t* h, cudaStream_t stream)
{
    Plan* p = plan_of(h);
    const Spec& s = p->spec;
    const double* coef = s.fun ? h->coe : h->weights;
    const Band b = make_band(h, p, coef, 0, h->numTiles - 1);
    p->last_path = launch_band(b, stream);
    p->last_mode = 0;
    check("Error computing slab", h->deviceNum);
    return p->last_path;
}

void plan_compute(cuSten_t* h, bool offload)
{
    cudaSetDevice(h->deviceNum);
    check("Setting current device", h->deviceNum);
    Plan* p = plan_of(h);
    if (!p)
    {
        printf("\ncuSten: handle was not created by cuStenCreate2D*\nprogram terminated ...\n\n");
        exit(EXIT_FAILURE);
    }
    const Spec& s = p->spec;
    const double* coef = s.fun ? h->coe : h->weights;
    const MemKind kin = classify(h->dataInput[0]);
    const MemKind kout = classify(h->dataOutput[0]);
    MemKind kcoef = coef ? classify(coef) : MK_DEVICE;
    // WENO: velocities in plain host memory send the call down the staged road like a host-resident field does
    const bool weno_host_vel = s.weno && (classify(h->uVel[0]) == MK_HOST || classify(h->vVel[0]) == MK_HOST);

    // the previous call left work on all three streams: this one starts behind all of it
    if (p->spread) join_all(h, p);

    // coefficients living in plain host memory are snapshotted to the device, stream-ordered
    if (kcoef == MK_HOST)
    {
        if (p->coef_cap < (size_t)p->ncoef)
        {
            if (p->d_coef)
            {
                for (int k = 0; k < 3; ++k) cudaStreamSynchronize(h->streams[k]);
                cudaFree(p->d_coef);
            }
            p->coef_cap = p->ncoef > 64 ? (size_t)p->ncoef : 64;
            cudaMalloc(&p->d_coef, p->coef_cap * sizeof(double));
            check("Allocating coefficient buffer", h->deviceNum);
        }
        cudaMemcpyAsync(p->d_coef, coef, (size_t)p->ncoef * sizeof(double), cudaMemcpyHostToDevice, h->streams[0]);
        // every stream that may launch a kernel must see the snapshot
        cudaEventRecord(h->events[1], h->streams[0]);
        cudaStreamWaitEvent(h->streams[1], h->events[1], 0);
        cudaStreamWaitEvent(h->streams[2], h->events[1], 0);
        coef = p->d_coef;
        kcoef = MK_DEVICE;
    }

    if (kin == MK_HOST || kout == MK_HOST || weno_host_vel) compute_staged(h, p, coef);
    else if (kin == MK_MANAGED || kout == MK_MANAGED) compute_managed(h, p, coef, kcoef, offload);
    else compute_resident(h, p, coef);
    p->joined_now = 0;
}

// ---- host-logic probe (no CUDA) --------------------------------------------------------------------------------
// Builds the plan of a variant on fake base addresses and reports the bands Compute would launch: per tile
// (merged == 0) or as one band over all tiles (merged == 1).  Offsets are in doubles from the input / output base.
int debug_bands(int variant, int numTiles, int nx, int ny, int H, int L, int R, int V, int T, int B, int merged,
                int slab, int slab_first, int slab_last, BandDesc* out, int max_out)
{
    static const Spec specs[12] = {{DIR_X, 1, 0, 0}, {DIR_X, 0, 0, 0}, {DIR_X, 1, 1, 0}, {DIR_X, 0, 1, 0},
                                   {DIR_Y, 1, 0, 0}, {DIR_Y, 0, 0, 0}, {DIR_Y, 1, 1, 0}, {DIR_Y, 0, 1, 0},
                                   {DIR_XY, 1, 0, 0}, {DIR_XY, 0, 0, 0}, {DIR_XY, 1, 1, 0}, {DIR_XY, 0, 1, 0}};
    if (variant < 0 || variant >= 12) return -1;
    const Spec sp = specs[variant];
    double* const in = reinterpret_cast<double*>((uintptr_t)1 << 40);
    double* const outp = reinterpret_cast<double*>((uintptr_t)1 << 41);
    double* const top_halo = reinterpret_cast<double*>((uintptr_t)1 << 42);
    double* const bot_halo = reinterpret_cast<double*>((uintptr_t)1 << 43);
    cuSten_t h;
    const int HH = sp.dir == DIR_Y ? 1 : H, VV = sp.dir == DIR_X ? 1 : V;
    plan_create_impl(&h, sp, 3, 0, numTiles, nx, ny, 32, 8, outp, in, nullptr, HH, sp.dir == DIR_Y ? 0 : L,
                     sp.dir == DIR_Y ? 0 : R, VV, sp.dir == DIR_X ? 0 : T, sp.dir == DIR_X ? 0 : B, HH * VV, nullptr, true);
    Plan* p = plan_of(&h);
    if (slab)
    {
        p->slab_enabled = 1;
        p->slab_top = top_halo;
        p->slab_bottom = bot_halo;
        p->slab_first = slab_first;
        p->slab_last = slab_last;
    }
    int n = 0;
    auto emit = [&](const Band& b) {
        if (n >= max_out) return;
        BandDesc& d = out[n++];
        d.in_off = b.in - in;
        d.out_off = b.out - outp;
        d.top_kind = !b.have_top ? 0 : (b.top == top_halo ? 2 : 1);
        d.top_off = d.top_kind == 1 ? b.top - in : 0;
        d.bottom_kind = !b.have_bottom ? 0 : (b.bottom == bot_halo ? 2 : 1);
        d.bottom_off = d.bottom_kind == 1 ? b.bottom - in : 0;
        d.rows = b.rows; d.nx = b.nx; d.L = b.L; d.R = b.R; d.T = b.T; d.B = b.B; d.H = b.H; d.V = b.V;
        d.wrap_x = b.wrap_x; d.xlo = b.xlo; d.xhi = b.xhi; d.ylo = b.ylo; d.yhi = b.yhi; d.zero_right = b.zero_right;
        d.contiguous = tiles_contiguous(&h) ? 1 : 0;
    };
    if (merged) emit(make_band(&h, p, nullptr, 0, h.numTiles - 1));
    else
        for (int t = 0; t < h.numTiles; ++t) emit(make_band(&h, p, nullptr, t, t));
    free(p);
    free(h.streams);
    free(h.events);
    free(h.dataInput);
    free(h.dataOutput);
    free(h.boundaryTop);
    free(h.boundaryBottom);
    return n;
}

}  // namespace custen
