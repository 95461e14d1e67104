// extern "C" boundary of libcusten_b200.so (declared in include/custen_c.h).
#include "../../include/custen_c.h"
#include "../../include/cuSten_fun.h"
#include "builtin_funs.cuh"
#include "plan.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

using namespace custen;

static inline cuSten_t* H(cuSten_c_handle* h) { return reinterpret_cast<cuSten_t*>(h); }

// ---- user-function fixtures: device-side pointer variables, read back like a C++ caller would -----------------
__device__ cuStenFunX fp_second_diff_x = custen_funs::second_diff_x;
__device__ cuStenFunX fp_weighted9_x = custen_funs::weighted9_x;
__device__ cuStenFunY fp_weighted9_y = custen_funs::weighted9_y;
__device__ cuStenFunY fp_weighted3_y = custen_funs::weighted3_y;
__device__ cuStenFunXY fp_weighted_xy = custen_funs::weighted_xy;
__device__ cuStenFunXY fp_cubic_xy = custen_funs::cubic_xy;

// ... and registered, so that the library's own fixtures take the inlined road (custen_set_tuning can force the
// opaque-pointer road for comparison).
CUSTEN_REGISTER_FUN_X_AS(custen_funs::second_diff_x, second_diff_x)
CUSTEN_REGISTER_FUN_X_AS(custen_funs::weighted9_x, weighted9_x)
CUSTEN_REGISTER_FUN_Y_AS(custen_funs::weighted9_y, weighted9_y)
CUSTEN_REGISTER_FUN_Y_AS(custen_funs::weighted3_y, weighted3_y)
CUSTEN_REGISTER_FUN_XY_AS(custen_funs::weighted_xy, weighted_xy)
CUSTEN_REGISTER_FUN_XY_AS(custen_funs::cubic_xy, cubic_xy)

extern "C" {

#define C_COMMON(V)                                                                                        \
    void custenSwap2D##V(cuSten_c_handle* h, double* dataInput) { cuStenSwap2D##V(H(h), dataInput); }       \
    void custenDestroy2D##V(cuSten_c_handle* h) { cuStenDestroy2D##V(H(h)); }                               \
    void custenCompute2D##V(cuSten_c_handle* h, int offload) { cuStenCompute2D##V(H(h), offload != 0); }

#define PFX cuSten_c_handle *h, int dev, int tiles, int nx, int ny, int bx, int by, double *out, double *in
#define PFA H(h), dev, tiles, nx, ny, bx, by, out, in

void custenCreate2DXp(PFX, double* w, int n, int l, int r) { cuStenCreate2DXp(PFA, w, n, l, r); }
C_COMMON(Xp)
void custenCreate2DXnp(PFX, double* w, int n, int l, int r) { cuStenCreate2DXnp(PFA, w, n, l, r); }
C_COMMON(Xnp)
void custenCreate2DXpFun(PFX, double* c, int n, int l, int r, int nc, double* f) { cuStenCreate2DXpFun(PFA, c, n, l, r, nc, f); }
C_COMMON(XpFun)
void custenCreate2DXnpFun(PFX, double* c, int n, int l, int r, int nc, double* f) { cuStenCreate2DXnpFun(PFA, c, n, l, r, nc, f); }
C_COMMON(XnpFun)

void custenCreate2DYp(PFX, double* w, int n, int t, int b) { cuStenCreate2DYp(PFA, w, n, t, b); }
C_COMMON(Yp)
void custenCreate2DYnp(PFX, double* w, int n, int t, int b) { cuStenCreate2DYnp(PFA, w, n, t, b); }
C_COMMON(Ynp)
void custenCreate2DYpFun(PFX, double* c, int n, int t, int b, int nc, double* f) { cuStenCreate2DYpFun(PFA, c, n, t, b, nc, f); }
C_COMMON(YpFun)
void custenCreate2DYnpFun(PFX, double* c, int n, int t, int b, double* f) { cuStenCreate2DYnpFun(PFA, c, n, t, b, f); }
C_COMMON(YnpFun)

void custenCreate2DXYp(PFX, double* w, int hh, int l, int r, int vv, int t, int b) { cuStenCreate2DXYp(PFA, w, hh, l, r, vv, t, b); }
C_COMMON(XYp)
void custenCreate2DXYnp(PFX, double* w, int hh, int l, int r, int vv, int t, int b) { cuStenCreate2DXYnp(PFA, w, hh, l, r, vv, t, b); }
C_COMMON(XYnp)
void custenCreate2DXYpFun(PFX, double* c, int hh, int l, int r, int vv, int t, int b, double* f)
{
    cuStenCreate2DXYpFun(PFA, c, hh, l, r, vv, t, b, f);
}
C_COMMON(XYpFun)
void custenCreate2DXYnpFun(PFX, double* c, int hh, int l, int r, int vv, int t, int b, double* f)
{
    cuStenCreate2DXYnpFun(PFA, c, hh, l, r, vv, t, b, f);
}
C_COMMON(XYnpFun)

void custenCreate2DXYWENOADVp(cuSten_c_handle* h, int dev, int tiles, int nx, int ny, int bx, int by, double dx, double dy,
                              double* u, double* v, double* out, double* in)
{
    cuStenCreate2DXYWENOADVp(H(h), dev, tiles, nx, ny, bx, by, dx, dy, u, v, out, in);
}
C_COMMON(XYWENOADVp)

void custenCheckError(const char* action) { checkError(action); }

// ---- additive -------------------------------------------------------------------------------------------------

size_t custen_handle_size(void) { return sizeof(cuSten_t); }

void custen_device_synchronize(void)
{
    cudaDeviceSynchronize();
    checkError("cudaDeviceSynchronize");
}

double* custen_builtin_fun(const char* name)
{
    void* fp = nullptr;
#define LOOKUP(N)                                                     \
    if (!strcmp(name, #N))                                            \
    {                                                                 \
        cudaMemcpyFromSymbol(&fp, fp_##N, sizeof(void*));             \
        checkError("reading device function pointer " #N);            \
        return (double*)fp;                                           \
    }
    LOOKUP(second_diff_x)
    LOOKUP(weighted9_x)
    LOOKUP(weighted9_y)
    LOOKUP(weighted3_y)
    LOOKUP(weighted_xy)
    LOOKUP(cubic_xy)
#undef LOOKUP
    return nullptr;
}

int custen_last_path(cuSten_c_handle* h)
{
    Plan* p = plan_of(H(h));
    return p ? p->last_path : -1;
}
int custen_last_mode(cuSten_c_handle* h)
{
    Plan* p = plan_of(H(h));
    return p ? p->last_mode : -1;
}
uint64_t custen_launch_count(void) { return launches_total(); }

void custen_set_tuning(int force_fallback, int force_tile, int chunk_rows, int ctas_per_sm, int force_opaque)
{
    Tuning& t = tuning();
    t.force_fallback = force_fallback;
    t.force_tile = force_tile;
    t.chunk_rows = chunk_rows;
    t.ctas_per_sm = ctas_per_sm;
    t.force_opaque = force_opaque;
}

void custen_set_managed_policy(int policy) { set_managed_policy(policy); }
void custen_set_handle_managed_policy(cuSten_c_handle* h, int policy) { set_handle_managed_policy(H(h), policy); }

// unified-memory plumbing for callers without a CUDA runtime binding of their own (tools/um_probe.py, tests)
int custen_mem_advise(const void* p, size_t bytes, int advice, int device)
{
    const cudaError_t e = cudaMemAdvise(p, bytes, (cudaMemoryAdvise)advice, device);
    cudaGetLastError();
    return (int)e;
}
int custen_mem_prefetch(const void* p, size_t bytes, int device)
{
    const cudaError_t e = cudaMemPrefetchAsync(p, bytes, device, 0);
    cudaGetLastError();
    return (int)e;
}

void custen_set_slab(cuSten_c_handle* h, const double* top, const double* bottom, int is_first, int is_last)
{
    Plan* p = plan_of(H(h));
    if (!p) return;
    p->slab_enabled = 1;
    p->slab_top = top;
    p->slab_bottom = bottom;
    p->slab_first = is_first;
    p->slab_last = is_last;
}

void custen_ipc_export(const void* dev_ptr, void* handle64, size_t* offset_out)
{
    cudaIpcMemHandle_t hd;
    cudaIpcGetMemHandle(&hd, const_cast<void*>(dev_ptr));
    checkError("cudaIpcGetMemHandle");
    memcpy(handle64, &hd, sizeof hd);
    // base of the allocation, through the driver entry point (no link-time dependency on libcuda)
    size_t off = 0;
    typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn)
    {
        unsigned long long base = 0;
        size_t size = 0;
        if (((GetRange)fn)(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) == 0)
            off = (size_t)((uintptr_t)dev_ptr - (uintptr_t)base);
    }
    cudaGetLastError();
    if (offset_out) *offset_out = off;
}
void* custen_ipc_open(const void* handle64)
{
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, sizeof hd);
    void* p = nullptr;
    cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    checkError("cudaIpcOpenMemHandle");
    return p;
}
void custen_ipc_close(void* mapped_ptr)
{
    cudaIpcCloseMemHandle(mapped_ptr);
    checkError("cudaIpcCloseMemHandle");
}

// Neighbour barrier of the slab layer: one thread publishes this rank's epoch into the flag words of the slab
// above and below (peer memory over NVLink) and spins until both neighbours have published theirs.  Ordered on
// the handle's compute stream, so "my previous sweep is complete" is what the epoch announces.
__global__ void custen_peer_barrier_kernel(unsigned long long* up_slot, unsigned long long* down_slot,
                                           volatile unsigned long long* from_up, volatile unsigned long long* from_down,
                                           unsigned long long epoch)
{
    __threadfence_system();
    if (up_slot) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(up_slot), "l"(epoch) : "memory");
    if (down_slot) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(down_slot), "l"(epoch) : "memory");
    unsigned long long v;
    if (up_slot)
        do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(from_up) : "memory"); } while (v < epoch);
    if (down_slot)
        do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(from_down) : "memory"); } while (v < epoch);
    __threadfence_system();
}

void custen_peer_barrier(cuSten_c_handle* h, void* up_flags, void* down_flags, void* my_flags, uint64_t epoch)
{
    // flag layout per rank: word 0 is written by the slab above, word 1 by the slab below
    unsigned long long* mine = (unsigned long long*)my_flags;
    unsigned long long* up = (unsigned long long*)up_flags;
    unsigned long long* down = (unsigned long long*)down_flags;
    // no handle: the legacy default stream, which every blocking stream of every handle orders against
    cudaStream_t st = h ? H(h)->streams[0] : (cudaStream_t)0;
    custen_peer_barrier_kernel<<<1, 1, 0, st>>>(up ? up + 1 : nullptr, down ? down + 0 : nullptr, mine + 0, mine + 1,
                                                (unsigned long long)epoch);
    checkError("custen_peer_barrier");
}

// field[r][c] = lo + (hi - lo) * u(seed, (row0 + r) * nx + c): counter-based, so any row can be regenerated on the CPU
__global__ void custen_fill_hash_kernel(double* f, long long row0, long long count, int nx, unsigned long long seed, double lo,
                                        double hi)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < count; i += stride)
    {
        unsigned long long z = seed + (unsigned long long)(row0 * nx + i) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
        f[i] = __dadd_rn(lo, __dmul_rn(hi - lo, u));   // two roundings, like the numpy twin (no FMA contraction)
    }
}
void custen_fill_hash(double* dev_field, long long row0, int rows, int nx, unsigned long long seed, double lo, double hi)
{
    custen_fill_hash_kernel<<<148 * 8, 256>>>(dev_field, row0, (long long)rows * nx, nx, seed, lo, hi);
    checkError("custen_fill_hash");
}

void* custen_device_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaMalloc(&p, bytes);
    checkError("cudaMalloc");
    cudaMemset(p, 0, bytes);
    return p;
}
void custen_device_free(void* p) { cudaFree(p); }

int custen_debug_bands(int variant, int numTiles, int nx, int ny, int H, int L, int R, int V, int T, int B, int merged,
                       int slab, int slab_first, int slab_last, void* out, int max_out)
{
    return debug_bands(variant, numTiles, nx, ny, H, L, R, V, T, B, merged, slab, slab_first, slab_last,
                       reinterpret_cast<BandDesc*>(out), max_out);
}

void* custen_event_create(void)
{
    cudaEvent_t e;
    cudaEventCreate(&e);
    checkError("cudaEventCreate");
    return e;
}
void custen_event_record(void* ev, cuSten_c_handle* h, int stream_idx)
{
    cudaEventRecord((cudaEvent_t)ev, H(h)->streams[stream_idx]);
    checkError("cudaEventRecord");
}
void custen_event_synchronize(void* ev)
{
    cudaEventSynchronize((cudaEvent_t)ev);
    checkError("cudaEventSynchronize");
}
float custen_event_elapsed_ms(void* start, void* stop)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, (cudaEvent_t)start, (cudaEvent_t)stop);
    checkError("cudaEventElapsedTime");
    return ms;
}
void custen_event_destroy(void* ev) { cudaEventDestroy((cudaEvent_t)ev); }

void* custen_host_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    checkError("cudaHostAlloc");
    return p;
}
void custen_host_free(void* p) { cudaFreeHost(p); }

// ---- pinned host memory next to a GPU, and what the host link can carry ------------------------------------------------
// NUMA node the GPU hangs off (sysfs), or -1 when the platform does not say.
int custen_device_numa_node(int device)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess)
    {
        cudaGetLastError();
        return -1;
    }
    for (char* c = bus; *c; ++c)
        if (*c >= 'A' && *c <= 'Z') *c += 'a' - 'A';
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// Pinned host buffer whose pages are bound to the GPU's NUMA node (mmap + mbind + cudaHostRegister), so that the
// staged pipeline of several ranks does not funnel every transfer through one socket's memory.  *node_out: the node
// the pages were bound to, or -1 when binding was not possible (then this is an ordinary pinned allocation).
// Free with custen_host_free_near.
void* custen_host_alloc_near(size_t bytes, int device, int* node_out)
{
    const int node = custen_device_numa_node(device);
    if (node_out) *node_out = -1;
    const size_t len = (bytes + 4095) & ~(size_t)4095;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    if (node >= 0 && node < 64)
    {
        unsigned long mask = 1ul << node;
        const int MPOL_BIND_ = 2;
        if (syscall(SYS_mbind, p, len, MPOL_BIND_, &mask, sizeof(mask) * 8 + 1, 0) == 0 && node_out) *node_out = node;
    }
    memset(p, 0, len);   // first touch under the policy
    if (cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess)
    {
        cudaGetLastError();
        munmap(p, len);
        return nullptr;
    }
    return p;
}
void custen_host_free_near(void* p, size_t bytes)
{
    if (!p) return;
    cudaHostUnregister(p);
    cudaGetLastError();
    munmap(p, (bytes + 4095) & ~(size_t)4095);
}

// Plain cudaMemcpyAsync in both directions at once between pinned host buffers and device scratch, `iters` times:
// the ceiling the out-of-core pipeline is measured against (bench.py e2e.link_frac).  Returns milliseconds.
float custen_link_probe(const void* host_src, void* host_dst, size_t bytes, int iters, int device)
{
    cudaSetDevice(device);
    const size_t chunk = bytes < ((size_t)256 << 20) ? bytes : ((size_t)256 << 20);
    void *d_in = nullptr, *d_out = nullptr;
    cudaMalloc(&d_in, chunk);
    cudaMalloc(&d_out, chunk);
    cudaStream_t s_in, s_out;
    cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    checkError("custen_link_probe: set-up");
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s_in);
    cudaStreamWaitEvent(s_out, e0, 0);
    for (int it = 0; it < iters; ++it)
        for (size_t o = 0; o < bytes; o += chunk)
        {
            const size_t n = bytes - o < chunk ? bytes - o : chunk;
            cudaMemcpyAsync(d_in, (const char*)host_src + o, n, cudaMemcpyHostToDevice, s_in);
            cudaMemcpyAsync((char*)host_dst + o, d_out, n, cudaMemcpyDeviceToHost, s_out);
        }
    cudaEventRecord(e2, s_out);
    cudaStreamWaitEvent(s_in, e2, 0);
    cudaEventRecord(e1, s_in);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    cudaStreamDestroy(s_in);
    cudaStreamDestroy(s_out);
    cudaFree(d_in);
    cudaFree(d_out);
    checkError("custen_link_probe");
    return ms;
}

void* custen_managed_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaMallocManaged(&p, bytes);
    checkError("cudaMallocManaged");
    return p;
}
void custen_managed_free(void* p) { cudaFree(p); }

}  // extern "C"
