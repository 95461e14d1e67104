// cuSten-B200 device kernels (sm_100a) and their launcher.
//
// Semantics follow the reference kernels, restated as one index map instead of per-block cases:
//   weights variants   out(y,x) = sum_{j<V} sum_{i<H} w[j*H+i] * in[(y-T+j), (x-L+i)]      (FMA chain, j outer, i inner)
//                      cuSten/src/kernels/2d_x_p_kernel.cu:166-172, 2d_y_p_kernel.cu:121-127, 2d_xy_p_kernel.cu:507-520
//   Fun variants       out(y,x) = f(tile, coe, loc[, jump[, H, V]])
//                      2d_x_np_fun_kernel.cu:126-128, 2d_y_p_fun_kernel.cu:136-138, 2d_xy_p_fun_kernel.cu:521-526
//   non-periodic masks 2d_x_np_kernel.cu:137-176, 2d_y_np_kernel.cu:178-249, 2d_xy_np_kernel.cu:136-960
//
// Three kernel families:
//   stream_acc_kernel   persistent, warp-specialised.  One producer warp feeds a ring of shared-memory
//                       stages with 1-D TMA bulk copies (cp.async.bulk + mbarrier), one grid row per copy,
//                       so periodic wrap, tile seams and remote halo rows are all just a source address.
//                       Consumer threads own two adjacent columns and sweep down the rows keeping V partial
//                       sums in registers: every input element is read from HBM once and from shared memory
//                       once per row, every output written once with a 128-bit store.
//   stream_tile_kernel  same producer; the consumers keep the last V-1 rows in front of each stage so that a
//                       contiguous window exists in shared memory and hand it to the user function.
//   fallback_kernel     plain halo tile with scalar loads for shapes the TMA path cannot take
//                       (odd nx, rows not 16-byte aligned, very wide stencils).
#include "stream_kernels.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>

namespace custen {

// ------------------------------------------------------------------------------------------------
// fallback family
// ------------------------------------------------------------------------------------------------

constexpr int FB_BX = 32;
constexpr int FB_BY = 8;

// MODE: 0 weights, 1 FunX, 2 FunY, 3 FunXY, 4 WENO advection
template <int MODE>
__global__ void __launch_bounds__(FB_BX* FB_BY) fallback_kernel(const __grid_constant__ Band b)
{
    extern __shared__ double fb_smem[];
    const int Reff = b.H - 1 - b.L, Beff = b.V - 1 - b.T;
    const int PW = FB_BX + b.L + Reff;
    const int PH = FB_BY + b.T + Beff;
    double* tile = fb_smem;
    double* cf = fb_smem + PW * PH;

    const int tid = threadIdx.y * FB_BX + threadIdx.x;
    for (int k = tid; k < b.ncoef; k += FB_BX * FB_BY) cf[k] = b.coef[k];

    const int x0 = blockIdx.x * FB_BX;
    const int y0 = blockIdx.y * FB_BY;
    for (int e = tid; e < PW * PH; e += FB_BX * FB_BY)
    {
        const int c = e % PW, r = e / PW;
        int gx = x0 - b.L + c;
        const int gy = y0 - b.T + r;
        bool ok = band_row_exists(b, gy);
        if (gx < 0 || gx >= b.nx)
        {
            if (b.wrap_x) gx = ((gx % b.nx) + b.nx) % b.nx;
            else ok = false;
        }
        tile[e] = ok ? band_row(b, gy)[gx] : 0.0;
    }
    __syncthreads();

    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= b.nx || y >= b.rows) return;
    if (y < b.ylo || y >= b.yhi) return;
    double* o = b.out + (ptrdiff_t)y * b.nx + x;
    if (x < b.xlo) return;
    if (x >= b.xhi)
    {
        if (b.zero_right) *o = 0.0;
        return;
    }

    const int tl = threadIdx.y * PW + threadIdx.x;  // top-left of this point's window
    double sum = 0.0;
    if (MODE == 0)
    {
        for (int j = 0; j < b.V; ++j)
            for (int i = 0; i < b.H; ++i) sum = fma(cf[j * b.H + i], tile[tl + j * PW + i], sum);
    }
    else if (MODE == 1) sum = ((FunX)b.func)(tile, cf, tl + b.L);
    else if (MODE == 2) sum = ((FunY)b.func)(tile, cf, tl + b.T * PW, PW);
    else if (MODE == 3) sum = ((FunXY)b.func)(tile, cf, tl, PW, b.H, b.V);
    else sum = OpWeno::apply(b, tile, cf, tl, PW, (ptrdiff_t)y * b.nx + x);
    *o = sum;
}

// ------------------------------------------------------------------------------------------------
// host side: dispatch
// ------------------------------------------------------------------------------------------------

static std::atomic<uint64_t> g_launches{0};
uint64_t launches_total() { return g_launches.load(std::memory_order_relaxed); }
void launches_add(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

Tuning& tuning()
{
    static Tuning t = [] {
        Tuning x{};
        if (const char* e = getenv("CUSTEN_FORCE_FALLBACK")) x.force_fallback = atoi(e);
        if (const char* e = getenv("CUSTEN_FORCE_TILE")) x.force_tile = atoi(e);
        if (const char* e = getenv("CUSTEN_CHUNK_ROWS")) x.chunk_rows = atoi(e);
        if (const char* e = getenv("CUSTEN_CTAS_PER_SM")) x.ctas_per_sm = atoi(e);
        if (const char* e = getenv("CUSTEN_FORCE_OPAQUE")) x.force_opaque = atoi(e);
        return x;
    }();
    return t;
}

static int sm_count()
{
    static int per_dev[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!per_dev[dev]) cudaDeviceGetAttribute(&per_dev[dev], cudaDevAttrMultiProcessorCount, dev);
    return per_dev[dev];
}

// Slab time stepping on the fallback road (odd nx, unaligned rows): the plain-load kernel has no producer warp to do the
// waiting, so one thread waits in front of it and one publishes behind it.
__global__ void slab_wait_kernel(const __grid_constant__ Band b)
{
    const unsigned long long sweep = *(volatile unsigned long long*)b.sync_local;
    if (b.wait_up) slab_wait(b.wait_up, sweep, b.sync_local);
    if (b.wait_down) slab_wait(b.wait_down, sweep, b.sync_local);
}
__global__ void slab_signal_kernel(const __grid_constant__ Band b)
{
    __threadfence_system();
    volatile unsigned long long* loc = b.sync_local;
    const unsigned long long done = loc[0] + 1ull;
    loc[0] = done;
    if (b.signal_up) st_release_sys(b.signal_up, done);
    if (b.signal_down) st_release_sys(b.signal_down, done);
}

static int launch_fallback(const Band& b, cudaStream_t st)
{
    if (b.sync_local) slab_wait_kernel<<<1, 1, 0, st>>>(b);
    const int Reff = b.H - 1 - b.L, Beff = b.V - 1 - b.T;
    const size_t smem = ((size_t)(FB_BX + b.L + Reff) * (FB_BY + b.T + Beff) + b.ncoef) * sizeof(double);
    dim3 grid((b.nx + FB_BX - 1) / FB_BX, (b.rows + FB_BY - 1) / FB_BY), block(FB_BX, FB_BY);
    const int mode = b.weno ? 4 : b.func ? (b.dir == DIR_X ? 1 : b.dir == DIR_Y ? 2 : 3) : 0;
#define FB_CASE(M)                                                                                          \
    case M:                                                                                                 \
        if (smem > 48 * 1024)                                                                               \
            cudaFuncSetAttribute(fallback_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        fallback_kernel<M><<<grid, block, smem, st>>>(b);                                                   \
        break;
    switch (mode)
    {
        FB_CASE(0) FB_CASE(1) FB_CASE(2) FB_CASE(3) FB_CASE(4)
    }
#undef FB_CASE
    if (b.sync_local) slab_signal_kernel<<<1, 1, 0, st>>>(b);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PATH_FALLBACK;
}

// ---- streaming launch geometry -----------------------------------------------------------------------------
// Register-accumulator family: geometry per window shape, from a sweep on B200 (tools/tune_stream.cu):
// consumer threads, columns per thread, rows per stage, ring depth, CTAs per SM.  A stage height that is a
// multiple of V lets the accumulator slots rotate by renaming instead of by register moves.
template <int H, int V> struct AccGeom { static constexpr int NT = 512, CPT = 1, SR = 8, NS = 3, MAXCPS = 1; };
template <int H> struct AccGeom<H, 1> { static constexpr int NT = 256, CPT = 2, SR = 8, NS = 3, MAXCPS = 1; };
template <> struct AccGeom<1, 3> { static constexpr int NT = 256, CPT = 1, SR = 6, NS = 3, MAXCPS = 2; };
template <> struct AccGeom<1, 5> { static constexpr int NT = 512, CPT = 1, SR = 10, NS = 3, MAXCPS = 1; };
template <> struct AccGeom<1, 7> { static constexpr int NT = 512, CPT = 1, SR = 7, NS = 3, MAXCPS = 1; };
template <> struct AccGeom<1, 9> { static constexpr int NT = 512, CPT = 1, SR = 9, NS = 3, MAXCPS = 1; };
// (5,5): the generic geometry; NT128 x 2 columns x 2 CTAs measured 320 Gpoints/s against 350 for it once placement is pinned

// Work decomposition: column strips x row chunks, chunk height chosen so that the item count is (just under)
// a whole number of waves of resident CTAs.
// The occupancy query and the shared-memory opt-in depend only on (kernel, device, request): they are made once and
// remembered, so a steady-state launch issues no runtime calls besides the launch itself (and can be stream-captured).
struct LaunchMemo
{
    const void* kernel;
    int device, threads, max_cps, tuned_cps;
    size_t smem_in, smem_out;
    int cps;
};
static LaunchMemo g_memo[256];
static std::atomic<int> g_memo_n{0};
static std::atomic_flag g_memo_lock = ATOMIC_FLAG_INIT;

static void resolve_occupancy(const void* kernel, int threads, size_t& smem, int max_cps, int& cps)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const int tuned = tuning().ctas_per_sm;
    const int n = g_memo_n.load(std::memory_order_acquire);
    for (int i = 0; i < n; ++i)
    {
        const LaunchMemo& m = g_memo[i];
        if (m.kernel == kernel && m.device == dev && m.threads == threads && m.smem_in == smem && m.max_cps == max_cps &&
            m.tuned_cps == tuned)
        {
            smem = m.smem_out;
            cps = m.cps;
            return;
        }
    }
    const size_t smem_in = smem;
    // the opt-in is one number per (kernel, device): it only ever grows, so that every remembered request stays launchable
    size_t opted = 0;
    for (int i = 0; i < n; ++i)
        if (g_memo[i].kernel == kernel && g_memo[i].device == dev && g_memo[i].smem_out > opted) opted = g_memo[i].smem_out;
    if (smem > opted)
    {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        opted = smem;
    }
    cps = tuned;
    if (cps <= 0)
    {
        cps = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kernel, threads, smem);
        if (max_cps > 0 && cps > max_cps) cps = max_cps;
        if (cps < 1) cps = 1;
        // The geometry wants exactly `cps` CTAs on every SM.  If more would fit, the hardware scheduler is free to
        // double up on some SMs and leave others empty (measured: bistable 290 / 350 Gpoints/s for the same launch), so
        // the request is padded until cps + 1 no longer fit.
        const size_t sm_bytes = 228 * 1024, per_cta_reserved = 1024;
        const size_t need = sm_bytes / (size_t)(cps + 1) - per_cta_reserved + 256;
        if (smem < need && (need + per_cta_reserved) * (size_t)cps <= sm_bytes && need <= 227 * 1024)
        {
            smem = need;
            if (smem > opted) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
    }
    while (g_memo_lock.test_and_set(std::memory_order_acquire)) {}
    const int k = g_memo_n.load(std::memory_order_relaxed);
    if (k < 256)
    {
        g_memo[k] = LaunchMemo{kernel, dev, threads, max_cps, tuned, smem_in, smem, cps};
        g_memo_n.store(k + 1, std::memory_order_release);
    }
    g_memo_lock.clear(std::memory_order_release);
}

LaunchGeom plan_stream_launch(StreamArgs& a, const void* kernel, int threads, size_t smem, int max_cps)
{
    int cps = 1;
    resolve_occupancy(kernel, threads, smem, max_cps, cps);
    const int ncta = sm_count() * cps;
    const Band& b = a.b;
    int ch = tuning().chunk_rows;
    if (ch <= 0)
    {
        const int target = 192;
        long items_t = (long)a.nstrips * ((b.rows + target - 1) / target);
        long waves = (items_t + ncta / 2) / ncta;
        if (waves < 1) waves = 1;
        long nch = (waves * ncta) / a.nstrips;
        if (nch < 1) nch = 1;
        ch = (int)((b.rows + nch - 1) / nch);
        if (ch < 16) ch = 16;
    }
    if (ch > b.rows) ch = b.rows;
    a.chunk_rows = ch;
    a.nchunks = (b.rows + ch - 1) / ch;
    a.nitems = a.nstrips * a.nchunks;
    a.edge_last = b.sync_local != nullptr;
    LaunchGeom g;
    g.grid = a.nitems < ncta ? a.nitems : ncta;
    g.threads = threads;
    g.smem = smem;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return g;
}

template <int H, int V, int LODD>
static void launch_acc(StreamArgs& a, cudaStream_t st)
{
    typedef AccGeom<H, V> G;
    a.TW = G::CPT * G::NT;
    a.PW = a.Lp + a.TW + a.Rp;
    a.nstrips = (a.b.nx + a.TW - 1) / a.TW;
    a.PFX = 0;
    a.stage_doubles = G::SR * a.PW;
    auto kernel = stream_acc_kernel<G::NT, G::SR, G::NS, H, V, LODD, G::CPT, 1>;
    const size_t smem = SMEM_STAGE_OFF + (size_t)G::NS * a.stage_doubles * sizeof(double);
    const LaunchGeom g = plan_stream_launch(a, (const void*)kernel, G::NT + 32, smem, G::MAXCPS);
    kernel<<<g.grid, g.threads, g.smem, st>>>(a);
}

// ---- registry of inlined user functions ---------------------------------------------------------------------

static FunRegistration* g_registry = nullptr;

void register_fun(FunRegistration* r)
{
    r->next = g_registry;  // static-initialisation time: single threaded, no CUDA calls here
    for (int d = 0; d < kMaxRegDevices; ++d) r->dev_ptr[d] = nullptr;
    g_registry = r;
}

// A device function's address is a per-device fact: it is read (cudaMemcpyFromSymbol, in the registering translation
// unit) once per device, on the device the launch is about to happen on.
static InlineLauncher find_inline(const void* func, int dir)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxRegDevices) return nullptr;
    for (FunRegistration* r = g_registry; r; r = r->next)
    {
        if (r->dir != dir) continue;
        if (!r->dev_ptr[dev]) r->dev_ptr[dev] = r->resolve();
        if (r->dev_ptr[dev] == func) return r->launch;
    }
    return nullptr;
}

static int round_even(int v) { return (v + 1) & ~1; }

int launch_band(const Band& b, cudaStream_t st)
{
    if (b.rows <= 0 || b.nx <= 0) return PATH_NONE;
    const Tuning& tu = tuning();
    const int Reff = b.H - 1 - b.L, Beff = b.V - 1 - b.T;

    // Conditions of the TMA path: 16-byte aligned rows, halos that fit, sane window.
    bool ok = !tu.force_fallback;
    ok = ok && (b.nx % 2 == 0) && (((uintptr_t)b.in | (uintptr_t)b.out) % 16 == 0);
    ok = ok && (!b.have_top || ((uintptr_t)b.top % 16 == 0)) && (!b.have_bottom || ((uintptr_t)b.bottom % 16 == 0));
    ok = ok && Reff >= 0 && Beff >= 0 && b.L >= 0 && b.T >= 0;
    ok = ok && b.L <= 8 && Reff <= 8 && b.T <= 8 && Beff <= 8 && b.ncoef <= MAX_SMEM_COEF;
    ok = ok && b.nx >= 16 && b.rows >= b.T && b.rows >= Beff;
    ok = ok && (b.B >= Beff || !b.have_bottom) && (b.R >= 0);
    if (!ok) return launch_fallback(b, st);

    StreamArgs a{};
    a.b = b;
    a.Lp = round_even(b.L);
    a.Rp = round_even(Reff);
    a.Beff = Beff;

    if (b.weno)
    {
        launch_tile_geom<TileWeno, 1, OpWeno>(a, st);
        return PATH_STREAM_TILE;
    }
    const bool lodd = (b.L & 1) != 0;
    if (!b.func && !tu.force_tile)
    {
#define ACC_CASE(HH, VV)                                   \
    if (b.H == HH && b.V == VV)                            \
    {                                                      \
        if (lodd) launch_acc<HH, VV, 1>(a, st);            \
        else launch_acc<HH, VV, 0>(a, st);                 \
        return PATH_STREAM_ACC;                            \
    }
        ACC_CASE(3, 1) ACC_CASE(5, 1) ACC_CASE(7, 1) ACC_CASE(9, 1)
        ACC_CASE(1, 3) ACC_CASE(1, 5) ACC_CASE(1, 7) ACC_CASE(1, 9)
        ACC_CASE(3, 3) ACC_CASE(5, 5)
#undef ACC_CASE
    }

    if (!b.func)
    {
        launch_tile_instance<false, 2, OpWeights>(a, st);
        return PATH_STREAM_TILE;
    }
    if (!tu.force_opaque)
    {
        if (InlineLauncher il = find_inline(b.func, b.dir))
        {
            il(a, st);
            return PATH_STREAM_INLINE;
        }
    }
    // opaque pointer: no minimum-blocks bound, the callee's register need is unknown until device link
    if (b.dir == DIR_X) launch_tile_opaque<OpPtrX>(a, st);
    else if (b.dir == DIR_Y) launch_tile_opaque<OpPtrY>(a, st);
    else launch_tile_opaque<OpPtrXY>(a, st);
    return PATH_STREAM_TILE;
}

}  // namespace custen
