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
#include "engine.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>

namespace custen {

// ------------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------------

typedef double (*FunX)(double*, double*, int);
typedef double (*FunY)(double*, double*, int, int);
typedef double (*FunXY)(double*, double*, int, int, int, int);

__device__ __forceinline__ const double* band_row(const Band& b, int r)
{
    // r is band-local: [-T, 0) -> top strip, [0, rows) -> the band, [rows, rows+B) -> bottom strip
    if (r < 0) return b.top + (ptrdiff_t)(r + b.T) * b.nx;
    if (r >= b.rows) return b.bottom + (ptrdiff_t)(r - b.rows) * b.nx;
    return b.in + (ptrdiff_t)r * b.nx;
}

__device__ __forceinline__ bool band_row_exists(const Band& b, int r)
{
    if (r < 0) return b.have_top && r >= -b.T;
    if (r >= b.rows) return b.have_bottom && r < b.rows + b.B;
    return true;
}

// ------------------------------------------------------------------------------------------------
// fallback family
// ------------------------------------------------------------------------------------------------

constexpr int FB_BX = 32;
constexpr int FB_BY = 8;

// MODE: 0 weights, 1 FunX, 2 FunY, 3 FunXY
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
    else sum = ((FunXY)b.func)(tile, cf, tl, PW, b.H, b.V);
    *o = sum;
}

// ------------------------------------------------------------------------------------------------
// streaming families: PTX helpers
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void consumer_bar(int nthreads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// streaming families: geometry shared by host and device
// ------------------------------------------------------------------------------------------------

struct StreamArgs
{
    Band b;
    int TW;             // strip width in columns (two per consumer thread)
    int Lp, Rp;         // halo widths rounded up to even (keeps every copy 16-byte aligned)
    int PW;             // shared-memory row pitch in doubles = Lp + TW + Rp
    int Beff;           // V - 1 - T: rows below the centre that the window reaches
    int PFX;            // rows kept in front of each stage (tile family: V - 1, acc family: 0)
    int nstrips, nchunks, chunk_rows, nitems;
    int stage_doubles;  // (PFX + SR) * PW
};

struct StageDesc
{
    int x0;      // first column of the strip
    int row0;    // band-local input row held by stage row 0
    int nrows;   // valid rows in this stage; < 0 terminates the consumers
    int out_lo;  // output rows this work item may write: [out_lo, out_hi)
    int out_hi;
    int pad[3];
};

constexpr int SMEM_BAR_OFF = 0;       // full[NS], empty[NS]
constexpr int SMEM_DESC_OFF = 128;    // NS descriptors of 32 B
constexpr int SMEM_COEF_OFF = 512;    // up to 128 coefficients
constexpr int SMEM_STAGE_OFF = 1536;  // stage ring
constexpr int MAX_SMEM_COEF = 128;

// Producer: one warp.  Lane l of the warp owns stage row l: it works out where that grid row lives
// (band, top strip, bottom strip; wrapped columns) and issues up to three bulk copies for it.
template <int SR, int NS>
__device__ __forceinline__ void producer_loop(const StreamArgs& a, unsigned char* smem, int lane)
{
    const Band& b = a.b;
    const uint32_t full0 = smem_u32(smem + SMEM_BAR_OFF);
    const uint32_t empty0 = full0 + 8 * NS;
    StageDesc* desc = reinterpret_cast<StageDesc*>(smem + SMEM_DESC_OFF);
    const uint32_t stage0 = smem_u32(smem + SMEM_STAGE_OFF);
    const uint32_t stage_bytes = (uint32_t)a.stage_doubles * 8u;
    const uint32_t pitch_bytes = (uint32_t)a.PW * 8u;

    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x)
    {
        const int strip = item % a.nstrips;
        const int chunk = item / a.nstrips;
        const int x0 = strip * a.TW;
        const int out_lo = chunk * a.chunk_rows;
        const int out_hi = min(out_lo + a.chunk_rows, b.rows);
        const int in_lo = out_lo - b.T;
        const int in_hi = out_hi + a.Beff;

        // unwrapped column range this strip needs: [u0, u1)
        const int xe = min(x0 + a.TW, b.nx);
        const int u0 = x0 - a.Lp, u1 = xe + a.Rp;

        for (int r0 = in_lo; r0 < in_hi; r0 += SR)
        {
            const int nrows = min(SR, in_hi - r0);
            mbar_wait(empty0 + 8 * s, ph ^ 1);

            const int r = r0 + lane;
            const bool live = lane < nrows && band_row_exists(b, r);
            // pieces: [u0,0) wrapped from the right edge, [max(u0,0), min(u1,nx)), [nx,u1) wrapped from the left edge
            const int m0 = max(u0, 0), m1 = min(u1, b.nx);
            const int lw = (b.wrap_x && u0 < 0) ? -u0 : 0;
            const int rw = (b.wrap_x && u1 > b.nx) ? u1 - b.nx : 0;
            uint32_t bytes = live ? (uint32_t)(m1 - m0 + lw + rw) * 8u : 0u;
            const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);

            const uint32_t bar = full0 + 8 * s;
            if (lane == 0)
            {
                StageDesc d;
                d.x0 = x0;
                d.row0 = r0;
                d.nrows = nrows;
                d.out_lo = out_lo;
                d.out_hi = out_hi;
                desc[s] = d;
                if (total) mbar_arrive_expect_tx(bar, total);
                else mbar_arrive(bar);
            }
            __syncwarp();
            if (live)
            {
                const double* src = band_row(b, r);
                const uint32_t dst = stage0 + s * stage_bytes + (uint32_t)(a.PFX + lane) * pitch_bytes;
                bulk_g2s(dst + (uint32_t)(m0 - u0) * 8u, src + m0, (uint32_t)(m1 - m0) * 8u, bar);
                if (lw) bulk_g2s(dst, src + (b.nx - lw), (uint32_t)lw * 8u, bar);
                if (rw) bulk_g2s(dst + (uint32_t)(b.nx - u0) * 8u, src, (uint32_t)rw * 8u, bar);
            }
            if (++s == NS) { s = 0; ph ^= 1; }
        }
    }
    // terminate the consumers
    mbar_wait(empty0 + 8 * s, ph ^ 1);
    if (lane == 0)
    {
        StageDesc d;
        d.x0 = 0; d.row0 = 0; d.nrows = -1; d.out_lo = 0; d.out_hi = 0;
        desc[s] = d;
        mbar_arrive(full0 + 8 * s);
    }
}

template <int NS>
__device__ __forceinline__ void stream_prologue(unsigned char* smem, int consumer_arrivals)
{
    if (threadIdx.x == 0)
    {
        const uint32_t full0 = smem_u32(smem + SMEM_BAR_OFF);
        for (int i = 0; i < NS; ++i)
        {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(full0 + 8 * (NS + i), consumer_arrivals);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// Store one finished pair of outputs honouring the non-periodic column masks.
struct ColMask
{
    bool s0, s1;  // store the computed value
    bool z0, z1;  // store 0.0 instead (Xnp right strip)
};
__device__ __forceinline__ ColMask make_colmask(const Band& b, int gx)
{
    ColMask m;
    const bool in0 = gx < b.nx, in1 = gx + 1 < b.nx;
    m.s0 = in0 && gx >= b.xlo && gx < b.xhi;
    m.s1 = in1 && gx + 1 >= b.xlo && gx + 1 < b.xhi;
    m.z0 = in0 && b.zero_right && gx >= b.xhi;
    m.z1 = in1 && b.zero_right && gx + 1 >= b.xhi;
    return m;
}
__device__ __forceinline__ void store_pair(double* p, double vx, double vy, const ColMask& m)
{
    if (m.s0 && m.s1) { *reinterpret_cast<double2*>(p) = make_double2(vx, vy); return; }
    if (m.s0) p[0] = vx; else if (m.z0) p[0] = 0.0;
    if (m.s1) p[1] = vy; else if (m.z1) p[1] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// stream_acc_kernel: weights variants, compile-time H x V, weights in registers
// ------------------------------------------------------------------------------------------------

template <int NT, int SR, int NS, int H, int V, int LODD>
__global__ void __launch_bounds__(NT + 32) stream_acc_kernel(const __grid_constant__ StreamArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    stream_prologue<NS>(smem, NT / 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == NT / 32)
    {
        producer_loop<SR, NS>(a, smem, lane);
        return;
    }

    const Band& b = a.b;
    const int t = threadIdx.x;
    const uint32_t full0 = smem_u32(smem + SMEM_BAR_OFF);
    const uint32_t empty0 = full0 + 8 * NS;
    const StageDesc* desc = reinterpret_cast<const StageDesc*>(smem + SMEM_DESC_OFF);
    const double* stage0 = reinterpret_cast<const double*>(smem + SMEM_STAGE_OFF);

    double w[H * V];
#pragma unroll
    for (int k = 0; k < H * V; ++k) w[k] = __ldg(b.coef + k);

    double ax[V], ay[V];
#pragma unroll
    for (int j = 0; j < V; ++j) ax[j] = ay[j] = 0.0;

    constexpr int NQ = (LODD + H + 2) / 2;  // 16-byte loads covering columns [2t, 2t + LODD + H]

    int s = 0;
    uint32_t ph = 0;
    for (;;)
    {
        mbar_wait(full0 + 8 * s, ph);
        const StageDesc d = desc[s];
        if (d.nrows < 0) break;
        const double* buf = stage0 + (size_t)s * a.stage_doubles;
        const int gx = d.x0 + 2 * t;
        const ColMask cm = make_colmask(b, gx);
        double* obase = b.out + (ptrdiff_t)(d.row0 - a.Beff) * b.nx + gx;

#pragma unroll
        for (int i = 0; i < SR; ++i)
        {
            if (i < d.nrows)
            {
                const double2* rp = reinterpret_cast<const double2*>(buf + i * a.PW) + t;
                double win[2 * NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                {
                    const double2 v = rp[q];
                    win[2 * q] = v.x;
                    win[2 * q + 1] = v.y;
                }
#pragma unroll
                for (int j = 0; j < V; ++j)
                {
#pragma unroll
                    for (int ii = 0; ii < H; ++ii)
                    {
                        ax[j] = fma(w[j * H + ii], win[LODD + ii], ax[j]);
                        ay[j] = fma(w[j * H + ii], win[LODD + ii + 1], ay[j]);
                    }
                }
                const int yo = d.row0 + i - a.Beff;
                if (yo >= d.out_lo && yo < d.out_hi && yo >= b.ylo && yo < b.yhi)
                    store_pair(obase + (ptrdiff_t)i * b.nx, ax[V - 1], ay[V - 1], cm);
#pragma unroll
                for (int j = V - 1; j > 0; --j)
                {
                    ax[j] = ax[j - 1];
                    ay[j] = ay[j - 1];
                }
                ax[0] = 0.0;
                ay[0] = 0.0;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if (++s == NS) { s = 0; ph ^= 1; }
    }
}

// ------------------------------------------------------------------------------------------------
// stream_tile_kernel: Fun variants (and weights with run-time H x V)
// MODE: 0 weights, 1 FunX, 2 FunY, 3 FunXY
// ------------------------------------------------------------------------------------------------

template <int NT, int SR, int NS, int MODE>
__global__ void __launch_bounds__(NT + 32) stream_tile_kernel(const __grid_constant__ StreamArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    stream_prologue<NS>(smem, 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == NT / 32)
    {
        producer_loop<SR, NS>(a, smem, lane);
        return;
    }

    const Band& b = a.b;
    const int t = threadIdx.x;
    const uint32_t full0 = smem_u32(smem + SMEM_BAR_OFF);
    const uint32_t empty0 = full0 + 8 * NS;
    const StageDesc* desc = reinterpret_cast<const StageDesc*>(smem + SMEM_DESC_OFF);
    double* stage0 = reinterpret_cast<double*>(smem + SMEM_STAGE_OFF);
    double* cf = reinterpret_cast<double*>(smem + SMEM_COEF_OFF);

    for (int k = t; k < b.ncoef; k += NT) cf[k] = b.coef[k];
    consumer_bar(NT);

    const int PW = a.PW, PFX = a.PFX;
    const int dlt = a.Lp - b.L;  // window of column c starts at pitch column c + dlt

    int s = 0;
    uint32_t ph = 0;
    for (;;)
    {
        mbar_wait(full0 + 8 * s, ph);
        const StageDesc d = desc[s];
        if (d.nrows < 0) break;
        double* buf = stage0 + (size_t)s * a.stage_doubles;

        // the thread's two columns are NT apart so that a warp reads consecutive 8-byte words
#pragma unroll
        for (int half = 0; half < 2; ++half)
        {
            const int c = t + half * NT;
            const int gx = d.x0 + c;
            if (gx >= b.nx) continue;
            const bool st = gx >= b.xlo && gx < b.xhi;
            const bool zr = b.zero_right && gx >= b.xhi;
            if (!st && !zr) continue;
            double* o = b.out + (ptrdiff_t)(d.row0 - a.Beff) * b.nx + gx;
            for (int i = 0; i < d.nrows; ++i)
            {
                const int yo = d.row0 + i - a.Beff;
                if (yo < d.out_lo || yo >= d.out_hi || yo < b.ylo || yo >= b.yhi) continue;
                const int tl = i * PW + c + dlt;  // top-left of the window (buffer row i == input row yo - T)
                double sum = 0.0;
                if (st)
                {
                    if (MODE == 0)
                    {
                        for (int j = 0; j < b.V; ++j)
                            for (int ii = 0; ii < b.H; ++ii) sum = fma(cf[j * b.H + ii], buf[tl + j * PW + ii], sum);
                    }
                    else if (MODE == 1) sum = ((FunX)b.func)(buf, cf, tl + b.L);
                    else if (MODE == 2) sum = ((FunY)b.func)(buf, cf, tl + b.T * PW, PW);
                    else sum = ((FunXY)b.func)(buf, cf, tl, PW, b.H, b.V);
                }
                o[(ptrdiff_t)i * b.nx] = sum;
            }
        }

        // carry the last V-1 rows over to the front of the next stage
        if (PFX > 0)
        {
            const int sn = (s + 1 == NS) ? 0 : s + 1;
            double* nxt = stage0 + (size_t)sn * a.stage_doubles;
            const double* src = buf + d.nrows * PW;
            for (int e = t; e < PFX * PW; e += NT) nxt[e] = src[e];
        }
        consumer_bar(NT);
        if (t == 0) mbar_arrive(empty0 + 8 * s);
        if (++s == NS) { s = 0; ph ^= 1; }
    }
}

// ------------------------------------------------------------------------------------------------
// host side: dispatch
// ------------------------------------------------------------------------------------------------

static std::atomic<uint64_t> g_launches{0};
uint64_t launches_total() { return g_launches.load(std::memory_order_relaxed); }

Tuning& tuning()
{
    static Tuning t = [] {
        Tuning x{};
        if (const char* e = getenv("CUSTEN_FORCE_FALLBACK")) x.force_fallback = atoi(e);
        if (const char* e = getenv("CUSTEN_FORCE_TILE")) x.force_tile = atoi(e);
        if (const char* e = getenv("CUSTEN_CHUNK_ROWS")) x.chunk_rows = atoi(e);
        if (const char* e = getenv("CUSTEN_CTAS_PER_SM")) x.ctas_per_sm = atoi(e);
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

static int launch_fallback(const Band& b, cudaStream_t st)
{
    const int Reff = b.H - 1 - b.L, Beff = b.V - 1 - b.T;
    const size_t smem = ((size_t)(FB_BX + b.L + Reff) * (FB_BY + b.T + Beff) + b.ncoef) * sizeof(double);
    dim3 grid((b.nx + FB_BX - 1) / FB_BX, (b.rows + FB_BY - 1) / FB_BY), block(FB_BX, FB_BY);
    const int mode = b.func ? (b.dir == DIR_X ? 1 : b.dir == DIR_Y ? 2 : 3) : 0;
#define FB_CASE(M)                                                                                          \
    case M:                                                                                                 \
        if (smem > 48 * 1024)                                                                               \
            cudaFuncSetAttribute(fallback_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        fallback_kernel<M><<<grid, block, smem, st>>>(b);                                                   \
        break;
    switch (mode)
    {
        FB_CASE(0) FB_CASE(1) FB_CASE(2) FB_CASE(3)
    }
#undef FB_CASE
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PATH_FALLBACK;
}

constexpr int K_NT = 128;  // consumer threads per CTA -> 256-column strips
constexpr int K_SR = 8;    // rows per stage
constexpr int K_NS = 4;    // stages in the ring

template <typename K>
static void launch_stream(K kernel, StreamArgs& a, cudaStream_t st)
{
    const size_t smem = SMEM_STAGE_OFF + (size_t)K_NS * a.stage_doubles * sizeof(double);
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int cps = tuning().ctas_per_sm;
    if (cps <= 0)
    {
        cps = (int)((220 * 1024) / (smem + 1024));
        if (cps > 3) cps = 3;
        if (cps < 1) cps = 1;
    }
    const int ncta = sm_count() * cps;

    // Choose the chunk height so that the item count is (just under) a whole number of waves.
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
    const int grid = a.nitems < ncta ? a.nitems : ncta;
    kernel<<<grid, K_NT + 32, smem, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    a.TW = 2 * K_NT;
    a.Lp = round_even(b.L);
    a.Rp = round_even(Reff);
    a.PW = a.Lp + a.TW + a.Rp;
    a.Beff = Beff;
    a.nstrips = (b.nx + a.TW - 1) / a.TW;

    const bool lodd = (b.L & 1) != 0;
    if (!b.func && !tu.force_tile)
    {
        a.PFX = 0;
        a.stage_doubles = K_SR * a.PW;
#define ACC_CASE(HH, VV)                                                                      \
    if (b.H == HH && b.V == VV)                                                               \
    {                                                                                         \
        if (lodd) launch_stream(stream_acc_kernel<K_NT, K_SR, K_NS, HH, VV, 1>, a, st);       \
        else launch_stream(stream_acc_kernel<K_NT, K_SR, K_NS, HH, VV, 0>, a, st);            \
        return PATH_STREAM_ACC;                                                               \
    }
        ACC_CASE(3, 1) ACC_CASE(5, 1) ACC_CASE(7, 1) ACC_CASE(9, 1)
        ACC_CASE(1, 3) ACC_CASE(1, 5) ACC_CASE(1, 7) ACC_CASE(1, 9)
        ACC_CASE(3, 3) ACC_CASE(5, 5)
#undef ACC_CASE
    }

    a.PFX = b.V - 1;
    a.stage_doubles = (a.PFX + K_SR) * a.PW;
    if (!b.func) launch_stream(stream_tile_kernel<K_NT, K_SR, K_NS, 0>, a, st);
    else if (b.dir == DIR_X) launch_stream(stream_tile_kernel<K_NT, K_SR, K_NS, 1>, a, st);
    else if (b.dir == DIR_Y) launch_stream(stream_tile_kernel<K_NT, K_SR, K_NS, 2>, a, st);
    else launch_stream(stream_tile_kernel<K_NT, K_SR, K_NS, 3>, a, st);
    return PATH_STREAM_TILE;
}

}  // namespace custen
