// cuSten-B200 streaming kernels (sm_100a), header form.
//
// Lives in a header because the Fun variants have two roads to the user's function:
//   * through the opaque device pointer the reference API carries (one indirect call per point), and
//   * inlined, when the translation unit that defines the function registers it with CUSTEN_REGISTER_FUN_*
//     (include/cuSten_fun.h): the kernel template below is then instantiated around the function itself and
//     Compute picks that instance when cuSten_t::devFunc equals the registered pointer.
//
// Design (all variants):
//   persistent CTAs, one producer warp + NT consumer threads.  The producer walks the CTA's work items
//   (column strip x row chunk) and feeds a ring of NS shared-memory stages with 1-D TMA bulk copies
//   (cp.async.bulk, completion on an mbarrier), ONE GRID ROW PER COPY.  Because the source of every row is
//   just an address, periodic wrap in x (up to three pieces per row), wrap / tile seams / remote halo rows
//   in y (band.top / band.bottom) are index arithmetic, never extra copies or special-cased blocks.
//   Every input element leaves HBM once; outputs are written once.
#ifndef CUSTEN_B200_STREAM_KERNELS_CUH
#define CUSTEN_B200_STREAM_KERNELS_CUH

#include "engine.h"
#include "weno_op.cuh"

namespace custen {

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

// ---- PTX helpers -------------------------------------------------------------------------------------------

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

// ---- slab time stepping: neighbour counters (Band::wait_* / signal_* / sync_local) ------------------------------

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
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Spin until the neighbour has published at least `target` finished sweeps.  A neighbour that never arrives (a bug, a
// dead rank) must not hang the GPU: after sync_local[3] nanoseconds the wait gives up and raises the sticky flag
// sync_local[2], which the host reports (custen_slab_error).  The proxy fence orders the TMA (async proxy) reads of
// the neighbour's rows behind the acquire.
__device__ __forceinline__ void slab_wait(const unsigned long long* flag, unsigned long long target,
                                          unsigned long long* local)
{
    if (ld_acquire_sys(flag) < target)
    {
        const unsigned long long limit = *(volatile unsigned long long*)(local + 3);
        const unsigned long long t0 = global_timer_ns();
        unsigned polls = 0;
        while (ld_acquire_sys(flag) < target)
        {
            if ((++polls & 255u) == 0 && limit && global_timer_ns() - t0 > limit)
            {
                *(volatile unsigned long long*)(local + 2) = 1ull;
                break;
            }
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}
// End of a sweep: called by the NT consumer threads once their last store has been issued.
__device__ __forceinline__ void slab_epilogue(const Band& b, int nconsumers)
{
    if (!b.sync_local) return;
    __threadfence_system();
    consumer_bar(nconsumers);
    if (threadIdx.x == 0)
    {
        __threadfence_system();
        const unsigned long long prev = atomicAdd(b.sync_local + 1, 1ull);
        if (prev == (unsigned long long)gridDim.x - 1)
        {
            __threadfence_system();
            volatile unsigned long long* loc = b.sync_local;
            loc[1] = 0ull;
            const unsigned long long done = loc[0] + 1ull;
            loc[0] = done;
            if (b.signal_up) st_release_sys(b.signal_up, done);
            if (b.signal_down) st_release_sys(b.signal_down, done);
        }
    }
}

// ---- geometry shared by host and device ----------------------------------------------------------------------

struct StreamArgs
{
    Band b;
    int TW;             // strip width in columns
    int Lp, Rp;         // halo widths rounded up to even (keeps every copy 16-byte aligned)
    int PW;             // shared-memory row pitch in doubles = Lp + TW + Rp
    int Beff;           // V - 1 - T: rows below the centre that the window reaches
    int PFX;            // rows kept in front of each stage (tile family: V - 1, acc family: 0)
    int nstrips, nchunks, chunk_rows, nitems;
    int stage_doubles;  // (PFX + SR) * PW
    int edge_last;      // slab time stepping: the first and last row chunk (the ones that touch halo rows) are swept last
    int warp_carry;     // tile family: who moves the PFX rows in front of a stage over from the previous stage.  1 = the
                        // producer warp, and the consumer warps hand a stage back one by one (no block-wide barrier: short
                        // windows, PFX <= SR / 4, and the compute-bound WENO operator); 0 = all consumers together behind a
                        // block-wide barrier (tall windows, where one warp is too slow: 9-row windows ran at 294 instead of
                        // 394 Gpoints/s with the producer copying)
};

// Order in which a CTA walks the row chunks.  In slab mode the two chunks that need a neighbour's rows come last, so
// that by the time a producer has to wait for a neighbour, the neighbour has normally long finished its previous sweep.
__device__ __forceinline__ int chunk_of(const StreamArgs& a, int q)
{
    if (!a.edge_last || a.nchunks <= 2) return q;
    return q < a.nchunks - 2 ? q + 1 : (q == a.nchunks - 2 ? 0 : a.nchunks - 1);
}

struct StageDesc
{
    int x0;      // first column of the strip
    int row0;    // band-local input row held by stage row 0
    int nrows;   // valid rows in this stage
    int out_lo;  // output rows this work item may write: [out_lo, out_hi)
    int out_hi;
    int first;   // 1: first stage of a work item (register accumulators restart)
    int pad[2];
};

constexpr int SMEM_BAR_OFF = 0;       // full[NS], empty[NS]
constexpr int SMEM_DESC_OFF = 128;    // NS descriptors of 32 B
constexpr int SMEM_COEF_OFF = 512;    // up to 128 coefficients
constexpr int SMEM_STAGE_OFF = 1536;  // stage ring
constexpr int MAX_SMEM_COEF = 128;

// Stages the producer will emit for this CTA (static round-robin over work items): the consumers count the same
// number, so the ring needs no termination message.
template <int SR>
__device__ __forceinline__ int cta_stage_count(const StreamArgs& a)
{
    int total = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x)
    {
        const int chunk = chunk_of(a, item / a.nstrips);
        const int out_lo = chunk * a.chunk_rows;
        const int out_hi = min(out_lo + a.chunk_rows, a.b.rows);
        total += (out_hi - out_lo + a.b.T + a.Beff + SR - 1) / SR;
    }
    return total;
}

// Producer: one warp.  Lane l owns stage row l: it works out where that grid row lives (band, top strip,
// bottom strip; wrapped columns) and issues up to three bulk copies for it.
template <int SR, int NS>
__device__ __forceinline__ void producer_loop(const StreamArgs& a, unsigned char* smem, int lane)
{
    static_assert(SR <= 32, "one producer lane per stage row");
    const Band& b = a.b;
    const uint32_t full0 = smem_u32(smem + SMEM_BAR_OFF);
    const uint32_t empty0 = full0 + 8 * NS;
    StageDesc* desc = reinterpret_cast<StageDesc*>(smem + SMEM_DESC_OFF);
    const uint32_t stage0 = smem_u32(smem + SMEM_STAGE_OFF);
    const uint32_t stage_bytes = (uint32_t)a.stage_doubles * 8u;
    const uint32_t pitch_bytes = (uint32_t)a.PW * 8u;

    // slab time stepping: this launch is sweep number `sweep` of the slab; a neighbour must have finished its sweep
    // `sweep - 1` before its halo rows are read (they are that sweep's output) and before the rows it reads from this
    // slab (the guard rows, that sweep's input) are overwritten
    const bool slab = b.sync_local != nullptr;
    const unsigned long long sweep = slab ? *(volatile unsigned long long*)b.sync_local : 0ull;
    bool ok_up = !slab || !b.wait_up, ok_down = !slab || !b.wait_down;

    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x)
    {
        const int strip = item % a.nstrips;
        const int chunk = chunk_of(a, item / a.nstrips);
        const int x0 = strip * a.TW;
        const int out_lo = chunk * a.chunk_rows;
        const int out_hi = min(out_lo + a.chunk_rows, b.rows);
        const int in_lo = out_lo - b.T;
        const int in_hi = out_hi + a.Beff;
        if (!ok_up && (in_lo < 0 || out_lo < b.guard_top))
        {
            slab_wait(b.wait_up, sweep, b.sync_local);
            ok_up = true;
        }
        if (!ok_down && (in_hi > b.rows || out_hi > b.rows - b.guard_bottom))
        {
            slab_wait(b.wait_down, sweep, b.sync_local);
            ok_down = true;
        }

        // unwrapped column range this strip needs: [u0, u1)
        const int xe = min(x0 + a.TW, b.nx);
        const int u0 = x0 - a.Lp, u1 = xe + a.Rp;
        // pieces: [u0,0) wrapped from the right edge, [max(u0,0), min(u1,nx)), [nx,u1) wrapped from the left edge
        const int m0 = max(u0, 0), m1 = min(u1, b.nx);
        const int lw = (b.wrap_x && u0 < 0) ? -u0 : 0;
        const int rw = (b.wrap_x && u1 > b.nx) ? u1 - b.nx : 0;
        const uint32_t row_bytes = (uint32_t)(m1 - m0 + lw + rw) * 8u;

        for (int r0 = in_lo; r0 < in_hi; r0 += SR)
        {
            const int nrows = min(SR, in_hi - r0);
            mbar_wait(empty0 + 8 * s, ph ^ 1);

            const int r = r0 + lane;
            const bool live = lane < nrows && band_row_exists(b, r);
            // warp_carry: the PFX rows in front of the stage are the last rows of the previous stage, copied over by this
            // warp (not for the first stage of an item, whose first PFX windows produce no output)
            const bool carry = a.warp_carry && a.PFX > 0 && r0 != in_lo;
            const uint32_t total = __reduce_add_sync(0xffffffffu, live ? row_bytes : 0u);

            const uint32_t bar = full0 + 8 * s;
            if (lane == 0)
            {
                StageDesc d;
                d.x0 = x0;
                d.row0 = r0;
                d.nrows = nrows;
                d.out_lo = out_lo;
                d.out_hi = out_hi;
                d.first = (r0 == in_lo);
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
            if (a.warp_carry)
            {
                // The stage's barrier takes 33 arrivals in this mode: the one above (with the byte count of the bulk
                // copies) and one per lane, after the carry rows are in place.  The previous stage (always a full one) must
                // have landed before its last PFX rows are read; its slot cannot be refilled under the copy, the
                // refill being issued by this same warp later.
                if (carry)
                {
                    const int sp = s == 0 ? NS - 1 : s - 1;
                    mbar_wait(full0 + 8 * sp, s == 0 ? ph ^ 1 : ph);
                    const double2* src = reinterpret_cast<const double2*>(smem + SMEM_STAGE_OFF + (size_t)sp * stage_bytes +
                                                                          (size_t)SR * pitch_bytes);
                    double2* dst2 = reinterpret_cast<double2*>(smem + SMEM_STAGE_OFF + (size_t)s * stage_bytes);
                    const int n2 = a.PFX * a.PW / 2;
                    for (int e = lane; e < n2; e += 32) dst2[e] = src[e];
                    __syncwarp();
                }
                mbar_arrive(bar);   // every lane, for the rows it copied
            }
            if (++s == NS) { s = 0; ph ^= 1; }
        }
    }
}

template <int NS>
__device__ __forceinline__ void stream_prologue(unsigned char* smem, int consumer_arrivals, int producer_arrivals = 1)
{
    if (threadIdx.x == 0)
    {
        const uint32_t full0 = smem_u32(smem + SMEM_BAR_OFF);
        for (int i = 0; i < NS; ++i)
        {
            mbar_init(full0 + 8 * i, producer_arrivals);
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

// ---- stream_acc_kernel: weights variants, compile-time H x V, weights in registers ----------------------------
// A consumer thread owns CPT adjacent columns (TW = CPT * NT) and sweeps down the rows.  For each arriving row it
// reads its window from shared memory (128-bit loads when CPT == 2) and feeds V partial sums, one per output row still in
// flight; the sum whose last tap row just arrived is stored with a 128-bit store and its slot restarts at 0.0.
// Each output's chain is therefore fma(w, v, sum) from sum = 0.0, rows top to bottom, taps left to right:
// the reference's order (2d_xy_p_kernel.cu:507-520), hence bit-identical results.

template <int NT, int SR, int NS, int H, int V, int LODD, int CPT, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB) stream_acc_kernel(const __grid_constant__ StreamArgs a)
{
    static_assert(CPT == 1 || CPT == 2, "columns per thread");
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

    double acc[V][CPT];
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[j][c] = 0.0;

    // CPT == 2: 16-byte loads covering pitch columns [2t, 2t + LODD + H]; CPT == 1: H 8-byte loads from t + LODD
    constexpr int NQ = (LODD + H + 2) / 2;
    constexpr int NW = CPT == 2 ? 2 * NQ : H;
    constexpr int W0 = CPT == 2 ? LODD : 0;
    constexpr bool ROT = (SR % V == 0) && V > 1;

    int s = 0;
    uint32_t ph = 0;
    const int nstages = cta_stage_count<SR>(a);
    for (int sc = 0; sc < nstages; ++sc)
    {
        mbar_wait(full0 + 8 * s, ph);
        const StageDesc d = desc[s];
        const double* buf = stage0 + (size_t)s * a.stage_doubles;
        const int gx = d.x0 + CPT * t;
        const ColMask cm = make_colmask(b, gx);
        double* obase = b.out + (ptrdiff_t)(d.row0 - a.Beff) * b.nx + gx;
        if (ROT && d.first)
        {
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[j][c] = 0.0;
        }

#pragma unroll
        for (int i = 0; i < SR; ++i)
        {
            if (i < d.nrows)
            {
                double win[NW];
                if (CPT == 2)
                {
                    const double2* rp = reinterpret_cast<const double2*>(buf + i * a.PW) + t;
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                    {
                        const double2 v = rp[q];
                        win[2 * q] = v.x;
                        win[2 * q + 1] = v.y;
                    }
                }
                else
                {
                    const double* rp = buf + i * a.PW + t + LODD;
#pragma unroll
                    for (int q = 0; q < H; ++q) win[q] = rp[q];
                }
                // Tap row j of this input row belongs to the output whose chain is in slot (i - j) mod V when the
                // stage height is a multiple of V (slots then line up from stage to stage and nothing moves);
                // otherwise slot j, with a shift after every row.
#pragma unroll
                for (int j = 0; j < V; ++j)
                {
                    constexpr int VV = V;
                    const int sl = ROT ? ((i - j) % VV + VV) % VV : j;
#pragma unroll
                    for (int ii = 0; ii < H; ++ii)
                    {
#pragma unroll
                        for (int c = 0; c < CPT; ++c) acc[sl][c] = fma(w[j * H + ii], win[W0 + ii + c], acc[sl][c]);
                    }
                }
                const int done = ROT ? ((i - (V - 1)) % V + V) % V : V - 1;  // slot whose last tap row just arrived
                const int yo = d.row0 + i - a.Beff;
                if (yo >= d.out_lo && yo < d.out_hi && yo >= b.ylo && yo < b.yhi)
                {
                    double* p = obase + (ptrdiff_t)i * b.nx;
                    if (CPT == 2) store_pair(p, acc[done][0], acc[done][CPT - 1], cm);
                    else if (cm.s0) *p = acc[done][0];
                    else if (cm.z0) *p = 0.0;
                }
                if (ROT)
                {
#pragma unroll
                    for (int c = 0; c < CPT; ++c) acc[done][c] = 0.0;
                }
                else
                {
#pragma unroll
                    for (int j = V - 1; j > 0; --j)
#pragma unroll
                        for (int c = 0; c < CPT; ++c) acc[j][c] = acc[j - 1][c];
#pragma unroll
                    for (int c = 0; c < CPT; ++c) acc[0][c] = 0.0;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if (++s == NS) { s = 0; ph ^= 1; }
    }
    slab_epilogue(b, NT);
}

// ---- stream_tile_kernel: a contiguous window in shared memory, handed to an operator --------------------------
// One column per consumer thread (TW = NT): a warp reads consecutive 8-byte words, conflict free.  Each stage
// buffer has PFX = V-1 rows in front of the rows the producer fills; the last V-1 rows of the previous stage are
// copied there (StreamArgs::warp_carry: by the producer warp once that stage has landed, or by all consumers behind a
// block-wide barrier), so every window is contiguous with pitch PW and the user function sees exactly the tile layout
// the reference gives it (data, loc, jump).

struct OpWeights  // run-time H x V weights (shapes without a register-accumulator instance)
{
    static __device__ __forceinline__ double apply(const Band& b, double* buf, double* cf, int tl, int PW, ptrdiff_t gidx)
    {
        double sum = 0.0;
        for (int j = 0; j < b.V; ++j)
            for (int i = 0; i < b.H; ++i) sum = fma(cf[j * b.H + i], buf[tl + j * PW + i], sum);
        return sum;
    }
};
struct OpPtrX  // opaque device pointer, X contract: loc = centre
{
    static __device__ __forceinline__ double apply(const Band& b, double* buf, double* cf, int tl, int PW, ptrdiff_t gidx)
    {
        return ((FunX)b.func)(buf, cf, tl + b.L);
    }
};
struct OpPtrY  // loc = centre, jump = pitch
{
    static __device__ __forceinline__ double apply(const Band& b, double* buf, double* cf, int tl, int PW, ptrdiff_t gidx)
    {
        return ((FunY)b.func)(buf, cf, tl + b.T * PW, PW);
    }
};
struct OpPtrXY  // loc = TOP-LEFT of the window (2d_xy_p_fun_kernel.cu:521)
{
    static __device__ __forceinline__ double apply(const Band& b, double* buf, double* cf, int tl, int PW, ptrdiff_t gidx)
    {
        return ((FunXY)b.func)(buf, cf, tl, PW, b.H, b.V);
    }
};
// Registered functions: the call is direct, so the compiler inlines the user's code into the sweep.
// LC / TC / HC / VC > 0 pin the stencil extents at compile time (loops in the user function then unroll).
template <FunX F, int LC>
struct OpInlineX
{
    static __device__ __forceinline__ double apply(const Band& b, double* buf, double* cf, int tl, int PW, ptrdiff_t gidx)
    {
        return F(buf, cf, tl + (LC >= 0 ? LC : b.L));
    }
};
template <FunY F, int TC>
struct OpInlineY
{
    static __device__ __forceinline__ double apply(const Band& b, double* buf, double* cf, int tl, int PW, ptrdiff_t gidx)
    {
        return F(buf, cf, tl + (TC >= 0 ? TC : b.T) * PW, PW);
    }
};
template <FunXY F, int HC, int VC>
struct OpInlineXY
{
    static __device__ __forceinline__ double apply(const Band& b, double* buf, double* cf, int tl, int PW, ptrdiff_t gidx)
    {
        return F(buf, cf, tl, PW, HC > 0 ? HC : b.H, VC > 0 ? VC : b.V);
    }
};

struct OpWeno
{
    // The streaming kernel walks a thread's rows in a rolled loop for this operator (one point is ~1300 instructions:
    // unrolled eight times the body overflowed the instruction cache) and fetches the next row's velocities from
    // global memory before it starts on the current row.
    static constexpr bool kRowLoop = true;
    static __device__ __forceinline__ double apply_uv(const Band& b, const double* buf, int tl, int PW, double u, double v)
    {
        const int c = tl + 3 * PW + 3;  // centre of the 7 x 7 window
        const double Fx = weno_line(buf, c, 1, u, b.p0);
        const double Fy = weno_line(buf, c, PW, v, b.p1);
        return u * Fx + v * Fy;
    }
    static __device__ __forceinline__ double apply(const Band& b, double* buf, double* cf, int tl, int PW, ptrdiff_t gidx)
    {
        return apply_uv(b, buf, tl, PW, b.aux0[gidx], b.aux1[gidx]);
    }
};
template <class Op, class = void> struct op_row_loop { static constexpr bool value = false; };
template <class Op> struct op_row_loop<Op, decltype((void)Op::kRowLoop)> { static constexpr bool value = true; };

template <int NT, int SR, int NS, int MINB, class Op>
__global__ void __launch_bounds__(NT + 32, MINB) stream_tile_kernel(const __grid_constant__ StreamArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    stream_prologue<NS>(smem, a.warp_carry ? NT : 1, a.warp_carry ? 33 : 1);

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
    const int dlt = a.Lp - b.L;  // the window of strip column c starts at pitch column c + dlt

    int s = 0;
    uint32_t ph = 0;
    const int nstages = cta_stage_count<SR>(a);
    for (int sc = 0; sc < nstages; ++sc)
    {
        mbar_wait(full0 + 8 * s, ph);
        const StageDesc d = desc[s];
        double* buf = stage0 + (size_t)s * a.stage_doubles;

        const int gx = d.x0 + t;
        const bool st = gx < b.nx && gx >= b.xlo && gx < b.xhi;
        const bool zr = gx < b.nx && b.zero_right && gx >= b.xhi;
        // rows of this stage that produce an output this item owns
        const int ybase = d.row0 - a.Beff;
        const int i_lo = max(0, max(d.out_lo, b.ylo) - ybase);
        const int i_hi = min(d.nrows, min(d.out_hi, b.yhi) - ybase);
        if (st)
        {
            double* o = b.out + (ptrdiff_t)ybase * b.nx + gx;
            const int tl0 = t + dlt;  // top-left of the window of stage row 0 (buffer row i <-> input row yo - T)
            if constexpr (op_row_loop<Op>::value)
            {
                if (i_lo < i_hi)
                {
                    const double* pu = b.aux0 + (ptrdiff_t)(ybase + i_lo) * b.nx + gx;
                    const double* pv = b.aux1 + (ptrdiff_t)(ybase + i_lo) * b.nx + gx;
                    double u = *pu, v = *pv;
#pragma unroll 1
                    for (int i = i_lo; i < i_hi; ++i)
                    {
                        double un = u, vn = v;
                        if (i + 1 < i_hi)
                        {
                            pu += b.nx;
                            pv += b.nx;
                            un = *pu;
                            vn = *pv;
                        }
                        o[(ptrdiff_t)i * b.nx] = Op::apply_uv(b, buf, tl0 + i * PW, PW, u, v);
                        u = un;
                        v = vn;
                    }
                }
            }
            else if (i_lo == 0 && i_hi == SR)
            {
#pragma unroll
                for (int i = 0; i < SR; ++i)
                    o[(ptrdiff_t)i * b.nx] = Op::apply(b, buf, cf, tl0 + i * PW, PW, (ptrdiff_t)(ybase + i) * b.nx + gx);
            }
            else
            {
                for (int i = i_lo; i < i_hi; ++i)
                    o[(ptrdiff_t)i * b.nx] = Op::apply(b, buf, cf, tl0 + i * PW, PW, (ptrdiff_t)(ybase + i) * b.nx + gx);
            }
        }
        else if (zr)
        {
            double* o = b.out + (ptrdiff_t)ybase * b.nx + gx;
            for (int i = i_lo; i < i_hi; ++i) o[(ptrdiff_t)i * b.nx] = 0.0;
        }

        if (a.warp_carry)
        {
            // the producer warp moves the carry rows: every consumer thread hands the stage back on its own (one arrival
            // per thread rather than "__syncwarp, lane 0 arrives": same speed, and compute-sanitizer's racecheck only
            // follows a thread's own arrival - with the per-warp form it reports the refill as racing with the other
            // lanes' reads, profiles/r2_compute_sanitizer.md)
            mbar_arrive(empty0 + 8 * s);
        }
        else
        {
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
        }
        if (++s == NS) { s = 0; ph ^= 1; }
    }
    slab_epilogue(b, NT);
}

// ---- launch plumbing shared by the library and by registering translation units ------------------------------

// Tile-family geometries (measured on B200, tools/tune_stream.cu): wide strips and one CTA per SM for 3-row
// windows, 256-column strips and two CTAs per SM otherwise.
struct TileSmall { static constexpr int NT = 256, SR = 8, NS = 3, MAXCPS = 2; };
struct TileBig { static constexpr int NT = 512, SR = 16, NS = 3, MAXCPS = 1; };
// WENO is compute-bound (18 single-precision powf per point): as many warps as the register file allows, and a
// consumer-warp count that is a multiple of four - 24 warps + the producer warp.  With 23 (736 threads) one scheduler
// had five consumer warps, ran ahead and waited a tenth of the time on the next stage's barrier (ncu source view):
// 736 -> 768 threads is +2.4 % on the reference example's fields and +5.6 % on random fields.  480 / 608-thread CTAs,
// two or three smaller CTAs per SM and three-stage rings measured within 3 % or slower
// (profiles/r2_weno_geom_v2.log, _v3.log, _v4.log).
struct TileWeno { static constexpr int NT = 768, SR = 8, NS = 2, MAXCPS = 1; };

struct LaunchGeom
{
    int grid;
    int threads;
    size_t smem;
};

// Defined in kernels.cu: raises the kernel's dynamic shared-memory limit, asks the occupancy calculator how many
// CTAs of `kernel` fit on an SM (an opaque user function may need any number of registers), caps that at
// max_cps, splits the band into work items accordingly and fills the item fields of `a`.
LaunchGeom plan_stream_launch(StreamArgs& a, const void* kernel, int threads, size_t smem, int max_cps);

// Fill the strip / pitch / stage fields of `a` for a tile-family geometry.
template <class G>
inline size_t tile_geometry(StreamArgs& a)
{
    a.TW = G::NT;
    a.PW = a.Lp + a.TW + a.Rp;
    a.nstrips = (a.b.nx + a.TW - 1) / a.TW;
    a.PFX = a.b.V - 1;
    a.warp_carry = (a.PFX * 4 <= G::SR || a.b.weno) ? 1 : 0;
    a.stage_doubles = (a.PFX + G::SR) * a.PW;
    return SMEM_STAGE_OFF + (size_t)G::NS * a.stage_doubles * sizeof(double);
}

template <class G, int MINB, class Op>
inline void launch_tile_geom(StreamArgs& a, cudaStream_t st)
{
    auto kernel = stream_tile_kernel<G::NT, G::SR, G::NS, MINB, Op>;
    const size_t smem = tile_geometry<G>(a);
    const LaunchGeom g = plan_stream_launch(a, (const void*)kernel, G::NT + 32, smem, G::MAXCPS);
    kernel<<<g.grid, g.threads, g.smem, st>>>(a);
}

// Opaque user functions: their register need is only known at device link, where a kernel compiled for a large CTA
// (a low per-thread register cap) would fail to link against a register-hungry callee (tried: with a 72-register
// cap nvlink refuses the library's own weighted_xy fixture, which ptxas gave 94 registers).  So the opaque road stays on
// 288-thread CTAs (cap 224 registers) and takes as many of them per SM as the occupancy calculator allows.
struct TileOpq256 { static constexpr int NT = 256, SR = 8, NS = 3, MAXCPS = 3; };

template <class Op>
inline void launch_tile_opaque(StreamArgs& a, cudaStream_t st)
{
    launch_tile_geom<TileOpq256, 1, Op>(a, st);
}

// BIG selects the wide geometry when the window is at most 3 rows tall (its stages then fit in shared memory).
template <bool BIG, int MINB, class Op>
inline void launch_tile_instance(StreamArgs& a, cudaStream_t st)
{
    if (BIG && a.b.V <= 3) launch_tile_geom<TileBig, 1, Op>(a, st);
    else launch_tile_geom<TileSmall, MINB, Op>(a, st);
}

// Signature of a registered (inlined) launcher: picks the instance for the band's extents and launches it.
typedef void (*InlineLauncher)(StreamArgs& a, cudaStream_t st);

constexpr int INLINE_MINB = 2;  // inlined instances are compiled for two CTAs per SM

template <FunX F>
inline void launch_inline_x(StreamArgs& a, cudaStream_t st)
{
    if (a.b.L == 1) launch_tile_instance<false, INLINE_MINB, OpInlineX<F, 1>>(a, st);
    else if (a.b.L == 4) launch_tile_instance<false, INLINE_MINB, OpInlineX<F, 4>>(a, st);
    else launch_tile_instance<false, INLINE_MINB, OpInlineX<F, -1>>(a, st);
}
template <FunY F>
inline void launch_inline_y(StreamArgs& a, cudaStream_t st)
{
    if (a.b.T == 1) launch_tile_instance<false, INLINE_MINB, OpInlineY<F, 1>>(a, st);
    else if (a.b.T == 4) launch_tile_instance<false, INLINE_MINB, OpInlineY<F, 4>>(a, st);
    else launch_tile_instance<false, INLINE_MINB, OpInlineY<F, -1>>(a, st);
}
template <FunXY F>
inline void launch_inline_xy(StreamArgs& a, cudaStream_t st)
{
    if (a.b.H == 3 && a.b.V == 3) launch_tile_instance<true, INLINE_MINB, OpInlineXY<F, 3, 3>>(a, st);
    else if (a.b.H == 5 && a.b.V == 5) launch_tile_instance<false, INLINE_MINB, OpInlineXY<F, 5, 5>>(a, st);
    else launch_tile_instance<false, INLINE_MINB, OpInlineXY<F, 0, 0>>(a, st);
}

// Registry of inlined instances, keyed by the device address of the user function.
constexpr int kMaxRegDevices = 16;
struct FunRegistration
{
    int dir;                               // Dir
    const void* (*resolve)();              // reads the device pointer on the current device (cudaMemcpyFromSymbol)
    InlineLauncher launch;
    const void* dev_ptr[kMaxRegDevices];   // per device, filled on first use there
    FunRegistration* next;
};
void register_fun(FunRegistration* r);     // defined in kernels.cu

}  // namespace custen

#endif
