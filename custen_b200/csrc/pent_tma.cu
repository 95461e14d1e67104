// The batched pentadiagonal solve of the Cahn-Hilliard ADI step, fed by the TMA engine (see the design note below).
//
// The math warp of this kernel is alone on its scheduler and issues in order, so the instruction ORDER written here is
// the optimisation: the shared-memory traffic of a row has to sit in the idle issue slots between the dependent FP64
// operations of the recurrence.  Left to itself ptxas schedules for latency hiding by other warps and moves all of it
// behind (or in front of) the chain, where it costs ~13 cycles per row (ncu source view).  Every operation of the math
// warp is therefore a volatile PTX statement; nvcc and ptxas keep their order (checked with cuobjdump -sass; the same
// source built with -Xptxas -O1 gives the same order and is ~2 % slower).
#include "pent_solve.h"

#include <cuda.h>

#include <cstdint>
#include <type_traits>

namespace custen_cahn {

// ---- the same solve fed by the TMA engine -----------------------------------------------------------------------------
// A warp that is alone on its scheduler issues in order, so every instruction that is not one of the recurrence's
// dependent FP64 operations (8 cycles each, tools/fp64_latency.cu: 4 per row forward, 2 backward) has to fit in the gaps
// between them.  The cp.async version (k_pent_solve_smem, cahn.cu) spends ~17 instructions per row (per-lane 64-bit address
// arithmetic for every LDGSTS and STG) and runs at ~57 cycles per row and direction.  Here a CTA is two warps on two schedulers:
//   * the math warp touches shared memory only: per group of TG rows it waits for the group's slot, keeps a rolling
//     window of TW rows (right-hand sides and coefficients) in registers, runs the chain, writes the results back into
//     the slot and arrives on the slot's `done` barrier.  Per row: LDS rhs, 2 x LDS.128 coefficients, the chain, STS;
//   * one lane of the copy warp moves everything through the async proxy: one 2-D tensor copy brings TG rows x 32
//     systems into a slot, one 1-D bulk copy brings the group's coefficients (a table pre-interleaved at factor time),
//     one tensor store takes the results back.  Loads run TAHEAD groups ahead; a slot is refilled once its store has
//     finished reading it, TLAG groups after it was issued.
// Rows 0, 1 (forward) and m-1, m-2 (backward) need no special code: their missing terms have zero coefficients in the
// tables and x - 0*y is exact, so the operation sequence per row is the reference's (cuPentBatch.cu:153-195).
constexpr int TW = 8;           // rows held in registers by the math warp
constexpr int TSLOTS = 16;      // slots in the ring (9 KB each)
constexpr int TLAG = 4;         // stores that may still be reading their slot
constexpr int TAHEAD = TSLOTS - TLAG;
constexpr size_t TMA_SMEM =
    (size_t)TSLOTS * (TG * 32 + TG * 4) * sizeof(double) + 2 * TSLOTS * sizeof(unsigned long long) + 32 * sizeof(double);

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_group(const CUtensorMap* tm, double* sdata, double* scoef, const double* tab,
                                               unsigned tab_bytes, int sys0, int row0, unsigned bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TG * 32 * 8 + tab_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_addr(sdata)),
                 "l"(tm), "r"(sys0), "r"(row0), "r"(bar)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(scoef)),
                 "l"(tab), "r"(tab_bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_store_group(const CUtensorMap* tm, const double* slot, int sys0, int row0)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(sys0), "r"(row0),
                 "r"(smem_addr(slot))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// ---- the math warp ------------------------------------------------------------------------------------------------------
// Every operation of the math warp is a volatile PTX statement, so the statement order below is the schedule.  It is a
// software pipeline over sweep positions p (forward: row p of the
// group, backward: row TG-1-p):
//     chain operation of position p            (each waits ~8 cycles for its predecessor)
//     "fillers" of position p-1 in the gaps:   store x[p-1] into its slot, fetch position p-1+TW into the registers
//                                              position p-1 just vacated, and (forward) the off-chain product of p+1
// so nothing that depends on a fresh result sits between two chain operations (ncu showed the store and the fetch
// address of a row serialised behind its own result: 12 cycles per row forward, 10 backward).
// The tables hold NEGATED coefficients (PTX fma has no operand negation): fma(-a, b, c) == fma(a, -b, c) exactly.
__device__ __forceinline__ double vfma(double a, double b, double c)
{
    double d;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
    return d;
}
__device__ __forceinline__ double vmul(double a, double b)
{
    double d;
    asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b));
    return d;
}
template <int OFF>
__device__ __forceinline__ double lds_f64(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ double2 lds_v2f64(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f64(unsigned addr, double v)
{
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v) : "memory");
}

// Rolling window of TW sweep positions held in registers: position p lives in window entry p % TW.
struct FwdWin
{
    double2 c0[TW], c1[TW];  // {-ds, -dl}, {-d, 1/d}
    double r[TW];
};
struct BwdWin
{
    double2 c[TW];           // {-du, -dw}
    double r[TW];
};
struct Addr
{
    unsigned rb_prv, rb_cur, rb_nxt;   // this lane's column in the previous / current / next slot's rows
    unsigned cf_cur, cf_nxt;           // coefficient rows of the current / next slot
};
struct Carry
{
    double x1, x2;   // the two most recent results (forward: x[i-1], x[i-2]; backward: x[i+1], x[i+2])
    double t;        // forward only: r[i] - ds[i] x[i-2] of the position about to be solved
};

template <bool FWD>
__host__ __device__ constexpr int row_at(int pos) { return FWD ? pos : TG - 1 - pos; }

// Fillers of position P - 1 (P = 0: the last position of the previous group): where its result goes and which
// position's data its registers take next.
template <bool FWD, int P>
struct Fill
{
    static constexpr int Q = P - 1;                                  // the position being retired
    static constexpr int W = (Q + TW) % TW;                          // its window entry
    static constexpr bool ST_PRV = Q < 0;                            // store into the previous slot
    static constexpr int ST_ROW = row_at<FWD>(Q < 0 ? TG - 1 : Q);
    static constexpr int NP = Q + TW;                                // position fetched into entry W
    static constexpr bool LD_NXT = NP >= TG;                         // ... from the next slot
    static constexpr int LD_ROW = row_at<FWD>(LD_NXT ? NP - TG : NP);
};

template <int P>
__device__ __forceinline__ void fwd_pos(FwdWin& q, const Addr& a, Carry& k)
{
    typedef Fill<true, P> F;
    constexpr int S = P % TW, S1 = (P + 1) % TW;
    const unsigned rb_ld = F::LD_NXT ? a.rb_nxt : a.rb_cur, cf_ld = F::LD_NXT ? a.cf_nxt : a.cf_cur;
    // x = (r - ds x2 - dl x1) / d as q = u (1/d), rem = fma(-d, q, u), x = fma(1/d, rem, q)   (cuPentBatch.cu:153-172)
    const double u = vfma(q.c0[S].y, k.x1, k.t);
    sts_f64<F::ST_ROW * 256>(F::ST_PRV ? a.rb_prv : a.rb_cur, k.x1);
    const double tn = vfma(q.c0[S1].x, k.x1, q.r[S1]);              // r - ds x2 of position P + 1
    const double qq = vmul(u, q.c1[S].y);
    q.r[F::W] = lds_f64<F::LD_ROW * 256>(rb_ld);
    const double rem = vfma(qq, q.c1[S].x, u);
    q.c0[F::W] = lds_v2f64<F::LD_ROW * 32>(cf_ld);
    const double x = vfma(q.c1[S].y, rem, qq);
    q.c1[F::W] = lds_v2f64<F::LD_ROW * 32 + 16>(cf_ld);
    k.x2 = k.x1;
    k.x1 = x;
    k.t = tn;
}
template <int P>
__device__ __forceinline__ void bwd_pos(BwdWin& q, const Addr& a, Carry& k)
{
    typedef Fill<false, P> F;
    constexpr int S = P % TW;
    const unsigned rb_ld = F::LD_NXT ? a.rb_nxt : a.rb_cur, cf_ld = F::LD_NXT ? a.cf_nxt : a.cf_cur;
    // x = (r - du x1) - dw x2   (cuPentBatch.cu:176-195)
    const double t = vfma(q.c[S].x, k.x1, q.r[S]);
    sts_f64<F::ST_ROW * 256>(F::ST_PRV ? a.rb_prv : a.rb_cur, k.x1);
    q.r[F::W] = lds_f64<F::LD_ROW * 256>(rb_ld);
    const double x = vfma(q.c[S].y, k.x2, t);
    q.c[F::W] = lds_v2f64<F::LD_ROW * 16>(cf_ld);
    k.x2 = k.x1;
    k.x1 = x;
}
// positions LO .. HI of a group, unrolled at compile time
template <bool FWD, int LO, int HI>
struct Positions
{
    template <class Win>
    static __device__ __forceinline__ void run(Win& q, const Addr& a, Carry& k)
    {
        if constexpr (FWD) fwd_pos<LO>(q, a, k);
        else bwd_pos<LO>(q, a, k);
        if constexpr (LO < HI) Positions<FWD, LO + 1, HI>::run(q, a, k);
    }
};
// the first TW positions of a sweep
__device__ __forceinline__ void window_first_load(FwdWin& q, unsigned rb, unsigned cf)
{
#pragma unroll
    for (int p = 0; p < TW; ++p)
    {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(q.r[p]) : "r"(rb + p * 256));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q.c0[p].x), "=d"(q.c0[p].y) : "r"(cf + p * 32));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q.c1[p].x), "=d"(q.c1[p].y) : "r"(cf + p * 32 + 16));
    }
}
__device__ __forceinline__ void window_first_load(BwdWin& q, unsigned rb, unsigned cf)
{
#pragma unroll
    for (int p = 0; p < TW; ++p)
    {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(q.r[p]) : "r"(rb + (TG - 1 - p) * 256));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q.c[p].x), "=d"(q.c[p].y) : "r"(cf + (TG - 1 - p) * 16));
    }
}

// Slot and barrier parity of the i-th group a CTA handles, counted over both sweeps.
__device__ __forceinline__ unsigned slot_of(unsigned i) { return i % TSLOTS; }
__device__ __forceinline__ unsigned parity_of(unsigned i) { return (i / TSLOTS) & 1; }

// A value the optimiser must keep in a register (ptxas would otherwise re-derive shared-memory addresses at every use).
__device__ __forceinline__ unsigned pinned(unsigned v)
{
    unsigned r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
// One poll of an mbarrier phase (no spinning): its latency can then sit underneath other work.
__device__ __forceinline__ unsigned mbar_test_parity(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// results of a group -> async proxy (the tensor store reads them), then tell the copy warp
__device__ __forceinline__ void publish(unsigned done_bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(done_bar) : "memory");
}

// Math warp, one sweep over the ngroups = nrows / TG row groups.  FWD: top to bottom, 4 coefficients per row; !FWD:
// bottom to top, 2 per row.  The result of a group's last position is stored by the first position of the next group
// (software pipeline), so a group is published one position late; `dummy` is a 256-byte row that takes the store
// of the (non-existent) position before the first one.
template <bool FWD>
__device__ __forceinline__ void math_sweep(unsigned sdata, unsigned scoef, unsigned full, unsigned done, unsigned dummy,
                                           int ngroups, int lane, unsigned& it)
{
    typedef typename std::conditional<FWD, FwdWin, BwdWin>::type Window;
    const unsigned b_rb = pinned(sdata + lane * 8), b_cf = pinned(scoef), b_full = pinned(full), b_done = pinned(done);
    constexpr int LAST_ROW = row_at<FWD>(TG - 1);
    Window q;
    unsigned s_cur = slot_of(it), ph_cur = parity_of(it);
    mbar_wait_parity(b_full + 8 * s_cur, ph_cur);
    window_first_load(q, b_rb + s_cur * (TG * 256), b_cf + s_cur * (TG * 32));
    Carry k;
    k.x1 = k.x2 = 0.0;
    k.t = 0.0;
    if constexpr (FWD) k.t = vfma(q.c0[0].x, 0.0, q.r[0]);
    unsigned s_nxt = s_cur + 1 == TSLOTS ? 0 : s_cur + 1;
    unsigned ph_nxt = s_cur + 1 == TSLOTS ? ph_cur ^ 1 : ph_cur;
    unsigned landed = ngroups > 1 ? mbar_test_parity(b_full + 8 * s_nxt, ph_nxt) : 1u;
    unsigned rb_prv = dummy + lane * 8 - LAST_ROW * 256;   // position -1 of the sweep "stores" into the dummy row
    unsigned done_prv = 0;
    for (int g = 0; g < ngroups; ++g)
    {
        // the next group's copy must have landed: its first rows are fetched underneath this group's chain.  (After the
        // last group the fetches re-read the current slot; the values are not used.)
        const bool has_next = g + 1 < ngroups;
        if (has_next && !landed) mbar_wait_parity(b_full + 8 * s_nxt, ph_nxt);
        const unsigned s_src = has_next ? s_nxt : s_cur;
        Addr a;
        a.rb_prv = rb_prv;
        a.rb_cur = b_rb + s_cur * (TG * 256);
        a.rb_nxt = b_rb + s_src * (TG * 256);
        a.cf_cur = b_cf + s_cur * (TG * 32);
        a.cf_nxt = b_cf + s_src * (TG * 32);
        const unsigned s_nn = s_nxt + 1 == TSLOTS ? 0 : s_nxt + 1;
        const unsigned ph_nn = s_nxt + 1 == TSLOTS ? ph_nxt ^ 1 : ph_nxt;
        Positions<FWD, 0, 0>::run(q, a, k);                 // ... which also stores the previous group's last result
        if (g > 0) publish(done_prv);
        Positions<FWD, 1, TG / 2 - 1>::run(q, a, k);
        landed = mbar_test_parity(b_full + 8 * s_nn, ph_nn);
        Positions<FWD, TG / 2, TG - 1>::run(q, a, k);
        rb_prv = a.rb_cur;
        done_prv = b_done + 8 * s_cur;
        s_cur = s_nxt;
        ph_cur = ph_nxt;
        s_nxt = s_nn;
        ph_nxt = ph_nn;
    }
    // the sweep's last result is still in a register
    sts_f64<LAST_ROW * 256>(rb_prv, k.x1);
    publish(done_prv);
    it += ngroups;
}

// Copy lane, one sweep.  Group g of the sweep is the CTA's group i0 + g; its rows start at row_of(g).
template <bool FWD>
__device__ __forceinline__ void copy_sweep(const CUtensorMap* tm, const double* __restrict__ tab, double (*sdata)[TG * 32],
                                           double (*scoef)[TG * 4], unsigned long long* full, unsigned long long* done,
                                           int nrows, int ngroups, int sys0, unsigned& it)
{
    constexpr int PER_ROW = FWD ? 4 : 2;
    constexpr unsigned TAB_BYTES = TG * PER_ROW * 8;
    auto row_of = [&](int g) { return FWD ? g * TG : nrows - TG * (g + 1); };
    auto load = [&](int g) {
        const unsigned sl = slot_of(it + g);
        tma_load_group(tm, sdata[sl], scoef[sl], tab + (size_t)g * TG * PER_ROW, TAB_BYTES, sys0, row_of(g),
                       smem_addr(&full[sl]));
    };
    for (int g = 0; g < TAHEAD && g < ngroups; ++g) load(g);
    for (int g = 0; g < ngroups; ++g)
    {
        const unsigned sl = slot_of(it + g);
        mbar_wait_parity(smem_addr(&done[sl]), parity_of(it + g));
        tma_store_group(tm, sdata[sl], sys0, row_of(g));
        // group g + TAHEAD goes into the slot group g - TLAG left: at most the TLAG newest stores may still be reading
        if (g + TAHEAD < ngroups)
        {
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(TLAG) : "memory");
            load(g + TAHEAD);
        }
    }
    it += ngroups;
    // everything this sweep wrote must be in memory before the next sweep (or the next kernel) reads it
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// The tensor covers all nrows = m + 2 rows of the array (nrows % TG == 0 and nBatch % 32 == 0 are required, so every box
// is in bounds: the copy engine faults on a store that reaches outside the tensor, tools/tma_probe.cu).  The two rows
// below the reduced system (the last two unknowns of the cyclic system, handled by k_solve_end) pass through unchanged:
// tabF holds {0, 0, -1, 1} for them (x = rb / 1), tabB holds {0, 0} for them and for row m-1, {-du, 0} for row m-2.
// tabF: {-ds, -dl, -d, 1/d} per row (zeros where the recurrence has no term); tabB: {-du, -dw} per row in the order the
// backward sweep visits the groups: entry TG j + k = row nrows - TG (j+1) + k.
// (x - 0*y is exact; the one representable difference to skipping the term is that a -0.0 may come back as +0.0.)
__global__ void __launch_bounds__(64) k_pent_solve_tma(const __grid_constant__ CUtensorMap tm, const double* __restrict__ tabF,
                                                       const double* __restrict__ tabB, int nrows)
{
    extern __shared__ __align__(128) unsigned char tma_smem[];
    double (*sdata)[TG * 32] = reinterpret_cast<double (*)[TG * 32]>(tma_smem);
    double (*scoef)[TG * 4] = reinterpret_cast<double (*)[TG * 4]>(tma_smem + (size_t)TSLOTS * TG * 32 * sizeof(double));
    unsigned long long* full =
        reinterpret_cast<unsigned long long*>(tma_smem + (size_t)TSLOTS * (TG * 32 + TG * 4) * sizeof(double));
    unsigned long long* done = full + TSLOTS;
    double* dummy = reinterpret_cast<double*>(done + TSLOTS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sys0 = blockIdx.x * 32;
    const int ngroups = nrows / TG;
    if (threadIdx.x == 0)
    {
        for (int i = 0; i < TSLOTS; ++i)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full[i])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_addr(&done[i])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned it = 0;
    if (warp == 0)
    {
        const unsigned a_data = smem_addr(sdata), a_coef = smem_addr(scoef), a_full = smem_addr(full), a_done = smem_addr(done);
        const unsigned a_dummy = smem_addr(dummy);
        math_sweep<true>(a_data, a_coef, a_full, a_done, a_dummy, ngroups, lane, it);
        math_sweep<false>(a_data, a_coef, a_full, a_done, a_dummy, ngroups, lane, it);
    }
    else if (lane == 0)
    {
        copy_sweep<true>(&tm, tabF, sdata, scoef, full, done, nrows, ngroups, sys0, it);
        copy_sweep<false>(&tm, tabB, sdata, scoef, full, done, nrows, ngroups, sys0, it);
    }
}

// the two tables of k_pent_solve_tma, from the factors (one thread per table row; trows = m + 2 rounded up to TG)
__global__ void k_build_tables(const double* __restrict__ ds, const double* __restrict__ dl, const double* __restrict__ d,
                               const double* __restrict__ du, const double* __restrict__ dw, const double* __restrict__ rinv,
                               double* __restrict__ tabF, double* __restrict__ tabB, int m, int trows)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= trows) return;
    // forward, row i
    const bool in = i < m;
    tabF[4 * i + 0] = (in && i >= 2) ? -ds[i] : 0.0;
    tabF[4 * i + 1] = (in && i >= 1) ? -dl[i] : 0.0;
    tabF[4 * i + 2] = in ? -d[i] : -1.0;
    tabF[4 * i + 3] = in ? rinv[i] : 1.0;
    // backward, entry i = group j, position k -> row trows - TG (j+1) + k
    const int j = i / TG, k = i % TG;
    const int row = trows - TG * (j + 1) + k;
    tabB[2 * i + 0] = (row >= 0 && row <= m - 2) ? -du[row] : 0.0;
    tabB[2 * i + 1] = (row >= 0 && row <= m - 3) ? -dw[row] : 0.0;
}

// 2-D tensor map over the interleaved right-hand sides: dimension 0 = system (nBatch, contiguous), dimension 1 = row
// (all n rows); box = 32 systems x TG rows.  cuTensorMapEncodeTiled is a host-side encoder; it is reached through
// the runtime's driver entry point query, so libcuda is not a link-time dependency.
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_rhs_map(CUtensorMap* tm, double* data, int nBatch, int nrows)
{
    static TensorMapEncodeFn encode = nullptr;
    static bool looked = false;
    if (!looked)
    {
        looked = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = (TensorMapEncodeFn)fn;
        else
            cudaGetLastError();
    }
    if (!encode) return false;
    // every box must lie inside the tensor: whole groups of rows, whole warps of systems
    if (nBatch % 32 || nrows % TG || nrows < 2 * TG || ((uintptr_t)data & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)nBatch, (cuuint64_t)nrows};
    const cuuint64_t strides[1] = {(cuuint64_t)nBatch * sizeof(double)};
    const cuuint32_t box[2] = {32, (cuuint32_t)TG};
    const cuuint32_t estr[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void pent_tma_build_tables(const double* ds, const double* dl, const double* d, const double* du, const double* dw,
                           const double* rinv, double* tabF, double* tabB, int m, int trows)
{
    k_build_tables<<<(trows + 127) / 128, 128>>>(ds, dl, d, du, dw, rinv, tabF, tabB, m, trows);
}

bool pent_tma_solve(double* data, int nBatch, int n, const double* tabF, const double* tabB, cudaStream_t stream)
{
    CUtensorMap tm;
    if (!make_rhs_map(&tm, data, nBatch, n)) return false;
    // the opt-in to more than 48 KB of dynamic shared memory is per device
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev])
    {
        cudaFuncSetAttribute(k_pent_solve_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM);
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    k_pent_solve_tma<<<nBatch / 32, 64, TMA_SMEM, stream>>>(tm, tabF, tabB, n);
    return true;
}

}  // namespace custen_cahn
