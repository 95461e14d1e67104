// The batched pentadiagonal solve of the Cahn-Hilliard ADI step, fed by the TMA engine (see the design note below).
//
// This file is its own translation unit because it is compiled with `-Xptxas -O1` (Makefile): the instruction ORDER
// written here is the optimisation.  The math warp is alone on its scheduler and issues in order, so the shared-memory
// traffic of a row has to sit in the idle issue slots between the dependent FP64 operations of the recurrence; at its
// default level ptxas schedules for latency hiding by other warps and moves all of it behind (or in front of) the
// chain, where it costs ~13 cycles per row (ncu source view).  At -O1 it keeps the order of the PTX.
#include "pent_solve.h"

#include <cuda.h>

#include <cstdint>
#include <type_traits>

namespace custen_cahn {

// ---- the same solve fed by the TMA engine -----------------------------------------------------------------------------
// A warp that is alone on its scheduler issues in order, so every instruction that is not one of the recurrence's
// dependent FP64 operations (8 cycles each, tools/fp64_latency.cu: 4 per row forward, 2 backward) has to fit in the gaps
// between them.  The cp.async version above spends ~17 instructions per row (per-lane 64-bit address arithmetic for every
// LDGSTS and STG) and runs at ~57 cycles per row and direction.  Here a CTA is two warps on two schedulers:
//   * the math warp touches shared memory only: per group of TG rows it waits for the group's slot, pulls right-hand
//     sides and coefficients into registers (one group ahead of the arithmetic), runs the chain, writes the results back
//     into the slot and arrives on the slot's `done` barrier.  Per row: LDS rhs, 2 x LDS.128 coefficients, the chain, STS;
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
constexpr size_t TMA_SMEM = (size_t)TSLOTS * (TG * 32 + TG * 4) * sizeof(double) + 2 * TSLOTS * sizeof(unsigned long long);

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

// Shared-memory accesses of the math warp, as PTX so that their order and their addresses are under control.
// The warp issues in order and ptxas schedules for latency hiding by other warps, which do not exist here: left alone it
// issues a group's ~50 shared-memory loads in one clump during which the chain stands still (ncu: 13 cycles per row).
// The loads of the NEXT group's row k are therefore given a (fake) dependence on the CURRENT group's result x_k -
// address = base + lo32(x_k) * zero, with `zero` a kernel argument that is 0 at run time but unknown to the compiler -
// so that they can only be placed behind row k's chain operations, i.e. in the idle issue slots of row k+1's.
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
__device__ __forceinline__ unsigned after(unsigned addr, double x, unsigned zero)
{
    unsigned a;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"((unsigned)__double2loint(x)), "r"(zero), "r"(addr));
    return a;
}

// Registers of the math warp: a rolling window of TW rows (right-hand sides and coefficients).  Row K of a group sits
// in window position K % TW (forward; backward: (TG - 1 - K) % TW); once solved, its registers take the row TW
// positions further along the sweep, from the current slot or, near the end of the group, from the next one.
struct FwdWin
{
    double2 c0[TW], c1[TW];  // {ds, dl}, {d, 1/d}
    double r[TW];
};
struct BwdWin
{
    double2 c[TW];           // {du, dw}
    double r[TW];
};
struct Addr
{
    unsigned rb_cur, rb_nxt;   // this lane's column in the current / next slot's rows
    unsigned cf_cur, cf_nxt;   // coefficient rows of the current / next slot
};

template <int K>
__device__ __forceinline__ void fwd_row(FwdWin& q, const Addr& a, unsigned zero, double& p1, double& p2)
{
    constexpr int S = K % TW;
    // cuPentBatch.cu:153-172
    const double x = div_by(q.r[S] - q.c0[S].x * p2 - q.c0[S].y * p1, q.c1[S].x, q.c1[S].y);
    sts_f64<K * 256>(a.rb_cur, x);
    p2 = p1;
    p1 = x;
    constexpr bool SAME = K + TW < TG;
    constexpr int N = SAME ? K + TW : K + TW - TG;   // row fetched next, in the current or the next slot
    const unsigned rb = after(SAME ? a.rb_cur : a.rb_nxt, x, zero), cf = after(SAME ? a.cf_cur : a.cf_nxt, x, zero);
    q.r[S] = lds_f64<N * 256>(rb);
    q.c0[S] = lds_v2f64<N * 32>(cf);
    q.c1[S] = lds_v2f64<N * 32 + 16>(cf);
}
template <int K>
__device__ __forceinline__ void bwd_row(BwdWin& q, const Addr& a, unsigned zero, double& p1, double& p2)
{
    constexpr int S = (TG - 1 - K) % TW;
    // cuPentBatch.cu:176-195; p1 = x[i+1], p2 = x[i+2]
    const double x = q.r[S] - q.c[S].x * p1 - q.c[S].y * p2;
    sts_f64<K * 256>(a.rb_cur, x);
    p2 = p1;
    p1 = x;
    constexpr bool SAME = K - TW >= 0;
    constexpr int N = SAME ? K - TW : K - TW + TG;
    const unsigned rb = after(SAME ? a.rb_cur : a.rb_nxt, x, zero), cf = after(SAME ? a.cf_cur : a.cf_nxt, x, zero);
    q.r[S] = lds_f64<N * 256>(rb);
    q.c[S] = lds_v2f64<N * 16>(cf);
}
// rows LO .. HI of a group, unrolled at compile time: forward ascending, backward descending
template <int LO, int HI>
struct Rows
{
    static __device__ __forceinline__ void fwd(FwdWin& q, const Addr& a, unsigned z, double& p1, double& p2)
    {
        fwd_row<LO>(q, a, z, p1, p2);
        if constexpr (LO < HI) Rows<LO + 1, HI>::fwd(q, a, z, p1, p2);
    }
    static __device__ __forceinline__ void bwd(BwdWin& q, const Addr& a, unsigned z, double& p1, double& p2)
    {
        bwd_row<HI>(q, a, z, p1, p2);
        if constexpr (LO < HI) Rows<LO, HI - 1>::bwd(q, a, z, p1, p2);
    }
};
// the first TW rows of a sweep
__device__ __forceinline__ void window_first_load(FwdWin& q, unsigned rb, unsigned cf)
{
#pragma unroll
    for (int k = 0; k < TW; ++k)
    {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(q.r[k]) : "r"(rb + k * 256));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q.c0[k].x), "=d"(q.c0[k].y) : "r"(cf + k * 32));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q.c1[k].x), "=d"(q.c1[k].y) : "r"(cf + k * 32 + 16));
    }
}
__device__ __forceinline__ void window_first_load(BwdWin& q, unsigned rb, unsigned cf)
{
#pragma unroll
    for (int k = 0; k < TW; ++k)   // window position k = row TG - 1 - k
    {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(q.r[k]) : "r"(rb + (TG - 1 - k) * 256));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q.c[k].x), "=d"(q.c[k].y) : "r"(cf + (TG - 1 - k) * 16));
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

// Math warp, one sweep over the ngroups = nrows / TG row groups.  FWD: top to bottom, 4 coefficients per row; !FWD:
// bottom to top, 2 per row.  Slot indices and barrier parities advance incrementally; the barrier of the group after
// next is polled in the middle of a group so that the poll's latency is not paid between two groups.
template <bool FWD>
__device__ __forceinline__ void math_sweep(unsigned sdata, unsigned scoef, unsigned full, unsigned done, int ngroups, int lane,
                                           unsigned zero, unsigned& it)
{
    typedef typename std::conditional<FWD, FwdWin, BwdWin>::type Window;
    const unsigned b_rb = pinned(sdata + lane * 8), b_cf = pinned(scoef), b_full = pinned(full), b_done = pinned(done);
    double p1 = 0.0, p2 = 0.0;
    Window q;
    unsigned s_cur = slot_of(it), ph_cur = parity_of(it);
    mbar_wait_parity(b_full + 8 * s_cur, ph_cur);
    window_first_load(q, b_rb + s_cur * (TG * 256), b_cf + s_cur * (TG * 32));
    unsigned s_nxt = s_cur + 1 == TSLOTS ? 0 : s_cur + 1;
    unsigned ph_nxt = s_cur + 1 == TSLOTS ? ph_cur ^ 1 : ph_cur;
    unsigned landed = ngroups > 1 ? mbar_test_parity(b_full + 8 * s_nxt, ph_nxt) : 1u;
    for (int g = 0; g < ngroups; ++g)
    {
        // the next group's copy must have landed: its rows are fetched underneath this group's chain.  (After the last
        // group the fetches re-read the current slot; the values are not used.)
        const bool has_next = g + 1 < ngroups;
        if (has_next && !landed) mbar_wait_parity(b_full + 8 * s_nxt, ph_nxt);
        const unsigned s_src = has_next ? s_nxt : s_cur;
        Addr a;
        a.rb_cur = b_rb + s_cur * (TG * 256);
        a.rb_nxt = b_rb + s_src * (TG * 256);
        a.cf_cur = b_cf + s_cur * (TG * 32);
        a.cf_nxt = b_cf + s_src * (TG * 32);
        const unsigned s_nn = s_nxt + 1 == TSLOTS ? 0 : s_nxt + 1;
        const unsigned ph_nn = s_nxt + 1 == TSLOTS ? ph_nxt ^ 1 : ph_nxt;
        if constexpr (FWD)
        {
            Rows<0, TG / 2 - 1>::fwd(q, a, zero, p1, p2);
            landed = mbar_test_parity(b_full + 8 * s_nn, ph_nn);
            Rows<TG / 2, TG - 1>::fwd(q, a, zero, p1, p2);
        }
        else
        {
            Rows<TG / 2, TG - 1>::bwd(q, a, zero, p1, p2);
            landed = mbar_test_parity(b_full + 8 * s_nn, ph_nn);
            Rows<0, TG / 2 - 1>::bwd(q, a, zero, p1, p2);
        }
        // results -> async proxy (the tensor store reads them), then tell the copy warp
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b_done + 8 * s_cur) : "memory");
        s_cur = s_nxt;
        ph_cur = ph_nxt;
        s_nxt = s_nn;
        ph_nxt = ph_nn;
    }
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
// tabF holds {0, 0, 1, 1} for them (x = rb / 1), tabB holds {0, 0} for them and for row m-1, {du, 0} for row m-2.
// tabF: {ds, dl, d, 1/d} per row (zeros where the recurrence has no term); tabB: {du, dw} per row in the order the
// backward sweep visits the groups: entry TG j + k = row nrows - TG (j+1) + k.
// (x - 0*y is exact; the one representable difference to skipping the term is that a -0.0 may come back as +0.0.)
__global__ void __launch_bounds__(64) k_pent_solve_tma(const __grid_constant__ CUtensorMap tm, const double* __restrict__ tabF,
                                                       const double* __restrict__ tabB, int nrows, unsigned zero)
{
    extern __shared__ __align__(128) unsigned char tma_smem[];
    double (*sdata)[TG * 32] = reinterpret_cast<double (*)[TG * 32]>(tma_smem);
    double (*scoef)[TG * 4] = reinterpret_cast<double (*)[TG * 4]>(tma_smem + (size_t)TSLOTS * TG * 32 * sizeof(double));
    unsigned long long* full =
        reinterpret_cast<unsigned long long*>(tma_smem + (size_t)TSLOTS * (TG * 32 + TG * 4) * sizeof(double));
    unsigned long long* done = full + TSLOTS;
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
        math_sweep<true>(a_data, a_coef, a_full, a_done, ngroups, lane, zero, it);
        math_sweep<false>(a_data, a_coef, a_full, a_done, ngroups, lane, zero, it);
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
    tabF[4 * i + 0] = (in && i >= 2) ? ds[i] : 0.0;
    tabF[4 * i + 1] = (in && i >= 1) ? dl[i] : 0.0;
    tabF[4 * i + 2] = in ? d[i] : 1.0;
    tabF[4 * i + 3] = in ? rinv[i] : 1.0;
    // backward, entry i = group j, position k -> row trows - TG (j+1) + k
    const int j = i / TG, k = i % TG;
    const int row = trows - TG * (j + 1) + k;
    tabB[2 * i + 0] = (row >= 0 && row <= m - 2) ? du[row] : 0.0;
    tabB[2 * i + 1] = (row >= 0 && row <= m - 3) ? dw[row] : 0.0;
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

bool pent_tma_solve(double* data, int nBatch, int n, const double* tabF, const double* tabB)
{
    CUtensorMap tm;
    if (!make_rhs_map(&tm, data, nBatch, n)) return false;
    static bool configured = false;
    if (!configured)
    {
        cudaFuncSetAttribute(k_pent_solve_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM);
        configured = true;
    }
    k_pent_solve_tma<<<nBatch / 32, 64, TMA_SMEM>>>(tm, tabF, tabB, n, 0u);
    return true;
}

}  // namespace custen_cahn
