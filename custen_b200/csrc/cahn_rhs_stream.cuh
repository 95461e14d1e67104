// Cahn-Hilliard ADI: the right-hand side of a time step as a row-streaming kernel (tolerance-mode road, cahn_part.cu).
#ifndef CUSTEN_B200_CAHN_RHS_STREAM_CUH
#define CUSTEN_B200_CAHN_RHS_STREAM_CUH

#include "cahn_rhs.cuh"

namespace custen_cahn {

// ---- the same right-hand side as a row-streaming kernel (tolerance-mode road) -----------------------------------------------
// k_rhs_fused spends its time on instruction issue and shared-memory traffic (ncu: 80 M warp instructions for 16.8 M
// points, 67 % of the shared-memory wavefront peak, FP64 pipe 40 %).  Here a CTA owns a strip of 512 columns x 32 rows and
// streams the 36 rows it needs through a ring in shared memory: rows of c and cOld are moved in with the async proxy
// (bulk copies on mbarriers issued by warp 0 a few stages ahead, the periodic wrap / the neighbouring slabs' halo rows
// resolved per row), and each of the 128 threads owns 4 columns and marches down the rows with its 5-row window of cBar and 3-row window
// of (c^3 - c) in registers - nothing is staged twice, nothing is exchanged between threads.  The zero weights of the
// two stencils (13 of 25 and 5 of 9 taps are non-zero: cuPentCahnADI.cu:452-476, :164-188) are skipped, which leaves
// every partial sum as it was: fma(0, v, acc) == acc for finite v.  Per point: 27 FP64 operations, 2 shared-memory
// loads of 16 bytes, 24 bytes of HBM traffic.
constexpr int RS_NT = 128;          // consumer threads
constexpr int RS_W = 4 * RS_NT;     // strip width
constexpr int RS_PW = RS_W + 4;     // row pitch in shared memory: 2 halo columns either side
constexpr int RS_BR = 32;           // output rows per CTA
constexpr int RS_SR = 4;            // rows per stage
constexpr int RS_NS = 3;            // stages in the ring
constexpr int RS_ROWS = RS_BR + 4;  // rows streamed per CTA
constexpr size_t RS_STAGE_DOUBLES = (size_t)RS_SR * 2 * RS_PW;
constexpr size_t RS_SMEM = RS_NS * RS_STAGE_DOUBLES * sizeof(double) + 2 * RS_NS * sizeof(unsigned long long);

namespace rs {
__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_wait(unsigned a, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra RS_WAIT_%=;\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void g2s(unsigned dst, const double* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
}  // namespace rs

__global__ void __launch_bounds__(RS_NT, 2) k_rhs_stream(const double* __restrict__ cOld, const double* __restrict__ cCurr,
                                                             const RhsHalo halo, double* __restrict__ out, int n, int rows,
                                                             const RhsCoef k)
{
    extern __shared__ __align__(128) unsigned char rs_smem[];
    double* ring = reinterpret_cast<double*>(rs_smem);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring + RS_NS * RS_STAGE_DOUBLES);   // full[NS], empty[NS]
    const int tid = threadIdx.x;
    const int xs = blockIdx.x * RS_W, y0 = blockIdx.y * RS_BR;
    const int ws = min(RS_W, n - xs);   // this strip's width (a multiple of 4)
    if (tid == 0)
    {
        for (int s = 0; s < RS_NS; ++s)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rs::saddr(bars + s)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rs::saddr(bars + RS_NS + s)), "r"(RS_NT) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- filling the ring: warp 0, between its own rows; lane = (row of the stage, array, piece) ----
    // (a fifth, dedicated producer warp would cost a whole CTA of occupancy: registers are handed out four warps at a time)
    const int lane = tid & 31, warp = tid >> 5;
    auto fill = [&](int f) {
        const int slot = f % RS_NS;
        const unsigned full = rs::saddr(bars + slot), empty = rs::saddr(bars + RS_NS + slot);
        if (f >= RS_NS) rs::bar_wait(empty, (unsigned)((f / RS_NS - 1) & 1));
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((unsigned)RS_SR * 2u * (unsigned)(ws + 4) * 8u)
                         : "memory");
        __syncwarp();
        if (lane < RS_SR * 6)
        {
            const int r = lane / 6, which = lane % 6, arr = which / 3, piece = which % 3;   // piece 0 body, 1 left halo, 2 right halo
            int gy = y0 - 2 + f * RS_SR + r;
            const double* base = arr ? cOld : cCurr;
            if (gy < 0 || gy >= rows)
            {
                if (halo.c_up != nullptr)
                {
                    const bool up = gy < 0;
                    base = up ? (arr ? halo.o_up : halo.c_up) : (arr ? halo.o_down : halo.c_down);
                    gy = up ? gy + 2 : gy - rows;
                }
                else
                    gy = gy < 0 ? gy + rows : gy - rows;
            }
            const double* grow = base + (size_t)gy * n;
            double* srow = ring + slot * RS_STAGE_DOUBLES + (size_t)(r * 2 + arr) * RS_PW;
            if (piece == 0)
                rs::g2s(rs::saddr(srow + 2), grow + xs, (unsigned)ws * 8u, full);
            else if (piece == 1)
                rs::g2s(rs::saddr(srow), grow + (xs == 0 ? n - 2 : xs - 2), 16u, full);
            else
                rs::g2s(rs::saddr(srow + 2 + ws), grow + (xs + ws >= n ? 0 : xs + ws), 16u, full);
        }
        __syncwarp();
    };
    constexpr int NFILL = RS_ROWS / RS_SR;
    if (warp == 0)
        for (int f = 0; f < RS_NS && f < NFILL; ++f) fill(f);

    // ---- consumers ----
    const int x0 = 4 * tid;
    const bool active = x0 < ws;
    // row kk of the stream is grid row y0 - 2 + kk; output row y0 + j needs cBar rows j .. j + 4, (c^3 - c) rows j + 1 ..
    // j + 3 and c - cOld of row j + 2, so it is computed when row kk = j + 4 has arrived
    double B[5][8] = {};   // cBar rows kk - 4 .. kk, columns x0 - 2 .. x0 + 5
    double F[3][6] = {};   // c^3 - c rows kk - 3 .. kk - 1 (at the top of iteration kk), columns x0 - 1 .. x0 + 4
    double D[3][4] = {};   // c - cOld of the own columns, same rows as F
#pragma unroll
    for (int kk = 0; kk < RS_ROWS; ++kk)
    {
        const int f = kk / RS_SR, slot = f % RS_NS;
        if (kk % RS_SR == 0)
        {
            // the slot of the stage this warp has just left is refilled with the stage NS - 1 ahead of the current one
            if (warp == 0 && f >= 1 && f - 1 + RS_NS < NFILL) fill(f - 1 + RS_NS);
            rs::bar_wait(rs::saddr(bars + slot), (unsigned)((f / RS_NS) & 1));
        }
        if (active)
        {
            const double* srow = ring + slot * RS_STAGE_DOUBLES + (size_t)((kk % RS_SR) * 2) * RS_PW + x0;
            double c[8], o[8];
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const double2 vc = *reinterpret_cast<const double2*>(srow + 2 * j);
                const double2 vo = *reinterpret_cast<const double2*>(srow + RS_PW + 2 * j);
                c[2 * j] = vc.x; c[2 * j + 1] = vc.y;
                o[2 * j] = vo.x; o[2 * j + 1] = vo.y;
            }
            // slide the windows (the row loop is unrolled: these are renames, not moves)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int i = 0; i < 8; ++i) B[a][i] = B[a + 1][i];
#pragma unroll
            for (int i = 0; i < 8; ++i) B[4][i] = 2.0 * c[i] - o[i];
            if (kk >= 4)
            {
                // output row y0 + kk - 4
                double res[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                {
                    const int i = q + 2;   // own column in B
                    double lin = 0.0;
                    lin = fma(k.wl[2], B[0][i], lin);
                    lin = fma(k.wl[6], B[1][i - 1], lin);
                    lin = fma(k.wl[7], B[1][i], lin);
                    lin = fma(k.wl[8], B[1][i + 1], lin);
                    lin = fma(k.wl[10], B[2][i - 2], lin);
                    lin = fma(k.wl[11], B[2][i - 1], lin);
                    lin = fma(k.wl[12], B[2][i], lin);
                    lin = fma(k.wl[13], B[2][i + 1], lin);
                    lin = fma(k.wl[14], B[2][i + 2], lin);
                    lin = fma(k.wl[16], B[3][i - 1], lin);
                    lin = fma(k.wl[17], B[3][i], lin);
                    lin = fma(k.wl[18], B[3][i + 1], lin);
                    lin = fma(k.wl[22], B[4][i], lin);
                    const int e = q + 1;   // own column in F
                    double non = 0.0;
                    non += k.cn[1] * F[0][e];
                    non += k.cn[3] * F[1][e - 1];
                    non += k.cn[4] * F[1][e];
                    non += k.cn[5] * F[1][e + 1];
                    non += k.cn[7] * F[2][e];
                    double h = lin;
                    h += -(2.0 / 3.0) * D[1][q] + non;
                    res[q] = h;
                }
                double* orow = out + (size_t)(y0 + kk - 4) * n + xs + x0;
                *reinterpret_cast<double2*>(orow) = make_double2(res[0], res[1]);
                *reinterpret_cast<double2*>(orow + 2) = make_double2(res[2], res[3]);
            }
            // row kk enters the two lagging windows
#pragma unroll
            for (int a = 0; a < 2; ++a)
            {
#pragma unroll
                for (int i = 0; i < 6; ++i) F[a][i] = F[a + 1][i];
#pragma unroll
                for (int i = 0; i < 4; ++i) D[a][i] = D[a + 1][i];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i)
            {
                const double u = c[i + 1];
                F[2][i] = (u * u * u) - u;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) D[2][i] = c[i + 2] - o[i + 2];
        }
        if (kk % RS_SR == RS_SR - 1)
        {
            // every thread hands the stage back itself (compute-sanitizer's racecheck follows a thread's own arrival, not
            // "__syncwarp, then lane 0 arrives for the warp": with that form it reports the refill as racing with the other
            // lanes' reads)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rs::saddr(bars + RS_NS + slot)) : "memory");
        }
    }
}

}  // namespace custen_cahn

#endif
