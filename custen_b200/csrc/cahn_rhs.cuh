// Cahn-Hilliard ADI: the right-hand side of a time step in one pass, shared by the bit-identical road (cahn.cu, writes
// rhs^T for the interleaved solve) and the tolerance-mode road (cahn_part.cu, writes rhs in the grid's own layout, on a
// y-slab whose halo rows may lie in the neighbouring GPUs' memory).
#ifndef CUSTEN_B200_CAHN_RHS_CUH
#define CUSTEN_B200_CAHN_RHS_CUH

#include <cstddef>

namespace custen_cahn {

// ---- the whole right-hand side in one pass ---------------------------------------------------
// findCBar + both stencils + findRHS + transpose: reads c and cOld once (with a 2-point periodic halo), writes rhs^T.
// Per point the arithmetic is that of the separate passes, operation for operation, so the result has the same bits:
//   cBar = 2 c - cOld                                                    (k_cbar; cuPentCahnADI.cu:58-69)
//   lin  = sum_{j,i} wl[5j+i] * cBar(y-2+j, x-2+i), one fma chain from 0.0, j outer, i inner
//                                                                        (stream_acc_kernel; 2d_xy_p_kernel.cu:507-520)
//   non  = cubic_xy(c tile, coeN, top-left of the 3 x 3 window)          (the registered user function, builtin_funs.cuh)
//   rhs  = lin + (-(2/3)(c - cOld) + non)                                (k_rhs_transpose; cuPentCahnADI.cu:72-86)
// A CTA owns a 32 x 32 tile; a thread owns a 2-column x 4-row patch, so its windows slide through registers and every
// shared-memory read is a 128-bit load (8 rows x 3 loads for the eight 5 x 5 windows, 6 rows x 2 loads for the eight
// 3 x 3 ones): shared-memory bandwidth stays below the FP64 pipe's time.
// The 25 + 9 coefficients travel as kernel arguments: FP64 instructions read them straight from the constant bank.
struct RhsCoef
{
    double wl[25];
    double cn[9];
};
constexpr int FT = 32;            // tile edge
constexpr int FP = FT + 4;        // cBar tile: halo of 2 on each side (even pitch: 16-byte aligned pairs)
constexpr int FPC = FT + 6;       // c tile: stored one column to the right so that the 3 x 3 windows' pairs are aligned too
// The two rows above a y-slab's first row and below its last one, for c and cOld (row -2 at *_up, row -1 at *_up + n; row
// `rows` at *_down, `rows + 1` at *_down + n).  All null: the array is the whole periodic grid.
struct RhsHalo
{
    const double *c_up, *c_down, *o_up, *o_down;
};
// TRANSPOSED: out[x * rows + y] (n == rows, the whole grid); otherwise out[y * n + x] on a rows x n slab (rows % 32 == 0
// and n % 32 == 0 on that road).
template <bool TRANSPOSED>
__global__ void __launch_bounds__(128, 6) k_rhs_fused(const double* __restrict__ cOld, const double* __restrict__ cCurr,
                                                      const RhsHalo halo, double* __restrict__ outT, int n, int rows,
                                                      const RhsCoef k)
{
    __shared__ __align__(16) double sc[FP * FPC];   // c      (row r, column col at r * FPC + col + 1)
    __shared__ __align__(16) double sb[FP * FP];    // cBar   (row r, column col at r * FP + col)
    __shared__ double tile[FT][FT + 1];             // cOld of the tile's points, then their rhs, for the transposed store
    const int tid = threadIdx.x;
    const int bx = blockIdx.x * FT, by = blockIdx.y * FT;
    // all of a thread's loads are issued before the first one is used (the loop is unrolled and split in two passes):
    // with six CTAs per SM the tile's load latency has to be paid once, not once per element
    constexpr int NLD = (FP * FP + 127) / 128;
    double vc[NLD], vo[NLD];
#pragma unroll
    for (int it = 0; it < NLD; ++it)
    {
        const int e = tid + it * 128;
        vc[it] = vo[it] = 0.0;
        if (e < FP * FP)
        {
            const int r = e / FP, col = e - r * FP;
            int gy = by - 2 + r, gx = bx - 2 + col;
            gx = gx < 0 ? gx + n : (gx >= n ? gx - n : gx);
            if (gx >= n) gx -= n;   // ragged last tile: rows / columns past the edge are loaded (wrapped), never written
            if (!TRANSPOSED && halo.c_up != nullptr && (gy < 0 || gy >= rows))
            {
                const bool up = gy < 0;
                const size_t i = (size_t)(up ? gy + 2 : gy - rows) * n + gx;
                vc[it] = (up ? halo.c_up : halo.c_down)[i];
                vo[it] = (up ? halo.o_up : halo.o_down)[i];
            }
            else
            {
                gy = gy < 0 ? gy + rows : (gy >= rows ? gy - rows : gy);
                if (gy >= rows) gy -= rows;
                const size_t i = (size_t)gy * n + gx;
                vc[it] = cCurr[i];
                vo[it] = cOld[i];
            }
        }
    }
#pragma unroll
    for (int it = 0; it < NLD; ++it)
    {
        const int e = tid + it * 128;
        if (e < FP * FP)
        {
            const int r = e / FP, col = e - r * FP;
            const double c = vc[it], co = vo[it];
            sc[r * FPC + col + 1] = c;
            sb[e] = 2.0 * c - co;
            // cOld of the tile's own points waits in the transpose buffer until its owner turns it into the rhs
            if (r >= 2 && r < FT + 2 && col >= 2 && col < FT + 2) tile[r - 2][col - 2] = co;
        }
    }
    __syncthreads();

    // outputs (r0 + o, x0 + q), o < 4, q < 2; output (r, x) sits at tile coordinates (r + 2, x + 2)
    const int x0 = 2 * (tid & 15), r0 = 4 * (tid >> 4);
    double lin[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int jr = 0; jr < 8; ++jr)   // cBar tile row r0 + jr is tap row j = jr - o of output row o
    {
        const double2* p = reinterpret_cast<const double2*>(sb + (r0 + jr) * FP + x0);
        const double2 a0 = p[0], a1 = p[1], a2 = p[2];
        const double v[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};   // tile columns x0 .. x0 + 5
#pragma unroll
        for (int o = 0; o < 4; ++o)
        {
            const int j = jr - o;
            if (j >= 0 && j < 5)
            {
#pragma unroll
                for (int i = 0; i < 5; ++i)
                {
                    lin[o][0] = fma(k.wl[j * 5 + i], v[i], lin[o][0]);
                    lin[o][1] = fma(k.wl[j * 5 + i], v[i + 1], lin[o][1]);
                }
            }
        }
    }
    // the user function of the nonlinear term, custen_funs::cubic_xy (cuPentCahnADI.cu:164-188), on the same windows:
    // acc = 0; for j: for i: acc += coe[3j + i] * ((v * v * v) - v)
    double non[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int jr = 0; jr < 6; ++jr)   // c tile row r0 + 1 + jr is tap row j = jr - o of output row o
    {
        const double2* p = reinterpret_cast<const double2*>(sc + (r0 + 1 + jr) * FPC + x0 + 2);
        const double2 a0 = p[0], a1 = p[1];
        const double u[4] = {a0.x, a0.y, a1.x, a1.y};               // tile columns x0 + 1 .. x0 + 4
        double t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = (u[i] * u[i] * u[i]) - u[i];
#pragma unroll
        for (int o = 0; o < 4; ++o)
        {
            const int j = jr - o;
            if (j >= 0 && j < 3)
            {
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
                    non[o][0] += k.cn[j * 3 + i] * t[i];
                    non[o][1] += k.cn[j * 3 + i] * t[i + 1];
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int q = 0; q < 2; ++q)
        {
            const int r = r0 + o, x = x0 + q;
            const double c = sc[(r + 2) * FPC + x + 3];
            const double co = tile[r][x];
            double h = lin[o][q];
            h += -(2.0 / 3.0) * (c - co) + non[o][q];
            tile[r][x] = h;
        }
    __syncthreads();
    const int tx = tid & 31;
    for (int r = tid >> 5; r < FT; r += 4)
    {
        if (TRANSPOSED)
        {
            const int x = by + tx, y = bx + r;   // row y of outT is column bx + r of the grid
            if (x < rows && y < n) outT[(size_t)y * rows + x] = tile[tx][r];
        }
        else
        {
            const int x = bx + tx, y = by + r;
            if (x < n && y < rows) outT[(size_t)y * n + x] = tile[r][tx];
        }
    }
}

}  // namespace custen_cahn

#endif
