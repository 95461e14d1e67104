/* cuSten-B200 CPU oracle — TEST INFRASTRUCTURE, never part of the product path.
 *
 * A plain-C restatement of what the reference's 2D stencil kernels compute, one point at a time, with the
 * reference's tap order and a fused multiply-add per tap (the sm_100 build of the reference is a pure DFMA
 * chain from sum = 0.0).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.
 *
 * Follows, per variant (paths under /root/reference/cuSten/src/kernels/):
 *   Xp     2d_x_p_kernel.cu:107-172      Xnp     2d_x_np_kernel.cu:97-177 (right strip <- 0.0, left strip untouched)
 *   Yp     2d_y_p_kernel.cu:104-192      Ynp     2d_y_np_kernel.cu:106-251
 *   XYp    2d_xy_p_kernel.cu:129-528     XYnp    2d_xy_np_kernel.cu:136-960
 *   XnpFun 2d_x_np_fun_kernel.cu:128,148,169   YpFun 2d_y_p_fun_kernel.cu:136-138
 *   YnpFun 2d_y_np_fun_kernel.cu:140-238       XYpFun 2d_xy_p_fun_kernel.cu:521-526 (loc = top-left)
 * Cahn-Hilliard pieces (paths under /root/reference/cuPentCahnADI/src/): see the second half of the file.
 *
 * Pinning (the reference ships no golden vectors):
 *   - on the CPU, against the reference's own serial restatement of XYp / XYpFun and of the ADI solve,
 *     serialCahnADI.c:478-622 and :628-720, compiled from /root/reference into oracle/_ref/libserialcahn.so
 *     (tests/test_oracle_cpu.py; agreement to rounding, that file is built without FMA);
 *   - on the GPU box, bit for bit against the reference's CUDA kernels rebuilt for sm_100
 *     (oracle/_ref/libcusten_ref.so, tests/test_parity_gpu.py) and against tests/golden/*.npz, which were
 *     produced by that library (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

enum { DIR_X = 0, DIR_Y = 1, DIR_XY = 2 };
enum {
    FUN_NONE = 0,
    FUN_SECOND_DIFF_X = 1,
    FUN_WEIGHTED9_X = 2,
    FUN_WEIGHTED9_Y = 3,
    FUN_WEIGHTED3_Y = 4,
    FUN_WEIGHTED_XY = 5,
    FUN_CUBIC_XY = 6
};

/* ---- the user-function fixtures, as nvcc contracts them (custen_b200/csrc/builtin_funs.cuh) ---- */

static double f_second_diff_x(const double* d, const double* coe, int loc)
{
    return (d[loc - 1] - 2 * d[loc] + d[loc + 1]) * coe[0];
}
static double f_weighted9_x(const double* d, const double* coe, int loc)
{
    double acc = 0.0;
    for (int k = 0; k < 9; ++k) acc = fma(coe[k], d[loc - 4 + k], acc);
    return acc;
}
static double f_weighted_y(const double* d, const double* coe, int loc, int jump, int n)
{
    double acc = 0.0;
    const int half = (n - 1) / 2;
    for (int k = 0; k < n; ++k) acc = fma(coe[k], d[loc + (k - half) * jump], acc);
    return acc;
}
static double f_weighted_xy(const double* d, const double* coe, int loc, int jump, int nx, int ny)
{
    double acc = 0.0;
    int c = 0;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) acc = fma(coe[c++], d[loc + j * jump + i], acc);
    return acc;
}
static double f_cubic_xy(const double* d, const double* coe, int loc, int jump, int nx, int ny)
{
    double acc = 0.0;
    int c = 0;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i)
        {
            const double v = d[loc + j * jump + i];
            const double u = fma(v * v, v, -v); /* (v*v*v) - v with the last multiply fused into the subtract */
            acc = fma(coe[c++], u, acc);
        }
    return acc;
}

static int wrap(int i, int n)
{
    i %= n;
    return i < 0 ? i + n : i;
}

/* One sweep of any of the 12 variants over an ny x nx row-major grid.
 *   dir       DIR_X / DIR_Y / DIR_XY        fun  FUN_* (FUN_NONE = weights variant)
 *   periodic  bit 0: x wraps, bit 1: y wraps.  The reference's variants are 3 (periodic) or 0 (non-periodic);
 *             1 = "x wraps, y does not" is what one y-slab of a periodic grid looks like once its halo rows
 *             have been attached above and below (multi-GPU tests).
 *   H,L,R     window width, taps left, taps right (X: H = numSten; Y: H = 1, L = R = 0)
 *   V,T,B     window height, taps above, taps below (X: V = 1, T = B = 0)
 * `out` must be pre-filled by the caller: regions the reference does not write are left alone.
 * Returns 0, or -1 for an unsupported combination. */
int custen_oracle_sweep(int dir, int periodic, int fun, const double* in, double* out, int nx, int ny,
                        const double* coef, int H, int L, int R, int V, int T, int B)
{
    if (dir == DIR_X) { V = 1; T = B = 0; }
    if (dir == DIR_Y) { H = 1; L = R = 0; }
    const int Reff = (H - 1 - L) > R ? (H - 1 - L) : R;
    const int Beff = (V - 1 - T) > B ? (V - 1 - T) : B;
    const int WW = L + Reff + 1, WH = T + Beff + 1;
    double* win = (double*)malloc(sizeof(double) * (size_t)WW * WH);
    if (!win) return -1;

    const int px = periodic & 1, py = (periodic >> 1) & 1;
    int xlo = 0, xhi = nx, ylo = 0, yhi = ny;
    if (!px && dir != DIR_Y) { xlo = L; xhi = nx - R; }
    if (!py && dir != DIR_X) { ylo = T; yhi = ny - B; }
    const int zero_right = (!px && dir == DIR_X && fun == FUN_NONE);

    int rc = 0;
    for (int y = ylo; y < yhi && rc == 0; ++y)
    {
        for (int x = 0; x < nx; ++x)
        {
            if (x < xlo) continue;
            if (x >= xhi)
            {
                if (zero_right) out[(size_t)y * nx + x] = 0.0;
                continue;
            }
            /* gather the window: row j, column i <- in[y-T+j, x-L+i] through the index map */
            for (int j = 0; j < WH; ++j)
                for (int i = 0; i < WW; ++i)
                {
                    int gy = y - T + j, gx = x - L + i;
                    double v = 0.0;
                    if (px) gx = wrap(gx, nx);
                    if (py) gy = wrap(gy, ny);
                    if (gy >= 0 && gy < ny && gx >= 0 && gx < nx) v = in[(size_t)gy * nx + gx];
                    win[j * WW + i] = v;
                }
            double r = 0.0;
            switch (fun)
            {
                case FUN_NONE:
                    for (int j = 0; j < V; ++j)
                        for (int i = 0; i < H; ++i) r = fma(coef[j * H + i], win[j * WW + i], r);
                    break;
                case FUN_SECOND_DIFF_X: r = f_second_diff_x(win, coef, L); break;
                case FUN_WEIGHTED9_X: r = f_weighted9_x(win, coef, L); break;
                case FUN_WEIGHTED9_Y: r = f_weighted_y(win, coef, T * WW, WW, 9); break;
                case FUN_WEIGHTED3_Y: r = f_weighted_y(win, coef, T * WW, WW, 3); break;
                case FUN_WEIGHTED_XY: r = f_weighted_xy(win, coef, 0, WW, H, V); break;
                case FUN_CUBIC_XY: r = f_cubic_xy(win, coef, 0, WW, H, V); break;
                default: rc = -1; break;
            }
            out[(size_t)y * nx + x] = r;
        }
    }
    free(win);
    return rc;
}
