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


/* ---- WENO5 advection (13th variant): u dphi/dx + v dphi/dy, periodic --------------------------------------------
 * Follows 2d_xyADVWENO_p_kernel.cu:51-86 (wenoSten) and :283-390 (upwinded one-sided differences).  The reference
 * squares through the single-precision powf; CUDA's powf is not correctly rounded and is not reproducible on a CPU,
 * so this restatement (libm powf) agrees with the GPU kernels only to single-precision rounding of the smoothness
 * terms: it is a sanity oracle for the WENO variant (tolerance in tests/test_weno_gpu.py); bit-level parity for that
 * variant is against the reference's own CUDA kernel. */
static double weno5_cpu(double v1, double v2, double v3, double v4, double v5)
{
    const double epsilon = 1e-06;
    const double phi1 = (1.0 / 3.0) * v1 - (7.0 / 6.0) * v2 + (11.0 / 6.0) * v3;
    const double phi2 = -(1.0 / 6.0) * v2 + (5.0 / 6.0) * v3 + (1.0 / 3.0) * v4;
    const double phi3 = (1.0 / 3.0) * v3 + (5.0 / 6.0) * v4 - (1.0 / 6.0) * v5;
    const double s1 = (13.0 / 12.0) * powf(v1 - 2.0 * v2 + v3, 2.0) + 0.25 * powf(v1 - 4.0 * v2 + 3.0 * v3, 2.0);
    const double s2 = (13.0 / 12.0) * powf(v2 - 2.0 * v3 + v4, 2.0) + 0.25 * powf(v2 - v4, 2.0);
    const double s3 = (13.0 / 12.0) * powf(v3 - 2.0 * v4 + v5, 2.0) + 0.25 * powf(3.0 * v3 - 4.0 * v4 + v5, 2.0);
    const double a1 = 0.1 / powf(s1 + epsilon, 2.0);
    const double a2 = 0.6 / powf(s2 + epsilon, 2.0);
    const double a3 = 0.3 / powf(s3 + epsilon, 2.0);
    const double denom = 1.0 / (a1 + a2 + a3);
    return phi1 * (a1 * denom) + phi2 * (a2 * denom) + phi3 * (a3 * denom);
}

static double weno_line_cpu(const double* in, int nx, int ny, int x, int y, int along_x, double vel, double coe)
{
    double a[7];
    for (int k = -3; k <= 3; ++k)
    {
        const int gx = along_x ? wrap(x + k, nx) : x;
        const int gy = along_x ? y : wrap(y + k, ny);
        a[k + 3] = in[(size_t)gy * nx + gx];
    }
    if (vel > 0.0)
        return weno5_cpu((a[1] - a[0]) * coe, (a[2] - a[1]) * coe, (a[3] - a[2]) * coe, (a[4] - a[3]) * coe, (a[5] - a[4]) * coe);
    return weno5_cpu((a[6] - a[5]) * coe, (a[5] - a[4]) * coe, (a[4] - a[3]) * coe, (a[3] - a[2]) * coe, (a[2] - a[1]) * coe);
}

int custen_oracle_weno(const double* in, const double* u, const double* v, double* out, int nx, int ny, double dx, double dy)
{
    const double cx = 1.0 / dx, cy = 1.0 / dy;
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x)
        {
            const size_t i = (size_t)y * nx + x;
            const double fx = weno_line_cpu(in, nx, ny, x, y, 1, u[i], cx);
            const double fy = weno_line_cpu(in, nx, ny, x, y, 0, v[i], cy);
            out[i] = u[i] * fx + v[i] * fy;
        }
    return 0;
}
