// User-function fixtures for the Fun variants.
//
// The Fun API takes a __device__ function pointer that must be device-linked with the library
// (relocatable device code), so a shared library can only run user functions that were linked into it.
// These are the functions the reference's own programs use, written afresh against the contracts in
// include/cuSten.h; the same header is compiled into the reference shim (oracle/ref_shim.cu) so both
// libraries call identical user code, and oracle/custen_oracle.c restates them for the CPU.
//
//   second_diff_x   3-point second difference scaled by coe[0]        (shape of examples/src/2d_x_np_fun.cu:49-54)
//   weighted_x      sum_k coe[k] * data[loc - nl + k], nl = (n-1)/2 with n = 9 taps
//   weighted9_y     9 rows centred on loc, coe[0..9)                  (shape of examples/src/2d_y_p_fun.cu:49-62)
//   weighted3_y     3 rows centred on loc, coe[0..3)
//   weighted_xy     sum_j sum_i coe[j*nx+i] * data[loc + j*jump + i]  (shape of examples/src/2d_xy_p_fun.cu:58-80)
//   cubic_xy        same sum over (c^3 - c)                           (shape of cuPentCahnADI/src/cuPentCahnADI.cu:164-188)
#ifndef CUSTEN_B200_BUILTIN_FUNS_CUH
#define CUSTEN_B200_BUILTIN_FUNS_CUH

namespace custen_funs {

__device__ inline double second_diff_x(double* data, double* coe, int loc)
{
    return (data[loc - 1] - 2 * data[loc] + data[loc + 1]) * coe[0];
}

__device__ inline double weighted9_x(double* data, double* coe, int loc)
{
    double acc = 0.0;
    for (int k = 0; k < 9; ++k) acc += coe[k] * data[loc - 4 + k];
    return acc;
}

__device__ inline double weighted9_y(double* data, double* coe, int loc, int jump)
{
    double acc = 0.0;
    for (int k = 0; k < 9; ++k) acc += coe[k] * data[loc + (k - 4) * jump];
    return acc;
}

__device__ inline double weighted3_y(double* data, double* coe, int loc, int jump)
{
    double acc = 0.0;
    for (int k = 0; k < 3; ++k) acc += coe[k] * data[loc + (k - 1) * jump];
    return acc;
}

__device__ inline double weighted_xy(double* data, double* coe, int loc, int jump, int nx, int ny)
{
    double acc = 0.0;
    int c = 0;
    for (int j = 0; j < ny; ++j)
    {
        const int row = loc + j * jump;
        for (int i = 0; i < nx; ++i) acc += coe[c++] * data[row + i];
    }
    return acc;
}

__device__ inline double cubic_xy(double* data, double* coe, int loc, int jump, int nx, int ny)
{
    double acc = 0.0;
    int c = 0;
    for (int j = 0; j < ny; ++j)
    {
        const int row = loc + j * jump;
        for (int i = 0; i < nx; ++i)
        {
            const double v = data[row + i];
            acc += coe[c++] * ((v * v * v) - v);
        }
    }
    return acc;
}

}  // namespace custen_funs

#endif
