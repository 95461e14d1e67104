// cuPentCahnADI re-hosted on cuSten-B200: same command line, same parameters, same initial condition and same printed
// quantity as the reference's GPU timing twin (cuPentSpeedUp/cuPentCahnADITiming/src/cuPentCahnADI.cu:158-180, :485-576):
//
//     ./cuPentCahnADI <n> [steps]
//
// n x n periodic grid, D = 1, gamma = 0.01, lx = 16 pi, dt = 0.1 dx, c0 = U(-0.1, 0.1) from the C library's unseeded
// rand() in row-major order (:142-146, :295-305), time loop `while (t < 10)` unless a step count is given; prints the
// seconds the loop took.  Everything numerical happens in libcusten_b200 (custen_cahn_*), which reproduces the
// reference's GPU solver bit for bit.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/custen_c.h"

static double double_rand(double lo, double hi)
{
    const double scale = (double)rand() / (double)RAND_MAX;
    return lo + scale * (hi - lo);
}

int main(int argc, char* argv[])
{
    if (argc < 2)
    {
        printf("usage: %s n [steps]\n", argv[0]);
        return 1;
    }
    const int n = atoi(argv[1]);
    const double lx = 16.0 * M_PI, dx = lx / n, dt = 0.1 * dx, T = 10.0;
    int steps = 0;
    if (argc > 2) steps = atoi(argv[2]);
    else
        for (double t = 0.0; t < T; t += dt) ++steps;

    std::vector<double> c0((size_t)n * n);
    for (size_t i = 0; i < c0.size(); ++i) c0[i] = double_rand(-0.1, 0.1);

    void* solver = custen_cahn_create(n, 1.0, 0.01, lx, 0.1, 0);
    custen_cahn_set_field(solver, c0.data());
    const float ms = custen_cahn_time_steps(solver, steps);
    printf("%f \n", ms / 1000);

    custen_cahn_get_field(solver, c0.data());
    double mean = 0.0, amax = 0.0;
    for (double v : c0) { mean += v; amax = fmax(amax, fabs(v)); }
    fprintf(stderr, "n = %d, %d steps, %.4f ms/step, mean(c) = %.3e, max|c| = %.4f\n", n, steps, ms / steps,
            mean / c0.size(), amax);
    custen_cahn_destroy(solver);
    return 0;
}
