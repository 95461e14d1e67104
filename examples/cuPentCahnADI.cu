// cuPentCahnADI re-hosted on cuSten-B200: the command line, parameters, initial condition and printed quantity of the
// reference's GPU timing twin (cuPentSpeedUp/cuPentCahnADITiming/src/cuPentCahnADI.cu:158-180, :485-576), plus the
// snapshot output of the reference's demo driver (cuPentCahnADI/src/cuPentCahnADI.cu:103-140, :592-601):
//
//     ./cuPentCahnADI <n> [steps] [print_every] [output_dir]
//
// n x n periodic grid, D = 1, gamma = 0.01, lx = 16 pi, dt = 0.1 dx, c0 = U(-0.1, 0.1) from the C library's unseeded
// rand() in row-major order (:142-146, :295-305), time loop `while (t < 10)` unless a step count is given; prints the
// seconds the loop took.  With print_every > 0 a snapshot of c is written into output_dir (default "output") every
// print_every steps and after the last one (the reference: every 100), as cahn_hilliard_<time>.bin;
// examples/cahn_analysis.py turns a directory of snapshots into the coarsening statistics s(t) and 1/k1 of the
// reference's plotting.py.  Everything numerical happens in libcusten_b200 (custen_cahn_*).
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
        printf("usage: %s n [steps] [print_every] [output_dir]\n", argv[0]);
        return 1;
    }
    const int n = atoi(argv[1]);
    const double lx = 16.0 * M_PI, dx = lx / n, dt = 0.1 * dx, T = 10.0;
    int steps = 0;
    if (argc > 2) steps = atoi(argv[2]);
    if (steps <= 0)
        for (double t = 0.0; t < T; t += dt) ++steps;
    const int print_every = argc > 3 ? atoi(argv[3]) : 0;
    const char* outdir = argc > 4 ? argv[4] : "output";

    std::vector<double> c0((size_t)n * n);
    for (size_t i = 0; i < c0.size(); ++i) c0[i] = double_rand(-0.1, 0.1);

    void* solver = custen_cahn_create(n, 1.0, 0.01, lx, 0.1, 0);
    custen_cahn_set_field(solver, c0.data());
    float ms = 0.0f;
    if (print_every <= 0)
        ms = custen_cahn_time_steps(solver, steps);
    else
    {
        // the reference's loop: time += dt per step, a snapshot whenever timeCount % print == 0, one more at the end
        double time = 0.0;
        for (int done = 0; done < steps;)
        {
            const int chunk = steps - done < print_every ? steps - done : print_every;
            ms += custen_cahn_time_steps(solver, chunk);
            for (int k = 0; k < chunk; ++k) time += dt;
            done += chunk;
            if (done % print_every == 0 && custen_cahn_write_snapshot(solver, outdir, time) != 0)
            {
                fprintf(stderr, "cannot write a snapshot into %s\n", outdir);
                return 2;
            }
        }
        if (custen_cahn_write_snapshot(solver, outdir, time) != 0) return 2;
    }
    printf("%f \n", ms / 1000);

    custen_cahn_get_field(solver, c0.data());
    double mean = 0.0, amax = 0.0;
    for (double v : c0) { mean += v; amax = fmax(amax, fabs(v)); }
    fprintf(stderr, "n = %d, %d steps, %.4f ms/step, mean(c) = %.3e, max|c| = %.4f\n", n, steps, ms / steps,
            mean / c0.size(), amax);
    custen_cahn_destroy(solver);
    return 0;
}
