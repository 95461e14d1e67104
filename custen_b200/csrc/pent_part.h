// Partitioned cyclic pentadiagonal solve, tolerance mode (pent_part.cu): interface shared with cahn.cu.
#ifndef CUSTEN_B200_PENT_PART_H
#define CUSTEN_B200_PENT_PART_H

#include <cuda_runtime.h>

namespace custen_cahn {

struct PartPlan;   // tables of one (n, np, coefficients): local factors, spikes, banded inverse of the interface ring

// Partition height for an n-row system: `wanted` if it divides n (a multiple of 32, at most 256, at least two
// partitions), else the largest of 128 / 64 / 32 that does; 0 when none does (the caller keeps the bit-identical solve).
int part_choose_np(int n, int wanted);

// coef5 = (a, b, c, d, e): the diagonals -2 .. +2 of the periodic system.  device_tables = false builds the host copies
// only (no CUDA call), for part_solve_host.
PartPlan* part_plan_create(int n, int np, const double coef5[5], bool device_tables);
void part_plan_destroy(PartPlan* plan);
int part_plan_np(const PartPlan* plan);
int part_plan_partitions(const PartPlan* plan);   // P = n / np over the whole system
int part_plan_reach(const PartPlan* plan);        // largest partition offset the interface coupling reaches
const double* part_plan_wv(const PartPlan* plan); // device table [np][4] = {W0, W1, V0, V1} for the consumers' correction

bool part_solve_supported(int nsys, int nrows_local, int np);

// Local solves in place, unknowns along the array's rows: data[row * nsys + sys], nrows_local rows (this GPU's share of
// the n-row systems, whole partitions), plus the interface values G[(p_local * 4 + k) * nsys + sys], k = first, second,
// last-but-one, last unknown of the partition's local solution.  qx != nullptr: the data is the partition-local result
// of part_solve_cols on the same array and still lacks that solve's correction; it is applied while the tiles are
// staged (qx = that solve's interface unknowns from part_reduce, q_nsys = its number of systems = nrows_local).
// Returns false (nothing enqueued) when the layout cannot take the tiles.
bool part_solve_rows(const PartPlan* plan, double* data, int nsys, int nrows_local, double* G, const double* qx, int q_nsys,
                     cudaStream_t stream);

// Local solves in place, unknowns contiguous: data[sys * ld + x], x < ld = n (whole systems on this GPU), nsys systems.
// G as above with P = n / np partitions.
bool part_solve_cols(const PartPlan* plan, double* data, int nsys, int ld, double* G, cudaStream_t stream);

// q[(p_local * 4 + k) * nsys + sys] = the four interface unknowns partition p_local needs for its correction
// x = g - (W0 q0 + W1 q1 + V0 q2 + V1 q3).  gptr_dev: device array of `world` pointers to the ranks' G arrays.
void part_reduce(const PartPlan* plan, const double* const* gptr_dev, int world, int rank, int P_loc, int nsys, double* q,
                 cudaStream_t stream);

void part_solve_host(const PartPlan* plan, const double* rhs, double* x);

}  // namespace custen_cahn

#endif
