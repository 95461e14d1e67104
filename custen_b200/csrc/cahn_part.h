// Cahn-Hilliard ADI, tolerance-mode road (cahn_part.cu): one y-slab of the grid per GPU; world = 1 is the engine of the
// single-GPU solver (cahn.cu, solver 2).
#ifndef CUSTEN_B200_CAHN_PART_H
#define CUSTEN_B200_CAHN_PART_H

#include <cuda_runtime.h>

#include "pent_part.h"

namespace custen_cahn {

struct PartSlab;

// nullptr when the partitioned layout cannot take the grid (the caller keeps the bit-identical road)
PartSlab* part_slab_create(int n, int rank, int world, double D, double gamma, double lx, double dt_over_dx, int device,
                           int np_wanted);
void part_slab_connect(PartSlab* s, char* up_block, char* down_block);   // neighbours' blocks as this process sees them
void part_slab_connect_ipc(PartSlab* s, const void* up_handle64, const void* down_handle64);
void part_slab_step(PartSlab* s, int nsteps);                            // asynchronous, on the slab's stream
float part_slab_time_steps(PartSlab* s, int nsteps);
void part_slab_set_fields(PartSlab* s, const double* c_rows_host, const double* cold_rows_host);   // cold null: = c
void part_slab_load_device(PartSlab* s, const double* c_dev, const double* cold_dev);
void part_slab_store_device(PartSlab* s, double* c_dev, double* cold_dev);
void part_slab_get_field(PartSlab* s, double* rows_host);
void part_slab_synchronize(PartSlab* s);
void part_slab_set_graph(PartSlab* s, int on);
void part_slab_set_timeout(PartSlab* s, double seconds);
int part_slab_error(PartSlab* s);
int part_slab_np(const PartSlab* s);
int part_slab_device(const PartSlab* s);
long part_slab_steps(const PartSlab* s);
char* part_slab_block(PartSlab* s);
cudaStream_t part_slab_stream(PartSlab* s);
void part_slab_destroy(PartSlab* s);

void part_set_default_np(int np);
void part_set_rhs_stream(int on);
int part_default_np();

}  // namespace custen_cahn

#endif
