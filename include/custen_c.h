/* cuSten-B200: plain-C boundary of libcusten_b200.so.
 *
 * One extern "C" entry point per function of the reference's API for the 2D X / Y / XY path, with the
 * reference's parameter lists (plain pointers and ints, no C++ or torch types), so that an FFI (ctypes, cgo,
 * JNI ...) binds exactly what a C++ caller of the reference links against:
 *
 *   custenCreate2D<V>   <->  cuStenCreate2D<V>    cuSten/src/struct/cuSten_struct_functions.h:67-836
 *   custenSwap2D<V>     <->  cuStenSwap2D<V>      (same header)
 *   custenDestroy2D<V>  <->  cuStenDestroy2D<V>   (same header)
 *   custenCompute2D<V>  <->  cuStenCompute2D<V>   cuSten/src/kernels/stencil_kernels.h:52-189
 *   custenCheckError    <->  checkError           cuSten/src/util/util.h:43
 *   <V> in { Xp, Xnp, XpFun, XnpFun, Yp, Ynp, YpFun, YnpFun, XYp, XYnp, XYpFun, XYnpFun, XYWENOADVp }
 *
 * The handle is the reference's cuSten_t (cuSten/src/struct/cuSten_struct_type.h:84-122): caller-allocated,
 * custen_handle_size() bytes, public fields at the reference's offsets.  Error behaviour is the reference's:
 * every function returns void and a CUDA error terminates the process with a message (util/error.cu:43-53).
 * Compute is asynchronous; completion is observed with custen_device_synchronize() (the reference's callers
 * use cudaDeviceSynchronize, examples/src/2d_xy_p.cu:164-167).
 *
 * Everything below the "additive" line has no reference counterpart.
 */
#ifndef CUSTEN_B200_CUSTEN_C_H
#define CUSTEN_B200_CUSTEN_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cuSten_c_handle cuSten_c_handle; /* == cuSten_t, opaque to C callers */

#define CUSTEN_C_COMMON(V)                                                   \
    void custenSwap2D##V(cuSten_c_handle* pt_cuSten, double* dataInput);     \
    void custenDestroy2D##V(cuSten_c_handle* pt_cuSten);                     \
    void custenCompute2D##V(cuSten_c_handle* pt_cuSten, int offload);

#define CUSTEN_C_PREFIX cuSten_c_handle *pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y, \
                        double *dataOutput, double *dataInput

void custenCreate2DXp(CUSTEN_C_PREFIX, double* weights, int numSten, int numStenLeft, int numStenRight);
CUSTEN_C_COMMON(Xp)
void custenCreate2DXnp(CUSTEN_C_PREFIX, double* weights, int numSten, int numStenLeft, int numStenRight);
CUSTEN_C_COMMON(Xnp)
void custenCreate2DXpFun(CUSTEN_C_PREFIX, double* coe, int numSten, int numStenLeft, int numStenRight, int numCoe, double* func);
CUSTEN_C_COMMON(XpFun)
void custenCreate2DXnpFun(CUSTEN_C_PREFIX, double* coe, int numSten, int numStenLeft, int numStenRight, int numCoe, double* func);
CUSTEN_C_COMMON(XnpFun)

void custenCreate2DYp(CUSTEN_C_PREFIX, double* weights, int numSten, int numStenTop, int numStenBottom);
CUSTEN_C_COMMON(Yp)
void custenCreate2DYnp(CUSTEN_C_PREFIX, double* weights, int numSten, int numStenTop, int numStenBottom);
CUSTEN_C_COMMON(Ynp)
void custenCreate2DYpFun(CUSTEN_C_PREFIX, double* coe, int numSten, int numStenTop, int numStenBottom, int numCoe, double* func);
CUSTEN_C_COMMON(YpFun)
void custenCreate2DYnpFun(CUSTEN_C_PREFIX, double* coe, int numSten, int numStenTop, int numStenBottom, double* func);
CUSTEN_C_COMMON(YnpFun)

void custenCreate2DXYp(CUSTEN_C_PREFIX, double* weights, int numStenHoriz, int numStenLeft, int numStenRight,
                       int numStenVert, int numStenTop, int numStenBottom);
CUSTEN_C_COMMON(XYp)
void custenCreate2DXYnp(CUSTEN_C_PREFIX, double* weights, int numStenHoriz, int numStenLeft, int numStenRight,
                        int numStenVert, int numStenTop, int numStenBottom);
CUSTEN_C_COMMON(XYnp)
void custenCreate2DXYpFun(CUSTEN_C_PREFIX, double* coe, int numStenHoriz, int numStenLeft, int numStenRight,
                          int numStenVert, int numStenTop, int numStenBottom, double* func);
CUSTEN_C_COMMON(XYpFun)
void custenCreate2DXYnpFun(CUSTEN_C_PREFIX, double* coe, int numStenHoriz, int numStenLeft, int numStenRight,
                           int numStenVert, int numStenTop, int numStenBottom, double* func);
CUSTEN_C_COMMON(XYnpFun)

/* 13th variant: periodic WENO5 advection (cuStenCreate2DXYWENOADVp, cuSten_struct_functions.h:309) */
void custenCreate2DXYWENOADVp(cuSten_c_handle* pt_cuSten, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X,
                              int BLOCK_Y, double dx, double dy, double* u, double* v, double* dataOutput, double* dataInput);
CUSTEN_C_COMMON(XYWENOADVp)

void custenCheckError(const char* action);

/* ------------------------------------------------ additive ------------------------------------------------ */

size_t custen_handle_size(void);                 /* sizeof(cuSten_t) */
void custen_device_synchronize(void);            /* cudaDeviceSynchronize + checkError */

/* Device function pointers of the user-function fixtures linked into this library (see
 * custen_b200/csrc/builtin_funs.cuh); the value is what cudaMemcpyFromSymbol would give a C++ caller
 * (examples/src/2d_xy_p_fun.cu:177-178).  Returns NULL for an unknown name.
 * Names: second_diff_x weighted9_x weighted9_y weighted3_y weighted_xy cubic_xy */
double* custen_builtin_fun(const char* name);

/* Which kernel family / residency mode served the last Compute on this handle (engine.h Path, plan.h). */
int custen_last_path(cuSten_c_handle* pt_cuSten);
int custen_last_mode(cuSten_c_handle* pt_cuSten);
uint64_t custen_launch_count(void);              /* kernels launched by this library so far */
void custen_set_tuning(int force_fallback, int force_tile, int chunk_rows, int ctas_per_sm, int force_opaque);
/* Unified-memory grids (what the reference requires, cuSten/src/kernels/2d_xy_p_kernel.cu:561-572).  0 (default):
 * offload == DEVICE skips the prefetches when the previous call left the grid on the GPU; offload == HOST sweeps the
 * grid in place over the host link (preferred location CPU + accessed-by advice) instead of migrating every tile both
 * ways.  1: always the reference's prefetch pipeline. */
void custen_set_managed_policy(int policy);
/* The same for one handle only (-1: follow the process-wide setting again). */
void custen_set_handle_managed_policy(cuSten_c_handle* pt_cuSten, int policy);
/* cudaMemAdvise / cudaMemPrefetchAsync (legacy stream) for FFI callers; return the cudaError_t. device -1 = the CPU. */
int custen_mem_advise(const void* p, size_t bytes, int advice, int device);
int custen_mem_prefetch(const void* p, size_t bytes, int device);

/* Multi-GPU y-slab layer: the handle's grid is one slab of a taller global grid.  `top` / `bottom` point at the
 * numStenTop rows above / numStenBottom rows below the slab (a local halo buffer filled by an exchange, or a
 * neighbour GPU's memory mapped through custen_ipc_open); is_first / is_last say whether the slab touches the
 * physical top / bottom edge of the global grid (matters for the non-periodic masks). */
void custen_set_slab(cuSten_c_handle* pt_cuSten, const double* top, const double* bottom, int is_first, int is_last);

/* CUDA IPC plumbing for the slab layer (one process per GPU): 64-byte handles travel over torch.distributed. */
/* custen_ipc_export writes the handle of the allocation `dev_ptr` lives in and the byte offset of `dev_ptr`
 * inside that allocation; the importer adds the offset to what custen_ipc_open returns. */
void custen_ipc_export(const void* dev_ptr, void* handle64, size_t* offset_out);
void* custen_ipc_open(const void* handle64);
void custen_ipc_close(void* mapped_ptr);

/* Neighbour barrier for the "peer" halo transport: flag words live in custen_device_alloc'ed memory (2 x u64 per
 * rank, zero-initialised, exported with custen_ipc_export); up_flags / down_flags are the neighbours' mapped flag
 * blocks (NULL where there is no neighbour).  Enqueued on the handle's compute stream, or on the legacy default
 * stream when pt_cuSten is NULL. */
void custen_peer_barrier(cuSten_c_handle* pt_cuSten, void* up_flags, void* down_flags, void* my_flags, uint64_t epoch);
void* custen_device_alloc(size_t bytes);
void custen_device_free(void* p);

/* Host-logic probe, no CUDA calls: the bands (tile, seam pointers, write masks) Compute would launch for a variant
 * (index into Xp Xnp XpFun XnpFun Yp Ynp YpFun YnpFun XYp XYnp XYpFun XYnpFun).  `out` receives up to max_out
 * records of { long long in_off, out_off, top_off, bottom_off; int top_kind, bottom_kind, rows, nx, L, R, T, B, H, V,
 * wrap_x, xlo, xhi, ylo, yhi, zero_right, contiguous; }; returns the number written. */
int custen_debug_bands(int variant, int numTiles, int nx, int ny, int H, int L, int R, int V, int T, int B, int merged,
                       int slab, int slab_first, int slab_last, void* out, int max_out);

/* Event timing on the stream a handle launches on (streams[idx] of the handle). */
void* custen_event_create(void);
void custen_event_record(void* ev, cuSten_c_handle* pt_cuSten, int stream_idx);
void custen_event_synchronize(void* ev);
float custen_event_elapsed_ms(void* start, void* stop);
void custen_event_destroy(void* ev);

/* Pinned host buffers for the out-of-core path; unified-memory buffers (what the reference's callers use,
 * examples/src/2d_x_p.cu:73-75). */
void* custen_host_alloc(size_t bytes);
void custen_host_free(void* p);
/* Pinned host memory bound to the NUMA node of `device` (mmap + mbind + cudaHostRegister); *node_out = that node, or -1
 * when binding was not possible.  custen_link_probe: plain cudaMemcpyAsync both ways at once over those buffers, ms. */
int custen_device_numa_node(int device);
void* custen_host_alloc_near(size_t bytes, int device, int* node_out);
void custen_host_free_near(void* p, size_t bytes);
float custen_link_probe(const void* host_src, void* host_dst, size_t bytes, int iters, int device);
void* custen_managed_alloc(size_t bytes);
void custen_managed_free(void* p);

/* ---- Multi-GPU y-slab layer with time stepping (custen_b200/csrc/slab.cu; new, no reference counterpart) ----------
 * A slab owns rows [rank ny/world, (rank+1) ny/world) of BOTH field buffers of a global ny x nx grid; the halo rows of
 * a sweep are read in place from the neighbour GPUs' memory - the reference's boundaryTop / boundaryBottom kernel
 * arguments (cuSten/src/kernels/2d_xy_p_kernel.cu:67-68) pointed at peer memory - and custen_slab_swap re-aliases input,
 * output and seams like cuStenSwap2D* (cuSten/src/struct/custenCreateDestroy2DXYp.cu:253-310).  The neighbour wait is
 * inside the sweep kernel (only the work items that touch halo rows wait; the last CTA publishes the slab's sweep
 * count), so a step is one launch and steps replay from a CUDA graph.
 * variant: index into Xp Xnp XpFun XnpFun Yp Ynp YpFun YnpFun XYp XYnp XYpFun XYnpFun; coef_host: ncoef doubles
 * (weights, or the function's coefficients) in host memory; func: device function pointer for the Fun variants, valid
 * on `device` (custen_builtin_fun, or cudaMemcpyFromSymbol in the caller).
 * One process per GPU: create -> export (64-byte IPC handle + offset, moved by the caller) -> connect (NULL where the
 * slab touches a physical edge of a non-periodic grid) -> [fill custen_slab_field(.,0)] -> host barrier -> run. */
void* custen_slab_create(int variant, int device, int rank, int world, int nx, int ny_global, const double* coef_host,
                         int ncoef, int H, int L, int R, int V, int T, int B, double* func);
void custen_slab_export(void* slab, void* handle64, size_t* offset_out);
void custen_slab_connect(void* slab, const void* up_handle64, size_t up_offset, const void* down_handle64, size_t down_offset);
double* custen_slab_field(void* slab, int which);      /* 0: input of the next sweep (= latest result), 1: its output */
int custen_slab_rows(void* slab);
void custen_slab_compute(void* slab);                   /* one sweep, asynchronous */
void custen_slab_swap(void* slab);
void custen_slab_run(void* slab, int nsteps);           /* nsteps x (compute + swap), pairs replayed from a CUDA graph */
void custen_slab_run_plain(void* slab, int nsteps);     /* the same, launch by launch */
float custen_slab_time_run(void* slab, int nsteps);     /* milliseconds (CUDA events on the slab's stream); synchronises */
void custen_slab_synchronize(void* slab);
int custen_slab_error(void* slab);                      /* 1: a neighbour wait timed out */
void custen_slab_set_timeout(void* slab, double seconds); /* default 20 s; 0 = wait for ever */
int custen_slab_last_path(void* slab);
void custen_slab_destroy(void* slab);

/* The same layer for ONE process driving several GPUs (what an existing single-process cuSten program can call;
 * SURVEY.md section 7 step 5).  devices[i] carries slab i; peers are reached with cudaDeviceEnablePeerAccess.  A device may
 * appear more than once (its slabs then share a stream and run in rank order) - the layout the one-GPU tests use.
 * builtin_fun names one of the library's fixtures, else funcs[i] is the user function's address on devices[i]. */
void* custen_mg_create(int ndev, const int* devices, int variant, int nx, int ny, const double* coef_host, int ncoef, int H,
                       int L, int R, int V, int T, int B, const char* builtin_fun, double* const* funcs);
void custen_mg_scatter(void* mg, const double* host_field);          /* ny x nx -> the slabs' current input */
void custen_mg_gather(void* mg, double* host_field, int which);      /* synchronises; which as custen_slab_field */
void custen_mg_fill_output(void* mg, double value);
void custen_mg_compute(void* mg);
void custen_mg_swap(void* mg);
void custen_mg_run(void* mg, int nsteps);
void custen_mg_synchronize(void* mg);
int custen_mg_error(void* mg);
void* custen_mg_slab(void* mg, int i);
void custen_mg_destroy(void* mg);

/* Synthetic input, regenerable anywhere: field[r][c] = lo + (hi - lo) * u(seed, (row0 + r) * nx + c), u = the top 53
 * bits of splitmix64(seed + index) / 2^53 (tests/cases.py hash_field is the numpy twin).  Legacy stream. */
void custen_fill_hash(double* dev_field, long long row0, int rows, int nx, unsigned long long seed, double lo, double hi);

/* Cahn-Hilliard ADI solver re-hosted on the engine (reference program cuPentCahnADI/src/cuPentCahnADI.cu and its
 * timing twin cuPentSpeedUp/cuPentCahnADITiming/src/cuPentCahnADI.cu, whose main() this replaces):
 * n x n periodic grid of side lx, dt = dt_over_dx * lx / n; the reference uses D = 1, gamma = 0.01, dt_over_dx = 0.1.
 * set_field sets c(t=0) = c(t=-dt) like the reference's initial condition (cuPentCahnADI.cu:295-305). */
void* custen_cahn_create(int nx, double D, double gamma, double lx, double dt_over_dx, int device);
void custen_cahn_set_field(void* solver, const double* c0_host);
void custen_cahn_step(void* solver, int nsteps);               /* asynchronous */
void custen_cahn_get_field(void* solver, double* out_host);     /* synchronises */
float custen_cahn_time_steps(void* solver, int nsteps);         /* milliseconds for nsteps steps (CUDA events) */
void custen_cahn_destroy(void* solver);
void custen_cahn_set_fused(int on);                             /* 1: fused right-hand-side pass (default), 0: via cuStenCompute2D* */
void custen_cahn_set_graph(int on);                             /* 1: replay the fused step from a CUDA graph (default), 0: kernel by kernel */
/* The custen_cahn_set_* switches are the defaults for solvers created AFTERWARDS; every solver keeps its own copy
 * (custen_cahn_config changes one solver).
 * solver 2 (default): partitioned tolerance-mode solve (custen_b200/csrc/pent_part.cu, step in cahn_part.cu) - within
 * 1e-13 of the reference's solver per step (cuPentBatch.cu:119-198 + BatchHyper.cu:195-259), not bit-identical; falls
 * back to 0 where the layout cannot take it (n not a multiple of 64).  0: TMA-fed solve in the reference's operation
 * order (bit-identical), 1: its cp.async twin. */
void custen_cahn_set_solver(int which);
void custen_cahn_set_partition_rows(int np);                    /* rows per partition of solver 2: 32, 64, 128 or 256 (default 128) */
int custen_cahn_config(void* solver, int key, int value);       /* key 0 solver, 1 fused, 2 graph, 3 table rows; value < 0 queries */
/* The partitioned algorithm's arithmetic on the host for ONE periodic pentadiagonal system (diagonals coef5 = a b c d e
 * at offsets -2 .. +2): pins tables and algorithm against a dense solve without a GPU.  Returns the number of
 * interface coupling blocks kept, 0 if (n, np) is not a valid partitioning. */
int custen_pent_part_host(int n, int np, const double* coef5, const double* rhs, double* x);
int custen_pent_part_choose_np(int n, int wanted);
/* The same systems through the DEVICE kernels: layout 0 rhs[sys * n + i], 1 rhs[i * nsys + sys], 2 the ADI pair on an
 * n x n array (along x, then along y); bit-identical to custen_pent_part_host per system. */
int custen_pent_part_device(int n, int np, const double* coef5, int nsys, const double* rhs_host, double* x_host, int layout);
void custen_cahn_set_rhs_stream(int on);                        /* tuning / tests, solver 2: 1 row-streaming right-hand side (default), 0 tile kernel */
void custen_cahn_set_table_rows(int rows);                      /* tuning / tests: coefficient rows staged per refill */

/* Snapshot of c(t) into <directory>/cahn_hilliard_<time, ten decimals>.bin (the reference's Print_Out naming,
 * cuPentCahnADI.cu:103-140, in raw binary instead of HDF5: "CUSTENC1" | int64 nx | int64 ny | double time | nx*ny
 * doubles); 0 on success.  custen_cahn_dt: the time step the solver uses. */
int custen_cahn_write_snapshot(void* solver, const char* directory, double time);
double custen_cahn_dt(void* solver);
/* c(t) and c(t - dt) separately (restart from a saved pair of fields); c_old_host NULL: both are c_host */
void custen_cahn_set_fields(void* solver, const double* c_host, const double* c_old_host);

/* The tolerance-mode solver on y-slabs over several GPUs (BASELINE.json config 5 at 2-8 GPUs; new - the reference is
 * single-GPU; custen_b200/csrc/cahn_part.cu).  Slab g of `world` owns rows [g n/world, (g+1) n/world) in the grid's own
 * layout.  Per step a GPU reads from its two neighbours only (a) the 2 + 2 halo rows of c and c(t - dt) for the
 * right-hand side and (b) the 4 interface values per system of the y-direction partitions next to the seam - no
 * transposes, no all-to-all; the slabs order themselves with counters in device memory, so steps are enqueued without
 * host synchronisation and replay from CUDA graphs.  Results are bit-identical to custen_cahn_* with solver 2 on one GPU.
 *
 *   one process per GPU   create -> export (a 64-byte cudaIpcMemHandle_t; the caller moves it: MPI, torch.distributed, a
 *                         file) -> connect(up's handle, down's handle) -> [barrier] -> set_field -> [barrier] -> step ...
 *   one process, G GPUs   custen_cahn_mg_*: the same slabs, peers reached through cudaDeviceEnablePeerAccess.
 * create returns NULL when the partitioned layout cannot take the grid: n / world must be a multiple of the partition
 * height (custen_cahn_set_partition_rows: 32, 64, 128 or 256) and the interface coupling must not reach past the
 * nearest neighbour. */
void* custen_cahn_slab_create(int nx, int rank, int world, double D, double gamma, double lx, double dt_over_dx, int device);
void custen_cahn_slab_export(void* slab, void* handle64);
void custen_cahn_slab_connect(void* slab, const void* up_handle64, const void* down_handle64);
void custen_cahn_slab_connect_local(void* slab, void* up_slab, void* down_slab);
void custen_cahn_slab_set_field(void* slab, const double* rows_host);   /* this slab's rows; c(t = 0) = c(t = -dt) */
void custen_cahn_slab_set_fields(void* slab, const double* c_rows_host, const double* c_old_rows_host);
void custen_cahn_slab_get_field(void* slab, double* rows_host);          /* synchronises this slab's stream */
void custen_cahn_slab_step(void* slab, int nsteps);                      /* asynchronous */
float custen_cahn_slab_time_steps(void* slab, int nsteps);               /* milliseconds (CUDA events on the slab's stream) */
void custen_cahn_slab_synchronize(void* slab);
int custen_cahn_slab_error(void* slab);                                  /* neighbour waits that timed out (0 = fine) */
void custen_cahn_slab_set_timeout(void* slab, double seconds);           /* default 10 s */
void custen_cahn_slab_set_graph(void* slab, int on);
int custen_cahn_slab_partition_rows(void* slab);
void custen_cahn_slab_destroy(void* slab);

void* custen_cahn_mg_create(int nx, int ngpus, const int* devices, double D, double gamma, double lx, double dt_over_dx);
void custen_cahn_mg_set_field(void* mg, const double* c0_host);         /* whole n x n grid */
void custen_cahn_mg_get_field(void* mg, double* out_host);
void custen_cahn_mg_step(void* mg, int nsteps);                         /* asynchronous on every GPU */
float custen_cahn_mg_time_steps(void* mg, int nsteps);                  /* milliseconds, the slowest GPU */
int custen_cahn_mg_error(void* mg);
void custen_cahn_mg_set_graph(void* mg, int on);
void custen_cahn_mg_destroy(void* mg);

#ifdef __cplusplus
}
#endif

#endif
