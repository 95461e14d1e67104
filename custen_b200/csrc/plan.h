// Host-side plan and tile scheduler of cuSten-B200.
//
// Re-creates, stream-ordered and without host blocking, the reference's per-call tile loop
// (cuSten/src/kernels/2d_xy_p_kernel.cu:542-657, state machine in SURVEY.md appendix B) and its plan
// builders (cuSten/src/struct/custenCreateDestroy2D*.cu).  One generic implementation serves the 12
// variants; the thin per-variant entry points live in api_cpp.cu.
#ifndef CUSTEN_B200_PLAN_H
#define CUSTEN_B200_PLAN_H

#include "../../include/cuSten.h"
#include "engine.h"

namespace custen {

struct Spec
{
    int dir;        // Dir
    int periodic;   // 1 periodic, 0 non-periodic
    int fun;        // 1 function-pointer variant
    int weno;       // 1 WENO advection variant (cuStenCreate2DXYWENOADVp)
};

enum MemKind : int { MK_DEVICE = 0, MK_MANAGED = 1, MK_HOST = 2 };

constexpr int kSlots = 3;  // staging ring depth for host-resident grids
constexpr int kMaxSpans = 4;  // unified-memory arrays of one handle: in, out, (u, v)

// Private state hung behind the public `streams` array (slot numStreams = magic, the next = Plan*), so that
// sizeof(cuSten_t) and every public field offset stay as in the reference.
struct Plan
{
    Spec spec;
    int ncoef;
    int last_path;           // Path of the most recent launch
    int last_mode;           // 0 resident single launch, 1 resident per tile, 2 managed pipeline, 3 staged,
                             // 4 managed + already resident, 5 managed + zero-copy over the host link
    // unified-memory ranges this handle put under "preferred location CPU / accessed by GPU" advice (HOST offload)
    const void* zc_ptr[kMaxSpans];
    size_t zc_bytes[kMaxSpans];
    int zc_n;
    // where the previous unified-memory call left which arrays: 0 unknown, 1 on the GPU, 2 at home on the CPU
    const void* res_ptr[kMaxSpans];
    size_t res_bytes[kMaxSpans];
    int res_n, res_where;
    const double* coef_on_gpu;   // unified-memory coefficients this handle has already prefetched to the GPU
    // slab extension (multi-GPU layer): rows above / below the grid come from these buffers
    const double* slab_top;
    const double* slab_bottom;
    int slab_first, slab_last;   // this slab touches the physical top / bottom of the global grid
    int slab_enabled;
    // slab time stepping (slab.cu): neighbour completion counters handed to the kernels through Band (engine.h)
    const unsigned long long* sync_wait_up;
    const unsigned long long* sync_wait_down;
    unsigned long long* sync_signal_up;
    unsigned long long* sync_signal_down;
    unsigned long long* sync_local;
    // staging for host-resident grids
    double* d_in[kSlots];
    double* d_out[kSlots];
    double* d_aux[kSlots];       // WENO variant: the tile's u rows followed by its v rows
    double* d_coef;
    size_t stage_rows;           // rows each d_in slot can hold
    cudaEvent_t ev_loaded[kSlots], ev_done[kSlots], ev_unloaded[kSlots];
    int events_ready;
    size_t coef_cap;             // doubles d_coef can hold
    // ordering between consecutive calls on one handle: `spread` = the previous call left work on streams[1] / [2]
    // as well as streams[0]; the next call then starts with a three-way join (ev_join)
    cudaEvent_t ev_join[3];
    int join_ready, spread, joined_now;
    int managed_policy;          // per handle: -1 follow the process default, else as custen_set_managed_policy
};

Plan* plan_of(cuSten_t* h);

void plan_create(cuSten_t* h, Spec spec, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y,
                 double* dataOutput, double* dataInput, double* coef, int H, int L, int R, int V, int T, int B,
                 int numCoe, double* func);
void plan_create_weno(cuSten_t* h, int deviceNum, int numTiles, int nx, int ny, int BLOCK_X, int BLOCK_Y, double dx,
                      double dy, double* u, double* v, double* dataOutput, double* dataInput);
void plan_swap(cuSten_t* h, double* dataInput);
void plan_destroy(cuSten_t* h);
void plan_compute(cuSten_t* h, bool offload);
// Slab layer: sweep the handle's whole (device-resident) grid as one band on `stream`, halo rows and neighbour
// counters as set in the plan's slab_* / sync_* fields.  Returns the Path used.
int plan_launch_slab(cuSten_t* h, cudaStream_t stream);

MemKind classify(const void* p);
// 0 (default): unified-memory grids take the resident / zero-copy roads when they apply; 1: always the reference's
// prefetch pipeline
void set_managed_policy(int policy);
// the same per handle (-1: follow the process default again)
void set_handle_managed_policy(cuSten_t* h, int policy);

// What one launch would cover, for tests of the host logic (see debug_bands in plan.cu).
struct BandDesc
{
    long long in_off, out_off, top_off, bottom_off;  // in doubles, relative to the input / output base
    int top_kind, bottom_kind;                       // 0 absent, 1 inside the input array, 2 slab halo buffer
    int rows, nx, L, R, T, B, H, V;
    int wrap_x, xlo, xhi, ylo, yhi, zero_right, contiguous;
};
int debug_bands(int variant, int numTiles, int nx, int ny, int H, int L, int R, int V, int T, int B, int merged,
                int slab, int slab_first, int slab_last, BandDesc* out, int max_out);

}  // namespace custen

#endif
