// cuSten-B200 engine internals: what one sweep over a band of rows looks like to the kernels.
//
// A "band" is a run of consecutive grid rows handled by one launch: a tile of the reference's
// numTiles loop (cuSten/src/kernels/2d_xy_p_kernel.cu:582-616), a y-slab of the multi-GPU layer,
// or the whole grid when the tiles are contiguous and resident.  The rows just above / below the
// band are reached through `top` / `bottom`, exactly the role of the reference's
// boundaryTop / boundaryBottom kernel arguments (2d_xy_p_kernel.cu:67-68): they may alias the same
// array (periodic wrap, tile seams), a staging buffer, or a neighbour GPU's memory.
#ifndef CUSTEN_B200_ENGINE_H
#define CUSTEN_B200_ENGINE_H

#include <cuda_runtime.h>
#include <stdint.h>

namespace custen {

enum Dir : int { DIR_X = 0, DIR_Y = 1, DIR_XY = 2 };

struct Band
{
    const double* in;      // first row of the band (row pitch = nx)
    const double* top;     // T rows directly above the band, or nullptr when have_top == 0
    const double* bottom;  // B rows directly below the band, or nullptr when have_bottom == 0
    double* out;           // first output row of the band
    const double* coef;    // weights (weights variants) or coe (Fun variants), device-readable
    const void* func;      // user __device__ function (Fun variants), else nullptr
    int nx;                // points per row
    int rows;              // rows in the band
    int L, R, T, B;        // taps left / right / above / below the centre
    int H, V;              // window width / height actually summed (H = numSten for X, ...)
    int ncoef;             // number of doubles behind coef
    int dir;               // Dir
    int wrap_x;            // 1: x is periodic (index map wraps), 0: columns outside [0,nx) do not exist
    int have_top;          // 1: `top` is readable
    int have_bottom;       // 1: `bottom` is readable
    int xlo, xhi;          // columns written: xlo <= x < xhi
    int ylo, yhi;          // band-local rows written: ylo <= y < yhi
    int zero_right;        // Xnp quirk (2d_x_np_kernel.cu:164-176): columns x >= xhi receive 0.0
    // WENO advection variant only (2d_xyADVWENO_p_kernel.cu): velocities (same row pitch as `out`) and 1/dx, 1/dy
    const double* aux0;
    const double* aux1;
    double p0, p1;
    int weno;
    // Slab time stepping (multi-GPU layer, slab.cu): completion counters of this slab and of its neighbours.  All null
    // when unused.  The sweep's producer warp waits for a neighbour only before it fetches that neighbour's halo rows
    // (or lets the consumers overwrite rows that neighbour may still be reading); the last CTA to finish publishes this
    // slab's new sweep count to both neighbours.  No separate barrier kernel.
    const unsigned long long* wait_up;    // word in this slab's memory the slab above publishes its sweep count into
    const unsigned long long* wait_down;  // ... the slab below
    unsigned long long* signal_up;        // words in the neighbours' memory this slab publishes its own count into
    unsigned long long* signal_down;
    unsigned long long* sync_local;       // [0] sweeps completed by this slab, [1] CTAs finished in the current launch,
                                          // [2] sticky time-out flag, [3] time-out in ns (0 = wait for ever)
    int guard_top, guard_bottom;          // the band's first / last rows that a neighbour reads as its halo
};

// Which kernel family served a launch (reported through the C ABI for tests / bench).
enum Path : int
{
    PATH_NONE = 0,
    PATH_STREAM_ACC = 1,   // TMA-fed row streaming, register accumulators (weights variants)
    PATH_STREAM_TILE = 2,  // TMA-fed row streaming, shared window handed to the user function
    PATH_FALLBACK = 3,     // plain-load halo tile (odd nx, unaligned rows, exotic shapes)
    PATH_STREAM_INLINE = 4 // as PATH_STREAM_TILE, user function registered and inlined (include/cuSten_fun.h)
};

struct Tuning
{
    int force_fallback;    // 1: always use PATH_FALLBACK (tests cross-check the two families)
    int chunk_rows;        // 0 = choose automatically
    int ctas_per_sm;       // 0 = choose automatically
    int force_tile;        // 1: weights variants use PATH_STREAM_TILE (tests)
    int force_opaque;      // 1: Fun variants always call through the device pointer, registered or not
};

Tuning& tuning();

// Enqueue one band sweep on `stream`; returns the Path used.  Never synchronises.
int launch_band(const Band& band, cudaStream_t stream);

// Counters (process-wide, relaxed): kernels launched by this library.
uint64_t launches_total();
void launches_add(uint64_t n);   // kernels replayed from a captured graph

}  // namespace custen

#endif
