// cuSten-B200, additive: register a user __device__ function so the Fun variants can inline it.
//
// The reference hands the kernels an opaque device function pointer (cuSten_t::devFunc, obtained with
// cudaMemcpyFromSymbol, examples/src/2d_xy_p_fun.cu:177-178) and pays one indirect call per grid point.  That
// road is kept unchanged.  A translation unit that DEFINES the function can additionally write, at file scope,
//
//     CUSTEN_REGISTER_FUN_XY(myFunction)      // or _X / _Y for the X / Y contracts
//
// which instantiates the streaming kernel around `myFunction` in that translation unit and records the pair
// (device address of myFunction -> that instance).  cuStenCompute2D*Fun then launches the inlined instance
// whenever cuSten_t::devFunc equals a registered address; arguments, results (bit for bit: same code, same
// compiler, same FMA contraction) and the rest of the API are unchanged.  Unregistered pointers keep working
// through the indirect call.
#ifndef CUSTEN_B200_CUSTEN_FUN_H
#define CUSTEN_B200_CUSTEN_FUN_H

#include "cuSten.h"
#include "../custen_b200/csrc/stream_kernels.cuh"

namespace custen {
struct FunRegistrar
{
    FunRegistration reg;
    FunRegistrar(int dir, const void* (*resolve)(), InlineLauncher launch)
    {
        reg.dir = dir;
        reg.resolve = resolve;
        reg.launch = launch;
        register_fun(&reg);
    }
};
}  // namespace custen

#define CUSTEN_REGISTER_FUN_IMPL(fn, tag, TYPE, DIRV, LAUNCH)                                      \
    __device__ TYPE custen_reg_ptr_##tag = fn;                                                     \
    static const void* custen_reg_resolve_##tag()                                                  \
    {                                                                                              \
        void* p = nullptr;                                                                         \
        cudaMemcpyFromSymbol(&p, custen_reg_ptr_##tag, sizeof(void*));                             \
        return p;                                                                                  \
    }                                                                                              \
    static custen::FunRegistrar custen_reg_obj_##tag(DIRV, custen_reg_resolve_##tag, custen::LAUNCH<fn>);

#define CUSTEN_REGISTER_FUN_X(fn) CUSTEN_REGISTER_FUN_IMPL(fn, fn, cuStenFunX, custen::DIR_X, launch_inline_x)
#define CUSTEN_REGISTER_FUN_Y(fn) CUSTEN_REGISTER_FUN_IMPL(fn, fn, cuStenFunY, custen::DIR_Y, launch_inline_y)
#define CUSTEN_REGISTER_FUN_XY(fn) CUSTEN_REGISTER_FUN_IMPL(fn, fn, cuStenFunXY, custen::DIR_XY, launch_inline_xy)
// for functions whose name is qualified (ns::fn): give the tag explicitly
#define CUSTEN_REGISTER_FUN_X_AS(fn, tag) CUSTEN_REGISTER_FUN_IMPL(fn, tag, cuStenFunX, custen::DIR_X, launch_inline_x)
#define CUSTEN_REGISTER_FUN_Y_AS(fn, tag) CUSTEN_REGISTER_FUN_IMPL(fn, tag, cuStenFunY, custen::DIR_Y, launch_inline_y)
#define CUSTEN_REGISTER_FUN_XY_AS(fn, tag) CUSTEN_REGISTER_FUN_IMPL(fn, tag, cuStenFunXY, custen::DIR_XY, launch_inline_xy)

#endif
