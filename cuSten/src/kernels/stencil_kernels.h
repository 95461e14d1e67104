// Path-compatibility shim for the reference layout (cuSten/src/kernels/stencil_kernels.h); see include/cuSten.h.
#include "../../../include/cuSten.h"
