// Path-compatibility shim for the reference layout (cuSten/src/struct/cuSten_struct_functions.h); see include/cuSten.h.
#include "../../../include/cuSten.h"
