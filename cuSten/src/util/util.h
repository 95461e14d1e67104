// Path-compatibility shim for the reference layout (cuSten/src/util/util.h); see include/cuSten.h.
#include "../../../include/cuSten.h"
