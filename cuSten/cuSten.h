// Path-compatibility shim: programs written against the reference say #include "<...>/cuSten/cuSten.h"
// (e.g. examples/src/2d_x_p.cu:34).  Everything lives in include/cuSten.h.
#include "../include/cuSten.h"
