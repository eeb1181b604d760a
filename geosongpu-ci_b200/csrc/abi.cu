// extern "C" entry points of the stencils: generated forwarding glue (abi_glue.inc) checked
// against the generated public header.
#include "../../include/b200stencil.h"
#include "impl.cuh"

#include "abi_glue.inc"
