// ORACLE build shim (test infrastructure): the OptiX SDK is not in this image.  The reference's host-only sources that
// oracle/ref_host.cpp compiles reach <optix_stubs.h> only through an inline error helper (utils/cuda/error.h:27-31) that
// none of them calls; this header gives that helper the two names it mentions so that the sources compile unmodified.
#pragma once
#include <cuda.h>
enum OptixResult { OPTIX_SUCCESS = 0 };
inline const char *optixGetErrorString(OptixResult) { return "optix is not available in the oracle build"; }
