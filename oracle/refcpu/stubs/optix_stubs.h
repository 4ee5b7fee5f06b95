// ORACLE / reference build shim (test infrastructure): the OptiX SDK is not in this image.  The reference's headers that the
// refcpu build compiles (utils/cuda/error.h, utils/optix/{context,as}.h) mention these five names; acceleration structures are
// replaced by oracle/refcpu/shim_optix.cpp (the reference's vendored lib/bvh).  <optix.h> forwards here.
#pragma once
#include <cuda.h>
enum OptixResult { OPTIX_SUCCESS = 0 };
inline const char *optixGetErrorString(OptixResult) { return "optix is not available in the oracle build"; }
typedef unsigned long long OptixTraversableHandle;
typedef struct OptixDeviceContext_t *OptixDeviceContext;
struct OptixBuildInput;
