// ORACLE / reference build (test infrastructure only): glue shared by the host shims of oracle/refcpu/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace refcpu {

// first bytes of the "PTX" that shim_ptxgen.cpp hands to the reference's cuda::Module: the path of a host shared object
constexpr const char *JIT_TAG = "//REFCPU_JIT ";

struct Callbacks {
    void (*tex2d)(unsigned long long tex, float u, float v, float out[4]);
    void (*tex3d)(unsigned long long tex, float u, float v, float w, float out[4]);
};

// phase 0 = before the thread runs, 1 = after
typedef void (*ThreadHook)(int linear_thread, int phase, void *user);
typedef void (*LaunchFn)(void **args, const int grid[3], const int block[3], int sequential, ThreadHook hook, void *user);

struct LaunchMode {
    bool sequential = false;  // false: blocks on OpenMP threads; true: one thread after the other on the calling thread
    ThreadHook hook = nullptr;
    void *user = nullptr;
};
void set_launch_mode(const LaunchMode &m);

void sample_tex2d(unsigned long long tex, float u, float v, float out[4]);
void sample_tex3d(unsigned long long tex, float u, float v, float w, float out[4]);
size_t arena_bytes_used();

// ---- OptiX stand-in (shim_optix.cpp): the reference's vendored lib/bvh over the flattened world-space triangles -------------
struct WorldTriangles {
    std::vector<float> verts;           // 9 floats per triangle
    std::vector<uint32_t> inst_prim;    // (instance id, primitive index) per triangle
};
const WorldTriangles &world_triangles(unsigned long long tlas_handle);
// rays in the reference SoA layout; writes the reference's IntersectionSOA (flag: bit 30 hit | instance id; detail t, prim, u, v bits)
void closest_hit(unsigned long long tlas_handle, int n, const float *o_med, const float *d_t1, uint32_t *flag, uint32_t *detail);
void any_hit(unsigned long long tlas_handle, int n, const float *o_med, const float *d_t1, uint8_t *occluded);

}  // namespace refcpu
