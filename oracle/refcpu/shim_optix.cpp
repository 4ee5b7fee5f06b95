// ORACLE / reference build (test infrastructure only).
//
// btrc::optix::Context without the OptiX SDK: acceleration-structure creation records the geometry, and closest / any hit run
// on the REFERENCE'S OWN vendored lib/bvh (SweepSahBuilder + SingleRayTraverser + Triangle<float>::intersect,
// /root/reference/lib/bvh/include) over the instances flattened to world space -- the stand-in for OptiX that SURVEY 8(c) adopts
// (OptiX is closed source; the contract is geometric: closest t, double-sided, t in [0, t1]; ties: smallest t, then smallest
// (instance, primitive), the rule the oracle and the CUDA kernels use).
// Replaces utils/optix/context.cpp:158-294 (accel build).  Vertices go to world space with Transform3D::apply_to_point's
// arithmetic (utils/math/transform.cpp:66-72) on the 3x4 matrix the scene hands to OptiX (core/scene.cpp:196-206).
#include <btrc/utils/optix/context.h>

#include <bvh/bvh.hpp>
#include <bvh/primitive_intersectors.hpp>
#include <bvh/ray.hpp>
#include <bvh/single_ray_traverser.hpp>
#include <bvh/sweep_sah_builder.hpp>
#include <bvh/triangle.hpp>
#include <bvh/vector.hpp>

#include <cstring>
#include <memory>
#include <mutex>
#include <optional>

#include "refcpu_internal.h"

namespace {

using Scalar = float;
using V3 = bvh::Vector3<Scalar>;
using Tri = bvh::Triangle<Scalar>;
using Ray = bvh::Ray<Scalar>;
using Bvh = bvh::Bvh<Scalar>;

struct Blas { std::vector<btrc::Vec3f> tri_verts; };  // 3 per triangle
struct Tlas {
    refcpu::WorldTriangles world;
    std::vector<Tri> tris;
    Bvh bvh;
};

std::mutex reg_mutex;
std::vector<std::unique_ptr<Blas>> blas_list;  // handle = index + 1
std::vector<std::unique_ptr<Tlas>> tlas_list;  // handle = (1 << 32) + index

constexpr unsigned long long TLAS_BIT = 1ull << 32;

// closest hit with the deterministic tie rule
struct ClosestTie {
    using Result = bvh::ClosestPrimitiveIntersector<Bvh, Tri>::Result;
    static constexpr bool any_hit = false;
    const Bvh &bvh;
    const Tlas &tlas;
    mutable bool have = false;
    mutable Scalar best_t = 0;
    mutable uint32_t best_inst = 0, best_prim = 0;
    std::optional<Result> intersect(size_t index, const Ray &ray) const {
        const size_t i = bvh.primitive_indices[index];
        auto hit = tlas.tris[i].intersect(ray);
        if (!hit) return std::nullopt;
        const uint32_t inst = tlas.world.inst_prim[2 * i], prim = tlas.world.inst_prim[2 * i + 1];
        const Scalar t = hit->t;
        if (have && !(t < best_t) && (t > best_t || inst > best_inst || (inst == best_inst && prim >= best_prim))) return std::nullopt;
        have = true; best_t = t; best_inst = inst; best_prim = prim;
        return std::make_optional(Result{i, *hit});
    }
};

}  // namespace

namespace refcpu {

const WorldTriangles &world_triangles(unsigned long long handle) { return tlas_list.at(handle - TLAS_BIT)->world; }

void closest_hit(unsigned long long handle, int n, const float *o_med, const float *d_t1, uint32_t *flag, uint32_t *detail) {
    const Tlas &tl = *tlas_list.at(handle - TLAS_BIT);
    bvh::SingleRayTraverser<Bvh> traverser(tl.bvh);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        flag[i] = 0;
        detail[4 * i] = detail[4 * i + 1] = detail[4 * i + 2] = detail[4 * i + 3] = 0;
        if (tl.tris.empty()) continue;
        Ray ray(V3(o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2]), V3(d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2]), 0.0f, d_t1[4 * i + 3]);
        ClosestTie isect{tl.bvh, tl};
        if (auto hit = traverser.traverse(ray, isect)) {
            const size_t k = hit->primitive_index;
            flag[i] = (1u << 30) | tl.world.inst_prim[2 * k];  // wavefront/common.h:12-14
            const float t = hit->intersection.t, u = hit->intersection.u, v = hit->intersection.v;
            std::memcpy(&detail[4 * i], &t, 4);
            detail[4 * i + 1] = tl.world.inst_prim[2 * k + 1];
            std::memcpy(&detail[4 * i + 2], &u, 4);
            std::memcpy(&detail[4 * i + 3], &v, 4);
        }
    }
}

void any_hit(unsigned long long handle, int n, const float *o_med, const float *d_t1, uint8_t *occluded) {
    const Tlas &tl = *tlas_list.at(handle - TLAS_BIT);
    bvh::SingleRayTraverser<Bvh> traverser(tl.bvh);
    bvh::AnyPrimitiveIntersector<Bvh, Tri> isect(tl.bvh, tl.tris.data());
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        if (tl.tris.empty()) { occluded[i] = 0; continue; }
        Ray ray(V3(o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2]), V3(d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2]), 0.0f, d_t1[4 * i + 3]);
        occluded[i] = traverser.traverse(ray, isect) ? 1 : 0;
    }
}

}  // namespace refcpu

BTRC_OPTIX_BEGIN

Context::Context() : context_(nullptr) {}
Context::Context(CUcontext) : context_(reinterpret_cast<OptixDeviceContext>(uintptr_t(1))) {}
Context::Context(Context &&other) noexcept : Context() { swap(other); }
Context &Context::operator=(Context &&other) noexcept { swap(other); return *this; }
Context::~Context() {}
void Context::swap(Context &other) noexcept { std::swap(context_, other.context_); }
Context::operator bool() const { return context_ != nullptr; }
Context::operator OptixDeviceContext() const { return context_; }
void Context::set_message_callback(MessageCallback callback) { message_callback_ = std::move(callback); }

template<typename Index>
TriangleAS Context::create_triangle_as_impl(std::span<const Vec3f> vertices, std::span<const Index> indices)
{
    auto b = std::make_unique<Blas>();
    if(indices.empty())
        b->tri_verts.assign(vertices.begin(), vertices.end());
    else
        for(auto i : indices) b->tri_verts.push_back(vertices[i]);
    std::lock_guard<std::mutex> lk(reg_mutex);
    blas_list.push_back(std::move(b));
    return TriangleAS(static_cast<OptixTraversableHandle>(blas_list.size()), cuda::Buffer<>());
}

TriangleAS Context::create_triangle_as(std::span<const Vec3f> vertices, std::span<const int16_t> indices)
{
    return create_triangle_as_impl(vertices, indices);
}

TriangleAS Context::create_triangle_as(std::span<const Vec3f> vertices, std::span<const int32_t> indices)
{
    return create_triangle_as_impl(vertices, indices);
}

InstanceAS Context::create_instance_as(std::span<const Instance> instances)
{
    auto t = std::make_unique<Tlas>();
    for(auto &inst : instances)
    {
        const Blas &b = *blas_list.at(inst.handle - 1);
        const auto &m = inst.local_to_world;
        const size_t ntri = b.tri_verts.size() / 3;
        for(size_t k = 0; k < ntri; ++k)
        {
            V3 w[3];
            for(int c = 0; c < 3; ++c)
            {
                const Vec3f &p = b.tri_verts[3 * k + c];
                w[c] = V3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3],
                          m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
                          m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
                for(int a = 0; a < 3; ++a) t->world.verts.push_back(w[c][a]);
            }
            t->world.inst_prim.push_back(inst.id);
            t->world.inst_prim.push_back(static_cast<uint32_t>(k));
            t->tris.emplace_back(w[0], w[1], w[2]);
        }
    }
    if(!t->tris.empty())
    {
        auto [bboxes, centers] = bvh::compute_bounding_boxes_and_centers(t->tris.data(), t->tris.size());
        auto global_bbox = bvh::compute_bounding_boxes_union(bboxes.get(), t->tris.size());
        bvh::SweepSahBuilder<Bvh> builder(t->bvh);
        builder.build(global_bbox, bboxes.get(), centers.get(), t->tris.size());
    }
    std::lock_guard<std::mutex> lk(reg_mutex);
    tlas_list.push_back(std::move(t));
    return InstanceAS(TLAS_BIT + tlas_list.size() - 1, cuda::Buffer<>());
}

BTRC_OPTIX_END
