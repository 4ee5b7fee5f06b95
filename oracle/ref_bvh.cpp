// ORACLE (test infrastructure only): thin C-ABI driver around the REFERENCE'S OWN vendored lib/bvh
// (madmann91/bvh v1, /root/reference/lib/bvh/include), compiled from the sources where they lie
// by oracle/Makefile into oracle/_ref/libbvhref.so.  No lib/bvh source is copied into this repo.
//
// It exposes what lib/bvh/test/benchmark.cpp:120-181 does: build a BVH (SweepSahBuilder, or
// BinnedSahBuilder) over Triangle<float> and run SingleRayTraverser + ClosestPrimitiveIntersector
// on a ray batch, returning primitive id, (t, u, v) and the Statistics counters
// (single_ray_traverser.hpp:132-135) that SURVEY 8(d) uses for the algorithmic BVH bytes.
#include <bvh/bvh.hpp>
#include <bvh/binned_sah_builder.hpp>
#include <bvh/primitive_intersectors.hpp>
#include <bvh/ray.hpp>
#include <bvh/single_ray_traverser.hpp>
#include <bvh/sweep_sah_builder.hpp>
#include <bvh/triangle.hpp>
#include <bvh/vector.hpp>

#include <cstdint>
#include <cstring>
#include <vector>

using Scalar = float;
using Vector3 = bvh::Vector3<Scalar>;
using Triangle = bvh::Triangle<Scalar>;
using Ray = bvh::Ray<Scalar>;
using Bvh = bvh::Bvh<Scalar>;

struct ref_bvh {
    std::vector<Triangle> tris;
    Bvh bvh;
};

extern "C" {

// verts: 9 floats per triangle (p0, p1, p2).  builder: 0 = sweep_sah, 1 = binned_sah.
ref_bvh *ref_bvh_build(int ntri, const float *verts, int builder) {
    auto *r = new ref_bvh();
    r->tris.reserve(ntri);
    for (int i = 0; i < ntri; ++i) {
        const float *v = verts + 9 * (size_t)i;
        r->tris.emplace_back(Vector3(v[0], v[1], v[2]), Vector3(v[3], v[4], v[5]), Vector3(v[6], v[7], v[8]));
    }
    auto [bboxes, centers] = bvh::compute_bounding_boxes_and_centers(r->tris.data(), r->tris.size());
    auto global_bbox = bvh::compute_bounding_boxes_union(bboxes.get(), r->tris.size());
    if (builder == 0) {
        bvh::SweepSahBuilder<Bvh> b(r->bvh);
        b.build(global_bbox, bboxes.get(), centers.get(), r->tris.size());
    } else {
        bvh::BinnedSahBuilder<Bvh, 16> b(r->bvh);
        b.build(global_bbox, bboxes.get(), centers.get(), r->tris.size());
    }
    return r;
}

void ref_bvh_free(ref_bvh *r) { delete r; }

// Rays in the reference SoA layout (float4 o|medium, float4 d|t1).  prim[i] = -1 on miss.
// stats[0] += traversal steps, stats[1] += primitive intersections.
void ref_bvh_closest(ref_bvh *r, int n, const float *o_med, const float *d_t1, int32_t *prim, float *tuv, uint64_t *stats) {
    bvh::ClosestPrimitiveIntersector<Bvh, Triangle> intersector(r->bvh, r->tris.data());
    bvh::SingleRayTraverser<Bvh> traverser(r->bvh);
    uint64_t steps = 0, isects = 0;
#pragma omp parallel for reduction(+ : steps, isects) schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        Ray ray(Vector3(o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2]),
                Vector3(d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2]), 0.0f, d_t1[4 * i + 3]);
        bvh::SingleRayTraverser<Bvh>::Statistics st;
        auto hit = traverser.traverse(ray, intersector, st);
        steps += st.traversal_steps;
        isects += st.intersections;
        if (hit) {
            prim[i] = (int32_t)hit->primitive_index;
            tuv[3 * i] = hit->intersection.t; tuv[3 * i + 1] = hit->intersection.u; tuv[3 * i + 2] = hit->intersection.v;
        } else {
            prim[i] = -1;
            tuv[3 * i] = tuv[3 * i + 1] = tuv[3 * i + 2] = 0;
        }
    }
    if (stats) { stats[0] += steps; stats[1] += isects; }
}

// lib/bvh/test/benchmark.cpp:120-181 restated as a library call: primary-ray render of |normal| of the hit
// triangle, the exact procedure whose output lib/bvh/test/scene/cornell_box_reference.png pins.
void ref_bvh_render_normals(ref_bvh *r, const float *eye, const float *dir_in, const float *up, float fov,
                            int width, int height, float *pixels, float *rays_o_med, float *rays_d_t1) {
    auto dir = bvh::normalize(Vector3(dir_in[0], dir_in[1], dir_in[2]));
    auto image_u = bvh::normalize(bvh::cross(dir, Vector3(up[0], up[1], up[2])));
    auto image_v = bvh::normalize(bvh::cross(image_u, dir));
    auto image_w = std::tan(fov * Scalar(3.14159265 * (1.0 / 180.0) * 0.5));
    auto ratio = Scalar(height) / Scalar(width);
    image_u = image_u * image_w;
    image_v = image_v * image_w * ratio;
    bvh::ClosestPrimitiveIntersector<Bvh, Triangle> intersector(r->bvh, r->tris.data());
    bvh::SingleRayTraverser<Bvh> traverser(r->bvh);
#pragma omp parallel for collapse(2)
    for (int i = 0; i < width; ++i) {
        for (int j = 0; j < height; ++j) {
            size_t index = 3 * ((size_t)width * j + i);
            auto u = 2 * (static_cast<Scalar>(i) + Scalar(0.5)) / static_cast<Scalar>(width) - Scalar(1);
            auto v = 2 * (static_cast<Scalar>(j) + Scalar(0.5)) / static_cast<Scalar>(height) - Scalar(1);
            Ray ray(Vector3(eye[0], eye[1], eye[2]), bvh::normalize(image_u * u + image_v * v + dir));
            if (rays_o_med) {
                size_t k = 4 * ((size_t)width * j + i);
                rays_o_med[k] = ray.origin[0]; rays_o_med[k + 1] = ray.origin[1]; rays_o_med[k + 2] = ray.origin[2]; rays_o_med[k + 3] = 0;
                rays_d_t1[k] = ray.direction[0]; rays_d_t1[k + 1] = ray.direction[1]; rays_d_t1[k + 2] = ray.direction[2]; rays_d_t1[k + 3] = ray.tmax;
            }
            auto hit = traverser.traverse(ray, intersector);
            if (!hit) pixels[index] = pixels[index + 1] = pixels[index + 2] = 0;
            else {
                auto normal = bvh::normalize(r->tris[hit->primitive_index].n);
                pixels[index] = std::fabs(normal[0]);
                pixels[index + 1] = std::fabs(normal[1]);
                pixels[index + 2] = std::fabs(normal[2]);
            }
        }
    }
}

}  // extern "C"
