// nb_bvh.h -- host-side BVH builder for the B200 render path (C++17, no CUDA).
//
// Stands where Accel::build() is a no-op in the reference (ref: src/accel.cpp:19-21, called from
// Scene::activate, src/scene.cpp:28).  Produces the SoA device layout consumed by nb_kernels.cuh:
//   nodes : 64 B per inner node = 4 x float4
//             n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//             n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//             n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//             n3 = (ref0, ref1, 0, 0) as int bits; ref >= 0: inner node index,
//                  ref < 0: leaf, ~ref = (first_leaf_triangle << 3) | (count - 1), count in 1..8
//           an absent child (trees with < 2 leaves) is a leaf holding one degenerate all-zero triangle, which
//           Moeller-Trumbore always rejects (det == 0).
//   tris  : 48 B per leaf-ordered triangle = 3 x float4 = (p0.xyz, prim_id bits), (p1.xyz, 0), (p2.xyz, 0)
//           -- exactly the 12 B of indices + 36 B of vertices Mesh::rayIntersect reads (ref: src/mesh.cpp:40-41),
//           pre-gathered so a leaf is one contiguous stream.
// Node 0 is the root.  The first `top_nodes` nodes are in breadth-first order (top of the tree, staged
// into shared memory by TMA), the rest in depth-first order for locality.
#pragma once
#include <cstdint>
#include <vector>

namespace nb {

struct BvhInput {
    const float *verts;        // global vertex array, 4 floats per vertex (xyz + pad)
    const uint32_t *faces;     // global face array, 4 uint32 per triangle (i0, i1, i2, mesh)
    uint32_t nprims;
};

struct BvhOutput {
    std::vector<float> nodes;  // 16 floats per node
    std::vector<float> tris;   // 12 floats per leaf-ordered triangle
    uint32_t nnodes = 0;
    uint32_t top_nodes = 0;    // number of leading nodes in BFS order
    int depth = 0;
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
    double build_seconds = 0;
};

// max_leaf in 1..8; bfs_nodes: how many leading nodes to lay out breadth-first (kSiblingPairs: sibling-pair layout instead,
// the two inner children of a node share one aligned 128-byte line); nthreads <= 0: hardware concurrency
constexpr uint32_t kSiblingPairs = 0xffffffffu;
// max_depth: BvhOutput::depth is guaranteed to stay below it (the per-lane traversal stack of the kernels, nb::kStack)
void build_bvh(const BvhInput &in, BvhOutput &out, int max_leaf = 4, uint32_t bfs_nodes = 2048, int nthreads = 0, int max_depth = 64, int sah_bins = 32);

// ---- on-disk cache of the built hierarchy ("binary blob of V/N/UV/F + BVH"; the step the reference does
// at every start, ref: src/obj.cpp:43-112 + Accel::build src/accel.cpp:19-21).  The key is a 64-bit FNV-1a hash over every
// vertex and index the builder reads plus the build parameters and the layout version, so a stale or foreign file can only
// miss.  File = header { magic, key, nnodes, ntris, top_nodes, depth, scene box } + nodes + leaf-ordered triangles.
uint64_t bvh_cache_key(const BvhInput &in, int max_leaf, uint32_t bfs_nodes, int max_depth = 64, int sah_bins = 32);
bool bvh_cache_load(const char *path, uint64_t key, BvhOutput &out);       // false: missing, unreadable, other key, truncated
bool bvh_cache_save(const char *path, uint64_t key, const BvhOutput &out); // false: directory not writable (not an error)

}  // namespace nb
