// nb_lbvh.cuh -- device-side LBVH builder ("GPU LBVH ... the step immediately before the
// path", where Accel::build is a no-op in the reference: src/accel.cpp:19-21).
//
// Morton codes of triangle centroids -> radix sort (cub::DeviceRadixSort; library plumbing) -> Karras 2012 binary
// radix tree (one thread per internal node) -> bottom-up padded AABBs with arrival counters -> subtrees of at most
// max_leaf triangles collapse into leaves -> the same 64 B node / 48 B triangle layout the host SAH builder emits
// (nb_bvh.h).  Builds 10 M triangles in tens of milliseconds instead of seconds; the tree is of lower quality than
// the SAH tree (more node visits per ray), so it is opt-in: nb_set_option(ctx, "builder", 1).
// Results do not depend on the tree, which is how this builder is parity-tested.
#pragma once
#include <cub/cub.cuh>
#include "nb_device.cuh"

namespace nb {

struct LbvhScratch {
    uint64_t *keys = nullptr, *keys_sorted = nullptr;   // (morton30 << 32) | prim
    float4 *leaf_lo = nullptr, *leaf_hi = nullptr;      // per sorted leaf: padded triangle box
    float4 *node_lo = nullptr, *node_hi = nullptr;      // per internal node: box
    int2 *children = nullptr;                           // per internal node: (left, right); >= 0 internal, < 0 => ~leaf
    int *parent = nullptr;                              // [0, n-1) internal parents, [n-1, 2n-1) leaf parents
    int2 *range = nullptr;                              // per internal node: first, last sorted leaf
    unsigned *arrive = nullptr;                         // per internal node arrival counter
    unsigned *emit = nullptr, *emit_index = nullptr;    // per internal node: kept as inner node / new index
    void *cub_tmp = nullptr; size_t cub_bytes = 0;
};

__device__ __forceinline__ unsigned expand_bits10(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void lbvh_keys_kernel(const float4 *verts, const uint4 *faces, unsigned n, float3 clo, float3 cinv, uint64_t *keys) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 f = faces[i];
    const float4 a = verts[f.x], b = verts[f.y], c = verts[f.z];
    const float cx = 0.5f * (fminf(a.x, fminf(b.x, c.x)) + fmaxf(a.x, fmaxf(b.x, c.x)));
    const float cy = 0.5f * (fminf(a.y, fminf(b.y, c.y)) + fmaxf(a.y, fmaxf(b.y, c.y)));
    const float cz = 0.5f * (fminf(a.z, fminf(b.z, c.z)) + fmaxf(a.z, fmaxf(b.z, c.z)));
    const unsigned qx = (unsigned) fminf(fmaxf((cx - clo.x) * cinv.x * 1024.0f, 0.0f), 1023.0f);
    const unsigned qy = (unsigned) fminf(fmaxf((cy - clo.y) * cinv.y * 1024.0f, 0.0f), 1023.0f);
    const unsigned qz = (unsigned) fminf(fmaxf((cz - clo.z) * cinv.z * 1024.0f, 0.0f), 1023.0f);
    const uint64_t m = (uint64_t) ((expand_bits10(qx) << 2) | (expand_bits10(qy) << 1) | expand_bits10(qz));
    keys[i] = (m << 32) | (uint64_t) i;
}

__global__ void lbvh_leaf_boxes_kernel(const float4 *verts, const uint4 *faces, const uint64_t *keys_sorted, unsigned n, float pad,
                                       float4 *leaf_lo, float4 *leaf_hi) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned prim = (unsigned) (keys_sorted[i] & 0xffffffffull);
    const uint4 f = faces[prim];
    const float4 a = verts[f.x], b = verts[f.y], c = verts[f.z];
    leaf_lo[i] = make_float4(fminf(a.x, fminf(b.x, c.x)) - pad, fminf(a.y, fminf(b.y, c.y)) - pad, fminf(a.z, fminf(b.z, c.z)) - pad, 0.f);
    leaf_hi[i] = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)) + pad, fmaxf(a.y, fmaxf(b.y, c.y)) + pad, fmaxf(a.z, fmaxf(b.z, c.z)) + pad, 0.f);
}

// length of the common prefix of keys i and j (keys are unique: the triangle index is part of the key); -1 out of range
__device__ __forceinline__ int lbvh_delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll((long long) (keys[i] ^ keys[j]));
}

// Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees" (HPG 2012), algorithm of fig. 4
__global__ void lbvh_hierarchy_kernel(const uint64_t *keys, int n, int2 *children, int *parent, int2 *range) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? ~gamma : gamma;
    const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    range[i] = make_int2(lo, hi);
    if (left >= 0) parent[left] = i; else parent[(n - 1) + gamma] = i;
    if (right >= 0) parent[right] = i; else parent[(n - 1) + gamma + 1] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void lbvh_fit_kernel(int n, const int2 *children, const int *parent, const float4 *leaf_lo, const float4 *leaf_hi,
                                float4 *node_lo, float4 *node_hi, unsigned *arrive) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = parent[(n - 1) + leaf];
    while (node >= 0) {
        if (atomicAdd(&arrive[node], 1u) == 0u) return;      // first arrival: the sibling subtree is not finished yet
        __threadfence();
        const int2 ch = children[node];
        const float4 llo = ch.x >= 0 ? node_lo[ch.x] : leaf_lo[~ch.x], lhi = ch.x >= 0 ? node_hi[ch.x] : leaf_hi[~ch.x];
        const float4 rlo = ch.y >= 0 ? node_lo[ch.y] : leaf_lo[~ch.y], rhi = ch.y >= 0 ? node_hi[ch.y] : leaf_hi[~ch.y];
        node_lo[node] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.f);
        node_hi[node] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.f);
        __threadfence();
        node = parent[node];
    }
}

// Depth of the radix tree: every leaf climbs to the root and the longest climb is kept.  The 62-bit keys bound it by 62
// (coincident centroids separate only in the index bits), close to the per-lane traversal stack (kStack = 64), so the
// build checks it instead of trusting the bound.
__global__ void lbvh_depth_kernel(int n, const int *parent, int *max_depth) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int d = 1;
    for (int node = parent[(n - 1) + leaf]; node >= 0; node = parent[node]) ++d;
    atomicMax(max_depth, d);
}

// An internal node stays an inner node of the output iff it covers more than max_leaf triangles.
__global__ void lbvh_mark_kernel(int n, const int2 *range, int max_leaf, unsigned *emit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    emit[i] = (range[i].y - range[i].x + 1) > max_leaf ? 1u : 0u;
}

__global__ void lbvh_emit_nodes_kernel(int n, const int2 *children, const int2 *range, const unsigned *emit, const unsigned *emit_index,
                                       const float4 *leaf_lo, const float4 *leaf_hi, const float4 *node_lo, const float4 *node_hi,
                                       float4 *out_nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !emit[i]) return;
    const int2 ch = children[i];
    float4 lo[2], hi[2]; int ref[2];
    const int c2[2] = { ch.x, ch.y };
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int c = c2[k];
        if (c < 0) {                                   // single triangle
            lo[k] = leaf_lo[~c]; hi[k] = leaf_hi[~c];
            ref[k] = (int) ~(((unsigned) (~c) << 3) | 0u);
        } else {
            lo[k] = node_lo[c]; hi[k] = node_hi[c];
            if (emit[c]) ref[k] = (int) emit_index[c];
            else { const int2 r = range[c]; ref[k] = (int) ~(((unsigned) r.x << 3) | (unsigned) (r.y - r.x)); }   // collapsed subtree
        }
    }
    float4 *o = out_nodes + (size_t) emit_index[i] * 4;
    o[0] = make_float4(lo[0].x, hi[0].x, lo[0].y, hi[0].y);
    o[1] = make_float4(lo[1].x, hi[1].x, lo[1].y, hi[1].y);
    o[2] = make_float4(lo[0].z, hi[0].z, lo[1].z, hi[1].z);
    o[3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.f, 0.f);
}

__global__ void lbvh_emit_tris_kernel(const float4 *verts, const uint4 *faces, const uint64_t *keys_sorted, unsigned n, float4 *out_tris) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned prim = (unsigned) (keys_sorted[i] & 0xffffffffull);
    const uint4 f = faces[prim];
    const float4 a = verts[f.x], b = verts[f.y], c = verts[f.z];
    out_tris[(size_t) i * 3 + 0] = make_float4(a.x, a.y, a.z, __uint_as_float(prim));
    out_tris[(size_t) i * 3 + 1] = make_float4(b.x, b.y, b.z, 0.f);
    out_tris[(size_t) i * 3 + 2] = make_float4(c.x, c.y, c.z, 0.f);
}

}  // namespace nb
