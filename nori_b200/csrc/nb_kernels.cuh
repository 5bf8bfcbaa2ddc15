// nb_kernels.cuh -- the sm_100a kernels of the render hot path.
//
//   K1 raygen      : pcg32 + PerspectiveCamera::sampleRay        (ref: src/main.cpp:41-46, src/perspective.cpp:76-97)
//   K2 traversal   : BVH closest-hit / any-hit                    (ref: src/accel.cpp:23-43, src/mesh.cpp:39-76, include/nori/bbox.h:323-350)
//   K3 hit record  : Intersection fill                            (ref: src/accel.cpp:45-96)
//   K4 integrator  : normals / ao / whitted / path_{mats,ems,mis} (interface ref: include/nori/integrator.h:42)
//   K5 film splat  : ImageBlock::put(pos, value)                  (ref: src/block.cpp:62-91)
//   K6 film merge  : ImageBlock::put(block)                       (ref: src/block.cpp:93-102)
//
// K1..K5 are fused into ONE persistent-threads kernel (render_kernel): every lane owns one light path at a
// time.  The warp runs in lock step over RAYS: a shading phase in which all 32 lanes advance their path by one
// vertex (hit record, integrator step, next ray; finished paths are splatted and replaced from a warp-local
// pool of (pixel, sample) items refilled from a global atomic counter, free lanes ranked by ballot/popc), then a
// traversal phase in which every lane walks its ray through the BVH to completion (while-while, per-lane stack,
// speculative: a found leaf is parked while the lane keeps descending).  Rays never leave registers.
// Optionally the top of the BVH is staged into shared memory once per CTA with a TMA bulk copy (cp.async.bulk).
#pragma once
#include "nb_device.cuh"
#include "nb_wide.h"

namespace nb {

// Resident CTAs per SM the render kernel is compiled for (128 threads each).  10 -> 48 registers/thread, 40 warps/SM:
// measured fastest on B200 (profiles/r1_v1_regcap_sweep.txt) -- the walk is latency bound, occupancy beats spills.
#ifndef NB_MIN_BLOCKS
#define NB_MIN_BLOCKS 10
#endif
// Path integrators (whitted / path_*) carry more live state and heavier shading code; they get their own cap.
#ifndef NB_MIN_BLOCKS_PATH
#define NB_MIN_BLOCKS_PATH 11
#endif
#ifndef NB_SPEC_VOTE
#define NB_SPEC_VOTE 1
#endif
#ifndef NB_WALK_NOINLINE
#define NB_WALK_NOINLINE 0
#endif
#ifndef NB_PHASED_WAVES
#define NB_PHASED_WAVES 1
#endif
#ifndef NB_PHASED_AO
#define NB_PHASED_AO 1       // ajax-ao 9.59 -> 9.20 ms, random10m-ao 11.8 -> 11.2 ms (profiles/r1_final_summary.md)
#endif
#ifndef NB_WAVEFRONT
#define NB_WAVEFRONT 0       // 1 only in nb_wave.cu: occlusion rays (ao, next-event estimation, simple) go to the engine's shadow queue
#endif
#ifndef NB_SPLAT_HOIST
#define NB_SPLAT_HOIST 1
#endif
#ifndef NB_WATCHDOG_CYCLES
#define NB_WATCHDOG_CYCLES 40000000000LL   // ~20 s at 1.9 GHz
#endif
#ifndef NB_STACK
#define NB_STACK 64
#endif
constexpr int kStack = NB_STACK;    // per-lane traversal stack; the builders guarantee depth < kStack (nb_bvh.cpp, nb_lbvh.cuh)
#ifndef NB_WIDE
#define NB_WIDE 0            // 1: the walk runs on the 8-wide compressed hierarchy (nb_wide.h) instead of the binary one
#endif
#ifndef NB_WIDE_POSTPONE
#define NB_WIDE_POSTPONE 0   // wide walk: triangles are tested only once this many lanes of the warp hold some (0: at once)
#endif
#if NB_WIDE
typedef uint2 StackT; constexpr int kStackN = kWideStack;
#else
typedef int StackT; constexpr int kStackN = kStack;
#endif
constexpr int kBlockEdgeMax = 32 + 2 * 8;

// Tile t of the image's grid of 32x32 tiles (ref: BlockGenerator, src/block.cpp:109-152; the order is cosmetic there).  GPU
// t % N owns tile t; the numbering comes from a table the host builds for N ranks (nb_api.cu: build_tile_order; entry = bx | by << 16)
// so that ownership follows the Latin pattern (bx + 3 * by) % N: every row and column of tiles is dealt evenly to all GPUs.
// Row-major numbering hands whole COLUMNS of tiles to a GPU when N divides the row length (768 / 32 = 24 tiles on 8 GPUs: the
// ranks holding the object's columns ran 15 % longer, profiles/r2_call8_static_guided_ab.txt).
__device__ __forceinline__ void tile_xy(const uint32_t *tab, int t, int &bx, int &by) {
    const uint32_t v = __ldg(&tab[t]);
    bx = (int) (v & 0xffffu); by = (int) (v >> 16);
}

struct SceneDev {
    const float4 *nodes;            // 4 x float4 per node
    const float4 *tris;             // 3 x float4 per leaf-ordered triangle
    const uint4 *faces;             // per global triangle: i0, i1, i2 (global vertex ids), mesh
    const float4 *verts;            // per global vertex: xyz
    const float4 *normals;          // per global vertex: xyz (meshes with normals)
    const float2 *uvs;              // per global vertex
    const DevMesh *meshes;
    const float *emitter_cdf;
    const int32_t *emitters;        // mesh ids of emitters
    int32_t n_emitters;
    uint32_t n_nodes, n_prims;
};

struct RenderParams {
    SceneDev sc;
    float s2c[16], c2w[16];
    int32_t W, H;
    float invW, invH, nearClip, farClip;
    float ftable[33];
    float fradius, lookup;
    int32_t border;
    uint32_t spp;
    int32_t seed_mode;
    uint64_t seed;
    int32_t integrator, rr_start, max_depth;
    int32_t tile_rank, tile_nranks;
    int32_t ntx, nty;               // tiles in x / y
    int32_t n_my_tiles;
    int32_t block_edge;             // 32 + 2*border
    uint32_t chunk;                 // samples per work unit
    uint32_t nchunks;
    uint32_t n_units;
    float4 *blocks;                 // n_my_tiles x block_edge x block_edge
    unsigned long long *counters;   // [0] next unit, [1] rays, [2] node visits, [3] tri tests, [4] hits shaded
    int32_t smem_nodes;             // nodes staged in shared memory (0 = none)
    int32_t block_stream_skip;      // per-block seeding served by skip-ahead (fixed draws per sample)
    int32_t tail_lanes;             // wavefront engine: a walk is suspended once <= tail_lanes lanes are still walking (0 = never)
    float light_pos[3], light_energy[3];   // point light of the `simple` integrator (appended: older fields keep their offsets)
    // wavefront engine (nb_wave.cu / nb_wave.cuh; unused by the other kernels)
    float4 *occ_queue;              // occlusion queue: 3 x float4 per ray (origin | mint, direction | maxt, radiance | slot)
    uint32_t occ_capacity;          // rays the queue holds (= wf_pool)
    uint32_t sample_offset;         // progressive frames: this pass renders the sample streams sample_offset .. sample_offset + spp - 1 of every pixel
    float4 *wf_cols;                // path pool, structure of arrays: kWfCols columns of wf_pool float4 each
    uint32_t *wf_ext;               // queue of pool slots whose extension ray waits to be traced
    uint32_t *wf_ctr;               // engine counters (nb_wave.cuh: WF_*)
    uint32_t wf_pool;               // path slots (multiple of 128)
    const uint32_t *tile_tab;       // tile t -> bx | by << 16, Z-order numbering (tile_xy)
    uint32_t split_units, split_sample, chunk_a, nchunks_a;   // fused kernel, guided schedule: coarse units first (samples [0, split_sample))
    uint32_t wf_chunk;              // samples per work unit (<= 8)
    unsigned long long wf_total;    // sample indices to hand out (virtual: ragged tiles / last chunk included)
};

// ------------------------------------------------------------------ traversal state (per lane)
struct Ray {
    float ox, oy, oz, dx, dy, dz, mint, maxt;
};

struct Trav {
    float idx, idy, idz, oodx, oody, oodz;   // 1/d and o/d for the slab test
    float hu, hv;                             // barycentrics of the closest hit
    uint32_t hprim;                           // NB_MISS if none
    int node;                                 // current node ref; INT_MAX-like sentinel when finished
    int sp;
};

constexpr int kDone = 0x7fffffff;

__device__ __forceinline__ void trav_begin(const Ray &r, Trav &t) {
    const float ooeps = 1e-24f;   // avoid inf * 0 in the slab test (the reference special-cases d == 0: bbox.h:331-333)
    float dx = fabsf(r.dx) > ooeps ? r.dx : copysignf(ooeps, r.dx);
    float dy = fabsf(r.dy) > ooeps ? r.dy : copysignf(ooeps, r.dy);
    float dz = fabsf(r.dz) > ooeps ? r.dz : copysignf(ooeps, r.dz);
    t.idx = 1.0f / dx; t.idy = 1.0f / dy; t.idz = 1.0f / dz;
    t.oodx = r.ox * t.idx; t.oody = r.oy * t.idy; t.oodz = r.oz * t.idz;
    t.hprim = 0xffffffffu; t.hu = 0.f; t.hv = 0.f;
    t.node = 0; t.sp = 0;
}

template <bool TMA_TOP>
__device__ __forceinline__ float4 ld_node(const SceneDev &sc, const float4 *snodes, int smem_nodes, int node, int k) {
    if (TMA_TOP && node < smem_nodes) return snodes[node * 4 + k];
    return __ldg(&sc.nodes[(size_t) node * 4 + k]);
}

// The while-while BVH walk of one ray, run to completion.
// any_hit: stop at the first accepted triangle (Accel::rayIntersect(..., shadowRay=true), ref: src/accel.cpp:35-36).
// Closest hit keeps shrinking r.maxt (ref: src/accel.cpp:37); among equal t the highest triangle index wins,
// which is what the reference's ascending loop with "t <= maxt" produces (ref: src/mesh.cpp:75) -- so the result
// does not depend on the order in which candidate leaves are tested.
// NB_SPECULATIVE: a lane that reaches a leaf parks it and keeps walking inner nodes until it finds a second leaf
// (or every lane of the warp holds one); triangle work then runs with more lanes active.
#ifndef NB_SPECULATIVE
#define NB_SPECULATIVE 1
#endif

template <bool COUNT>
__device__ __forceinline__ bool leaf_test(const SceneDev &sc, int leaf, Ray &r, Trav &t, bool any_hit, unsigned &n_tris) {
    const unsigned payload = ~(unsigned) leaf;
    const unsigned first = payload >> 3, count = (payload & 7u) + 1u;
    const V3 o = mk(r.ox, r.oy, r.oz), d = mk(r.dx, r.dy, r.dz);
    for (unsigned i = 0; i < count; ++i) {
        const float4 a = __ldg(&sc.tris[(size_t) (first + i) * 3 + 0]);
        const float4 b = __ldg(&sc.tris[(size_t) (first + i) * 3 + 1]);
        const float4 c = __ldg(&sc.tris[(size_t) (first + i) * 3 + 2]);
        if (COUNT) n_tris++;
        // Moeller-Trumbore, operation for operation as ref: src/mesh.cpp:39-76
        const V3 p0 = xyz(a);
        const V3 edge1 = xyz(b) - p0, edge2 = xyz(c) - p0;
        const V3 pvec = cross(d, edge2);
        const float det = dot(edge1, pvec);
        if (det > -1e-8f && det < 1e-8f) continue;
        const float inv_det = 1.0f / det;
        const V3 tvec = o - p0;
        const float u = dot(tvec, pvec) * inv_det;
        if (u < 0.0f || u > 1.0f) continue;
        const V3 qvec = cross(tvec, edge1);
        const float v = dot(d, qvec) * inv_det;
        if (v < 0.0f || u + v > 1.0f) continue;
        const float tt = dot(edge2, qvec) * inv_det;
        if (!(tt >= r.mint && tt <= r.maxt)) continue;
        const uint32_t prim = __float_as_uint(a.w);
        if (any_hit) { t.hprim = prim; return true; }
        if (t.hprim == 0xffffffffu || tt < r.maxt || prim > t.hprim) { r.maxt = tt; t.hu = u; t.hv = v; t.hprim = prim; }
    }
    return false;
}

#if NB_WIDE
// One triangle of the wide layout (same 48-byte record, same operation sequence as leaf_test).  Returns true when an any-hit
// query is answered.
__device__ __forceinline__ bool tri_test_one(const SceneDev &sc, unsigned idx, Ray &r, Trav &t, bool any_hit) {
    const float4 a = __ldg(&sc.tris[(size_t) idx * 3 + 0]);
    const float4 b = __ldg(&sc.tris[(size_t) idx * 3 + 1]);
    const float4 c = __ldg(&sc.tris[(size_t) idx * 3 + 2]);
    const V3 o = mk(r.ox, r.oy, r.oz), d = mk(r.dx, r.dy, r.dz);
    const V3 p0 = xyz(a);
    const V3 edge1 = xyz(b) - p0, edge2 = xyz(c) - p0;
    const V3 pvec = cross(d, edge2);
    const float det = dot(edge1, pvec);
    if (det > -1e-8f && det < 1e-8f) return false;
    const float inv_det = 1.0f / det;
    const V3 tvec = o - p0;
    const float u = dot(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    const V3 qvec = cross(tvec, edge1);
    const float v = dot(d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return false;
    const float tt = dot(edge2, qvec) * inv_det;
    if (!(tt >= r.mint && tt <= r.maxt)) return false;
    const uint32_t prim = __float_as_uint(a.w);
    if (any_hit) { t.hprim = prim; return true; }
    if (t.hprim == 0xffffffffu || tt < r.maxt || prim > t.hprim) { r.maxt = tt; t.hu = u; t.hv = v; t.hprim = prim; }
    return false;
}

// The walk over the 8-wide compressed hierarchy (nb_wide.h).  Same contract as the binary trav_run below: closest hit with the
// reference's tie rule or any hit, hit in (t.hprim, t.hu, t.hv, r.maxt); with tail > 0 (wavefront engine) the walk is
// suspended -- its pending groups pushed on the stack, t.node = 1 -- once at most `tail` lanes of the warp are still walking.
template <bool COUNT, bool TMA_TOP>
__device__ __forceinline__ void trav_run(const SceneDev &sc, const float4 *, int, Ray &r, Trav &t,
                                         StackT *stack, bool any_hit, unsigned &n_nodes, unsigned &n_tris, int tail = 0) {
    WideRay R;
    R.idx = t.idx; R.idy = t.idy; R.idz = t.idz; R.ox = r.ox; R.oy = r.oy; R.oz = r.oz;
    R.negx = t.idx < 0.f; R.negy = t.idy < 0.f; R.negz = t.idz < 0.f;
    const uint32_t octinv = 7u - ((R.negx ? 1u : 0u) | (R.negy ? 2u : 0u) | (R.negz ? 4u : 0u));
    R.octinv4 = octinv * 0x01010101u;
    uint2 ng = make_uint2(0u, 0x80000000u), tg = make_uint2(0u, 0u);
    int sp = 0;
    if (t.node != kDone && t.node != 0) { ng = make_uint2(0u, 0u); sp = t.sp; }        // resumed: its groups are on the stack
    bool suspended = false;
    const uint4 *wn = reinterpret_cast<const uint4 *>(sc.nodes);
    for (;;) {
        if (ng.y > 0x00ffffffu) {
            const uint32_t imask = ng.y & 0xffu;
            const int bit = wide_bfind(ng.y);
            const uint32_t base = ng.x;
            ng.y &= ~(1u << bit);
            if (ng.y > 0x00ffffffu) stack[sp++] = ng;
            const uint32_t slot = ((uint32_t) (bit - 24)) ^ octinv;
            const uint32_t ni = base + (uint32_t) __popc(imask & ~(0xffffffffu << slot));
            uint32_t cb, tb, im;
            const uint32_t hm = wide_node_test([&](uint32_t *w) {
                const uint4 *q = wn + (size_t) ni * 5;
#pragma unroll
                for (int k = 0; k < 5; ++k) { const uint4 v = __ldg(q + k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
            }, R, r.mint, r.maxt, cb, tb, im);
            if (COUNT) n_nodes++;
            ng = make_uint2(cb, (hm & 0xff000000u) | im);
            tg = make_uint2(tb, hm & 0x00ffffffu);
        } else {
            tg = ng; ng = make_uint2(0u, 0u);
        }
#if NB_WIDE_POSTPONE
        // postpone the triangles while few lanes hold some and this lane still has nodes to visit
        if (tg.y != 0u && ng.y > 0x00ffffffu && __popc(__ballot_sync(__activemask(), tg.y != 0u)) < NB_WIDE_POSTPONE) { stack[sp++] = tg; tg.y = 0u; }
#endif
        bool answered = false;
        while (tg.y != 0u) {
            const int ti = wide_bfind(tg.y);
            tg.y &= ~(1u << ti);
            if (COUNT) n_tris++;
            if (tri_test_one(sc, tg.x + (unsigned) ti, r, t, any_hit)) { answered = true; break; }
        }
        if (answered) break;
        if (ng.y <= 0x00ffffffu) {
            if (sp > 0) ng = stack[--sp];
            else break;
        }
#if NB_WAVEFRONT
        if (tail > 0 && __popc(__activemask()) <= tail) {
            if (ng.y != 0u) stack[sp++] = ng;          // a node group with hits left, or a popped triangle group
            suspended = true;
            break;
        }
#endif
    }
    if (suspended) { t.node = 1; t.sp = sp; } else { t.node = kDone; t.sp = 0; }
}
#else
template <bool COUNT, bool TMA_TOP>
__device__ __forceinline__ void trav_run(const SceneDev &sc, const float4 *snodes, int smem_nodes, Ray &r, Trav &t,
                                         StackT *stack, bool any_hit, unsigned &n_nodes, unsigned &n_tris, int tail = 0) {
    int node = t.node, sp = t.sp;
#define NB_POP() (sp ? stack[--sp] : kDone)
    int parked = 0;                  // postponed leaf ref (leaf refs are negative; 0 = none)
    bool suspended = false;
#if NB_SPECULATIVE >= 2
    int parked2 = 0;
#endif
    while (node != kDone || parked != 0) {
        // ---- inner nodes
        while (node >= 0 && node != kDone) {
            const float4 n0 = ld_node<TMA_TOP>(sc, snodes, smem_nodes, node, 0);
            const float4 n1 = ld_node<TMA_TOP>(sc, snodes, smem_nodes, node, 1);
            const float4 n2 = ld_node<TMA_TOP>(sc, snodes, smem_nodes, node, 2);
            const float4 n3 = ld_node<TMA_TOP>(sc, snodes, smem_nodes, node, 3);
            if (COUNT) n_nodes++;
            // slab tests (explicit fma: may only cull; boxes are padded by the builder)
            float c0lox = __fmaf_rn(n0.x, t.idx, -t.oodx), c0hix = __fmaf_rn(n0.y, t.idx, -t.oodx);
            float c0loy = __fmaf_rn(n0.z, t.idy, -t.oody), c0hiy = __fmaf_rn(n0.w, t.idy, -t.oody);
            float c0loz = __fmaf_rn(n2.x, t.idz, -t.oodz), c0hiz = __fmaf_rn(n2.y, t.idz, -t.oodz);
            float c1lox = __fmaf_rn(n1.x, t.idx, -t.oodx), c1hix = __fmaf_rn(n1.y, t.idx, -t.oodx);
            float c1loy = __fmaf_rn(n1.z, t.idy, -t.oody), c1hiy = __fmaf_rn(n1.w, t.idy, -t.oody);
            float c1loz = __fmaf_rn(n2.z, t.idz, -t.oodz), c1hiz = __fmaf_rn(n2.w, t.idz, -t.oodz);
            float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), r.mint));
            float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), r.maxt));
            float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), r.mint));
            float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), r.maxt));
            const bool h0 = c0min <= c0max, h1 = c1min <= c1max;
            const int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swap = c1min < c0min;
                node = swap ? r1 : r0;
                stack[sp++] = swap ? r0 : r1;
            } else if (h0 || h1) {
                node = h0 ? r0 : r1;
            } else {
                node = NB_POP();
            }
#if NB_SPECULATIVE
            if (node < 0 && parked == 0) { parked = node; node = NB_POP(); }
#if NB_SPECULATIVE >= 2
            if (node < 0 && parked2 == 0) { parked2 = node; node = NB_POP(); }
            if (__ballot_sync(__activemask(), parked2 == 0) == 0u) break;
#elif NB_SPEC_VOTE
            if (__ballot_sync(__activemask(), parked == 0) == 0u) break;      // every lane still walking holds a leaf
#endif
#endif
        }
        // ---- leaves
#if NB_SPECULATIVE
        if (parked != 0) {
            if (leaf_test<COUNT>(sc, parked, r, t, any_hit, n_tris)) break;
            parked = 0;
        }
#if NB_SPECULATIVE >= 2
        if (parked2 != 0) {
            if (leaf_test<COUNT>(sc, parked2, r, t, any_hit, n_tris)) break;
            parked2 = 0;
        }
#endif
        if (node < 0) { parked = node; node = NB_POP(); }
#else
        if (node == kDone) break;
        if (leaf_test<COUNT>(sc, node, r, t, any_hit, n_tris)) break;
        node = NB_POP();
#endif
#if NB_WAVEFRONT
        // ---- suspension (wavefront engine, nb_wave.cuh): when only a few lanes of the warp are still walking, they keep
        // their (node, stack) and the warp goes on to retire / refill the finished lanes; the walk resumes afterwards.  Checked at the END of an iteration so that every call makes progress (a check on entry livelocks as soon
        // as a wave starts with <= tail lanes).
        if (tail > 0 && (node != kDone || parked != 0) && __popc(__activemask()) <= tail) {
            if (parked != 0) { if (node != kDone) stack[sp++] = node; node = parked; }
            suspended = true;
            break;
        }
#endif
    }
    if (suspended) { t.node = node; t.sp = sp; } else { t.node = kDone; t.sp = 0; }
#undef NB_POP
}
#endif   // NB_WIDE

// Resumable walk (wavefront engine): tr.node == kDone on entry means a fresh ray; otherwise (node, sp, stack) and the
// closest hit so far continue from where a tail cut suspended them.  The slab-test reciprocals are recomputed per call
// so that only (node, sp, hit) stay live across the shading phase.
template <bool COUNT, bool TMA_TOP>
__device__ __forceinline__ void walk_wave(const float4 *nodes, const float4 *tris, const float4 *snodes, int smem_nodes,
                                          Ray &ray, Trav &tr, StackT *stack, bool any_hit, int tail, unsigned &nn, unsigned &nt) {
    SceneDev sc;
    sc.nodes = nodes; sc.tris = tris;
    const int node = tr.node, sp = tr.sp;
    const uint32_t hp = tr.hprim; const float hu = tr.hu, hv = tr.hv;
    Trav t; trav_begin(ray, t);
    if (node != kDone) { t.node = node; t.sp = sp; t.hprim = hp; t.hu = hu; t.hv = hv; }
    trav_run<COUNT, TMA_TOP>(sc, snodes, smem_nodes, ray, t, stack, any_hit, nn, nt, tail);
    tr.node = t.node; tr.sp = t.sp; tr.hprim = t.hprim; tr.hu = t.hu; tr.hv = t.hv;
}

// Entry to the walk (one call per ray).  NB_WALK_NOINLINE=1 compiles it as a real function with its own register
// allocation and traversal stack; measured 4-15 % SLOWER than inlining on B200 (profiles/r1_v6_variant_matrix.txt),
// so the default inlines it.
struct WalkResult { float t, u, v; uint32_t prim; unsigned n_nodes, n_tris; };

#if NB_WALK_NOINLINE
#define NB_WALK_ATTR __noinline__
#else
#define NB_WALK_ATTR __forceinline__
#endif
template <bool COUNT, bool TMA_TOP>
__device__ NB_WALK_ATTR WalkResult walk(const float4 *nodes, const float4 *tris, const float4 *snodes, int smem_nodes,
                                        float ox, float oy, float oz, float mint, float dx, float dy, float dz, float maxt, bool any_hit) {
    StackT stack[kStackN];
    SceneDev sc;
    sc.nodes = nodes; sc.tris = tris;
    Ray r; r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz; r.mint = mint; r.maxt = maxt;
    Trav t; trav_begin(r, t);
    unsigned nn = 0, nt = 0;
    trav_run<COUNT, TMA_TOP>(sc, snodes, smem_nodes, r, t, stack, any_hit, nn, nt);
    WalkResult w; w.t = r.maxt; w.u = t.hu; w.v = t.hv; w.prim = t.hprim; w.n_nodes = nn; w.n_tris = nt;
    return w;
}

// ------------------------------------------------------------------ K3: intersection record (ref: src/accel.cpp:45-96)
struct Its {
    V3 p; float t; float uvx, uvy; Frame sh; int mesh;
};

__device__ __forceinline__ void fill_its(const SceneDev &sc, uint32_t prim, float t, float u, float v, Its &its, V3 *geo_n) {
    const uint4 f = __ldg(&sc.faces[prim]);
    const V3 p0 = xyz(__ldg(&sc.verts[f.x])), p1 = xyz(__ldg(&sc.verts[f.y])), p2 = xyz(__ldg(&sc.verts[f.z]));
    const float b0 = 1 - (u + v), b1 = u, b2 = v;
    its.t = t; its.mesh = (int) f.w;
    its.p = lin3(b0, p0, b1, p1, b2, p2);
    const uint32_t flags = __ldg(&sc.meshes[f.w].flags);
    its.uvx = u; its.uvy = v;
    if (flags & 2u) {
        const float2 t0 = __ldg(&sc.uvs[f.x]), t1 = __ldg(&sc.uvs[f.y]), t2 = __ldg(&sc.uvs[f.z]);
        its.uvx = b0 * t0.x + b1 * t1.x + b2 * t2.x;
        its.uvy = b0 * t0.y + b1 * t1.y + b2 * t2.y;
    }
    if (flags & 1u) {
        const V3 n0 = xyz(__ldg(&sc.normals[f.x])), n1 = xyz(__ldg(&sc.normals[f.y])), n2 = xyz(__ldg(&sc.normals[f.z]));
        its.sh = frame_from_n(normalize(lin3(b0, n0, b1, n1, b2, n2)));
        if (geo_n) *geo_n = normalize(cross(p1 - p0, p2 - p0));
    } else {
        const V3 gn = normalize(cross(p1 - p0, p2 - p0));
        its.sh = frame_from_n(gn);
        if (geo_n) *geo_n = gn;
    }
}

// ------------------------------------------------------------------ K1: camera ray (ref: src/perspective.cpp:76-97, include/nori/transform.h:55-68)
__device__ __forceinline__ void sample_ray(const RenderParams &P, float sx, float sy, Ray &ray) {
    const float *m = P.s2c, *c = P.c2w;
    const float px = sx * P.invW, py = sy * P.invH, pz = 0.0f;
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = ((m[4 * i + 0] * px + m[4 * i + 1] * py) + m[4 * i + 2] * pz) + m[4 * i + 3] * 1.0f;
    const V3 nearP = mk(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
    const V3 d = normalize(nearP);
    const float invZ = 1.0f / d.z;
    const float w = ((c[12] * 0.0f + c[13] * 0.0f) + c[14] * 0.0f) + c[15] * 1.0f;
    ray.ox = (((c[0] * 0.0f + c[1] * 0.0f) + c[2] * 0.0f) + c[3] * 1.0f) / w;
    ray.oy = (((c[4] * 0.0f + c[5] * 0.0f) + c[6] * 0.0f) + c[7] * 1.0f) / w;
    ray.oz = (((c[8] * 0.0f + c[9] * 0.0f) + c[10] * 0.0f) + c[11] * 1.0f) / w;
    ray.dx = c[0] * d.x + (c[1] * d.y + c[2] * d.z);
    ray.dy = c[4] * d.x + (c[5] * d.y + c[6] * d.z);
    ray.dz = c[8] * d.x + (c[9] * d.y + c[10] * d.z);
    ray.mint = P.nearClip * invZ;
    ray.maxt = P.farClip * invZ;
}

// ------------------------------------------------------------------ K5: film splat (ref: src/block.cpp:62-91)
// blocks are stored per owned tile at a fixed block_edge pitch; clipping uses the tile's real size
// (edge tiles are smaller, ref: src/block.cpp:129).  One 128-bit RED per touched pixel.
__device__ __forceinline__ void splat(const RenderParams &P, int tile_slot, int tox, int toy, int tsx, int tsy,
                                      float sx, float sy, V3 value) {
    if (value.x < 0 || !isfinite(value.x) || value.y < 0 || !isfinite(value.y) || value.z < 0 || !isfinite(value.z)) return;
    const int bd = P.border;
    const float posx = sx - 0.5f - (float) (tox - bd), posy = sy - 0.5f - (float) (toy - bd);
    int x0 = (int) ceilf(posx - P.fradius), y0 = (int) ceilf(posy - P.fradius);
    int x1 = (int) floorf(posx + P.fradius), y1 = (int) floorf(posy + P.fradius);
    x0 = max(x0, 0); y0 = max(y0, 0);
    x1 = min(x1, tsx + 2 * bd - 1); y1 = min(y1, tsy + 2 * bd - 1);
    float4 *blk = P.blocks + (size_t) tile_slot * P.block_edge * P.block_edge;
    const int nx = x1 - x0 + 1;
    if (NB_SPLAT_HOIST && nx <= 6) {
        // common case (radius <= 2.5): the column weights are looked up once and kept in registers
        float wxs[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) wxs[i] = i < nx ? P.ftable[(int) (fabsf((float) (x0 + i) - posx) * P.lookup)] : 0.0f;
        for (int y = y0; y <= y1; ++y) {
            const float wy = P.ftable[(int) (fabsf((float) y - posy) * P.lookup)];
            float4 *row = blk + y * P.block_edge + x0;
#pragma unroll
            for (int i = 0; i < 6; ++i)
                if (i < nx) atomicAdd(&row[i], make_float4(value.x * wxs[i] * wy, value.y * wxs[i] * wy, value.z * wxs[i] * wy, 1.0f * wxs[i] * wy));
        }
        return;
    }
    for (int y = y0; y <= y1; ++y) {
        const float wy = P.ftable[(int) (fabsf((float) y - posy) * P.lookup)];
        for (int x = x0; x <= x1; ++x) {
            const float wx = P.ftable[(int) (fabsf((float) x - posx) * P.lookup)];
            float4 add = make_float4(value.x * wx * wy, value.y * wx * wy, value.z * wx * wy, 1.0f * wx * wy);
            atomicAdd(&blk[y * P.block_edge + x], add);
        }
    }
}

#ifndef NB_SPLAT_TILE
#define NB_SPLAT_TILE 0      // A/B: per-warp shared-memory film tile, see splat_tile()
#endif
#if NB_SPLAT_TILE
// Film-splat pre-reduction: a warp works through one work unit (an 8x4 pixel patch x a chunk of samples) at a time, and every
// sample of that patch splats into the same (8+4) x (4+4) pixels.  The warp therefore keeps that rectangle in shared memory
// (96 x float4 = 1.5 KB per warp), adds the samples of its CURRENT patch there (shared-memory float atomics: lanes collide),
// and writes it out with one 128-bit RED per touched pixel when it moves on -- 96 global REDs per unit instead of 16 per
// sample.  Paths that finish after their warp has moved to another patch, and filters with a border other than 2, take the
// global path of splat().
constexpr int kTileW = 12, kTileH = 8;
struct WarpTile { int slot, px0, py0; bool valid; };      // warp-uniform: owned-tile slot and pixel origin of the patch

__device__ __forceinline__ void tile_flush(const RenderParams &P, float4 *wt, WarpTile &T, unsigned lane) {
    if (T.valid) {
        const int tile_id = P.tile_rank + T.slot * P.tile_nranks;
        int tbx, tby; tile_xy(P.tile_tab, tile_id, tbx, tby);
        const int tox = tbx * 32, toy = tby * 32;
        float4 *blk = P.blocks + (size_t) T.slot * P.block_edge * P.block_edge;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = (int) lane + 32 * k;
            const float4 v = wt[i];
            if (v.w != 0.f || v.x != 0.f || v.y != 0.f || v.z != 0.f) {
                // tile pixel (i % 12, i / 12) is block pixel (patch origin - tile origin + i) : the block's border equals the tile's margin
                const int bx = T.px0 - tox + i % kTileW, by = T.py0 - toy + i / kTileW;
                atomicAdd(&blk[by * P.block_edge + bx], v);
                wt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    __syncwarp();
}

// true if the sample was added to the warp's tile
__device__ __forceinline__ bool splat_tile(const RenderParams &P, float4 *wt, const WarpTile &T, int tile_slot, int tsx, int tsy,
                                           float sx, float sy, V3 value) {
    const int px = (int) sx, py = (int) sy;
    if (!T.valid || P.border != 2 || tile_slot != T.slot || px < T.px0 || px >= T.px0 + 8 || py < T.py0 || py >= T.py0 + 4) return false;
    if (value.x < 0 || !isfinite(value.x) || value.y < 0 || !isfinite(value.y) || value.z < 0 || !isfinite(value.z)) return true;
    const int tile_id = P.tile_rank + tile_slot * P.tile_nranks;
    int tbx, tby; tile_xy(P.tile_tab, tile_id, tbx, tby);
    const int tox = tbx * 32, toy = tby * 32;
    const float posx = sx - 0.5f - (float) (tox - 2), posy = sy - 0.5f - (float) (toy - 2);       // block coordinates, as in splat()
    int x0 = (int) ceilf(posx - P.fradius), y0 = (int) ceilf(posy - P.fradius);
    int x1 = (int) floorf(posx + P.fradius), y1 = (int) floorf(posy + P.fradius);
    x0 = max(x0, 0); y0 = max(y0, 0);
    x1 = min(x1, tsx + 3); y1 = min(y1, tsy + 3);
    const int ox = T.px0 - tox, oy = T.py0 - toy;                                                  // block coordinates of tile pixel (0, 0)
    for (int y = y0; y <= y1; ++y) {
        const float wy = P.ftable[(int) (fabsf((float) y - posy) * P.lookup)];
        for (int x = x0; x <= x1; ++x) {
            const float wx = P.ftable[(int) (fabsf((float) x - posx) * P.lookup)];
            float *t = reinterpret_cast<float *>(&wt[(y - oy) * kTileW + (x - ox)]);
            atomicAdd(t + 0, value.x * wx * wy); atomicAdd(t + 1, value.y * wx * wy); atomicAdd(t + 2, value.z * wx * wy); atomicAdd(t + 3, 1.0f * wx * wy);
        }
    }
    return true;
}
#endif

#if NB_WAVEFRONT
// Wavefront engine: one thread per pool slot (nb_wave.cuh: wf_logic_kernel).  An occlusion ray goes to the shadow queue with
// the radiance it would add to the path and the slot it belongs to; wf_trace_kernel adds it to the slot's L if unoccluded.
// A path emits at most one such ray per iteration, so the queue (one entry per slot) cannot overflow.
__device__ __forceinline__ bool occ_push(const RenderParams &P, const Ray &r, V3 contrib, float, float, int) {
    const unsigned mask = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if ((int) lane == leader) base = atomicAdd(&P.wf_ctr[3], (unsigned) __popc(mask));     // WF_SHADOW_COUNT
    base = __shfl_sync(mask, base, leader);
    const unsigned idx = base + (unsigned) __popc(mask & ((1u << lane) - 1u));
    float4 *q = P.occ_queue + (size_t) idx * 3u;
    q[0] = make_float4(r.ox, r.oy, r.oz, r.mint);
    q[1] = make_float4(r.dx, r.dy, r.dz, r.maxt);
    q[2] = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(blockIdx.x * blockDim.x + threadIdx.x));
    return true;
}
#endif

// ------------------------------------------------------------------ emitter sampling [authored]; DiscretePDF::sample ref: include/nori/dpdf.h:93-99
__device__ __forceinline__ uint32_t cdf_sample(const float *cdf, uint32_t n, float x) {
    uint32_t lo = 0, hi = n + 1;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (__ldg(&cdf[mid]) < x) lo = mid + 1; else hi = mid; }
    long long idx = (long long) lo - 1; if (idx < 0) idx = 0;
    if (idx > (long long) n - 1) idx = (long long) n - 1;
    return (uint32_t) idx;
}

struct EmitSample { V3 y, n; float pdfA; V3 Le; };

__device__ __noinline__ void sample_emitter(const SceneDev &sc, float xe, float xt, float xa, float xb, EmitSample &es) {
    int k = (int) (xe * (float) sc.n_emitters); if (k > sc.n_emitters - 1) k = sc.n_emitters - 1;
    const int mi = __ldg(&sc.emitters[k]);
    const DevMesh &m = sc.meshes[mi];
    const uint32_t f = cdf_sample(sc.emitter_cdf + m.cdf_offset, m.nf, xt);
    const uint4 fi = __ldg(&sc.faces[m.prim_offset + f]);
    const V3 p0 = xyz(__ldg(&sc.verts[fi.x])), p1 = xyz(__ldg(&sc.verts[fi.y])), p2 = xyz(__ldg(&sc.verts[fi.z]));
    const float su = sqrtf(1.0f - xa);
    const float b0 = 1.0f - su, b1 = xb * su; const float b2 = 1.0f - b0 - b1;
    es.y = lin3(b0, p0, b1, p1, b2, p2);
    if (m.flags & 1u) {
        const V3 n0 = xyz(__ldg(&sc.normals[fi.x])), n1 = xyz(__ldg(&sc.normals[fi.y])), n2 = xyz(__ldg(&sc.normals[fi.z]));
        es.n = normalize(lin3(b0, n0, b1, n1, b2, n2));
    } else {
        es.n = normalize(cross(p1 - p0, p2 - p0));
    }
    es.pdfA = 1.0f / (m.area_sum * (float) sc.n_emitters);
    es.Le = mk(m.radiance[0], m.radiance[1], m.radiance[2]);
}

// ------------------------------------------------------------------ K4: per-path state machine
enum { ST_IDLE = 0, ST_EXTEND = 1, ST_SHADOW = 2, ST_SHADOW_AO = 3 };

struct Path {
    Pcg32 rng;
    float sx, sy;            // film position of the camera sample
    V3 L, T;                 // radiance, throughput
    V3 contrib;              // pending next-event contribution (added if the shadow ray is unoccluded)
    V3 next_d;               // direction of the extension ray that follows the shadow ray
    float prev_pdf;
    int depth;
    int tile_slot;           // owned-tile index of the sample (film block)
    short tox, toy;          // tile origin
    unsigned char tsx, tsy;  // tile size
    unsigned char stage;
    bool prev_specular, has_next;
#if NB_WAVEFRONT
    unsigned deferred;       // occlusion rays this lane handed to the queue (they count as traced rays)
#endif
};

// One shading step after a ray finished.  Returns true when the path is complete (L final).
// Sampler draw order is part of the spec and identical to oracle.c:Li.
template <int INTEG>
__device__ __forceinline__ bool shade(const RenderParams &P, Path &ps, Ray &ray, const Trav &tr, unsigned &n_hits) {
    const SceneDev &sc = P.sc;
    if (ps.stage == ST_SHADOW_AO) {
        if (tr.hprim == 0xffffffffu) ps.L = mk(1, 1, 1);
        return true;
    }
    if (ps.stage == ST_SHADOW) {
        if (tr.hprim == 0xffffffffu) ps.L = ps.L + ps.contrib;
        if (!ps.has_next) return true;
        ray.dx = ps.next_d.x; ray.dy = ps.next_d.y; ray.dz = ps.next_d.z; ray.mint = NB_EPSILON; ray.maxt = NB_INF;
        ps.stage = ST_EXTEND;
        return false;
    }
    // ---- ST_EXTEND: closest hit finished
    if (tr.hprim == 0xffffffffu) return true;
    Its its;
    fill_its(sc, tr.hprim, ray.maxt, tr.hu, tr.hv, its, nullptr);
    n_hits++;
    if (INTEG == 0) {                                   // normals
        ps.L = mk(fabsf(its.sh.n.x), fabsf(its.sh.n.y), fabsf(its.sh.n.z));
        return true;
    }
    if (INTEG == 1) {                                   // ao
        const float x = pcg_next_float(ps.rng), y = pcg_next_float(ps.rng);
        const V3 w = to_world(its.sh, square_to_cosine_hemisphere(x, y));
        ray.ox = its.p.x; ray.oy = its.p.y; ray.oz = its.p.z; ray.dx = w.x; ray.dy = w.y; ray.dz = w.z;
        ray.mint = NB_EPSILON; ray.maxt = NB_INF;
#if NB_WAVEFRONT
        if (occ_push(P, ray, mk(1, 1, 1), ps.sx, ps.sy, ps.tile_slot)) { ps.deferred++; return true; }   // L stays 0; visibility arrives through the queue
#endif
        ps.stage = ST_SHADOW_AO;
        return false;
    }
    if (INTEG == 6) {                                   // simple: one point light (oracle.c: ORC_INT_SIMPLE, same operation order)
        const V3 dvec = mk(P.light_pos[0], P.light_pos[1], P.light_pos[2]) - its.p;
        const float dist2 = dot(dvec, dvec);
        const float dist = sqrtf(dist2);
        const V3 wo_w = mk(dvec.x / dist, dvec.y / dist, dvec.z / dist);
        const float cosT = dot(its.sh.n, wo_w);
        if (!(cosT > 0.0f)) return true;
        const float g = cosT / dist2 * 0.025330295910584444f;      // 1 / (4 pi^2)
        ps.contrib = mk(P.light_energy[0] * g, P.light_energy[1] * g, P.light_energy[2] * g);
        ps.has_next = false;
        ray.ox = its.p.x; ray.oy = its.p.y; ray.oz = its.p.z; ray.dx = wo_w.x; ray.dy = wo_w.y; ray.dz = wo_w.z;
        ray.mint = NB_EPSILON; ray.maxt = dist - NB_EPSILON;
#if NB_WAVEFRONT
        if (occ_push(P, ray, ps.contrib, ps.sx, ps.sy, ps.tile_slot)) { ps.deferred++; return true; }
#endif
        ps.stage = ST_SHADOW;                           // resolved by the ST_SHADOW branch above: L += contrib if unoccluded
        return false;
    }
    const DevMesh &m = sc.meshes[its.mesh];
    const V3 wi = to_local(its.sh, neg(mk(ray.dx, ray.dy, ray.dz)));
    const bool diffuse = bsdf_is_diffuse(m);

    if (m.emitter_type == 1 && wi.z > 0.0f) {           // emitted radiance
        const V3 Le = mk(m.radiance[0], m.radiance[1], m.radiance[2]);
        float w = 1.0f;
        bool add;
        if (INTEG == 2) add = diffuse;
        else if (INTEG == 3) add = true;
        else if (INTEG == 4) add = ps.prev_specular;
        else {
            add = true;
            if (!ps.prev_specular) {
                const float pdfA = 1.0f / (m.area_sum * (float) sc.n_emitters);
                const float pdf_em = pdfA * (its.t * its.t) / wi.z;
                w = ps.prev_pdf / (ps.prev_pdf + pdf_em);
            }
        }
        if (add) { ps.L.x += ps.T.x * Le.x * w; ps.L.y += ps.T.y * Le.y * w; ps.L.z += ps.T.z * Le.z * w; }
    }

    ray.ox = its.p.x; ray.oy = its.p.y; ray.oz = its.p.z;

    if (INTEG == 2 && !diffuse) {                       // whitted: specular chain with 0.95 continuation
        const float x = pcg_next_float(ps.rng);
        if (x >= 0.95f) return true;
        const float sx = pcg_next_float(ps.rng), sy = pcg_next_float(ps.rng);
        V3 wo; int measure; float spdf;
        const V3 f = bsdf_sample(m, wi, sx, sy, wo, measure, spdf);
        if (is_zero(f)) return true;
        ps.T = mk(ps.T.x * f.x / 0.95f, ps.T.y * f.y / 0.95f, ps.T.z * f.z / 0.95f);
        const V3 d = to_world(its.sh, wo);
        ray.dx = d.x; ray.dy = d.y; ray.dz = d.z; ray.mint = NB_EPSILON; ray.maxt = NB_INF;
        ps.depth++;
        return ps.depth >= P.max_depth;
    }

    if (INTEG != 2 && ps.depth >= P.rr_start) {          // Russian roulette
        float q = max3(ps.T); if (q > 0.99f) q = 0.99f;
        const float x = pcg_next_float(ps.rng);
        if (x >= q) return true;
        ps.T = mk(ps.T.x / q, ps.T.y / q, ps.T.z / q);
    }

    bool want_shadow = false;
    Ray sray = ray;
    if ((INTEG == 2 || INTEG == 4 || INTEG == 5) && diffuse && sc.n_emitters > 0) {   // next-event estimation
        const float xe = pcg_next_float(ps.rng), xt = pcg_next_float(ps.rng);
        const float xa = pcg_next_float(ps.rng), xb = pcg_next_float(ps.rng);
        EmitSample es; sample_emitter(sc, xe, xt, xa, xb, es);
        const V3 dvec = es.y - its.p;
        const float dist2 = dot(dvec, dvec);
        const float dist = sqrtf(dist2);
        const V3 wo_w = mk(dvec.x / dist, dvec.y / dist, dvec.z / dist);
        const float cosL = -dot(es.n, wo_w);
        if (cosL > 0.0f) {
            const V3 wo = to_local(its.sh, wo_w);
            float bpdf;
            const V3 f = bsdf_eval_pdf(m, wi, wo, bpdf);
            if (!is_zero(f)) {
                const float pdf_sa = es.pdfA * dist2 / cosL;
                float w = 1.0f;
                if (INTEG == 5) w = pdf_sa / (pdf_sa + bpdf);
                const float g = wo.z / pdf_sa * w;
                ps.contrib = mk(ps.T.x * f.x * es.Le.x * g, ps.T.y * f.y * es.Le.y * g, ps.T.z * f.z * es.Le.z * g);
                sray.dx = wo_w.x; sray.dy = wo_w.y; sray.dz = wo_w.z; sray.mint = NB_EPSILON; sray.maxt = dist - NB_EPSILON;
                want_shadow = true;
            }
        }
    }
    ps.has_next = false;
    if (INTEG != 2) {                                   // BSDF sampling -> extension ray
        const float sx = pcg_next_float(ps.rng), sy = pcg_next_float(ps.rng);
        V3 wo; int measure; float spdf;
        const V3 f = bsdf_sample(m, wi, sx, sy, wo, measure, spdf);
        if (!is_zero(f)) {
            ps.T = ps.T * f;
            ps.prev_specular = (measure == 2);
            ps.prev_pdf = ps.prev_specular ? 0.0f : spdf;            // == bsdf_pdf(wi, wo) of the sampled direction
            ps.next_d = to_world(its.sh, wo);
            ps.depth++;
            ps.has_next = ps.depth < P.max_depth;
        }
    }
#if NB_WAVEFRONT
    // the shadow ray goes to the queue with its contribution; the path carries on with its extension ray at once
    if (want_shadow && occ_push(P, sray, ps.contrib, ps.sx, ps.sy, ps.tile_slot)) { ps.deferred++; want_shadow = false; }
#endif
    if (want_shadow) { ray = sray; ps.stage = ST_SHADOW; return false; }
    if (!ps.has_next) return true;
    ray.dx = ps.next_d.x; ray.dy = ps.next_d.y; ray.dz = ps.next_d.z; ray.mint = NB_EPSILON; ray.maxt = NB_INF;
    ps.stage = ST_EXTEND;
    return false;
}

// Start the camera path of work item (pixel, sample): ref src/main.cpp:41-46
__device__ __forceinline__ void begin_path(const RenderParams &P, Path &ps, Ray &ray, int px, int py, uint32_t sample) {
    if (P.seed_mode == 0) {
        const uint64_t pix = (uint64_t) py * (uint64_t) P.W + (uint64_t) px;
        pcg_seed(ps.rng, (P.seed << 32) + pix, (uint64_t) sample + (uint64_t) P.sample_offset);
    } else if (P.block_stream_skip) {
        // reference per-block stream (ref: src/independent.cpp:36-41) for integrators with a FIXED number of draws per
        // sample (normals: the 4 camera draws): sample j of the block sits at stream position 4*j, reached by skip-ahead
        const int tox = (px / 32) * 32, toy = (py / 32) * 32;
        const int tsx = min(32, P.W - tox);
        const uint64_t j = ((uint64_t) (py - toy) * (uint64_t) tsx + (uint64_t) (px - tox)) * (uint64_t) P.spp + (uint64_t) sample;
        pcg_seed(ps.rng, (uint64_t) tox, (uint64_t) toy);
        pcg_advance(ps.rng, 4ull * j);
    }
    ps.sx = (float) px + pcg_next_float(ps.rng);
    ps.sy = (float) py + pcg_next_float(ps.rng);
    pcg_next_float(ps.rng); pcg_next_float(ps.rng);          // apertureSample (consumed, unused by the pinhole)
    sample_ray(P, ps.sx, ps.sy, ray);
    ps.L = mk(0, 0, 0); ps.T = mk(1, 1, 1);
    ps.depth = 0; ps.prev_specular = true; ps.prev_pdf = 0.0f; ps.has_next = false;
    ps.stage = ST_EXTEND;
}

// ------------------------------------------------------------------ TMA bulk copy of the BVH top into shared memory
__device__ __forceinline__ void tma_stage_nodes(float4 *snodes, const float4 *gnodes, int n_nodes, uint64_t *mbar) {
    const uint32_t bytes = (uint32_t) n_nodes * 64u;
    const uint32_t mbar_a = (uint32_t) __cvta_generic_to_shared(mbar);
    const uint32_t dst_a = (uint32_t) __cvta_generic_to_shared(snodes);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_a), "l"(gnodes), "r"(bytes), "r"(mbar_a) : "memory");
    }
    // everyone waits for phase 0
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(mbar_a) : "memory");
    }
}

// ------------------------------------------------------------------ L2 warm-up of the scene arrays
// A frame starts with a cold L2 (another frame's blocks, the film, or the bench's flush went through it), and a walk that
// demand-misses its first nodes to HBM pays the DRAM latency once per level.  One bulk prefetch instruction per 4 KB chunk
// (cp.async.bulk.prefetch.L2, Hopper+) pulls the node and triangle arrays into L2 at full HBM speed -- 57 MB in ~10 us --
// while the render kernel is being launched behind it.  Matters most when the frame is short (8 GPUs: 1.3 ms per frame).
__global__ void l2_prefetch_kernel(const char *a, unsigned long long a_bytes, const char *b, unsigned long long b_bytes) {
    constexpr unsigned long long kChunk = 4096ull;
    const unsigned long long na = (a_bytes + kChunk - 1) / kChunk, nb_ = (b_bytes + kChunk - 1) / kChunk;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < na + nb_;
         i += (unsigned long long) gridDim.x * blockDim.x) {
        const bool first = i < na;
        const unsigned long long off = (first ? i : i - na) * kChunk;
        const unsigned long long total = first ? a_bytes : b_bytes;
        const char *p = (first ? a : b) + off;
        const unsigned bytes = (unsigned) ((total - off < kChunk ? total - off : kChunk) & ~15ull);    // multiple of 16 B
        if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
    }
}

// ------------------------------------------------------------------ the fused persistent kernel (K1..K5)
template <int INTEG, bool COUNT, bool TMA_TOP>
__global__ void __launch_bounds__(128, (INTEG <= 1 || INTEG == 6) ? NB_MIN_BLOCKS : NB_MIN_BLOCKS_PATH) render_kernel(const __grid_constant__ RenderParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *snodes = reinterpret_cast<float4 *>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    const int smem_nodes = TMA_TOP ? P.smem_nodes : 0;
    if (TMA_TOP && smem_nodes > 0) tma_stage_nodes(snodes, P.sc.nodes, smem_nodes, &mbar);

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
#if NB_SPLAT_TILE
    __shared__ float4 wtiles[4][kTileW * kTileH];
    float4 *wt = wtiles[threadIdx.x >> 5];
    for (int k = 0; k < 3; ++k) wt[lane + 32 * k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    WarpTile WT; WT.valid = false; WT.slot = 0; WT.px0 = 0; WT.py0 = 0;
#endif
    Path ps; Ray ray; Trav tr;
    ps.stage = ST_IDLE; tr.node = kDone; tr.sp = 0; tr.hprim = 0xffffffffu;
    unsigned n_rays = 0, n_nodes = 0, n_tris = 0, n_hits = 0;
    unsigned wave_nodes = 0; unsigned long long wave_max_sum = 0, n_waves = 0;

    // warp-uniform work-unit state
    bool exhausted = false, traced = false;
    uint32_t next_item = 0, n_items = 0, valid_mask = 0, n_valid = 0, sample_base = 0;
    int u_tile_slot = 0, u_tox = 0, u_toy = 0, u_tsx = 0, u_tsy = 0, u_px0 = 0, u_py0 = 0;

    for (;;) {
        // ---- shading phase (lock step: every lane's ray is finished here)
        if (ps.stage != ST_IDLE && traced) {
            const bool finished = shade<INTEG>(P, ps, ray, tr, n_hits);
            if (finished) {
#if NB_SPLAT_TILE
                if (!splat_tile(P, wt, WT, ps.tile_slot, ps.tsx, ps.tsy, ps.sx, ps.sy, ps.L))
#endif
                splat(P, ps.tile_slot, ps.tox, ps.toy, ps.tsx, ps.tsy, ps.sx, ps.sy, ps.L);
                ps.stage = ST_IDLE;
            } else {
                n_rays++;
            }
        }
        // ---- regeneration: free lanes take the next items of the warp's unit (ballot/popc compaction)
        bool need = (ps.stage == ST_IDLE);
        unsigned need_mask = __ballot_sync(0xffffffffu, need);
        while (need_mask != 0u && !exhausted) {
            if (next_item >= n_items) {
                uint32_t u = 0;
                if (lane == 0) u = (uint32_t) atomicAdd(&P.counters[0], 1ULL);
                u = __shfl_sync(0xffffffffu, u, 0);
                if (u >= P.n_units) { exhausted = true; break; }
                // unit -> (owned tile, 8x4 patch, sample chunk); patches vary fastest (concurrent warps splat into different pixels).
                // Guided schedule: the first split_units units are COARSE (chunk_a samples of a patch: the warp stays on its 32
                // pixels and its walks share their nodes in L1), the rest FINE (chunk samples): the frame ends on small units.
                uint32_t chunk = P.chunk, nchunks = P.nchunks, sample0 = P.split_sample;
                if (u < P.split_units) { chunk = P.chunk_a; nchunks = P.nchunks_a; sample0 = 0u; }
                else u -= P.split_units;
                const uint32_t patch = u % 32u;
                const uint32_t rest = u / 32u;
                const uint32_t chunk_id = rest % nchunks;
                u_tile_slot = (int) (rest / nchunks);
                const int tile_id = P.tile_rank + u_tile_slot * P.tile_nranks;
                int bx, by; tile_xy(P.tile_tab, tile_id, bx, by);
                u_tox = bx * 32; u_toy = by * 32;
                u_tsx = min(32, P.W - u_tox); u_tsy = min(32, P.H - u_toy);
                u_px0 = u_tox + (int) (patch & 3u) * 8; u_py0 = u_toy + (int) (patch >> 2) * 4;
#if NB_SPLAT_TILE
                if (!WT.valid || WT.slot != u_tile_slot || WT.px0 != u_px0 || WT.py0 != u_py0) {   // another patch (a new chunk of the same patch keeps the tile)
                    __syncwarp();
                    tile_flush(P, wt, WT, lane);
                    WT.valid = true; WT.slot = u_tile_slot; WT.px0 = u_px0; WT.py0 = u_py0;
                }
#endif
                const int lx = u_px0 + (int) (lane & 7u), ly = u_py0 + (int) (lane >> 3);
                valid_mask = __ballot_sync(0xffffffffu, lx < P.W && ly < P.H);
                n_valid = __popc(valid_mask);
                sample_base = sample0 + chunk_id * chunk;
                const uint32_t ns = min(chunk, P.spp - sample_base);
                n_items = n_valid * ns; next_item = 0;
                continue;
            }
            const uint32_t avail = n_items - next_item;
            const uint32_t rank = __popc(need_mask & lt_mask);
            if (need && rank < avail) {
                const uint32_t item = next_item + rank;
                const uint32_t pix_slot = item % n_valid, s = sample_base + item / n_valid;
                const int pl = __fns(valid_mask, 0, pix_slot + 1);      // lane index of the pix_slot-th valid pixel
                ps.tile_slot = u_tile_slot;
                ps.tox = (short) u_tox; ps.toy = (short) u_toy;
                ps.tsx = (unsigned char) u_tsx; ps.tsy = (unsigned char) u_tsy;
                begin_path(P, ps, ray, u_px0 + (pl & 7), u_py0 + (pl >> 3), s);
                n_rays++;
                need = false;
            }
            next_item += min((uint32_t) __popc(need_mask), avail);
            need_mask = __ballot_sync(0xffffffffu, need);
        }
        if (__ballot_sync(0xffffffffu, ps.stage != ST_IDLE) == 0u) break;     // work exhausted and every path retired
        // ---- traversal phase: every active lane walks its ray to completion.  Path integrators run PHASED waves
        // (NB_PHASED_WAVES): if any lane holds a shadow ray the wave traces shadow rays only, else extension rays only;
        // lanes holding the other kind keep their ray and wait.  Shading after a wave is then stage-homogeneous
        // (all NEE-resolve or all hit-shade) instead of a 50/50 mix, and any-hit waves are not held up by closest-hit.
        traced = (ps.stage != ST_IDLE);
        if (NB_PHASED_WAVES && (INTEG >= 2 || (NB_PHASED_AO && INTEG == 1))) {
            const bool any_shadow = __ballot_sync(0xffffffffu, ps.stage >= ST_SHADOW) != 0u;
            traced = (any_shadow ? (ps.stage >= ST_SHADOW) : (ps.stage == ST_EXTEND));
        }
        if (traced) {
            const WalkResult w = walk<COUNT, TMA_TOP>(P.sc.nodes, P.sc.tris, snodes, smem_nodes, ray.ox, ray.oy, ray.oz, ray.mint,
                                                      ray.dx, ray.dy, ray.dz, ray.maxt, ps.stage != ST_EXTEND);
            ray.maxt = w.t; tr.hu = w.u; tr.hv = w.v; tr.hprim = w.prim;
            if (COUNT) { n_nodes += w.n_nodes; n_tris += w.n_tris; wave_nodes = w.n_nodes; }
        }
        if (COUNT) {
            // lock-step diagnostics: per wave, the longest walk (what the warp pays) vs the sum over lanes (what it needs)
            unsigned mx = wave_nodes;
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) { wave_max_sum += mx; n_waves++; }
            wave_nodes = 0;
        }
    }

#if NB_SPLAT_TILE
    __syncwarp();
    tile_flush(P, wt, WT, lane);
#endif
    // counters: warp-reduce then one atomic per warp
    unsigned long long v1 = n_rays, v2 = n_nodes, v3 = n_tris, v4 = n_hits;
    for (int o = 16; o > 0; o >>= 1) {
        v1 += __shfl_down_sync(0xffffffffu, v1, o); v2 += __shfl_down_sync(0xffffffffu, v2, o);
        v3 += __shfl_down_sync(0xffffffffu, v3, o); v4 += __shfl_down_sync(0xffffffffu, v4, o);
    }
    if (lane == 0) {
        atomicAdd(&P.counters[1], v1);
        if (COUNT) { atomicAdd(&P.counters[2], v2); atomicAdd(&P.counters[3], v3); atomicAdd(&P.counters[5], wave_max_sum); atomicAdd(&P.counters[6], n_waves); }
        atomicAdd(&P.counters[4], v4);
    }
}

// Reference seeding mode (Independent::prepare, ref: src/independent.cpp:36-41): one SEQUENTIAL pcg32 stream per
// 32x32 block, consumed pixel by pixel, sample by sample (ref: src/main.cpp:38-53).  One thread per block;
// this is the plumbing configuration (BASELINE configs[0]), not the throughput path.
template <int INTEG, bool COUNT>
__global__ void __launch_bounds__(32) render_block_mode_kernel(const __grid_constant__ RenderParams P) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.n_my_tiles) return;
    StackT stack[kStackN];
    Path ps; Ray ray; Trav tr;
    unsigned n_rays = 0, n_nodes = 0, n_tris = 0, n_hits = 0;
    const int tile_id = P.tile_rank + slot * P.tile_nranks;
    int bx, by; tile_xy(P.tile_tab, tile_id, bx, by);
    const int tox = bx * 32, toy = by * 32, tsx = min(32, P.W - tox), tsy = min(32, P.H - toy);
    pcg_seed(ps.rng, (uint64_t) tox, (uint64_t) toy);
    ps.tile_slot = slot; ps.tox = (short) tox; ps.toy = (short) toy; ps.tsx = (unsigned char) tsx; ps.tsy = (unsigned char) tsy;
    for (int y = 0; y < tsy; ++y) for (int x = 0; x < tsx; ++x) for (uint32_t i = 0; i < P.spp; ++i) {
        begin_path(P, ps, ray, tox + x, toy + y, i);
        for (;;) {
            trav_begin(ray, tr); n_rays++;
            trav_run<COUNT, false>(P.sc, nullptr, 0, ray, tr, stack, ps.stage != ST_EXTEND, n_nodes, n_tris);
            if (shade<INTEG>(P, ps, ray, tr, n_hits)) break;
        }
        splat(P, slot, tox, toy, tsx, tsy, ps.sx, ps.sy, ps.L);
    }
    atomicAdd(&P.counters[1], (unsigned long long) n_rays);
    if (COUNT) { atomicAdd(&P.counters[2], (unsigned long long) n_nodes); atomicAdd(&P.counters[3], (unsigned long long) n_tris); }
    atomicAdd(&P.counters[4], (unsigned long long) n_hits);
}

// Li of n independent camera paths, the device counterpart of the loop of the reference's t-test in scene mode
// (ref: src/ttest.cpp:153-167): pixelSample = next2D() * outputSize, apertureSample = next2D(), value = Li, and the
// luminance of value (ref: src/common.cpp:206-208) is what the test accumulates.  The reference walks ONE sequential
// sampler stream through all paths, which cannot be split; here path k owns the stream seed((seed << 32) + k, 0)
// (oracle.c: orc_li_samples).  One thread per path: a test utility, not a throughput path.
template <int INTEG>
__global__ void __launch_bounds__(128) li_samples_kernel(const __grid_constant__ RenderParams P, unsigned long long n, float *lum) {
    StackT stack[kStackN];
    Path ps; Ray ray; Trav tr;
    unsigned n_rays = 0, n_nodes = 0, n_tris = 0, n_hits = 0;
    for (unsigned long long k = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (unsigned long long) gridDim.x * blockDim.x) {
        pcg_seed(ps.rng, (P.seed << 32) + k, 0ull);
        ps.sx = pcg_next_float(ps.rng) * (float) P.W;
        ps.sy = pcg_next_float(ps.rng) * (float) P.H;
        pcg_next_float(ps.rng); pcg_next_float(ps.rng);
        sample_ray(P, ps.sx, ps.sy, ray);
        ps.L = mk(0, 0, 0); ps.T = mk(1, 1, 1);
        ps.depth = 0; ps.prev_specular = true; ps.prev_pdf = 0.0f; ps.has_next = false;
        ps.stage = ST_EXTEND;
        for (;;) {
            trav_begin(ray, tr); n_rays++;
            trav_run<false, false>(P.sc, nullptr, 0, ray, tr, stack, ps.stage != ST_EXTEND, n_nodes, n_tris);
            if (shade<INTEG>(P, ps, ray, tr, n_hits)) break;
        }
        lum[k] = ps.L.x * 0.212671f + ps.L.y * 0.715160f + ps.L.z * 0.072169f;
    }
    atomicAdd(&P.counters[1], (unsigned long long) n_rays);
    atomicAdd(&P.counters[4], (unsigned long long) n_hits);
}


// ------------------------------------------------------------------ K6: merge finished blocks into the full film (ref: src/block.cpp:93-102)
__global__ void merge_blocks_kernel(const float4 *blocks, int n_tiles, int rank, int nranks, const uint32_t *tile_tab, int W, int H,
                                    int border, int block_edge, float4 *film) {
    const int per_block = block_edge * block_edge;
    const long long gid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long) n_tiles * per_block) return;
    const int slot = (int) (gid / per_block), r = (int) (gid % per_block);
    const int y = r / block_edge, x = r % block_edge;
    const int tile_id = rank + slot * nranks;
    int bx, by; tile_xy(tile_tab, tile_id, bx, by);
    const int tox = bx * 32, toy = by * 32;
    const int tsx = min(32, W - tox), tsy = min(32, H - toy);
    if (x >= tsx + 2 * border || y >= tsy + 2 * border) return;
    const float4 v = blocks[gid];
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) return;
    const int cols = W + 2 * border;
    atomicAdd(&film[(size_t) (toy + y) * cols + (tox + x)], v);
}

// Same merge for the gathered blocks of ALL ranks in one launch: blocks = [nranks][stride_tiles][edge][edge] (ranks
// padded to stride_tiles); slot k of rank r is tile r + k * nranks.
__global__ void merge_all_blocks_kernel(const float4 *blocks, int nranks, int stride_tiles, int total_tiles, const uint32_t *tile_tab, int W, int H,
                                        int border, int block_edge, float4 *film) {
    const int per_block = block_edge * block_edge;
    const long long gid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long) nranks * stride_tiles * per_block) return;
    const int r = (int) (gid % per_block);
    const int slot_all = (int) (gid / per_block);
    const int rank = slot_all / stride_tiles, slot = slot_all % stride_tiles;
    const int tile_id = rank + slot * nranks;
    if (tile_id >= total_tiles) return;
    const int y = r / block_edge, x = r % block_edge;
    int bx, by; tile_xy(tile_tab, tile_id, bx, by);
    const int tox = bx * 32, toy = by * 32;
    const int tsx = min(32, W - tox), tsy = min(32, H - toy);
    if (x >= tsx + 2 * border || y >= tsy + 2 * border) return;
    const float4 v = blocks[gid];
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) return;
    atomicAdd(&film[(size_t) (toy + y) * (W + 2 * border) + (tox + x)], v);
}

// ------------------------------------------------------------------ batched Scene::rayIntersect (ref: include/nori/scene.h:63-85)
struct HitOut { float t, u, v; uint32_t prim; uint32_t mesh; };

template <bool COUNT>
__global__ void __launch_bounds__(128) intersect_kernel(SceneDev sc, const float4 *rays, unsigned long long n, HitOut *hits,
                                                        int shadow, float *full16, unsigned long long *counters) {
    StackT stack[kStackN];
    unsigned n_nodes = 0, n_tris = 0, n_rays = 0;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long) gridDim.x * blockDim.x) {
        const float4 a = __ldg(&rays[2 * i]), b = __ldg(&rays[2 * i + 1]);
        Ray r; r.ox = a.x; r.oy = a.y; r.oz = a.z; r.mint = a.w; r.dx = b.x; r.dy = b.y; r.dz = b.z; r.maxt = b.w;
        Trav t; trav_begin(r, t); n_rays++;
        trav_run<COUNT, false>(sc, nullptr, 0, r, t, stack, shadow != 0, n_nodes, n_tris);
        if (hits) {
            HitOut h;
            if (t.hprim != 0xffffffffu && !shadow) { h.t = r.maxt; h.u = t.hu; h.v = t.hv; h.prim = t.hprim; h.mesh = __ldg(&sc.faces[t.hprim]).w; }
            else { h.t = 0.f; h.u = 0.f; h.v = 0.f; h.prim = (t.hprim != 0xffffffffu) ? 0u : 0xffffffffu; h.mesh = h.prim; }
            hits[i] = h;
        }
        if (full16) {
            float *o = full16 + 16 * i;
            if (t.hprim == 0xffffffffu) { for (int k = 0; k < 16; ++k) o[k] = 0.f; o[15] = -1.0f; }
            else {
                Its its; fill_its(sc, t.hprim, r.maxt, t.hu, t.hv, its, nullptr);
                o[0] = its.p.x; o[1] = its.p.y; o[2] = its.p.z; o[3] = its.t; o[4] = its.uvx; o[5] = its.uvy;
                o[6] = its.sh.s.x; o[7] = its.sh.s.y; o[8] = its.sh.s.z; o[9] = its.sh.t.x; o[10] = its.sh.t.y; o[11] = its.sh.t.z;
                o[12] = its.sh.n.x; o[13] = its.sh.n.y; o[14] = its.sh.n.z; o[15] = (float) its.mesh;
            }
        }
    }
    if (counters) {
        atomicAdd(&counters[1], (unsigned long long) n_rays);
        if (COUNT) { atomicAdd(&counters[2], (unsigned long long) n_nodes); atomicAdd(&counters[3], (unsigned long long) n_tris); }
    }
}

// ImageBlock::toBitmap (ref: src/block.cpp:45-51, include/nori/color.h:100-105)
__global__ void film_to_rgb_kernel(const float4 *film, int W, int H, int border, float *rgb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const int y = i / W, x = i % W;
    const float4 p = film[(size_t) (y + border) * (W + 2 * border) + (x + border)];
    float r = 0.f, g = 0.f, b = 0.f;
    if (p.w != 0.f) { r = p.x / p.w; g = p.y / p.w; b = p.z / p.w; }
    rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
}

// ImageBlock::toBitmap + Color3f::toSRGB + the 8-bit quantisation of Bitmap::savePNG in one pass over the film
// (ref: src/block.cpp:45-51, src/common.cpp:166-180, src/bitmap.cpp:100-110).  x^(1/2.4) = exp(log(x) / 2.4) through the
// deterministic polynomials shared with the host (csrc/host/common.cpp), so the bytes equal the host loop's.
__global__ void film_to_srgb8_kernel(const float4 *film, int W, int H, int border, unsigned char *rgb8) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const int y = i / W, x = i % W;
    const float4 p = film[(size_t) (y + border) * (W + 2 * border) + (x + border)];
    float c[3] = { 0.f, 0.f, 0.f };
    if (p.w != 0.f) { c[0] = p.x / p.w; c[1] = p.y / p.w; c[2] = p.z / p.w; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float v = c[k];
        const float t = v <= 0.0031308f ? 12.92f * v : (1.0f + 0.055f) * det_expf(det_logf(v) * (1.0f / 2.4f)) - 0.055f;
        rgb8[3 * (size_t) i + k] = (unsigned char) fminf(255.f, fmaxf(0.f, 255.f * t));
    }
}

}  // namespace nb
