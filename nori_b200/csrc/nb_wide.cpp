// nb_wide.cpp -- collapse of the binary hierarchy (nb_bvh.h layout, from either builder) into the 8-wide compressed layout
// of nb_wide.h, plus a host reference walk over it for the CPU tests.  Host code, no CUDA.
#include "nb_wide.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <limits>

namespace nb {
namespace {

struct Box3 { float lo[3], hi[3]; };

inline float area(const Box3 &b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return (dx < 0 || dy < 0 || dz < 0) ? 0.f : 2.f * (dx * dy + dy * dz + dz * dx);
}

inline Box3 child_box(const float *nd, int c) {        // nb_bvh.h: n0 = c0 x,y ; n1 = c1 x,y ; n2 = c0 z, c1 z
    Box3 b;
    if (c == 0) { b.lo[0] = nd[0]; b.hi[0] = nd[1]; b.lo[1] = nd[2]; b.hi[1] = nd[3]; b.lo[2] = nd[8]; b.hi[2] = nd[9]; }
    else { b.lo[0] = nd[4]; b.hi[0] = nd[5]; b.lo[1] = nd[6]; b.hi[1] = nd[7]; b.lo[2] = nd[10]; b.hi[2] = nd[11]; }
    return b;
}

inline int32_t child_ref(const float *nd, int c) { int32_t r; std::memcpy(&r, &nd[12 + c], 4); return r; }

struct Child { Box3 box; int32_t ref; };               // ref >= 0: binary inner node, < 0: leaf (~ref = first << 3 | count - 1)

}  // namespace

bool build_wide(const float *bnodes, uint32_t n_bnodes, const float *btris, uint32_t n_btris, WideOutput &out, const char **err) {
    auto t0 = std::chrono::steady_clock::now();
    static const char *e_leaf = "a binary leaf holds more than 3 triangles (wide nodes encode 1..3 per leaf child: build with max_leaf <= 3)";
    static const char *e_deep = "wide hierarchy too deep for the traversal stack";
    static const char *e_bad = "malformed binary hierarchy";
    out.nodes.clear(); out.tris.clear(); out.nnodes = 0; out.depth = 0;
    if (n_bnodes == 0) { if (err) *err = e_bad; return false; }
    out.nodes.reserve((size_t) n_bnodes * 8);
    out.tris.reserve((size_t) n_btris * 12);

    // work list: (binary node whose two children seed the wide node, wide node index, depth)
    struct Item { uint32_t bnode; uint32_t wnode; int depth; };
    std::vector<Item> todo;
    out.nodes.resize(20, 0u);
    out.nnodes = 1;
    todo.push_back({ 0u, 0u, 1 });
    while (!todo.empty()) {
        const Item it = todo.back(); todo.pop_back();
        out.depth = std::max(out.depth, it.depth);
        if (it.bnode >= n_bnodes) { if (err) *err = e_bad; return false; }
        // ---- greedy collapse: open the inner child with the largest surface area until there are 8 children
        Child ch[8]; int nch = 0;
        const float *root = bnodes + (size_t) it.bnode * 16;
        for (int c = 0; c < 2; ++c) { ch[nch].box = child_box(root, c); ch[nch].ref = child_ref(root, c); ++nch; }
        while (nch < 8) {
            int best = -1; float best_a = -1.f;
            for (int i = 0; i < nch; ++i) if (ch[i].ref >= 0) { const float a = area(ch[i].box); if (a > best_a) { best_a = a; best = i; } }
            if (best < 0) break;
            if ((uint32_t) ch[best].ref >= n_bnodes) { if (err) *err = e_bad; return false; }
            const float *nd = bnodes + (size_t) ch[best].ref * 16;
            const Child a = { child_box(nd, 0), child_ref(nd, 0) }, b = { child_box(nd, 1), child_ref(nd, 1) };
            ch[best] = a; ch[nch++] = b;
        }
        // ---- node box and quantisation grid
        Box3 nb;
        for (int a = 0; a < 3; ++a) { nb.lo[a] = std::numeric_limits<float>::infinity(); nb.hi[a] = -std::numeric_limits<float>::infinity(); }
        for (int i = 0; i < nch; ++i) for (int a = 0; a < 3; ++a) { nb.lo[a] = std::min(nb.lo[a], ch[i].box.lo[a]); nb.hi[a] = std::max(nb.hi[a], ch[i].box.hi[a]); }
        uint32_t ebyte[3]; double step[3];
        for (int a = 0; a < 3; ++a) {
            const double ext = (double) nb.hi[a] - (double) nb.lo[a];
            int ex = 1;                                                     // biased exponent; 2^(ex - 127) * 255 >= ext
            if (ext > 0) { int fe; std::frexp(ext / 255.0, &fe); ex = fe + 127; }     // ext / 255 = m * 2^fe, m in [0.5, 1)  =>  2^fe >= ext / 255
            ex = std::min(254, std::max(1, ex));
            while (ex < 254 && std::ldexp(1.0, ex - 127) * 255.0 < ext) ++ex;
            ebyte[a] = (uint32_t) ex; step[a] = std::ldexp(1.0, ex - 127);
        }
        // ---- slots: child i goes where its centroid lies relative to the node's (bit a of the slot = high side along axis a)
        int slot_of[8]; bool used[8] = { false, false, false, false, false, false, false, false };
        {
            double ctr[3]; for (int a = 0; a < 3; ++a) ctr[a] = 0.5 * ((double) nb.lo[a] + (double) nb.hi[a]);
            double score[8][8];
            for (int i = 0; i < nch; ++i) for (int s = 0; s < 8; ++s) {
                double v = 0;
                for (int a = 0; a < 3; ++a) v += (((s >> a) & 1) ? 1.0 : -1.0) * (0.5 * ((double) ch[i].box.lo[a] + (double) ch[i].box.hi[a]) - ctr[a]);
                score[i][s] = v;
            }
            bool placed[8] = { false, false, false, false, false, false, false, false };
            for (int k = 0; k < nch; ++k) {
                int bi = -1, bs = -1; double bv = -std::numeric_limits<double>::infinity();
                for (int i = 0; i < nch; ++i) if (!placed[i]) for (int s = 0; s < 8; ++s) if (!used[s] && score[i][s] > bv) { bv = score[i][s]; bi = i; bs = s; }
                placed[bi] = true; used[bs] = true; slot_of[bi] = bs;
            }
        }
        // ---- emit: inner children consecutive in slot order, triangles of leaf children consecutive in slot order
        int by_slot[8]; for (int s = 0; s < 8; ++s) by_slot[s] = -1;
        for (int i = 0; i < nch; ++i) by_slot[slot_of[i]] = i;
        uint32_t imask = 0, n_inner = 0;
        for (int s = 0; s < 8; ++s) if (by_slot[s] >= 0 && ch[by_slot[s]].ref >= 0) { imask |= 1u << s; ++n_inner; }
        const uint32_t child_base = out.nnodes;
        out.nnodes += n_inner;
        out.nodes.resize((size_t) out.nnodes * 20, 0u);
        const uint32_t tri_base = (uint32_t) (out.tris.size() / 12);
        uint32_t w[20]; for (int k = 0; k < 20; ++k) w[k] = 0u;
        std::memcpy(&w[0], &nb.lo[0], 4); std::memcpy(&w[1], &nb.lo[1], 4); std::memcpy(&w[2], &nb.lo[2], 4);
        w[3] = ebyte[0] | (ebyte[1] << 8) | (ebyte[2] << 16) | (imask << 24);
        w[4] = child_base; w[5] = tri_base;
        uint32_t inner_seen = 0, tri_off = 0;
        for (int s = 0; s < 8; ++s) {
            const int i = by_slot[s];
            if (i < 0) continue;
            uint32_t meta;
            if (ch[i].ref >= 0) {
                meta = 0x20u | (24u + (uint32_t) s);
                todo.push_back({ (uint32_t) ch[i].ref, child_base + inner_seen, it.depth + 1 });
                ++inner_seen;
            } else {
                const uint32_t payload = ~(uint32_t) ch[i].ref;
                const uint32_t first = payload >> 3, count = (payload & 7u) + 1u;
                if (count > 3) { if (err) *err = e_leaf; return false; }
                if (first + count > n_btris || tri_off + count > 24) { if (err) *err = e_bad; return false; }
                meta = (((1u << count) - 1u) << 5) | tri_off;
                out.tris.insert(out.tris.end(), btris + (size_t) first * 12, btris + (size_t) (first + count) * 12);
                tri_off += count;
            }
            w[6 + (s >> 2)] |= meta << (8 * (s & 3));
            // quantise OUTWARDS: floor for the low planes, ceil for the high planes, then make sure in double arithmetic
            uint32_t q[6];
            for (int a = 0; a < 3; ++a) {
                const double lo = ((double) ch[i].box.lo[a] - (double) nb.lo[a]) / step[a], hi = ((double) ch[i].box.hi[a] - (double) nb.lo[a]) / step[a];
                long ql = (long) std::floor(lo), qh = (long) std::ceil(hi);
                while (ql > 0 && (double) nb.lo[a] + (double) ql * step[a] > (double) ch[i].box.lo[a]) --ql;
                while (qh < 255 && (double) nb.lo[a] + (double) qh * step[a] < (double) ch[i].box.hi[a]) ++qh;
                q[a] = (uint32_t) std::min(255l, std::max(0l, ql)); q[3 + a] = (uint32_t) std::min(255l, std::max(0l, qh));
            }
            const int sh = 8 * (s & 3), hf = s >> 2;
            w[8 + hf] |= q[0] << sh; w[10 + hf] |= q[1] << sh; w[12 + hf] |= q[2] << sh;
            w[14 + hf] |= q[3] << sh; w[16 + hf] |= q[4] << sh; w[18 + hf] |= q[5] << sh;
        }
        std::memcpy(out.nodes.data() + (size_t) it.wnode * 20, w, sizeof w);
    }
    if (2 * out.depth + 2 > kWideStack) { if (err) *err = e_deep; return false; }
    out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

// ------------------------------------------------------------------ host reference walk (tests only: the device walk is the product)
namespace {
struct HostHit { float t, u, v; uint32_t prim; };

// Moeller-Trumbore exactly as the kernels' leaf_test (ref: src/mesh.cpp:39-76); the library is compiled -ffp-contract=off
inline bool tri_test(const float *tri, const float o[3], const float d[3], float mint, float maxt, float &t, float &u, float &v) {
    const float p0[3] = { tri[0], tri[1], tri[2] };
    const float e1[3] = { tri[4] - p0[0], tri[5] - p0[1], tri[6] - p0[2] }, e2[3] = { tri[8] - p0[0], tri[9] - p0[1], tri[10] - p0[2] };
    auto dot = [](const float *a, const float *b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); };
    auto cross = [](const float *a, const float *b, float *c) { c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; };
    float pvec[3]; cross(d, e2, pvec);
    const float det = dot(e1, pvec);
    if (det > -1e-8f && det < 1e-8f) return false;
    const float inv = 1.0f / det;
    const float tvec[3] = { o[0] - p0[0], o[1] - p0[1], o[2] - p0[2] };
    u = dot(tvec, pvec) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    float qvec[3]; cross(tvec, e1, qvec);
    v = dot(d, qvec) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot(e2, qvec) * inv;
    return t >= mint && t <= maxt;
}
}  // namespace
}  // namespace nb

// Host-only diagnostic (declared in include/nori_b200.h): binary build (as nb_debug_build_bvh, max_leaf 3) + wide collapse;
// returns the wide layout.  info = { wide nodes, triangles, wide depth, binary nodes }.
extern "C" int nb_debug_build_bvh(const float *verts4, const uint32_t *faces4, uint32_t nprims, int max_leaf, int64_t bfs_nodes,
                                  float *nodes_out, uint64_t nodes_cap, float *tris_out, uint64_t tris_cap, uint32_t info[4]);

extern "C" int nb_debug_build_wide(const float *verts4, const uint32_t *faces4, uint32_t nprims, uint32_t *nodes_out, uint64_t nodes_cap,
                                   float *tris_out, uint64_t tris_cap, uint32_t info[4]) {
    if (!info) return 1;
    uint32_t bi[4];
    if (nb_debug_build_bvh(verts4, faces4, nprims, 3, 2048, nullptr, 0, nullptr, 0, bi)) return 1;
    std::vector<float> bn((size_t) bi[0] * 16), bt((size_t) bi[1] * 12);
    if (nb_debug_build_bvh(verts4, faces4, nprims, 3, 2048, bn.data(), bn.size(), bt.data(), bt.size(), bi)) return 1;
    nb::WideOutput w; const char *err = nullptr;
    if (!nb::build_wide(bn.data(), bi[0], bt.data(), bi[1], w, &err)) { fprintf(stderr, "build_wide: %s\n", err ? err : "?"); return 3; }
    info[0] = w.nnodes; info[1] = (uint32_t) (w.tris.size() / 12); info[2] = (uint32_t) w.depth; info[3] = bi[0];
    if (nodes_out) { if (nodes_cap < w.nodes.size()) return 2; std::memcpy(nodes_out, w.nodes.data(), w.nodes.size() * 4); }
    if (tris_out) { if (tris_cap < w.tris.size()) return 2; std::memcpy(tris_out, w.tris.data(), w.tris.size() * 4); }
    return 0;
}

// Host reference walk over a wide hierarchy: rays = 8 floats each (o, mint, d, maxt), hits = (t, u, v, prim bits) per ray
// (prim = 0xffffffff: miss).  Uses the SAME node step as the kernels (nb_wide.h: wide_node_test).  counts (nullable) =
// { node visits, triangle tests }.
extern "C" int nb_debug_wide_intersect(const uint32_t *nodes, uint32_t nnodes, const float *tris, const float *rays, uint64_t nrays,
                                       int any_hit, float *hits4, uint64_t counts[2]) {
    using namespace nb;
    uint64_t n_nodes = 0, n_tris = 0;
    for (uint64_t r = 0; r < nrays; ++r) {
        const float *ry = rays + 8 * r;
        const float o[3] = { ry[0], ry[1], ry[2] }, d[3] = { ry[4], ry[5], ry[6] };
        const float mint = ry[3]; float maxt = ry[7];
        WideRay R;
        const float ooeps = 1e-24f;
        const float dx = std::fabs(d[0]) > ooeps ? d[0] : std::copysign(ooeps, d[0]), dy = std::fabs(d[1]) > ooeps ? d[1] : std::copysign(ooeps, d[1]),
                    dz = std::fabs(d[2]) > ooeps ? d[2] : std::copysign(ooeps, d[2]);
        R.idx = 1.0f / dx; R.idy = 1.0f / dy; R.idz = 1.0f / dz; R.ox = o[0]; R.oy = o[1]; R.oz = o[2];
        R.negx = R.idx < 0.f; R.negy = R.idy < 0.f; R.negz = R.idz < 0.f;
        const uint32_t oct = (R.negx ? 1u : 0u) | (R.negy ? 2u : 0u) | (R.negz ? 4u : 0u);
        R.octinv4 = (7u - oct) * 0x01010101u;
        float hu = 0, hv = 0; uint32_t hprim = 0xffffffffu;
        uint32_t stack[2 * kWideStack]; int sp = 0;
        uint32_t ngx = 0, ngy = 0x80000000u, tgx = 0, tgy = 0;
        bool done = false;
        while (!done) {
            if (ngy > 0x00ffffffu) {
                const uint32_t hits = ngy, imask = ngy & 0xffu;
                const int bit = wide_bfind(hits);
                const uint32_t base = ngx;
                ngy &= ~(1u << bit);
                if (ngy > 0x00ffffffu) { if (sp >= 2 * kWideStack) return 4; stack[sp++] = ngx; stack[sp++] = ngy; }
                const uint32_t slot = ((uint32_t) (bit - 24)) ^ (R.octinv4 & 7u);
                const uint32_t rel = (uint32_t) wide_popc(imask & ~(0xffffffffu << slot));
                const uint32_t ni = base + rel;
                if (ni >= nnodes) return 5;
                uint32_t cb, tb, im;
                const uint32_t hm = wide_node_test([&](uint32_t *w) { std::memcpy(w, nodes + (size_t) ni * 20, 80); }, R, mint, maxt, cb, tb, im);
                ++n_nodes;
                ngx = cb; ngy = (hm & 0xff000000u) | im; tgx = tb; tgy = hm & 0x00ffffffu;
            } else { tgx = ngx; tgy = ngy; ngx = 0; ngy = 0; }
            while (tgy != 0) {
                const int ti = wide_bfind(tgy);
                tgy &= ~(1u << ti);
                const float *tri = tris + (size_t) (tgx + (uint32_t) ti) * 12;
                float t, u, v; ++n_tris;
                if (!tri_test(tri, o, d, mint, maxt, t, u, v)) continue;
                uint32_t prim; std::memcpy(&prim, &tri[3], 4);
                if (any_hit) { hprim = prim; done = true; break; }
                if (hprim == 0xffffffffu || t < maxt || prim > hprim) { maxt = t; hu = u; hv = v; hprim = prim; }
            }
            if (done) break;
            if (ngy <= 0x00ffffffu) {
                if (sp > 0) { ngy = stack[--sp]; ngx = stack[--sp]; }
                else break;
            }
        }
        float *h = hits4 + 4 * r;
        h[0] = hprim != 0xffffffffu && !any_hit ? maxt : 0.f; h[1] = any_hit ? 0.f : hu; h[2] = any_hit ? 0.f : hv;
        std::memcpy(&h[3], &hprim, 4);
    }
    if (counts) { counts[0] = n_nodes; counts[1] = n_tris; }
    return 0;
}
