// nb_bvh.cpp -- binned-SAH BVH2 builder (host, multi-threaded).  See nb_bvh.h for the layout contract.
//
// The reference ships no hierarchy (brute force, ref: src/accel.cpp:30-43); results must not depend
// on the tree, so every box is padded by a small multiple of the coordinate magnitude: the slab test may
// then only cull candidates that Moeller-Trumbore (ref: src/mesh.cpp:39-76) would reject anyway.
#include "nb_bvh.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>

namespace nb {
namespace {

constexpr int kBins = 32;           // array capacity; the number of bins in use is Builder::nbins (32 by default: -1 % render time, -9 % node visits on the Cornell box against 16, profiles/r2_call5_walk_variants_ab.txt)
constexpr float kInf = std::numeric_limits<float>::infinity();

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = kInf; hi[a] = -kInf; } }
    void grow(const Box &b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const float *p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx < 0 ? 0.f : 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

struct PrimRef { Box b; float c[3]; };

struct BNode {
    Box box;
    uint32_t left = 0, right = 0;   // valid if count == 0
    uint32_t start = 0, count = 0;  // leaf if count > 0
    int depth = 0;
};

// Runs fn(begin, end, chunk) over [b, e) split into `chunks` contiguous pieces on separate threads.
template <typename F> void parallel_chunks(uint32_t b, uint32_t e, int chunks, F fn) {
    std::vector<std::thread> th;
    const uint32_t n = e - b;
    for (int c = 0; c < chunks; ++c) {
        const uint32_t cb = b + (uint32_t) ((uint64_t) n * c / chunks), ce = b + (uint32_t) ((uint64_t) n * (c + 1) / chunks);
        th.emplace_back([=] { fn(cb, ce, c); });
    }
    for (auto &t : th) t.join();
}

struct Builder {
    const PrimRef *prims;
    uint32_t *order;
    uint32_t *scratch = nullptr;     // same size as order: parallel partition target
    int nthreads = 1;
    static constexpr uint32_t kParallelNode = 1u << 19;   // nodes with at least this many triangles are processed by all threads
    std::vector<BNode> nodes;
    std::atomic<uint32_t> nnodes{0};
    std::atomic<int> threads_free{0};
    int max_leaf;
    int nbins = 32;                  // SAH bins per axis (<= kBins)
    int max_depth = 64;              // the walk's per-lane stack: every leaf must end up shallower than this

    uint32_t alloc() { return nnodes.fetch_add(1); }

    void build(uint32_t ni, uint32_t start, uint32_t end, int depth) {
        Box box, cbox;
        box.reset(); cbox.reset();
        const bool wide = (end - start) >= kParallelNode && nthreads > 1 && threads_free.load(std::memory_order_relaxed) >= nthreads - 1;
        if (wide) {
            std::vector<Box> pb((size_t) nthreads), pc((size_t) nthreads);
            parallel_chunks(start, end, nthreads, [&](uint32_t b, uint32_t e, int c) {
                Box bb, cc; bb.reset(); cc.reset();
                for (uint32_t i = b; i < e; ++i) { const PrimRef &p = prims[order[i]]; bb.grow(p.b); cc.grow(p.c); }
                pb[(size_t) c] = bb; pc[(size_t) c] = cc;
            });
            for (int c = 0; c < nthreads; ++c) { box.grow(pb[(size_t) c]); cbox.grow(pc[(size_t) c].lo); cbox.grow(pc[(size_t) c].hi); }
        } else {
            for (uint32_t i = start; i < end; ++i) { const PrimRef &p = prims[order[i]]; box.grow(p.b); cbox.grow(p.c); }
        }
        BNode &nd = nodes[ni];
        nd.box = box; nd.depth = depth;
        const uint32_t n = end - start;
        if (n <= (uint32_t) max_leaf) { nd.start = start; nd.count = n; return; }

        int best_axis = -1, best_bin = -1;
        float best_cost = kInf;
        // A subtree of n triangles finished by balanced index splits is at most ceil(log2 n) + 1 levels deep, so a SAH split
        // is taken only while its children could still be finished that way below max_depth: the traversal stack can never
        // overflow, whatever the geometry (depth + 2 + ceil(log2 n) < max_depth).
        int lg = 0; while ((1u << lg) < n) ++lg;
        if (depth + 2 + lg < max_depth) {
            Box bins[3][kBins]; uint32_t cnt[3][kBins];
            const int NB = nbins;
            float scale[3];
            for (int a = 0; a < 3; ++a) {
                float ext = cbox.hi[a] - cbox.lo[a];
                scale[a] = ext > 0 ? NB / ext : 0.f;
                for (int k = 0; k < NB; ++k) { bins[a][k].reset(); cnt[a][k] = 0; }
            }
            auto bin_range = [&](uint32_t b, uint32_t e, Box (*bn)[kBins], uint32_t (*cn)[kBins]) {
                for (uint32_t i = b; i < e; ++i) {
                    const PrimRef &p = prims[order[i]];
                    for (int a = 0; a < 3; ++a) {
                        if (scale[a] == 0.f) continue;
                        int k = std::min(NB - 1, std::max(0, (int) ((p.c[a] - cbox.lo[a]) * scale[a])));
                        bn[a][k].grow(p.b); cn[a][k]++;
                    }
                }
            };
            if (wide) {
                struct Part { Box b[3][kBins]; uint32_t c[3][kBins]; };
                std::vector<Part> parts((size_t) nthreads);
                parallel_chunks(start, end, nthreads, [&](uint32_t b, uint32_t e, int c) {
                    Part &pt = parts[(size_t) c];
                    for (int a = 0; a < 3; ++a) for (int k = 0; k < NB; ++k) { pt.b[a][k].reset(); pt.c[a][k] = 0; }
                    bin_range(b, e, pt.b, pt.c);
                });
                for (auto &pt : parts) for (int a = 0; a < 3; ++a) for (int k = 0; k < NB; ++k)
                    if (pt.c[a][k]) { bins[a][k].grow(pt.b[a][k]); cnt[a][k] += pt.c[a][k]; }
            } else {
                bin_range(start, end, bins, cnt);
            }
            for (int a = 0; a < 3; ++a) {
                if (scale[a] == 0.f) continue;
                float ra[kBins]; uint32_t rc[kBins];
                Box acc; acc.reset(); uint32_t c = 0;
                for (int k = NB - 1; k >= 0; --k) { if (cnt[a][k]) acc.grow(bins[a][k]); c += cnt[a][k]; ra[k] = acc.area(); rc[k] = c; }
                acc.reset(); c = 0;
                for (int k = 0; k < NB - 1; ++k) {
                    if (cnt[a][k]) acc.grow(bins[a][k]);
                    c += cnt[a][k];
                    if (c == 0 || rc[k + 1] == 0) continue;
                    float cost = acc.area() * (float) c + ra[k + 1] * (float) rc[k + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k; }
                }
            }
        }
        uint32_t mid;
        if (best_axis < 0) {
            mid = start + n / 2;
        } else {
            const int a = best_axis;
            const float sc = nbins / (cbox.hi[a] - cbox.lo[a]), lo = cbox.lo[a];
            const int NBp = nbins;
            auto goes_left = [&](uint32_t id) {
                int k = std::min(NBp - 1, std::max(0, (int) ((prims[id].c[a] - lo) * sc)));
                return k <= best_bin;
            };
            if (wide && scratch) {
                // two-pass parallel partition through the scratch array (stable within chunks)
                std::vector<uint32_t> nl((size_t) nthreads, 0), nr((size_t) nthreads, 0);
                parallel_chunks(start, end, nthreads, [&](uint32_t b, uint32_t e, int c) {
                    uint32_t l = 0; for (uint32_t i = b; i < e; ++i) l += goes_left(order[i]) ? 1u : 0u;
                    nl[(size_t) c] = l; nr[(size_t) c] = (e - b) - l;
                });
                uint32_t total_l = 0; for (uint32_t v : nl) total_l += v;
                std::vector<uint32_t> ol((size_t) nthreads), orr((size_t) nthreads);
                uint32_t accl = start, accr = start + total_l;
                for (int c = 0; c < nthreads; ++c) { ol[(size_t) c] = accl; orr[(size_t) c] = accr; accl += nl[(size_t) c]; accr += nr[(size_t) c]; }
                parallel_chunks(start, end, nthreads, [&](uint32_t b, uint32_t e, int c) {
                    uint32_t l = ol[(size_t) c], r = orr[(size_t) c];
                    for (uint32_t i = b; i < e; ++i) { const uint32_t id = order[i]; if (goes_left(id)) scratch[l++] = id; else scratch[r++] = id; }
                });
                parallel_chunks(start, end, nthreads, [&](uint32_t b, uint32_t e, int) { std::memcpy(order + b, scratch + b, sizeof(uint32_t) * (e - b)); });
                mid = start + total_l;
            } else {
                uint32_t *first = order + start, *last = order + end;
                uint32_t *m = std::partition(first, last, goes_left);
                mid = (uint32_t) (m - order);
            }
            if (mid == start || mid == end) mid = start + n / 2;
        }
        const uint32_t l = alloc(), r = alloc();
        nodes[ni].left = l; nodes[ni].right = r;
        bool spawned = false;
        std::thread th;
        if (n > 65536 && threads_free.load(std::memory_order_relaxed) > 0) {
            if (threads_free.fetch_sub(1) > 0) {
                spawned = true;
                th = std::thread([this, l, start, mid, depth] { build(l, start, mid, depth + 1); });
            } else {
                threads_free.fetch_add(1);
            }
        }
        if (!spawned) build(l, start, mid, depth + 1);
        build(r, mid, end, depth + 1);
        if (spawned) { th.join(); threads_free.fetch_add(1); }
    }
};

inline float as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

}  // namespace

void build_bvh(const BvhInput &in, BvhOutput &out, int max_leaf, uint32_t bfs_nodes, int nthreads, int max_depth, int sah_bins) {
    auto t0 = std::chrono::steady_clock::now();
    max_leaf = std::min(8, std::max(1, max_leaf));
    const uint32_t n = in.nprims;
    std::vector<PrimRef> prims(n);
    std::vector<uint32_t> order(n);
    float maxabs = 0.f;
    Box scene; scene.reset();
    for (uint32_t i = 0; i < n; ++i) {
        PrimRef &p = prims[i];
        p.b.reset();
        for (int k = 0; k < 3; ++k) {
            const float *v = in.verts + 4 * (size_t) in.faces[4 * (size_t) i + k];
            p.b.grow(v);
            for (int a = 0; a < 3; ++a) maxabs = std::max(maxabs, std::fabs(v[a]));
        }
        for (int a = 0; a < 3; ++a) p.c[a] = 0.5f * (p.b.lo[a] + p.b.hi[a]);
        scene.grow(p.b);
        order[i] = i;
    }
    const float pad = 4e-6f * maxabs;
    for (uint32_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) { prims[i].b.lo[a] -= pad; prims[i].b.hi[a] += pad; }
    for (int a = 0; a < 3; ++a) { out.scene_lo[a] = n ? scene.lo[a] : 0.f; out.scene_hi[a] = n ? scene.hi[a] : 0.f; }

    Builder b;
    std::vector<uint32_t> scratch(n >= Builder::kParallelNode ? n : 0);
    b.prims = prims.data(); b.order = order.data(); b.max_leaf = max_leaf; b.max_depth = std::max(8, max_depth); b.nbins = std::min(kBins, std::max(4, sah_bins));
    b.scratch = scratch.empty() ? nullptr : scratch.data();
    b.nodes.resize(n ? 2 * (size_t) n : 1);
    if (nthreads <= 0) nthreads = (int) std::thread::hardware_concurrency();
    nthreads = std::max(1, std::min(nthreads, 64));
    b.nthreads = nthreads;
    b.threads_free = std::max(0, nthreads - 1);
    if (n) { uint32_t root = b.alloc(); b.build(root, 0, n, 0); }
    if (getenv("NB_BVH_TIMING")) fprintf(stderr, "[bvh] refs+build %.2fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());

    // ---- relayout: inner nodes only; BFS for the first bfs_nodes, DFS below ----
    out.nodes.clear(); out.tris.clear(); out.depth = 0;
    auto absent = [&](float *nd, int c) {   // degenerate point box at the origin (the leaf behind it rejects everything)
        if (c == 0) { nd[0] = nd[1] = nd[2] = nd[3] = nd[8] = nd[9] = 0.f; }
        else { nd[4] = nd[5] = nd[6] = nd[7] = nd[10] = nd[11] = 0.f; }
    };
    auto put_box = [&](float *nd, int c, const Box &bx) {
        if (c == 0) { nd[0] = bx.lo[0]; nd[1] = bx.hi[0]; nd[2] = bx.lo[1]; nd[3] = bx.hi[1]; nd[8] = bx.lo[2]; nd[9] = bx.hi[2]; }
        else { nd[4] = bx.lo[0]; nd[5] = bx.hi[0]; nd[6] = bx.lo[1]; nd[7] = bx.hi[1]; nd[10] = bx.lo[2]; nd[11] = bx.hi[2]; }
    };
    auto emit_leaf = [&](const BNode &lf) -> int32_t {
        const uint32_t first = (uint32_t) (out.tris.size() / 12);
        for (uint32_t i = 0; i < lf.count; ++i) {
            const uint32_t prim = order[lf.start + i];
            for (int k = 0; k < 3; ++k) {
                const float *v = in.verts + 4 * (size_t) in.faces[4 * (size_t) prim + k];
                out.tris.push_back(v[0]); out.tris.push_back(v[1]); out.tris.push_back(v[2]);
                out.tris.push_back(k == 0 ? as_float(prim) : 0.f);
            }
        }
        return (int32_t) ~((first << 3) | (lf.count - 1));
    };

    if (n == 0 || b.nodes[0].count > 0) {
        // degenerate trees: a root whose child 0 is the only leaf (or nothing at all)
        out.nodes.assign(16, 0.f);
        float *nd = out.nodes.data();
        absent(nd, 0); absent(nd, 1);
        // An absent child is a one-triangle leaf holding a degenerate (all-zero) triangle: det == 0, always rejected
        // (ref: src/mesh.cpp:52).  (An inverted box would NOT be culled by a min/max slab test.)
        int32_t r0 = 0, r1 = 0;
        if (n) { put_box(nd, 0, b.nodes[0].box); r0 = emit_leaf(b.nodes[0]); }
        const uint32_t dummy = (uint32_t) (out.tris.size() / 12);
        for (int k = 0; k < 12; ++k) out.tris.push_back(0.f);
        r1 = (int32_t) ~((dummy << 3) | 0u);
        if (!n) r0 = r1;
        std::memcpy(&nd[12], &r0, 4); std::memcpy(&nd[13], &r1, 4);
        out.nnodes = 1; out.top_nodes = 1; out.depth = 1;
    } else {
        // final index assignment
        std::vector<uint32_t> final_order;   // build-node ids of inner nodes in final order
        final_order.reserve(n);
        std::vector<int32_t> final_index(b.nnodes.load(), -1);
        constexpr uint32_t kHole = 0xffffffffu;      // padding slot (sibling-pair layout)
        if (bfs_nodes == kSiblingPairs) {
            // Sibling-pair layout: the two inner children of a node occupy one aligned 128-byte line (two 64 B nodes), so
            // the walk's later pop of the far child hits the line its near sibling already brought into L1.
            final_index[0] = 0; final_order.push_back(0);
            std::vector<uint32_t> stack; stack.push_back(0);
            while (!stack.empty()) {
                uint32_t id = stack.back(); stack.pop_back();
                const BNode &nd = b.nodes[id];
                const bool li = b.nodes[nd.left].count == 0, ri = b.nodes[nd.right].count == 0;
                if (li && ri && (final_order.size() & 1u)) final_order.push_back(kHole);
                if (li) { final_index[nd.left] = (int32_t) final_order.size(); final_order.push_back(nd.left); }
                if (ri) { final_index[nd.right] = (int32_t) final_order.size(); final_order.push_back(nd.right); }
                if (ri) stack.push_back(nd.right);
                if (li) stack.push_back(nd.left);
            }
            out.top_nodes = 0;
        } else {
        std::vector<uint32_t> queue; queue.push_back(0);
        size_t qh = 0;
        while (qh < queue.size() && final_order.size() < bfs_nodes) {
            uint32_t id = queue[qh++];
            final_index[id] = (int32_t) final_order.size();
            final_order.push_back(id);
            const BNode &nd = b.nodes[id];
            if (b.nodes[nd.left].count == 0) queue.push_back(nd.left);
            if (b.nodes[nd.right].count == 0) queue.push_back(nd.right);
        }
        out.top_nodes = (uint32_t) final_order.size();
        std::vector<uint32_t> stack;
        for (size_t q = queue.size(); q-- > qh;) stack.push_back(queue[q]);   // remaining subtrees, DFS each (in queue order)
        while (!stack.empty()) {
            uint32_t id = stack.back(); stack.pop_back();
            final_index[id] = (int32_t) final_order.size();
            final_order.push_back(id);
            const BNode &nd = b.nodes[id];
            if (b.nodes[nd.right].count == 0) stack.push_back(nd.right);
            if (b.nodes[nd.left].count == 0) stack.push_back(nd.left);
        }
        }
        out.nnodes = (uint32_t) final_order.size();
        out.nodes.assign((size_t) out.nnodes * 16, 0.f);
        // leaf triangle offsets in final node order (sequential prefix), then nodes + triangles are written in parallel
        std::vector<uint32_t> leaf_first((size_t) out.nnodes * 2, 0);
        uint32_t tri_cursor = 0;
        for (uint32_t fi = 0; fi < out.nnodes; ++fi) {
            if (final_order[fi] == kHole) continue;
            const BNode &nd = b.nodes[final_order[fi]];
            out.depth = std::max(out.depth, nd.depth + 2);
            const uint32_t ch[2] = { nd.left, nd.right };
            for (int c = 0; c < 2; ++c) {
                const BNode &cn = b.nodes[ch[c]];
                if (cn.count > 0) { leaf_first[(size_t) fi * 2 + c] = tri_cursor; tri_cursor += cn.count; }
            }
        }
        out.tris.assign((size_t) tri_cursor * 12, 0.f);
        auto fill = [&](uint32_t fb, uint32_t fe, int) {
            for (uint32_t fi = fb; fi < fe; ++fi) {
                if (final_order[fi] == kHole) continue;      // padding slot: never referenced, stays zero
                const BNode &nd = b.nodes[final_order[fi]];
                float *o = out.nodes.data() + (size_t) fi * 16;
                const uint32_t ch[2] = { nd.left, nd.right };
                for (int c = 0; c < 2; ++c) {
                    const BNode &cn = b.nodes[ch[c]];
                    put_box(o, c, cn.box);
                    int32_t ref;
                    if (cn.count > 0) {
                        const uint32_t first = leaf_first[(size_t) fi * 2 + c];
                        for (uint32_t i = 0; i < cn.count; ++i) {
                            const uint32_t prim = order[cn.start + i];
                            float *t = out.tris.data() + (size_t) (first + i) * 12;
                            for (int k = 0; k < 3; ++k) {
                                const float *v = in.verts + 4 * (size_t) in.faces[4 * (size_t) prim + k];
                                t[4 * k] = v[0]; t[4 * k + 1] = v[1]; t[4 * k + 2] = v[2]; t[4 * k + 3] = k == 0 ? as_float(prim) : 0.f;
                            }
                        }
                        ref = (int32_t) ~((first << 3) | (cn.count - 1));
                    } else {
                        ref = final_index[ch[c]];
                    }
                    std::memcpy(&o[12 + c], &ref, 4);
                }
            }
        };
        if (out.nnodes >= 65536 && nthreads > 1) parallel_chunks(0, out.nnodes, nthreads, fill);
        else fill(0, out.nnodes, 0);
    }
    out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ------------------------------------------------------------------ on-disk cache (see nb_bvh.h)
namespace {
struct CacheHeader {
    char magic[8];              // "NBBVH002": layout version of nodes / tris
    uint64_t key;
    uint32_t nnodes, ntris, top_nodes; int32_t depth;
    float scene_lo[3], scene_hi[3];
    uint64_t payload_hash;      // FNV-1a of the node and triangle bytes: a truncated or damaged file misses
};
inline uint64_t fnv1a(const void *data, size_t n, uint64_t h) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    // 8 bytes per step (the arrays are 4-byte words; a tail of < 8 bytes is folded bytewise)
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; std::memcpy(&w, p + i, 8); h = (h ^ w) * 0x100000001b3ull; }
    for (; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}
}  // namespace

uint64_t bvh_cache_key(const BvhInput &in, int max_leaf, uint32_t bfs_nodes, int max_depth, int sah_bins) {
    uint64_t h = 0xcbf29ce484222325ull;
    const uint32_t params[6] = { in.nprims, (uint32_t) max_leaf, bfs_nodes, 2u /* layout version */, (uint32_t) max_depth, (uint32_t) sah_bins };
    h = fnv1a(params, sizeof params, h);
    uint32_t max_v = 0;
    for (uint32_t i = 0; i < in.nprims; ++i) for (int k = 0; k < 3; ++k) max_v = std::max(max_v, in.faces[4 * (size_t) i + k]);
    h = fnv1a(in.faces, sizeof(uint32_t) * 4 * (size_t) in.nprims, h);
    if (in.nprims) h = fnv1a(in.verts, sizeof(float) * 4 * ((size_t) max_v + 1), h);
    return h;
}

bool bvh_cache_load(const char *path, uint64_t key, BvhOutput &out) {
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) return false;
    CacheHeader hd;
    bool ok = fread(&hd, sizeof hd, 1, f) == 1 && std::memcmp(hd.magic, "NBBVH002", 8) == 0 && hd.key == key;
    if (ok) {
        out.nodes.resize((size_t) hd.nnodes * 16); out.tris.resize((size_t) hd.ntris * 12);
        ok = fread(out.nodes.data(), sizeof(float), out.nodes.size(), f) == out.nodes.size() &&
             fread(out.tris.data(), sizeof(float), out.tris.size(), f) == out.tris.size();
        if (ok) {
            uint64_t h = fnv1a(out.nodes.data(), out.nodes.size() * sizeof(float), 0xcbf29ce484222325ull);
            h = fnv1a(out.tris.data(), out.tris.size() * sizeof(float), h);
            ok = h == hd.payload_hash;
        }
    }
    fclose(f);
    if (!ok) { out.nodes.clear(); out.tris.clear(); return false; }
    out.nnodes = hd.nnodes; out.top_nodes = hd.top_nodes; out.depth = hd.depth; out.build_seconds = 0;
    for (int a = 0; a < 3; ++a) { out.scene_lo[a] = hd.scene_lo[a]; out.scene_hi[a] = hd.scene_hi[a]; }
    return true;
}

bool bvh_cache_save(const char *path, uint64_t key, const BvhOutput &out) {
    if (!path || !*path) return false;
    const std::string tmp = std::string(path) + ".tmp";     // written aside and renamed: readers never see half a file
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    CacheHeader hd; std::memset(&hd, 0, sizeof hd);
    std::memcpy(hd.magic, "NBBVH002", 8);
    hd.key = key; hd.nnodes = out.nnodes; hd.ntris = (uint32_t) (out.tris.size() / 12); hd.top_nodes = out.top_nodes; hd.depth = out.depth;
    for (int a = 0; a < 3; ++a) { hd.scene_lo[a] = out.scene_lo[a]; hd.scene_hi[a] = out.scene_hi[a]; }
    hd.payload_hash = fnv1a(out.nodes.data(), out.nodes.size() * sizeof(float), 0xcbf29ce484222325ull);
    hd.payload_hash = fnv1a(out.tris.data(), out.tris.size() * sizeof(float), hd.payload_hash);
    bool ok = fwrite(&hd, sizeof hd, 1, f) == 1 && fwrite(out.nodes.data(), sizeof(float), out.nodes.size(), f) == out.nodes.size() &&
              fwrite(out.tris.data(), sizeof(float), out.tris.size(), f) == out.tris.size();
    ok = fclose(f) == 0 && ok;
    if (ok) ok = std::rename(tmp.c_str(), path) == 0;
    if (!ok) std::remove(tmp.c_str());
    return ok;
}

}  // namespace nb

// Host-only diagnostic of the hierarchy cache (no context, no GPU; declared in include/nori_b200.h): what nb_build_accel does
// with nb_set_accel_cache -- key, load or build + save.  info = { nodes, leaf triangles, top nodes, depth, cache hit (0/1) }.
extern "C" int nb_debug_bvh_cache(const float *verts4, const uint32_t *faces4, uint32_t nprims, int max_leaf, int64_t bfs_nodes,
                                  const char *path, float *nodes_out, uint64_t nodes_cap, float *tris_out, uint64_t tris_cap, uint32_t info[5]) {
    if ((nprims && (!verts4 || !faces4)) || !info || !path || max_leaf < 1 || max_leaf > 8) return 1;
    nb::BvhInput in; in.verts = verts4; in.faces = faces4; in.nprims = nprims;
    const uint32_t bfs = bfs_nodes < 0 ? nb::kSiblingPairs : (uint32_t) bfs_nodes;
    const uint64_t key = nb::bvh_cache_key(in, max_leaf, bfs);
    nb::BvhOutput out;
    const bool hit = nb::bvh_cache_load(path, key, out);
    if (!hit) { nb::build_bvh(in, out, max_leaf, bfs, 0); nb::bvh_cache_save(path, key, out); }
    info[0] = out.nnodes; info[1] = (uint32_t) (out.tris.size() / 12); info[2] = out.top_nodes; info[3] = (uint32_t) out.depth; info[4] = hit ? 1u : 0u;
    if (nodes_out) { if (nodes_cap < out.nodes.size()) return 2; std::memcpy(nodes_out, out.nodes.data(), out.nodes.size() * sizeof(float)); }
    if (tris_out) { if (tris_cap < out.tris.size()) return 2; if (!out.tris.empty()) std::memcpy(tris_out, out.tris.data(), out.tris.size() * sizeof(float)); }
    return 0;
}

// Host-only diagnostic entry (declared in include/nori_b200.h): runs the SAH builder on caller-provided arrays and
// returns the device layout, so that the hierarchy can be checked without a GPU (tests/test_bvh_cpu.py).
extern "C" int nb_debug_build_bvh(const float *verts4, const uint32_t *faces4, uint32_t nprims, int max_leaf, int64_t bfs_nodes,
                                  float *nodes_out, uint64_t nodes_cap, float *tris_out, uint64_t tris_cap, uint32_t info[4]) {
    if ((nprims && (!verts4 || !faces4)) || !info || max_leaf < 1 || max_leaf > 8) return 1;
    nb::BvhInput in; in.verts = verts4; in.faces = faces4; in.nprims = nprims;
    nb::BvhOutput out;
    nb::build_bvh(in, out, max_leaf, bfs_nodes < 0 ? nb::kSiblingPairs : (uint32_t) bfs_nodes, 0);
    info[0] = out.nnodes; info[1] = (uint32_t) (out.tris.size() / 12); info[2] = out.top_nodes; info[3] = (uint32_t) out.depth;
    if (nodes_out) { if (nodes_cap < out.nodes.size()) return 2; std::memcpy(nodes_out, out.nodes.data(), out.nodes.size() * sizeof(float)); }
    if (tris_out) { if (tris_cap < out.tris.size()) return 2; if (!out.tris.empty()) std::memcpy(tris_out, out.tris.data(), out.tris.size() * sizeof(float)); }
    return 0;
}
