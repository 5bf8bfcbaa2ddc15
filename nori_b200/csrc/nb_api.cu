// nb_api.cu -- C-ABI of libnori_b200.so (see include/nori_b200.h for the contract and reference citations).
//
// No CPU fallback exists in this file: every entry point either runs the sm_100a kernels or fails with an
// error.  The oracle (oracle/) is never linked or called from here.
#include "nb_ctx.h"
#include "nb_bvh.h"
#include "nb_kernels.cuh"
#include "nb_lbvh.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace nbi {
thread_local std::string g_err;
int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return 1;
}
}  // namespace nbi
using nbi::fail; using nbi::HostMesh; using nbi::DevBuf;

namespace nbm {     // nb_multi.inl (end of this file)
int replicate_scene(nb_ctx *c, bool with_header);
int render_group(nb_ctx *c, float4 *film, cudaStream_t s0, nb_stats *st);
void release_group(nb_ctx *c);
bool grouped(const nb_ctx *c);
int render_group_stats(nb_ctx *c, cudaStream_t s0, nb_stats *st);
}

namespace {

int tiles_for(const nb_ctx *c, int rank, int nranks, int *ntx_out, int *nty_out) {
    int ntx = (c->W + NB_BLOCK_SIZE - 1) / NB_BLOCK_SIZE, nty = (c->H + NB_BLOCK_SIZE - 1) / NB_BLOCK_SIZE;
    if (ntx_out) *ntx_out = ntx;
    if (nty_out) *nty_out = nty;
    int total = ntx * nty;
    return total > rank ? (total - rank + nranks - 1) / nranks : 0;
}

int ensure_device(nb_ctx *c) { CK(cudaSetDevice(c->device)); return 0; }

// Numbering of the image's 32x32 tiles for a group of `nranks` GPUs: tile_tab[t] = bx | by << 16 (nb_kernels.cuh: tile_xy), GPU
// t % nranks owns tile t.  Ownership follows the Latin pattern (bx + shift * by) % nranks, shift = 3 (5 or 7 when 3 divides nranks) -- every row and every column of tiles
// deals its tiles evenly to all ranks, the pattern with the lowest load imbalance on the BASELINE frames among row-major,
// diagonal, Z-order and hashed assignments -- and rank r's k-th tile (row by row) is tile k * nranks + r.
// The few tiles by which the pattern misses the counts the numbering implies (rank r owns ceil((total - r) / nranks) tiles)
// move from the ranks with a surplus (their last tiles) to the ranks with a deficit.  oracle.c and multigpu.py build the same.
void build_tile_order(int ntx, int nty, int nranks, std::vector<uint32_t> &tab) {
    const int total = ntx * nty;
    const int shift = nranks % 3 ? 3 : nranks % 5 ? 5 : 7;      // coprime to the group size (groups of up to 16): no rank owns a column
    std::vector<std::vector<uint32_t>> lists((size_t) nranks);
    for (int by = 0; by < nty; ++by) for (int bx = 0; bx < ntx; ++bx) lists[(size_t) ((bx + shift * by) % nranks)].push_back((uint32_t) bx | ((uint32_t) by << 16));
    std::vector<uint32_t> pool;
    auto target = [&](int r) { return total > r ? (size_t) ((total - r + nranks - 1) / nranks) : (size_t) 0; };
    for (int r = 0; r < nranks; ++r) while (lists[(size_t) r].size() > target(r)) { pool.push_back(lists[(size_t) r].back()); lists[(size_t) r].pop_back(); }
    size_t head = 0;
    for (int r = 0; r < nranks; ++r) while (lists[(size_t) r].size() < target(r)) lists[(size_t) r].push_back(pool[head++]);
    tab.assign((size_t) total, 0u);
    for (int r = 0; r < nranks; ++r) for (size_t q = 0; q < lists[(size_t) r].size(); ++q) tab[q * (size_t) nranks + (size_t) r] = lists[(size_t) r][q];
}

// Work units = (owned tile, 8x4 pixel patch, chunk of samples).  Chunk 8 renders 11 % faster than chunk 1 (a warp stays on its
// 32 pixels, its walks share their nodes in L1) but a frame must END on small units, and a unit must stay a small part of
// the frame (profiles/r2_call7_work_unit_size.txt, r2_call8_static_guided_ab.txt).  GUIDED schedule: the first `guided` percent
// of the samples go in COARSE units -- the largest power of two <= 8 that still leaves >= 12 coarse units per resident warp
// -- the rest in FINE units sized to leave >= 16 per warp.  Small shares (8 GPUs on the 9 ms frame) degenerate to the plain
// fine schedule.  An explicit `chunk` option switches the schedule off.  Pure function of its arguments (nb_debug_unit_plan
// runs it on the host); render_kernel decodes unit u as: u < split_units -> coarse (samples [0, split_sample) in chunks of
// chunk_a), else fine (u - split_units; samples [split_sample, spp) in chunks of chunk); then patch = u % 32,
// chunk id = (u / 32) % nchunks, tile slot = (u / 32) / nchunks.
struct UnitPlan { uint32_t chunk, nchunks, split_sample, chunk_a, nchunks_a, split_units, n_units; };

bool plan_units(int n_my_tiles, uint32_t spp, int64_t warps, int64_t opt_chunk, int64_t opt_guided, int64_t opt_coarse, UnitPlan &out) {
    int64_t chunk = opt_chunk;
    uint32_t split_sample = 0, coarse = 1;
    if (chunk <= 0) {
        const int64_t guided = std::max<int64_t>(0, std::min<int64_t>(100, opt_guided));
        const int64_t want = opt_coarse > 0 ? opt_coarse : 8;
        for (int64_t cc = want; cc >= 2 && !split_sample; cc /= 2) {
            const uint32_t ss = (uint32_t) ((uint64_t) spp * (uint64_t) guided / 100 / (uint64_t) cc) * (uint32_t) cc;
            if (ss >= (uint32_t) cc && (int64_t) n_my_tiles * 32 * (ss / cc) >= 12 * warps) { split_sample = ss; coarse = (uint32_t) cc; }
        }
        const int64_t units_fine = (int64_t) n_my_tiles * 32 * (spp - split_sample);
        chunk = std::max<int64_t>(1, std::min<int64_t>(8, units_fine / (16 * warps)));
        if (split_sample && chunk >= (int64_t) coarse) { split_sample = 0; chunk = std::max<int64_t>(1, std::min<int64_t>(8, (int64_t) n_my_tiles * 32 * spp / (16 * warps))); }
    }
    const uint32_t spp_fine = spp - split_sample;
    out.chunk = (uint32_t) std::max<int64_t>(1, std::min<int64_t>(chunk, (int64_t) std::max<uint32_t>(spp_fine, 1u)));
    out.nchunks = spp_fine ? (spp_fine + out.chunk - 1) / out.chunk : 0;
    out.split_sample = split_sample; out.chunk_a = coarse; out.nchunks_a = split_sample / coarse;
    const unsigned long long units_a = (unsigned long long) n_my_tiles * 32ULL * out.nchunks_a;
    const unsigned long long units = units_a + (unsigned long long) n_my_tiles * 32ULL * out.nchunks;
    if (units > 0xf0000000ULL) return false;      // (+ one claim per warp past the end stays below 2^32)
    out.split_units = (uint32_t) units_a;
    out.n_units = (uint32_t) units;
    return true;
}

int ensure_tile_table(nb_ctx *c, int nranks) {
    if (c->tile_tab_d && c->tab_W == c->W && c->tab_H == c->H && c->tab_N == nranks) return 0;
    const int ntx = (c->W + NB_BLOCK_SIZE - 1) / NB_BLOCK_SIZE, nty = (c->H + NB_BLOCK_SIZE - 1) / NB_BLOCK_SIZE;
    build_tile_order(ntx, nty, nranks, c->tile_tab_h);
    if (c->tile_tab_d) cudaFree(c->tile_tab_d);
    c->tile_tab_d = nullptr;
    CK(cudaMalloc(&c->tile_tab_d, sizeof(uint32_t) * std::max<size_t>(c->tile_tab_h.size(), 1)));
    CK(cudaMemcpyAsync(c->tile_tab_d, c->tile_tab_h.data(), sizeof(uint32_t) * c->tile_tab_h.size(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->tab_W = c->W; c->tab_H = c->H; c->tab_N = nranks;
    return 0;
}

int fill_scene(nb_ctx *c, nb::SceneDev &sc) {
    if (!c->built) return fail("nb_build_accel has not been called");
    sc.nodes = c->nodes.d; sc.tris = c->tris.d; sc.faces = c->faces.d; sc.verts = c->verts.d; sc.normals = c->normals.d;
    sc.uvs = c->uvs.d; sc.meshes = c->dmeshes.d; sc.emitter_cdf = c->cdf.d; sc.emitters = c->emitters.d;
    sc.n_emitters = (int32_t) c->emitters.n; sc.n_nodes = c->n_nodes; sc.n_prims = c->n_prims;
    return 0;
}

constexpr size_t kArenaAlign = 256, kArenaSlack = 64 * 256;      // slack: a group of up to 64 ranks pads the arena to N equal shards

// offsets of the nine tables inside the arena, from their element counts (the same on every rank of a group)
size_t arena_layout(const size_t counts[9], size_t offs[9]) {
    const size_t elem[9] = { sizeof(float4), sizeof(float4), sizeof(float4), sizeof(float4), sizeof(float2), sizeof(uint4), sizeof(nb::DevMesh), sizeof(float), sizeof(int32_t) };
    size_t off = 0;
    for (int i = 0; i < 9; ++i) { offs[i] = off; off += (counts[i] * elem[i] + kArenaAlign - 1) / kArenaAlign * kArenaAlign; }
    return off;
}

void arena_counts(const nb_ctx *c, size_t counts[9]) {
    const size_t n[9] = { c->nodes.n, c->tris.n, c->verts.n, c->normals.n, c->uvs.n, c->faces.n, c->dmeshes.n, c->cdf.n, c->emitters.n };
    for (int i = 0; i < 9; ++i) counts[i] = n[i];
}

void arena_views(nb_ctx *c, const size_t counts[9], const size_t offs[9]) {
    c->nodes.set_view(c->arena_d, c->arena_h, offs[0], counts[0]); c->tris.set_view(c->arena_d, c->arena_h, offs[1], counts[1]);
    c->verts.set_view(c->arena_d, c->arena_h, offs[2], counts[2]); c->normals.set_view(c->arena_d, c->arena_h, offs[3], counts[3]);
    c->uvs.set_view(c->arena_d, c->arena_h, offs[4], counts[4]); c->faces.set_view(c->arena_d, c->arena_h, offs[5], counts[5]);
    c->dmeshes.set_view(c->arena_d, c->arena_h, offs[6], counts[6]); c->cdf.set_view(c->arena_d, c->arena_h, offs[7], counts[7]);
    c->emitters.set_view(c->arena_d, c->arena_h, offs[8], counts[8]);
}

void arena_release(nb_ctx *c) {
    c->nodes.release(); c->tris.release(); c->verts.release(); c->normals.release(); c->uvs.release(); c->faces.release();
    c->dmeshes.release(); c->cdf.release(); c->emitters.release();
    if (c->arena_d) cudaFree(c->arena_d);
    if (c->arena_h) cudaFreeHost(c->arena_h);
    c->arena_d = nullptr; c->arena_h = nullptr; c->arena_bytes = 0;
}

// Moves the nine separately allocated tables of a finished build (device data + pinned mirrors) into one arena.
int arena_pack(nb_ctx *c) {
    size_t counts[9], offs[9];
    arena_counts(c, counts);
    const size_t total = arena_layout(counts, offs);
    char *ad = nullptr, *ah = nullptr;
    CK(cudaMalloc(&ad, total + kArenaSlack));
    if (cudaMallocHost(&ah, total + kArenaSlack) != cudaSuccess) { cudaFree(ad); return fail("cudaMallocHost(%zu) failed", total); }
    memset(ah, 0, total + kArenaSlack);
    cudaStream_t s = c->stream;
    cudaError_t e = cudaMemsetAsync(ad, 0, total + kArenaSlack, s);
#define PK(buf, i) do { if (e == cudaSuccess && c->buf.n) { e = cudaMemcpyAsync(ad + offs[i], c->buf.d, c->buf.bytes(), cudaMemcpyDeviceToDevice, s); \
                        memcpy(ah + offs[i], c->buf.h, c->buf.bytes()); } } while (0)
    PK(nodes, 0); PK(tris, 1); PK(verts, 2); PK(normals, 3); PK(uvs, 4); PK(faces, 5); PK(dmeshes, 6); PK(cdf, 7); PK(emitters, 8);
#undef PK
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { cudaFree(ad); cudaFreeHost(ah); return fail("arena_pack failed: %s", cudaGetErrorString(e)); }
    arena_release(c);
    c->arena_d = ad; c->arena_h = ah; c->arena_bytes = total;
    arena_views(c, counts, offs);
    return 0;
}

template <int INTEG>
void launch_render(const nb::RenderParams &P, bool count, bool block_mode, int grid, size_t smem, cudaStream_t s) {
    if (block_mode) {
        int g = (P.n_my_tiles + 31) / 32;
        if (count) nb::render_block_mode_kernel<INTEG, true><<<g, 32, 0, s>>>(P);
        else nb::render_block_mode_kernel<INTEG, false><<<g, 32, 0, s>>>(P);
    } else {
        if (count) nb::render_kernel<INTEG, true, false><<<grid, 128, smem, s>>>(P);
        else if (P.smem_nodes > 0) nb::render_kernel<INTEG, false, true><<<grid, 128, smem, s>>>(P);
        else nb::render_kernel<INTEG, false, false><<<grid, 128, smem, s>>>(P);
    }
}

template <int INTEG>
cudaError_t occupancy(int *blocks, bool count, bool tma, size_t smem) {
    auto kc = nb::render_kernel<INTEG, true, false>;
    auto kt = nb::render_kernel<INTEG, false, true>;
    auto kp = nb::render_kernel<INTEG, false, false>;
    auto k = count ? kc : (tma ? kt : kp);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, k, 128, smem);
}

// Renders this context's tiles into c->blocks (or blocks_out) on stream s.
int render_blocks(nb_ctx *c, float4 *blocks_out, cudaStream_t s, nb_stats *st, int *n_tiles_out) {
    if (ensure_device(c)) return 1;
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    nb::RenderParams P;
    memset(&P, 0, sizeof P);
    if (fill_scene(c, P.sc)) return 1;
    if (c->integ.type < NB_INT_NORMALS || c->integ.type > NB_INT_SIMPLE) return fail("unsupported integrator type %d (no CPU fallback)", c->integ.type);
    if (c->integ.type == NB_INT_SIMPLE && !c->have_light) return fail("the simple integrator needs nb_set_point_light");
    memcpy(P.light_pos, c->light_pos, sizeof P.light_pos); memcpy(P.light_energy, c->light_energy, sizeof P.light_energy);
    memcpy(P.s2c, c->s2c, sizeof P.s2c); memcpy(P.c2w, c->c2w, sizeof P.c2w);
    P.W = c->W; P.H = c->H; P.invW = 1.0f / (float) c->W; P.invH = 1.0f / (float) c->H;   // cwiseInverse, ref: src/perspective.cpp:27
    P.nearClip = c->nearClip; P.farClip = c->farClip;
    memcpy(P.ftable, c->ftable, sizeof P.ftable);
    P.fradius = c->fradius; P.lookup = NB_FILTER_RESOLUTION / c->fradius; P.border = c->border;   // ref: src/block.cpp:20,27
    const uint32_t spp = c->prog_active ? c->prog_pass : c->spp;       // a progressive pass renders prog_pass sample streams from prog_done on
    P.spp = spp; P.sample_offset = c->prog_active ? c->prog_done : 0u; P.seed_mode = c->seed_mode; P.seed = c->seed;
    P.integrator = c->integ.type; P.rr_start = c->integ.rr_start > 0 ? c->integ.rr_start : 3;
    P.max_depth = c->integ.max_depth > 0 ? c->integ.max_depth : (1 << 20);
    P.tile_rank = c->tile_rank; P.tile_nranks = c->tile_nranks;
    P.n_my_tiles = tiles_for(c, c->tile_rank, c->tile_nranks, &P.ntx, &P.nty);
    if (ensure_tile_table(c, c->tile_nranks)) return 1;
    P.tile_tab = c->tile_tab_d;
    P.block_edge = NB_BLOCK_SIZE + 2 * c->border;
    if (P.block_edge > nb::kBlockEdgeMax) return fail("filter radius %.3f too large (border %d > 8)", c->fradius, c->border);
    UnitPlan plan;
    if (!plan_units(P.n_my_tiles, spp, (int64_t) c->sm_count * NB_MIN_BLOCKS * 4, c->opt_chunk, c->opt_guided, c->opt_coarse, plan)) return fail("too many work units");
    P.chunk = plan.chunk; P.nchunks = plan.nchunks; P.split_sample = plan.split_sample; P.chunk_a = plan.chunk_a; P.nchunks_a = plan.nchunks_a;
    P.split_units = plan.split_units; P.n_units = plan.n_units;
    const size_t blk_elems = (size_t) P.n_my_tiles * P.block_edge * P.block_edge;
    if (!blocks_out) {
        if (blk_elems > c->blocks_cap) {
            if (c->blocks) cudaFree(c->blocks);
            c->blocks = nullptr; c->blocks_cap = 0;
            CK(cudaMalloc(&c->blocks, sizeof(float4) * (blk_elems ? blk_elems : 1)));
            c->blocks_cap = blk_elems;
        }
        blocks_out = c->blocks;
    }
    P.blocks = blocks_out;
    P.counters = c->counters;
    // Reference per-block streams are sequential by construction; only `normals` consumes a fixed number of draws per
    // sample (4), which lets the parallel kernel jump to each sample's stream position (pcg32 skip-ahead).
    P.block_stream_skip = (c->seed_mode == NB_SEED_PER_BLOCK && c->integ.type == NB_INT_NORMALS) ? 1 : 0;
    const bool block_mode = c->seed_mode == NB_SEED_PER_BLOCK && !P.block_stream_skip;
    const bool count = c->opt_count != 0;
    P.smem_nodes = (block_mode || count) ? 0 : (int) std::min<int64_t>(std::min<int64_t>(c->opt_smem_nodes, c->top_nodes), 3400);
    const size_t smem = (size_t) P.smem_nodes * 64;

    int occ = 0;
    cudaError_t oe = cudaSuccess;
    switch (c->integ.type) {
        case 0: oe = occupancy<0>(&occ, count, P.smem_nodes > 0, smem); break; case 1: oe = occupancy<1>(&occ, count, P.smem_nodes > 0, smem); break;
        case 2: oe = occupancy<2>(&occ, count, P.smem_nodes > 0, smem); break; case 3: oe = occupancy<3>(&occ, count, P.smem_nodes > 0, smem); break;
        case 4: oe = occupancy<4>(&occ, count, P.smem_nodes > 0, smem); break; case 5: oe = occupancy<5>(&occ, count, P.smem_nodes > 0, smem); break;
        default: oe = occupancy<6>(&occ, count, P.smem_nodes > 0, smem); break;
    }
    if (oe != cudaSuccess) return fail("occupancy query failed: %s", cudaGetErrorString(oe));
    if (occ < 1) return fail("render kernel does not fit on an SM (smem %zu B)", smem);
    if (c->opt_blocks_per_sm > 0) occ = (int) std::min<int64_t>(occ, c->opt_blocks_per_sm);
    const int grid = c->sm_count * occ;    // persistent: a whole number of CTAs per SM

    CK(cudaEventRecord(c->ev[0], s));
    CK(cudaMemsetAsync(c->counters, 0, sizeof(unsigned long long) * 8, s));
    if (blk_elems && !c->prog_active) CK(cudaMemsetAsync(blocks_out, 0, sizeof(float4) * blk_elems, s));   // a progressive pass accumulates
    CK(cudaEventRecord(c->ev[1], s));
    // L2 warm-up (inside the timed region): only when the walk's arrays fit L2 with room to spare, and never for the
    // one-thread-per-block plumbing mode
    const size_t walk_bytes = c->nodes.bytes() + c->tris.bytes();
    if (P.n_my_tiles > 0 && !block_mode && c->opt_prefetch != 0 && walk_bytes > 0 && walk_bytes <= (size_t) 96 << 20) {
        nb::l2_prefetch_kernel<<<std::min(c->sm_count * 4, (int) ((walk_bytes / 4096 + 127) / 128 + 1)), 128, 0, s>>>(
            reinterpret_cast<const char *>(c->nodes.d), c->nodes.bytes(), reinterpret_cast<const char *>(c->tris.d), c->tris.bytes());
        CK(cudaGetLastError());
    }
    if (P.n_my_tiles > 0) {
        switch (c->integ.type) {
            case 0: launch_render<0>(P, count, block_mode, grid, smem, s); break;
            case 1: launch_render<1>(P, count, block_mode, grid, smem, s); break;
            case 2: launch_render<2>(P, count, block_mode, grid, smem, s); break;
            case 3: launch_render<3>(P, count, block_mode, grid, smem, s); break;
            case 4: launch_render<4>(P, count, block_mode, grid, smem, s); break;
            case 5: launch_render<5>(P, count, block_mode, grid, smem, s); break;
            default: launch_render<6>(P, count, block_mode, grid, smem, s); break;
        }
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(c->ev[2], s));
    if (n_tiles_out) *n_tiles_out = P.n_my_tiles;
    if (st) {
        memset(st, 0, sizeof *st);
        // samples rendered by this context
        unsigned long long ns = 0;
        for (int k = 0; k < P.n_my_tiles; ++k) {
            const uint32_t tv = c->tile_tab_h[(size_t) (c->tile_rank + k * c->tile_nranks)]; const int bx = (int) (tv & 0xffffu), by = (int) (tv >> 16);
            ns += (unsigned long long) std::min(32, c->W - bx * 32) * std::min(32, c->H - by * 32);
        }
        st->samples = ns * spp;
        st->launches = P.n_my_tiles > 0 ? 1 : 0;
    }
    return 0;
}

// ---- wavefront engine (nb_wave.cu): launchers with C linkage, RenderParams passed as bytes
extern "C" cudaError_t nb_wv_launch_logic(const void *params, size_t bytes, int integ, cudaStream_t s);
extern "C" cudaError_t nb_wv_launch_trace(const void *params, size_t bytes, int count, int grid, cudaStream_t s);
extern "C" cudaError_t nb_wv_occupancy(int count, int *blocks_trace);
extern "C" int nb_wv_columns(void);

// Same contract as render_blocks.  A pool of wf_pool path slots lives in device memory; one ITERATION is one logic launch
// (every slot: splat + regenerate, or one integrator step; rays go to the extension / occlusion queues) and one trace launch
// (persistent warps with dynamic fetch drain both queues).  The host enqueues `wf_check` iterations at a time and then reads
// the live-slot counter; the loop ends when no slot holds a path.  This call therefore synchronises with the stream.
int render_blocks_wave(nb_ctx *c, float4 *blocks_out, cudaStream_t s, nb_stats *st, int *n_tiles_out) {
    if (c->seed_mode == NB_SEED_PER_BLOCK) return render_blocks(c, blocks_out, s, st, n_tiles_out);   // sequential streams: fused engine
    if (ensure_device(c)) return 1;
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    nb::RenderParams P;
    memset(&P, 0, sizeof P);
    if (fill_scene(c, P.sc)) return 1;
    if (c->integ.type < NB_INT_NORMALS || c->integ.type > NB_INT_SIMPLE) return fail("unsupported integrator type %d (no CPU fallback)", c->integ.type);
    if (c->integ.type == NB_INT_SIMPLE && !c->have_light) return fail("the simple integrator needs nb_set_point_light");
    memcpy(P.light_pos, c->light_pos, sizeof P.light_pos); memcpy(P.light_energy, c->light_energy, sizeof P.light_energy);
    memcpy(P.s2c, c->s2c, sizeof P.s2c); memcpy(P.c2w, c->c2w, sizeof P.c2w);
    P.W = c->W; P.H = c->H; P.invW = 1.0f / (float) c->W; P.invH = 1.0f / (float) c->H;
    P.nearClip = c->nearClip; P.farClip = c->farClip;
    memcpy(P.ftable, c->ftable, sizeof P.ftable);
    P.fradius = c->fradius; P.lookup = NB_FILTER_RESOLUTION / c->fradius; P.border = c->border;
    const uint32_t spp = c->prog_active ? c->prog_pass : c->spp;       // a progressive pass renders prog_pass sample streams from prog_done on
    P.spp = spp; P.sample_offset = c->prog_active ? c->prog_done : 0u; P.seed_mode = c->seed_mode; P.seed = c->seed;
    P.integrator = c->integ.type; P.rr_start = c->integ.rr_start > 0 ? c->integ.rr_start : 3;
    P.max_depth = c->integ.max_depth > 0 ? c->integ.max_depth : (1 << 20);
    P.tile_rank = c->tile_rank; P.tile_nranks = c->tile_nranks;
    P.n_my_tiles = tiles_for(c, c->tile_rank, c->tile_nranks, &P.ntx, &P.nty);
    if (ensure_tile_table(c, c->tile_nranks)) return 1;
    P.tile_tab = c->tile_tab_d;
    P.block_edge = NB_BLOCK_SIZE + 2 * c->border;
    if (P.block_edge > nb::kBlockEdgeMax) return fail("filter radius %.3f too large (border %d > 8)", c->fradius, c->border);
    P.wf_chunk = (uint32_t) std::max<int64_t>(1, std::min<int64_t>(c->opt_chunk > 0 ? c->opt_chunk : 8, std::min<int64_t>(8, spp)));
    P.chunk = P.wf_chunk;
    P.nchunks = (spp + P.wf_chunk - 1) / P.wf_chunk;
    P.wf_total = (unsigned long long) P.n_my_tiles * 32ull * P.nchunks * 32ull * P.wf_chunk;
    // pool: as many slots as asked for, never more than there are samples, a multiple of 128
    unsigned long long pool = (unsigned long long) std::max<int64_t>(128, c->opt_wf_pool);
    pool = std::min<unsigned long long>(pool, std::max<unsigned long long>(P.wf_total, 128ull));
    pool = (pool + 127ull) / 128ull * 128ull;
    P.wf_pool = (uint32_t) pool;
    const int ncols = nb_wv_columns();
    if (pool > c->wf_cap) {
        if (c->wf_cols) cudaFree(c->wf_cols);
        if (c->wf_ext) cudaFree(c->wf_ext);
        if (c->wf_shadow) cudaFree(c->wf_shadow);
        c->wf_cols = nullptr; c->wf_ext = nullptr; c->wf_shadow = nullptr; c->wf_cap = 0;
        CK(cudaMalloc(&c->wf_cols, sizeof(float4) * (size_t) ncols * pool));
        CK(cudaMalloc(&c->wf_ext, sizeof(uint32_t) * pool));
        CK(cudaMalloc(&c->wf_shadow, sizeof(float4) * 3 * pool));
        c->wf_cap = pool;
    }
    if (!c->wf_ctr) { CK(cudaMalloc(&c->wf_ctr, sizeof(uint32_t) * 8)); CK(cudaMallocHost(&c->wf_ctr_h, sizeof(uint32_t) * 8)); }
    P.wf_cols = c->wf_cols; P.wf_ext = c->wf_ext; P.wf_ctr = c->wf_ctr; P.occ_queue = c->wf_shadow; P.occ_capacity = P.wf_pool;
    P.tail_lanes = (int32_t) c->opt_occ_tail;
    const size_t blk_elems = (size_t) P.n_my_tiles * P.block_edge * P.block_edge;
    if (!blocks_out) {
        if (blk_elems > c->blocks_cap) {
            if (c->blocks) cudaFree(c->blocks);
            c->blocks = nullptr; c->blocks_cap = 0;
            CK(cudaMalloc(&c->blocks, sizeof(float4) * (blk_elems ? blk_elems : 1)));
            c->blocks_cap = blk_elems;
        }
        blocks_out = c->blocks;
    }
    P.blocks = blocks_out;
    P.counters = c->counters;
    const bool count = c->opt_count != 0;
    int occ_t = 0;
    cudaError_t oe = nb_wv_occupancy(count ? 1 : 0, &occ_t);
    if (oe != cudaSuccess) return fail("occupancy query failed: %s", cudaGetErrorString(oe));
    if (occ_t < 1) return fail("wavefront trace kernel does not fit on an SM");
    if (c->opt_blocks_per_sm > 0) occ_t = (int) std::min<int64_t>(occ_t, c->opt_blocks_per_sm);
    // persistent trace grid: a whole number of CTAs per SM, but no more warps than rays to fetch 32 at a time
    const int grid_t = (int) std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long) c->sm_count * occ_t, (pool + 127ull) / 128ull));

    CK(cudaEventRecord(c->ev[0], s));
    CK(cudaMemsetAsync(c->counters, 0, sizeof(unsigned long long) * 8, s));
    CK(cudaMemsetAsync(c->wf_ctr, 0, sizeof(uint32_t) * 8, s));
    CK(cudaMemsetAsync(c->wf_cols + (size_t) 4 * pool, 0, sizeof(float4) * pool, s));      // column 4 holds the slot state: all empty
    if (blk_elems && !c->prog_active) CK(cudaMemsetAsync(blocks_out, 0, sizeof(float4) * blk_elems, s));   // a progressive pass accumulates
    CK(cudaEventRecord(c->ev[1], s));
    unsigned long long launches = 0;
    const int check = (int) std::max<int64_t>(1, c->opt_wf_check);
    bool done = P.n_my_tiles == 0;
    for (long long it = 0; !done; ) {
        for (int k = 0; k < check; ++k, ++it) {
            CK(cudaMemsetAsync(c->wf_ctr + 2, 0, sizeof(uint32_t) * 5, s));      // queue lengths, fetch cursors, live slots; the sample cursor keeps running
            cudaError_t e = nb_wv_launch_logic(&P, sizeof P, c->integ.type, s);
            if (e != cudaSuccess) return fail("wavefront logic launch failed: %s", cudaGetErrorString(e));
            e = nb_wv_launch_trace(&P, sizeof P, count ? 1 : 0, grid_t, s);
            if (e != cudaSuccess) return fail("wavefront trace launch failed: %s", cudaGetErrorString(e));
            launches += 2;
        }
        CK(cudaMemcpyAsync(c->wf_ctr_h, c->wf_ctr, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (c->wf_ctr_h[7]) return fail("device watchdog fired in the wavefront trace kernel (results invalid)");
        if (c->wf_ctr_h[6] == 0) done = true;                  // no slot holds a path after the last logic pass
        if (it > 4000000) return fail("wavefront engine: no progress after %lld iterations", it);
    }
    CK(cudaEventRecord(c->ev[2], s));
    if (n_tiles_out) *n_tiles_out = P.n_my_tiles;
    if (st) {
        memset(st, 0, sizeof *st);
        unsigned long long ns = 0;
        for (int k = 0; k < P.n_my_tiles; ++k) {
            const uint32_t tv = c->tile_tab_h[(size_t) (c->tile_rank + k * c->tile_nranks)]; const int bx = (int) (tv & 0xffffu), by = (int) (tv >> 16);
            ns += (unsigned long long) std::min(32, c->W - bx * 32) * std::min(32, c->H - by * 32);
        }
        st->samples = ns * spp;
        st->launches = launches;
    }
    return 0;
}

int render_tiles(nb_ctx *c, float4 *blocks_out, cudaStream_t s, nb_stats *st, int *n_tiles_out) {
    if (c->opt_engine == 2) return render_blocks_wave(c, blocks_out, s, st, n_tiles_out);
    return render_blocks(c, blocks_out, s, st, n_tiles_out);
}

int merge(nb_ctx *c, const float4 *blocks, int n_tiles, int rank, int nranks, float4 *film, cudaStream_t s) {
    if (n_tiles <= 0) return 0;
    int edge = NB_BLOCK_SIZE + 2 * c->border;
    long long total = (long long) n_tiles * edge * edge;
    int grid = (int) ((total + 255) / 256);
    if (ensure_tile_table(c, nranks)) return 1;
    nb::merge_blocks_kernel<<<grid, 256, 0, s>>>(blocks, n_tiles, rank, nranks, c->tile_tab_d, c->W, c->H, c->border, edge, film);
    CK(cudaGetLastError());
    return 0;
}

int finish_stats(nb_ctx *c, cudaStream_t s, nb_stats *st, int extra_launches) {
    CK(cudaMemcpyAsync(c->counters_h, c->counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(c->ev[3], s));
    CK(cudaStreamSynchronize(s));
    if (c->counters_h[6] >> 63) return fail("device watchdog fired: a persistent loop of an experimental engine did not terminate (results invalid)");
    if (st) {
        st->rays = c->counters_h[1]; st->node_visits = c->counters_h[2]; st->tri_tests = c->counters_h[3];
        st->hits_shaded = c->counters_h[4];
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); st->kernel_ms = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[3])); st->total_ms = ms;
        st->launches += extra_launches;
    }
    return 0;
}

// Device LBVH build (nb_lbvh.cuh).  verts / faces must already be on the device.  Fills c->nodes / c->tris (device + pinned
// mirrors) and the node count.
int build_lbvh_device(nb_ctx *c, size_t nf, float pad, const float clo[3], const float chi[3]) {
    const int n = (int) nf;
    cudaStream_t s = c->stream;
    nb::LbvhScratch L;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(L.keys); cudaFree(L.keys_sorted); cudaFree(L.leaf_lo); cudaFree(L.leaf_hi); cudaFree(L.node_lo); cudaFree(L.node_hi);
        cudaFree(L.children); cudaFree(L.parent); cudaFree(L.range); cudaFree(L.arrive); cudaFree(L.emit); cudaFree(L.emit_index); cudaFree(L.cub_tmp);
    };
#define LCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail("%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
    LCK(cudaMalloc(&L.keys, sizeof(uint64_t) * n)); LCK(cudaMalloc(&L.keys_sorted, sizeof(uint64_t) * n));
    LCK(cudaMalloc(&L.leaf_lo, sizeof(float4) * n)); LCK(cudaMalloc(&L.leaf_hi, sizeof(float4) * n));
    LCK(cudaMalloc(&L.node_lo, sizeof(float4) * n)); LCK(cudaMalloc(&L.node_hi, sizeof(float4) * n));
    LCK(cudaMalloc(&L.children, sizeof(int2) * n)); LCK(cudaMalloc(&L.parent, sizeof(int) * 2 * n)); LCK(cudaMalloc(&L.range, sizeof(int2) * n));
    LCK(cudaMalloc(&L.arrive, sizeof(unsigned) * n)); LCK(cudaMalloc(&L.emit, sizeof(unsigned) * n)); LCK(cudaMalloc(&L.emit_index, sizeof(unsigned) * n));
    size_t sort_bytes = 0, scan_bytes = 0;
    LCK(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, L.keys, L.keys_sorted, n, 0, 62, s));
    LCK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, L.emit, L.emit_index, n - 1, s));
    L.cub_bytes = std::max(sort_bytes, scan_bytes);
    LCK(cudaMalloc(&L.cub_tmp, L.cub_bytes));
    const int B = 256, G = (n + B - 1) / B;
    float3 lo3 = make_float3(clo[0], clo[1], clo[2]);
    float3 inv3 = make_float3(chi[0] > clo[0] ? 1.0f / (chi[0] - clo[0]) : 0.f, chi[1] > clo[1] ? 1.0f / (chi[1] - clo[1]) : 0.f,
                              chi[2] > clo[2] ? 1.0f / (chi[2] - clo[2]) : 0.f);
    nb::lbvh_keys_kernel<<<G, B, 0, s>>>(c->verts.d, c->faces.d, (unsigned) n, lo3, inv3, L.keys);
    LCK(cub::DeviceRadixSort::SortKeys(L.cub_tmp, L.cub_bytes, L.keys, L.keys_sorted, n, 0, 62, s));
    nb::lbvh_leaf_boxes_kernel<<<G, B, 0, s>>>(c->verts.d, c->faces.d, L.keys_sorted, (unsigned) n, pad, L.leaf_lo, L.leaf_hi);
    nb::lbvh_hierarchy_kernel<<<G, B, 0, s>>>(L.keys_sorted, n, L.children, L.parent, L.range);
    int *depth_d = reinterpret_cast<int *>(c->counters);          // scratch: the counters are cleared by every render
    LCK(cudaMemsetAsync(depth_d, 0, sizeof(int), s));
    nb::lbvh_depth_kernel<<<G, B, 0, s>>>(n, L.parent, depth_d);
    int tree_depth = 0;
    LCK(cudaMemcpyAsync(&tree_depth, depth_d, sizeof(int), cudaMemcpyDeviceToHost, s));
    LCK(cudaMemsetAsync(L.arrive, 0, sizeof(unsigned) * n, s));
    nb::lbvh_fit_kernel<<<G, B, 0, s>>>(n, L.children, L.parent, L.leaf_lo, L.leaf_hi, L.node_lo, L.node_hi, L.arrive);
    nb::lbvh_mark_kernel<<<G, B, 0, s>>>(n, L.range, (int) std::max<int64_t>(1, std::min<int64_t>(8, c->opt_max_leaf)), L.emit);
    LCK(cub::DeviceScan::ExclusiveSum(L.cub_tmp, L.cub_bytes, L.emit, L.emit_index, n - 1, s));
    unsigned last_idx = 0, last_emit = 0;
    LCK(cudaMemcpyAsync(&last_idx, L.emit_index + (n - 2), sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    LCK(cudaMemcpyAsync(&last_emit, L.emit + (n - 2), sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    LCK(cudaStreamSynchronize(s));
    const unsigned nnodes = last_idx + last_emit;
    if (nnodes == 0) { cleanup(); return fail("LBVH: empty hierarchy"); }
    if (tree_depth >= nb::kStack - 1) { cleanup(); return fail("LBVH too deep (%d levels; the traversal stack holds %d)", tree_depth, nb::kStack); }
    LCK(c->nodes.alloc((size_t) nnodes * 4)); LCK(c->tris.alloc((size_t) n * 3));
    nb::lbvh_emit_nodes_kernel<<<G, B, 0, s>>>(n, L.children, L.range, L.emit, L.emit_index, L.leaf_lo, L.leaf_hi, L.node_lo, L.node_hi, c->nodes.d);
    nb::lbvh_emit_tris_kernel<<<G, B, 0, s>>>(c->verts.d, c->faces.d, L.keys_sorted, (unsigned) n, c->tris.d);
    LCK(cudaGetLastError());
    LCK(cudaMemcpyAsync(c->nodes.h, c->nodes.d, c->nodes.bytes(), cudaMemcpyDeviceToHost, s));
    LCK(cudaMemcpyAsync(c->tris.h, c->tris.d, c->tris.bytes(), cudaMemcpyDeviceToHost, s));
    LCK(cudaStreamSynchronize(s));
#undef LCK
    cleanup();
    c->n_nodes = nnodes; c->top_nodes = 0; c->bvh_depth = tree_depth;
    return rc;
}

}  // namespace

extern "C" {

const char *nb_last_error(void) { return nbi::g_err.c_str(); }
int nb_abi_version(void) { return NB_ABI_VERSION; }
int nb_node_bytes(void) { return NB_WIDE ? 80 : 64; }

nb_ctx *nb_create(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { fail("no CUDA device available (%s); libnori_b200 has no CPU fallback", cudaGetErrorString(e)); return nullptr; }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= ndev) { fail("device %d out of range (%d devices)", device, ndev); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail("cudaSetDevice(%d) failed", device); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { fail("cudaGetDeviceProperties failed"); return nullptr; }
    if (prop.major < 10) { fail("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return nullptr; }
    nb_ctx *c = new nb_ctx();
    c->device = device; c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
    ok = ok && cudaMalloc(&c->counters, sizeof(unsigned long long) * 8) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->counters_h, sizeof(unsigned long long) * 8) == cudaSuccess;
    if (!ok) { fail("context allocation failed: %s", cudaGetErrorString(cudaGetLastError())); nb_destroy(c); return nullptr; }
    // default Gaussian filter (ref: src/perspective.cpp:71-73, src/rfilter.cpp:16-30)
    for (int i = 0; i < NB_FILTER_RESOLUTION; ++i) {
        float pos = (2.0f * i) / NB_FILTER_RESOLUTION, alpha = -1.0f / (2.0f * 0.5f * 0.5f);
        c->ftable[i] = std::max(0.0f, std::exp(alpha * pos * pos) - std::exp(alpha * 2.0f * 2.0f));
    }
    c->ftable[NB_FILTER_RESOLUTION] = 0.0f;
    return c;
}

void nb_destroy(nb_ctx *c) {
    if (!c) return;
    for (nb_ctx *f : c->followers) { f->leader = nullptr; nb_destroy(f); }
    c->followers.clear();
    cudaSetDevice(c->device);
    nbm::release_group(c);
    arena_release(c);
    if (c->tile_tab_d) cudaFree(c->tile_tab_d);
    if (c->shard_h) cudaFreeHost(c->shard_h);
    if (c->blocks) cudaFree(c->blocks);
    if (c->film) cudaFree(c->film);
    if (c->wf_cols) cudaFree(c->wf_cols);
    if (c->wf_ext) cudaFree(c->wf_ext);
    if (c->wf_shadow) cudaFree(c->wf_shadow);
    if (c->wf_ctr) cudaFree(c->wf_ctr);
    if (c->wf_ctr_h) cudaFreeHost(c->wf_ctr_h);
    if (c->counters) cudaFree(c->counters);
    if (c->counters_h) cudaFreeHost(c->counters_h);
    for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int nb_add_mesh(nb_ctx *c, const float *V, uint32_t nv, const float *N, const float *UV, const uint32_t *F, uint32_t nf,
                const nb_bsdf_desc *bsdf, const nb_emitter_desc *emitter) {
    if (!c) { fail("null context"); return -1; }
    if ((!V && nv) || (!F && nf)) { fail("nb_add_mesh: null vertex/index array"); return -1; }
    for (size_t i = 0; i < (size_t) nf * 3; ++i) if (F[i] >= nv) { fail("nb_add_mesh: index %u out of range (nv=%u)", F[i], nv); return -1; }
    HostMesh m;
    m.nv = nv; m.nf = nf;
    m.V.assign(V, V + (size_t) nv * 3);
    if (N) m.N.assign(N, N + (size_t) nv * 3);
    if (UV) m.UV.assign(UV, UV + (size_t) nv * 2);
    m.F.assign(F, F + (size_t) nf * 3);
    if (bsdf) {
        if (bsdf->type < NB_BSDF_DIFFUSE || bsdf->type > NB_BSDF_MICROFACET) { fail("unsupported BSDF type %d (no CPU fallback)", bsdf->type); return -1; }
        m.bsdf = *bsdf;
    } else {   // default diffuse, ref: src/mesh.cpp:23-29, src/diffuse.cpp:19
        memset(&m.bsdf, 0, sizeof m.bsdf); m.bsdf.type = NB_BSDF_DIFFUSE; m.bsdf.albedo[0] = m.bsdf.albedo[1] = m.bsdf.albedo[2] = 0.5f;
    }
    if (emitter) {
        if (emitter->type != NB_EMITTER_NONE && emitter->type != NB_EMITTER_AREA) { fail("unsupported emitter type %d", emitter->type); return -1; }
        m.emitter = *emitter;
    } else memset(&m.emitter, 0, sizeof m.emitter);
    c->meshes.push_back(std::move(m));
    c->built = false;
    return (int) c->meshes.size() - 1;
}

int nb_clear_meshes(nb_ctx *c) {
    if (!c) return fail("null context");
    c->meshes.clear(); c->built = false;
    c->have_light = false;          // a new scene starts without the previous scene's point light
    for (nb_ctx *f : c->followers) { f->built = false; f->have_light = false; }
    return 0;
}

static int upload_local(nb_ctx *c) {        // the whole arena in ONE host->device copy
    if (!c->arena_d || !c->arena_h) return fail("no host copy of the scene on this context");
    CK(cudaMemcpyAsync(c->arena_d, c->arena_h, c->arena_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int nb_upload_scene(nb_ctx *c) {
    if (!c) return fail("null context");
    if (c->leader) return fail("nb_upload_scene: call it on the group's leader context");
    if (!c->built) return fail("nb_build_accel has not been called");
    if (ensure_device(c)) return 1;
    // a group uploads SHARDED: every rank sends 1/N of the arena over its own PCIe link, one in-place ncclAllGather over
    // NVLink completes it on every device (nb_multi.inl)
    if (nbm::grouped(c)) return nbm::replicate_scene(c, false);
    return upload_local(c);
}

static int build_accel_local(nb_ctx *c);

int nb_build_accel(nb_ctx *c) {
    if (!c) return fail("null context");
    if (c->leader) return fail("nb_build_accel: call it on the group's leader context");
    if (ensure_device(c)) return 1;
    // In a group the hierarchy is built ONCE, by rank 0, and replicated over NVLink; the other ranks only receive.
    if (nbm::grouped(c) && c->comm_rank != 0) return nbm::replicate_scene(c, true);
    if (build_accel_local(c)) return 1;
    if (nbm::grouped(c)) return nbm::replicate_scene(c, true);
    return 0;
}

#if NB_WIDE
// Collapses a binary hierarchy (nb_bvh.h layout) into the 8-wide compressed one (nb_wide.h) and makes it the context's
// nodes / triangles (device + pinned mirror).  bnodes / btris may alias the context's current mirrors.
static int to_wide(nb_ctx *c, const float *bnodes, uint32_t n_bnodes, const float *btris, uint32_t n_btris) {
    nb::WideOutput w; const char *err = nullptr;
    if (!nb::build_wide(bnodes, n_bnodes, btris, n_btris, w, &err)) return fail("wide hierarchy: %s", err ? err : "conversion failed");
    CK(c->nodes.alloc((size_t) w.nnodes * 5)); CK(c->tris.alloc(w.tris.size() / 4));
    memcpy(c->nodes.h, w.nodes.data(), w.nodes.size() * sizeof(uint32_t));
    if (!w.tris.empty()) memcpy(c->tris.h, w.tris.data(), w.tris.size() * sizeof(float));
    CK(cudaMemcpyAsync(c->nodes.d, c->nodes.h, c->nodes.bytes(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->tris.d, c->tris.h, c->tris.bytes(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->n_nodes = w.nnodes; c->top_nodes = 0; c->bvh_depth = w.depth; c->build_seconds += w.seconds;
    return 0;
}
#endif

static int build_accel_local(nb_ctx *c) {
    size_t nv = 0, nf = 0, ncdf = 0, nem = 0;
    for (auto &m : c->meshes) { nv += m.nv; nf += m.nf; if (m.emitter.type == NB_EMITTER_AREA) { ncdf += m.nf + 1; nem++; } }
    if (nf >= (1u << 28)) return fail("too many triangles (%zu)", nf);
    arena_release(c);
    bool any_n = false, any_uv = false;
    for (auto &m : c->meshes) { any_n = any_n || !m.N.empty(); any_uv = any_uv || !m.UV.empty(); }
    // normals / texture coordinates only when some mesh has them (the kernels read them behind the per-mesh flags)
    CK(c->verts.alloc(nv)); CK(c->normals.alloc(any_n ? nv : 0)); CK(c->uvs.alloc(any_uv ? nv : 0)); CK(c->faces.alloc(nf));
    CK(c->dmeshes.alloc(c->meshes.size())); CK(c->cdf.alloc(ncdf)); CK(c->emitters.alloc(nem));
    size_t vo = 0, fo = 0, co = 0, eo = 0;
    for (size_t mi = 0; mi < c->meshes.size(); ++mi) {
        const HostMesh &m = c->meshes[mi];
        for (uint32_t i = 0; i < m.nv; ++i) {
            c->verts.h[vo + i] = make_float4(m.V[3 * i], m.V[3 * i + 1], m.V[3 * i + 2], 0.f);
            if (any_n) c->normals.h[vo + i] = m.N.empty() ? make_float4(0, 0, 0, 0) : make_float4(m.N[3 * i], m.N[3 * i + 1], m.N[3 * i + 2], 0.f);
            if (any_uv) c->uvs.h[vo + i] = m.UV.empty() ? make_float2(0, 0) : make_float2(m.UV[2 * i], m.UV[2 * i + 1]);
        }
        for (uint32_t f = 0; f < m.nf; ++f)
            c->faces.h[fo + f] = make_uint4((uint32_t) vo + m.F[3 * f], (uint32_t) vo + m.F[3 * f + 1], (uint32_t) vo + m.F[3 * f + 2], (uint32_t) mi);
        nb::DevMesh &d = c->dmeshes.h[mi];
        memset(&d, 0, sizeof d);
        d.bsdf_type = m.bsdf.type; memcpy(d.albedo, m.bsdf.albedo, sizeof d.albedo);
        d.alpha = m.bsdf.alpha; d.intIOR = m.bsdf.intIOR; d.extIOR = m.bsdf.extIOR; d.ks = m.bsdf.ks;
        d.emitter_type = m.emitter.type; memcpy(d.radiance, m.emitter.radiance, sizeof d.radiance);
        d.prim_offset = (uint32_t) fo; d.nf = m.nf;
        d.flags = (m.N.empty() ? 0u : 1u) | (m.UV.empty() ? 0u : 2u);
        if (m.emitter.type == NB_EMITTER_AREA) {
            // DiscretePDF over triangle areas: ref src/mesh.cpp:31-37, include/nori/dpdf.h:40-84
            float *cdf = c->cdf.h + co;
            cdf[0] = 0.0f;
            for (uint32_t f = 0; f < m.nf; ++f) {
                const float *p0 = &m.V[3 * m.F[3 * f]], *p1 = &m.V[3 * m.F[3 * f + 1]], *p2 = &m.V[3 * m.F[3 * f + 2]];
                float e1[3] = { p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2] }, e2[3] = { p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2] };
                float cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
                float area = 0.5f * std::sqrt(cx * cx + (cy * cy + cz * cz));
                cdf[f + 1] = cdf[f] + area;
            }
            d.area_sum = cdf[m.nf];
            if (d.area_sum > 0) {
                float norm = 1.0f / d.area_sum;
                for (uint32_t f = 1; f <= m.nf; ++f) cdf[f] *= norm;
                cdf[m.nf] = 1.0f;
            }
            d.cdf_offset = (uint32_t) co;
            c->emitters.h[eo++] = (int32_t) mi;
            co += m.nf + 1;
        }
        vo += m.nv; fo += m.nf;
    }
    c->n_prims = (uint32_t) nf;
    if (c->opt_builder == 1 && nf > 8) {
        // ---- device LBVH: upload the mesh tables, then build on the GPU
        cudaStream_t s = c->stream;
#define UP(buf) CK(cudaMemcpyAsync(buf.d, buf.h, buf.bytes(), cudaMemcpyHostToDevice, s))
        UP(c->verts); UP(c->normals); UP(c->uvs); UP(c->faces); UP(c->dmeshes); UP(c->cdf); UP(c->emitters);
#undef UP
        float maxabs = 0.f, clo[3] = { INFINITY, INFINITY, INFINITY }, chi[3] = { -INFINITY, -INFINITY, -INFINITY };
        for (size_t f = 0; f < nf; ++f) {
            const uint4 fc = c->faces.h[f];
            const float4 v[3] = { c->verts.h[fc.x], c->verts.h[fc.y], c->verts.h[fc.z] };
            const float *p[3] = { &v[0].x, &v[1].x, &v[2].x };
            for (int a = 0; a < 3; ++a) {
                const float lo = std::min(p[0][a], std::min(p[1][a], p[2][a])), hi = std::max(p[0][a], std::max(p[1][a], p[2][a]));
                const float ctr = 0.5f * (lo + hi);
                clo[a] = std::min(clo[a], ctr); chi[a] = std::max(chi[a], ctr);
                maxabs = std::max(maxabs, std::max(std::fabs(lo), std::fabs(hi)));
            }
        }
        cudaEvent_t e0 = c->ev[0], e1 = c->ev[3];
        CK(cudaEventRecord(e0, s));
        if (build_lbvh_device(c, nf, 4e-6f * maxabs, clo, chi)) return 1;
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        c->build_seconds = ms * 1e-3; c->builder_used = 1;
#if NB_WIDE
        // the device builder leaves the binary layout in the pinned mirrors: collapse it on the host, replace nodes / triangles
        if (to_wide(c, reinterpret_cast<const float *>(c->nodes.h), c->n_nodes, reinterpret_cast<const float *>(c->tris.h), (uint32_t) (c->tris.n / 3))) return 1;
#endif
        if (arena_pack(c)) return 1;
        c->built = true;
        return 0;
    }
    nb::BvhInput in; in.verts = reinterpret_cast<const float *>(c->verts.h); in.faces = reinterpret_cast<const uint32_t *>(c->faces.h);
    in.nprims = (uint32_t) nf;
    nb::BvhOutput out;
    const uint32_t bfs = c->opt_bfs_nodes < 0 ? nb::kSiblingPairs : (uint32_t) c->opt_bfs_nodes;
    // on-disk hierarchy cache (nb_set_accel_cache): keyed on every vertex / index the builder reads + the build parameters
    uint64_t key = 0; bool hit = false;
    if (!c->accel_cache.empty()) { key = nb::bvh_cache_key(in, (int) c->opt_max_leaf, bfs, nb::kStack, (int) c->opt_sah_bins); hit = nb::bvh_cache_load(c->accel_cache.c_str(), key, out); }
    if (!hit) {
        nb::build_bvh(in, out, (int) c->opt_max_leaf, bfs, 0, nb::kStack, (int) c->opt_sah_bins);
        if (!c->accel_cache.empty()) nb::bvh_cache_save(c->accel_cache.c_str(), key, out);
    }
    c->accel_cache_hit = hit;
    c->n_nodes = out.nnodes; c->top_nodes = out.top_nodes; c->bvh_depth = out.depth;
    c->build_seconds = out.build_seconds; c->builder_used = 0;
    if (out.depth >= nb::kStack) return fail("BVH too deep (%d)", out.depth);
#if NB_WIDE
    if (to_wide(c, out.nodes.data(), out.nnodes, out.tris.data(), (uint32_t) (out.tris.size() / 12))) return 1;
#else
    CK(c->nodes.alloc((size_t) out.nnodes * 4)); CK(c->tris.alloc(out.tris.size() / 4));
    memcpy(c->nodes.h, out.nodes.data(), out.nodes.size() * sizeof(float));
    if (!out.tris.empty()) memcpy(c->tris.h, out.tris.data(), out.tris.size() * sizeof(float));
#endif
    {   // mesh tables to the device (nodes / tris follow inside arena_pack's copies: their device buffers are still empty)
        cudaStream_t s = c->stream;
#define UP(buf) CK(cudaMemcpyAsync(buf.d, buf.h, buf.bytes(), cudaMemcpyHostToDevice, s))
        UP(c->nodes); UP(c->tris); UP(c->verts); UP(c->normals); UP(c->uvs); UP(c->faces); UP(c->dmeshes); UP(c->cdf); UP(c->emitters);
#undef UP
        CK(cudaStreamSynchronize(s));
    }
    if (arena_pack(c)) return 1;
    c->built = true;
    return 0;
}

int nb_set_accel_cache(nb_ctx *c, const char *path) {
    if (!c) return fail("null context");
    c->accel_cache = path ? path : "";
    return 0;
}

int nb_accel_cache_hit(nb_ctx *c) { return (c && c->built && c->accel_cache_hit) ? 1 : 0; }

int nb_build_stats(nb_ctx *c, double *seconds, int *builder) {
    if (!c) return fail("null context");
    if (!c->built) return fail("nb_build_accel has not been called");
    if (seconds) *seconds = c->build_seconds;
    if (builder) *builder = c->builder_used;
    return 0;
}

int nb_set_camera(nb_ctx *c, const float s2c[16], const float c2w[16], int width, int height, float nearClip, float farClip) {
    if (!c) return fail("null context");
    if (!s2c || !c2w) return fail("nb_set_camera: null matrix");
    if (width <= 0 || height <= 0 || width > 32767 || height > 32767) return fail("invalid output size %dx%d", width, height);
    memcpy(c->s2c, s2c, sizeof c->s2c); memcpy(c->c2w, c2w, sizeof c->c2w);
    c->W = width; c->H = height; c->nearClip = nearClip; c->farClip = farClip; c->have_camera = true; c->film_valid = false;
    for (nb_ctx *f : c->followers) if (nb_set_camera(f, s2c, c2w, width, height, nearClip, farClip)) return 1;
    return 0;
}

int nb_set_filter(nb_ctx *c, const float table[NB_FILTER_RESOLUTION + 1], float radius) {
    if (!c) return fail("null context");
    if (!table) return fail("nb_set_filter: null table");
    if (!(radius > 0)) return fail("invalid filter radius %f", radius);
    int border = (int) std::ceil(radius - 0.5f);     // ref: src/block.cpp:20
    if (border > 8) return fail("filter radius %f too large", radius);
    memcpy(c->ftable, table, sizeof c->ftable);
    c->fradius = radius; c->border = border; c->film_valid = false;
    for (nb_ctx *f : c->followers) if (nb_set_filter(f, table, radius)) return 1;
    return 0;
}

int nb_set_sampler(nb_ctx *c, uint32_t spp, int seed_mode, uint64_t seed) {
    if (!c) return fail("null context");
    if (spp == 0) return fail("sampleCount must be >= 1");
    if (seed_mode != NB_SEED_PER_SAMPLE && seed_mode != NB_SEED_PER_BLOCK) return fail("unknown seed mode %d", seed_mode);
    c->spp = spp; c->seed_mode = seed_mode; c->seed = seed;
    for (nb_ctx *f : c->followers) if (nb_set_sampler(f, spp, seed_mode, seed)) return 1;
    return 0;
}

int nb_set_integrator(nb_ctx *c, const nb_integrator_desc *d) {
    if (!c || !d) return fail("null argument");
    if (d->type < NB_INT_NORMALS || d->type > NB_INT_SIMPLE) return fail("unsupported integrator type %d (no CPU fallback)", d->type);
    c->integ = *d;
    for (nb_ctx *f : c->followers) if (nb_set_integrator(f, d)) return 1;
    return 0;
}

int nb_set_point_light(nb_ctx *c, const float position[3], const float energy[3]) {
    if (!c || !position || !energy) return fail("null argument");
    memcpy(c->light_pos, position, sizeof c->light_pos); memcpy(c->light_energy, energy, sizeof c->light_energy);
    c->have_light = true;
    for (nb_ctx *f : c->followers) if (nb_set_point_light(f, position, energy)) return 1;
    return 0;
}

int nb_set_tiles(nb_ctx *c, int rank, int nranks) {
    if (!c) return fail("null context");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("invalid tile shard (%d of %d)", rank, nranks);
    if (nbm::grouped(c) && (rank != c->comm_rank || nranks != c->comm_nranks))
        return fail("nb_set_tiles: a context in a %d-GPU group renders shard (%d, %d); the group owns the tile assignment", c->comm_nranks, c->comm_rank, c->comm_nranks);
    c->tile_rank = rank; c->tile_nranks = nranks;
    return 0;
}

int nb_tile_count(nb_ctx *c, int rank, int nranks, int *ntiles, int *block_edge) {
    if (!c) return fail("null context");
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("invalid tile shard (%d of %d)", rank, nranks);
    if (ntiles) *ntiles = tiles_for(c, rank, nranks, nullptr, nullptr);
    if (block_edge) *block_edge = NB_BLOCK_SIZE + 2 * c->border;
    return 0;
}

// Host-only diagnostic (no context, no device): the work-unit plan of one render launch (plan_units above).
int nb_debug_unit_plan(int n_tiles, uint32_t spp, int64_t resident_warps, int64_t chunk, int64_t guided, int64_t coarse, uint32_t out[7]) {
    if (n_tiles < 0 || spp < 1 || resident_warps < 1 || !out) return 1;
    UnitPlan p;
    if (!plan_units(n_tiles, spp, resident_warps, chunk, guided < 0 ? 75 : guided, coarse, p)) return 2;
    out[0] = p.chunk; out[1] = p.nchunks; out[2] = p.split_sample; out[3] = p.chunk_a; out[4] = p.nchunks_a; out[5] = p.split_units; out[6] = p.n_units;
    return 0;
}

// Host-only diagnostic (no context, no device): the tile numbering nb_set_tiles / nb_render use for a group of `nranks`.
int nb_debug_tile_order(int width, int height, int nranks, uint32_t *bx_by_out, uint64_t cap) {
    if (width < 1 || height < 1 || nranks < 1) return 1;
    const int ntx = (width + NB_BLOCK_SIZE - 1) / NB_BLOCK_SIZE, nty = (height + NB_BLOCK_SIZE - 1) / NB_BLOCK_SIZE;
    if (!bx_by_out) return 0;
    if (cap < (uint64_t) ntx * (uint64_t) nty) return 2;
    std::vector<uint32_t> tab;
    build_tile_order(ntx, nty, nranks, tab);
    for (size_t i = 0; i < tab.size(); ++i) bx_by_out[i] = tab[i];
    return 0;
}

int nb_render_blocks_device(nb_ctx *c, float *blocks_dev, void *stream, nb_stats *st) {
    if (!c || !blocks_dev) return fail("null argument");
    cudaStream_t s = stream ? (cudaStream_t) stream : c->stream;
    if (render_tiles(c, reinterpret_cast<float4 *>(blocks_dev), s, st, nullptr)) return 1;
    if (!st) return 0;               // fire and forget: no counter read-back, no host synchronisation
    return finish_stats(c, s, st, 0);
}

int nb_merge_blocks_device(nb_ctx *c, const float *blocks_dev, int rank, int nranks, float *film_dev, void *stream) {
    if (!c || !blocks_dev || !film_dev) return fail("null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("invalid tile shard (%d of %d)", rank, nranks);
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    if (ensure_device(c)) return 1;
    cudaStream_t s = stream ? (cudaStream_t) stream : c->stream;
    int n = tiles_for(c, rank, nranks, nullptr, nullptr);
    return merge(c, reinterpret_cast<const float4 *>(blocks_dev), n, rank, nranks, reinterpret_cast<float4 *>(film_dev), s);
}

int nb_merge_all_blocks_device(nb_ctx *c, const float *blocks_dev, int nranks, int stride_tiles, float *film_dev, void *stream) {
    if (!c || !blocks_dev || !film_dev) return fail("null argument");
    if (nranks < 1 || stride_tiles < 0) return fail("invalid arguments");
    if (ensure_device(c)) return 1;
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    cudaStream_t s = stream ? (cudaStream_t) stream : c->stream;
    int ntx = 0, nty = 0;
    tiles_for(c, 0, 1, &ntx, &nty);
    const int edge = NB_BLOCK_SIZE + 2 * c->border;
    const long long total = (long long) nranks * stride_tiles * edge * edge;
    if (total == 0) return 0;
    if (ensure_tile_table(c, nranks)) return 1;
    nb::merge_all_blocks_kernel<<<(int) ((total + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4 *>(blocks_dev), nranks, stride_tiles,
                                                                           ntx * nty, c->tile_tab_d, c->W, c->H, c->border, edge, reinterpret_cast<float4 *>(film_dev));
    CK(cudaGetLastError());
    return 0;
}

int nb_render_device(nb_ctx *c, float *film_dev, void *stream, nb_stats *st) {
    if (!c || !film_dev) return fail("null argument");
    cudaStream_t s = stream ? (cudaStream_t) stream : c->stream;
    if (nbm::grouped(c)) {
        if (c->leader) return fail("nb_render_device: call it on the group's leader context");
        if (ensure_device(c)) return 1;
        return nbm::render_group(c, reinterpret_cast<float4 *>(film_dev), s, st);
    }
    int n_tiles = 0;
    if (render_tiles(c, nullptr, s, st, &n_tiles)) return 1;
    const size_t film_elems = (size_t) (c->W + 2 * c->border) * (c->H + 2 * c->border);
    CK(cudaMemsetAsync(film_dev, 0, sizeof(float4) * film_elems, s));
    if (merge(c, c->blocks, n_tiles, c->tile_rank, c->tile_nranks, reinterpret_cast<float4 *>(film_dev), s)) return 1;
    if (!st) return 0;               // fire and forget: no counter read-back, no host synchronisation
    return finish_stats(c, s, st, n_tiles > 0 ? 1 : 0);
}

int nb_last_kernel_ms(nb_ctx *c, double *ms) {
    if (!c || !ms) return fail("null argument");
    if (ensure_device(c)) return 1;
    float f = 0;
    CK(cudaEventElapsedTime(&f, c->ev[1], c->ev[2]));      // fails with cudaErrorNotReady while the render is still running
    *ms = f;
    return 0;
}

int nb_render(nb_ctx *c, float *film_host, nb_stats *st) {
    if (!c || (!film_host && !(nbm::grouped(c) && c->comm_rank != 0))) return fail("null argument");
    if (c->prog_active) return fail("nb_render inside a progressive frame (nb_render_end first)");
    if (ensure_device(c)) return 1;
    const size_t film_elems = (size_t) (c->W + 2 * c->border) * (c->H + 2 * c->border);
    if (film_elems > c->film_cap) {
        if (c->film) cudaFree(c->film);
        c->film = nullptr; c->film_cap = 0;
        CK(cudaMalloc(&c->film, sizeof(float4) * (film_elems ? film_elems : 1)));
        c->film_cap = film_elems;
    }
    cudaStream_t s = c->stream;
    if (nbm::grouped(c)) {
        // N devices behind the same call: tiles sharded tile_id % N, ONE NCCL gather of the finished blocks, ONE merge
        if (c->leader) return fail("nb_render: call it on the group's leader context");
        nb_stats tmp; if (!st) st = &tmp;
        if (nbm::render_group(c, c->film, s, nullptr)) return 1;
        if (c->comm_rank == 0) CK(cudaMemcpyAsync(film_host, c->film, sizeof(float4) * film_elems, cudaMemcpyDeviceToHost, s));
        // statistics after the film copy is queued: one synchronisation for the whole frame
        if (nbm::render_group_stats(c, s, st)) return 1;
        st->d2h_bytes = c->comm_rank == 0 ? sizeof(float4) * film_elems : 0; st->h2d_bytes = sizeof(nb::RenderParams);
        c->film_valid = c->comm_rank == 0;
        return 0;
    }
    int n_tiles = 0;
    if (render_tiles(c, nullptr, s, st, &n_tiles)) return 1;
    CK(cudaMemsetAsync(c->film, 0, sizeof(float4) * film_elems, s));
    if (merge(c, c->blocks, n_tiles, c->tile_rank, c->tile_nranks, c->film, s)) return 1;
    CK(cudaMemcpyAsync(film_host, c->film, sizeof(float4) * film_elems, cudaMemcpyDeviceToHost, s));
    if (finish_stats(c, s, st, n_tiles > 0 ? 1 : 0)) return 1;
    if (st) { st->d2h_bytes = sizeof(float4) * film_elems; st->h2d_bytes = sizeof(nb::RenderParams); }
    c->film_valid = true;
    return 0;
}

int nb_li_samples(nb_ctx *c, uint64_t n, float *lum_host, nb_stats *st) {
    if (!c || (n && !lum_host)) return fail("null argument");
    if (ensure_device(c)) return 1;
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    if (c->integ.type < NB_INT_NORMALS || c->integ.type > NB_INT_SIMPLE) return fail("unsupported integrator type %d (no CPU fallback)", c->integ.type);
    if (c->integ.type == NB_INT_SIMPLE && !c->have_light) return fail("the simple integrator needs nb_set_point_light");
    if (n > 0xffffffffULL) return fail("nb_li_samples: at most 2^32 - 1 paths");
    nb::RenderParams P;
    memset(&P, 0, sizeof P);
    if (fill_scene(c, P.sc)) return 1;
    memcpy(P.light_pos, c->light_pos, sizeof P.light_pos); memcpy(P.light_energy, c->light_energy, sizeof P.light_energy);
    memcpy(P.s2c, c->s2c, sizeof P.s2c); memcpy(P.c2w, c->c2w, sizeof P.c2w);
    P.W = c->W; P.H = c->H; P.invW = 1.0f / (float) c->W; P.invH = 1.0f / (float) c->H;
    P.nearClip = c->nearClip; P.farClip = c->farClip;
    P.seed = c->seed;
    P.integrator = c->integ.type; P.rr_start = c->integ.rr_start > 0 ? c->integ.rr_start : 3;
    P.max_depth = c->integ.max_depth > 0 ? c->integ.max_depth : (1 << 20);
    P.counters = c->counters;
    if (st) memset(st, 0, sizeof *st);
    if (n == 0) return 0;
    cudaStream_t s = c->stream;
    float *dl = nullptr;
    CK(cudaMalloc(&dl, sizeof(float) * n));
    int rc = 0;
    const int grid = (int) std::min<uint64_t>((n + 127) / 128, (uint64_t) c->sm_count * 16);
    cudaError_t e = cudaEventRecord(c->ev[0], s);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->counters, 0, sizeof(unsigned long long) * 8, s);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev[1], s);
    if (e == cudaSuccess) {
        switch (c->integ.type) {
            case 0: nb::li_samples_kernel<0><<<grid, 128, 0, s>>>(P, n, dl); break;
            case 1: nb::li_samples_kernel<1><<<grid, 128, 0, s>>>(P, n, dl); break;
            case 2: nb::li_samples_kernel<2><<<grid, 128, 0, s>>>(P, n, dl); break;
            case 3: nb::li_samples_kernel<3><<<grid, 128, 0, s>>>(P, n, dl); break;
            case 4: nb::li_samples_kernel<4><<<grid, 128, 0, s>>>(P, n, dl); break;
            case 5: nb::li_samples_kernel<5><<<grid, 128, 0, s>>>(P, n, dl); break;
            default: nb::li_samples_kernel<6><<<grid, 128, 0, s>>>(P, n, dl); break;
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(c->ev[2], s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lum_host, dl, sizeof(float) * n, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) rc = fail("nb_li_samples failed: %s", cudaGetErrorString(e));
    if (!rc) rc = finish_stats(c, s, st, 0);
    else cudaStreamSynchronize(s);
    cudaFree(dl);
    if (!rc && st) { st->samples = n; st->launches = 1; st->d2h_bytes = sizeof(float) * n; st->h2d_bytes = sizeof(nb::RenderParams); }
    return rc;
}

// defined in nb_aux.cu (own translation unit: see the note there)
cudaError_t nb_aux_launch_bsdf_query(const void *devmesh, size_t devmesh_bytes, unsigned long long n, const float *wi, int wi_stride,
                                     const float *a, int mode, float *out, int grid, cudaStream_t s);

namespace {
int bsdf_query(nb_ctx *c, const nb_bsdf_desc *b, const float *wi, int wi_per_query, const float *a, int a_width, uint64_t n,
               float *out, int out_width, int mode) {
    if (!c || !b || (n && (!wi || !a || !out))) return fail("null argument");
    if (b->type < NB_BSDF_DIFFUSE || b->type > NB_BSDF_MICROFACET) return fail("unsupported BSDF type %d (no CPU fallback)", b->type);
    if (ensure_device(c)) return 1;
    if (n == 0) return 0;
    nb::DevMesh m;
    memset(&m, 0, sizeof m);
    m.bsdf_type = b->type; memcpy(m.albedo, b->albedo, sizeof m.albedo);
    m.alpha = b->alpha; m.intIOR = b->intIOR; m.extIOR = b->extIOR; m.ks = b->ks;
    const size_t wi_floats = wi_per_query ? 3 * n : 3;
    float *dwi = nullptr, *da = nullptr, *dout = nullptr;
    cudaError_t e = cudaMalloc(&dwi, sizeof(float) * wi_floats);
    if (e == cudaSuccess) e = cudaMalloc(&da, sizeof(float) * a_width * n);
    if (e == cudaSuccess) e = cudaMalloc(&dout, sizeof(float) * out_width * n);
    cudaStream_t s = c->stream;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dwi, wi, sizeof(float) * wi_floats, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(da, a, sizeof(float) * a_width * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        const int grid = (int) std::min<uint64_t>((n + 127) / 128, (uint64_t) c->sm_count * 16);
        e = nb_aux_launch_bsdf_query(&m, sizeof m, n, dwi, wi_per_query ? 3 : 0, da, mode, dout, grid, s);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, sizeof(float) * out_width * n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(dwi); cudaFree(da); cudaFree(dout);
    if (e != cudaSuccess) return fail("BSDF query failed: %s", cudaGetErrorString(e));
    return 0;
}
}  // namespace

int nb_bsdf_sample(nb_ctx *c, const nb_bsdf_desc *b, const float *wi, int wi_per_query, const float *xi, uint64_t n, float *out8) {
    return bsdf_query(c, b, wi, wi_per_query, xi, 2, n, out8, 8, 0);
}

int nb_bsdf_eval_pdf(nb_ctx *c, const nb_bsdf_desc *b, const float *wi, int wi_per_query, const float *wo, uint64_t n, float *out4) {
    return bsdf_query(c, b, wi, wi_per_query, wo, 3, n, out4, 4, 1);
}

int nb_intersect_device(nb_ctx *c, const nb_ray *rays_dev, uint64_t n, nb_hit *hits_dev, int shadow, void *stream, nb_stats *st) {
    if (!c) return fail("null context");
    if (ensure_device(c)) return 1;
    nb::SceneDev sc;
    if (fill_scene(c, sc)) return 1;
    cudaStream_t s = stream ? (cudaStream_t) stream : c->stream;
    CK(cudaEventRecord(c->ev[0], s));
    CK(cudaMemsetAsync(c->counters, 0, sizeof(unsigned long long) * 8, s));
    CK(cudaEventRecord(c->ev[1], s));
    if (n) {
        int grid = (int) std::min<uint64_t>((n + 127) / 128, (uint64_t) c->sm_count * 16);
        if (c->opt_count) nb::intersect_kernel<true><<<grid, 128, 0, s>>>(sc, reinterpret_cast<const float4 *>(rays_dev), n, reinterpret_cast<nb::HitOut *>(hits_dev), shadow, nullptr, c->counters);
        else nb::intersect_kernel<false><<<grid, 128, 0, s>>>(sc, reinterpret_cast<const float4 *>(rays_dev), n, reinterpret_cast<nb::HitOut *>(hits_dev), shadow, nullptr, c->counters);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(c->ev[2], s));
    if (st) { memset(st, 0, sizeof *st); st->launches = n ? 1 : 0; }
    return finish_stats(c, s, st, 0);
}

int nb_intersect(nb_ctx *c, const nb_ray *rays, uint64_t n, nb_hit *hits, int shadow, nb_stats *st) {
    if (!c || (n && (!rays || !hits))) return fail("null argument");
    if (ensure_device(c)) return 1;
    if (n == 0) { if (st) memset(st, 0, sizeof *st); return 0; }
    nb_ray *dr = nullptr; nb_hit *dh = nullptr;
    CK(cudaMalloc(&dr, sizeof(nb_ray) * n));
    if (cudaMalloc(&dh, sizeof(nb_hit) * n) != cudaSuccess) { cudaFree(dr); return fail("cudaMalloc failed"); }
    int rc = 0;
    if (cudaMemcpyAsync(dr, rays, sizeof(nb_ray) * n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = fail("H2D copy failed");
    if (!rc) rc = nb_intersect_device(c, dr, n, dh, shadow, c->stream, st);
    if (!rc && cudaMemcpy(hits, dh, sizeof(nb_hit) * n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail("D2H copy failed");
    cudaFree(dr); cudaFree(dh);
    if (!rc && st) { st->h2d_bytes = sizeof(nb_ray) * n; st->d2h_bytes = sizeof(nb_hit) * n; }
    return rc;
}

int nb_intersect_full(nb_ctx *c, const nb_ray *rays, uint64_t n, float *out16) {
    if (!c || (n && (!rays || !out16))) return fail("null argument");
    if (ensure_device(c)) return 1;
    if (n == 0) return 0;
    nb::SceneDev sc;
    if (fill_scene(c, sc)) return 1;
    nb_ray *dr = nullptr; float *df = nullptr;
    CK(cudaMalloc(&dr, sizeof(nb_ray) * n));
    if (cudaMalloc(&df, sizeof(float) * 16 * n) != cudaSuccess) { cudaFree(dr); return fail("cudaMalloc failed"); }
    int rc = 0;
    // copies and kernel on the SAME (non-blocking) stream: a pageable cudaMemcpy on the legacy stream does not order with it
    if (cudaMemcpyAsync(dr, rays, sizeof(nb_ray) * n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = fail("H2D copy failed");
    if (!rc) {
        int grid = (int) std::min<uint64_t>((n + 127) / 128, (uint64_t) c->sm_count * 16);
        nb::intersect_kernel<false><<<grid, 128, 0, c->stream>>>(sc, reinterpret_cast<const float4 *>(dr), n, nullptr, 0, df, nullptr);
        if (cudaGetLastError() != cudaSuccess) rc = fail("intersect kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (!rc && cudaMemcpyAsync(out16, df, sizeof(float) * 16 * n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = fail("D2H copy failed");
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) rc = fail("intersect kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(dr); cudaFree(df);
    return rc;
}

int nb_film_to_rgb(nb_ctx *c, const float *film_host, float *rgb_host) {
    if (!c || !film_host || !rgb_host) return fail("null argument");
    if (ensure_device(c)) return 1;
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    const size_t film_elems = (size_t) (c->W + 2 * c->border) * (c->H + 2 * c->border);
    float4 *df = nullptr; float *dr = nullptr;
    CK(cudaMalloc(&df, sizeof(float4) * film_elems));
    if (cudaMalloc(&dr, sizeof(float) * 3 * c->W * c->H) != cudaSuccess) { cudaFree(df); return fail("cudaMalloc failed"); }
    int rc = 0;
    if (cudaMemcpyAsync(df, film_host, sizeof(float4) * film_elems, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = fail("H2D copy failed");
    if (!rc) {
        nb::film_to_rgb_kernel<<<(c->W * c->H + 255) / 256, 256, 0, c->stream>>>(df, c->W, c->H, c->border, dr);
        if (cudaGetLastError() != cudaSuccess) rc = fail("film_to_rgb kernel failed");
    }
    if (!rc && cudaMemcpyAsync(rgb_host, dr, sizeof(float) * 3 * c->W * c->H, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = fail("D2H copy failed");
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) rc = fail("film_to_rgb kernel failed");
    cudaFree(df); cudaFree(dr);
    return rc;
}

// ---- progressive frames (stands where NoriScreen refreshes from the shared ImageBlock while the tiles render, ref: src/gui.cpp:120-138)
int nb_render_begin(nb_ctx *c) {
    if (!c) return fail("null context");
    if (nbm::grouped(c)) return fail("progressive frames run on a single-device context");
    if (ensure_device(c)) return 1;
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    if (c->seed_mode != NB_SEED_PER_SAMPLE) return fail("progressive frames need per-(pixel, sample) sampler streams (NB_SEED_PER_SAMPLE)");
    const int n_tiles = tiles_for(c, c->tile_rank, c->tile_nranks, nullptr, nullptr);
    const int edge = NB_BLOCK_SIZE + 2 * c->border;
    const size_t blk_elems = (size_t) n_tiles * edge * edge;
    if (blk_elems > c->blocks_cap) {
        if (c->blocks) cudaFree(c->blocks);
        c->blocks = nullptr; c->blocks_cap = 0;
        CK(cudaMalloc(&c->blocks, sizeof(float4) * (blk_elems ? blk_elems : 1)));
        c->blocks_cap = blk_elems;
    }
    if (blk_elems) CK(cudaMemsetAsync(c->blocks, 0, sizeof(float4) * blk_elems, c->stream));
    c->prog_active = true; c->prog_done = 0; c->prog_pass = 0; c->film_valid = false;
    memset(&c->prog_stats, 0, sizeof c->prog_stats);
    return 0;
}

int nb_render_pass(nb_ctx *c, uint32_t n_samples, nb_stats *st) {
    if (!c) return fail("null context");
    if (!c->prog_active) return fail("nb_render_pass: call nb_render_begin first");
    if (ensure_device(c)) return 1;
    if (n_samples == 0 || c->prog_done + n_samples > c->spp) return fail("nb_render_pass: %u samples asked, %u of %u left", n_samples, c->spp - c->prog_done, c->spp);
    c->prog_pass = n_samples;
    nb_stats ps; memset(&ps, 0, sizeof ps);
    int n_tiles = 0;
    if (render_tiles(c, nullptr, c->stream, &ps, &n_tiles)) { c->prog_active = false; return 1; }
    if (finish_stats(c, c->stream, &ps, 0)) { c->prog_active = false; return 1; }
    c->prog_done += n_samples;
    nb_stats &a = c->prog_stats;
    a.samples += ps.samples; a.rays += ps.rays; a.node_visits += ps.node_visits; a.tri_tests += ps.tri_tests; a.hits_shaded += ps.hits_shaded;
    a.kernel_ms += ps.kernel_ms; a.total_ms += ps.total_ms; a.launches += ps.launches;
    if (st) *st = a;
    return 0;
}

int nb_render_preview(nb_ctx *c, float *film_host, uint8_t *rgb8_host) {
    if (!c) return fail("null context");
    if (!c->prog_active) return fail("nb_render_preview: call nb_render_begin first");
    if (ensure_device(c)) return 1;
    const size_t film_elems = (size_t) (c->W + 2 * c->border) * (c->H + 2 * c->border);
    if (film_elems > c->film_cap) {
        if (c->film) cudaFree(c->film);
        c->film = nullptr; c->film_cap = 0;
        CK(cudaMalloc(&c->film, sizeof(float4) * (film_elems ? film_elems : 1)));
        c->film_cap = film_elems;
    }
    cudaStream_t s = c->stream;
    const int n_tiles = tiles_for(c, c->tile_rank, c->tile_nranks, nullptr, nullptr);
    CK(cudaMemsetAsync(c->film, 0, sizeof(float4) * film_elems, s));
    if (merge(c, c->blocks, n_tiles, c->tile_rank, c->tile_nranks, c->film, s)) return 1;
    if (film_host) CK(cudaMemcpyAsync(film_host, c->film, sizeof(float4) * film_elems, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    c->film_valid = true;
    if (rgb8_host) return nb_last_film_to_srgb8(c, rgb8_host);
    return 0;
}

int nb_render_end(nb_ctx *c) {
    if (!c) return fail("null context");
    c->prog_active = false; c->prog_done = 0; c->prog_pass = 0;
    return 0;
}

int nb_last_film_to_srgb8(nb_ctx *c, uint8_t *rgb8_host) {
    if (!c || !rgb8_host) return fail("null argument");
    if (c->leader) return fail("nb_last_film_to_srgb8: call it on the group's leader context");
    if (ensure_device(c)) return 1;
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    const size_t film_elems = (size_t) (c->W + 2 * c->border) * (c->H + 2 * c->border);
    if (!c->film || !c->film_valid || c->film_cap < film_elems) return fail("nb_last_film_to_srgb8: no film on the device (call nb_render first)");
    const size_t n = (size_t) c->W * c->H * 3;
    unsigned char *d8 = nullptr;
    CK(cudaMalloc(&d8, n));
    nb::film_to_srgb8_kernel<<<(c->W * c->H + 255) / 256, 256, 0, c->stream>>>(c->film, c->W, c->H, c->border, d8);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(rgb8_host, d8, n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d8);
    if (e != cudaSuccess) return fail("film_to_srgb8 kernel failed: %s", cudaGetErrorString(e));
    return 0;
}

int nb_set_option(nb_ctx *c, const char *key, int64_t value) {
    if (!c || !key) return fail("null argument");
    std::string k(key);
    if (k == "blocks_per_sm") c->opt_blocks_per_sm = value;
    else if (k == "smem_nodes") {
#if NB_WIDE
        if (value != 0) return fail("smem_nodes stages 64-byte binary nodes; this build walks the 8-wide hierarchy");
#endif
        c->opt_smem_nodes = value;
    }
    else if (k == "chunk") c->opt_chunk = value;
    else if (k == "count") c->opt_count = value;
    else if (k == "coarse") { if (value != 0 && value != 2 && value != 4 && value != 8 && value != 16) return fail("coarse must be 2, 4, 8 or 16 (samples per coarse work unit; 0 = default 8)"); c->opt_coarse = value ? value : 8; }
    else if (k == "guided") { if (value < -1 || value > 100) return fail("guided must be in [0, 100] (percent of the samples in coarse work units; -1 = default 75)"); c->opt_guided = value < 0 ? 75 : value; }
    else if (k == "prefetch") c->opt_prefetch = value;
    else if (k == "engine") { if (value != 0 && value != 2) return fail("engine must be 0 (fused kernel) or 2 (wavefront)"); c->opt_engine = value; }
    else if (k == "wf_pool") { if (value < 128 || value > (1ll << 28)) return fail("wf_pool must be in [128, 2^28]"); c->opt_wf_pool = value; }
    else if (k == "wf_check") { if (value < 1 || value > 1024) return fail("wf_check must be in [1, 1024]"); c->opt_wf_check = value; }
    else if (k == "occ_tail") { if (value < 0 || value > 31) return fail("occ_tail must be in [0, 31]"); c->opt_occ_tail = value; }
    else if (k == "max_leaf") {
#if NB_WIDE
        if (value < 1 || value > 3) return fail("max_leaf must be in [1, 3] (a leaf child of a wide node holds 1..3 triangles)");
#endif
        c->opt_max_leaf = value; c->built = false;
    }
    else if (k == "bfs_nodes") { c->opt_bfs_nodes = value; c->built = false; }
    else if (k == "sah_bins") { if (value < 4 || value > 32) return fail("sah_bins must be in [4, 32]"); c->opt_sah_bins = value; c->built = false; }
    else if (k == "builder") { if (value != 0 && value != 1) return fail("builder must be 0 (host SAH) or 1 (device LBVH)"); c->opt_builder = value; c->built = false; }
    else return fail("unknown option \"%s\"", key);
    for (nb_ctx *f : c->followers) if (nb_set_option(f, key, value)) return 1;
    return 0;
}

int nb_debug_counters(nb_ctx *c, uint64_t out[8]) {
    if (!c || !out) return fail("null argument");
    for (int i = 0; i < 8; ++i) out[i] = c->counters_h[i];
    return 0;
}

int nb_scene_info(nb_ctx *c, uint64_t *ntris, uint64_t *nnodes, uint64_t *scene_bytes, int *bvh_depth) {
    if (!c) return fail("null context");
    if (!c->built) return fail("nb_build_accel has not been called");
    if (ntris) *ntris = c->n_prims;
    if (nnodes) *nnodes = c->n_nodes;
    if (scene_bytes) *scene_bytes = c->nodes.bytes() + c->tris.bytes() + c->verts.bytes() + c->normals.bytes() + c->uvs.bytes() + c->faces.bytes();
    if (bvh_depth) *bvh_depth = c->bvh_depth;
    return 0;
}

}  // extern "C"

#include "nb_multi.inl"
