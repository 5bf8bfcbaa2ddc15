// nb_wave.cuh -- the wavefront engine (nb_set_option(ctx, "engine", 2)): the render hot path as a stream of rays.
//
// The fused kernel (nb_kernels.cuh: render_kernel) keeps one light path per lane and walks its rays in lock step; on the
// path tracers that costs 10 / 32 active lanes, 650 B of register spills per thread around the walk and an instruction
// working set of 78 KB (profiles/r1_final_ncu_render_kernel_cbox-mis_phased.txt).  Here the same per-path
// arithmetic -- shade<INTEG>() and begin_path() of nb_kernels.cuh, unchanged, so the sampler draw order and every rounding
// are the oracle's -- is split the way the north star asks for ("persistent-threads wavefront tracer"):
//
//   path pool     P slots, structure of arrays in HBM (kWfCols float4 columns: ray, hit, sampler state, radiance, throughput)
//   wf_logic      one thread per slot, every lane busy: finished paths are splatted (ImageBlock::put, ref: src/block.cpp:62-91)
//                 and replaced by the next camera sample (ref: src/main.cpp:41-46) in place; paths with a traced ray take one
//                 integrator step (emission, Russian roulette, next-event estimation, BSDF sample).  Rays leave through two
//                 queues: extension rays as slot indices, occlusion rays with the radiance they would add and their slot.
//   wf_trace      persistent warps with DYNAMIC FETCH: a lane keeps its walk (node, stack) across refills, the warp leaves the
//                 walk when at most tail_lanes lanes are still busy, finished lanes write their result and take the next ray
//                 of the queue (one atomic per refill, ballot/popc ranking) -- the dense any-hit / closest-hit kernels the
//                 lock-step model asked for.  Occlusion rays first (their result is a 12-byte add to
//                 the slot's radiance), then extension rays (result: t, u, v, triangle).
//
// One iteration = one wf_logic + one wf_trace launch; the host loop (nb_wave.cu) runs iterations until no slot holds a path.
// Rays are materialised here -- SURVEY 8d's 48 B per ray apply.
#pragma once
#include "nb_kernels.cuh"

namespace nb {

constexpr int kWfCols = 7;
// columns of the pool (each wf_pool float4):
//   0: ray origin, mint            1: ray direction, maxt (closest hit: t after the trace)
//   2: hit u, v, triangle (bits), -
//   3: pcg32 state (2 words), film position sx, sy
//   4: sample index within the pixel (its pcg32 stream: inc = 2 * sample + 1), packed depth | stage | prev_specular, prev_pdf, owned-tile slot
//   5: radiance L.xyz, -           6: throughput T.xyz, -
enum { WF_NEXT_LO = 0, WF_NEXT_HI = 1, WF_EXT_COUNT = 2, WF_SHADOW_COUNT = 3, WF_EXT_FETCH = 4, WF_SHADOW_FETCH = 5, WF_LIVE = 6, WF_WATCHDOG = 7, WF_NCTR = 8 };
enum { WST_EMPTY = 0, WST_TRACED = 1, WST_WAIT = 2 };      // slot state: no path / extension ray traced or in the queue / waits for its last occlusion ray

__device__ __forceinline__ unsigned wf_pack(int depth, int stage, bool prev_specular) {
    return ((unsigned) depth << 16) | ((unsigned) stage << 8) | (prev_specular ? 1u : 0u);
}

// Sample index -> (owned tile, pixel, sample): 32x32 tiles, inside a tile 8x4 pixel patches x chunks of wf_chunk samples, the
// 32 pixels of a patch fastest -- consecutive indices are the coherent camera rays the fused kernel gives one warp.
// Returns false for the padding of ragged tiles / of the last sample chunk.
__device__ __forceinline__ bool wf_decode(const RenderParams &P, unsigned long long idx, int &tile_slot, int &px, int &py, uint32_t &sample) {
    const uint32_t per_unit = 32u * P.wf_chunk;
    const unsigned long long unit = idx / per_unit;
    const uint32_t item = (uint32_t) (idx % per_unit);
    const uint32_t patch = (uint32_t) (unit % 32ull);
    const unsigned long long rest = unit / 32ull;
    const uint32_t chunk_id = (uint32_t) (rest % P.nchunks);
    tile_slot = (int) (rest / P.nchunks);
    const int tile_id = P.tile_rank + tile_slot * P.tile_nranks;
    int tbx, tby; tile_xy(P.tile_tab, tile_id, tbx, tby);
    const int tox = tbx * 32, toy = tby * 32;
    const uint32_t pl = item & 31u;
    px = tox + (int) (patch & 3u) * 8 + (int) (pl & 7u);
    py = toy + (int) (patch >> 2) * 4 + (int) (pl >> 3);
    sample = chunk_id * P.wf_chunk + (item >> 5);
    return px < P.W && py < P.H && sample < P.spp;
}

// ------------------------------------------------------------------ logic: one thread per pool slot
template <int INTEG>
__global__ void __launch_bounds__(128, 4) wf_logic_kernel(const __grid_constant__ RenderParams P) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;          // grid covers the pool exactly (wf_pool % 128 == 0)
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const size_t N = P.wf_pool;
    float4 *C = P.wf_cols;
    unsigned n_rays = 0, n_hits = 0;

    float4 c4 = C[4 * N + slot];
    const unsigned packed = __float_as_uint(c4.y);
    int stage = (int) ((packed >> 8) & 0xffu);
    Path ps; Ray ray; Trav tr;
    ps.deferred = 0;
    ps.tile_slot = __float_as_int(c4.w);
    bool emit_ext = false;           // this slot leaves the pass with an extension ray to trace
    bool done = false;               // path complete: splat, then take the next sample

    if (stage == WST_TRACED) {
        const float4 c0 = C[0 * N + slot], c1 = C[1 * N + slot], c2 = C[2 * N + slot], c3 = C[3 * N + slot], c5 = C[5 * N + slot], c6 = C[6 * N + slot];
        ray.ox = c0.x; ray.oy = c0.y; ray.oz = c0.z; ray.mint = c0.w; ray.dx = c1.x; ray.dy = c1.y; ray.dz = c1.z; ray.maxt = c1.w;
        tr.hu = c2.x; tr.hv = c2.y; tr.hprim = __float_as_uint(c2.z); tr.node = kDone; tr.sp = 0;
        ps.rng.state = ((uint64_t) __float_as_uint(c3.y) << 32) | (uint64_t) __float_as_uint(c3.x);
        ps.rng.inc = ((uint64_t) __float_as_uint(c4.x) << 1) | 1ull;
        ps.sx = c3.z; ps.sy = c3.w;
        ps.depth = (int) (packed >> 16); ps.prev_specular = (packed & 1u) != 0u; ps.prev_pdf = c4.z;
        ps.L = mk(c5.x, c5.y, c5.z); ps.T = mk(c6.x, c6.y, c6.z);
        ps.has_next = false; ps.stage = ST_EXTEND;
        const bool finished = shade<INTEG>(P, ps, ray, tr, n_hits);
        if (!finished) {
            emit_ext = true;
        } else if (ps.deferred != 0) {
            stage = WST_WAIT;        // the path's last occlusion ray is in the queue: L is complete after this iteration's trace
            C[3 * N + slot] = make_float4(c3.x, c3.y, ps.sx, ps.sy);
            C[4 * N + slot] = make_float4(c4.x, __uint_as_float(wf_pack(ps.depth, WST_WAIT, false)), 0.f, c4.w);
            C[5 * N + slot] = make_float4(ps.L.x, ps.L.y, ps.L.z, 0.f);
        } else {
            done = true;
        }
    } else if (stage == WST_WAIT) {
        const float4 c3 = C[3 * N + slot], c5 = C[5 * N + slot];
        ps.sx = c3.z; ps.sy = c3.w; ps.L = mk(c5.x, c5.y, c5.z);
        done = true;
    }
    if (done) {
        const int tile_id = P.tile_rank + ps.tile_slot * P.tile_nranks;
        int tbx, tby; tile_xy(P.tile_tab, tile_id, tbx, tby);
    const int tox = tbx * 32, toy = tby * 32;
        splat(P, ps.tile_slot, tox, toy, min(32, P.W - tox), min(32, P.H - toy), ps.sx, ps.sy, ps.L);
        stage = WST_EMPTY;
    }
    // ---- regeneration (ref: src/main.cpp:41-46): free slots take the next sample indices, one atomic per round and warp
    uint32_t sample = 0;
    bool need = (stage == WST_EMPTY);
    bool was_empty = need && !done;      // an empty slot that stays empty needs no store
    unsigned need_mask = __ballot_sync(0xffffffffu, need);
    while (need_mask != 0u) {
        unsigned lo = 0, hi = 0;
        if (lane == 0) {
            const unsigned long long b = atomicAdd(reinterpret_cast<unsigned long long *>(&P.wf_ctr[WF_NEXT_LO]), (unsigned long long) __popc(need_mask));
            lo = (unsigned) b; hi = (unsigned) (b >> 32);
        }
        lo = __shfl_sync(0xffffffffu, lo, 0); hi = __shfl_sync(0xffffffffu, hi, 0);
        const unsigned long long base = ((unsigned long long) hi << 32) | lo;
        if (need) {
            const unsigned long long idx = base + (unsigned long long) __popc(need_mask & lt_mask);
            if (idx >= P.wf_total) {
                need = false;                                  // samples exhausted: the slot stays empty
            } else {
                int tile_slot, px, py;
                if (wf_decode(P, idx, tile_slot, px, py, sample)) {
                    ps.tile_slot = tile_slot;
                    begin_path(P, ps, ray, px, py, sample);
                    ps.deferred = 0;
                    stage = WST_TRACED; emit_ext = true; need = false; was_empty = false;
                }
            }
        }
        need_mask = __ballot_sync(0xffffffffu, need);
    }
    if (stage == WST_EMPTY && !was_empty) C[4 * N + slot] = make_float4(0.f, __uint_as_float(wf_pack(0, WST_EMPTY, false)), 0.f, 0.f);
    // ---- extension rays: state back to the pool, slot index into the queue (warp-aggregated append)
    const unsigned ext_mask = __ballot_sync(0xffffffffu, emit_ext);
    if (ext_mask != 0u) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&P.wf_ctr[WF_EXT_COUNT], (unsigned) __popc(ext_mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (emit_ext) {
            n_rays++;
            P.wf_ext[base + (unsigned) __popc(ext_mask & lt_mask)] = slot;
            const uint32_t smp = (uint32_t) (ps.rng.inc >> 1);
            C[0 * N + slot] = make_float4(ray.ox, ray.oy, ray.oz, ray.mint);
            C[1 * N + slot] = make_float4(ray.dx, ray.dy, ray.dz, ray.maxt);
            C[3 * N + slot] = make_float4(__uint_as_float((unsigned) ps.rng.state), __uint_as_float((unsigned) (ps.rng.state >> 32)), ps.sx, ps.sy);
            C[4 * N + slot] = make_float4(__uint_as_float(smp), __uint_as_float(wf_pack(ps.depth, WST_TRACED, ps.prev_specular)), ps.prev_pdf, __int_as_float(ps.tile_slot));
            C[5 * N + slot] = make_float4(ps.L.x, ps.L.y, ps.L.z, 0.f);
            C[6 * N + slot] = make_float4(ps.T.x, ps.T.y, ps.T.z, 0.f);
        }
    }
    // ---- bookkeeping: live slots (loop condition of the host), ray / hit counters
    const unsigned live_mask = __ballot_sync(0xffffffffu, stage != WST_EMPTY);
    n_rays += ps.deferred;
    unsigned v1 = n_rays, v4 = n_hits;
    for (int o = 16; o > 0; o >>= 1) { v1 += __shfl_down_sync(0xffffffffu, v1, o); v4 += __shfl_down_sync(0xffffffffu, v4, o); }
    if (lane == 0) {
        if (live_mask) atomicAdd(&P.wf_ctr[WF_LIVE], (unsigned) __popc(live_mask));
        if (v1) atomicAdd(&P.counters[1], (unsigned long long) v1);
        if (v4) atomicAdd(&P.counters[4], (unsigned long long) v4);
    }
}

// ------------------------------------------------------------------ trace: persistent warps, dynamic fetch
// phase 0: the occlusion queue (any hit; an unoccluded ray adds its radiance to the slot's L -- one ray per slot and
// iteration, so the read-modify-write needs no atomic); phase 1: the extension queue (closest hit -> t, u, v, triangle).
template <bool COUNT>
__global__ void __launch_bounds__(128, NB_MIN_BLOCKS) wf_trace_kernel(const __grid_constant__ RenderParams P) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const size_t N = P.wf_pool;
    float4 *C = P.wf_cols;
    StackT stack[kStackN + 1];
    unsigned n_nodes = 0, n_tris = 0;
    const long long wd_t0 = clock64();
    for (int phase = 0; phase < 2; ++phase) {
        const bool any_hit = (phase == 0);
        const unsigned n_in = any_hit ? P.wf_ctr[WF_SHADOW_COUNT] : P.wf_ctr[WF_EXT_COUNT];
        unsigned *fetch = &P.wf_ctr[any_hit ? WF_SHADOW_FETCH : WF_EXT_FETCH];
        Ray ray; Trav tr;
        tr.node = kDone; tr.sp = 0; tr.hprim = 0xffffffffu; tr.hu = 0.f; tr.hv = 0.f;
        V3 contrib = mk(0, 0, 0); uint32_t slot = 0;
        bool active = false, exhausted = (n_in == 0u);
        for (;;) {
            if (__any_sync(0xffffffffu, clock64() - wd_t0 > NB_WATCHDOG_CYCLES)) { if (lane == 0) atomicOr(&P.wf_ctr[WF_WATCHDOG], 1u); return; }
            // ---- refill (all 32 lanes are converged here)
            const unsigned idle_mask = __ballot_sync(0xffffffffu, !active);
            if (idle_mask != 0u && !exhausted) {
                const unsigned n_idle = __popc(idle_mask);
                unsigned start = 0;
                if (lane == 0) start = atomicAdd(fetch, n_idle);
                start = __shfl_sync(0xffffffffu, start, 0);
                const unsigned idx = start + (unsigned) __popc(idle_mask & lt_mask);
                if (!active && idx < n_in) {
                    if (any_hit) {
                        const float4 *q = P.occ_queue + (size_t) idx * 3u;
                        const float4 a = q[0], b = q[1], c = q[2];
                        ray.ox = a.x; ray.oy = a.y; ray.oz = a.z; ray.mint = a.w; ray.dx = b.x; ray.dy = b.y; ray.dz = b.z; ray.maxt = b.w;
                        contrib = mk(c.x, c.y, c.z); slot = __float_as_uint(c.w);
                    } else {
                        slot = P.wf_ext[idx];
                        const float4 a = C[0 * N + slot], b = C[1 * N + slot];
                        ray.ox = a.x; ray.oy = a.y; ray.oz = a.z; ray.mint = a.w; ray.dx = b.x; ray.dy = b.y; ray.dz = b.z; ray.maxt = b.w;
                    }
                    tr.node = kDone;                              // fresh ray for walk_wave
                    active = true;
                }
                if (start + n_idle >= n_in) exhausted = true;
            }
            if (__ballot_sync(0xffffffffu, active) == 0u) break;
            // ---- walk until at most tail_lanes lanes are left (to completion once the queue is drained)
            if (active) {
                unsigned nn = 0, nt = 0;
                walk_wave<COUNT, false>(P.sc.nodes, P.sc.tris, nullptr, 0, ray, tr, stack, any_hit, exhausted ? 0 : P.tail_lanes, nn, nt);
                if (COUNT) { n_nodes += nn; n_tris += nt; }
                if (tr.node == kDone) {
                    if (any_hit) {
                        if (tr.hprim == 0xffffffffu) {
                            float4 l = C[5 * N + slot];
                            l.x += contrib.x; l.y += contrib.y; l.z += contrib.z;
                            C[5 * N + slot] = l;
                        }
                    } else {
                        C[2 * N + slot] = make_float4(tr.hu, tr.hv, __uint_as_float(tr.hprim), 0.f);
                        reinterpret_cast<float *>(&C[1 * N + slot])[3] = ray.maxt;      // t of the closest hit (unchanged on a miss)
                    }
                    active = false;
                }
            }
        }
    }
    if (COUNT) {
        unsigned long long v2 = n_nodes, v3 = n_tris;
        for (int o = 16; o > 0; o >>= 1) { v2 += __shfl_down_sync(0xffffffffu, v2, o); v3 += __shfl_down_sync(0xffffffffu, v3, o); }
        if (lane == 0) { atomicAdd(&P.counters[2], v2); atomicAdd(&P.counters[3], v3); }
    }
}

}  // namespace nb
