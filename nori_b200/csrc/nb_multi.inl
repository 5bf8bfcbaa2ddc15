// nb_multi.inl -- multi-GPU behind the C-ABI (included at the end of nb_api.cu: it needs nb_ctx and the launchers).
//
// Replaces the reference's tile loop + merge (ref: src/main.cpp:85-113, src/block.cpp:93-102) for N devices: 32x32 image
// tiles are sharded tile_id % N (BlockGenerator's role, ref: src/block.cpp:119-152); every device renders its tiles into
// packed ImageBlocks; ONE grouped ncclSend/ncclRecv per frame gathers the finished blocks on rank 0 over NVLink (NCCL has
// no native gather); ONE merge launch adds them into the film.  The scene (one arena, nb_ctx.h) is built once (rank 0) and
// replicated with ONE ncclBroadcast over NVLink; re-uploads are sharded (1/N per PCIe link + ONE in-place ncclAllGather).
//
// Two ways to form the group, same code underneath:
//   nb_create_multi(devices, n)            one process drives n devices (ncclCommInitAll) -- what `nori --gpus n` uses
//   nb_comm_init_rank(ctx, id, rank, n)    one process per device (torchrun); the 128-byte id travels by any transport
//
// NCCL is dlopen'ed (libnccl.so.2): single-GPU users need no NCCL at all, and a process that already holds a copy
// (torch's bundled one) shares it instead of loading a second.
#include <dlfcn.h>
#include <nccl.h>
#include <thread>

namespace nbm {

struct NcclApi {
    void *handle = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &nccl() {
    static NcclApi a = [] {
        NcclApi x;
        const char *names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char *n : names) { x.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (x.handle) break; }
        if (!x.handle) { x.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return x; }
#define NB_SYM(field, name) do { *(void **) (&x.field) = dlsym(x.handle, name); if (!x.field) { x.err = std::string("libnccl lacks ") + name; return x; } } while (0)
        NB_SYM(GetUniqueId, "ncclGetUniqueId"); NB_SYM(CommInitRank, "ncclCommInitRank"); NB_SYM(CommInitAll, "ncclCommInitAll");
        NB_SYM(CommDestroy, "ncclCommDestroy"); NB_SYM(GroupStart, "ncclGroupStart"); NB_SYM(GroupEnd, "ncclGroupEnd");
        NB_SYM(Send, "ncclSend"); NB_SYM(Recv, "ncclRecv"); NB_SYM(Broadcast, "ncclBroadcast"); NB_SYM(AllGather, "ncclAllGather"); NB_SYM(GetErrorString, "ncclGetErrorString");
#undef NB_SYM
        return x;
    }();
    return a;
}

#define NCK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail("%s failed: %s", #call, nbm::nccl().GetErrorString(r_)); } while (0)

int render_group_stats(nb_ctx *c, cudaStream_t s0, nb_stats *st);

int need_nccl() {
    NcclApi &a = nccl();
    if (!a.err.empty()) return fail("multi-GPU rendering needs NCCL: %s", a.err.c_str());
    return 0;
}

// the contexts of the group that live in this process: leader + followers, or the single context of a process-per-GPU rank
std::vector<nb_ctx *> locals(nb_ctx *c) {
    std::vector<nb_ctx *> L; L.push_back(c);
    for (nb_ctx *f : c->followers) L.push_back(f);
    return L;
}

bool grouped(const nb_ctx *c) { return c->comm != nullptr && c->comm_nranks > 1; }

// Replicates the built scene of rank 0 on every rank (with_header): a 16-word header (table sizes + hierarchy facts), then
// the scene arena in ONE ncclBroadcast over NVLink.  Ranks other than 0 hold device-only tables; a process-per-GPU rank also
// keeps a pinned copy of ITS shard of the arena for later re-uploads.
// Re-upload (!with_header): SHARDED -- every rank copies 1/N of the arena host->device over its own PCIe link, one in-place
// ncclAllGather over NVLink completes the arena on every device.
int replicate_scene(nb_ctx *c, bool with_header) {
    if (need_nccl()) return 1;
    NcclApi &N = nccl();
    std::vector<nb_ctx *> L = locals(c);
    const int nranks = c->comm_nranks;
    constexpr int kHdr = 16;
    if (with_header) {
        std::vector<unsigned long long *> hd(L.size(), nullptr);
        auto cleanup = [&]() { for (size_t i = 0; i < L.size(); ++i) if (hd[i]) { cudaSetDevice(L[i]->device); cudaFree(hd[i]); } };
        for (size_t i = 0; i < L.size(); ++i) {
            nb_ctx *x = L[i];
            CK(cudaSetDevice(x->device));
            CK(cudaMalloc(&hd[i], sizeof(unsigned long long) * kHdr));
            if (x->comm_rank == 0) {
                if (!x->built) { cleanup(); return fail("replicate_scene: rank 0 has no built scene"); }
                unsigned long long h[kHdr] = { x->nodes.n, x->tris.n, x->verts.n, x->normals.n, x->uvs.n, x->faces.n, x->dmeshes.n, x->cdf.n,
                                               x->emitters.n, x->n_nodes, x->n_prims, x->top_nodes, (unsigned long long) x->bvh_depth,
                                               (unsigned long long) x->builder_used, x->arena_bytes, 0 };
                CK(cudaMemcpyAsync(hd[i], h, sizeof h, cudaMemcpyHostToDevice, x->stream));
                CK(cudaStreamSynchronize(x->stream));
            }
        }
        NCK(N.GroupStart());
        for (size_t i = 0; i < L.size(); ++i) {
            CK(cudaSetDevice(L[i]->device));
            NCK(N.Broadcast(hd[i], hd[i], sizeof(unsigned long long) * kHdr, ncclChar, 0, (ncclComm_t) L[i]->comm, L[i]->stream));
        }
        NCK(N.GroupEnd());
        for (size_t i = 0; i < L.size(); ++i) {
            nb_ctx *x = L[i];
            if (x->comm_rank == 0) continue;
            CK(cudaSetDevice(x->device));
            unsigned long long h[kHdr];
            CK(cudaMemcpyAsync(h, hd[i], sizeof h, cudaMemcpyDeviceToHost, x->stream));
            CK(cudaStreamSynchronize(x->stream));
            size_t counts[9], offs[9];
            for (int k = 0; k < 9; ++k) counts[k] = (size_t) h[k];
            const size_t total = arena_layout(counts, offs);
            if (total != (size_t) h[14]) { cleanup(); return fail("replicate_scene: arena layout mismatch between ranks (%zu vs %llu bytes)", total, h[14]); }
            arena_release(x);
            CK(cudaMalloc(&x->arena_d, total + kArenaSlack));
            x->arena_bytes = total;
            arena_views(x, counts, offs);
            x->n_nodes = (uint32_t) h[9]; x->n_prims = (uint32_t) h[10]; x->top_nodes = (uint32_t) h[11]; x->bvh_depth = (int) h[12];
            x->builder_used = (int) h[13]; x->build_seconds = 0; x->built = true;
        }
        cleanup();
        NCK(N.GroupStart());
        for (nb_ctx *x : L) {
            CK(cudaSetDevice(x->device));
            NCK(N.Broadcast(x->arena_d, x->arena_d, x->arena_bytes, ncclChar, 0, (ncclComm_t) x->comm, x->stream));
        }
        NCK(N.GroupEnd());
    }
    for (nb_ctx *x : L) if (!x->built || !x->arena_d) return fail("replicate_scene: rank %d has no scene buffers (nb_build_accel first)", x->comm_rank);
    const size_t B = c->arena_bytes;
    const size_t S = ((B + (size_t) nranks - 1) / (size_t) nranks + kArenaAlign - 1) / kArenaAlign * kArenaAlign;     // shard bytes; nranks * S <= B + slack
    if (with_header) {
        // process-per-GPU ranks > 0: keep this rank's shard in pinned host memory (the source of later sharded uploads)
        for (nb_ctx *x : L) {
            if (x->comm_rank == 0 || x->leader) continue;
            CK(cudaSetDevice(x->device));
            if (S > x->shard_cap) {
                if (x->shard_h) cudaFreeHost(x->shard_h);
                x->shard_h = nullptr; x->shard_cap = 0;
                CK(cudaMallocHost(&x->shard_h, S));
                x->shard_cap = S;
            }
            const size_t off = (size_t) x->comm_rank * S;
            const size_t n = off < B ? std::min(S, B - off) : 0;
            if (n) CK(cudaMemcpyAsync(x->shard_h, x->arena_d + off, n, cudaMemcpyDeviceToHost, x->stream));
        }
    } else {
        for (nb_ctx *x : L) {
            CK(cudaSetDevice(x->device));
            const size_t off = (size_t) x->comm_rank * S;
            const size_t n = off < B ? std::min(S, B - off) : 0;
            const char *src = x->arena_h ? x->arena_h + off : (x->leader && x->leader->arena_h ? x->leader->arena_h + off : x->shard_h);
            if (n && !src) return fail("replicate_scene: rank %d has no host copy of its scene shard", x->comm_rank);
            if (n) CK(cudaMemcpyAsync(x->arena_d + off, src, n, cudaMemcpyHostToDevice, x->stream));
        }
        NCK(N.GroupStart());
        for (nb_ctx *x : L) {
            CK(cudaSetDevice(x->device));
            NCK(N.AllGather(x->arena_d + (size_t) x->comm_rank * S, x->arena_d, S, ncclChar, (ncclComm_t) x->comm, x->stream));
        }
        NCK(N.GroupEnd());
    }
    for (nb_ctx *x : L) { CK(cudaSetDevice(x->device)); CK(cudaStreamSynchronize(x->stream)); }
    CK(cudaSetDevice(c->device));
    return 0;
}

// One frame on the whole group.  film (device memory of rank 0; ignored elsewhere) receives the merged un-normalised film.
// `s0` is the stream of the calling context (the other local contexts use their own).  Stats: summed over the LOCAL
// contexts; kernel_ms is the slowest local render kernel, total_ms the calling context's whole step.
int render_group(nb_ctx *c, float4 *film, cudaStream_t s0, nb_stats *st) {
    if (need_nccl()) return 1;
    NcclApi &N = nccl();
    std::vector<nb_ctx *> L = locals(c);
    const int nranks = c->comm_nranks;
    if (!c->have_camera) return fail("nb_set_camera has not been called");
    const int edge = NB_BLOCK_SIZE + 2 * c->border;
    int ntx = 0, nty = 0;
    const int n_max = tiles_for(c, 0, nranks, &ntx, &nty);          // rank 0 owns the most tiles (tile_id % nranks)
    const size_t per_rank = (size_t) n_max * edge * edge;            // float4 elements
    for (nb_ctx *x : L) {
        CK(cudaSetDevice(x->device));
        x->tile_rank = x->comm_rank; x->tile_nranks = nranks;
        if (x->comm_rank == 0) {
            const size_t want = per_rank * (size_t) nranks;
            if (want > x->gather_cap) {
                if (x->gather) cudaFree(x->gather);
                x->gather = nullptr; x->gather_cap = 0;
                CK(cudaMalloc(&x->gather, sizeof(float4) * (want ? want : 1)));
                x->gather_cap = want;
            }
        } else if (per_rank > x->send_cap) {
            if (x->send_blocks) cudaFree(x->send_blocks);
            x->send_blocks = nullptr; x->send_cap = 0;
            CK(cudaMalloc(&x->send_blocks, sizeof(float4) * (per_rank ? per_rank : 1)));
            x->send_cap = per_rank;
        }
    }
    // ---- every device renders its tiles (asynchronous launches, one stream per device)
    if (L.size() > 1 && c->opt_engine == 2) {
        // the wavefront engine's host loop synchronises with its stream: one host thread per device, or they would run in turn
        std::vector<std::thread> th;
        std::vector<std::string> errs(L.size());
        for (size_t i = 0; i < L.size(); ++i)
            th.emplace_back([&, i] {
                nb_ctx *x = L[i];
                cudaStream_t s = (x == c) ? s0 : x->stream;
                if (render_tiles(x, x->comm_rank == 0 ? x->gather : x->send_blocks, s, &x->last_st, nullptr)) errs[i] = nb_last_error();
            });
        for (auto &t : th) t.join();
        for (const std::string &e : errs) if (!e.empty()) return fail("%s", e.c_str());
    } else {
        for (nb_ctx *x : L) {
            cudaStream_t s = (x == c) ? s0 : x->stream;
            if (render_tiles(x, x->comm_rank == 0 ? x->gather : x->send_blocks, s, &x->last_st, nullptr)) return 1;
        }
    }
    // ---- ONE exchange per frame: finished ImageBlocks to rank 0
    if (per_rank) {
        NCK(N.GroupStart());
        for (nb_ctx *x : L) {
            CK(cudaSetDevice(x->device));
            cudaStream_t s = (x == c) ? s0 : x->stream;
            if (x->comm_rank == 0) {
                for (int p = 1; p < nranks; ++p)
                    NCK(N.Recv(x->gather + (size_t) p * per_rank, per_rank * 4, ncclFloat, p, (ncclComm_t) x->comm, s));
            } else {
                NCK(N.Send(x->send_blocks, per_rank * 4, ncclFloat, 0, (ncclComm_t) x->comm, s));
            }
        }
        NCK(N.GroupEnd());
    }
    // ---- rank 0: ONE merge launch (ImageBlock::put(ImageBlock&), ref: src/block.cpp:93-102)
    for (nb_ctx *x : L) {
        if (x->comm_rank != 0) continue;
        if (!film) return fail("rank 0 needs a film buffer");
        CK(cudaSetDevice(x->device));
        cudaStream_t s = (x == c) ? s0 : x->stream;
        const size_t film_elems = (size_t) (x->W + 2 * x->border) * (x->H + 2 * x->border);
        CK(cudaMemsetAsync(film, 0, sizeof(float4) * film_elems, s));
        const long long total = (long long) nranks * n_max * edge * edge;
        if (total) {
            if (ensure_tile_table(x, nranks)) return 1;
            nb::merge_all_blocks_kernel<<<(int) ((total + 255) / 256), 256, 0, s>>>(x->gather, nranks, n_max, ntx * nty, x->tile_tab_d, x->W, x->H, x->border, edge, film);
            CK(cudaGetLastError());
        }
    }
    CK(cudaSetDevice(c->device));
    if (!st) return 0;                 // enqueue only; render_group_stats() synchronises later
    return render_group_stats(c, s0, st);
}

// Counters and timings of the local contexts after render_group (synchronises their streams).
int render_group_stats(nb_ctx *c, cudaStream_t s0, nb_stats *st) {
    std::vector<nb_ctx *> L = locals(c);
    nb_stats sum; memset(&sum, 0, sizeof sum);
    for (nb_ctx *x : L) {
        CK(cudaSetDevice(x->device));
        cudaStream_t s = (x == c) ? s0 : x->stream;
        if (finish_stats(x, s, &x->last_st, (x->comm_rank == 0) ? 1 : 0)) return 1;
        const nb_stats &l = x->last_st;
        sum.samples += l.samples; sum.rays += l.rays; sum.node_visits += l.node_visits; sum.tri_tests += l.tri_tests;
        sum.hits_shaded += l.hits_shaded; sum.launches += l.launches;
        sum.kernel_ms = std::max(sum.kernel_ms, l.kernel_ms);
        if (x == c) sum.total_ms = l.total_ms;
    }
    CK(cudaSetDevice(c->device));
    *st = sum;
    return 0;
}

void release_group(nb_ctx *c) {
    if (c->comm) { NcclApi &N = nccl(); if (N.CommDestroy) { cudaSetDevice(c->device); N.CommDestroy((ncclComm_t) c->comm); } c->comm = nullptr; }
    if (c->gather) { cudaFree(c->gather); c->gather = nullptr; c->gather_cap = 0; }
    if (c->send_blocks) { cudaFree(c->send_blocks); c->send_blocks = nullptr; c->send_cap = 0; }
}

}  // namespace nbm

extern "C" {

nb_ctx *nb_create_multi(const int *devices, int ndev) {
    if (!devices || ndev < 1) { fail("nb_create_multi: need at least one device"); return nullptr; }
    for (int i = 0; i < ndev; ++i) for (int j = 0; j < i; ++j) if (devices[i] == devices[j]) { fail("nb_create_multi: device %d listed twice", devices[i]); return nullptr; }
    nb_ctx *lead = nb_create(devices[0]);
    if (!lead) return nullptr;
    if (ndev == 1) return lead;
    if (nbm::need_nccl()) { nb_destroy(lead); return nullptr; }
    for (int i = 1; i < ndev; ++i) {
        nb_ctx *f = nb_create(devices[i]);
        if (!f) { nb_destroy(lead); return nullptr; }
        f->leader = lead;
        lead->followers.push_back(f);
    }
    std::vector<ncclComm_t> comms((size_t) ndev);
    ncclResult_t r = nbm::nccl().CommInitAll(comms.data(), ndev, devices);
    if (r != ncclSuccess) { fail("ncclCommInitAll failed: %s", nbm::nccl().GetErrorString(r)); nb_destroy(lead); return nullptr; }
    std::vector<nb_ctx *> L = nbm::locals(lead);
    for (int i = 0; i < ndev; ++i) { L[(size_t) i]->comm = comms[(size_t) i]; L[(size_t) i]->comm_rank = i; L[(size_t) i]->comm_nranks = ndev; }
    cudaSetDevice(lead->device);
    return lead;
}

int nb_device_count(nb_ctx *c) {
    if (!c) return 0;
    return c->leader ? c->leader->comm_nranks : c->comm_nranks;
}

int nb_comm_get_unique_id(uint8_t id[NB_COMM_ID_BYTES]) {
    if (!id) return fail("null argument");
    if (nbm::need_nccl()) return 1;
    static_assert(sizeof(ncclUniqueId) == NB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NCK(nbm::nccl().GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return 0;
}

int nb_comm_init_rank(nb_ctx *c, const uint8_t id[NB_COMM_ID_BYTES], int rank, int nranks) {
    if (!c || !id) return fail("null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("invalid rank %d of %d", rank, nranks);
    if (!c->followers.empty() || c->leader) return fail("nb_comm_init_rank: context already belongs to an nb_create_multi group");
    if (c->comm) return fail("nb_comm_init_rank: context already has a communicator");
    if (nbm::need_nccl()) return 1;
    if (ensure_device(c)) return 1;
    ncclUniqueId u; memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    NCK(nbm::nccl().CommInitRank(&comm, nranks, u, rank));
    c->comm = comm; c->comm_rank = rank; c->comm_nranks = nranks;
    c->tile_rank = rank; c->tile_nranks = nranks;
    return 0;
}

int nb_render_gather(nb_ctx *c, float *film_dev, void *stream, nb_stats *st) {
    if (!c) return fail("null context");
    if (c->leader) return fail("nb_render_gather: call it on the group's leader context");
    if (ensure_device(c)) return 1;
    cudaStream_t s = stream ? (cudaStream_t) stream : c->stream;
    if (!nbm::grouped(c)) {            // a group of one: the plain device render
        if (!film_dev) return fail("null film");
        return nb_render_device(c, film_dev, stream, st);
    }
    if (c->comm_rank == 0 && !film_dev) return fail("rank 0 needs a film buffer");
    return nbm::render_group(c, reinterpret_cast<float4 *>(film_dev), s, st);
}

}  // extern "C"
