// nb_ctx.h -- private definition of nb_ctx, shared by the translation units of libnori_b200.so (nb_api.cu: kernels and
// single-device entry points; nb_multi.cu: NCCL communicators and the multi-device render).  Not part of the C-ABI.
#pragma once
#include "../../include/nori_b200.h"
#include "nb_device.cuh"

#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace nbi {

int fail(const char *fmt, ...);     // sets the thread-local message of nb_last_error(), returns 1

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return nbi::fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct HostMesh {
    std::vector<float> V, N, UV;
    std::vector<uint32_t> F;
    uint32_t nv = 0, nf = 0;
    nb_bsdf_desc bsdf;
    nb_emitter_desc emitter;
};

template <typename T>
struct DevBuf {
    T *d = nullptr; T *h = nullptr; size_t n = 0;   // device + pinned host mirror (h stays null for device-only buffers)
    bool view = false;                               // a window into the context's scene arena: not owned
    void release() { if (!view) { if (d) cudaFree(d); if (h) cudaFreeHost(h); } d = nullptr; h = nullptr; n = 0; view = false; }
    cudaError_t alloc(size_t count, bool with_host = true) {
        release(); n = count;
        size_t bytes = sizeof(T) * (count ? count : 1);
        cudaError_t e = cudaMalloc(&d, bytes); if (e != cudaSuccess) return e;
        return with_host ? cudaMallocHost(&h, bytes) : cudaSuccess;
    }
    void set_view(char *dev_base, char *host_base, size_t offset, size_t count) {
        release();
        d = reinterpret_cast<T *>(dev_base + offset); h = host_base ? reinterpret_cast<T *>(host_base + offset) : nullptr;
        n = count; view = true;
    }
    size_t bytes() const { return sizeof(T) * n; }
};

}  // namespace nbi

struct nb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    int sm_count = 0;
    std::vector<nbi::HostMesh> meshes;
    // scene tables (device + pinned mirrors)
    nbi::DevBuf<float4> nodes, tris, verts, normals;
    nbi::DevBuf<float2> uvs;
    nbi::DevBuf<uint4> faces;
    nbi::DevBuf<nb::DevMesh> dmeshes;
    nbi::DevBuf<float> cdf;
    nbi::DevBuf<int32_t> emitters;
    // After a build the nine tables are windows into ONE allocation (the scene arena): one host->device copy uploads the scene,
    // one ncclBroadcast replicates it, and a group re-uploads it SHARDED -- every rank sends 1/N over its own PCIe link and one
    // in-place ncclAllGather over NVLink completes it (nb_multi.inl).
    char *arena_d = nullptr, *arena_h = nullptr; size_t arena_bytes = 0;
    char *shard_h = nullptr; size_t shard_cap = 0;   // process-per-GPU ranks > 0: pinned copy of this rank's shard of the arena
    uint32_t n_nodes = 0, n_prims = 0, top_nodes = 0; int bvh_depth = 0; bool built = false;
    double build_seconds = 0;
    // camera / film / sampler / integrator
    float s2c[16], c2w[16]; int W = 0, H = 0; float nearClip = 1e-4f, farClip = 1e4f; bool have_camera = false;
    float ftable[33]; float fradius = 2.0f; int border = 2;
    uint32_t spp = 1; int seed_mode = NB_SEED_PER_SAMPLE; uint64_t seed = 0;
    nb_integrator_desc integ = { NB_INT_NORMALS, 3, 0, 0 };
    float light_pos[3] = { 0, 0, 0 }, light_energy[3] = { 0, 0, 0 }; bool have_light = false;
    int tile_rank = 0, tile_nranks = 1;
    std::vector<uint32_t> tile_tab_h; uint32_t *tile_tab_d = nullptr; int tab_W = 0, tab_H = 0, tab_N = 0;   // tile numbering for tab_N ranks (ensure_tile_table)
    // work buffers
    float4 *blocks = nullptr; size_t blocks_cap = 0;
    float4 *film = nullptr; size_t film_cap = 0; bool film_valid = false;   // film_valid: holds the last nb_render's film
    unsigned long long *counters = nullptr;          // 8 x u64 device
    unsigned long long *counters_h = nullptr;        // pinned
    // options
    int64_t opt_blocks_per_sm = 0, opt_smem_nodes = 0, opt_chunk = 0, opt_count = 0, opt_max_leaf = 3,
            opt_bfs_nodes = 2048, opt_builder = 0, opt_engine = 0, opt_occ_tail = 20;
    // wavefront engine (nb_wave.cu): path pool (structure of arrays), extension queue, occlusion queue, counters
    float4 *wf_cols = nullptr, *wf_shadow = nullptr; uint32_t *wf_ext = nullptr, *wf_ctr = nullptr, *wf_ctr_h = nullptr; size_t wf_cap = 0;
    int64_t opt_wf_pool = 1 << 21, opt_wf_check = 4;
    bool prog_active = false; uint32_t prog_done = 0, prog_pass = 0; nb_stats prog_stats = {};   // progressive frame (nb_render_begin .. nb_render_end)
    int64_t opt_sah_bins = 32;         // SAH bins per axis of the host builder
    int64_t opt_coarse = 8;            // guided schedule: samples per coarse work unit (upper bound)
    int64_t opt_guided = 75;           // fused kernel: percent of the samples scheduled in coarse work units (0 = plain schedule)
    int64_t opt_prefetch = 0;          // L2 warm-up of nodes + triangles before the render kernel (l2_prefetch_kernel)
    int builder_used = 0;   // 0 host SAH, 1 device LBVH
    std::string accel_cache; bool accel_cache_hit = false;   // on-disk hierarchy cache (nb_set_accel_cache)
    // ---- multi-GPU (nb_multi.inl)
    std::vector<nb_ctx *> followers;   // nb_create_multi: the contexts on the other devices (owned by this leader)
    nb_ctx *leader = nullptr;          // set on followers
    void *comm = nullptr;              // ncclComm_t of this context (single-process group or nb_comm_init_rank)
    int comm_rank = 0, comm_nranks = 1;
    float4 *gather = nullptr; size_t gather_cap = 0;   // rank 0: blocks of all ranks, [nranks][n_max][edge][edge]
    float4 *send_blocks = nullptr; size_t send_cap = 0; // ranks > 0: own blocks padded to n_max tiles
    nb_stats last_st = {};             // statistics of this context's last render inside a group
};
