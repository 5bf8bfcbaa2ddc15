/*
 * nori_b200.h -- C-ABI of the Blackwell-native render hot path for Nori (libnori_b200.so).
 *
 * The reference (wjakob/nori @ 092f581) has NO FFI: plugins are translation units linked
 * into one executable (ref: CMakeLists.txt:40-91, include/nori/object.h:141-149).  The
 * boundary below is therefore defined by the three seams of the reference's own host code
 * that the GPU path replaces; each entry point cites the reference
 * interface it stands in for.  Plain C: pointers and sizes only, no C++/torch types, no
 * exceptions, no callbacks.  Every int-returning function returns 0 on success and
 * nonzero on failure with a thread-local message in nb_last_error() -- the host wrapper
 * rethrows it as NoriException (ref: include/nori/common.h:135-140).
 *
 * Threading: one nb_ctx is driven by one host thread at a time (the reference's render
 * thread, ref: src/main.cpp:78).  There is no CPU fallback anywhere behind this API:
 * nb_create fails if no CUDA device is usable, and unsupported plugin types are errors.
 */
#ifndef NORI_B200_H
#define NORI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_ABI_VERSION 2
#define NB_BLOCK_SIZE 32          /* ref: include/nori/block.h:17  NORI_BLOCK_SIZE */
#define NB_FILTER_RESOLUTION 32   /* ref: include/nori/rfilter.h:12 */
#define NB_MISS 0xffffffffu

/* BSDF plugins the device path knows (ref: src/diffuse.cpp, src/mirror.cpp, src/dielectric.cpp, src/microfacet.cpp) */
enum { NB_BSDF_DIFFUSE = 0, NB_BSDF_MIRROR = 1, NB_BSDF_DIELECTRIC = 2, NB_BSDF_MICROFACET = 3 };
/* Emitter plugins (interface ref: include/nori/emitter.h:16-24; "area" is authored) */
enum { NB_EMITTER_NONE = 0, NB_EMITTER_AREA = 1 };
/* Integrator plugins named by the shipped scenes (ref: scenes/pa1/bunny.xml:8, pa3/ajax-ao.xml:8,
 * pa4/cbox/cbox-whitted.xml:4, pa5/cbox/cbox_{mats,ems,mis}.xml:4); interface ref: include/nori/integrator.h:42 */
enum { NB_INT_NORMALS = 0, NB_INT_AO = 1, NB_INT_WHITTED = 2, NB_INT_PATH_MATS = 3, NB_INT_PATH_EMS = 4, NB_INT_PATH_MIS = 5,
       NB_INT_SIMPLE = 6 /* point light, needs nb_set_point_light (ref: scenes/pa3/ajax-simple.xml:8-11) */ };
/* Sampler stream assignment.  PER_BLOCK is the reference's Independent::prepare (ref: src/independent.cpp:36-41):
 * one sequential pcg32 stream per 32x32 block.  PER_SAMPLE seeds one stream per (pixel, sample) through the
 * generate()/advance() hooks of the Sampler API (ref: include/nori/sampler.h:66-83) and is the parallel-friendly mode. */
enum { NB_SEED_PER_SAMPLE = 0, NB_SEED_PER_BLOCK = 1 };

typedef struct nb_bsdf_desc {
    int32_t type;
    float   albedo[3];      /* Diffuse "albedo" (ref: src/diffuse.cpp:19) / Microfacet "kd" (ref: src/microfacet.cpp:27) */
    float   alpha;          /* Microfacet "alpha" (ref: src/microfacet.cpp:18) */
    float   intIOR, extIOR; /* ref: src/microfacet.cpp:21-24, src/dielectric.cpp:17-20 */
    float   ks;             /* 1 - max(kd) (ref: src/microfacet.cpp:36) */
} nb_bsdf_desc;

typedef struct nb_emitter_desc {
    int32_t type;
    float   radiance[3];
} nb_emitter_desc;

typedef struct nb_integrator_desc {
    int32_t type;
    int32_t rr_start;   /* first bounce index at which Russian roulette applies (path_*); <=0 -> 3 */
    int32_t max_depth;  /* cap on path vertices; <=0 -> 1<<20 */
    int32_t reserved;
} nb_integrator_desc;

/* Batched equivalent of the arguments of Accel::rayIntersect (ref: include/nori/accel.h:54, include/nori/ray.h:30-34) */
typedef struct nb_ray { float o[3]; float mint; float d[3]; float maxt; } nb_ray;     /* 32 B */
typedef struct nb_hit { float t, u, v; uint32_t prim; uint32_t mesh; } nb_hit;        /* 20 B; prim = global triangle index, NB_MISS = none */

typedef struct nb_stats {
    uint64_t samples;       /* camera samples rendered by this context (its tiles only) */
    uint64_t rays;          /* camera + extension + shadow rays traced (device counter) */
    uint64_t node_visits;   /* BVH node visits (device counter) */
    uint64_t tri_tests;     /* ray/triangle tests (device counter) */
    uint64_t hits_shaded;   /* closest hits for which an intersection record was filled */
    double   kernel_ms;     /* device time of the render kernel(s) of the last nb_render*, CUDA events on the launch stream */
    double   total_ms;      /* device time of the whole last nb_render* call incl. clears, merge and copies */
    uint64_t launches;      /* kernels launched by the last call */
    uint64_t h2d_bytes, d2h_bytes;  /* bytes copied by the last call */
} nb_stats;

typedef struct nb_ctx nb_ctx;

/* Context owns all device memory and streams for ONE device.  device < 0 selects the current device.
 * Replaces: the Accel instance a Scene owns (ref: src/scene.cpp:16-18). */
nb_ctx *nb_create(int device);
void    nb_destroy(nb_ctx *);
const char *nb_last_error(void);
int     nb_abi_version(void);
/* Bytes one node visit of the walk fetches: 64 (binary node: two child boxes + two references) or 80 (8-wide compressed node,
 * nori_b200/csrc/nb_wide.h) -- the per-visit figure of the algorithmic-bytes accounting (bench.py). */
int     nb_node_bytes(void);

/* ---- N GPUs behind the same calls (replaces the TBB tile loop + merge for N devices, ref: src/main.cpp:85-113,
 * src/block.cpp:93-102).  nb_create_multi returns ONE context (on devices[0]) that owns a context per further device
 * and the NCCL communicators between them (ncclCommInitAll; libnccl.so.2 is loaded on first use).  Every scene / camera /
 * sampler / integrator / option call on it applies to all devices; nb_build_accel builds the hierarchy once and replicates
 * the scene arrays over NVLink (ncclBroadcast); nb_render / nb_render_device shard the 32x32 tiles tile_id % ndev, gather
 * the finished ImageBlocks on devices[0] with ONE grouped ncclSend/ncclRecv, merge them with ONE launch and return the
 * full film.  ndev == 1 is exactly nb_create(devices[0]).  nb_device_count: devices behind a context (1 for nb_create). */
nb_ctx *nb_create_multi(const int *devices, int ndev);
int     nb_device_count(nb_ctx *);
/* One process per GPU (torchrun, MPI): rank 0 draws a communicator id, the caller ships the 128 bytes to every rank by any
 * means, each rank attaches its context.  From then on the context is rank `rank` of an `nranks`-GPU group: nb_build_accel
 * builds on rank 0 only and replicates over NVLink (the other ranks need no meshes), nb_upload_scene crosses PCIe once,
 * and nb_render / nb_render_gather render this rank's tile shard, gather on rank 0 and merge there (film is written on
 * rank 0 only and may be NULL elsewhere).  All ranks must make the same sequence of these calls (they are collective). */
#define NB_COMM_ID_BYTES 128
int nb_comm_get_unique_id(uint8_t id[NB_COMM_ID_BYTES]);
int nb_comm_init_rank(nb_ctx *, const uint8_t id[NB_COMM_ID_BYTES], int rank, int nranks);
/* Device-film variant of the group render; stats == NULL only enqueues (no host synchronisation). */
int nb_render_gather(nb_ctx *, float *film_dev, void *stream, nb_stats *stats);

/* Replaces Accel::addMesh (ref: src/accel.cpp:12-17; called from Scene::addChild, ref: src/scene.cpp:48-53).
 * V: 3*nv floats packed xyz == m_V 3xN column-major (ref: include/nori/mesh.h:160); N (3*nv) and UV (2*nv)
 * nullable; F: 3*nf uint32 == m_F.  Caller keeps ownership, callee copies.  Lifts the single-mesh
 * limit (ref: src/accel.cpp:13-14).  bsdf NULL -> default diffuse 0.5 (ref: src/mesh.cpp:23-29).
 * Returns the mesh id (>=0) or -1. */
int nb_add_mesh(nb_ctx *, const float *V, uint32_t nv, const float *N, const float *UV,
                const uint32_t *F, uint32_t nf, const nb_bsdf_desc *bsdf, const nb_emitter_desc *emitter);
int nb_clear_meshes(nb_ctx *);

/* Replaces Accel::build (ref: src/accel.cpp:19-21; called from Scene::activate, ref: src/scene.cpp:28):
 * host SAH BVH build + SoA upload. */
int nb_build_accel(nb_ctx *);
/* Seconds the last nb_build_accel spent building the hierarchy and which builder ran: 0 = host binned SAH (default),
 * 1 = device LBVH (nb_set_option(ctx, "builder", 1); Morton sort + Karras radix tree on the GPU, ~100x faster to
 * build, lower tree quality).  Results are identical with either tree. */
int nb_build_stats(nb_ctx *, double *seconds, int *builder);
/* On-disk cache of the built hierarchy (host SAH builder): with a path set, nb_build_accel first looks for a file whose key
 * -- a 64-bit hash of every vertex and index the builder reads, the build parameters and the layout version -- matches, and
 * loads nodes + leaf-ordered triangles from it instead of building; otherwise it builds and writes the file (silently
 * skipped if the directory is read-only).  NULL or "" disables.  nb_accel_cache_hit: 1 if the last build was served by it. */
int nb_set_accel_cache(nb_ctx *, const char *path);
int nb_accel_cache_hit(nb_ctx *);
/* Re-uploads the already built scene arrays from pinned host memory (used to time host->device traffic). */
int nb_upload_scene(nb_ctx *);

/* PerspectiveCamera state (ref: src/perspective.cpp:22-39,41-68): row-major 4x4 sampleToCamera and
 * cameraToWorld, output size, clip planes.  Ray generation follows ref: src/perspective.cpp:76-97. */
int nb_set_camera(nb_ctx *, const float s2c[16], const float c2w[16], int width, int height, float nearClip, float farClip);

/* Reconstruction filter as ImageBlock tabulates it (ref: src/block.cpp:19-27): table[i] = filter->eval(radius*i/32),
 * table[32] = 0, evaluated by the HOST plugin so any ReconstructionFilter plugin works unchanged. */
int nb_set_filter(nb_ctx *, const float table[NB_FILTER_RESOLUTION + 1], float radius);

/* Sampler state (ref: src/independent.cpp:23-25 "sampleCount"). */
int nb_set_sampler(nb_ctx *, uint32_t spp, int seed_mode, uint64_t seed);

/* Integrator selection; unsupported type -> error (no CPU fallback). */
int nb_set_integrator(nb_ctx *, const nb_integrator_desc *);
/* Properties of the `simple` integrator (ref: scenes/pa3/ajax-simple.xml:9-10): <point name="position">, <color name="energy">.
 * Li = energy / (4 pi^2) * max(0, cos theta) / |x - position|^2 * V(x <-> position).  Required before rendering NB_INT_SIMPLE. */
int nb_set_point_light(nb_ctx *, const float position[3], const float energy[3]);

/* Li of n independent camera paths over the whole image plane -- the loop of the reference's t-test in scene mode
 * (ref: src/ttest.cpp:153-167): pixelSample = next2D() * outputSize, apertureSample = next2D(), value = Li(ray);
 * lum_host[k] receives value.getLuminance() (ref: src/common.cpp:206-208) of path k.  The reference consumes one
 * sequential sampler stream across all paths; here path k owns the pcg32 stream seed((seed << 32) + k, 0), seed being
 * the one given to nb_set_sampler.  Needs scene, camera and integrator; film, filter and tiling are not used. */
int nb_li_samples(nb_ctx *, uint64_t n, float *lum_host, nb_stats *stats /* nullable */);

/* Batched BSDF::sample() and BSDF::eval() + pdf() (ref: include/nori/bsdf.h:59-87) in local shading coordinates, for the
 * callers of the BSDF plugins outside the render loop -- the reference's t-test in BSDF mode and its chi^2 test
 * (ref: src/ttest.cpp:104-125, src/chi2test.cpp:113-153).  No scene is needed.  wi: 3 floats, shared by the whole batch
 * (wi_per_query = 0; both tests fix wi) or 3 per query (wi_per_query = 1).
 *   nb_bsdf_sample:   xi 2 floats per query; out8 = wo.xyz, weight.rgb (= eval * cos / pdf, 0 for a failed sample),
 *                     pdf(wi, wo) (0 for discrete lobes), measure (1 solid angle, 2 discrete)
 *   nb_bsdf_eval_pdf: wo 3 floats per query; out4 = eval.rgb, pdf  (solid-angle measure) */
int nb_bsdf_sample(nb_ctx *, const nb_bsdf_desc *, const float *wi, int wi_per_query, const float *xi, uint64_t n, float *out8);
int nb_bsdf_eval_pdf(nb_ctx *, const nb_bsdf_desc *, const float *wi, int wi_per_query, const float *wo, uint64_t n, float *out4);

/* Tile sharding across GPUs: this context renders only 32x32 tiles with tile_id % nranks == rank
 * (the numbering depends on nranks: ownership follows the Latin pattern (bx + shift * by) % nranks, shift = 3 -- 5 or 7 when 3 divides nranks --, over the ceil(W/32) x ceil(H/32) tile grid -- every row and column of tiles is dealt evenly to all ranks -- and rank r's k-th tile, row by row, is tile k * nranks + r).  Default (0, 1) = all tiles.  Replaces BlockGenerator::next
 * as the work scheduler (ref: src/block.cpp:119-152). */
int nb_set_tiles(nb_ctx *, int rank, int nranks);

/* Replaces the body of render() (ref: src/main.cpp:61-119): BlockGenerator + TBB loop + renderBlock +
 * ImageBlock::put x2.  Writes the (H+2b) x (W+2b) x 4 fp32 row-major UN-normalised weighted film including
 * the border -- byte compatible with ImageBlock storage (ref: include/nori/block.h:35, src/block.cpp:36) so that
 * toBitmap / EXR / GUI code is unchanged.  film_host is caller-owned HOST memory. */
int nb_render(nb_ctx *, float *film_host, nb_stats *stats);

/* Same, but the film stays on the device (film_dev: device pointer on the context's device) and all work is
 * enqueued on `stream` (a cudaStream_t passed as void*; NULL = the context's own stream).  Used by the
 * multi-GPU driver, which exchanges films with NCCL. */
int nb_render_device(nb_ctx *, float *film_dev, void *stream, nb_stats *stats);   /* stats == NULL: enqueue only, no host sync */
/* Device time of the render kernel(s) of this context's last (finished) render call: CUDA events recorded around them on
 * the launch stream.  For callers that enqueue frames without statistics and read the kernel time afterwards. */
int nb_last_kernel_ms(nb_ctx *, double *ms);

/* Finished ImageBlocks of this context's tiles, packed: blocks_dev receives ntiles_mine x (32+2b) x (32+2b) x 4 fp32
 * (tile order = ascending tile_id of the tiles owned by (rank, nranks)); the frame-end exchange gathers these. */
int nb_render_blocks_device(nb_ctx *, float *blocks_dev, void *stream, nb_stats *stats);   /* stats == NULL: enqueue only, no host sync */
/* Number of tiles owned by (rank, nranks) for the current camera, and the block edge (32 + 2*border). */
int nb_tile_count(nb_ctx *, int rank, int nranks, int *ntiles, int *block_edge);
/* Adds packed blocks of (rank, nranks) into a full film on the device: the merge of ImageBlock::put(ImageBlock&)
 * (ref: src/block.cpp:93-102).  film_dev must be zeroed by the caller before the first merge. */
int nb_merge_blocks_device(nb_ctx *, const float *blocks_dev, int rank, int nranks, float *film_dev, void *stream);

/* The same merge for the gathered blocks of ALL ranks in one launch: blocks_dev = [nranks][stride_tiles][edge][edge][4]
 * (every rank padded to stride_tiles blocks, as gathered over NCCL). */
int nb_merge_all_blocks_device(nb_ctx *, const float *blocks_dev, int nranks, int stride_tiles, float *film_dev, void *stream);

/* Batched Scene::rayIntersect (ref: include/nori/scene.h:63-65,82-85 -> src/accel.cpp:23-43).  HOST buffers.
 * shadow != 0: any-hit query, hits[i].prim = 0 if occluded else NB_MISS. */
int nb_intersect(nb_ctx *, const nb_ray *rays, uint64_t n, nb_hit *hits, int shadow, nb_stats *stats);
/* Device-buffer variant (rays_dev / hits_dev on the context's device). */
int nb_intersect_device(nb_ctx *, const nb_ray *rays_dev, uint64_t n, nb_hit *hits_dev, int shadow, void *stream, nb_stats *stats);
/* Full intersection records (ref: src/accel.cpp:45-96): out16[i] = p(3) t uv(2) shFrame.s(3) .t(3) .n(3) mesh. */
int nb_intersect_full(nb_ctx *, const nb_ray *rays, uint64_t n, float *out16);

/* Film normalisation: ImageBlock::toBitmap (ref: src/block.cpp:45-51, include/nori/color.h:100-105). HOST buffers. */
int nb_film_to_rgb(nb_ctx *, const float *film_host, float *rgb_host);

/* Output path on the device: normalisation (ImageBlock::toBitmap, ref: src/block.cpp:45-51), sRGB tonemap
 * (Color3f::toSRGB, ref: src/common.cpp:166-180) and 8-bit quantisation (ref: src/bitmap.cpp:100-110) of the film the last
 * nb_render left on the device, in one kernel; rgb8_host receives W x H x 3 bytes, ready for the PNG writer.  The bytes
 * equal the host loop's (both sides evaluate x^(1/2.4) with the same polynomials). */
int nb_last_film_to_srgb8(nb_ctx *, uint8_t *rgb8_host);

/* Progressive frames -- the role of NoriScreen, which redraws from the shared ImageBlock while tiles are still being rendered
 * (ref: src/gui.cpp:120-138): the frame is rendered in passes of n sample streams per pixel; the accumulators stay on the
 * device, and after any pass nb_render_preview merges them into a film (film_host, nullable: (H+2b) x (W+2b) x 4 floats)
 * and/or the tonemapped 8-bit image (rgb8_host, nullable: W x H x 3).  The un-normalised film of k passes is exactly the
 * film of the first k * n samples, so every preview is a valid image and the last one equals nb_render's (up to the order
 * of the film atomics).  nb_render_pass stats accumulate over the passes.  Single-device contexts, NB_SEED_PER_SAMPLE. */
int nb_render_begin(nb_ctx *);
int nb_render_pass(nb_ctx *, uint32_t n_samples, nb_stats *stats /* nullable */);
int nb_render_preview(nb_ctx *, float *film_host, uint8_t *rgb8_host);
int nb_render_end(nb_ctx *);

/* Tuning knobs (optional; sane defaults): key/value, see DESIGN.md section 6.  Unknown key -> error. */
int nb_set_option(nb_ctx *, const char *key, int64_t value);
/* Raw device counters of the last call (diagnostics; meaningful with option "count" = 1): [1] rays, [2] node visits,
 * [3] triangle tests, [4] hits shaded, [5] sum over lock-step waves of the LONGEST walk in the warp, [6] waves. */
int nb_debug_counters(nb_ctx *, uint64_t out[8]);
/* Host-only diagnostic (no context, no GPU): runs the SAH builder that nb_build_accel uses on verts4 (xyz + pad per vertex)
 * / faces4 (i0, i1, i2, mesh per triangle) and returns the device layout -- 16 floats per node, 12 floats per leaf-ordered
 * triangle (nori_b200/csrc/nb_bvh.h).  info = { nodes, leaf triangles, breadth-first top nodes, depth }.  Call with null
 * outputs to size the arrays (capacities in floats).  Returns 0, 1 (bad argument) or 2 (capacity too small). */
int nb_debug_build_bvh(const float *verts4, const uint32_t *faces4, uint32_t nprims, int max_leaf, int64_t bfs_nodes,
                       float *nodes_out, uint64_t nodes_cap, float *tris_out, uint64_t tris_cap, uint32_t info[4]);
/* Host-only diagnostic of the tile numbering (no context, no device): out[t] = bx | by << 16 of tile t for a group of
 * `nranks` (see nb_set_tiles); ceil(width/32) * ceil(height/32) entries.  Returns 0, 1 (bad argument) or 2 (capacity). */
int nb_debug_tile_order(int width, int height, int nranks, uint32_t *bx_by_out, uint64_t cap);
/* Host-only diagnostic of the work-unit schedule of one render launch (no context, no device): a rank that owns `n_tiles`
 * tiles, `spp` samples per pixel, `resident_warps` = SMs x resident CTAs x 4, and the options chunk / guided / coarse of
 * nb_set_option (guided < 0 = default).  out = { fine chunk, fine chunks per patch, first fine sample, coarse chunk, coarse
 * chunks per patch, coarse units, units }.  Returns 0, 1 (bad argument) or 2 (more than 0xf0000000 units). */
int nb_debug_unit_plan(int n_tiles, uint32_t spp, int64_t resident_warps, int64_t chunk, int64_t guided, int64_t coarse, uint32_t out[7]);
/* Host-only diagnostic of the hierarchy cache: key -> load, else build + save, exactly as nb_build_accel does with
 * nb_set_accel_cache.  info = { nodes, leaf triangles, top nodes, depth, hit (0/1) }. */
int nb_debug_bvh_cache(const float *verts4, const uint32_t *faces4, uint32_t nprims, int max_leaf, int64_t bfs_nodes, const char *path,
                       float *nodes_out, uint64_t nodes_cap, float *tris_out, uint64_t tris_cap, uint32_t info[5]);
/* Host-only diagnostics of the 8-wide compressed hierarchy (nori_b200/csrc/nb_wide.h): binary SAH build (max_leaf 3) + collapse;
 * nodes_out = 20 words per wide node, tris_out = 12 floats per triangle; info = { wide nodes, triangles, wide depth, binary
 * nodes }.  nb_debug_wide_intersect walks it on the HOST with the kernels' own node step (rays: o, mint, d, maxt = 8 floats;
 * hits4: t, u, v, triangle bits (NB_MISS = none); counts (nullable) = { node visits, triangle tests }): the CPU tests use it
 * to check the builder and the walk's logic against the oracle's brute-force loop without a GPU. */
int nb_debug_build_wide(const float *verts4, const uint32_t *faces4, uint32_t nprims, uint32_t *nodes_out, uint64_t nodes_cap,
                        float *tris_out, uint64_t tris_cap, uint32_t info[4]);
int nb_debug_wide_intersect(const uint32_t *nodes, uint32_t nnodes, const float *tris, const float *rays, uint64_t nrays,
                            int any_hit, float *hits4, uint64_t counts[2]);
/* Scene geometry summary after nb_build_accel. */
int nb_scene_info(nb_ctx *, uint64_t *ntris, uint64_t *nnodes, uint64_t *scene_bytes, int *bvh_depth);

#ifdef __cplusplus
}
#endif
#endif
