// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// render.h -- the render() driver of the host (ref: src/main.cpp:58-148) on top of the C-ABI in include/nori_b200.h.
#pragma once
#include "block.h"
#include "scene.h"

struct nb_ctx;
struct nb_stats;
struct nb_bsdf_desc;

NORI_NAMESPACE_BEGIN

struct RenderOptions {
    int device = 0;            // CUDA device of this process (first device of the group when gpus > 1)
    int gpus = 1;              // devices driven by this process: tiles sharded tile_id % gpus, ONE NCCL gather per frame (nb_create_multi)
    int tileRank = 0, tileRanks = 1;   // tile shard (multi-GPU: one process per GPU)
    bool quiet = false;
    bool deviceBuilder = false;        // build the hierarchy on the GPU (LBVH) instead of the host SAH builder
    int previewEvery = 0;              // > 0: render progressively, that many samples per pass, and rewrite <scene>_preview.png after every pass
    std::string previewName;           // file name stem of the preview image
    std::string accelCache;            // file caching the built hierarchy (nb_set_accel_cache); empty = build every time
};

/// Builds the GPU context for a scene: meshes + plugin descriptors (from the factory's creation records), BVH,
/// camera, tabulated filter, sampler, integrator.  Throws NoriException on anything the device path cannot run.
nb_ctx *createDeviceScene(const Scene *scene, const ImageBlock &film, const RenderOptions &opt);

/// POD descriptor of a BSDF plugin instance (from its creation record); throws for plugins without a device implementation.
void describeBSDF(const BSDF *bsdf, nb_bsdf_desc *out);

/// Replaces the body of render(): fills `result` (the full-image ImageBlock) through nb_render.
/// srgb8 (optional): the W x H x 3 tonemapped 8-bit image, produced on the device from the film it still holds
/// (nb_last_film_to_srgb8) -- what the PNG writer consumes.
void renderScene(Scene *scene, ImageBlock &result, const RenderOptions &opt, nb_stats *stats = nullptr, std::vector<uint8_t> *srgb8 = nullptr);

/// Full driver: render + toBitmap + EXR/PNG next to the scene file (ref: src/main.cpp:127-147)
void render(Scene *scene, const std::string &filename, const RenderOptions &opt);

NORI_NAMESPACE_END
