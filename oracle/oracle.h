/*
 * oracle.h -- CPU restatement of the Nori render hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for nori_b200.  It is NOT part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (libnori_b200.so) never links or calls it.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  Parts of the path that have NO implementation in the
 * reference (integrators, area emitter, warps, Microfacet/Dielectric bodies -- the
 * reference ships them as stubs) are "authored": they follow
 * the reference's interfaces and are pinned by the reference's own statistical
 * fixtures (scenes/pa4/tests, scenes/pa5/tests), restated in tests/.
 *
 * Parity pin status (details in DESIGN.md):
 *   - pcg32            : pinned by the pcg-random.org KAT (seed 42/54) -- the reference
 *                        repo itself holds no RNG vectors (ext/pcg32 is an empty submodule).
 *   - Microfacet, warps: pinned by scenes/pa5/tests/{ttest,chi2test}-microfacet.xml.
 *   - whitted/path_*   : pinned by scenes/pa4/tests/test-mesh*.xml, scenes/pa5/tests/test-*.xml.
 *   - image-level      : "parity unpinned" against the reference binary (the reference
 *                        cannot be compiled here: every ext/ submodule is empty).
 */
#ifndef NORI_ORACLE_H
#define NORI_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- plain-data descriptors (layout mirrors include/nori_b200.h on purpose, so
 *      one set of ctypes structures in the tests can feed both sides) ---- */
enum { ORC_BSDF_DIFFUSE = 0, ORC_BSDF_MIRROR = 1, ORC_BSDF_DIELECTRIC = 2, ORC_BSDF_MICROFACET = 3 };
enum { ORC_EMITTER_NONE = 0, ORC_EMITTER_AREA = 1 };
enum { ORC_INT_NORMALS = 0, ORC_INT_AO = 1, ORC_INT_WHITTED = 2,
       ORC_INT_PATH_MATS = 3, ORC_INT_PATH_EMS = 4, ORC_INT_PATH_MIS = 5, ORC_INT_SIMPLE = 6 };
enum { ORC_SEED_PER_SAMPLE = 0, ORC_SEED_PER_BLOCK = 1 };
enum { ORC_ACCEL_BRUTE = 0, ORC_ACCEL_BVH = 1 };

typedef struct {
    int32_t type;
    float   albedo[3];      /* Diffuse: albedo; Microfacet: kd */
    float   alpha;          /* Microfacet roughness */
    float   intIOR, extIOR; /* Microfacet / Dielectric */
    float   ks;             /* Microfacet: 1 - max(kd) (src/microfacet.cpp:36) */
} orc_bsdf;

typedef struct {
    int32_t type;
    float   radiance[3];
} orc_emitter;

typedef struct {
    int32_t type;
    int32_t rr_start;   /* bounce index from which Russian roulette applies (path_*) */
    int32_t max_depth;  /* hard cap on path vertices (safety; 0 = default 1<<20) */
    int32_t reserved;
} orc_integrator;

typedef struct { float o[3]; float mint; float d[3]; float maxt; } orc_ray;   /* 32 B */
typedef struct { float t, u, v; uint32_t prim; uint32_t mesh; } orc_hit;      /* prim = global triangle index; 0xffffffff = miss */

typedef struct {
    uint64_t samples;      /* camera samples (W*H*spp) */
    uint64_t rays;         /* camera + extension + shadow rays traced */
    uint64_t node_visits;  /* BVH node visits (0 for brute force) */
    uint64_t tri_tests;    /* ray/triangle tests */
    double   seconds;      /* tile loop only (== src/main.cpp:83..118 Timer placement) */
} orc_stats;

typedef struct orc_scene orc_scene;

/* ---- scene assembly (mirrors Scene::addChild/activate, src/scene.cpp:27-79) ---- */
orc_scene *orc_scene_create(void);
void       orc_scene_destroy(orc_scene *);
/* V: 3*nv packed xyz (== m_V 3xN col-major, include/nori/mesh.h:160), N/UV nullable, F: 3*nf */
int  orc_scene_add_mesh(orc_scene *, const float *V, uint32_t nv, const float *N, const float *UV,
                        const uint32_t *F, uint32_t nf, const orc_bsdf *, const orc_emitter *);
int  orc_scene_build(orc_scene *);   /* builds the CPU BVH + emitter CDFs */
void orc_scene_set_camera(orc_scene *, const float s2c[16], const float c2w[16], int W, int H, float nearClip, float farClip);
void orc_scene_set_filter(orc_scene *, const float table[33], float radius);
void orc_scene_set_sampler(orc_scene *, uint32_t spp, int seed_mode, uint64_t seed);
void orc_scene_set_integrator(orc_scene *, const orc_integrator *);
void orc_scene_set_point_light(orc_scene *, const float position[3], const float energy[3]);  /* `simple` integrator */
void orc_scene_set_tiles(orc_scene *, int rank, int nranks);  /* render only tiles with id % nranks == rank */

/* ---- the path ---- */
/* film: (H+2b) x (W+2b) x 4 fp32, row-major, un-normalised weighted film incl. border
 * (== ImageBlock storage, include/nori/block.h:35, src/block.cpp:36). */
int  orc_render(orc_scene *, float *film, int accel_kind, int nthreads, orc_stats *);
int  orc_intersect(orc_scene *, const orc_ray *, uint64_t n, orc_hit *, int shadow, int accel_kind, orc_stats *);
/* full intersection record, for hit-fill parity: out[16] = p(3) t uv(2) sh.s(3) sh.t(3) sh.n(3) mesh */
int  orc_intersect_full(orc_scene *, const orc_ray *, uint64_t n, float *out16, int accel_kind);
/* Li of n camera paths the way ttest scene mode draws them (src/ttest.cpp:153-167): one sequential
 * default-constructed... see oracle.c.  lum[n] receives the luminance of each path. */
int  orc_ttest_scene(orc_scene *, uint64_t n, int accel_kind, double *lum);
/* the same loop with one stream per path (path k: seed((seed << 32) + k, 0)): the oracle of nb_li_samples */
int  orc_li_samples(orc_scene *, uint64_t n, int accel_kind, float *lum);
int  orc_film_to_rgb(const float *film, int W, int H, int border, float *rgb);  /* toBitmap, src/block.cpp:45-51 */
int  orc_block_order(int W, int H, int block, int32_t *xy);  /* BlockGenerator spiral (src/block.cpp:109-152); returns count */

/* ---- unit-level entry points used by the fixture tests ---- */
typedef struct { uint64_t state, inc; } orc_pcg32;
void     orc_pcg32_init(orc_pcg32 *);                                  /* default ctor */
void     orc_pcg32_seed(orc_pcg32 *, uint64_t initstate, uint64_t initseq);
uint32_t orc_pcg32_next_uint(orc_pcg32 *);
float    orc_pcg32_next_float(orc_pcg32 *);
void     orc_pcg32_advance(orc_pcg32 *, int64_t delta);

void  orc_sincos2pi(float u, float *s, float *c);
float orc_logf(float x);
float orc_expf(float x);
void  orc_square_to_cosine_hemisphere(const float xi[2], float out[3]);
float orc_square_to_cosine_hemisphere_pdf(const float v[3]);
void  orc_square_to_beckmann(const float xi[2], float alpha, float out[3]);
float orc_square_to_beckmann_pdf(const float m[3], float alpha);
float orc_fresnel(float cosThetaI, float extIOR, float intIOR);
void  orc_coordinate_system(const float a[3], float b[3], float c[3]);
/* returns the sample weight (eval*cos/pdf); wo/eta/measure written (measure: 1 solid angle, 2 discrete) */
void  orc_bsdf_sample(const orc_bsdf *, const float wi[3], const float xi[2], float wo[3], float *eta, int *measure, float weight[3]);
void  orc_bsdf_eval(const orc_bsdf *, const float wi[3], const float wo[3], float out[3]);
float orc_bsdf_pdf(const orc_bsdf *, const float wi[3], const float wo[3]);
void  orc_bsdf_eval_pdf_batch(const orc_bsdf *, const float wi[3], const float *wo, uint64_t n, float *out4);
/* n samples from one continuing rng stream (ref: src/ttest.cpp:116-118, src/chi2test.cpp:113-115); wo_out nullable */
int   orc_bsdf_sample_batch(const orc_bsdf *, const float wi[3], uint64_t n, orc_pcg32 *rng, float *wo_out, float *weight_out);
void  orc_filter_table(int kind, float radius, float stddev, float B, float C, float table[33], float *radius_out);
void  orc_camera_matrices(float fov, float nearClip, float farClip, int W, int H, float s2c[16]);
void  orc_sample_ray(const orc_scene *, float sx, float sy, orc_ray *out);

#ifdef __cplusplus
}
#endif
#endif
