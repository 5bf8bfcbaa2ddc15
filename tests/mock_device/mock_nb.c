/* mock_nb.c -- TEST DOUBLE of the C-ABI (include/nori_b200.h) for CPU-only tests of the HOST objects that sit on top of
 * it (the `ttest` / `chi2test` scene objects in nori_b200/csrc/host/stat_tests.cpp).
 *
 * It is built by tests/test_host_stat_objects_cpu.py into a temporary directory as "libnori_b200.so" next to a copy of
 * libnori_host.so, and is never placed in nori_b200/lib, never shipped, and renders nothing: nb_render fails.  What it
 * answers, in closed form and without the oracle:
 *   nb_bsdf_sample / nb_bsdf_eval_pdf : a DIFFUSE BSDF only (cosine-hemisphere sampling, eval = albedo / pi,
 *                                       pdf = cos / pi); the environment variable MOCK_NB_PDF_SCALE skews the sampled
 *                                       distribution so that a chi^2 test must reject it
 *   nb_li_samples                     : lum[k] = u_k + MOCK_NB_LI_SHIFT with u_k uniform in [0,1) (mean 0.5, variance 1/12)
 * so that the statistics, binning, quadrature, verdicts and error handling of the host objects can be exercised here. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "nori_b200.h"

struct nb_ctx { int dummy; };
static const char *g_err = "";

static double env_double(const char *name, double dflt) { const char *v = getenv(name); return v ? atof(v) : dflt; }

const char *nb_last_error(void) { return g_err; }
nb_ctx *nb_create(int device) { (void) device; return (nb_ctx *) calloc(1, sizeof(nb_ctx)); }
void nb_destroy(nb_ctx *c) { free(c); }
int nb_node_bytes(void) { return 64; }
nb_ctx *nb_create_multi(const int *devices, int ndev) { (void) ndev; return nb_create(devices ? devices[0] : 0); }
int nb_set_accel_cache(nb_ctx *c, const char *path) { (void) c; (void) path; return 0; }
int nb_last_film_to_srgb8(nb_ctx *c, uint8_t *rgb8) { (void) c; (void) rgb8; return 1; }
int nb_render_begin(nb_ctx *c) { (void) c; return 1; }
int nb_render_pass(nb_ctx *c, uint32_t n, nb_stats *st) { (void) c; (void) n; (void) st; return 1; }
int nb_render_preview(nb_ctx *c, float *f, uint8_t *r) { (void) c; (void) f; (void) r; return 1; }
int nb_render_end(nb_ctx *c) { (void) c; return 0; }
int nb_set_option(nb_ctx *c, const char *k, int64_t v) { (void) c; (void) k; (void) v; return 0; }
int nb_add_mesh(nb_ctx *c, const float *V, uint32_t nv, const float *N, const float *UV, const uint32_t *F, uint32_t nf,
                const nb_bsdf_desc *b, const nb_emitter_desc *e) { (void) c; (void) V; (void) nv; (void) N; (void) UV; (void) F; (void) nf; (void) b; (void) e; return 0; }
int nb_build_accel(nb_ctx *c) { (void) c; return 0; }
int nb_set_camera(nb_ctx *c, const float s2c[16], const float c2w[16], int w, int h, float n, float f) { (void) c; (void) s2c; (void) c2w; (void) w; (void) h; (void) n; (void) f; return 0; }
int nb_set_filter(nb_ctx *c, const float t[NB_FILTER_RESOLUTION + 1], float r) { (void) c; (void) t; (void) r; return 0; }
int nb_set_sampler(nb_ctx *c, uint32_t spp, int mode, uint64_t seed) { (void) c; (void) spp; (void) mode; (void) seed; return 0; }
int nb_set_integrator(nb_ctx *c, const nb_integrator_desc *d) { (void) c; (void) d; return 0; }
int nb_set_point_light(nb_ctx *c, const float p[3], const float e[3]) { (void) c; (void) p; (void) e; return 0; }
int nb_set_tiles(nb_ctx *c, int r, int n) { (void) c; (void) r; (void) n; return 0; }
int nb_render(nb_ctx *c, float *film, nb_stats *st) { (void) c; (void) film; (void) st; g_err = "mock device: nb_render is not available"; return 1; }

static uint64_t lcg(uint64_t *s) { *s = *s * 6364136223846793005ULL + 1442695040888963407ULL; return *s >> 11; }

int nb_li_samples(nb_ctx *c, uint64_t n, float *lum, nb_stats *st) {
    (void) c; (void) st;
    uint64_t s = 12345;
    const double shift = env_double("MOCK_NB_LI_SHIFT", 0.0);
    for (uint64_t k = 0; k < n; ++k) lum[k] = (float) ((double) lcg(&s) / 9007199254740992.0 + shift);
    return 0;
}

int nb_bsdf_sample(nb_ctx *c, const nb_bsdf_desc *b, const float *wi, int per_query, const float *xi, uint64_t n, float *out8) {
    (void) c;
    if (b->type != NB_BSDF_DIFFUSE) { g_err = "mock device: diffuse only"; return 1; }
    const double skew = env_double("MOCK_NB_PDF_SCALE", 1.0);     /* 1 = cosine-weighted; other values sample cos^skew */
    for (uint64_t k = 0; k < n; ++k) {
        const float *w = wi + (per_query ? 3 * k : 0);
        float *o = out8 + 8 * k;
        memset(o, 0, 8 * sizeof(float));
        o[7] = 1.0f;
        if (w[2] <= 0) continue;
        const double u = xi[2 * k], v = xi[2 * k + 1];
        const double cosT = pow(1.0 - u, 1.0 / (1.0 + skew)), sinT = sqrt(fmax(0.0, 1.0 - cosT * cosT)), phi = 2.0 * M_PI * v;
        o[0] = (float) (sinT * cos(phi)); o[1] = (float) (sinT * sin(phi)); o[2] = (float) cosT;
        o[3] = b->albedo[0]; o[4] = b->albedo[1]; o[5] = b->albedo[2];
        o[6] = (float) (cosT / M_PI);
    }
    return 0;
}

int nb_bsdf_eval_pdf(nb_ctx *c, const nb_bsdf_desc *b, const float *wi, int per_query, const float *wo, uint64_t n, float *out4) {
    (void) c;
    if (b->type != NB_BSDF_DIFFUSE) { g_err = "mock device: diffuse only"; return 1; }
    for (uint64_t k = 0; k < n; ++k) {
        const float *w = wi + (per_query ? 3 * k : 0), *v = wo + 3 * k;
        float *o = out4 + 4 * k;
        memset(o, 0, 4 * sizeof(float));
        if (w[2] <= 0 || v[2] <= 0) continue;
        o[0] = (float) (b->albedo[0] / M_PI); o[1] = (float) (b->albedo[1] / M_PI); o[2] = (float) (b->albedo[2] / M_PI);
        o[3] = (float) (v[2] / M_PI);
    }
    return 0;
}
