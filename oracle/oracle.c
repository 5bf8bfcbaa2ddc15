/*
 * oracle.c -- CPU restatement of the Nori render hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h).  Plain C99, scalar fp32, compiled with -ffp-contract=off so that every
 * multiply and add rounds separately -- the CUDA path is compiled with -fmad=false and
 * follows the same operation order, which is what makes image-level parity attainable
 * (SURVEY.md section 7 "Path divergence under float differences").
 *
 * Citations "ref:" are relative to /root/reference (wjakob/nori @ 092f581).
 * "[authored]" marks functions whose body does not exist in the reference (stubs that
 * throw); they follow the reference's interfaces and the spec in DESIGN.md section 3, and
 * are pinned by the reference's statistical fixtures restated under tests/.
 *
 * Operation-order notes (Eigen is absent from the container, so these are [recalled]
 * and unpinned; effect <= 1 ulp): 3-element dot products and squared norms reduce as
 * a0 + (a1 + a2) (Eigen's unrolled redux tree); linear combinations of vectors evaluate
 * left to right per coefficient; 4x4 * vec4 accumulates column by column.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>

/* ------------------------------------------------------------------ constants */
#define ORC_EPSILON   1e-4f                    /* ref: include/nori/common.h:38 */
#define ORC_PI        3.14159265358979323846f  /* ref: include/nori/common.h:43 */
#define ORC_INV_PI    0.31830988618379067154f  /* ref: include/nori/common.h:44 */
#define ORC_INV_FOURPI2 0.025330295910584444f  /* 1 / (4 pi^2) */
#define ORC_BLOCK     32                       /* ref: include/nori/block.h:17 */
#define ORC_FILTER_RES 32                      /* ref: include/nori/rfilter.h:12 */
#define ORC_MISS      0xffffffffu

typedef struct { float x, y, z; } v3;

static inline v3 v3make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3sub(v3 a, v3 b) { return v3make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3add(v3 a, v3 b) { return v3make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3scale(v3 a, float s) { return v3make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3neg(v3 a) { return v3make(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
static inline v3 cross3(v3 a, v3 b) {
    return v3make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 normalize3(v3 a) {
    float n = sqrtf(dot3(a, a));
    return v3make(a.x / n, a.y / n, a.z / n);
}
static inline v3 ld3(const float *p) { return v3make(p[0], p[1], p[2]); }

/* ------------------------------------------------------------------ pcg32
 * Restates wjakob/pcg32 @ 70099ead (pcg32.h; un-vendored submodule).
 * Call sites: ref: src/independent.cpp:36-55. */
#define PCG32_DEFAULT_STATE  0x853c49e6748fea9bULL
#define PCG32_DEFAULT_STREAM 0xda3e39cb94b95bdbULL
#define PCG32_MULT           0x5851f42d4c957f2dULL

void orc_pcg32_init(orc_pcg32 *r) { r->state = PCG32_DEFAULT_STATE; r->inc = PCG32_DEFAULT_STREAM; }

uint32_t orc_pcg32_next_uint(orc_pcg32 *r) {
    uint64_t old = r->state;
    r->state = old * PCG32_MULT + r->inc;
    uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xs >> rot) | (xs << ((~rot + 1u) & 31u));
}

void orc_pcg32_seed(orc_pcg32 *r, uint64_t initstate, uint64_t initseq) {
    r->state = 0u;
    r->inc = (initseq << 1u) | 1u;
    orc_pcg32_next_uint(r);
    r->state += initstate;
    orc_pcg32_next_uint(r);
}

float orc_pcg32_next_float(orc_pcg32 *r) {
    union { uint32_t u; float f; } x;
    x.u = (orc_pcg32_next_uint(r) >> 9) | 0x3f800000u;
    return x.f - 1.0f;
}

void orc_pcg32_advance(orc_pcg32 *r, int64_t delta_) {
    uint64_t cur_mult = PCG32_MULT, cur_plus = r->inc, acc_mult = 1u, acc_plus = 0u;
    uint64_t delta = (uint64_t) delta_;
    while (delta > 0) {
        if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta /= 2;
    }
    r->state = acc_mult * r->state + acc_plus;
}

/* ------------------------------------------------------------------ deterministic transcendentals
 * [authored] The reference calls libm (sincosf/logf/expf).  glibc and CUDA's libdevice differ
 * in the last ulp, which would decorrelate paths; both sides therefore evaluate the SAME
 * polynomial sequences (Cephes single-precision kernels) with separately rounded mul/add. */
void orc_sincos2pi(float u, float *so, float *co) {
    /* angle = 2*pi*u, u in [0,1).  Octant reduction is exact: u*8 is a power-of-two scale. */
    float u8 = u * 8.0f;
    int k = (int) u8;                 /* 0..7 (8 only if u==1, folded below) */
    float f = u8 - (float) k;         /* exact */
    int j = (k + 1) >> 1;             /* nearest multiple of pi/2 */
    float r = (k & 1) ? (f - 1.0f) : f;
    float x = r * 0.78539816339744830962f;   /* in [-pi/4, pi/4] */
    float z = x * x;
    float s = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * x + x;
    float c = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z
              - 0.5f * z + 1.0f;
    switch (j & 3) {
        case 0: *so = s;  *co = c;  break;
        case 1: *so = c;  *co = -s; break;
        case 2: *so = -s; *co = -c; break;
        default: *so = -c; *co = s; break;
    }
}

float orc_logf(float xin) {
    /* valid for normal positive floats (callers pass (0,1]) */
    union { float f; uint32_t u; } b; b.f = xin;
    int e = (int) ((b.u >> 23) & 0xff) - 126;        /* frexp exponent: x = m * 2^e, m in [0.5,1) */
    b.u = (b.u & 0x007fffffu) | 0x3f000000u;
    float x = b.f;
    if (x < 0.70710678118654752440f) { e -= 1; x = x + x - 1.0f; } else { x = x - 1.0f; }
    float z = x * x;
    float y = ((((((((7.0376836292e-2f * x - 1.1514610310e-1f) * x + 1.1676998740e-1f) * x
              - 1.2420140846e-1f) * x + 1.4249322787e-1f) * x - 1.6668057665e-1f) * x
              + 2.0000714765e-1f) * x - 2.4999993993e-1f) * x + 3.3333331174e-1f) * x * z;
    float fe = (float) e;
    y = y + (-2.12194440e-4f * fe);
    y = y + (-0.5f * z);
    z = x + y;
    z = z + 0.693359375f * fe;
    return z;
}

float orc_expf(float x) {
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    float z = floorf(1.44269504088896341f * x + 0.5f);
    x = x - z * 0.693359375f;
    x = x - z * -2.12194440e-4f;
    int n = (int) z;
    z = x * x;
    z = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x
        + 4.1665795894e-2f) * x + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    union { float f; uint32_t u; } s; s.u = (uint32_t) (n + 127) << 23;   /* 2^n, n in [-126,127] */
    return z * s.f;
}

/* ------------------------------------------------------------------ warps [authored]
 * ref interface: include/nori/warp.h:44-57; bodies throw in src/warp.cpp:53-67. */
static inline v3 sq2coshemi(float x, float y) {
    float r = sqrtf(x), s, c;
    orc_sincos2pi(y, &s, &c);
    float zz = 1.0f - x;
    return v3make(r * c, r * s, sqrtf(zz > 0.0f ? zz : 0.0f));
}
void orc_square_to_cosine_hemisphere(const float xi[2], float out[3]) {
    v3 r = sq2coshemi(xi[0], xi[1]); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float orc_square_to_cosine_hemisphere_pdf(const float v[3]) { return v[2] <= 0.0f ? 0.0f : v[2] * ORC_INV_PI; }

static inline v3 sq2beckmann(float x, float y, float alpha) {
    float s, c;
    orc_sincos2pi(x, &s, &c);
    float tan2 = -(alpha * alpha) * orc_logf(1.0f - y);
    float ct = 1.0f / sqrtf(1.0f + tan2);
    float st2 = 1.0f - ct * ct;
    float st = sqrtf(st2 > 0.0f ? st2 : 0.0f);
    return v3make(st * c, st * s, ct);
}
void orc_square_to_beckmann(const float xi[2], float alpha, float out[3]) {
    v3 r = sq2beckmann(xi[0], xi[1], alpha); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
/* Beckmann normal distribution D(m) (without the cos factor) */
static inline float beckmann_D(v3 m, float alpha) {
    if (m.z <= 0.0f) return 0.0f;
    float c2 = m.z * m.z;
    float tan2 = (1.0f - c2) / c2;
    float a2 = alpha * alpha;
    return orc_expf(-tan2 / a2) / (ORC_PI * a2 * (c2 * c2));
}
float orc_square_to_beckmann_pdf(const float m[3], float alpha) {
    v3 mm = ld3(m);
    return mm.z <= 0.0f ? 0.0f : beckmann_D(mm, alpha) * mm.z;
}

/* ref: src/common.cpp:259-288 */
float orc_fresnel(float cosThetaI, float extIOR, float intIOR) {
    float etaI = extIOR, etaT = intIOR;
    if (extIOR == intIOR) return 0.0f;
    if (cosThetaI < 0.0f) { float t = etaI; etaI = etaT; etaT = t; cosThetaI = -cosThetaI; }
    float eta = etaI / etaT, sinThetaTSqr = eta * eta * (1 - cosThetaI * cosThetaI);
    if (sinThetaTSqr > 1.0f) return 1.0f;
    float cosThetaT = sqrtf(1.0f - sinThetaTSqr);
    float Rs = (etaI * cosThetaI - etaT * cosThetaT) / (etaI * cosThetaI + etaT * cosThetaT);
    float Rp = (etaT * cosThetaI - etaI * cosThetaT) / (etaT * cosThetaI + etaI * cosThetaT);
    return (Rs * Rs + Rp * Rp) / 2.0f;
}

/* ref: src/common.cpp:248-257 ; Frame(n): include/nori/frame.h:36-38 */
typedef struct { v3 s, t, n; } frame;
static inline frame frame_from_n(v3 a) {
    frame f; f.n = a;
    v3 c;
    if (fabsf(a.x) > fabsf(a.y)) {
        float invLen = 1.0f / sqrtf(a.x * a.x + a.z * a.z);
        c = v3make(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(a.y * a.y + a.z * a.z);
        c = v3make(0.0f, a.z * invLen, -a.y * invLen);
    }
    f.t = c;
    f.s = cross3(c, a);
    return f;
}
void orc_coordinate_system(const float a[3], float b[3], float c[3]) {
    frame f = frame_from_n(ld3(a));
    b[0] = f.s.x; b[1] = f.s.y; b[2] = f.s.z; c[0] = f.t.x; c[1] = f.t.y; c[2] = f.t.z;
}
/* ref: include/nori/frame.h:41-50 */
static inline v3 to_local(const frame *f, v3 v) { return v3make(dot3(v, f->s), dot3(v, f->t), dot3(v, f->n)); }
static inline v3 to_world(const frame *f, v3 v) {
    return v3make(f->s.x * v.x + f->t.x * v.y + f->n.x * v.z,
                  f->s.y * v.x + f->t.y * v.y + f->n.y * v.z,
                  f->s.z * v.x + f->t.z * v.y + f->n.z * v.z);
}

/* ------------------------------------------------------------------ BSDFs
 * interface ref: include/nori/bsdf.h:59-87.  measure: 1 = solid angle, 2 = discrete. */
static inline int bsdf_is_diffuse(const orc_bsdf *b) {   /* ref: src/diffuse.cpp:73, src/microfacet.cpp:59-64 */
    return b->type == ORC_BSDF_DIFFUSE || b->type == ORC_BSDF_MICROFACET;
}

/* [authored] G1 of the Beckmann shadowing-masking rational approximation */
static inline float mf_G1(v3 wv, v3 wh, float alpha) {
    if (dot3(wv, wh) / wv.z <= 0.0f) return 0.0f;
    float c2 = wv.z * wv.z;
    float s2 = 1.0f - c2;
    if (s2 <= 0.0f) return 1.0f;
    float tanv = sqrtf(s2) / wv.z;
    float b = 1.0f / (alpha * tanv);
    if (b >= 1.6f) return 1.0f;
    float b2 = b * b;
    return (3.535f * b + 2.181f * b2) / (1.0f + 2.276f * b + 2.577f * b2);
}

static v3 bsdf_eval(const orc_bsdf *b, v3 wi, v3 wo) {
    v3 zero = { 0, 0, 0 };
    switch (b->type) {
        case ORC_BSDF_DIFFUSE:   /* ref: src/diffuse.cpp:23-33 */
            if (wi.z <= 0 || wo.z <= 0) return zero;
            return v3make(b->albedo[0] * ORC_INV_PI, b->albedo[1] * ORC_INV_PI, b->albedo[2] * ORC_INV_PI);
        case ORC_BSDF_MICROFACET: {   /* [authored] contract ref: src/microfacet.cpp:40-42 */
            if (wi.z <= 0 || wo.z <= 0) return zero;
            v3 wh = normalize3(v3add(wi, wo));
            float D = beckmann_D(wh, b->alpha);
            float F = orc_fresnel(dot3(wh, wi), b->extIOR, b->intIOR);
            float G = mf_G1(wi, wh, b->alpha) * mf_G1(wo, wh, b->alpha);
            float spec = b->ks * D * F * G / (4.0f * wi.z * wo.z);
            return v3make(b->albedo[0] * ORC_INV_PI + spec, b->albedo[1] * ORC_INV_PI + spec,
                          b->albedo[2] * ORC_INV_PI + spec);
        }
        default: return zero;     /* discrete BSDFs evaluate to zero: ref src/mirror.cpp:17-20, src/dielectric.cpp:23-26 */
    }
}

static float bsdf_pdf(const orc_bsdf *b, v3 wi, v3 wo) {
    switch (b->type) {
        case ORC_BSDF_DIFFUSE:   /* ref: src/diffuse.cpp:36-52 */
            if (wi.z <= 0 || wo.z <= 0) return 0.0f;
            return ORC_INV_PI * wo.z;
        case ORC_BSDF_MICROFACET: {   /* [authored] */
            if (wi.z <= 0 || wo.z <= 0) return 0.0f;
            v3 wh = normalize3(v3add(wi, wo));
            float D = beckmann_D(wh, b->alpha);
            float Jh = 1.0f / (4.0f * dot3(wh, wo));
            return b->ks * D * wh.z * Jh + (1.0f - b->ks) * wo.z * ORC_INV_PI;
        }
        default: return 0.0f;
    }
}

/* returns weight = eval*cos/pdf; writes wo, eta, measure.  weight==0 <=> invalid sample */
static v3 bsdf_sample(const orc_bsdf *b, v3 wi, float xi_x, float xi_y, v3 *wo, float *eta, int *measure) {
    v3 zero = { 0, 0, 0 };
    *eta = 1.0f; *measure = 1; *wo = v3make(0, 0, 1);
    switch (b->type) {
        case ORC_BSDF_DIFFUSE:   /* ref: src/diffuse.cpp:55-71 */
            if (wi.z <= 0) return zero;
            *wo = sq2coshemi(xi_x, xi_y);
            return v3make(b->albedo[0], b->albedo[1], b->albedo[2]);
        case ORC_BSDF_MIRROR:    /* ref: src/mirror.cpp:27-43 */
            if (wi.z <= 0) return zero;
            *wo = v3make(-wi.x, -wi.y, wi.z);
            *measure = 2;
            return v3make(1, 1, 1);
        case ORC_BSDF_DIELECTRIC: {   /* [authored] stub at ref: src/dielectric.cpp:33-35 */
            float cosI = wi.z;
            float F = orc_fresnel(cosI, b->extIOR, b->intIOR);
            *measure = 2;
            if (xi_x < F) {
                *wo = v3make(-wi.x, -wi.y, wi.z);
                return v3make(1, 1, 1);
            }
            float etaI = b->extIOR, etaT = b->intIOR;
            if (cosI < 0.0f) { float t = etaI; etaI = etaT; etaT = t; cosI = -cosI; }
            float e = etaI / etaT;
            float sin2T = e * e * (1 - cosI * cosI);
            float cosT = sqrtf(1.0f - sin2T);
            *wo = v3make(-e * wi.x, -e * wi.y, wi.z > 0 ? -cosT : cosT);
            *eta = etaT / etaI;
            return v3make(1, 1, 1);
        }
        case ORC_BSDF_MICROFACET: {   /* [authored] contract ref: src/microfacet.cpp:50-57 */
            if (wi.z <= 0) return zero;
            if (xi_x < b->ks) {
                float x = xi_x / b->ks;
                v3 wh = sq2beckmann(x, xi_y, b->alpha);
                float d2 = 2.0f * dot3(wh, wi);
                *wo = v3make(d2 * wh.x - wi.x, d2 * wh.y - wi.y, d2 * wh.z - wi.z);
            } else {
                float x = (xi_x - b->ks) / (1.0f - b->ks);
                *wo = sq2coshemi(x, xi_y);
            }
            if (wo->z <= 0) return zero;
            v3 f = bsdf_eval(b, wi, *wo);
            float p = bsdf_pdf(b, wi, *wo);
            if (!(p > 0.0f)) return zero;
            return v3make(f.x * wo->z / p, f.y * wo->z / p, f.z * wo->z / p);
        }
        default: return zero;
    }
}

void orc_bsdf_sample(const orc_bsdf *b, const float wi[3], const float xi[2], float wo[3], float *eta,
                     int *measure, float weight[3]) {
    v3 o; v3 w = bsdf_sample(b, ld3(wi), xi[0], xi[1], &o, eta, measure);
    wo[0] = o.x; wo[1] = o.y; wo[2] = o.z; weight[0] = w.x; weight[1] = w.y; weight[2] = w.z;
}
void orc_bsdf_eval(const orc_bsdf *b, const float wi[3], const float wo[3], float out[3]) {
    v3 f = bsdf_eval(b, ld3(wi), ld3(wo)); out[0] = f.x; out[1] = f.y; out[2] = f.z;
}
float orc_bsdf_pdf(const orc_bsdf *b, const float wi[3], const float wo[3]) { return bsdf_pdf(b, ld3(wi), ld3(wo)); }
/* eval + pdf of n outgoing directions for one wi: out4 = eval.rgb, pdf (the checker of nb_bsdf_eval_pdf; also lets the
 * chi^2 fixture integrate pdf() with as many nodes as a narrow lobe needs) */
void orc_bsdf_eval_pdf_batch(const orc_bsdf *b, const float wi[3], const float *wo, uint64_t n, float *out4) {
    for (uint64_t k = 0; k < n; ++k) {
        v3 f = bsdf_eval(b, ld3(wi), ld3(wo + 3 * k));
        out4[4 * k] = f.x; out4[4 * k + 1] = f.y; out4[4 * k + 2] = f.z; out4[4 * k + 3] = bsdf_pdf(b, ld3(wi), ld3(wo + 3 * k));
    }
}

/* ------------------------------------------------------------------ reconstruction filters
 * ref: src/rfilter.cpp:16-108 (eval) tabulated as in src/block.cpp:19-27. kind: 0 gaussian 1 mitchell 2 tent 3 box */
static float filter_eval(int kind, float x, float radius, float stddev, float B, float C) {
    switch (kind) {
        case 0: {
            float alpha = -1.0f / (2.0f * stddev * stddev);
            float v = expf(alpha * x * x) - expf(alpha * radius * radius);
            return v > 0.0f ? v : 0.0f;
        }
        case 1: {
            x = fabsf(2.0f * x / radius);
            float x2 = x * x, x3 = x2 * x;
            if (x < 1) return 1.0f / 6.0f * ((12 - 9 * B - 6 * C) * x3 + (-18 + 12 * B + 6 * C) * x2 + (6 - 2 * B));
            else if (x < 2) return 1.0f / 6.0f * ((-B - 6 * C) * x3 + (6 * B + 30 * C) * x2 + (-12 * B - 48 * C) * x + (8 * B + 24 * C));
            else return 0.0f;
        }
        case 2: { float v = 1.0f - fabsf(x); return v > 0.0f ? v : 0.0f; }
        default: return 1.0f;
    }
}
void orc_filter_table(int kind, float radius, float stddev, float B, float C, float table[33], float *radius_out) {
    if (kind == 2) radius = 1.0f;      /* ref: src/rfilter.cpp:81-83 */
    if (kind == 3) radius = 0.5f;      /* ref: src/rfilter.cpp:97-99 */
    for (int i = 0; i < ORC_FILTER_RES; ++i) {
        float pos = (radius * i) / ORC_FILTER_RES;
        table[i] = filter_eval(kind, pos, radius, stddev, B, C);
    }
    table[ORC_FILTER_RES] = 0.0f;
    *radius_out = radius;
}

/* ------------------------------------------------------------------ scene */
typedef struct {
    float *V, *N, *UV;      /* packed xyz / xyz / uv (nullable N, UV) */
    uint32_t *F;
    uint32_t nv, nf;
    uint32_t prim_offset;   /* global index of triangle 0 */
    orc_bsdf bsdf;
    orc_emitter emitter;
    float *cdf;             /* nf+1 entries, normalised (ref: include/nori/dpdf.h) */
    float area_sum;
} mesh_t;

typedef struct { float lo[3], hi[3]; int32_t left, right; uint32_t start, count; } bvh_node;

struct orc_scene {
    mesh_t *meshes; int nmeshes;
    uint32_t nprims;
    uint32_t *prim_mesh;     /* per global prim: mesh index */
    int *emitters; int nemitters;
    /* BVH */
    bvh_node *nodes; uint32_t nnodes; uint32_t *prim_order;
    /* camera: ref src/perspective.cpp:22-39,76-97 */
    float s2c[16], c2w[16]; int W, H; float invW, invH, nearClip, farClip;
    /* film / filter: ref src/block.cpp:15-37 */
    float ftable[ORC_FILTER_RES + 1]; float fradius; int border; float lookup;
    uint32_t spp; int seed_mode; uint64_t seed;
    orc_integrator integ;
    float light_pos[3], light_energy[3];   /* point light of the `simple` integrator (scenes/pa3/ajax-simple.xml:8-11) */
    int tile_rank, tile_nranks;
};

orc_scene *orc_scene_create(void) {
    orc_scene *s = (orc_scene *) calloc(1, sizeof(orc_scene));
    s->spp = 1; s->tile_nranks = 1; s->integ.rr_start = 3;
    float r; orc_filter_table(0, 2.0f, 0.5f, 0, 0, s->ftable, &r);
    orc_scene_set_filter(s, s->ftable, r);
    return s;
}
void orc_scene_destroy(orc_scene *s) {
    if (!s) return;
    for (int i = 0; i < s->nmeshes; ++i) {
        free(s->meshes[i].V); free(s->meshes[i].N); free(s->meshes[i].UV); free(s->meshes[i].F); free(s->meshes[i].cdf);
    }
    free(s->meshes); free(s->prim_mesh); free(s->emitters); free(s->nodes); free(s->prim_order); free(s);
}
static void *dup_mem(const void *p, size_t n) { if (!p) return NULL; void *r = malloc(n ? n : 1); memcpy(r, p, n); return r; }

int orc_scene_add_mesh(orc_scene *s, const float *V, uint32_t nv, const float *N, const float *UV,
                       const uint32_t *F, uint32_t nf, const orc_bsdf *b, const orc_emitter *e) {
    s->meshes = (mesh_t *) realloc(s->meshes, sizeof(mesh_t) * (s->nmeshes + 1));
    mesh_t *m = &s->meshes[s->nmeshes];
    memset(m, 0, sizeof(*m));
    m->V = (float *) dup_mem(V, sizeof(float) * 3 * nv);
    m->N = (float *) dup_mem(N, sizeof(float) * 3 * nv);
    m->UV = (float *) dup_mem(UV, sizeof(float) * 2 * nv);
    m->F = (uint32_t *) dup_mem(F, sizeof(uint32_t) * 3 * nf);
    m->nv = nv; m->nf = nf; m->prim_offset = s->nprims;
    if (b) m->bsdf = *b; else { m->bsdf.type = ORC_BSDF_DIFFUSE; m->bsdf.albedo[0] = m->bsdf.albedo[1] = m->bsdf.albedo[2] = 0.5f; } /* ref: src/mesh.cpp:23-29, src/diffuse.cpp:19 */
    if (e) m->emitter = *e;
    s->nprims += nf;
    return s->nmeshes++;
}
void orc_scene_set_camera(orc_scene *s, const float s2c[16], const float c2w[16], int W, int H, float nearClip, float farClip) {
    memcpy(s->s2c, s2c, sizeof(float) * 16); memcpy(s->c2w, c2w, sizeof(float) * 16);
    s->W = W; s->H = H; s->invW = 1.0f / (float) W; s->invH = 1.0f / (float) H;   /* cwiseInverse: ref src/perspective.cpp:27 */
    s->nearClip = nearClip; s->farClip = farClip;
}
void orc_scene_set_filter(orc_scene *s, const float table[33], float radius) {
    if (table != s->ftable) memcpy(s->ftable, table, sizeof(float) * (ORC_FILTER_RES + 1));
    s->fradius = radius;
    s->border = (int) ceilf(radius - 0.5f);            /* ref: src/block.cpp:20 */
    s->lookup = ORC_FILTER_RES / radius;               /* ref: src/block.cpp:27 */
}
void orc_scene_set_sampler(orc_scene *s, uint32_t spp, int seed_mode, uint64_t seed) { s->spp = spp; s->seed_mode = seed_mode; s->seed = seed; }
void orc_scene_set_integrator(orc_scene *s, const orc_integrator *i) { s->integ = *i; if (s->integ.rr_start <= 0) s->integ.rr_start = 3; }
void orc_scene_set_point_light(orc_scene *s, const float position[3], const float energy[3]) {
    memcpy(s->light_pos, position, sizeof s->light_pos); memcpy(s->light_energy, energy, sizeof s->light_energy);
}
void orc_scene_set_tiles(orc_scene *s, int rank, int nranks) { s->tile_rank = rank; s->tile_nranks = nranks < 1 ? 1 : nranks; }

static inline void tri_verts(const mesh_t *m, uint32_t f, v3 *p0, v3 *p1, v3 *p2) {
    const uint32_t *idx = m->F + 3 * (size_t) f;
    *p0 = ld3(m->V + 3 * (size_t) idx[0]); *p1 = ld3(m->V + 3 * (size_t) idx[1]); *p2 = ld3(m->V + 3 * (size_t) idx[2]);
}

/* ------------------------------------------------------------------ CPU BVH (binned SAH).
 * The reference's Accel is brute force (ref: src/accel.cpp:30-43); a student BVH is its intended
 * replacement and TBoundingBox::rayIntersect (ref: include/nori/bbox.h:323-350) its node test.
 * Boxes are padded so the slab test can only cull what Moeller-Trumbore would also reject. */
typedef struct { float lo[3], hi[3], c[3]; } prim_box;
typedef struct { orc_scene *s; prim_box *pb; uint32_t cap; } bvh_build;

static void box_reset(float lo[3], float hi[3]) { for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
static void box_grow(float lo[3], float hi[3], const float l2[3], const float h2[3]) {
    for (int a = 0; a < 3; ++a) { if (l2[a] < lo[a]) lo[a] = l2[a]; if (h2[a] > hi[a]) hi[a] = h2[a]; }
}
static float box_area(const float lo[3], const float hi[3]) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0) return 0.0f;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

static uint32_t bvh_alloc(bvh_build *b) {
    orc_scene *s = b->s;
    if (s->nnodes == b->cap) { b->cap = b->cap ? b->cap * 2 : 1024; s->nodes = (bvh_node *) realloc(s->nodes, sizeof(bvh_node) * b->cap); }
    return s->nnodes++;
}

static uint32_t bvh_build_rec(bvh_build *b, uint32_t start, uint32_t end, int depth) {
    orc_scene *s = b->s;
    uint32_t ni = bvh_alloc(b);
    float lo[3], hi[3], clo[3], chi[3];
    box_reset(lo, hi); box_reset(clo, chi);
    for (uint32_t i = start; i < end; ++i) {
        const prim_box *p = &b->pb[s->prim_order[i]];
        box_grow(lo, hi, p->lo, p->hi); box_grow(clo, chi, p->c, p->c);
    }
    uint32_t n = end - start;
    bvh_node nd; memcpy(nd.lo, lo, sizeof lo); memcpy(nd.hi, hi, sizeof hi);
    nd.left = nd.right = -1; nd.start = start; nd.count = n;
    if (n <= 4 || depth > 60) { s->nodes[ni] = nd; return ni; }
    enum { NB = 16 };
    int best_axis = -1, best_bin = -1; float best_cost = INFINITY;
    for (int a = 0; a < 3; ++a) {
        float ext = chi[a] - clo[a];
        if (!(ext > 0)) continue;
        float blo[NB][3], bhi[NB][3]; uint32_t cnt[NB];
        for (int k = 0; k < NB; ++k) { box_reset(blo[k], bhi[k]); cnt[k] = 0; }
        float scale = NB / ext;
        for (uint32_t i = start; i < end; ++i) {
            const prim_box *p = &b->pb[s->prim_order[i]];
            int k = (int) ((p->c[a] - clo[a]) * scale); if (k >= NB) k = NB - 1; if (k < 0) k = 0;
            box_grow(blo[k], bhi[k], p->lo, p->hi); cnt[k]++;
        }
        float la[NB], ra[NB]; uint32_t lc[NB], rc[NB];
        float tlo[3], thi[3]; uint32_t c = 0; box_reset(tlo, thi);
        for (int k = 0; k < NB; ++k) { if (cnt[k]) box_grow(tlo, thi, blo[k], bhi[k]); c += cnt[k]; la[k] = box_area(tlo, thi); lc[k] = c; }
        box_reset(tlo, thi); c = 0;
        for (int k = NB - 1; k >= 0; --k) { if (cnt[k]) box_grow(tlo, thi, blo[k], bhi[k]); c += cnt[k]; ra[k] = box_area(tlo, thi); rc[k] = c; }
        for (int k = 0; k < NB - 1; ++k) {
            if (lc[k] == 0 || rc[k + 1] == 0) continue;
            float cost = la[k] * lc[k] + ra[k + 1] * rc[k + 1];
            if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k; }
        }
    }
    uint32_t mid;
    if (best_axis < 0) {
        mid = start + n / 2;      /* identical centroids: split by index */
    } else {
        int a = best_axis; float ext = chi[a] - clo[a], scale = NB / ext;
        uint32_t i = start, j = end;
        while (i < j) {
            const prim_box *p = &b->pb[s->prim_order[i]];
            int k = (int) ((p->c[a] - clo[a]) * scale); if (k >= NB) k = NB - 1; if (k < 0) k = 0;
            if (k <= best_bin) ++i; else { --j; uint32_t t = s->prim_order[i]; s->prim_order[i] = s->prim_order[j]; s->prim_order[j] = t; }
        }
        mid = i;
        if (mid == start || mid == end) mid = start + n / 2;
    }
    uint32_t l = bvh_build_rec(b, start, mid, depth + 1);
    uint32_t r = bvh_build_rec(b, mid, end, depth + 1);
    nd.left = (int32_t) l; nd.right = (int32_t) r; nd.count = 0;
    s->nodes[ni] = nd;
    return ni;
}

int orc_scene_build(orc_scene *s) {
    free(s->prim_mesh); free(s->nodes); free(s->prim_order); free(s->emitters);
    s->nodes = NULL; s->nnodes = 0; s->nemitters = 0;
    s->prim_mesh = (uint32_t *) malloc(sizeof(uint32_t) * (s->nprims ? s->nprims : 1));
    s->prim_order = (uint32_t *) malloc(sizeof(uint32_t) * (s->nprims ? s->nprims : 1));
    s->emitters = (int *) malloc(sizeof(int) * (s->nmeshes ? s->nmeshes : 1));
    prim_box *pb = (prim_box *) malloc(sizeof(prim_box) * (s->nprims ? s->nprims : 1));
    float maxabs = 0.0f;
    for (int mi = 0; mi < s->nmeshes; ++mi) {
        mesh_t *m = &s->meshes[mi];
        for (uint32_t f = 0; f < m->nf; ++f) {
            uint32_t g = m->prim_offset + f;
            s->prim_mesh[g] = (uint32_t) mi; s->prim_order[g] = g;
            v3 p[3]; tri_verts(m, f, &p[0], &p[1], &p[2]);
            prim_box *b = &pb[g]; box_reset(b->lo, b->hi);
            for (int k = 0; k < 3; ++k) { float q[3] = { p[k].x, p[k].y, p[k].z }; box_grow(b->lo, b->hi, q, q);
                for (int a = 0; a < 3; ++a) if (fabsf(q[a]) > maxabs) maxabs = fabsf(q[a]); }
            for (int a = 0; a < 3; ++a) b->c[a] = 0.5f * (b->lo[a] + b->hi[a]);
        }
        /* emitter CDF over triangle areas: ref src/mesh.cpp:31-37, include/nori/dpdf.h:40-84 */
        free(m->cdf); m->cdf = NULL;
        if (m->emitter.type == ORC_EMITTER_AREA) {
            s->emitters[s->nemitters++] = mi;
            m->cdf = (float *) malloc(sizeof(float) * (m->nf + 1));
            m->cdf[0] = 0.0f;
            for (uint32_t f = 0; f < m->nf; ++f) {
                v3 p0, p1, p2; tri_verts(m, f, &p0, &p1, &p2);
                v3 c = cross3(v3sub(p1, p0), v3sub(p2, p0));
                float area = 0.5f * sqrtf(dot3(c, c));
                m->cdf[f + 1] = m->cdf[f] + area;
            }
            m->area_sum = m->cdf[m->nf];
            if (m->area_sum > 0) {
                float norm = 1.0f / m->area_sum;
                for (uint32_t f = 1; f <= m->nf; ++f) m->cdf[f] *= norm;
                m->cdf[m->nf] = 1.0f;
            }
        }
    }
    float pad = 4e-6f * maxabs;
    for (uint32_t g = 0; g < s->nprims; ++g) for (int a = 0; a < 3; ++a) { pb[g].lo[a] -= pad; pb[g].hi[a] += pad; }
    if (s->nprims) { bvh_build b = { s, pb, 0 }; bvh_build_rec(&b, 0, s->nprims, 0); }
    free(pb);
    return 0;
}

/* ------------------------------------------------------------------ ray / triangle: ref src/mesh.cpp:39-76 */
typedef struct { v3 o, d, dRcp; float mint, maxt; } ray_t;   /* ref: include/nori/ray.h:30-34 */
static inline void ray_update(ray_t *r) { r->dRcp = v3make(1.0f / r->d.x, 1.0f / r->d.y, 1.0f / r->d.z); } /* ref: ray.h:62-64 */

static inline int tri_intersect(v3 p0, v3 p1, v3 p2, const ray_t *ray, float *u, float *v, float *t) {
    v3 edge1 = v3sub(p1, p0), edge2 = v3sub(p2, p0);
    v3 pvec = cross3(ray->d, edge2);
    float det = dot3(edge1, pvec);
    if (det > -1e-8f && det < 1e-8f) return 0;
    float inv_det = 1.0f / det;
    v3 tvec = v3sub(ray->o, p0);
    *u = dot3(tvec, pvec) * inv_det;
    if (*u < 0.0 || *u > 1.0) return 0;
    v3 qvec = cross3(tvec, edge1);
    *v = dot3(ray->d, qvec) * inv_det;
    if (*v < 0.0 || *u + *v > 1.0) return 0;
    *t = dot3(edge2, qvec) * inv_det;
    return *t >= ray->mint && *t <= ray->maxt;
}

typedef struct { float t, u, v; uint32_t prim; } hit_t;
typedef struct { uint64_t rays, nodes, tris; } counters;

/* Closest/any hit, brute force: ref src/accel.cpp:30-43 (extended to several meshes in add order). */
static int search_brute(const orc_scene *s, ray_t ray, int shadow, hit_t *h, counters *c) {
    int found = 0; h->prim = ORC_MISS;
    for (int mi = 0; mi < s->nmeshes; ++mi) {
        const mesh_t *m = &s->meshes[mi];
        for (uint32_t f = 0; f < m->nf; ++f) {
            v3 p0, p1, p2; float u, v, t;
            tri_verts(m, f, &p0, &p1, &p2);
            if (tri_intersect(p0, p1, p2, &ray, &u, &v, &t)) {
                if (shadow) { c->tris += f + 1; return 1; }
                ray.maxt = h->t = t; h->u = u; h->v = v; h->prim = m->prim_offset + f; found = 1;
            }
        }
        c->tris += m->nf;
    }
    return found;
}

/* slab test: ref include/nori/bbox.h:353-380 */
static inline int box_intersect(const float lo[3], const float hi[3], const ray_t *ray, float *nearT_) {
    float nearT = -INFINITY, farT = INFINITY;
    const float o[3] = { ray->o.x, ray->o.y, ray->o.z }, d[3] = { ray->d.x, ray->d.y, ray->d.z },
                r[3] = { ray->dRcp.x, ray->dRcp.y, ray->dRcp.z };
    for (int i = 0; i < 3; ++i) {
        if (d[i] == 0) {
            if (o[i] < lo[i] || o[i] > hi[i]) return 0;
        } else {
            float t1 = (lo[i] - o[i]) * r[i], t2 = (hi[i] - o[i]) * r[i];
            if (t1 > t2) { float tmp = t1; t1 = t2; t2 = tmp; }
            if (t1 > nearT) nearT = t1;
            if (t2 < farT) farT = t2;
            if (!(nearT <= farT)) return 0;
        }
    }
    *nearT_ = nearT;
    return ray->mint <= farT && nearT <= ray->maxt;
}

/* BVH search.  Result is order independent: among equal minimal t the HIGHEST global triangle
 * index wins, which is what the ascending brute-force loop with "t <= maxt" yields
 * (ref: src/mesh.cpp:75, src/accel.cpp:37). */
static int search_bvh(const orc_scene *s, ray_t ray, int shadow, hit_t *h, counters *c) {
    int found = 0; h->prim = ORC_MISS;
    if (!s->nnodes) return 0;
    uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const bvh_node *nd = &s->nodes[stack[--sp]];
        float nt;
        c->nodes++;
        if (!box_intersect(nd->lo, nd->hi, &ray, &nt)) continue;
        if (nd->left < 0) {
            for (uint32_t i = 0; i < nd->count; ++i) {
                uint32_t g = s->prim_order[nd->start + i];
                const mesh_t *m = &s->meshes[s->prim_mesh[g]];
                v3 p0, p1, p2; float u, v, t;
                tri_verts(m, g - m->prim_offset, &p0, &p1, &p2);
                c->tris++;
                if (tri_intersect(p0, p1, p2, &ray, &u, &v, &t)) {
                    if (shadow) return 1;
                    if (!found || t < h->t || g > h->prim) { ray.maxt = h->t = t; h->u = u; h->v = v; h->prim = g; }
                    found = 1;
                }
            }
        } else {
            const bvh_node *l = &s->nodes[nd->left], *r = &s->nodes[nd->right];
            /* visit the child whose box centre is nearer along the ray first */
            float cl = 0, cr = 0;
            const float d[3] = { ray.d.x, ray.d.y, ray.d.z };
            for (int a = 0; a < 3; ++a) { cl += (l->lo[a] + l->hi[a]) * d[a]; cr += (r->lo[a] + r->hi[a]) * d[a]; }
            if (cl < cr) { stack[sp++] = (uint32_t) nd->right; stack[sp++] = (uint32_t) nd->left; }
            else { stack[sp++] = (uint32_t) nd->left; stack[sp++] = (uint32_t) nd->right; }
        }
    }
    return found;
}

static inline int scene_search(const orc_scene *s, const ray_t *ray, int shadow, int accel, hit_t *h, counters *c) {
    c->rays++;
    return accel == ORC_ACCEL_BRUTE ? search_brute(s, *ray, shadow, h, c) : search_bvh(s, *ray, shadow, h, c);
}

/* Intersection record fill: ref src/accel.cpp:45-96, include/nori/mesh.h:23-35 */
typedef struct { v3 p; float t; float uvx, uvy; frame sh, geo; int mesh; } its_t;

static void fill_its(const orc_scene *s, const hit_t *h, its_t *its) {
    int mi = (int) s->prim_mesh[h->prim];
    const mesh_t *m = &s->meshes[mi];
    uint32_t f = h->prim - m->prim_offset;
    const uint32_t *idx = m->F + 3 * (size_t) f;
    float b0 = 1 - (h->u + h->v), b1 = h->u, b2 = h->v;
    v3 p0 = ld3(m->V + 3 * (size_t) idx[0]), p1 = ld3(m->V + 3 * (size_t) idx[1]), p2 = ld3(m->V + 3 * (size_t) idx[2]);
    its->t = h->t; its->mesh = mi;
    its->p = v3make(b0 * p0.x + b1 * p1.x + b2 * p2.x, b0 * p0.y + b1 * p1.y + b2 * p2.y, b0 * p0.z + b1 * p1.z + b2 * p2.z);
    its->uvx = h->u; its->uvy = h->v;
    if (m->UV) {
        const float *t0 = m->UV + 2 * (size_t) idx[0], *t1 = m->UV + 2 * (size_t) idx[1], *t2 = m->UV + 2 * (size_t) idx[2];
        its->uvx = b0 * t0[0] + b1 * t1[0] + b2 * t2[0];
        its->uvy = b0 * t0[1] + b1 * t1[1] + b2 * t2[1];
    }
    its->geo = frame_from_n(normalize3(cross3(v3sub(p1, p0), v3sub(p2, p0))));
    if (m->N) {
        v3 n0 = ld3(m->N + 3 * (size_t) idx[0]), n1 = ld3(m->N + 3 * (size_t) idx[1]), n2 = ld3(m->N + 3 * (size_t) idx[2]);
        v3 n = v3make(b0 * n0.x + b1 * n1.x + b2 * n2.x, b0 * n0.y + b1 * n1.y + b2 * n2.y, b0 * n0.z + b1 * n1.z + b2 * n2.z);
        its->sh = frame_from_n(normalize3(n));
    } else {
        its->sh = its->geo;
    }
}

/* ------------------------------------------------------------------ camera: ref src/perspective.cpp:76-97,
 * include/nori/transform.h:55-68 */
static void sample_ray(const orc_scene *s, float sx, float sy, ray_t *ray) {
    const float *m = s->s2c, *c = s->c2w;
    float px = sx * s->invW, py = sy * s->invH, pz = 0.0f;
    float r[4];
    for (int i = 0; i < 4; ++i) r[i] = ((m[4 * i + 0] * px + m[4 * i + 1] * py) + m[4 * i + 2] * pz) + m[4 * i + 3] * 1.0f;
    v3 nearP = v3make(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
    v3 d = normalize3(nearP);
    float invZ = 1.0f / d.z;
    float w = ((c[12] * 0.0f + c[13] * 0.0f) + c[14] * 0.0f) + c[15] * 1.0f;
    ray->o = v3make((((c[0] * 0.0f + c[1] * 0.0f) + c[2] * 0.0f) + c[3] * 1.0f) / w,
                    (((c[4] * 0.0f + c[5] * 0.0f) + c[6] * 0.0f) + c[7] * 1.0f) / w,
                    (((c[8] * 0.0f + c[9] * 0.0f) + c[10] * 0.0f) + c[11] * 1.0f) / w);
    ray->d = v3make(c[0] * d.x + (c[1] * d.y + c[2] * d.z), c[4] * d.x + (c[5] * d.y + c[6] * d.z),
                    c[8] * d.x + (c[9] * d.y + c[10] * d.z));
    ray->mint = s->nearClip * invZ;
    ray->maxt = s->farClip * invZ;
    ray_update(ray);
}
void orc_sample_ray(const orc_scene *s, float sx, float sy, orc_ray *out) {
    ray_t r; sample_ray(s, sx, sy, &r);
    out->o[0] = r.o.x; out->o[1] = r.o.y; out->o[2] = r.o.z; out->d[0] = r.d.x; out->d[1] = r.d.y; out->d[2] = r.d.z;
    out->mint = r.mint; out->maxt = r.maxt;
}

/* ref: src/perspective.cpp:41-68 -- sampleToCamera = inverse(scale * translate * perspective).
 * Composed and inverted in double, rounded once to fp32 (Eigen's fp32 inverse is not
 * reproducible here; both sides of every parity test receive THESE numbers). */
void orc_camera_matrices(float fov, float nearClip, float farClip, int W, int H, float s2c[16]) {
    double aspect = (double) ((float) W / (float) H);
    double recip = 1.0 / ((double) farClip - (double) nearClip);
    double cot = 1.0 / tan((double) fov / 2.0 * (3.14159265358979323846 / 180.0));
    /* forward: x' = -0.5*(cot x/z - 1), y' = -0.5*aspect*(cot y/z - 1/aspect), z' = far(z-near)/(z(far-near)).
       inverse of the homogeneous matrix M = S*T*P computed analytically. */
    double P[16] = { cot, 0, 0, 0,  0, cot, 0, 0,  0, 0, farClip * recip, -(double) nearClip * farClip * recip,  0, 0, 1, 0 };
    double T[16] = { 1, 0, 0, -1,  0, 1, 0, -1.0 / aspect,  0, 0, 1, 0,  0, 0, 0, 1 };
    double S[16] = { -0.5, 0, 0, 0,  0, -0.5 * aspect, 0, 0,  0, 0, 1, 0,  0, 0, 0, 1 };
    double TP[16], M[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += T[4 * i + k] * P[4 * k + j]; TP[4 * i + j] = a; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += S[4 * i + k] * TP[4 * k + j]; M[4 * i + j] = a; }
    /* Gauss-Jordan inverse in double */
    double A[4][8];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { A[i][j] = M[4 * i + j]; A[i][4 + j] = (i == j); }
    for (int col = 0; col < 4; ++col) {
        int piv = col; for (int r = col + 1; r < 4; ++r) if (fabs(A[r][col]) > fabs(A[piv][col])) piv = r;
        if (piv != col) for (int j = 0; j < 8; ++j) { double t = A[col][j]; A[col][j] = A[piv][j]; A[piv][j] = t; }
        double d = A[col][col];
        for (int j = 0; j < 8; ++j) A[col][j] /= d;
        for (int r = 0; r < 4; ++r) if (r != col) { double f = A[r][col]; for (int j = 0; j < 8; ++j) A[r][j] -= f * A[col][j]; }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s2c[4 * i + j] = (float) A[i][4 + j];
}

/* ------------------------------------------------------------------ film: ref src/block.cpp:15-102 */
typedef struct { int ox, oy, sx, sy, border; int rows, cols; float *px; /* rows*cols*4 */ } block_t;

static int color_valid(v3 c) {   /* ref: src/common.cpp:196-203 */
    return !(c.x < 0 || !isfinite(c.x) || c.y < 0 || !isfinite(c.y) || c.z < 0 || !isfinite(c.z));
}

static void block_put(const orc_scene *s, block_t *b, float sx, float sy, v3 value) {
    if (!color_valid(value)) return;   /* ref: src/block.cpp:63-67 (warning text omitted) */
    float posx = sx - 0.5f - (float) (b->ox - b->border), posy = sy - 0.5f - (float) (b->oy - b->border);
    int x0 = (int) ceilf(posx - s->fradius), y0 = (int) ceilf(posy - s->fradius);
    int x1 = (int) floorf(posx + s->fradius), y1 = (int) floorf(posy + s->fradius);
    if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0;
    if (x1 > b->cols - 1) x1 = b->cols - 1; if (y1 > b->rows - 1) y1 = b->rows - 1;
    float wx[16], wy[16];
    for (int x = x0, i = 0; x <= x1; ++x) wx[i++] = s->ftable[(int) (fabsf((float) x - posx) * s->lookup)];
    for (int y = y0, i = 0; y <= y1; ++y) wy[i++] = s->ftable[(int) (fabsf((float) y - posy) * s->lookup)];
    for (int y = y0, yr = 0; y <= y1; ++y, ++yr)
        for (int x = x0, xr = 0; x <= x1; ++x, ++xr) {
            float *p = b->px + 4 * ((size_t) y * b->cols + x);
            /* Color4f(value) * wX * wY, then += : ref src/block.cpp:90 */
            p[0] += value.x * wx[xr] * wy[yr]; p[1] += value.y * wx[xr] * wy[yr];
            p[2] += value.z * wx[xr] * wy[yr]; p[3] += 1.0f * wx[xr] * wy[yr];
        }
}

/* BlockGenerator spiral: ref src/block.cpp:109-152 */
int orc_block_order(int W, int H, int bs, int32_t *xy) {
    int nbx = (int) ceilf(W / (float) bs), nby = (int) ceilf(H / (float) bs);
    int left = nbx * nby, n = 0;
    int bx = nbx / 2, by = nby / 2, dir = 0 /*ERight*/, stepsLeft = 1, numSteps = 1;
    while (left > 0) {
        if (xy) { xy[2 * n] = bx; xy[2 * n + 1] = by; }
        ++n;
        if (--left == 0) break;
        do {
            switch (dir) { case 0: ++bx; break; case 1: ++by; break; case 2: --bx; break; default: --by; break; }
            if (--stepsLeft == 0) {
                dir = (dir + 1) % 4;
                if (dir == 2 || dir == 0) ++numSteps;
                stepsLeft = numSteps;
            }
        } while (bx < 0 || by < 0 || bx >= nbx || by >= nby);
    }
    return n;
}

int orc_film_to_rgb(const float *film, int W, int H, int border, float *rgb) {   /* ref: src/block.cpp:45-51, include/nori/color.h:100-105 */
    int cols = W + 2 * border;
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const float *p = film + 4 * ((size_t) (y + border) * cols + (x + border));
        float *o = rgb + 3 * ((size_t) y * W + x);
        if (p[3] != 0) { o[0] = p[0] / p[3]; o[1] = p[1] / p[3]; o[2] = p[2] / p[3]; } else { o[0] = o[1] = o[2] = 0.0f; }
    }
    return 0;
}

/* ------------------------------------------------------------------ emitter sampling [authored]
 * interface ref: include/nori/emitter.h:16-24 (empty); DiscretePDF ref: include/nori/dpdf.h:93-99 */
static uint32_t cdf_sample(const float *cdf, uint32_t n, float x) {
    /* std::lower_bound over cdf[0..n] then index = max(0, pos-1), min(index, n-1) */
    uint32_t lo = 0, hi = n + 1;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (cdf[mid] < x) lo = mid + 1; else hi = mid; }
    int64_t idx = (int64_t) lo - 1; if (idx < 0) idx = 0;
    if (idx > (int64_t) n - 1) idx = (int64_t) n - 1;
    return (uint32_t) idx;
}

typedef struct { v3 y, n; float pdfA; v3 Le; } emit_sample;

static void sample_emitter(const orc_scene *s, float xe, float xt, float xa, float xb, emit_sample *es) {
    int k = (int) (xe * (float) s->nemitters); if (k > s->nemitters - 1) k = s->nemitters - 1;
    const mesh_t *m = &s->meshes[s->emitters[k]];
    uint32_t f = cdf_sample(m->cdf, m->nf, xt);
    const uint32_t *idx = m->F + 3 * (size_t) f;
    v3 p0 = ld3(m->V + 3 * (size_t) idx[0]), p1 = ld3(m->V + 3 * (size_t) idx[1]), p2 = ld3(m->V + 3 * (size_t) idx[2]);
    float su = sqrtf(1.0f - xa);
    float b0 = 1.0f - su, b1 = xb * su; float b2 = 1.0f - b0 - b1;
    es->y = v3make(b0 * p0.x + b1 * p1.x + b2 * p2.x, b0 * p0.y + b1 * p1.y + b2 * p2.y, b0 * p0.z + b1 * p1.z + b2 * p2.z);
    if (m->N) {
        v3 n0 = ld3(m->N + 3 * (size_t) idx[0]), n1 = ld3(m->N + 3 * (size_t) idx[1]), n2 = ld3(m->N + 3 * (size_t) idx[2]);
        es->n = normalize3(v3make(b0 * n0.x + b1 * n1.x + b2 * n2.x, b0 * n0.y + b1 * n1.y + b2 * n2.y, b0 * n0.z + b1 * n1.z + b2 * n2.z));
    } else {
        es->n = normalize3(cross3(v3sub(p1, p0), v3sub(p2, p0)));
    }
    es->pdfA = 1.0f / (m->area_sum * (float) s->nemitters);
    es->Le = v3make(m->emitter.radiance[0], m->emitter.radiance[1], m->emitter.radiance[2]);
}

/* ------------------------------------------------------------------ Integrator::Li [authored]
 * interface ref: include/nori/integrator.h:42.  One iterative state machine for all six
 * integrators; the order of sampler draws is part of the spec. */
static inline v3 v3mul(v3 a, v3 b) { return v3make(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float max3(v3 a) { float m = a.x > a.y ? a.x : a.y; return m > a.z ? m : a.z; }
static inline int v3iszero(v3 a) { return a.x == 0 && a.y == 0 && a.z == 0; }

static v3 Li(const orc_scene *s, orc_pcg32 *rng, ray_t ray, int accel, counters *c) {
    const int type = s->integ.type;
    const int max_depth = s->integ.max_depth > 0 ? s->integ.max_depth : (1 << 20);
    v3 L = { 0, 0, 0 }, T = { 1, 1, 1 };
    int prev_specular = 1; float prev_pdf = 0.0f;
    for (int depth = 0; depth < max_depth; ++depth) {
        hit_t h;
        if (!scene_search(s, &ray, 0, accel, &h, c)) break;
        its_t its; fill_its(s, &h, &its);
        const mesh_t *m = &s->meshes[its.mesh];

        if (type == ORC_INT_NORMALS) {
            L = v3make(fabsf(its.sh.n.x), fabsf(its.sh.n.y), fabsf(its.sh.n.z));
            break;
        }
        if (type == ORC_INT_AO) {
            float x = orc_pcg32_next_float(rng), y = orc_pcg32_next_float(rng);
            v3 w = to_world(&its.sh, sq2coshemi(x, y));
            ray_t sr; sr.o = its.p; sr.d = w; sr.mint = ORC_EPSILON; sr.maxt = INFINITY; ray_update(&sr);
            hit_t sh;
            if (!scene_search(s, &sr, 1, accel, &sh, c)) L = v3make(1, 1, 1);
            break;
        }

        if (type == ORC_INT_SIMPLE) {
            /* [authored] the course's point-light integrator named by scenes/pa3/ajax-simple.xml:8-11 (no source in the
             * reference): Li = Phi / (4 pi^2) * max(0, cos theta) / |x - p|^2 * V(x <-> p); theta is measured against
             * the shading normal; no sampler draws beyond the camera sample. */
            v3 dvec = v3sub(ld3(s->light_pos), its.p);
            float dist2 = dot3(dvec, dvec);
            float dist = sqrtf(dist2);
            v3 wo_w = v3make(dvec.x / dist, dvec.y / dist, dvec.z / dist);
            float cosT = dot3(its.sh.n, wo_w);
            if (cosT > 0.0f) {
                float g = cosT / dist2 * ORC_INV_FOURPI2;
                ray_t sr; sr.o = its.p; sr.d = wo_w; sr.mint = ORC_EPSILON; sr.maxt = dist - ORC_EPSILON; ray_update(&sr);
                hit_t sh;
                if (!scene_search(s, &sr, 1, accel, &sh, c))
                    L = v3make(s->light_energy[0] * g, s->light_energy[1] * g, s->light_energy[2] * g);
            }
            break;
        }

        v3 wi = to_local(&its.sh, v3neg(ray.d));
        const int diffuse = bsdf_is_diffuse(&m->bsdf);

        /* emitted radiance */
        if (m->emitter.type == ORC_EMITTER_AREA && wi.z > 0.0f) {
            v3 Le = v3make(m->emitter.radiance[0], m->emitter.radiance[1], m->emitter.radiance[2]);
            float w = 1.0f;
            int add = 0;
            if (type == ORC_INT_WHITTED) add = diffuse;
            else if (type == ORC_INT_PATH_MATS) add = 1;
            else if (type == ORC_INT_PATH_EMS) add = prev_specular;
            else { /* MIS */
                add = 1;
                if (!prev_specular) {
                    float pdfA = 1.0f / (m->area_sum * (float) s->nemitters);
                    float pdf_em = pdfA * (its.t * its.t) / wi.z;
                    w = prev_pdf / (prev_pdf + pdf_em);
                }
            }
            if (add) { L.x += T.x * Le.x * w; L.y += T.y * Le.y * w; L.z += T.z * Le.z * w; }
        }

        if (type == ORC_INT_WHITTED && !diffuse) {
            float x = orc_pcg32_next_float(rng);
            if (x >= 0.95f) break;
            float sx = orc_pcg32_next_float(rng), sy = orc_pcg32_next_float(rng);
            v3 wo; float eta; int measure;
            v3 f = bsdf_sample(&m->bsdf, wi, sx, sy, &wo, &eta, &measure);
            if (v3iszero(f)) break;
            T = v3make(T.x * f.x / 0.95f, T.y * f.y / 0.95f, T.z * f.z / 0.95f);
            ray.o = its.p; ray.d = to_world(&its.sh, wo); ray.mint = ORC_EPSILON; ray.maxt = INFINITY; ray_update(&ray);
            continue;
        }

        if (type != ORC_INT_WHITTED && depth >= s->integ.rr_start) {
            float q = max3(T); if (q > 0.99f) q = 0.99f;
            float x = orc_pcg32_next_float(rng);
            if (x >= q) break;
            T = v3make(T.x / q, T.y / q, T.z / q);
        }

        /* next-event estimation */
        if ((type == ORC_INT_WHITTED || type == ORC_INT_PATH_EMS || type == ORC_INT_PATH_MIS) && diffuse && s->nemitters > 0) {
            float xe = orc_pcg32_next_float(rng), xt = orc_pcg32_next_float(rng);
            float xa = orc_pcg32_next_float(rng), xb = orc_pcg32_next_float(rng);
            emit_sample es; sample_emitter(s, xe, xt, xa, xb, &es);
            v3 dvec = v3sub(es.y, its.p);
            float dist2 = dot3(dvec, dvec);
            float dist = sqrtf(dist2);
            v3 wo_w = v3make(dvec.x / dist, dvec.y / dist, dvec.z / dist);
            float cosL = -dot3(es.n, wo_w);
            if (cosL > 0.0f) {
                v3 wo = to_local(&its.sh, wo_w);
                v3 f = bsdf_eval(&m->bsdf, wi, wo);
                if (!v3iszero(f)) {
                    float pdf_sa = es.pdfA * dist2 / cosL;
                    float w = 1.0f;
                    if (type == ORC_INT_PATH_MIS) w = pdf_sa / (pdf_sa + bsdf_pdf(&m->bsdf, wi, wo));
                    float g = wo.z / pdf_sa * w;
                    v3 contrib = v3make(T.x * f.x * es.Le.x * g, T.y * f.y * es.Le.y * g, T.z * f.z * es.Le.z * g);
                    ray_t sr; sr.o = its.p; sr.d = wo_w; sr.mint = ORC_EPSILON; sr.maxt = dist - ORC_EPSILON; ray_update(&sr);
                    hit_t sh;
                    if (!scene_search(s, &sr, 1, accel, &sh, c)) L = v3add(L, contrib);
                }
            }
        }
        if (type == ORC_INT_WHITTED) break;

        /* BSDF sampling */
        {
            float sx = orc_pcg32_next_float(rng), sy = orc_pcg32_next_float(rng);
            v3 wo; float eta; int measure;
            v3 f = bsdf_sample(&m->bsdf, wi, sx, sy, &wo, &eta, &measure);
            if (v3iszero(f)) break;
            T = v3mul(T, f);
            prev_specular = (measure == 2);
            prev_pdf = prev_specular ? 0.0f : bsdf_pdf(&m->bsdf, wi, wo);
            ray.o = its.p; ray.d = to_world(&its.sh, wo); ray.mint = ORC_EPSILON; ray.maxt = INFINITY; ray_update(&ray);
        }
    }
    return L;
}

/* ------------------------------------------------------------------ render: ref src/main.cpp:27-56 (renderBlock), 58-119 (tile loop) */
static void render_block(const orc_scene *s, block_t *b, int accel, counters *c) {
    memset(b->px, 0, sizeof(float) * 4 * (size_t) b->rows * b->cols);
    orc_pcg32 rng;
    if (s->seed_mode == ORC_SEED_PER_BLOCK) orc_pcg32_seed(&rng, (uint64_t) b->ox, (uint64_t) b->oy);   /* ref: src/independent.cpp:36-41 */
    for (int y = 0; y < b->sy; ++y) for (int x = 0; x < b->sx; ++x) for (uint32_t i = 0; i < s->spp; ++i) {
        if (s->seed_mode == ORC_SEED_PER_SAMPLE) {
            uint64_t pix = (uint64_t) (y + b->oy) * (uint64_t) s->W + (uint64_t) (x + b->ox);
            orc_pcg32_seed(&rng, (s->seed << 32) + pix, (uint64_t) i);
        }
        float sx = (float) (x + b->ox) + orc_pcg32_next_float(&rng);
        float sy = (float) (y + b->oy) + orc_pcg32_next_float(&rng);
        orc_pcg32_next_float(&rng); orc_pcg32_next_float(&rng);     /* apertureSample: ref src/main.cpp:42 */
        ray_t ray; sample_ray(s, sx, sy, &ray);
        v3 value = Li(s, &rng, ray, accel, c);                      /* camera weight is 1: ref src/perspective.cpp:96 */
        block_put(s, b, sx, sy, value);
    }
}

typedef struct {
    orc_scene *s; int accel; int ntiles; const int32_t *order; block_t *blocks;
    volatile int next; counters c; pthread_mutex_t *mu;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *) arg;
    counters c = { 0, 0, 0 };
    for (;;) {
        int i = __sync_fetch_and_add(&j->next, 1);
        if (i >= j->ntiles) break;
        block_t *b = &j->blocks[i];
        if (b->px) render_block(j->s, b, j->accel, &c);
    }
    pthread_mutex_lock(j->mu);
    j->c.rays += c.rays; j->c.nodes += c.nodes; j->c.tris += c.tris;
    pthread_mutex_unlock(j->mu);
    return NULL;
}

int orc_render(orc_scene *s, float *film, int accel, int nthreads, orc_stats *st) {
    const int W = s->W, H = s->H, bd = s->border;
    const int cols = W + 2 * bd, rows = H + 2 * bd;
    int nbx = (W + ORC_BLOCK - 1) / ORC_BLOCK, nby = (H + ORC_BLOCK - 1) / ORC_BLOCK, ntiles = nbx * nby;
    int32_t *order = (int32_t *) malloc(sizeof(int32_t) * 2 * ntiles);
    orc_block_order(W, H, ORC_BLOCK, order);
    block_t *blocks = (block_t *) calloc(ntiles, sizeof(block_t));
    /* Tile shards (multi-GPU slices): tile t is owned by rank t % nranks.  Numbering of the device path (nb_api.cu:
     * build_tile_order): ownership follows the Latin pattern (bx + shift by) % nranks (shift = 3; 5 or 7 when 3 divides nranks), rank r's k-th tile (row by row) is tile
     * k * nranks + r; the few tiles by which the pattern misses the implied counts move from the ranks with a surplus (their last
     * tiles) to those with a deficit.  zrank[by * nbx + bx] = t */
    int32_t *zrank = (int32_t *) malloc(sizeof(int32_t) * ntiles);
    {
        const int N = s->tile_nranks;
        int32_t *lists = (int32_t *) malloc(sizeof(int32_t) * (size_t) N * ntiles);     /* lists[r * ntiles + k] = cell */
        int32_t *len = (int32_t *) calloc((size_t) N, sizeof(int32_t));
        int32_t *pool = (int32_t *) malloc(sizeof(int32_t) * ntiles);
        int npool = 0, head = 0;
        const int shift = N % 3 ? 3 : N % 5 ? 5 : 7;
        for (int y = 0; y < nby; ++y) for (int x = 0; x < nbx; ++x) { int r = (x + shift * y) % N; lists[(size_t) r * ntiles + len[r]++] = y * nbx + x; }
        for (int r = 0; r < N; ++r) { int target = ntiles > r ? (ntiles - r + N - 1) / N : 0; while (len[r] > target) pool[npool++] = lists[(size_t) r * ntiles + --len[r]]; }
        for (int r = 0; r < N; ++r) { int target = ntiles > r ? (ntiles - r + N - 1) / N : 0; while (len[r] < target) lists[(size_t) r * ntiles + len[r]++] = pool[head++]; }
        for (int r = 0; r < N; ++r) for (int k = 0; k < len[r]; ++k) zrank[lists[(size_t) r * ntiles + k]] = k * N + r;
        free(lists); free(len); free(pool);
    }
    uint64_t nsamples = 0;
    for (int i = 0; i < ntiles; ++i) {
        block_t *b = &blocks[i];
        int bx = order[2 * i], by = order[2 * i + 1];
        b->ox = bx * ORC_BLOCK; b->oy = by * ORC_BLOCK;
        b->sx = W - b->ox < ORC_BLOCK ? W - b->ox : ORC_BLOCK;   /* ref: src/block.cpp:129 */
        b->sy = H - b->oy < ORC_BLOCK ? H - b->oy : ORC_BLOCK;
        b->border = bd; b->cols = b->sx + 2 * bd; b->rows = b->sy + 2 * bd;
        int tile_id = zrank[by * nbx + bx];   /* tile numbering of the device path */
        if (tile_id % s->tile_nranks == s->tile_rank) {
            b->px = (float *) malloc(sizeof(float) * 4 * (size_t) b->rows * b->cols);
            nsamples += (uint64_t) b->sx * b->sy * s->spp;
        }
    }
    free(zrank);
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    job_t job; memset(&job, 0, sizeof job);
    job.s = s; job.accel = accel; job.ntiles = ntiles; job.order = order; job.blocks = blocks; job.mu = &mu;
    if (nthreads < 1) nthreads = 1;
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t) * nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, worker, &job);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    /* merge in BlockGenerator order: ref src/block.cpp:93-102 (deterministic regardless of thread schedule) */
    memset(film, 0, sizeof(float) * 4 * (size_t) rows * cols);
    for (int i = 0; i < ntiles; ++i) {
        block_t *b = &blocks[i];
        if (!b->px) continue;
        for (int y = 0; y < b->rows; ++y) {
            float *dst = film + 4 * ((size_t) (b->oy + y) * cols + b->ox);
            const float *src = b->px + 4 * (size_t) y * b->cols;
            for (int x = 0; x < 4 * b->cols; ++x) dst[x] += src[x];
        }
        free(b->px);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (st) {
        st->samples = nsamples; st->rays = job.c.rays; st->node_visits = job.c.nodes; st->tri_tests = job.c.tris;
        st->seconds = (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
    }
    free(th); free(blocks); free(order);
    return 0;
}

static ray_t ray_from(const orc_ray *r) {
    ray_t x; x.o = ld3(r->o); x.d = ld3(r->d); x.mint = r->mint; x.maxt = r->maxt; ray_update(&x); return x;
}

int orc_intersect(orc_scene *s, const orc_ray *rays, uint64_t n, orc_hit *hits, int shadow, int accel, orc_stats *st) {
    counters c = { 0, 0, 0 };
    for (uint64_t i = 0; i < n; ++i) {
        ray_t r = ray_from(&rays[i]); hit_t h; h.t = 0; h.u = 0; h.v = 0; h.prim = ORC_MISS;
        int found = scene_search(s, &r, shadow, accel, &h, &c);
        orc_hit *o = &hits[i];
        if (found && !shadow) { o->t = h.t; o->u = h.u; o->v = h.v; o->prim = h.prim; o->mesh = s->prim_mesh[h.prim]; }
        else { o->t = 0; o->u = 0; o->v = 0; o->prim = found ? 0u : ORC_MISS; o->mesh = found ? 0u : ORC_MISS; }
    }
    if (st) { st->rays = c.rays; st->node_visits = c.nodes; st->tri_tests = c.tris; st->samples = 0; st->seconds = 0; }
    return 0;
}

int orc_intersect_full(orc_scene *s, const orc_ray *rays, uint64_t n, float *out, int accel) {
    counters c = { 0, 0, 0 };
    for (uint64_t i = 0; i < n; ++i) {
        ray_t r = ray_from(&rays[i]); hit_t h; float *o = out + 16 * i;
        if (!scene_search(s, &r, 0, accel, &h, &c)) { for (int k = 0; k < 16; ++k) o[k] = 0; o[15] = -1.0f; continue; }
        its_t its; fill_its(s, &h, &its);
        o[0] = its.p.x; o[1] = its.p.y; o[2] = its.p.z; o[3] = its.t; o[4] = its.uvx; o[5] = its.uvy;
        o[6] = its.sh.s.x; o[7] = its.sh.s.y; o[8] = its.sh.s.z; o[9] = its.sh.t.x; o[10] = its.sh.t.y; o[11] = its.sh.t.z;
        o[12] = its.sh.n.x; o[13] = its.sh.n.y; o[14] = its.sh.n.z; o[15] = (float) its.mesh;
    }
    return 0;
}

/* Scene-mode t-test sampling: ref src/ttest.cpp:139-167.  One default-seeded sequential stream. */
int orc_ttest_scene(orc_scene *s, uint64_t n, int accel, double *lum) {
    orc_pcg32 rng; orc_pcg32_init(&rng);
    counters c = { 0, 0, 0 };
    for (uint64_t k = 0; k < n; ++k) {
        float sx = orc_pcg32_next_float(&rng) * (float) s->W, sy = orc_pcg32_next_float(&rng) * (float) s->H;
        orc_pcg32_next_float(&rng); orc_pcg32_next_float(&rng);
        ray_t ray; sample_ray(s, sx, sy, &ray);
        v3 v = Li(s, &rng, ray, accel, &c);
        lum[k] = (double) (v.x * 0.212671f + v.y * 0.715160f + v.z * 0.072169f);   /* ref: src/common.cpp:206-208 */
    }
    return 0;
}

/* The same loop with one pcg32 stream PER PATH (path k: seed((seed << 32) + k, 0)) -- the parallel form that
 * nb_li_samples evaluates on the device; lum is fp32 and must match it bit for bit. */
int orc_li_samples(orc_scene *s, uint64_t n, int accel, float *lum) {
    counters c = { 0, 0, 0 };
    for (uint64_t k = 0; k < n; ++k) {
        orc_pcg32 rng; orc_pcg32_seed(&rng, (s->seed << 32) + k, 0);
        float sx = orc_pcg32_next_float(&rng) * (float) s->W, sy = orc_pcg32_next_float(&rng) * (float) s->H;
        orc_pcg32_next_float(&rng); orc_pcg32_next_float(&rng);
        ray_t ray; sample_ray(s, sx, sy, &ray);
        v3 v = Li(s, &rng, ray, accel, &c);
        lum[k] = v.x * 0.212671f + v.y * 0.715160f + v.z * 0.072169f;
    }
    return 0;
}

/* BSDF-mode t-test sampling: ref src/ttest.cpp:104-125 (wi = sphericalDirection(angle, 0), ref src/common.cpp:224-236).
 * The caller owns the rng so that consecutive angles continue ONE stream, as the reference does. */
int orc_bsdf_sample_batch(const orc_bsdf *b, const float wi_[3], uint64_t n, orc_pcg32 *rng, float *wo_out, float *weight_out) {
    v3 wi = ld3(wi_);
    for (uint64_t k = 0; k < n; ++k) {
        float x = orc_pcg32_next_float(rng), y = orc_pcg32_next_float(rng);
        v3 wo; float eta; int measure;
        v3 w = bsdf_sample(b, wi, x, y, &wo, &eta, &measure);
        if (wo_out) { wo_out[3 * k] = wo.x; wo_out[3 * k + 1] = wo.y; wo_out[3 * k + 2] = wo.z; }
        weight_out[3 * k] = w.x; weight_out[3 * k + 1] = w.y; weight_out[3 * k + 2] = w.z;
    }
    return 0;
}
