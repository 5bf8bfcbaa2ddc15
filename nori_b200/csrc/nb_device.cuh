// nb_device.cuh -- device-side math of the render hot path (sm_100a).
//
// Built with -fmad=false: every fp32 multiply and add below rounds separately, in the SAME order as
// the CPU oracle, so that paths do not decorrelate.  Fused multiply-adds are used
// only where results are provably independent of them (BVH slab tests, which may only cull -- they
// are written with explicit __fmaf_rn).  Division and sqrt are IEEE (nvcc defaults -prec-div/-prec-sqrt).
//
// "ref:" citations are relative to /root/reference (wjakob/nori @ 092f581).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nb {

#define NB_EPSILON 1e-4f                       // ref: include/nori/common.h:38
#define NB_PI 3.14159265358979323846f          // ref: include/nori/common.h:43
#define NB_INV_PI 0.31830988618379067154f      // ref: include/nori/common.h:44
#define NB_INF __int_as_float(0x7f800000)

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
// 3-element reductions associate as a0 + (a1 + a2) (Eigen's unrolled redux; see oracle.c header)
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 normalize(V3 a) { float n = sqrtf(dot(a, a)); return mk(a.x / n, a.y / n, a.z / n); }
__device__ __forceinline__ float max3(V3 a) { float m = a.x > a.y ? a.x : a.y; return m > a.z ? m : a.z; }
__device__ __forceinline__ bool is_zero(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
__device__ __forceinline__ V3 lin3(float b0, V3 p0, float b1, V3 p1, float b2, V3 p2) {   // b0*p0 + b1*p1 + b2*p2, left to right
    return mk(b0 * p0.x + b1 * p1.x + b2 * p2.x, b0 * p0.y + b1 * p1.y + b2 * p2.y, b0 * p0.z + b1 * p1.z + b2 * p2.z);
}
__device__ __forceinline__ V3 xyz(float4 v) { return mk(v.x, v.y, v.z); }

// ------------------------------------------------------------------ pcg32 (wjakob/pcg32 @ 70099ead; call sites ref: src/independent.cpp:36-55)
struct Pcg32 { uint64_t state, inc; };
__device__ __forceinline__ uint32_t pcg_next_uint(Pcg32 &r) {
    uint64_t old = r.state;
    r.state = old * 0x5851f42d4c957f2dULL + r.inc;
    uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xs >> rot) | (xs << ((~rot + 1u) & 31u));
}
__device__ __forceinline__ void pcg_seed(Pcg32 &r, uint64_t initstate, uint64_t initseq) {
    r.state = 0u; r.inc = (initseq << 1u) | 1u;
    pcg_next_uint(r); r.state += initstate; pcg_next_uint(r);
}
// pcg32::advance -- O(log delta) skip-ahead of the LCG (Brown, "Random Number Generation with Arbitrary Strides")
__device__ __forceinline__ void pcg_advance(Pcg32 &r, uint64_t delta) {
    uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = r.inc, acc_mult = 1u, acc_plus = 0u;
    while (delta > 0) {
        if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    r.state = acc_mult * r.state + acc_plus;
}
__device__ __forceinline__ float pcg_next_float(Pcg32 &r) {
    return __uint_as_float((pcg_next_uint(r) >> 9) | 0x3f800000u) - 1.0f;
}

// ------------------------------------------------------------------ deterministic transcendentals (same polynomials as the oracle)
__device__ __forceinline__ void sincos2pi(float u, float &so, float &co) {
    float u8 = u * 8.0f;
    int k = (int) u8;
    float f = u8 - (float) k;
    int j = (k + 1) >> 1;
    float r = (k & 1) ? (f - 1.0f) : f;
    float x = r * 0.78539816339744830962f;
    float z = x * x;
    float s = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * x + x;
    float c = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    switch (j & 3) {
        case 0: so = s; co = c; break;
        case 1: so = c; co = -s; break;
        case 2: so = -s; co = -c; break;
        default: so = -c; co = s; break;
    }
}
__device__ __forceinline__ float det_logf(float xin) {
    uint32_t b = __float_as_uint(xin);
    int e = (int) ((b >> 23) & 0xff) - 126;
    float x = __uint_as_float((b & 0x007fffffu) | 0x3f000000u);
    if (x < 0.70710678118654752440f) { e -= 1; x = x + x - 1.0f; } else { x = x - 1.0f; }
    float z = x * x;
    float y = ((((((((7.0376836292e-2f * x - 1.1514610310e-1f) * x + 1.1676998740e-1f) * x - 1.2420140846e-1f) * x
              + 1.4249322787e-1f) * x - 1.6668057665e-1f) * x + 2.0000714765e-1f) * x - 2.4999993993e-1f) * x
              + 3.3333331174e-1f) * x * z;
    float fe = (float) e;
    y = y + (-2.12194440e-4f * fe);
    y = y + (-0.5f * z);
    z = x + y;
    z = z + 0.693359375f * fe;
    return z;
}
__device__ __forceinline__ float det_expf(float x) {
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    float z = floorf(1.44269504088896341f * x + 0.5f);
    x = x - z * 0.693359375f;
    x = x - z * -2.12194440e-4f;
    int n = (int) z;
    z = x * x;
    z = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x + 4.1665795894e-2f) * x
        + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    return z * __uint_as_float((uint32_t) (n + 127) << 23);
}

// ------------------------------------------------------------------ warps [authored]; interface ref: include/nori/warp.h:44-57
__device__ __forceinline__ V3 square_to_cosine_hemisphere(float x, float y) {
    float r = sqrtf(x), s, c;
    sincos2pi(y, s, c);
    float zz = 1.0f - x;
    return mk(r * c, r * s, sqrtf(zz > 0.0f ? zz : 0.0f));
}
__device__ __forceinline__ V3 square_to_beckmann(float x, float y, float alpha) {
    float s, c;
    sincos2pi(x, s, c);
    float tan2 = -(alpha * alpha) * det_logf(1.0f - y);
    float ct = 1.0f / sqrtf(1.0f + tan2);
    float st2 = 1.0f - ct * ct;
    float st = sqrtf(st2 > 0.0f ? st2 : 0.0f);
    return mk(st * c, st * s, ct);
}
__device__ __forceinline__ float beckmann_D(V3 m, float alpha) {
    if (m.z <= 0.0f) return 0.0f;
    float c2 = m.z * m.z;
    float tan2 = (1.0f - c2) / c2;
    float a2 = alpha * alpha;
    return det_expf(-tan2 / a2) / (NB_PI * a2 * (c2 * c2));
}

// ref: src/common.cpp:259-288
__device__ __forceinline__ float fresnel(float cosThetaI, float extIOR, float intIOR) {
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

// Frame(n) / coordinateSystem: ref include/nori/frame.h:36-50, src/common.cpp:248-257
struct Frame { V3 s, t, n; };
__device__ __forceinline__ Frame frame_from_n(V3 a) {
    Frame f; f.n = a;
    V3 c;
    if (fabsf(a.x) > fabsf(a.y)) {
        float invLen = 1.0f / sqrtf(a.x * a.x + a.z * a.z);
        c = mk(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(a.y * a.y + a.z * a.z);
        c = mk(0.0f, a.z * invLen, -a.y * invLen);
    }
    f.t = c;
    f.s = cross(c, a);
    return f;
}
__device__ __forceinline__ V3 to_local(const Frame &f, V3 v) { return mk(dot(v, f.s), dot(v, f.t), dot(v, f.n)); }
__device__ __forceinline__ V3 to_world(const Frame &f, V3 v) {
    return mk(f.s.x * v.x + f.t.x * v.y + f.n.x * v.z, f.s.y * v.x + f.t.y * v.y + f.n.y * v.z, f.s.z * v.x + f.t.z * v.y + f.n.z * v.z);
}

// ------------------------------------------------------------------ device scene tables
struct DevMesh {               // one per mesh (plugin parameters captured on the host)
    int32_t bsdf_type;
    float albedo[3];
    float alpha, intIOR, extIOR, ks;
    int32_t emitter_type;
    float radiance[3];
    uint32_t prim_offset, nf;
    uint32_t flags;            // 1: has normals, 2: has uvs
    float area_sum;            // sum of triangle areas (emitters)
    uint32_t cdf_offset;       // into emitter_cdf (nf + 1 entries)
    uint32_t pad;
};

// ------------------------------------------------------------------ BSDFs (interface ref: include/nori/bsdf.h:59-87)
__device__ __forceinline__ bool bsdf_is_diffuse(const DevMesh &m) { return m.bsdf_type == 0 || m.bsdf_type == 3; }

__device__ __forceinline__ float mf_G1(V3 wv, V3 wh, float alpha) {
    if (dot(wv, wh) / wv.z <= 0.0f) return 0.0f;
    float c2 = wv.z * wv.z;
    float s2 = 1.0f - c2;
    if (s2 <= 0.0f) return 1.0f;
    float tanv = sqrtf(s2) / wv.z;
    float b = 1.0f / (alpha * tanv);
    if (b >= 1.6f) return 1.0f;
    float b2 = b * b;
    return (3.535f * b + 2.181f * b2) / (1.0f + 2.276f * b + 2.577f * b2);
}

// eval() and pdf() of the same (wi, wo) in one pass.  The microfacet terms they share (half vector, Beckmann D) are
// computed once; every value is produced by the same operation sequence as the separate functions of the oracle
// (oracle.c: bsdf_eval / bsdf_pdf), so the results are bit-identical.
__device__ __noinline__ V3 bsdf_eval_pdf(const DevMesh &m, V3 wi, V3 wo, float &pdf) {
    pdf = 0.0f;
    if (m.bsdf_type == 0) {            // ref: src/diffuse.cpp:23-33, 36-52
        if (wi.z <= 0 || wo.z <= 0) return mk(0, 0, 0);
        pdf = NB_INV_PI * wo.z;
        return mk(m.albedo[0] * NB_INV_PI, m.albedo[1] * NB_INV_PI, m.albedo[2] * NB_INV_PI);
    }
    if (m.bsdf_type == 3) {            // [authored] contract ref: src/microfacet.cpp:40-47
        if (wi.z <= 0 || wo.z <= 0) return mk(0, 0, 0);
        const V3 wh = normalize(wi + wo);
        const float D = beckmann_D(wh, m.alpha);
        const float F = fresnel(dot(wh, wi), m.extIOR, m.intIOR);
        const float G = mf_G1(wi, wh, m.alpha) * mf_G1(wo, wh, m.alpha);
        const float spec = m.ks * D * F * G / (4.0f * wi.z * wo.z);
        const float Jh = 1.0f / (4.0f * dot(wh, wo));
        pdf = m.ks * D * wh.z * Jh + (1.0f - m.ks) * wo.z * NB_INV_PI;
        return mk(m.albedo[0] * NB_INV_PI + spec, m.albedo[1] * NB_INV_PI + spec, m.albedo[2] * NB_INV_PI + spec);
    }
    return mk(0, 0, 0);                // discrete: ref src/mirror.cpp:17-24, src/dielectric.cpp:23-30
}

// returns weight = eval*cos/pdf (0 <=> invalid); measure: 1 solid angle, 2 discrete
// `pdf` receives pdf(wi, wo) of the sampled direction (0 for discrete lobes and invalid samples)
__device__ __noinline__ V3 bsdf_sample(const DevMesh &m, V3 wi, float xi_x, float xi_y, V3 &wo, int &measure, float &pdf) {
    measure = 1; wo = mk(0, 0, 1); pdf = 0.0f;
    switch (m.bsdf_type) {
        case 0:                        // ref: src/diffuse.cpp:55-71
            if (wi.z <= 0) return mk(0, 0, 0);
            wo = square_to_cosine_hemisphere(xi_x, xi_y);
            pdf = wo.z <= 0 ? 0.0f : NB_INV_PI * wo.z;                  // == bsdf_pdf(diffuse): ref src/diffuse.cpp:36-52
            return mk(m.albedo[0], m.albedo[1], m.albedo[2]);
        case 1:                        // ref: src/mirror.cpp:27-43
            if (wi.z <= 0) return mk(0, 0, 0);
            wo = mk(-wi.x, -wi.y, wi.z);
            measure = 2;
            return mk(1, 1, 1);
        case 2: {                      // [authored] stub at ref: src/dielectric.cpp:33-35
            float cosI = wi.z;
            float F = fresnel(cosI, m.extIOR, m.intIOR);
            measure = 2;
            if (xi_x < F) { wo = mk(-wi.x, -wi.y, wi.z); return mk(1, 1, 1); }
            float etaI = m.extIOR, etaT = m.intIOR;
            if (cosI < 0.0f) { float t = etaI; etaI = etaT; etaT = t; cosI = -cosI; }
            float e = etaI / etaT;
            float sin2T = e * e * (1 - cosI * cosI);
            float cosT = sqrtf(1.0f - sin2T);
            wo = mk(-e * wi.x, -e * wi.y, wi.z > 0 ? -cosT : cosT);
            return mk(1, 1, 1);
        }
        case 3: {                      // [authored] contract ref: src/microfacet.cpp:50-57
            if (wi.z <= 0) return mk(0, 0, 0);
            if (xi_x < m.ks) {
                float x = xi_x / m.ks;
                V3 wh = square_to_beckmann(x, xi_y, m.alpha);
                float d2 = 2.0f * dot(wh, wi);
                wo = mk(d2 * wh.x - wi.x, d2 * wh.y - wi.y, d2 * wh.z - wi.z);
            } else {
                float x = (xi_x - m.ks) / (1.0f - m.ks);
                wo = square_to_cosine_hemisphere(x, xi_y);
            }
            if (wo.z <= 0) return mk(0, 0, 0);
            float p;
            const V3 f = bsdf_eval_pdf(m, wi, wo, p);
            if (!(p > 0.0f)) return mk(0, 0, 0);
            pdf = p;
            return mk(f.x * wo.z / p, f.y * wo.z / p, f.z * wo.z / p);
        }
    }
    return mk(0, 0, 0);
}

}  // namespace nb
