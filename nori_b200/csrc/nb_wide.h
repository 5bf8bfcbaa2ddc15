// nb_wide.h -- 8-wide compressed BVH (after Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs Through
// Compressed Wide BVHs", HPG 2017): layout, host-side collapse of the binary hierarchy, and the traversal step shared by
// the device kernels (nb_kernels.cuh) and a host reference used by the CPU tests.
//
// Why: the binary while-while walk runs at 13 / 32 active lanes (ncu, profiles/r2_ajax-ao_lines.txt: trav_run 55 % of all
// warp instructions at 13.1 lanes) because lanes need different numbers of inner-node steps between two leaves.  Here every
// lane that walks does the SAME work per step -- one 80-byte node = 8 quantised child boxes -- and the triangles of the
// children it hit are postponed into a bit mask, so the lanes of a warp stay in step; a ray visits ~4x fewer nodes and
// fetches 5 x 16 B per node instead of 4 x 16 B per binary node (~3x fewer bytes per ray).
//
// Results cannot depend on the hierarchy: child boxes are quantised OUTWARDS from boxes the builder
// already padded, and closest hits follow the tie rule of the reference's loop (equal t: highest triangle index wins,
// ref: src/mesh.cpp:75, src/accel.cpp:37), so the walk may test candidate triangles in any order.
//
// Node = 5 x 16 bytes:
//   q0: p.x, p.y, p.z (fp32 origin of the quantisation grid = low corner of the node's box), [e.x, e.y, e.z, imask]
//       e.a = biased exponent byte of the grid step 2^(e.a - 127) along axis a; imask bit s: the child in slot s is an inner node
//   q1: child_base (index of the first inner child; inner children are consecutive in slot order), tri_base (first triangle
//       of the node's leaf children, <= 24 of them consecutive), meta[0..3], meta[4..7]
//       meta[s]: 0 = empty slot; inner: 0b001_11000 | s  (low 5 bits 24 + s); leaf: (unary count: 1 -> 001, 2 -> 011, 3 -> 111) << 5 | first triangle offset
//   q2: qlo.x[0..7], qlo.y[0..7]    q3: qlo.z[0..7], qhi.x[0..7]    q4: qhi.y[0..7], qhi.z[0..7]      (one byte per slot)
//       child box along a = [p.a + qlo * step.a, p.a + qhi * step.a]
// Slots are assigned so that slot s lies towards (+/-x, +/-y, +/-z) by its bits (bit 0 = x high side ...): a ray with octant
// bits o (bit a set: direction negative along a) visits the hit children in the order of decreasing (s ^ (7 - o)).
// Triangles: the same 48-byte records as the binary layout (nb_bvh.h), reordered so that a node's leaf children are consecutive.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(__CUDACC__)
#define NB_HD __host__ __device__ __forceinline__
#else
#define NB_HD inline
#include <cmath>
#endif

namespace nb {

constexpr int kWideStack = 40;      // uint2 entries; build_wide checks 2 * depth + 2 <= kWideStack

struct WideOutput {
    std::vector<uint32_t> nodes;    // 20 words per node
    std::vector<float> tris;        // 12 floats per triangle, node-ordered
    uint32_t nnodes = 0;
    int depth = 0;                  // levels of wide nodes
    double seconds = 0;
};

// bnodes: 16 floats per binary node, btris: 12 floats per leaf-ordered triangle -- the layout of nb_bvh.h (from either
// builder); every binary leaf must hold <= 3 triangles.  Returns false (with a message) if the tree cannot be converted.
bool build_wide(const float *bnodes, uint32_t n_bnodes, const float *btris, uint32_t n_btris, WideOutput &out, const char **err);

// ------------------------------------------------------------------ shared helpers
NB_HD float wide_as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; std::memcpy(&f, &u, 4); return f;
#endif
}
NB_HD uint32_t wide_as_uint(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; std::memcpy(&u, &f, 4); return u;
#endif
}
NB_HD float wide_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return std::fmaf(a, b, c);
#endif
}
// byte i of x as an exact float: the byte lands in the mantissa of 2^23 (one PRMT), minus 2^23 (exact)
NB_HD float wide_byte(uint32_t x, int i) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650u + (unsigned) i)) - 8388608.0f;
#else
    return wide_as_float(((x >> (8 * i)) & 0xffu) | 0x4B000000u) - 8388608.0f;
#endif
}
NB_HD int wide_bfind(uint32_t x) {      // index of the highest set bit (x != 0)
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int) x);
#else
    return 31 - __builtin_clz(x);
#endif
}
NB_HD int wide_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
NB_HD uint32_t wide_sext_bytes(uint32_t x) {   // every byte whose bit 7 is set becomes 0xff, the others 0x00
    return ((x >> 7) & 0x01010101u) * 0xffu;
}

struct WideRay {                // per-ray constants of the node test
    float idx, idy, idz;        // 1 / d (zero components replaced by +-1e-24 as in trav_begin)
    float ox, oy, oz;           // origin
    uint32_t octinv4;           // (7 - octant) replicated into the four bytes
    bool negx, negy, negz;
};

// One node step: tests the 8 children of the node at `q` (5 x uint4 as 20 words via ld) against the ray segment [mint, maxt]
// and returns the hit mask: bits 24..31 = hit inner children (bit 24 + (slot ^ octinv)), bits 0..23 = triangles of the hit leaf
// children (bit = offset from tri_base).  child_base / tri_base / imask are returned through the references.
template <typename Load>
NB_HD uint32_t wide_node_test(const Load &ld, const WideRay &R, float mint, float maxt, uint32_t &child_base, uint32_t &tri_base, uint32_t &imask) {
    uint32_t w[20];
    ld(w);                                                 // 5 x 128-bit loads
    const uint32_t e = w[3];
    imask = e >> 24;
    child_base = w[4]; tri_base = w[5];
    // t = (p + q * 2^e - o) / d  =  q * (2^e / d) + (p - o) / d
    const float sx = wide_as_float((e & 0xffu) << 23) * R.idx, sy = wide_as_float(((e >> 8) & 0xffu) << 23) * R.idy, sz = wide_as_float(((e >> 16) & 0xffu) << 23) * R.idz;
    const float bx = (wide_as_float(w[0]) - R.ox) * R.idx, by = (wide_as_float(w[1]) - R.oy) * R.idy, bz = (wide_as_float(w[2]) - R.oz) * R.idz;
    uint32_t hitmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = w[6 + half];
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;            // low 5 bits >= 24  <=>  bits 3 and 4 set
        const uint32_t inner_mask4 = wide_sext_bytes(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (R.octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        // near / far planes by the sign of the direction
        const uint32_t lox = w[8 + half], loy = w[10 + half], loz = w[12 + half], hix = w[14 + half], hiy = w[16 + half], hiz = w[18 + half];
        const uint32_t nx = R.negx ? hix : lox, fx = R.negx ? lox : hix;
        const uint32_t ny = R.negy ? hiy : loy, fy = R.negy ? loy : hiy;
        const uint32_t nz = R.negz ? hiz : loz, fz = R.negz ? loz : hiz;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 4; ++i) {
            const float t0x = wide_fma(wide_byte(nx, i), sx, bx), t1x = wide_fma(wide_byte(fx, i), sx, bx);
            const float t0y = wide_fma(wide_byte(ny, i), sy, by), t1y = wide_fma(wide_byte(fy, i), sy, by);
            const float t0z = wide_fma(wide_byte(nz, i), sz, bz), t1z = wide_fma(wide_byte(fz, i), sz, bz);
            const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, mint));
            const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, maxt));
            if (cmin <= cmax) hitmask |= ((child_bits4 >> (8 * i)) & 0xffu) << ((bit_index4 >> (8 * i)) & 0xffu);
        }
    }
    return hitmask;
}

}  // namespace nb
