// nb_aux.cu -- device entry points that are NOT on the render path, compiled as their own translation unit so that
// adding one never perturbs the code ptxas generates for the render kernels in nb_api.cu (the noinline BSDF functions
// are shared per translation unit; a second caller in the same unit changed the path tracers' SASS).
// nb_device.cuh defines non-inline __device__ functions (bsdf_sample, bsdf_eval_pdf); a second translation unit needs
// its own copies under another namespace name, or the host-side stubs collide at link time.
#define nb nb_aux
#include "nb_device.cuh"
#include <cstring>

namespace nb {

// Batched BSDF::sample / BSDF::eval + pdf (ref: include/nori/bsdf.h:59-87) for the callers of the BSDF plugins that are not
// the render loop: the reference's t-test in BSDF mode and its chi^2 test (ref: src/ttest.cpp:104-125,
// src/chi2test.cpp:113-153).  wi_stride 0 = one incident direction for the whole batch (both tests fix wi), 3 = per query.
// mode 0: a = xi (2 per query), out = (wo.xyz, weight.rgb, pdf, measure) 8 floats; mode 1: a = wo (3 per query), out = (f.rgb, pdf).
__global__ void __launch_bounds__(128) bsdf_query_kernel(DevMesh m, unsigned long long n, const float *wi, int wi_stride,
                                                         const float *a, int mode, float *out) {
    for (unsigned long long k = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (unsigned long long) gridDim.x * blockDim.x) {
        const float *w = wi + (size_t) wi_stride * k;
        const V3 wiv = mk(w[0], w[1], w[2]);
        if (mode == 0) {
            V3 wo; int measure; float pdf;
            const V3 f = bsdf_sample(m, wiv, a[2 * k], a[2 * k + 1], wo, measure, pdf);
            float *o = out + 8 * k;
            o[0] = wo.x; o[1] = wo.y; o[2] = wo.z; o[3] = f.x; o[4] = f.y; o[5] = f.z; o[6] = pdf; o[7] = (float) measure;
        } else {
            float pdf;
            const V3 f = bsdf_eval_pdf(m, wiv, mk(a[3 * k], a[3 * k + 1], a[3 * k + 2]), pdf);
            float *o = out + 4 * k;
            o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = pdf;
        }
    }
}

}  // namespace nb

// launcher used by nb_api.cu (nb_bsdf_sample / nb_bsdf_eval_pdf)
// (devmesh: the bytes of a DevMesh -- the two units see the same struct under different namespace names)
extern "C" cudaError_t nb_aux_launch_bsdf_query(const void *devmesh, size_t devmesh_bytes, unsigned long long n, const float *wi, int wi_stride,
                                     const float *a, int mode, float *out, int grid, cudaStream_t s) {
    nb::DevMesh m;
    if (devmesh_bytes != sizeof m) return cudaErrorInvalidValue;
    std::memcpy(&m, devmesh, sizeof m);
    nb::bsdf_query_kernel<<<grid, 128, 0, s>>>(m, n, wi, wi_stride, a, mode, out);
    return cudaGetLastError();
}
