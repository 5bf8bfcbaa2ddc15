// nb_wavefront.cu -- the deferred-occlusion engine (nb_set_option(ctx, "engine", 1); experimental, off by default).
//
// Same kernels as nb_api.cu, compiled a second time with NB_DEFER_SHADOW = 1 into their own namespace: the render kernel
// no longer traces ambient-occlusion / next-event / point-light shadow rays itself but appends them, with the radiance
// they would contribute, to a queue in device memory (occ_push); occlusion_kernel then traces the queue with dynamic
// fetch and splats the unoccluded contributions (the reconstruction filter is linear, so a sample may reach the film in
// pieces).  Why: in the lock-step kernel an occlusion wave is only as wide as the number of lanes that hit something and
// lasts as long as its longest any-hit walk; a lock-step simulation on recorded walk lengths of the headline workload
// puts those waves at ~0.27 lane utilisation against ~0.75 for the coherent camera-ray waves.
// A separate translation unit because (a) the validated kernels of nb_api.cu must not change by a single instruction and
// (b) nb_device.cuh defines non-inline device functions whose host stubs would collide at link time.
#define nb nb_wf
#define NB_DEFER_SHADOW 1
#include "nb_kernels.cuh"
#include <cstring>

namespace {

template <int INTEG>
cudaError_t launch_render(const nb::RenderParams &P, bool count, int grid, cudaStream_t s) {
    if (count) nb::render_kernel<INTEG, true, false><<<grid, 128, 0, s>>>(P);
    else nb::render_kernel<INTEG, false, false><<<grid, 128, 0, s>>>(P);
    return cudaGetLastError();
}

template <int INTEG>
cudaError_t occupancy(int *blocks, bool count) {
    if (count) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, nb::render_kernel<INTEG, true, false>, 128, 0);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, nb::render_kernel<INTEG, false, false>, 128, 0);
}

}  // namespace

// (params: the bytes of a RenderParams -- both translation units see the same struct under different namespace names)
extern "C" cudaError_t nb_wf_occupancy(int integ, int count, int *blocks_render, int *blocks_occlusion) {
    cudaError_t e;
    switch (integ) {
        case 0: e = occupancy<0>(blocks_render, count != 0); break; case 1: e = occupancy<1>(blocks_render, count != 0); break;
        case 2: e = occupancy<2>(blocks_render, count != 0); break; case 3: e = occupancy<3>(blocks_render, count != 0); break;
        case 4: e = occupancy<4>(blocks_render, count != 0); break; case 5: e = occupancy<5>(blocks_render, count != 0); break;
        default: e = occupancy<6>(blocks_render, count != 0); break;
    }
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_occlusion, nb::occlusion_kernel, 128, 0);
}

extern "C" cudaError_t nb_wf_launch_render(const void *params, size_t bytes, int integ, int count, int grid, cudaStream_t s) {
    nb::RenderParams P;
    if (bytes != sizeof P) return cudaErrorInvalidValue;
    std::memcpy(&P, params, sizeof P);
    switch (integ) {
        case 0: return launch_render<0>(P, count != 0, grid, s); case 1: return launch_render<1>(P, count != 0, grid, s);
        case 2: return launch_render<2>(P, count != 0, grid, s); case 3: return launch_render<3>(P, count != 0, grid, s);
        case 4: return launch_render<4>(P, count != 0, grid, s); case 5: return launch_render<5>(P, count != 0, grid, s);
        default: return launch_render<6>(P, count != 0, grid, s);
    }
}

extern "C" cudaError_t nb_wf_launch_occlusion(const void *params, size_t bytes, int grid, cudaStream_t s) {
    nb::RenderParams P;
    if (bytes != sizeof P) return cudaErrorInvalidValue;
    std::memcpy(&P, params, sizeof P);
    nb::occlusion_kernel<<<grid, 128, 0, s>>>(P);
    return cudaGetLastError();
}
