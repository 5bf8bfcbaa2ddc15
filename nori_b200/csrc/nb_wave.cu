// nb_wave.cu -- launchers of the wavefront engine (nb_wave.cuh; nb_set_option(ctx, "engine", 2)).
//
// Its own translation unit for two reasons: the kernels of nb_api.cu must not change by an instruction when this engine is
// edited, and nb_device.cuh defines non-inline device functions whose host stubs would collide at link time -- hence the
// private namespace name.  shade<INTEG>() is compiled here with NB_WAVEFRONT = 1: occlusion rays go to the engine's shadow
// queue (nb_kernels.cuh: occ_push) instead of being traced by the thread that generated them.
#define nb nb_wv
#define NB_WAVEFRONT 1
#include "nb_wave.cuh"
#include <cstring>

namespace {
template <int INTEG>
cudaError_t launch_logic(const nb::RenderParams &P, cudaStream_t s) {
    nb::wf_logic_kernel<INTEG><<<P.wf_pool / 128u, 128, 0, s>>>(P);
    return cudaGetLastError();
}
}  // namespace

// (params: the bytes of a RenderParams -- the translation units see the same struct under different namespace names)
extern "C" cudaError_t nb_wv_launch_logic(const void *params, size_t bytes, int integ, cudaStream_t s) {
    nb::RenderParams P;
    if (bytes != sizeof P) return cudaErrorInvalidValue;
    std::memcpy(&P, params, sizeof P);
    switch (integ) {
        case 0: return launch_logic<0>(P, s); case 1: return launch_logic<1>(P, s); case 2: return launch_logic<2>(P, s);
        case 3: return launch_logic<3>(P, s); case 4: return launch_logic<4>(P, s); case 5: return launch_logic<5>(P, s);
        default: return launch_logic<6>(P, s);
    }
}

extern "C" cudaError_t nb_wv_launch_trace(const void *params, size_t bytes, int count, int grid, cudaStream_t s) {
    nb::RenderParams P;
    if (bytes != sizeof P) return cudaErrorInvalidValue;
    std::memcpy(&P, params, sizeof P);
    if (count) nb::wf_trace_kernel<true><<<grid, 128, 0, s>>>(P);
    else nb::wf_trace_kernel<false><<<grid, 128, 0, s>>>(P);
    return cudaGetLastError();
}

extern "C" cudaError_t nb_wv_occupancy(int count, int *blocks_trace) {
    if (count) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_trace, nb::wf_trace_kernel<true>, 128, 0);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_trace, nb::wf_trace_kernel<false>, 128, 0);
}

extern "C" int nb_wv_columns(void) { return nb::kWfCols; }
