// gpr_pushing_kernels.cu — pushing kernels; built with -fmad=false (see gpr_pushing.cuh).
#include "gpr_launch.h"

namespace gpr {

template <bool BOX, bool NOISE>
static cudaError_t launch_push_bn(bool reset, const PushArgs& a, cudaStream_t s) {
    const int threads = 128;
    const unsigned blocks = (unsigned)((a.B + threads - 1) / threads);
    if (reset)
        pushing_reset_kernel<BOX, NOISE><<<blocks, threads, 0, s>>>(a);
    else
        pushing_step_kernel<BOX, NOISE><<<blocks, threads, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_push(bool reset, bool box, bool noise, const PushArgs& a, cudaStream_t s) {
    if (box) return noise ? launch_push_bn<true, true>(reset, a, s) : launch_push_bn<true, false>(reset, a, s);
    return noise ? launch_push_bn<false, true>(reset, a, s) : launch_push_bn<false, false>(reset, a, s);
}

}  // namespace gpr
