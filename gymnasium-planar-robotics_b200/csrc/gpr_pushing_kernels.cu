// gpr_pushing_kernels.cu — pushing kernels; built with -fmad=false (see gpr_pushing.cuh).
#include <algorithm>

#include "gpr_launch.h"

namespace gpr {

template <bool BOX, bool NOISE>
static cudaError_t launch_push_bn(PushKernel which, const PushArgs& a, int num_sms, cudaStream_t s) {
    const int threads = 128;
    const unsigned blocks = (unsigned)((a.B + threads - 1) / threads);
    if (which == PUSH_RESET) {
        pushing_reset_kernel<BOX, NOISE><<<blocks, threads, 0, s>>>(a);
    } else if (which == PUSH_STEP) {
        pushing_step_kernel<BOX, NOISE><<<blocks, kPushCta, 0, s>>>(a);
    } else {
        // persistent warps pulling queued envs 32 at a time: at most one warp per 32 envs, at most a full residency
        const long long warps = ((long long)a.B + 31) / 32;
        const long long ctas = std::min<long long>((warps + kPushContactCta / 32 - 1) / (kPushContactCta / 32), (long long)num_sms * GPR_PUSH_CONTACT_MINB);
        pushing_contact_kernel<BOX, NOISE><<<(unsigned)std::max<long long>(ctas, 1), kPushContactCta, 0, s>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_push(PushKernel which, bool box, bool noise, const PushArgs& a, int num_sms, cudaStream_t s) {
    if (box) return noise ? launch_push_bn<true, true>(which, a, num_sms, s) : launch_push_bn<true, false>(which, a, num_sms, s);
    return noise ? launch_push_bn<false, true>(which, a, num_sms, s) : launch_push_bn<false, false>(which, a, num_sms, s);
}

}  // namespace gpr
