// gpr_planning_g.cu — planning kernels for ONE lane-group width (compile with -DGPR_G=<1|2|4|8|16|32>).
#include <algorithm>

#include "gpr_launch.h"

#ifndef GPR_G
#error "compile with -DGPR_G=<lane group width>"
#endif

namespace gpr {

template <int G, bool BOX, bool NOISE>
static cudaError_t launch_plan_gbn(PlanKernel which, const PlanArgs& a_in, int num_sms, cudaStream_t s) {
    const int threads = 256;
    const long long lanes = (long long)a_in.B * G;
    const unsigned blocks = (unsigned)((lanes + threads - 1) / threads);
    const unsigned step_blocks = (unsigned)((lanes + StepThreads<G>::value - 1) / StepThreads<G>::value);
    PlanArgs a = a_in;
    a.step_ctas = step_blocks;
    const bool extra = a.out_f64 != 0 || a.compact_index != nullptr;  // planning_autoreset_kernel waits until all of its warps have reported (reset_ctl)
    if (which == PLAN_RESET) {
        planning_reset_kernel<G, BOX, NOISE><<<blocks, threads, 0, s>>>(a);
    } else if (which == PLAN_STEP) {
        // (learn_jerk is a template parameter of the step kernel only: it sits inside the 40-cycle loop)
        // EXTRA: float64 outputs or compact transport asked for; the plain instantiation carries none of that code
        if (extra) {
            if (a.learn_jerk) planning_step_kernel<G, BOX, NOISE, true, true><<<step_blocks, StepThreads<G>::value, 0, s>>>(a);
            else planning_step_kernel<G, BOX, NOISE, false, true><<<step_blocks, StepThreads<G>::value, 0, s>>>(a);
        } else {
            if (a.learn_jerk) planning_step_kernel<G, BOX, NOISE, true, false><<<step_blocks, StepThreads<G>::value, 0, s>>>(a);
            else planning_step_kernel<G, BOX, NOISE, false, false><<<step_blocks, StepThreads<G>::value, 0, s>>>(a);
        }
    } else {
        // one warp per finished env, pulled through an atomic cursor: a fixed grid of 8 CTAs (4 warps each) per SM
        const unsigned ab = (unsigned)std::min<long long>(((long long)a.B + 3) / 4, (long long)num_sms * 8);
        if (!a.overlap) {
            if (extra) planning_autoreset_kernel<G, BOX, NOISE, true><<<ab, 128, 0, s>>>(a);
            else planning_autoreset_kernel<G, BOX, NOISE, false><<<ab, 128, 0, s>>>(a);
        } else {
            // programmatic dependent launch: resident as soon as every CTA of the step grid (the previous launch in
            // this stream) has started; the kernel synchronises with the step kernel through the work list itself
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ab);
            cfg.blockDim = dim3(128);
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            return extra ? cudaLaunchKernelEx(&cfg, planning_autoreset_kernel<G, BOX, NOISE, true>, a)
                         : cudaLaunchKernelEx(&cfg, planning_autoreset_kernel<G, BOX, NOISE, false>, a);
        }
    }
    return cudaGetLastError();
}

template <int G>
cudaError_t launch_plan_g(PlanKernel which, bool box, bool noise, const PlanArgs& a, int num_sms, cudaStream_t s) {
    if (box) return noise ? launch_plan_gbn<G, true, true>(which, a, num_sms, s) : launch_plan_gbn<G, true, false>(which, a, num_sms, s);
    return noise ? launch_plan_gbn<G, false, true>(which, a, num_sms, s) : launch_plan_gbn<G, false, false>(which, a, num_sms, s);
}

template cudaError_t launch_plan_g<GPR_G>(PlanKernel, bool, bool, const PlanArgs&, int, cudaStream_t);

}  // namespace gpr
