// gpr_launch.h — launcher prototypes: the kernels live in separate translation units (one per lane-group width G for
// planning, one for pushing compiled with -fmad=false, one for the small utility kernels) so they build in parallel.
#pragma once

#include <cuda_runtime.h>

#include "gpr_planning.cuh"
#include "gpr_pushing.cuh"

namespace gpr {

enum PlanKernel { PLAN_STEP = 0, PLAN_RESET = 1, PLAN_AUTORESET = 2 };

// defined (explicitly instantiated) in gpr_planning_g.cu, once per G in {1, 2, 4, 8, 16, 32}
template <int G>
cudaError_t launch_plan_g(PlanKernel which, bool box, bool noise, const PlanArgs& a, int num_sms, cudaStream_t s);

// gpr_pushing_kernels.cu
enum PushKernel { PUSH_STEP = 0, PUSH_RESET = 1, PUSH_CONTACT = 2 };
cudaError_t launch_push(PushKernel which, bool box, bool noise, const PushArgs& a, int num_sms, cudaStream_t s);

// gpr_misc_kernels.cu
cudaError_t launch_compute_reward(int kind, int N, int batch, double threshold, const void* achieved, const void* desired, bool f64,
                                  const uint8_t* mcol, const uint8_t* wcol, float* reward, uint8_t* terminated, cudaStream_t s);

}  // namespace gpr
