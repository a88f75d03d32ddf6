// gpr_pushing.cuh — fused kernels of BenchmarkPushingEnv's step path (placeholder until the planar contact model lands).
#pragma once

#include "gpr_device.cuh"

namespace gpr {

struct PushArgs {
    int B;
    const float2* action;
    gpr_outputs out;
    const uint8_t* reset_mask;
    const double2* inject_start;
    const double2* inject_goal;
    const double2* inject_object;
};

}  // namespace gpr
