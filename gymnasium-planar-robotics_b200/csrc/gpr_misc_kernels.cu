// gpr_misc_kernels.cu — small utility kernels.
#include "gpr_launch.h"

namespace gpr {

// HER relabelling (plan:502-534 / push:499-527 on arrays): one thread per transition.
template <typename T>
struct Vec2;
template <>
struct Vec2<float> {
    using type = float2;
};
template <>
struct Vec2<double> {
    using type = double2;
};

template <typename T>
__global__ void compute_reward_kernel(int kind, int N, int batch, double threshold, const T* __restrict__ achieved,
                                      const T* __restrict__ desired, const uint8_t* __restrict__ mcol,
                                      const uint8_t* __restrict__ wcol, float* __restrict__ reward,
                                      uint8_t* __restrict__ terminated) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const bool mc = mcol ? mcol[b] != 0 : false, wc = wcol ? wcol[b] != 0 : false;
    if (kind == GPR_ENV_PLANNING) {
        int reached = 0;
        using V = typename Vec2<T>::type;
        const V* ag = reinterpret_cast<const V*>(achieved) + (size_t)b * N;
        const V* dg = reinterpret_cast<const V*>(desired) + (size_t)b * N;
        for (int m = 0; m < N; ++m) {
            const V a2 = ag[m], d2 = dg[m];
            const double dx = dsub((double)a2.x, (double)d2.x), dy = dsub((double)a2.y, (double)d2.y);
            reached += sqrt_le(dadd(dmul(dx, dx), dmul(dy, dy)), threshold) ? 1 : 0;
        }
        float r;
        bool t, s;
        planning_reward(N, reached, mc, wc, r, t, s);
        if (reward) reward[b] = r;
        if (terminated) terminated[b] = t;
    } else {
        using V = typename Vec2<T>::type;
        const V a2 = reinterpret_cast<const V*>(achieved)[b], d2 = reinterpret_cast<const V*>(desired)[b];
        const double dx = dsub((double)a2.x, (double)d2.x), dy = dsub((double)a2.y, (double)d2.y);
        const bool reached = sqrt_le(dadd(dmul(dx, dx), dmul(dy, dy)), threshold);
        const float r = wc ? -50.f : (reached ? 0.f : -1.f);  // push:521-523
        if (reward) reward[b] = r;
        if (terminated) terminated[b] = wc;  // push:475
    }
}


cudaError_t launch_compute_reward(int kind, int N, int batch, double threshold, const void* achieved, const void* desired, bool f64,
                                  const uint8_t* mcol, const uint8_t* wcol, float* reward, uint8_t* terminated, cudaStream_t s) {
    const int threads = 256;
    const unsigned blocks = (unsigned)((batch + threads - 1) / threads);
    if (f64)
        compute_reward_kernel<double><<<blocks, threads, 0, s>>>(kind, N, batch, threshold, (const double*)achieved, (const double*)desired,
                                                                 mcol, wcol, reward, terminated);
    else
        compute_reward_kernel<float><<<blocks, threads, 0, s>>>(kind, N, batch, threshold, (const float*)achieved, (const float*)desired,
                                                                mcol, wcol, reward, terminated);
    return cudaGetLastError();
}

}  // namespace gpr
