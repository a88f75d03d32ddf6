// gpr_planning.cuh — fused kernels of BenchmarkPlanningEnv's step path.
//
//   planning_step_kernel : basic:1835-1950 in one launch — action clip, num_cycles x { plan:420-450 limit control,
//                          mj_step-equivalent integration, basic:459-788 wall check, basic:355-424 mover check, break on
//                          collision }, plan:536-573 observation, plan:575-602 info, plan:502-534 reward,
//                          plan:459-479 terminated, TimeLimit truncation, episode statistics and auto-reset
//                          (plan:355-418 rejection sampling with the counter-based RNG of include/gpr_rng.h).
//   planning_reset_kernel: basic:1770-1833 for a masked subset, optionally with injected starts / goals.
#pragma once

#include "gpr_device.cuh"

namespace gpr {

struct PlanArgs {
    int B, N;
    int learn_jerk, num_cycles, max_episode_steps, autoreset, max_reset_attempts, quirks;
    uint32_t env_base;
    uint64_t seed;
    double dt, v_max, a_max, j_max, act_lim;
    double v_max2_lo, a_max2_lo;  // max^2 * (1 - 1e-14), see ensure_max
    double threshold, min_goal_dist;
    double min_xy[2], span_xy[2];
    double sigma_p, sigma_v;
    double quirk_rsum[2];  // [safety] max over pairs of r_i + r_j (basic:409 broadcast quirk)
    LayoutArgs L;
    const double* c_wall;   // [2][GPR_MAX_MOVERS][2] device
    const double* c_mover;  // [2][GPR_MAX_MOVERS][2] device
    // state (SoA, float64)
    double2* pos;
    double2* vel;
    double2* acc;
    double2* goal;
    int32_t* elapsed;
    uint32_t* rng;
    uint8_t* needs_reset;
    float* ep_return;
    double* stats;        // 6 accumulators, see gpr_episode_stats
    uint32_t* fail_count;  // number of resets whose rejection loop hit max_reset_attempts
    // per-call I/O
    const float2* action;
    gpr_outputs out;
    // reset-kernel only
    const uint8_t* reset_mask;
    const double2* inject_start;
    const double2* inject_goal;
};

template <int G>
struct Lane {
    unsigned lane;   // lane in warp
    unsigned gmask;  // ballot mask of this lane's group
    int env;         // local env index
    int m;           // mover index within the env
    bool env_ok;     // env < B
    bool active;     // env_ok && m < N
    size_t idx;      // env * N + m
    uint32_t env_global;
};

// One observation row (plan:536-573) + the per-env reductions the reward needs.
template <int G, bool NOISE>
__device__ __forceinline__ void observe(const PlanArgs& a, const Lane<G>& ln, uint32_t event, double2 p, double2 v,
                                        double2 goal, double2& ag, double2& ov, int& reached_cnt) {
    ag = p;
    ov = v;
    if (NOISE) {
        float n4[4];
        gpr_normal4(a.seed, ln.env_global, event, GPR_RNG_OBS, (uint32_t)ln.m, n4);
        ag.x = dadd(p.x, dmul((double)n4[0], a.sigma_p));
        ag.y = dadd(p.y, dmul((double)n4[1], a.sigma_p));
        ov.x = dadd(v.x, dmul((double)n4[2], a.sigma_v));
        ov.y = dadd(v.y, dmul((double)n4[3], a.sigma_v));
    }
    const double dx = dsub(ag.x, goal.x), dy = dsub(ag.y, goal.y);
    const bool reached = ln.active && sqrt_le(dadd(dmul(dx, dx), dmul(dy, dy)), a.threshold);  // plan:521
    reached_cnt = __popc(__ballot_sync(FULL, reached) & ln.gmask);
}

template <int G>
__device__ __forceinline__ void store_obs(const PlanArgs& a, const Lane<G>& ln, float* O, float* AG, float* DG, double2 ov,
                                          double2 acc, double2 ag, double2 goal) {
    if (!ln.active) return;
    const int N = a.N;
    if (O) {
        const size_t row = (size_t)ln.env * (size_t)(2 * N * (1 + a.learn_jerk));
        reinterpret_cast<float2*>(O + row)[ln.m] = make_float2((float)ov.x, (float)ov.y);
        if (a.learn_jerk) reinterpret_cast<float2*>(O + row + 2 * N)[ln.m] = make_float2((float)acc.x, (float)acc.y);
    }
    if (AG) reinterpret_cast<float2*>(AG)[ln.idx] = make_float2((float)ag.x, (float)ag.y);
    if (DG) reinterpret_cast<float2*>(DG)[ln.idx] = make_float2((float)goal.x, (float)goal.y);
}

// plan:355-418 + basic:1797-1805 for the envs with `need` set; warp-collective (every lane of the warp calls it).
template <int G, bool BOX, bool NOISE>
__device__ __forceinline__ void reset_group(const PlanArgs& a, const Tables& tb, const Lane<G>& ln, bool need,
                                            uint32_t event, const double2* inj_start, const double2* inj_goal,
                                            double2& p, double2& v, double2& acc, double2& goal, bool& mc, bool& wc,
                                            bool& failed) {
    const int mm = ln.active ? ln.m : 0;
    const double cw0 = a.c_wall[(1 * GPR_MAX_MOVERS + mm) * 2 + 0], cw1 = a.c_wall[(1 * GPR_MAX_MOVERS + mm) * 2 + 1];
    const double cms0 = a.c_mover[(1 * GPR_MAX_MOVERS + mm) * 2 + 0], cms1 = a.c_mover[(1 * GPR_MAX_MOVERS + mm) * 2 + 1];
    const double cm0 = a.c_mover[(0 * GPR_MAX_MOVERS + mm) * 2 + 0], cm1 = a.c_mover[(0 * GPR_MAX_MOVERS + mm) * 2 + 1];
    const int cap = a.max_reset_attempts > 0 ? a.max_reset_attempts : 1;
    Rect rw, rm;
    failed = false;

    // ---- loop A: all starts at once (plan:369-385)
    bool pend = need && (inj_start == nullptr);
    if (need && inj_start != nullptr && ln.active) p = inj_start[ln.idx];
    for (int t = 0; t < cap; ++t) {
        if (!__any_sync(FULL, pend)) break;
        const gpr_u32x4 r = gpr_rng_block(a.seed, ln.env_global, event, GPR_RNG_RESET_SAMPLE + 2u * (uint32_t)t, (uint32_t)ln.m);
        const double x = dadd(a.min_xy[0], dmul(a.span_xy[0], gpr_uniform53(r.v[0], r.v[1])));  // plan:377
        const double y = dadd(a.min_xy[1], dmul(a.span_xy[1], gpr_uniform53(r.v[2], r.v[3])));
        const bool part = pend && ln.active;
        if (BOX) {
            rect_vertices_axis(x, y, cw0, cw1, rw);
            rect_vertices_axis(x, y, cms0, cms1, rm);
        }
        const bool bad = part && !wall_valid<BOX>(tb, a.L, x, y, cw0, rw);                               // plan:379
        const bool hit = pair_check<G, BOX>(ln.lane, ln.m, part, x, y, cms0, cms1, rm, a.quirks != 0, a.quirk_rsum[1]);  // plan:381
        const bool rej = (__ballot_sync(FULL, bad || hit) & ln.gmask) != 0u;
        if (pend) {
            p = make_double2(x, y);
            if (!rej) pend = false;
        }
    }
    failed |= pend;

    // ---- loop B: all goals at once (plan:395-413)
    pend = need && (inj_goal == nullptr);
    if (need && inj_goal != nullptr && ln.active) goal = inj_goal[ln.idx];
    for (int t = 0; t < cap; ++t) {
        if (!__any_sync(FULL, pend)) break;
        const gpr_u32x4 r =
            gpr_rng_block(a.seed, ln.env_global, event, GPR_RNG_RESET_SAMPLE + 2u * (uint32_t)t + 1u, (uint32_t)ln.m);
        const double x = dadd(a.min_xy[0], dmul(a.span_xy[0], gpr_uniform53(r.v[0], r.v[1])));  // plan:405
        const double y = dadd(a.min_xy[1], dmul(a.span_xy[1], gpr_uniform53(r.v[2], r.v[3])));
        const bool part = pend && ln.active;
        if (BOX) rect_vertices_axis(x, y, cw0, cw1, rw);
        bool bad = part && !wall_valid<BOX>(tb, a.L, x, y, cw0, rw);  // plan:406
        // plan:408-413: any pair closer than min_goal_dist (strict '<') rejects
        if (G > 1) {
            const unsigned base = ln.lane & ~(unsigned)(G - 1);
#pragma unroll
            for (int k = 1; k <= G / 2; ++k) {
                const int src = (int)(base | (unsigned)((ln.m + k) & (G - 1)));
                const double ox = __shfl_sync(FULL, x, src), oy = __shfl_sync(FULL, y, src);
                const bool opart = __shfl_sync(FULL, (int)part, src) != 0;
                const double dx = dsub(x, ox), dy = dsub(y, oy);
                if (part && opart && sqrt_lt(dadd(dmul(dx, dx), dmul(dy, dy)), a.min_goal_dist)) bad = true;
            }
        }
        const bool rej = (__ballot_sync(FULL, bad) & ln.gmask) != 0u;
        if (pend) {
            goal = make_double2(x, y);
            if (!rej) pend = false;
        }
    }
    failed |= pend;

    // ---- fresh MjData (plan:336-353): qvel = act = qacc = 0
    if (need) {
        v = make_double2(0.0, 0.0);
        acc = make_double2(0.0, 0.0);
    }
    // ---- basic:1799-1805: wall check WITH the safety offset, mover check WITHOUT, on independently noisy qpos
    double wx = p.x, wy = p.y, mx = p.x, my = p.y;
    const bool part = need && ln.active;
    if (NOISE) {
        float n4[4];
        gpr_normal4(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK, (uint32_t)ln.m, n4);
        wx = dadd(p.x, dmul((double)n4[0], a.sigma_p));
        wy = dadd(p.y, dmul((double)n4[1], a.sigma_p));
        mx = dadd(p.x, dmul((double)n4[2], a.sigma_p));
        my = dadd(p.y, dmul((double)n4[3], a.sigma_p));
        if (BOX) {
            float q[4];
            gpr_normal4(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK_WQUAT, (uint32_t)ln.m, q);
            rect_vertices(wx, wy, dadd(1.0, dmul((double)q[0], a.sigma_p)), dmul((double)q[1], a.sigma_p),
                          dmul((double)q[2], a.sigma_p), dmul((double)q[3], a.sigma_p), cw0, cw1, rw);
            if (G > 1) {
                gpr_normal4(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK_MQUAT, (uint32_t)ln.m, q);
                rect_vertices(mx, my, dadd(1.0, dmul((double)q[0], a.sigma_p)), dmul((double)q[1], a.sigma_p),
                              dmul((double)q[2], a.sigma_p), dmul((double)q[3], a.sigma_p), cm0, cm1, rm);
            }
        }
    } else if (BOX) {
        rect_vertices_axis(wx, wy, cw0, cw1, rw);
        rect_vertices_axis(mx, my, cm0, cm1, rm);
    }
    const bool bad = part && !wall_valid<BOX>(tb, a.L, wx, wy, cw0, rw);
    const bool hit = pair_check<G, BOX>(ln.lane, ln.m, part, mx, my, cm0, cm1, rm, a.quirks != 0, a.quirk_rsum[0]);
    if (need) {
        wc = (__ballot_sync(FULL, bad) & ln.gmask) != 0u;
        mc = (__ballot_sync(FULL, hit) & ln.gmask) != 0u;
    } else {
        (void)__ballot_sync(FULL, bad);
        (void)__ballot_sync(FULL, hit);
    }
}

// reward / terminated / is_success for one env from the group reductions (plan:502-534, 459-479, 596-601)
__device__ __forceinline__ void planning_reward(int N, int reached, bool mc, bool wc, float& reward, bool& term,
                                                bool& succ) {
    const bool coll = mc || wc;
    const bool all = reached == N;
    reward = coll ? -50.0f : (all ? 50.0f : -(float)(N - reached));
    term = coll || all;
    succ = all && !coll;
}

template <int G>
__device__ __forceinline__ Lane<G> make_lane(const PlanArgs& a) {
    Lane<G> ln;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    ln.lane = threadIdx.x & 31u;
    ln.gmask = group_mask<G>(ln.lane);
    ln.env = (int)(gtid / G);
    ln.m = (int)(gtid % G);
    ln.env_ok = ln.env < a.B;
    ln.active = ln.env_ok && ln.m < a.N;
    ln.idx = (size_t)ln.env * (size_t)a.N + (size_t)ln.m;
    ln.env_global = a.env_base + (uint32_t)ln.env;
    return ln;
}

template <int G, bool BOX, bool NOISE>
__global__ void __launch_bounds__(256) planning_step_kernel(const PlanArgs a) {
    __shared__ Tables tb;
    load_tables(tb, a.L);
    __syncthreads();
    const Lane<G> ln = make_lane<G>(a);
    const int mm = ln.active ? ln.m : 0;
    const double cw0 = a.c_wall[mm * 2 + 0], cw1 = a.c_wall[mm * 2 + 1];    // safety = 0 (basic:1888-1901)
    const double cm0 = a.c_mover[mm * 2 + 0], cm1 = a.c_mover[mm * 2 + 1];

    double2 p = make_double2(0, 0), v = p, acc = p, goal = p;
    double2 u = p;
    uint32_t event = 0;
    int elapsed = 0;
    bool pending_reset = false;
    if (ln.env_ok) {
        event = a.rng[ln.env];
        elapsed = a.elapsed[ln.env];
        if (a.autoreset == GPR_AUTORESET_NEXT_STEP) pending_reset = a.needs_reset[ln.env] != 0;
    }
    if (ln.active) {
        p = a.pos[ln.idx];
        v = a.vel[ln.idx];
        acc = a.acc[ln.idx];
        goal = a.goal[ln.idx];
        const float2 af = a.action[ln.idx];
        // basic:1869-1873 clip to the action Box
        u.x = fmin(fmax((double)af.x, -a.act_lim), a.act_lim);
        u.y = fmin(fmax((double)af.y, -a.act_lim), a.act_lim);
    }

    // ------------------------------------------------------------------ the 40-cycle loop (basic:1879-1905)
    bool alive = ln.env_ok && !pending_reset;
    bool mc = false, wc = false;
    Rect rw, rm;
    for (int cyc = 0; cyc < a.num_cycles; ++cyc) {
        if (!__any_sync(FULL, alive)) break;
        float n4[4] = {0.f, 0.f, 0.f, 0.f};
        if (NOISE) gpr_normal4(a.seed, ln.env_global, event, (uint32_t)cyc * 4u + GPR_RNG_BLOCK_VEL_WALL, (uint32_t)ln.m, n4);
        if (alive && ln.active) {
            // plan:420-450 _mujoco_step_callback
            double velx = v.x, vely = v.y;
            if (NOISE) {
                velx = dadd(v.x, dmul((double)n4[0], a.sigma_v));  // plan:430
                vely = dadd(v.y, dmul((double)n4[1], a.sigma_v));
            }
            double t0, t1, ax, ay;
            if (a.learn_jerk) {
                double atx, aty, jx, jy;
                ensure_max(acc.x, acc.y, a.a_max, a.a_max2_lo, u.x, u.y, a.dt, atx, aty, jx, jy);  // plan:434
                ensure_max(velx, vely, a.v_max, a.v_max2_lo, atx, aty, a.dt, t0, t1, ax, ay);      // plan:437
                if (atx != ax || aty != ay) {                                                      // plan:438
                    jx = ddiv(dsub(ax, acc.x), a.dt);
                    jy = ddiv(dsub(ay, acc.y), a.dt);
                }
                // mj_step, integrator actuator with actearly (plan:305-311): act += dt*ctrl; qacc = act
                acc.x = dadd(acc.x, dmul(a.dt, jx));
                acc.y = dadd(acc.y, dmul(a.dt, jy));
            } else {
                ensure_max(velx, vely, a.v_max, a.v_max2_lo, u.x, u.y, a.dt, t0, t1, ax, ay);  // plan:442
                acc.x = ax;  // dyntype none, gain = mass (plan:314-320): qacc = ctrl
                acc.y = ay;
            }
            // semi-implicit Euler (MuJoCo): qvel += dt*qacc; qpos += dt*qvel
            v.x = dadd(v.x, dmul(a.dt, acc.x));
            v.y = dadd(v.y, dmul(a.dt, acc.y));
            p.x = dadd(p.x, dmul(a.dt, v.x));
            p.y = dadd(p.y, dmul(a.dt, v.y));
        }
        // basic:1888-1894 wall check (noisy qpos, no safety offset)
        double wx = p.x, wy = p.y, mx = p.x, my = p.y;
        if (NOISE) {
            wx = dadd(p.x, dmul((double)n4[2], a.sigma_p));
            wy = dadd(p.y, dmul((double)n4[3], a.sigma_p));
            if (G > 1) {
                float k4[4];
                gpr_normal4(a.seed, ln.env_global, event, (uint32_t)cyc * 4u + GPR_RNG_BLOCK_MOVER, (uint32_t)ln.m, k4);
                mx = dadd(p.x, dmul((double)k4[0], a.sigma_p));
                my = dadd(p.y, dmul((double)k4[1], a.sigma_p));
            }
            if (BOX) {
                float q[4];
                gpr_normal4(a.seed, ln.env_global, event, (uint32_t)cyc * 4u + GPR_RNG_BLOCK_WALL_QUAT, (uint32_t)ln.m, q);
                rect_vertices(wx, wy, dadd(1.0, dmul((double)q[0], a.sigma_p)), dmul((double)q[1], a.sigma_p),
                              dmul((double)q[2], a.sigma_p), dmul((double)q[3], a.sigma_p), cw0, cw1, rw);
                if (G > 1) {
                    gpr_normal4(a.seed, ln.env_global, event, (uint32_t)cyc * 4u + GPR_RNG_BLOCK_MOVER_QUAT, (uint32_t)ln.m, q);
                    rect_vertices(mx, my, dadd(1.0, dmul((double)q[0], a.sigma_p)), dmul((double)q[1], a.sigma_p),
                                  dmul((double)q[2], a.sigma_p), dmul((double)q[3], a.sigma_p), cm0, cm1, rm);
                }
            }
        } else if (BOX) {
            rect_vertices_axis(wx, wy, cw0, cw1, rw);
            rect_vertices_axis(mx, my, cm0, cm1, rm);
        }
        const bool part = alive && ln.active;
        const bool bad = part && !wall_valid<BOX>(tb, a.L, wx, wy, cw0, rw);
        // basic:1895-1901 mover check (independent noisy qpos)
        const bool hit = pair_check<G, BOX>(ln.lane, ln.m, part, mx, my, cm0, cm1, rm, a.quirks != 0, a.quirk_rsum[0]);
        const bool wnow = (__ballot_sync(FULL, bad) & ln.gmask) != 0u;
        const bool mnow = (__ballot_sync(FULL, hit) & ln.gmask) != 0u;
        if (alive) {
            wc = wnow;
            mc = mnow;
            if (wc || mc) alive = false;  // basic:1904 break
        }
    }

    // ------------------------------------------------------------------ observation, info, reward (basic:1910-1929)
    const bool stepped = ln.env_ok && !pending_reset;
    double2 ag, ov;
    int reached;
    observe<G, NOISE>(a, ln, event, p, v, goal, ag, ov, reached);
    float reward;
    bool term, succ;
    planning_reward(a.N, reached, mc, wc, reward, term, succ);
    if (stepped) {
        event += 1u;
        elapsed += 1;
    }
    const bool trunc = stepped && a.max_episode_steps > 0 && elapsed >= a.max_episode_steps;  // gymnasium TimeLimit
    const bool done = stepped && (term || trunc);

    // episode statistics (one lane per env accumulates, one atomic per warp and counter)
    {
        const bool lead = ln.env_ok && ln.m == 0;
        float ret = 0.f;
        if (lead && stepped) ret = a.ep_return[ln.env] + reward;
        const bool fin = lead && done;
        double s_ep = fin ? 1.0 : 0.0, s_ret = fin ? (double)ret : 0.0, s_len = fin ? (double)elapsed : 0.0;
        double s_succ = (fin && succ) ? 1.0 : 0.0, s_mc = (fin && mc) ? 1.0 : 0.0, s_wc = (fin && wc) ? 1.0 : 0.0;
        if (__any_sync(FULL, fin)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s_ep += __shfl_xor_sync(FULL, s_ep, o);
                s_ret += __shfl_xor_sync(FULL, s_ret, o);
                s_len += __shfl_xor_sync(FULL, s_len, o);
                s_succ += __shfl_xor_sync(FULL, s_succ, o);
                s_mc += __shfl_xor_sync(FULL, s_mc, o);
                s_wc += __shfl_xor_sync(FULL, s_wc, o);
            }
            if (ln.lane == 0) {
                atomicAdd(a.stats + 0, s_ep);
                atomicAdd(a.stats + 1, s_ret);
                atomicAdd(a.stats + 2, s_len);
                atomicAdd(a.stats + 3, s_succ);
                atomicAdd(a.stats + 4, s_mc);
                atomicAdd(a.stats + 5, s_wc);
            }
        }
        if (lead && stepped) a.ep_return[ln.env] = done ? 0.f : ret;
    }

    // per-env scalars of the finished transition
    if (ln.env_ok && ln.m == 0 && stepped) {
        if (a.out.reward) a.out.reward[ln.env] = reward;
        if (a.out.terminated) a.out.terminated[ln.env] = term;
        if (a.out.truncated) a.out.truncated[ln.env] = trunc;
        if (a.out.is_success) a.out.is_success[ln.env] = succ;
        if (a.out.mover_collision) a.out.mover_collision[ln.env] = mc;
        if (a.out.wall_collision) a.out.wall_collision[ln.env] = wc;
    }

    // ------------------------------------------------------------------ auto-reset
    bool need = false;
    if (a.autoreset == GPR_AUTORESET_SAME_STEP) need = done;
    if (a.autoreset == GPR_AUTORESET_NEXT_STEP) need = ln.env_ok && pending_reset;
    if (__any_sync(FULL, need)) {
        if (need && a.autoreset == GPR_AUTORESET_SAME_STEP)
            store_obs<G>(a, ln, a.out.final_observation, a.out.final_achieved_goal, a.out.final_desired_goal, ov, acc, ag, goal);
        bool rmc = false, rwc = false, failed = false;
        reset_group<G, BOX, NOISE>(a, tb, ln, need, event, nullptr, nullptr, p, v, acc, goal, rmc, rwc, failed);
        double2 ag2, ov2;
        int reached2;
        observe<G, NOISE>(a, ln, event, p, v, goal, ag2, ov2, reached2);
        if (need) {
            ag = ag2;
            ov = ov2;
            event += 1u;
            elapsed = 0;
            if (failed && ln.m == 0) atomicAdd(a.fail_count, 1u);
            if (a.autoreset == GPR_AUTORESET_NEXT_STEP && ln.m == 0) {
                // gymnasium NEXT_STEP: this call only resets; reward 0, not done; info of the fresh episode
                float r2;
                bool t2, s2;
                planning_reward(a.N, reached2, rmc, rwc, r2, t2, s2);
                if (a.out.reward) a.out.reward[ln.env] = 0.f;
                if (a.out.terminated) a.out.terminated[ln.env] = 0;
                if (a.out.truncated) a.out.truncated[ln.env] = 0;
                if (a.out.is_success) a.out.is_success[ln.env] = s2;
                if (a.out.mover_collision) a.out.mover_collision[ln.env] = rmc;
                if (a.out.wall_collision) a.out.wall_collision[ln.env] = rwc;
            }
        }
    }
    store_obs<G>(a, ln, a.out.observation, a.out.achieved_goal, a.out.desired_goal, ov, acc, ag, goal);

    // ------------------------------------------------------------------ state write-back
    if (ln.active) {
        a.pos[ln.idx] = p;
        a.vel[ln.idx] = v;
        a.acc[ln.idx] = acc;
        if (need) a.goal[ln.idx] = goal;
    }
    if (ln.env_ok && ln.m == 0) {
        a.rng[ln.env] = event;
        a.elapsed[ln.env] = elapsed;
        if (a.autoreset == GPR_AUTORESET_NEXT_STEP) a.needs_reset[ln.env] = (done && !need) ? 1 : 0;
    }
}

template <int G, bool BOX, bool NOISE>
__global__ void __launch_bounds__(256) planning_reset_kernel(const PlanArgs a) {
    __shared__ Tables tb;
    load_tables(tb, a.L);
    __syncthreads();
    const Lane<G> ln = make_lane<G>(a);
    const bool need = ln.env_ok && (a.reset_mask == nullptr || a.reset_mask[ln.env] != 0);
    double2 p = make_double2(0, 0), v = p, acc = p, goal = p;
    uint32_t event = 0;
    if (ln.env_ok) event = a.rng[ln.env];
    if (ln.active && !need) {  // keep registers meaningful for non-reset envs (they are not written back)
        p = a.pos[ln.idx];
        goal = a.goal[ln.idx];
    }
    bool mc = false, wc = false, failed = false;
    reset_group<G, BOX, NOISE>(a, tb, ln, need, event, a.inject_start, a.inject_goal, p, v, acc, goal, mc, wc, failed);
    double2 ag, ov;
    int reached;
    observe<G, NOISE>(a, ln, event, p, v, goal, ag, ov, reached);
    if (!need) return;
    store_obs<G>(a, ln, a.out.observation, a.out.achieved_goal, a.out.desired_goal, ov, acc, ag, goal);
    if (ln.active) {
        a.pos[ln.idx] = p;
        a.vel[ln.idx] = v;
        a.acc[ln.idx] = acc;
        a.goal[ln.idx] = goal;
    }
    if (ln.m == 0) {
        float r;
        bool t, s;
        planning_reward(a.N, reached, mc, wc, r, t, s);
        if (a.out.is_success) a.out.is_success[ln.env] = s;
        if (a.out.mover_collision) a.out.mover_collision[ln.env] = mc;
        if (a.out.wall_collision) a.out.wall_collision[ln.env] = wc;
        a.rng[ln.env] = event + 1u;
        a.elapsed[ln.env] = 0;
        a.ep_return[ln.env] = 0.f;
        if (a.needs_reset) a.needs_reset[ln.env] = 0;
        if (failed) atomicAdd(a.fail_count, 1u);
    }
}

// HER relabelling (plan:502-534 / push:499-527 on arrays): one thread per transition.
__global__ void compute_reward_kernel(int kind, int N, int batch, double threshold, const float* __restrict__ achieved,
                                      const float* __restrict__ desired, const uint8_t* __restrict__ mcol,
                                      const uint8_t* __restrict__ wcol, float* __restrict__ reward,
                                      uint8_t* __restrict__ terminated) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const bool mc = mcol ? mcol[b] != 0 : false, wc = wcol ? wcol[b] != 0 : false;
    if (kind == GPR_ENV_PLANNING) {
        int reached = 0;
        const float2* ag = reinterpret_cast<const float2*>(achieved) + (size_t)b * N;
        const float2* dg = reinterpret_cast<const float2*>(desired) + (size_t)b * N;
        for (int m = 0; m < N; ++m) {
            const float2 a2 = ag[m], d2 = dg[m];
            const double dx = dsub((double)a2.x, (double)d2.x), dy = dsub((double)a2.y, (double)d2.y);
            reached += sqrt_le(dadd(dmul(dx, dx), dmul(dy, dy)), threshold) ? 1 : 0;
        }
        float r;
        bool t, s;
        planning_reward(N, reached, mc, wc, r, t, s);
        if (reward) reward[b] = r;
        if (terminated) terminated[b] = t;
    } else {
        const float2 a2 = reinterpret_cast<const float2*>(achieved)[b], d2 = reinterpret_cast<const float2*>(desired)[b];
        const double dx = dsub((double)a2.x, (double)d2.x), dy = dsub((double)a2.y, (double)d2.y);
        const bool reached = sqrt_le(dadd(dmul(dx, dx), dmul(dy, dy)), threshold);
        const float r = wc ? -50.f : (reached ? 0.f : -1.f);  // push:521-523
        if (reward) reward[b] = r;
        if (terminated) terminated[b] = wc;  // push:475
    }
}

}  // namespace gpr
