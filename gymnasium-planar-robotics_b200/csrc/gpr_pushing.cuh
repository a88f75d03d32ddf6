// gpr_pushing.cuh — fused kernels of BenchmarkPushingEnv's step path: one lane per environment (one mover + one object).
//
//   pushing_step_kernel : basic:1835-1950 — action clip, num_cycles x { push:419-455 limit control + yaw impedance,
//                         planar mj_step substitute with mover-object box contact and object-ground friction
//                         (include/gpr_push_physics.h), basic:459-788 wall check, break on wall collision },
//                         push:529-576 observation, push:578-608 info, push:499-527 reward, push:457-476 terminated,
//                         TimeLimit, episode statistics, auto-reset (push:373-417).
//   pushing_reset_kernel: basic:1770-1833 for a masked subset with optional injected positions.
//
// This translation unit is compiled with -fmad=false: the shared physics header uses plain float64 expressions and must
// round exactly like the CPU oracle (gcc -ffp-contract=off).
#pragma once

#include "../../include/gpr_push_physics.h"
#include "gpr_device.cuh"
#include "gpr_planning.cuh"  // wall_core / wall_fast / guess_cell / normal4_cold

namespace gpr {

struct PushArgs {
    int B;
    int learn_jerk, num_cycles, max_episode_steps, autoreset, max_reset_attempts;
    uint32_t env_base;
    uint64_t seed;
    double dt, inv_dt, v_max, a_max, j_max, act_lim;
    double v_max2_lo, a_max2_lo;
    double threshold;
    double min_xy[2], span_xy[2];          // mover spawn box (push:250-255)
    double obj_min[2], obj_span[2];        // object / goal box (push:257-258)
    double min_mo_dist;                    // push:279-288, strict '>' accepts
    double sigma_p, sigma_v, sigma_obj;    // sensor noise; object position noise 1e-5 (push:178)
    double c_wall[2][2];                   // [safety][xy] mover collision size incl. offsets (basic:487)
    double v_lazy2;                        // (v_max - velocity-noise bound)^2: below it neither clip nor noise matters
    float wxf, wyf, wall_delta, inv_dtf;   // float32 wall screen (see wall_core) and travel budget, as in PlanArgs
    LayoutArgs L;
    gpr_push_params P;
    // state (SoA, float64)
    double2* pos;      // mover x, y
    double2* vel;
    double2* acc;      // MuJoCo qacc of the mover (differs from `act` under contact)
    double2* act;      // jerk-mode integrator state
    double* mover_rot;  // [B,3] cos yaw, sin yaw, yaw rate
    double* obj_pos;    // [B,4] x, y, cos yaw, sin yaw
    double* obj_vel;    // [B,3] vx, vy, yaw rate
    float* warm;        // [B,GPR_PUSH_WARM] warm-start state of the contact solve (zero for every env in the free regime)
    double2* goal;     // object goal
    int32_t* elapsed;
    uint32_t* rng;
    uint8_t* needs_reset;
    float* ep_return;
    double* stats;
    uint32_t* fail_count;
    uint32_t* debug_errors;  // [DBG_NUM_SLOTS] GPR_DEBUG_BOUNDS builds only (see gpr_device.cuh)
    int write_goal;  // see PlanArgs
    int out_f64;     // GPR_OUT_FLOAT64
    // work queue of the envs that entered the contact regime (see pushing_step_kernel), double-buffered by step parity
    unsigned long long* queue;      // [B] (cycle << 32 | env)
    unsigned long long* queue_ctl;  // [2] entries appended
    uint32_t* queue_cursor;         // [2] entries claimed by pushing_contact_kernel
    int parity;
    // per-call I/O
    const float2* action;
    gpr_outputs out;
    const uint8_t* reset_mask;
    const double2* inject_start;
    const double2* inject_goal;
    const double2* inject_object;
};

struct PushState {
    gpr_body2 M, O;
    double2 acc, act, goal;
};

__device__ __forceinline__ void push_load(const PushArgs& a, int e, PushState& s) {
    const double2 p = a.pos[e], v = a.vel[e];
    s.M.x = p.x;
    s.M.y = p.y;
    s.M.vx = v.x;
    s.M.vy = v.y;
    s.M.c = a.mover_rot[3 * e + 0];
    s.M.s = a.mover_rot[3 * e + 1];
    s.M.w = a.mover_rot[3 * e + 2];
    s.O.x = a.obj_pos[4 * e + 0];
    s.O.y = a.obj_pos[4 * e + 1];
    s.O.c = a.obj_pos[4 * e + 2];
    s.O.s = a.obj_pos[4 * e + 3];
    s.O.vx = a.obj_vel[3 * e + 0];
    s.O.vy = a.obj_vel[3 * e + 1];
    s.O.w = a.obj_vel[3 * e + 2];
    s.acc = a.acc[e];
    s.act = a.act[e];
    s.goal = a.goal[e];
}

__device__ __forceinline__ void push_store(const PushArgs& a, int e, const PushState& s) {
    a.pos[e] = make_double2(s.M.x, s.M.y);
    a.vel[e] = make_double2(s.M.vx, s.M.vy);
    a.mover_rot[3 * e + 0] = s.M.c;
    a.mover_rot[3 * e + 1] = s.M.s;
    a.mover_rot[3 * e + 2] = s.M.w;
    a.obj_pos[4 * e + 0] = s.O.x;
    a.obj_pos[4 * e + 1] = s.O.y;
    a.obj_pos[4 * e + 2] = s.O.c;
    a.obj_pos[4 * e + 3] = s.O.s;
    a.obj_vel[3 * e + 0] = s.O.vx;
    a.obj_vel[3 * e + 1] = s.O.vy;
    a.obj_vel[3 * e + 2] = s.O.w;
    a.acc[e] = s.acc;
    a.act[e] = s.act;
    a.goal[e] = s.goal;
}

// basic:1888-1894 / 1799-1801 for the single mover: noisy qpos (position and, for the box shape, the yaw quaternion)
template <bool BOX, bool NOISE>
__device__ __forceinline__ bool push_wall_bad(const PushArgs& a, const Tables& tb, const gpr_body2& M, int safety,
                                              float nx, float ny, uint32_t env_global, uint32_t event, uint32_t qstream) {
    double wx = M.x, wy = M.y;
    if (NOISE) {
        wx = dadd(M.x, dmul((double)nx, a.sigma_p));
        wy = dadd(M.y, dmul((double)ny, a.sigma_p));
    }
    const double c0 = a.c_wall[safety][0], c1 = a.c_wall[safety][1];
    Rect r;
    if (BOX) {
        // yaw quaternion (cos(yaw/2), 0, 0, sin(yaw/2)) from (cos yaw, sin yaw) by the half-angle identities
        const double ch = dsqrt(dmul(0.5, dadd(1.0, M.c)));
        const double sh = ddiv(M.s, dmul(2.0, ch));
        double qw = ch, qx = 0.0, qy = 0.0, qz = sh;
        if (NOISE) {
            float q[4];
            gpr_normal4(a.seed, env_global, event, qstream, 0u, q);
            qw = dadd(qw, dmul((double)q[0], a.sigma_p));
            qx = dmul((double)q[1], a.sigma_p);
            qy = dmul((double)q[2], a.sigma_p);
            qz = dadd(qz, dmul((double)q[3], a.sigma_p));
        }
        rect_vertices(wx, wy, qw, qx, qy, qz, c0, c1, r);
    }
    return !wall_valid<BOX>(tb, a.L, wx, wy, c0, r);
}

// push:529-576 observation + push:499-527 / 457-476 / 578-608 reward, terminated, is_success
template <bool NOISE>
__device__ __forceinline__ void push_observe(const PushArgs& a, const PushState& s, uint32_t env_global, uint32_t event,
                                             double (&obs)[6], double2& ag, bool& reached) {
    double px = s.M.x, py = s.M.y, vx = s.M.vx, vy = s.M.vy;
    if (NOISE) {
        float n4[4];
        gpr_normal4(a.seed, env_global, event, GPR_RNG_OBS, 0u, n4);
        px = dadd(px, dmul((double)n4[0], a.sigma_p));
        py = dadd(py, dmul((double)n4[1], a.sigma_p));
        vx = dadd(vx, dmul((double)n4[2], a.sigma_v));
        vy = dadd(vy, dmul((double)n4[3], a.sigma_v));
    }
    obs[0] = px;
    obs[1] = py;
    obs[2] = vx;
    obs[3] = vy;
    obs[4] = s.acc.x;  // qacc, no noise (push:556)
    obs[5] = s.acc.y;
    ag = make_double2(s.O.x, s.O.y);
    if (a.sigma_obj != 0.0) {  // push:565: always-on object position noise
        float k4[4];
        gpr_normal4(a.seed, env_global, event, GPR_RNG_OBJECT, 0u, k4);
        ag.x = dadd(ag.x, dmul((double)k4[0], a.sigma_obj));
        ag.y = dadd(ag.y, dmul((double)k4[1], a.sigma_obj));
    }
    const double dx = dsub(ag.x, s.goal.x), dy = dsub(ag.y, s.goal.y);
    reached = sqrt_le(dadd(dmul(dx, dx), dmul(dy, dy)), a.threshold);  // push:519
}

__device__ __forceinline__ void push_store_obs(const PushArgs& a, int e, float* O, float* AG, float* DG, const double (&obs)[6],
                                               double2 ag, double2 goal) {
    if (O) {
        const size_t rowp = (size_t)e * (size_t)(2 + a.learn_jerk);  // in (x, y) pairs
        if (!a.out_f64 && !a.learn_jerk && (reinterpret_cast<uintptr_t>(O) & 15u) == 0u) {
            // the 16-byte row in ONE store: two 8-byte stores per lane leave every 32-byte sector half written twice, and
            // on the host route (rows in page-locked host memory) each partial write is a PCIe transaction of its own
            reinterpret_cast<float4*>(O)[e] = make_float4((float)obs[0], (float)obs[1], (float)obs[2], (float)obs[3]);
        } else {
            store_pair(a.out_f64 != 0, O, rowp + 0, obs[0], obs[1]);
            store_pair(a.out_f64 != 0, O, rowp + 1, obs[2], obs[3]);
            if (a.learn_jerk) store_pair(a.out_f64 != 0, O, rowp + 2, obs[4], obs[5]);
        }
    }
    if (AG) store_pair(a.out_f64 != 0, AG, (size_t)e, ag.x, ag.y);
    if (DG) store_pair(a.out_f64 != 0, DG, (size_t)e, goal.x, goal.y);
}

// push:373-417 + basic:1797-1805, split in three stages so that the object-placement loop (push:392-407) can be shared by
// the warp: acceptance is ~75% per draw on average, but a mover spawned near the layout centre leaves only a sliver (or
// nothing) of the object box outside min_mo_dist, and one lane looping thousands of draws would stall its whole warp.
//   stage A (lane-local)      : mover start, the first kSeqDraws object draws
//   stage B (warp-collective) : for every lane still without an accepted draw, all 32 lanes test 32 consecutive draws at
//                               once; the lowest accepted index wins — exactly the sequential loop's result
//   stage C (lane-local)      : goal, fresh MjData, reset-time wall check
constexpr int kSeqDraws = 4;

__device__ __forceinline__ bool push_object_draw(const PushArgs& a, uint32_t env_global, uint32_t event, int t, double mx,
                                                 double my, double& ox, double& oy) {
    double ux, uy;
    gpr_sample_xy(a.seed, env_global, event, GPR_RNG_RESET_OBJECT, 0u, (uint32_t)t, 0u, &ux, &uy);
    ox = dadd(a.obj_min[0], dmul(a.obj_span[0], ux));
    oy = dadd(a.obj_min[1], dmul(a.obj_span[1], uy));
    const double dx = dsub(ox, mx), dy = dsub(oy, my);
    return !sqrt_le(dadd(dmul(dx, dx), dmul(dy, dy)), a.min_mo_dist);  // push:407 strict '>'
}

// stage A.  Returns true when the object position is settled (injected or accepted).
__device__ __forceinline__ bool push_reset_a(const PushArgs& a, PushState& s, uint32_t env_global, uint32_t event,
                                             const double2* inj_start, const double2* inj_object, int e) {
    if (inj_start) {
        s.M.x = inj_start[e].x;
        s.M.y = inj_start[e].y;
    } else {  // push:387-389
        double ux, uy;
        gpr_sample_xy(a.seed, env_global, event, GPR_RNG_RESET_SAMPLE, 0u, 0u, 0u, &ux, &uy);
        s.M.x = dadd(a.min_xy[0], dmul(a.span_xy[0], ux));
        s.M.y = dadd(a.min_xy[1], dmul(a.span_xy[1], uy));
    }
    if (inj_object) {
        s.O.x = inj_object[e].x;
        s.O.y = inj_object[e].y;
        return true;
    }
    const int cap = a.max_reset_attempts > 0 ? a.max_reset_attempts : 1;
    bool ok = false;
    for (int t = 0; t < kSeqDraws && t < cap && !ok; ++t) ok = push_object_draw(a, env_global, event, t, s.M.x, s.M.y, s.O.x, s.O.y);
    return ok;
}

// stage B, warp-collective: EVERY lane of the warp must call it.  pend: this lane still needs an object position.
// Returns true (for a pending lane) if the loop ran out of attempts (the last draw is kept, like the oracle does).
__device__ __forceinline__ bool push_reset_b(const PushArgs& a, PushState& s, uint32_t env_global, uint32_t event, bool pend) {
    const unsigned lane = threadIdx.x & 31u;
    const int cap = a.max_reset_attempts > 0 ? a.max_reset_attempts : 1;
    bool failed = false;
    unsigned todo = __ballot_sync(FULL, pend);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t eg = __shfl_sync(FULL, env_global, leader), ev = __shfl_sync(FULL, event, leader);
        const double mx = __shfl_sync(FULL, s.M.x, leader), my = __shfl_sync(FULL, s.M.y, leader);
        bool found = false;
        double wx = 0.0, wy = 0.0;
        for (int t0 = kSeqDraws; t0 < cap && !found; t0 += 32) {
            const int t = t0 + (int)lane;
            double ox, oy;
            const bool acc = t < cap && push_object_draw(a, eg, ev, t, mx, my, ox, oy);
            const unsigned accm = __ballot_sync(FULL, acc);
            if (accm) {
                const int win = __ffs(accm) - 1;
                wx = __shfl_sync(FULL, ox, win);
                wy = __shfl_sync(FULL, oy, win);
                found = true;
            }
        }
        if (!found) push_object_draw(a, eg, ev, cap - 1, mx, my, wx, wy);  // keep the last draw (uniform across lanes)
        if ((int)lane == leader) {
            s.O.x = wx;
            s.O.y = wy;
            failed = !found;
        }
    }
    return failed;
}

// stage C
template <bool BOX, bool NOISE>
__device__ __forceinline__ void push_reset_c(const PushArgs& a, const Tables& tb, PushState& s, uint32_t env_global,
                                             uint32_t event, const double2* inj_goal, int e, bool& wc) {
    if (inj_goal) {
        s.goal = inj_goal[e];
    } else {  // push:409-411
        double ux, uy;
        gpr_sample_xy(a.seed, env_global, event, GPR_RNG_RESET_OBJECT, 1u, 0u, 0u, &ux, &uy);
        s.goal.x = dadd(a.obj_min[0], dmul(a.obj_span[0], ux));
        s.goal.y = dadd(a.obj_min[1], dmul(a.obj_span[1], uy));
    }
    // reload_model (push:353-371): fresh MjData, everything at rest, identity orientations
    s.M.vx = s.M.vy = s.M.w = 0.0;
    s.M.c = 1.0;
    s.M.s = 0.0;
    s.O.vx = s.O.vy = s.O.w = 0.0;
    s.O.c = 1.0;
    s.O.s = 0.0;
    s.acc = make_double2(0.0, 0.0);
    s.act = make_double2(0.0, 0.0);
    for (int i = 0; i < GPR_PUSH_WARM; ++i) a.warm[(size_t)e * GPR_PUSH_WARM + i] = 0.f;  // fresh MjData: no warm-start forces
    // basic:1799-1801 wall check with the safety offset on noisy qpos
    float n4[4] = {0.f, 0.f, 0.f, 0.f};
    if (NOISE) gpr_normal4(a.seed, env_global, event, GPR_RNG_RESET_CHECK, 0u, n4);
    wc = push_wall_bad<BOX, NOISE>(a, tb, s.M, 1, n4[0], n4[1], env_global, event, GPR_RNG_RESET_CHECK_WQUAT);
}

__device__ __forceinline__ void push_reward(bool reached, bool wc, float& reward, bool& term, bool& succ) {
    reward = wc ? -50.f : (reached ? 0.f : -1.f);  // push:521-523
    term = wc;                                      // push:475
    succ = reached && !wc;                          // push:602
}

// ---- the env-step kernels -------------------------------------------------------------------------------------------------
// Per cycle an environment is either FREE (object at rest and out of the mover's reach: gpr_push_substep_free, ~60
// instructions) or in the CONTACT regime (manifold + projected Gauss-Seidel sweeps: thousands of dependent float64
// instructions), and an object that has been touched keeps sliding for the rest of its episode.  With random actions about a
// fifth of the environments are in the contact regime at any time — nearly every warp would run the long path with a handful
// of active lanes.  The step is therefore split BY POPULATION into two launches (no CTA barrier anywhere):
//
//   pushing_step_kernel     one lane per env.  Runs the cycles of the free regime; an env that meets the contact condition
//                           at the top of cycle c (gpr_push_is_free, the very test gpr_push_substep makes) is parked: its
//                           state as of that cycle goes back to the state arrays and (c, env) is appended to a work queue.
//                           Every other env finishes here: observation, reward, flags, statistics, auto-reset.
//   pushing_contact_kernel  persistent warps pull queued envs 32 at a time and run their remaining cycles with the general
//                           substep, every lane in the long path, then finish them with the same code.
//
// Both kernels evaluate the same functions on the same values in the same order as the one-kernel formulation (and as the
// oracle): which kernel runs a cycle changes no result bit.
constexpr int kPushCta = 128;

#ifndef GPR_PUSH_MINB
#define GPR_PUSH_MINB 4
#endif

// push:419-455 _mujoco_step_callback for one cycle: returns the commanded acceleration (cx, cy) — the limited acceleration, or
// the integrator state `act` in jerk mode — and updates s.act.  n4/have0: this cycle's noise block 0, generated on demand.
template <bool BOX, bool NOISE>
__device__ __forceinline__ void push_control(const PushArgs& a, PushState& s, double ux, double uy, uint32_t env_global,
                                             uint32_t event, uint32_t s0, float (&n4)[4], bool& have0, double& cx, double& cy) {
    double dxv = ux, dyv = uy, jx = 0.0, jy = 0.0;
    if (a.learn_jerk) ensure_max(s.acc.x, s.acc.y, a.a_max, a.a_max2_lo, ux, uy, a.dt, a.inv_dt, dxv, dyv, jx, jy);  // push:432 (real qacc)
    double velx = s.M.vx, vely = s.M.vy;
    if (NOISE) {
        // the velocity noise (push:428) is generated only where it can matter
        const double sx = dadd(dmul(a.dt, dxv), s.M.vx), sy = dadd(dmul(a.dt, dyv), s.M.vy);
        if (BOX || !(dadd(dmul(sx, sx), dmul(sy, sy)) < a.v_lazy2)) {
            normal4_cold(a.seed, env_global, event, s0 + GPR_RNG_BLOCK_VEL_WALL, 0u, n4);
            have0 = true;
            velx = dadd(velx, dmul((double)n4[0], a.sigma_v));
            vely = dadd(vely, dmul((double)n4[1], a.sigma_v));
        }
    }
    double t0, t1, ax, ay;
    ensure_max(velx, vely, a.v_max, a.v_max2_lo, dxv, dyv, a.dt, a.inv_dt, t0, t1, ax, ay);  // push:435 / 440
    if (a.learn_jerk) {
        if (dxv != ax || dyv != ay) {  // push:436
            jx = ddiv_rcp(dsub(ax, s.acc.x), a.dt, a.inv_dt);
            jy = ddiv_rcp(dsub(ay, s.acc.y), a.dt, a.inv_dt);
        }
        // integrator actuator with actearly (push:305-311): act += dt*ctrl, force uses the new act
        s.act.x = dadd(s.act.x, dmul(a.dt, jx));
        s.act.y = dadd(s.act.y, dmul(a.dt, jy));
        cx = s.act.x;
        cy = s.act.y;
    } else {
        cx = ax;
        cy = ay;
    }
}

// basic:1888-1894 wall check of one cycle (no mover-mover check with one mover; push:607 asserts no mover collision).
// Circle shape: with a travel budget exactly as in planning_step_kernel (certified clearance, wall_core).
template <bool BOX, bool NOISE>
__device__ __forceinline__ bool push_wall_cycle(const PushArgs& a, const Tables& tb, const PushState& s, uint32_t env_global,
                                                uint32_t event, uint32_t s0, float (&n4)[4], bool have0, float cwf, int& gi,
                                                int& gj, float& travel, float& lim_w) {
    bool wc = false;
    if (BOX) {
        if (NOISE && !have0) normal4_cold(a.seed, env_global, event, s0 + GPR_RNG_BLOCK_VEL_WALL, 0u, n4);
        wc = push_wall_bad<BOX, NOISE>(a, tb, s.M, 0, n4[2], n4[3], env_global, event, s0 + GPR_RNG_BLOCK_WALL_QUAT);
    } else {
        const float avx = fabsf((float)s.M.vx), avy = fabsf((float)s.M.vy);
        travel += (fmaxf(avx, avy) + 0.5f * fminf(avx, avy)) * 1.0001f;
        if (!(travel < lim_w)) {
            float clear_w;
            const int f = wall_fast(a, tb, s.M.x, s.M.y, cwf, cwf, gi, gj, clear_w);
            if (f == 2) {  // too close to call in float32: the exact check on the noisy position
                if (NOISE && !have0) normal4_cold(a.seed, env_global, event, s0 + GPR_RNG_BLOCK_VEL_WALL, 0u, n4);
                wc = push_wall_bad<BOX, NOISE>(a, tb, s.M, 0, n4[2], n4[3], env_global, event, s0 + GPR_RNG_BLOCK_WALL_QUAT);
                guess_cell(a, s.M.x, s.M.y, gi, gj);
            } else {
                wc = f == 0;
            }
            lim_w = (travel + clear_w * a.inv_dtf) * 0.999999f;
        }
    }
    return wc;
}

// Everything after the cycle loop for one lane's env (warp-collective: EVERY lane of the warp calls it; `live` = this lane
// holds an env that has finished its cycles here): observation, reward, flags, statistics, auto-reset, state write-back.
// `filler` = this lane's env was parked for pushing_contact_kernel, which will write its result rows later: the lane writes
// zeros into the dense rows now so that the warp's stores stay CONTIGUOUS.  Holes would split every store into partially
// written 32-byte sectors, and into page-locked host memory each fragment is a PCIe write of its own (measured,
// tools/sm_store_bw.cu: 50 GB/s for whole sectors, 0.8 G fragments/s otherwise — the fragments were 0.12 ms per step).
template <bool BOX, bool NOISE>
__device__ __forceinline__ void push_finish(const PushArgs& a, const Tables& tb, int e, bool live, bool filler, bool pending,
                                            PushState& s, uint32_t env_global, uint32_t event, int elapsed, bool wc) {
    double obs[6] = {0, 0, 0, 0, 0, 0};
    double2 ag = make_double2(0, 0);
    bool reached = false;
    float reward = 0.f;
    bool term = false, succ = false;
    const bool stepped = live && !pending;
    if (stepped) {
        push_observe<NOISE>(a, s, env_global, event, obs, ag, reached);
        push_reward(reached, wc, reward, term, succ);
        event += 1u;
        elapsed += 1;
    }
    const bool trunc = stepped && a.max_episode_steps > 0 && elapsed >= a.max_episode_steps;
    const bool done = stepped && (term || trunc);
    if (stepped) {
        const float ret = a.ep_return[e] + reward;
        if (done) {
            atomicAdd(a.stats + 0, 1.0);
            atomicAdd(a.stats + 1, (double)ret);
            atomicAdd(a.stats + 2, (double)elapsed);
            if (succ) atomicAdd(a.stats + 3, 1.0);
            if (wc) atomicAdd(a.stats + 5, 1.0);
        }
        a.ep_return[e] = done ? 0.f : ret;
    }
    if (stepped || filler) {  // (a filler lane holds reward 0 and no flag)
        if (a.out.reward) a.out.reward[e] = reward;
        if (a.out.terminated) a.out.terminated[e] = term;
        if (a.out.truncated) a.out.truncated[e] = trunc;
        if (a.out.is_success) a.out.is_success[e] = succ;
        if (a.out.mover_collision) a.out.mover_collision[e] = 0;
        if (a.out.wall_collision) a.out.wall_collision[e] = stepped && wc;
    }
    // ---- auto-reset (push:373-417), stages A / B / C
    const bool need = live && ((a.autoreset == GPR_AUTORESET_SAME_STEP && done) || pending);
    bool settled = true;
    if (need) {
        if (a.autoreset == GPR_AUTORESET_SAME_STEP)
            push_store_obs(a, e, a.out.final_observation, a.out.final_achieved_goal, a.out.final_desired_goal, obs, ag, s.goal);
        settled = push_reset_a(a, s, env_global, event, nullptr, nullptr, e);
    }
    const bool failed = push_reset_b(a, s, env_global, event, need && !settled);
    if (need) {
        bool rwc;
        push_reset_c<BOX, NOISE>(a, tb, s, env_global, event, nullptr, e, rwc);
        push_observe<NOISE>(a, s, env_global, event, obs, ag, reached);
        event += 1u;
        elapsed = 0;
        if (failed) atomicAdd(a.fail_count, 1u);
        if (pending) {  // gymnasium NEXT_STEP: this call only resets
            float r2;
            bool t2, s2;
            push_reward(reached, rwc, r2, t2, s2);
            if (a.out.reward) a.out.reward[e] = 0.f;
            if (a.out.terminated) a.out.terminated[e] = 0;
            if (a.out.truncated) a.out.truncated[e] = 0;
            if (a.out.is_success) a.out.is_success[e] = s2;
            if (a.out.mover_collision) a.out.mover_collision[e] = 0;
            if (a.out.wall_collision) a.out.wall_collision[e] = rwc;
        }
    }
    if (live || filler)
        push_store_obs(a, e, a.out.observation, a.out.achieved_goal, (live && (a.write_goal || need)) ? a.out.desired_goal : nullptr,
                       obs, ag, s.goal);
    if (!live) return;
    push_store(a, e, s);
    a.rng[e] = event;
    a.elapsed[e] = elapsed;
    if (a.autoreset == GPR_AUTORESET_NEXT_STEP) a.needs_reset[e] = (done && !need) ? 1 : 0;
}

template <bool BOX, bool NOISE>
__global__ void __launch_bounds__(kPushCta, GPR_PUSH_MINB) pushing_step_kernel(const __grid_constant__ PushArgs a) {
    __shared__ Tables tb;
    load_tables(tb, a.L);
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // the other buffer belongs to the next step: clear it now
        a.queue_ctl[a.parity ^ 1] = 0ull;
        a.queue_cursor[a.parity ^ 1] = 0u;
    }
    const unsigned lane = threadIdx.x & 31u;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < a.B;  // (no early return: the reset's stage B is warp-collective)
    const uint32_t env_global = a.env_base + (uint32_t)e;
    PushState s;
    memset(&s, 0, sizeof(s));
    uint32_t event = 0;
    int elapsed = 0;
    bool pending = false;
    if (valid) {
        push_load(a, e, s);
        event = a.rng[e];
        elapsed = a.elapsed[e];
        pending = a.autoreset == GPR_AUTORESET_NEXT_STEP && a.needs_reset[e] != 0;
    }
    const bool stepped = valid && !pending;
    double ux = 0.0, uy = 0.0;
    if (stepped) {
        const float2 af = a.action[e];
        ux = fmin(fmax((double)af.x, -a.act_lim), a.act_lim);  // basic:1869-1873
        uy = fmin(fmax((double)af.y, -a.act_lim), a.act_lim);
    }
    int gi = 0, gj = 0;
    guess_cell(a, s.M.x, s.M.y, gi, gj);
    const float cwf = (float)a.c_wall[0][0];
    float travel = 0.f, lim_w = -1.f;
    bool active = stepped, wc = false;
    int parked_at = -1;
    for (int cyc = 0; cyc < a.num_cycles; ++cyc) {  // basic:1879
        if (!__any_sync(FULL, active)) break;
        if (!active) continue;
        if (!gpr_push_is_free(&a.P, &s.M, &s.O)) {  // contact regime from this cycle on: hand the env over
            parked_at = cyc;
            active = false;
            continue;
        }
        const uint32_t s0 = (uint32_t)cyc * 4u;
        float n4[4] = {0.f, 0.f, 0.f, 0.f};
        bool have0 = false;
        double cx, cy, qax, qay;
        push_control<BOX, NOISE>(a, s, ux, uy, env_global, event, s0, n4, have0, cx, cy);
        // mj_step (basic:1882), free regime.  A mover that has never been rotated (yaw exactly 0, yaw rate 0 — every env that
        // has not had a contact since its reset) and an object at exact rest make everything of gpr_push_substep_free except
        // the mover's x / y integration an exact no-op (tau = 0, every object term 0): integrate just that, with the very
        // operations the general function applies (ux + 0 * (1/m) = ux + 0, v += dt a, x += dt v).
        if (s.M.w == 0.0 && s.M.s == 0.0 && s.M.c == 1.0) {
            qax = dadd(cx, 0.0);
            qay = dadd(cy, 0.0);
            s.M.vx = dadd(s.M.vx, dmul(a.P.dt, qax));
            s.M.vy = dadd(s.M.vy, dmul(a.P.dt, qay));
            s.M.x = dadd(s.M.x, dmul(a.P.dt, s.M.vx));
            s.M.y = dadd(s.M.y, dmul(a.P.dt, s.M.vy));
        } else {
            gpr_push_substep_free(&a.P, &s.M, &s.O, cx, cy, &qax, &qay);
        }
        s.acc = make_double2(qax, qay);
        wc = push_wall_cycle<BOX, NOISE>(a, tb, s, env_global, event, s0, n4, have0, cwf, gi, gj, travel, lim_w);
        if (wc) active = false;  // basic:1904
    }
    // ---- parked envs: state as of the top of cycle `parked_at` + a queue entry (slots reserved with one atomic per warp)
    const bool parked = parked_at >= 0;
    const unsigned pm = __ballot_sync(FULL, parked);
    if (pm) {
        unsigned long long t = 0ull;
        if (lane == 0) t = atomicAdd(a.queue_ctl + a.parity, (unsigned long long)__popc(pm));
        const unsigned slot0 = (unsigned)__shfl_sync(FULL, t, 0);
        GPR_CHECK(a, !parked || (slot0 + __popc(pm & ((1u << lane) - 1u)) < (unsigned)a.B && e < a.B), DBG_LIST_SLOT);
        if (parked) {
            push_store(a, e, s);
            a.queue[slot0 + __popc(pm & ((1u << lane) - 1u))] = ((unsigned long long)(unsigned)parked_at << 32) | (unsigned long long)(uint32_t)e;
        }
    }
    push_finish<BOX, NOISE>(a, tb, e, valid && !parked, valid && parked, pending, s, env_global, event, elapsed, wc);
}

#ifndef GPR_PUSH_CONTACT_MINB
#define GPR_PUSH_CONTACT_MINB 6
#endif
constexpr int kPushContactCta = 64;
#ifndef GPR_PUSH_MIN_PULL
#define GPR_PUSH_MIN_PULL 8u
#endif

template <bool BOX, bool NOISE>
__global__ void __launch_bounds__(kPushContactCta, GPR_PUSH_CONTACT_MINB) pushing_contact_kernel(const __grid_constant__ PushArgs a) {
    __shared__ Tables tb;
    load_tables(tb, a.L);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t count = (uint32_t)a.queue_ctl[a.parity];  // final: pushing_step_kernel has completed
    uint32_t* const cursor = a.queue_cursor + a.parity;
    const float cwf = (float)a.c_wall[0][0];
    // Envs per pull.  The kernel is bound by the LATENCY of one warp's serial substep chain, lengthened by every distinct path
    // (0 / 1 / 2 contact points, sliding or resting corners) its lanes take: while the queue is short enough for the grid's
    // warps to take it in one round, a warp pulls fewer envs (down to GPR_PUSH_MIN_PULL) and leaves its other lanes idle.
    const uint32_t warps = gridDim.x * (kPushContactCta / 32);
    uint32_t pull = 32u;
    while (pull > GPR_PUSH_MIN_PULL && (unsigned long long)(pull >> 1) * warps >= count) pull >>= 1;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, pull);
        base = __shfl_sync(FULL, base, 0);
        if (base >= count) break;
        const bool live = lane < pull && base + lane < count;
        int e = 0, cyc0 = a.num_cycles;
        PushState s;
        memset(&s, 0, sizeof(s));
        uint32_t event = 0;
        int elapsed = 0;
        double ux = 0.0, uy = 0.0;
        float warm[GPR_PUSH_WARM];
#pragma unroll
        for (int i = 0; i < GPR_PUSH_WARM; ++i) warm[i] = 0.f;
        if (live) {
            const unsigned long long entry = a.queue[base + lane];
            e = (int)(uint32_t)entry;
            cyc0 = (int)(entry >> 32);
            GPR_CHECK(a, (uint32_t)entry < (uint32_t)a.B && cyc0 >= 0 && cyc0 < a.num_cycles, DBG_LIST_ENTRY);
            GPR_CHECK(a, base + lane < (uint32_t)a.B, DBG_LIST_SLOT);
            push_load(a, e, s);
#pragma unroll
            for (int i = 0; i < GPR_PUSH_WARM; ++i) warm[i] = a.warm[(size_t)e * GPR_PUSH_WARM + i];
            event = a.rng[e];
            elapsed = a.elapsed[e];
            const float2 af = a.action[e];
            ux = fmin(fmax((double)af.x, -a.act_lim), a.act_lim);  // basic:1869-1873
            uy = fmin(fmax((double)af.y, -a.act_lim), a.act_lim);
        }
        const uint32_t env_global = a.env_base + (uint32_t)e;
        int gi = 0, gj = 0;
        guess_cell(a, s.M.x, s.M.y, gi, gj);
        float travel = 0.f, lim_w = -1.f;
        bool wc = false;
        for (int cyc = cyc0; cyc < a.num_cycles; ++cyc) {
            const uint32_t s0 = (uint32_t)cyc * 4u;
            float n4[4] = {0.f, 0.f, 0.f, 0.f};
            bool have0 = false;
            double cx, cy, qax, qay;
            push_control<BOX, NOISE>(a, s, ux, uy, env_global, event, s0, n4, have0, cx, cy);
            gpr_push_substep(&a.P, &s.M, &s.O, cx, cy, &qax, &qay, warm);  // mj_step (basic:1882), general
            s.acc = make_double2(qax, qay);
            wc = push_wall_cycle<BOX, NOISE>(a, tb, s, env_global, event, s0, n4, have0, cwf, gi, gj, travel, lim_w);
            if (wc) break;  // basic:1904
        }
        __syncwarp();
        if (live) {
            // An env that is FREE again (object at exact rest, out of reach) will run its next cycles in
            // pushing_step_kernel, which never touches this state: leave it as the free substep would — all zero.
            const bool free_now = gpr_push_is_free(&a.P, &s.M, &s.O);
#pragma unroll
            for (int i = 0; i < GPR_PUSH_WARM; ++i) a.warm[(size_t)e * GPR_PUSH_WARM + i] = free_now ? 0.f : warm[i];
        }
        push_finish<BOX, NOISE>(a, tb, e, live, false, false, s, env_global, event, elapsed, wc);  // (a reset zeroes it again)
    }
}

template <bool BOX, bool NOISE>
__global__ void __launch_bounds__(128) pushing_reset_kernel(const __grid_constant__ PushArgs a) {
    __shared__ Tables tb;
    load_tables(tb, a.L);
    __syncthreads();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool need = e < a.B && !(a.reset_mask && !a.reset_mask[e]);
    const uint32_t env_global = a.env_base + (uint32_t)e;
    PushState s;
    memset(&s, 0, sizeof(s));
    uint32_t event = 0;
    bool settled = true;
    if (need) {
        event = a.rng[e];
        settled = push_reset_a(a, s, env_global, event, a.inject_start, a.inject_object, e);
    }
    const bool failed = push_reset_b(a, s, env_global, event, need && !settled);
    if (!need) return;
    bool wc;
    push_reset_c<BOX, NOISE>(a, tb, s, env_global, event, a.inject_goal, e, wc);
    double obs[6];
    double2 ag;
    bool reached;
    push_observe<NOISE>(a, s, env_global, event, obs, ag, reached);
    push_store_obs(a, e, a.out.observation, a.out.achieved_goal, a.out.desired_goal, obs, ag, s.goal);
    push_store(a, e, s);
    float r;
    bool t, succ;
    push_reward(reached, wc, r, t, succ);
    if (a.out.is_success) a.out.is_success[e] = succ;
    if (a.out.mover_collision) a.out.mover_collision[e] = 0;
    if (a.out.wall_collision) a.out.wall_collision[e] = wc;
    a.rng[e] = event + 1u;
    a.elapsed[e] = 0;
    a.ep_return[e] = 0.f;
    if (a.needs_reset) a.needs_reset[e] = 0;
    if (failed) atomicAdd(a.fail_count, 1u);
}

}  // namespace gpr
