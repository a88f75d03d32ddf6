// gpr_planning.cuh — fused kernels of BenchmarkPlanningEnv's step path.
//
//   planning_step_kernel     : basic:1835-1950 in one launch — action clip, num_cycles x { plan:420-450 limit control,
//                              mj_step-equivalent integration, basic:459-788 wall check (+ static obstacles, the typed
//                              form of the basic:1976-1986 hook), basic:355-424 mover check, break on collision },
//                              plan:536-573 observation, plan:575-602 info, plan:502-534 reward, plan:459-479 terminated,
//                              TimeLimit truncation, episode statistics; finished envs are PUBLISHED on a work list.
//   planning_autoreset_kernel: plan:355-418 + basic:1797-1805 for the envs on that list, launched as the step kernel's
//                              programmatic dependent: it consumes entries while the step kernel's last wave still runs
//                              (rejection sampling of starts and goals with the counter-based RNG of include/gpr_rng.h,
//                              the whole warp per env; checks / observation / stores per batch of 32/G envs).
//   planning_reset_kernel    : basic:1770-1833 for a masked subset, optionally with injected starts / goals.
//
// Exact optimisations that shape the code (DESIGN.md "Kernels"); every one leaves the oracle's results bit for bit:
//   * LAZY NOISE.  The sensor noise of the reference (sigma = 1e-5 by default) is drawn three times per mover and cycle
//     (plan:430, basic:1888-1901) but only ever feeds comparisons.  The portable normal generator is bounded
//     (|n| <= GPR_NORMAL_ABS_MAX), the RNG is counter-based (skipping a draw changes no other draw), so a comparison whose
//     noise-free margin exceeds the noise bound has the same outcome with and without the noise.  The kernels evaluate
//     each check noise-free first and generate the noise only inside the band where it can matter.
//   * FLOAT32 SCREENS WITH EXACT FALLBACK.  wall_core / pair_circle_fast / pair_box_screen decide a comparison in float32
//     when its margin exceeds noise bound + rounding slack; only the thin uncertain band re-evaluates in float64.
//   * TEMPORAL COHERENCE.  A check certifies a clearance; while a mover's accumulated travel stays below it the
//     reference's check is known to be negative and is skipped (planning_step_kernel).
//   * WARP-UNIFORM FREE-RUN PATH.  Control + integration without any clip or noise when no lane of the warp needs them.
//   * REJECTION SAMPLING IN BULK.  Acceptance of "all movers at once" sampling is ~1% for 4 movers on 3x3 tiles.  For
//     G <= 4 every lane screens two complete attempts (64 per warp iteration, sample_env_lanes); for wider groups all
//     32/G lane groups test different attempts (sample_env_groups).  The lowest accepted attempt index wins, which is
//     exactly the sequential loop's result.
#pragma once

#include "gpr_device.cuh"

namespace gpr {

struct PlanArgs {
    int B, N;
    int learn_jerk, num_cycles, max_episode_steps, autoreset, max_reset_attempts;
    int uniform_pairs;  // circle: one threshold for every pair (equal radii, or the basic:409 broadcast quirk)
    uint32_t env_base;
    uint64_t seed;
    double dt, inv_dt, v_max, a_max, j_max, act_lim;  // inv_dt = RN(1 / dt), see ddiv_rcp
    double v_max2_lo, a_max2_lo;  // max^2 * (1 - 1e-14), see ensure_max
    double v_lazy2;               // (v_max - vel-noise bound)^2: below it the velocity clip cannot trigger (or -1)
    double threshold, min_goal_dist;
    double min_xy[2], span_xy[2];
    double sigma_p, sigma_v;
    double pair_margin;     // bound on |noisy distance - distance| (0 without noise)
    double pair_t[2];       // [safety] common pair threshold when uniform_pairs
    double band_lo2[2][2];  // [safety][noisy] d^2 below  -> certain hit   (uniform_pairs)
    double band_hi2[2][2];  // [safety][noisy] d^2 above  -> certain miss
    float wall_delta;       // bound on the position noise per coordinate (+ float slack)
    float wxf, wyf;         // tile widths as float (prefilter only)
    // float32 prefilters: decide a comparison in float32 when its margin exceeds noise bound + float rounding slack,
    // fall back to the exact float64 evaluation otherwise
    float pair_lo2f[2][2], pair_hi2f[2][2];  // [safety][noisy] squared-distance bands (uniform_pairs)
    float pair_mgf[2];                       // [noisy] distance margin for per-mover radii
    float pair_thif[2][2];                   // [safety][noisy] upper edge of the band as a distance (uniform_pairs)
    float rot_extf;                          // box shape: bound on |sin| of the noise-induced rotation (0 without noise)
    float inv_dtf;                           // 1 / cycle_time, rounded down (travel budget in velocity units)
    float inv_wxf, inv_wyf;                  // 1 / tile width (cell guess of the float32 screens)
    float goal_lo2f, goal_hi2f;              // min_goal_dist bands
    float minxf, minyf, spanxf, spanyf;      // spawn box as float
    LayoutArgs L;
    // static obstacles (gpr_config.num_obstacles): x, y, size0, size1 — circle radius / box half sizes
    int n_obst;
    float obst_delta;  // float32 screen slack: position-noise bound + float rounding
    float obst_vmaxf;  // largest prescribed speed of an extra body (gpr_config.obstacle_vel), rounded up; 0 = all static
    double obst[GPR_MAX_OBSTACLES][4];
    double obst_vel[GPR_MAX_OBSTACLES][2];
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
    double* stats;         // 6 accumulators, see gpr_episode_stats
    uint32_t* fail_count;  // number of resets whose rejection loop hit max_reset_attempts
    uint32_t* debug_errors;  // [DBG_NUM_SLOTS] GPR_DEBUG_BOUNDS builds only (see gpr_device.cuh)
    // auto-reset work list: the step kernel appends finished envs, the auto-reset kernel consumes them — WHILE the step
    // kernel is still running when the two are launched as a programmatic dependent pair (see planning_autoreset_kernel).
    unsigned long long* reset_list;  // [B]  (RNG event << 32 | env); all ones = slot not published (consumers restore it)
    unsigned long long* reset_ctl;   // [2]  (warps of the step kernel that have reported << 32 | slots reserved),
                                     //      double-buffered by step parity
    uint32_t* reset_cursor;          // [2]  slots claimed by consumers
    unsigned step_ctas;      // grid size of the step kernel (set by the launcher)
    int overlap;             // host side: launch the auto-reset kernel as the step kernel's programmatic dependent
    int parity;
    int write_goal;  // 0: the step kernel leaves desired_goal rows of envs that were not reset alone (GPR_OUT_GOAL_ON_CHANGE)
    int out_f64;     // GPR_OUT_FLOAT64: observation / goal outputs are double arrays
    // COMPACT TRANSPORT of the sparse results (gpr_step_host, copy-engine route; SAME_STEP auto-reset): the terminal
    // observation of a finished env and the goal of its new episode go to row `slot` of these arrays — slot = the env's
    // position on the auto-reset work list — instead of row `env` of final_* / desired_goal, with compact_index[slot] = env;
    // the host copies the first `count` rows (count = entries published) and scatters them into the caller's arrays.
    // All NULL = off (the kernels then write final_* / desired_goal rows in place).
    float* compact_final_obs;
    float* compact_final_ag;
    float* compact_final_dg;
    float* compact_goal;
    int32_t* compact_index;
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
    GPR_CHECK(a, !ln.active || (ln.env >= 0 && ln.idx < (size_t)a.B * (size_t)a.N), DBG_LANE_INDEX);
    return ln;
}

__device__ __forceinline__ double noisy(double x, float n, double sigma) { return dadd(x, dmul((double)n, sigma)); }

// gpr_normal4 for the lazily generated noise: those sites are cold (a check only needs its noise inside a ~1e-4 m band),
// so one out-of-line copy keeps the hot loops small (instruction cache) instead of ~150 inlined instructions per site
GPR_COLD(GPR_INL_NORMAL) void normal4_cold(uint64_t seed, uint32_t env_global, uint32_t event, uint32_t stream,
                                                 uint32_t lane, float (&out)[4]) {
    gpr_normal4(seed, env_global, event, stream, lane, out);
}

// ---- circle pair check with the lazy-noise band (basic:392-409), warp-collective --------------------------------------
//   NOISY: the positions carry the mover-check noise of (event, stream) — generated only inside the uncertainty band.
template <int G, bool NOISY>
GPR_COLD(GPR_INL_PAIR) bool pair_circle(const PlanArgs& a, unsigned lane, int m, bool part, double x, double y,
                                            double r, int safety, uint32_t env_global, uint32_t event, uint32_t stream) {
    bool hit = false;
    if (G == 1) return false;
    const unsigned base = lane & ~(unsigned)(G - 1);
#pragma unroll
    for (int k = 1; k <= G / 2; ++k) {
        const int pm = (m + k) & (G - 1);
        const int src = (int)(base | (unsigned)pm);
        const double ox = __shfl_sync(FULL, x, src);
        const double oy = __shfl_sync(FULL, y, src);
        const bool opart = __shfl_sync(FULL, (int)part, src) != 0;
        const bool mine = part && opart && !(k == G / 2 && m >= G / 2);
        const double dx = dsub(x, ox), dy = dsub(y, oy);
        const double d2 = dadd(dmul(dx, dx), dmul(dy, dy));
        double t, lo2, hi2;
        if (a.uniform_pairs) {
            t = a.pair_t[safety];
            lo2 = a.band_lo2[safety][NOISY ? 1 : 0];
            hi2 = a.band_hi2[safety][NOISY ? 1 : 0];
        } else {
            const double orr = __shfl_sync(FULL, r, src);
            t = dadd(r, orr);  // basic:409
            const double mg = NOISY ? a.pair_margin : 0.0;
            const double tl = t - mg, th = t + mg;
            lo2 = tl > 0.0 ? tl * tl * (1.0 - 1e-14) : -1.0;
            hi2 = th * th * (1.0 + 1e-14);
        }
        if (mine) {
            if (d2 <= lo2) {
                hit = true;
            } else if (!(d2 >= hi2)) {  // uncertain band (or NaN): decide exactly like the reference
                double xi = x, yi = y, xj = ox, yj = oy;
                if (NOISY) {
                    float k4[4];
                    normal4_cold(a.seed, env_global, event, stream, (uint32_t)m, k4);
                    xi = noisy(x, k4[0], a.sigma_p);
                    yi = noisy(y, k4[1], a.sigma_p);
                    normal4_cold(a.seed, env_global, event, stream, (uint32_t)pm, k4);
                    xj = noisy(ox, k4[0], a.sigma_p);
                    yj = noisy(oy, k4[1], a.sigma_p);
                }
                const double ex = dsub(xi, xj), ey = dsub(yi, yj);
                if (dsqrt(dadd(dmul(ex, ex), dmul(ey, ey))) <= t) hit = true;
            }
        }
    }
    return hit;
}

// Wall check in float32 for the tracked cell (gi, gj) of an axis-aligned rectangle with half sizes (cx, cy) — the circle's
// bounding square of basic:545-558, or a bounding rectangle of a box-shaped mover.  Returns 1 valid / 0 invalid when every
// comparison of basic:507-558 has a margin larger than `wall_delta` (noise bound + float rounding slack) — the float32
// signs are then certain and equal the reference's float64 ones — and 2 when some comparison is too close to call (or the
// mover left the tracked cell): the caller then runs the exact float64 check (with the noise, if any).
//
// `clear` (>= 0): a certified L-infinity distance the centre may still travel without the check turning invalid.  The
// check is "the rectangle [x-cx, x+cx] x [y-cy, y+cy] overlaps no missing cell" (sides_ok: an unsafe side needs that
// neighbour, an unsafe corner also the diagonal one), so the clearance is the smallest L-infinity gap between the
// rectangle and a missing cell of the 3x3 neighbourhood, capped by the gap to the ring beyond it (>= tile width - c),
// minus wall_delta.  It does not depend on how close the verdict of THIS cycle was.
template <class Args>  // PlanArgs or PushArgs: needs wxf, wyf, wall_delta, L
__device__ __forceinline__ int wall_core(const Args& a, const Tables& tb, float fx, float fy, float cx, float cy, int gi,
                                         int gj, float& clear) {
    const float wc = a.wxf - cx, hc = a.wyf - cy;
    const float gw = fx - cx, ge = wc - fx, gs = fy - cy, gn = hc - fy;  // gaps to the W / E columns and S / N rows
    const bool inside = fx > 0.f && fx < a.wxf && fy > 0.f && fy < a.wyf;
    const uint32_t code = tb.cell[gi * a.L.ny + gj];
    float c = fminf(wc, hc);
    c = (code & CELL_W) ? c : fminf(c, gw);
    c = (code & CELL_E) ? c : fminf(c, ge);
    c = (code & CELL_S) ? c : fminf(c, gs);
    c = (code & CELL_N) ? c : fminf(c, gn);
    c = (code & CELL_SW) ? c : fminf(c, fmaxf(gw, gs));
    c = (code & CELL_NW) ? c : fminf(c, fmaxf(gw, gn));
    c = (code & CELL_SE) ? c : fminf(c, fmaxf(ge, gs));
    c = (code & CELL_NE) ? c : fminf(c, fmaxf(ge, gn));
    clear = (inside && (code & CELL_T)) ? fmaxf(c - a.wall_delta, 0.f) : 0.f;
    // certainty of this cycle's verdict: distance to the nearest comparison threshold of the tracked cell
    const float mx = fminf(fminf(fabsf(fx), fabsf(gw)), fminf(fabsf(ge), fabsf(fx - a.wxf)));
    const float my = fminf(fminf(fabsf(fy), fabsf(gs)), fminf(fabsf(gn), fabsf(fy - a.wyf)));
    if (!(fminf(mx, my) >= a.wall_delta) || !inside) return 2;
    const uint32_t u = (gw < 0.f ? 1u : 0u) | (ge < 0.f ? 2u : 0u) | (gs < 0.f ? 4u : 0u) | (gn < 0.f ? 8u : 0u);
    return ((code & CELL_3X3) || sides_ok(u, code)) ? 1 : 0;
}
template <class Args>
__device__ __forceinline__ int wall_fast(const Args& a, const Tables& tb, double x, double y, float cx, float cy, int gi,
                                         int gj, float& clear) {
    return wall_core(a, tb, (float)dsub(x, tb.xlo[gi]), (float)dsub(y, tb.ylo[gj]), cx, cy, gi, gj, clear);
}
// the same screen for a float32 position (rejection sampling): the cell is guessed from the float coordinates; a
// coordinate within wall_delta of a cell border comes back "too close to call" like any other near-threshold case
__device__ __forceinline__ int wall_screen_f(const PlanArgs& a, const Tables& tb, float xf, float yf, float cx, float cy) {
    const int gi = min(max(__float2int_rd(xf * a.inv_wxf), 0), a.L.nx - 1);
    const int gj = min(max(__float2int_rd(yf * a.inv_wyf), 0), a.L.ny - 1);
    float clear;
    return wall_core(a, tb, xf - tb.xlof[gi], yf - tb.ylof[gj], cx, cy, gi, gj, clear);
}

template <class Args>
__device__ __forceinline__ void guess_cell(const Args& a, double x, double y, int& gi, int& gj) {
    gi = min(max(__double2int_rd(dmul(x, a.L.inv_wx)), 0), a.L.nx - 1);
    gj = min(max(__double2int_rd(dmul(y, a.L.inv_wy)), 0), a.L.ny - 1);
}

// basic:1888-1894 for one mover (circle): float32 prefilter, exact float64 fallback with lazily generated noise.
//   n4/have : this cycle's noise block 0 (words 2,3 = wall noise), generated on demand
template <bool NOISE>
__device__ __forceinline__ bool wall_bad_circle(const PlanArgs& a, const Tables& tb, bool part, double x, double y,
                                                double c, float cf, int& gi, int& gj, uint32_t env_global, uint32_t event,
                                                uint32_t stream, int m, int w0, float (&n4)[4], bool& have,
                                                float& margin) {
    margin = 3.0e38f;
    if (!part) return false;
    const int f = wall_fast(a, tb, x, y, cf, cf, gi, gj, margin);
    if (f != 2) return f == 0;
    // (the clearance computed for the tracked cell stays valid: it is 0 whenever the mover is not strictly inside it)
    double wx = x, wy = y;
    if (NOISE) {
        if (!have) {
            normal4_cold(a.seed, env_global, event, stream, (uint32_t)m, n4);
            have = true;
        }
        wx = noisy(x, n4[w0], a.sigma_p);
        wy = noisy(y, n4[w0 + 1], a.sigma_p);
    }
    guess_cell(a, x, y, gi, gj);  // refresh the tracked cell
    Rect dummy;
    return !wall_valid<false>(tb, a.L, wx, wy, c, dummy);
}

// Circle pair check with a float32 prefilter (warp-collective).  xf/yf are this lane's float coordinates (1e30f when the
// lane does not take part); a pair is decided in float32 when its squared distance is outside the band, and only if some
// lane of the warp meets an undecided pair does the whole warp run the exact float64 check.
// `margin`: min over ALL pairs this mover is part of of (distance - upper band edge), 0 when a pair was hit or undecided.
// A pair cannot start to collide while its two movers have together travelled less than the pair's margin; each mover may
// therefore use up HALF of its own `margin` (which is <= the margin of each of its pairs) — a mover far from everything
// keeps a large budget even when two other movers of its env are about to touch.
template <int G, bool NOISY>
__device__ __forceinline__ bool pair_circle_fast(const PlanArgs& a, unsigned lane, int m, bool part, double x, double y,
                                                 double r, int safety, uint32_t env_global, uint32_t event,
                                                 uint32_t stream, float& margin) {
    margin = 3.0e38f;
    if (G == 1) return false;
    const float xf = part ? (float)x : 1e30f, yf = part ? (float)y : 1e30f;
    const float rf = (float)r;
    const unsigned base = lane & ~(unsigned)(G - 1);
    bool hit = false, unc = false;
#pragma unroll
    for (int k = 1; k <= G / 2; ++k) {
        const int src = (int)(base | (unsigned)((m + k) & (G - 1)));
        const float oxf = __shfl_sync(FULL, xf, src), oyf = __shfl_sync(FULL, yf, src);
        float lo2, hi2, thi;
        if (a.uniform_pairs) {
            lo2 = a.pair_lo2f[safety][NOISY ? 1 : 0];
            hi2 = a.pair_hi2f[safety][NOISY ? 1 : 0];
            thi = a.pair_thif[safety][NOISY ? 1 : 0];
        } else {
            const float t = rf + __shfl_sync(FULL, rf, src), mg = a.pair_mgf[NOISY ? 1 : 0];
            lo2 = t > mg ? (t - mg) * (t - mg) : -1.f;
            hi2 = (t + mg) * (t + mg);
            thi = (t + mg) * 1.000001f;
        }
        const bool both = part && oxf < 1e29f;
        const bool mine = both && !(k == G / 2 && m >= G / 2);
        const float dx = xf - oxf, dy = yf - oyf;
        const float d2 = dx * dx + dy * dy;
        if (mine) {
            if (d2 < lo2) hit = true;
            else if (!(d2 > hi2)) unc = true;
        }
        // this pair's margin, and — for the mover on the other side of the lane group — the margin of the pair (m-k, m) that
        // lane m-k has just computed: `margin` ends up as the minimum over ALL pairs this mover is part of
        const float mk = both ? fmaxf(sqrtf(d2) * 0.999999f - thi, 0.f) : 3.0e38f;
        margin = fminf(margin, mk);
        if (k < G / 2) margin = fminf(margin, __shfl_sync(FULL, mk, (int)(base | (unsigned)((m - k) & (G - 1)))));
    }
    if (__any_sync(FULL, unc)) hit = pair_circle<G, NOISY>(a, lane, m, part, x, y, r, safety, env_global, event, stream);
    return hit;
}

// Box shape: float32 screen of the mover-mover check (warp-collective).  Planning movers never rotate (the only rotation
// is the sensor noise on the quaternion), so each rectangle lies inside the axis-aligned box [x +- ex] x [y +- ey] with
// (ex, ey) = half sizes widened by the noise-rotation bound.  Two such boxes that are separated along x OR along y by more
// than the position-noise bound + float slack `mg` cannot have intersecting edges (geom:107-138): a certain miss, and
// `margin` is the largest of the two axis gaps — the L-infinity distance the pair may still close before that changes.
// `near` marks lanes with a pair the screen cannot clear (the caller then runs the exact rectangle test for the warp).
// `sure` (only with hx > 0: all boxes have ONE size (hx, hy), the true half sizes): the two boxes overlap along BOTH axes by
// more than the noise bound + slack.  Two equal rectangles cannot contain one another, so their edges cross: a certain hit
// (geom:107-138 reports it) — the exact 16-edge-pair test is only needed for pairs in the thin band around first touch.
template <int G>
__device__ __forceinline__ void pair_box_screen(unsigned lane, int m, bool part, double x, double y, float exf, float eyf,
                                                float mg, bool& near, float& margin, unsigned& kmask, bool& sure, float hx = 0.f,
                                                float hy = 0.f, float shrink = 0.f) {
    near = false;
    sure = false;
    margin = 3.0e38f;
    kmask = 0u;  // bit k: the pair (m, m+k) could not be cleared
    if (G == 1) return;
    const float xf = part ? (float)x : 1e30f, yf = part ? (float)y : 1e30f;
    const unsigned base = lane & ~(unsigned)(G - 1);
#pragma unroll
    for (int k = 1; k <= G / 2; ++k) {
        const int src = (int)(base | (unsigned)((m + k) & (G - 1)));
        const float oxf = __shfl_sync(FULL, xf, src), oyf = __shfl_sync(FULL, yf, src);
        const float tx = (exf + __shfl_sync(FULL, exf, src) + mg) * 1.000001f;
        const float ty = (eyf + __shfl_sync(FULL, eyf, src) + mg) * 1.000001f;
        const bool both = part && oxf < 1e29f;
        const float adx = fabsf(xf - oxf), ady = fabsf(yf - oyf);
        const float mgn = fmaxf(adx * 0.999999f - tx, ady * 0.999999f - ty);
        if (both) {
            // certain hit: overlap along both axes deeper than (noise bound + rotation of the noisy quaternion + slack)
            const bool hitc = hx > 0.f && adx * 1.000001f < 2.f * hx - shrink && ady * 1.000001f < 2.f * hy - shrink;
            sure = sure || hitc;
            near = near || (!(mgn > 0.f) && !hitc);
            kmask |= (!(mgn > 0.f) && !hitc) ? (1u << k) : 0u;
        }
        // (own pairs on both sides of the lane group, see pair_circle_fast)
        const float mk = both ? fmaxf(mgn, 0.f) : 3.0e38f;
        margin = fminf(margin, mk);
        if (k < G / 2) margin = fminf(margin, __shfl_sync(FULL, mk, (int)(base | (unsigned)((m - k) & (G - 1)))));
    }
}

// ---- static obstacles (the typed form of basic:1976-1986, see gpr_config.num_obstacles) -------------------------------
// Exact test of ONE mover against every obstacle, on the position (x, y) the caller has made noisy or not; (s0, s1) are the
// mover's collision sizes, rm its rectangle (box shape).  Mirrors gpro_check_obstacle_collision of the oracle.
template <bool BOX>
// t: time the extra bodies have been moving (gpr_config.obstacle_vel); 0 at reset() and for start / goal sampling.
static __device__ __noinline__ bool obstacle_hit_exact(const PlanArgs& a, double x, double y, double s0, const Rect& rm, double t = 0.0) {
#pragma unroll 1
    for (int k = 0; k < a.n_obst; ++k) {
        const double ox = dadd(a.obst[k][0], dmul(a.obst_vel[k][0], t)), oy = dadd(a.obst[k][1], dmul(a.obst_vel[k][1], t));
        if (!BOX) {
            const double dx = dsub(x, ox), dy = dsub(y, oy);
            if (dsqrt(dadd(dmul(dx, dx), dmul(dy, dy))) <= dadd(s0, a.obst[k][2])) return true;
        } else {
            const double h0 = a.obst[k][2], h1 = a.obst[k][3];
            if (fabs(dsub(x, ox)) <= h0 && fabs(dsub(y, oy)) <= h1) return true;  // centre inside: the edge test is blind to it
            Rect ro;
            rect_vertices_axis(ox, oy, h0, h1, ro);
            if (rects_intersect(rm, ro)) return true;
        }
    }
    return false;
}
// float32 screen: 0 = certain miss (clear = distance the mover may still travel before that can change), 1 = certain hit
// (circle only), 2 = too close to call.  (e0, e1): circle radius / bounding half extents of the (noise-rotated) box.
template <bool BOX>
__device__ __forceinline__ int obstacle_screen(const PlanArgs& a, double x, double y, float e0, float e1, float& clear, double t = 0.0) {
    clear = 3.0e38f;
    int verdict = 0;
#pragma unroll 1
    for (int k = 0; k < a.n_obst; ++k) {
        const double ox = dadd(a.obst[k][0], dmul(a.obst_vel[k][0], t)), oy = dadd(a.obst[k][1], dmul(a.obst_vel[k][1], t));
        const float dx = fabsf((float)dsub(x, ox)), dy = fabsf((float)dsub(y, oy));
        float gap;
        if (!BOX) {
            const float t = e0 + (float)a.obst[k][2];
            const float d = sqrtf(dx * dx + dy * dy);
            gap = d * 0.999999f - t * 1.000001f - a.obst_delta;
            if (d * 1.000001f + a.obst_delta < t * 0.999999f) verdict = max(verdict, 1);
            else if (!(gap > 0.f)) verdict = 2;
        } else {
            const float gx = dx * 0.999999f - (e0 + (float)a.obst[k][2]) * 1.000001f - a.obst_delta;
            const float gy = dy * 0.999999f - (e1 + (float)a.obst[k][3]) * 1.000001f - a.obst_delta;
            gap = fmaxf(gx, gy);
            if (!(gap > 0.f)) verdict = 2;
        }
        clear = fminf(clear, fmaxf(gap, 0.f));
    }
    if (verdict != 0) clear = 0.f;
    return verdict;
}

// One observation row (plan:536-573) + the per-env reductions the reward needs.
template <int G, bool NOISE>
__device__ __forceinline__ void observe(const PlanArgs& a, const Lane<G>& ln, uint32_t event, double2 p, double2 v,
                                        double2 goal, double2& ag, double2& ov, int& reached_cnt) {
    ag = p;
    ov = v;
    if (NOISE) {
        float n4[4];
        gpr_normal4(a.seed, ln.env_global, event, GPR_RNG_OBS, (uint32_t)ln.m, n4);
        ag.x = noisy(p.x, n4[0], a.sigma_p);
        ag.y = noisy(p.y, n4[1], a.sigma_p);
        ov.x = noisy(v.x, n4[2], a.sigma_v);
        ov.y = noisy(v.y, n4[3], a.sigma_v);
    }
    const double dx = dsub(ag.x, goal.x), dy = dsub(ag.y, goal.y);
    const bool reached = ln.active && sqrt_le(dadd(dmul(dx, dx), dmul(dy, dy)), a.threshold);  // plan:521
    reached_cnt = __popc(__ballot_sync(FULL, reached) & ln.gmask);
}

// one (x, y) pair of an output row: element index `pair` of a float2 (or, with GPR_OUT_FLOAT64, double2) array
__device__ __forceinline__ void store_pair(bool f64, float* base, size_t pair, double x, double y) {
    if (f64) reinterpret_cast<double2*>(base)[pair] = make_double2(x, y);
    else reinterpret_cast<float2*>(base)[pair] = make_float2((float)x, (float)y);
}

// Rows of an arbitrary row index (compact transport: row = work-list slot): the general, colder form of store_obs.  Out of
// line on purpose: inlined at its five call sites it cost the step kernel 3 % (registers and code around the hot loop).
template <int G>
static __device__ __noinline__ void store_obs_row(const PlanArgs& a, const Lane<G>& ln, size_t row, float* O, float* AG, float* DG,
                                              double2 ov, double2 acc, double2 ag, double2 goal) {
    if (!ln.active) return;
    const int N = a.N;
    const bool f64 = a.out_f64 != 0;
    if (O) {
        const size_t rowp = row * (size_t)(N * (1 + a.learn_jerk));
        store_pair(f64, O, rowp + ln.m, ov.x, ov.y);
        if (a.learn_jerk) store_pair(f64, O, rowp + N + ln.m, acc.x, acc.y);
    }
    if (AG) store_pair(f64, AG, row * (size_t)N + ln.m, ag.x, ag.y);
    if (DG) store_pair(f64, DG, row * (size_t)N + ln.m, goal.x, goal.y);
}

// EXTRA = false: the kernel instantiation for plain float32 outputs written in place (the vector envs' default and the
// benchmark configuration) carries none of the float64-output / compact-transport code: measured 1.5 % of the step each.
template <int G, bool EXTRA = true>
__device__ __forceinline__ void store_obs(const PlanArgs& a, const Lane<G>& ln, float* O, float* AG, float* DG, double2 ov,
                                          double2 acc, double2 ag, double2 goal) {
    if (!ln.active) return;
    const int N = a.N;
    if (EXTRA && a.out_f64) {  // (cold: the vector envs use float32 outputs)
        store_obs_row<G>(a, ln, (size_t)ln.env, O, AG, DG, ov, acc, ag, goal);
        return;
    }
    if (O) {
        const size_t row = (size_t)ln.env * (size_t)(2 * N * (1 + a.learn_jerk));
        reinterpret_cast<float2*>(O + row)[ln.m] = make_float2((float)ov.x, (float)ov.y);
        if (a.learn_jerk) reinterpret_cast<float2*>(O + row + 2 * N)[ln.m] = make_float2((float)acc.x, (float)acc.y);
    }
    if (AG) reinterpret_cast<float2*>(AG)[ln.idx] = make_float2((float)ag.x, (float)ag.y);
    if (DG) reinterpret_cast<float2*>(DG)[ln.idx] = make_float2((float)goal.x, (float)goal.y);
}

// Instruction-cache footprint of the auto-reset kernel.  With starts and goals each running their own inlined copy of the
// sampling loop, the kernel's hot code is ~23 KB spread over 64 KB — more than the 32 KB L1.5 instruction cache once the
// step kernel's warps share the SM (the two kernels overlap, see planning_autoreset_kernel) — and "no instruction" was its
// top stall reason (profiles/r1b_full_planning4.txt).  GPR_SAMPLER_ONE_COPY: for the circle shape the kind of sample is a
// run-time argument and both kinds go through ONE copy of the loop.  Measured on B200 (planning4, 65,536 / 1,048,576 envs):
// two copies 276 M / 449 M env-steps/s; one copy 318 M / 462 M; one copy with the Philox block out of line
// (GPR_SAMPLER_PHILOX_CALL=1) 303 M / 437 M; out-of-line Philox with two copies 283 M / 441 M.
#ifndef GPR_SAMPLER_PHILOX_CALL
#define GPR_SAMPLER_PHILOX_CALL 0
#endif
#ifndef GPR_SAMPLER_ONE_COPY
#define GPR_SAMPLER_ONE_COPY 1
#endif
#ifndef GPR_SAMPLER_ROLL_H
#define GPR_SAMPLER_ROLL_H 0
#endif
#if GPR_SAMPLER_PHILOX_CALL
static __device__ __noinline__ gpr_u32x4 sampler_block(uint64_t seed, uint32_t eg, uint32_t ev, uint32_t stream, uint32_t lane) {
    return gpr_rng_block_sampling(seed, eg, ev, stream, lane);
}
#else
__device__ __forceinline__ gpr_u32x4 sampler_block(uint64_t seed, uint32_t eg, uint32_t ev, uint32_t stream, uint32_t lane) {
    return gpr_rng_block_sampling(seed, eg, ev, stream, lane);
}
#endif

// ---- warp-cooperative rejection sampling (plan:369-385 starts / plan:395-413 goals) ----------------------------------
// For every lane group with `need`, find the FIRST attempt t (t = 0, 1, 2, ...) whose N positions pass the test, exactly as
// the reference's sequential while-loop would, but with all 32/G groups of the warp testing different attempts of the
// same environment concurrently.  KIND 0: wall check with safety offset + mover collision with safety offset;
// KIND 1: wall check with safety offset + pairwise distance >= min_goal_dist.
// One environment (global index eg, RNG event ev), whole warp: on return EVERY lane holds, for mover m = lane % G, the
// position of the first accepted attempt (or of the last attempt, with failed = true, if the cap was hit).
// Phase 2 of sample_env (exact float64 confirmation of the attempts that survived the float32 screen; ~1% of them):
// out of line, so the hot screening loop stays small.  Warp-collective.  Returns true (in every lane) when an attempt of
// this Philox block round was accepted; `out` then holds, for mover m = lane % G, the position of the first accepted one.
template <int G, bool BOX, int KIND>
GPR_COLD(GPR_INL_CONFIRM) bool sample_confirm(const PlanArgs& a, const Tables& tb, unsigned lane, unsigned gmask,
                                                   int cap, uint32_t blk, gpr_u32x4 r, unsigned cand_bits,
                                                   unsigned needx_bits, uint32_t eg, uint32_t ev, double2& out) {
    const int m = (int)(lane % G);
    const bool has_mover = m < a.N;
    const int mm = has_mover ? m : 0;
    const double cw0 = a.c_wall[(GPR_MAX_MOVERS + mm) * 2 + 0], cw1 = a.c_wall[(GPR_MAX_MOVERS + mm) * 2 + 1];
    const double cs0 = a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 0], cs1 = a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 1];
    const unsigned base = lane & ~(unsigned)(G - 1);
    double xs[2], ys[2];
    unsigned okmask[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = 2 * (int)blk + h;
        const bool cand = (cand_bits >> h) & 1u, needx = (needx_bits >> h) & 1u;
        const double x = dadd(a.min_xy[0], dmul(a.span_xy[0], gpr_uniform32(r.v[2 * h])));  // plan:377/405
        const double y = dadd(a.min_xy[1], dmul(a.span_xy[1], gpr_uniform32(r.v[2 * h + 1])));
        xs[h] = x;
        ys[h] = y;
        const bool part = has_mover && t < cap && cand;
        bool hit = false;
        if (G > 1 && __any_sync(FULL, needx)) {
            const bool px = part && needx;
            if (KIND == 0) {
                if (BOX) {
                    Rect rm;
                    rect_vertices_axis(x, y, cs0, cs1, rm);
                    hit = pair_check<G, true>(lane, m, px, x, y, cs0, cs1, rm, false, 0.0);
                } else {
                    hit = pair_circle<G, false>(a, lane, m, px, x, y, cs0, 1, eg, ev, 0u);  // plan:381
                }
            } else {
                // plan:408-413: any pair closer than min_goal_dist (strict '<') rejects
#pragma unroll
                for (int k = 1; k <= G / 2; ++k) {
                    const int src = (int)(base | (unsigned)((m + k) & (G - 1)));
                    const double ox = __shfl_sync(FULL, x, src), oy = __shfl_sync(FULL, y, src);
                    const bool opart = __shfl_sync(FULL, (int)px, src) != 0;
                    const double dx = dsub(x, ox), dy = dsub(y, oy);
                    if (px && opart && sqrt_lt(dadd(dmul(dx, dx), dmul(dy, dy)), a.min_goal_dist)) hit = true;
                }
            }
        }
        const unsigned hitm = __ballot_sync(FULL, hit);
        const bool alive_grp = cand && (hitm & gmask) == 0u;
        bool bad = false;
        if (alive_grp && part) {  // plan:379 / 406 wall check with safety offset
            Rect rw;
            if (BOX) rect_vertices_axis(x, y, cw0, cw1, rw);
            bad = !wall_valid<BOX>(tb, a.L, x, y, cw0, rw);
            if (a.n_obst > 0 && !bad) {  // starts and goals clear every obstacle by the safety offset
                Rect rm;
                if (BOX) rect_vertices_axis(x, y, cs0, cs1, rm);
                bad = obstacle_hit_exact<BOX>(a, x, y, cs0, rm);
            }
        }
        const unsigned badm = __ballot_sync(FULL, bad);
        okmask[h] = __ballot_sync(FULL, alive_grp && (badm & gmask) == 0u && m == 0);
    }
    if (!(okmask[0] | okmask[1])) return false;
    // sequential order is t = t0, t0+1, ...: slot-major, half-minor
    const int s0 = okmask[0] ? (__ffs(okmask[0]) - 1) / G : 1 << 20;
    const int s1 = okmask[1] ? (__ffs(okmask[1]) - 1) / G : 1 << 20;
    const int hw = (2 * s0 <= 2 * s1 + 1) ? 0 : 1;
    const int sw = hw == 0 ? s0 : s1;
    const double wx = __shfl_sync(FULL, hw == 0 ? xs[0] : xs[1], sw * G + m);
    const double wy = __shfl_sync(FULL, hw == 0 ? ys[0] : ys[1], sw * G + m);
    out = make_double2(wx, wy);
    return true;
}

template <int G, bool BOX, int KIND>
__device__ __forceinline__ void sample_env_groups(const PlanArgs& a, const Tables& tb, unsigned lane, unsigned gmask_, uint32_t eg,
                                                  uint32_t ev, double2& out, bool& failed) {
    constexpr int S = 32 / G;  // attempts tested per half-iteration
    const int slot = (int)(lane / G);
    const int m = (int)(lane % G);
    const bool has_mover = m < a.N;
    const int mm = has_mover ? m : 0;
    const float rf = (float)a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 0];   // circle radius / box half size x (with offset)
    const float sf = (float)a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 1];   // box half size y
    const int cap = a.max_reset_attempts > 0 ? a.max_reset_attempts : 1;
    const unsigned base = lane & ~(unsigned)(G - 1);
    bool found = false;
#pragma unroll 1
    for (int t0 = 0; t0 < cap && !found; t0 += 2 * S) {
        const uint32_t blk = (uint32_t)(t0 / 2 + slot);
        const gpr_u32x4 r = sampler_block(a.seed, eg, ev, GPR_RNG_RESET_SAMPLE + 2u * blk + (uint32_t)KIND, (uint32_t)m);
        // ---- phase 1, float32: ~99% of the attempts die on a pair that is far inside the rejection band
        unsigned cand_bits = 0u, needx_bits = 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = 2 * (int)blk + h;
            const bool part = has_mover && t < cap;
            const float xf = part ? fmaf(a.spanxf, (float)r.v[2 * h] * 2.3283064365386963e-10f, a.minxf) : 1e30f;
            const float yf = part ? fmaf(a.spanyf, (float)r.v[2 * h + 1] * 2.3283064365386963e-10f, a.minyf) : 1e30f;
            bool rej = false, unc = false;
            if (G > 1) {
#pragma unroll
                for (int k = 1; k <= G / 2; ++k) {
                    const int src = (int)(base | (unsigned)((m + k) & (G - 1)));
                    const float oxf = __shfl_sync(FULL, xf, src), oyf = __shfl_sync(FULL, yf, src);
                    float lo2, hi2;
                    if (KIND == 1) {
                        lo2 = a.goal_lo2f;
                        hi2 = a.goal_hi2f;
                    } else if (BOX) {
                        // axis-aligned rectangles (identity quaternion, plan:364): a gap along x or y is a certain miss;
                        // overlap along both axes is a certain hit when all boxes have one size (no containment case,
                        // which geom:107-138 would not report) — the rest goes to the exact test
                        const float tx = rf + __shfl_sync(FULL, rf, src), ty = sf + __shfl_sync(FULL, sf, src);
                        const float mg = a.pair_mgf[0];
                        const float adx = fabsf(xf - oxf), ady = fabsf(yf - oyf);
                        const bool mine = part && oxf < 1e29f;
                        if (mine) {
                            if (adx > tx + mg || ady > ty + mg) {
                            } else if (a.uniform_pairs && adx < tx - mg && ady < ty - mg) {
                                rej = true;
                            } else {
                                unc = true;
                            }
                        }
                        continue;
                    } else if (a.uniform_pairs) {
                        lo2 = a.pair_lo2f[1][0];
                        hi2 = a.pair_hi2f[1][0];
                    } else {
                        const float tt = rf + __shfl_sync(FULL, rf, src), mg = a.pair_mgf[0];
                        lo2 = tt > mg ? (tt - mg) * (tt - mg) : -1.f;
                        hi2 = (tt + mg) * (tt + mg);
                    }
                    const bool mine = part && oxf < 1e29f;  // (pairs seen from both ends agree; no need to dedupe)
                    const float dx = xf - oxf, dy = yf - oyf;
                    const float d2 = dx * dx + dy * dy;
                    if (mine) {
                        if (d2 < lo2) rej = true;
                        else if (!(d2 > hi2)) unc = true;
                    }
                }
            }
            const unsigned rejm = __ballot_sync(FULL, rej), uncm = __ballot_sync(FULL, unc);
            const bool cand = (rejm & gmask_) == 0u && t < cap;
            cand_bits |= cand ? (1u << h) : 0u;
            needx_bits |= (cand && (uncm & gmask_) != 0u) ? (1u << h) : 0u;
        }
        // ---- phase 2, float64 (exact), only when some group of the warp still has a candidate
        if (!__any_sync(FULL, cand_bits != 0u)) continue;
        found = sample_confirm<G, BOX, KIND>(a, tb, lane, gmask_, cap, blk, r, cand_bits, needx_bits, eg, ev, out);
    }
    if (!found) {
        // the reference would loop forever (plan:369); keep the last attempt's sample and report the failure
        double ux, uy;
        gpr_sample_xy(a.seed, eg, ev, GPR_RNG_RESET_SAMPLE, (uint32_t)KIND, (uint32_t)(cap - 1), (uint32_t)m, &ux, &uy);
        out = make_double2(dadd(a.min_xy[0], dmul(a.span_xy[0], ux)), dadd(a.min_xy[1], dmul(a.span_xy[1], uy)));
        failed = true;
    }
}

// ---- rejection sampling, one lane per PAIR OF ATTEMPTS (G <= 8) ---------------------------------------------------------
// The same sequence of attempts as sample_env_groups (same Philox blocks: block (2*b + KIND, mover m) carries attempts 2b
// and 2b+1 of mover m), but each lane draws ALL movers of its two attempts and screens them without any shuffle: 64
// attempts per warp iteration.  The float32 screens (pair distances, wall_screen_f) give one of three verdicts per attempt:
// certainly rejected / certainly accepted / too close to call.  The first attempt in index order that is accepted wins;
// a too-close-to-call attempt that comes before any accepted one is decided exactly (float64, confirm_attempt) first.
// Verdict on ONE attempt that survived the pair screen; lanes 0..G-1 hold its mover m = lane (position x, y as the oracle
// computes it).  Warp-collective.  pairs_certain: the float32 pair screen already proved that no pair is too close;
// otherwise the pairs are re-tested exactly (plan:381 / 408-413).  The wall check with the safety offset (plan:379 / 406)
// is made here for every candidate: float32 screen on the exact position, exact float64 test when too close to call.
template <int G, bool BOX, int KIND>
__device__ __forceinline__ bool confirm_attempt(const PlanArgs& a, const Tables& tb, unsigned lane, double x, double y,
                                                bool pairs_certain, uint32_t eg, uint32_t ev) {
    const int m = (int)(lane % G);
    const bool part = lane < (unsigned)G && m < a.N;
    const int mm = part ? m : 0;
    const double cw0 = a.c_wall[(GPR_MAX_MOVERS + mm) * 2 + 0], cw1 = a.c_wall[(GPR_MAX_MOVERS + mm) * 2 + 1];
    bool hit = false;
    if (G > 1 && !pairs_certain) {  // (uniform branch)
        const double cs0 = a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 0], cs1 = a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 1];
        if (KIND == 0) {
            if (BOX) {
                Rect rm;
                rect_vertices_axis(x, y, cs0, cs1, rm);
                hit = pair_check<G, true>(lane, m, part, x, y, cs0, cs1, rm, false, 0.0);
            } else {
                hit = pair_circle<G, false>(a, lane, m, part, x, y, cs0, 1, eg, ev, 0u);
            }
        } else {
            const unsigned base = lane & ~(unsigned)(G - 1);
#pragma unroll
            for (int k = 1; k <= G / 2; ++k) {
                const int src = (int)(base | (unsigned)((m + k) & (G - 1)));
                const double ox = __shfl_sync(FULL, x, src), oy = __shfl_sync(FULL, y, src);
                const bool opart = __shfl_sync(FULL, (int)part, src) != 0;
                const double dx = dsub(x, ox), dy = dsub(y, oy);
                if (part && opart && sqrt_lt(dadd(dmul(dx, dx), dmul(dy, dy)), a.min_goal_dist)) hit = true;
            }
        }
    }
    bool bad = false;
    if (part) {
        int gi, gj;
        guess_cell(a, x, y, gi, gj);
        float clear;
        // box shape: the mover's rectangle is axis-aligned here (identity quaternion, no noise), so the rectangle screen
        // with its own half sizes is exact whenever it is certain (an overlap with a missing cell makes a vertex unsafe
        // towards a missing neighbour; no overlap leaves nothing for the inner-corner edge test to find)
        const int f = wall_fast(a, tb, x, y, (float)cw0, BOX ? (float)cw1 : (float)cw0, gi, gj, clear);
        if (f == 2) {
            Rect rw;
            if (BOX) rect_vertices_axis(x, y, cw0, cw1, rw);
            bad = !wall_valid<BOX>(tb, a.L, x, y, cw0, rw);
        } else {
            bad = f == 0;
        }
        // starts and goals clear every obstacle by the safety offset (exact, no noise inside the sampling loop)
        if (a.n_obst > 0 && !bad) {
            const double cs0 = a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 0], cs1 = a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 1];
            Rect rm;
            if (BOX) rect_vertices_axis(x, y, cs0, cs1, rm);
            bad = obstacle_hit_exact<BOX>(a, x, y, cs0, rm);
        }
    }
    return __ballot_sync(FULL, hit || bad) == 0u;
}

#ifndef GPR_LANES_MAX_G
#define GPR_LANES_MAX_G 4  // widest lane group sampled with one lane per attempt pair (wider ones: one group per attempt;
                           // measured on B200: G=8 gains nothing from it — 32 Philox words per lane spill)
#endif
template <int G, bool BOX, int KIND>
__device__ __forceinline__ void sample_env_lanes(const PlanArgs& a, const Tables& tb, unsigned lane, uint32_t eg, uint32_t ev,
                                                 double2& out, bool& failed, int kind_rt = 0) {
    // KIND < 0 (circle shape only): the kind is a run-time argument, so that starts and goals share ONE copy of this loop
    // (instruction-cache footprint of the auto-reset kernel); they differ in thresholds and in the exact confirmation only
    static_assert(G <= 8, "one lane draws all movers: register budget");
    static_assert(KIND >= 0 || !BOX, "run-time kind: circle shape only");
    const int kind = KIND < 0 ? kind_rt : KIND;
    const int N = a.N;
    const int cap = a.max_reset_attempts > 0 ? a.max_reset_attempts : 1;
    const int m_self = (int)(lane % G);
    // collision sizes with the safety offset (plan:381), as floats for the pair screen
    float rf[G], sf[G];
#pragma unroll
    for (int m = 0; m < G; ++m) {
        const int mm = m < N ? m : 0;
        rf[m] = (float)a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 0];
        sf[m] = (float)a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 1];
    }
    const float mg = a.pair_mgf[0];
    // span * 2^-32: scaling by a power of two is exact, so fmaf(span * 2^-32, w, min) IS fmaf(span, w * 2^-32, min) bit for bit
    const float sxs = a.spanxf * 2.3283064365386963e-10f, sys = a.spanyf * 2.3283064365386963e-10f;
    const bool uni = kind == 1 || a.uniform_pairs != 0;
    const float lo2u = kind == 1 ? a.goal_lo2f : a.pair_lo2f[1][0], hi2u = kind == 1 ? a.goal_hi2f : a.pair_hi2f[1][0];
    bool found = false;
#pragma unroll 1
    for (int t0 = 0; t0 < cap && !found; t0 += 64) {
        const uint32_t blk = (uint32_t)(t0 / 2) + lane;
        gpr_u32x4 r[G];
#pragma unroll
        for (int m = 0; m < G; ++m) {
            if (m < N) r[m] = sampler_block(a.seed, eg, ev, GPR_RNG_RESET_SAMPLE + 2u * blk + (uint32_t)kind, (uint32_t)m);
            else r[m].v[0] = r[m].v[1] = r[m].v[2] = r[m].v[3] = 0u;
        }
        unsigned accb = 0u, uncb = 0u;  // this lane's pair-screen verdicts for its attempts h = 0, 1
#if GPR_SAMPLER_ROLL_H
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int h = 0; h < 2; ++h) {
            float xf[G], yf[G];
#pragma unroll
            for (int m = 0; m < G; ++m) {
                // (lanes of a group wider than num_movers: park the spare movers far away from everything, no branches)
                const uint32_t wx = h ? r[m].v[2] : r[m].v[0], wy = h ? r[m].v[3] : r[m].v[1];
                xf[m] = m < N ? fmaf(sxs, (float)wx, a.minxf) : 1e8f * (float)(m + 1);
                yf[m] = m < N ? fmaf(sys, (float)wy, a.minyf) : 0.f;
            }
            bool rej = false, unc = false;
            if (!(BOX && kind == 0) && uni) {
                // one threshold pair for every pair of movers: classify the SMALLEST squared distance once (a min per pair
                // instead of two compares and their logic) — rejected below lo2, too close to call up to hi2
                float dmin = 3.0e38f;
#pragma unroll
                for (int i = 0; i < G; ++i) {
#pragma unroll
                    for (int j = i + 1; j < G; ++j) {
                        const float dx = xf[i] - xf[j], dy = yf[i] - yf[j];
                        dmin = fminf(dmin, dx * dx + dy * dy);
                    }
                }
                rej = dmin < lo2u;
                unc = !(dmin > hi2u);
            } else {
#pragma unroll
            for (int i = 0; i < G; ++i) {
#pragma unroll
                for (int j = i + 1; j < G; ++j) {
                    const float dx = xf[i] - xf[j], dy = yf[i] - yf[j];
                    if (BOX && kind == 0) {
                        // axis-aligned rectangles: a gap along x or y is a certain miss, overlap along both axes a certain
                        // hit when all boxes have one size (else containment, which geom:107-138 does not report, is possible)
                        const float tx = rf[i] + rf[j], ty = sf[i] + sf[j];
                        const float adx = fabsf(dx), ady = fabsf(dy);
                        const bool miss = adx > tx + mg || ady > ty + mg;
                        const bool hitc = uni && adx < tx - mg && ady < ty - mg;
                        rej = rej || hitc;
                        unc = unc || (!miss && !hitc);
                    } else {
                        const float tt = rf[i] + rf[j];
                        const float lo2 = uni ? lo2u : (tt > mg ? (tt - mg) * (tt - mg) : -1.f);
                        const float hi2 = uni ? hi2u : (tt + mg) * (tt + mg);
                        const float d2 = dx * dx + dy * dy;
                        rej = rej || d2 < lo2;
                        unc = unc || !(d2 < lo2 || d2 > hi2);
                    }
                }
            }
            }
            const bool valid = 2 * (int)blk + h < cap;
            if (valid && !rej) {
                if (unc) uncb |= 1u << h;
                else accb |= 1u << h;
            }
        }
        // ---- candidates in index order (lane-major, h-minor): the first one that also passes the wall check wins
        unsigned acc0 = __ballot_sync(FULL, accb & 1u), acc1 = __ballot_sync(FULL, accb & 2u);
        unsigned unc0 = __ballot_sync(FULL, uncb & 1u), unc1 = __ballot_sync(FULL, uncb & 2u);
        while (!found && (acc0 | acc1 | unc0 | unc1)) {
            const int l0 = __ffs(acc0 | acc1 | unc0 | unc1) - 1;
            const unsigned bit = 1u << l0;
            const int h0 = ((acc0 | unc0) & bit) ? 0 : 1;
            const bool certain = ((h0 == 0 ? acc0 : acc1) & bit) != 0u;
            // positions of that attempt -> lane m holds mover m (every lane: m = lane % G, only lanes < G matter)
            double px = 0.0, py = 0.0;
#pragma unroll
            for (int m = 0; m < G; ++m) {
                const uint32_t wx = __shfl_sync(FULL, h0 ? r[m].v[2] : r[m].v[0], l0);
                const uint32_t wy = __shfl_sync(FULL, h0 ? r[m].v[3] : r[m].v[1], l0);
                if (m == m_self) {
                    px = dadd(a.min_xy[0], dmul(a.span_xy[0], gpr_uniform32(wx)));  // plan:377/405
                    py = dadd(a.min_xy[1], dmul(a.span_xy[1], gpr_uniform32(wy)));
                }
            }
            bool ok;
            if constexpr (KIND < 0) {
                ok = kind == 0 ? confirm_attempt<G, BOX, 0>(a, tb, lane, px, py, certain, eg, ev)
                               : confirm_attempt<G, BOX, 1>(a, tb, lane, px, py, certain, eg, ev);
            } else {
                ok = confirm_attempt<G, BOX, KIND>(a, tb, lane, px, py, certain, eg, ev);
            }
            if (ok) {
                out = make_double2(px, py);
                found = true;
            } else if (h0 == 0) {
                acc0 &= ~bit;
                unc0 &= ~bit;
            } else {
                acc1 &= ~bit;
                unc1 &= ~bit;
            }
        }
    }
    if (!found) {
        // the reference would loop forever (plan:369); keep the last attempt's sample and report the failure
        double ux, uy;
        gpr_sample_xy(a.seed, eg, ev, GPR_RNG_RESET_SAMPLE, (uint32_t)kind, (uint32_t)(cap - 1), (uint32_t)m_self, &ux, &uy);
        out = make_double2(dadd(a.min_xy[0], dmul(a.span_xy[0], ux)), dadd(a.min_xy[1], dmul(a.span_xy[1], uy)));
        failed = true;
    }
}

// One environment (global index eg, RNG event ev), whole warp: on return EVERY lane holds, for mover m = lane % G, the
// position of the first accepted attempt (or of the last attempt, with failed = true, if the cap was hit).
template <int G, bool BOX, int KIND>
__device__ __forceinline__ void sample_env(const PlanArgs& a, const Tables& tb, unsigned lane, unsigned gmask_, uint32_t eg,
                                           uint32_t ev, double2& out, bool& failed) {
    if constexpr (G <= GPR_LANES_MAX_G) {
        sample_env_lanes<G, BOX, KIND>(a, tb, lane, eg, ev, out, failed);
    } else {
        sample_env_groups<G, BOX, KIND>(a, tb, lane, gmask_, eg, ev, out, failed);
    }
}

template <int G, bool BOX, int KIND>
__device__ __forceinline__ void sample_positions(const PlanArgs& a, const Tables& tb, const Lane<G>& ln, bool need,
                                                 uint32_t event, double2& out, bool& failed) {
    unsigned todo = __ballot_sync(FULL, need && ln.m == 0);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t eg = __shfl_sync(FULL, ln.env_global, leader);
        const uint32_t ev = __shfl_sync(FULL, event, leader);
        const bool target = (ln.lane / G) == (unsigned)(leader / G);
        double2 res;
        bool f = false;
        sample_env<G, BOX, KIND>(a, tb, ln.lane, ln.gmask, eg, ev, res, f);
        if (target) {
            out = res;
            failed = failed || f;
        }
    }
}

// basic:1799-1805: wall check WITH the safety offset, mover check WITHOUT, on independently noisy qpos (warp-collective)
template <int G, bool BOX, bool NOISE>
__device__ __forceinline__ void reset_checks(const PlanArgs& a, const Tables& tb, const Lane<G>& ln, bool need,
                                             uint32_t event, double2 p, bool& mc, bool& wc, bool& oc) {
    const int mm = ln.active ? ln.m : 0;
    const double cw0 = a.c_wall[(GPR_MAX_MOVERS + mm) * 2 + 0], cw1 = a.c_wall[(GPR_MAX_MOVERS + mm) * 2 + 1];
    const double cm0 = a.c_mover[mm * 2 + 0], cm1 = a.c_mover[mm * 2 + 1];
    const bool part = need && ln.active;
    bool bad, hit;
    if (!BOX) {
        {
            int gi, gj;
            guess_cell(a, p.x, p.y, gi, gj);
            float n4[4];
            bool have = false;
            float mg_unused;
            bad = wall_bad_circle<NOISE>(a, tb, part, p.x, p.y, cw0, (float)cw0, gi, gj, ln.env_global, event,
                                         GPR_RNG_RESET_CHECK, ln.m, 0, n4, have, mg_unused);
        }
        // the mover-check noise of reset() lives in words 2,3 of the same block
        hit = false;
        if (G > 1) {
            const unsigned base = ln.lane & ~(unsigned)(G - 1);
#pragma unroll
            for (int k = 1; k <= G / 2; ++k) {
                const int pm = (ln.m + k) & (G - 1);
                const int src = (int)(base | (unsigned)pm);
                const double ox = __shfl_sync(FULL, p.x, src), oy = __shfl_sync(FULL, p.y, src);
                const double orr = __shfl_sync(FULL, cm0, src);
                const bool opart = __shfl_sync(FULL, (int)part, src) != 0;
                const bool mine = part && opart && !(k == G / 2 && ln.m >= G / 2);
                const double t = a.uniform_pairs ? a.pair_t[0] : dadd(cm0, orr);
                const double mg = NOISE ? a.pair_margin : 0.0;
                const double dx = dsub(p.x, ox), dy = dsub(p.y, oy);
                const double d2 = dadd(dmul(dx, dx), dmul(dy, dy));
                const double tl = t - mg, th = t + mg;
                if (mine) {
                    if (tl > 0.0 && d2 <= tl * tl * (1.0 - 1e-14)) {
                        hit = true;
                    } else if (!(d2 >= th * th * (1.0 + 1e-14))) {
                        double xi = p.x, yi = p.y, xj = ox, yj = oy;
                        if (NOISE) {
                            float k4[4];
                            normal4_cold(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK, (uint32_t)ln.m, k4);
                            xi = noisy(p.x, k4[2], a.sigma_p);
                            yi = noisy(p.y, k4[3], a.sigma_p);
                            normal4_cold(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK, (uint32_t)pm, k4);
                            xj = noisy(ox, k4[2], a.sigma_p);
                            yj = noisy(oy, k4[3], a.sigma_p);
                        }
                        const double ex = dsub(xi, xj), ey = dsub(yi, yj);
                        if (dsqrt(dadd(dmul(ex, ex), dmul(ey, ey))) <= t) hit = true;
                    }
                }
            }
        }
    } else {
        // float32 screens first (bounding rectangle for the walls, axis gaps for the pairs — the mover is axis-aligned up to
        // the sensor noise on its quaternion, see planning_step_kernel); exact tests only where they cannot decide
        const float ext_w = a.rot_extf * (float)(cw0 + cw1), ext_m = a.rot_extf * (float)(cm0 + cm1);
        int f = 1;
        if (part) {
            int gi, gj;
            guess_cell(a, p.x, p.y, gi, gj);
            float clear;
            f = wall_fast(a, tb, p.x, p.y, (float)cw0 + ext_w, (float)cw1 + ext_w, gi, gj, clear);
        }
        const bool wneed = part && f != 1;  // (0 is not a verdict: the bounding rectangle is conservative)
        bool near, sure;
        float clear_p;
        unsigned kmask;
        pair_box_screen<G>(ln.lane, ln.m, part, p.x, p.y, (float)cm0 + ext_m, (float)cm1 + ext_m, a.pair_mgf[NOISE ? 1 : 0], near,
                           clear_p, kmask, sure, a.uniform_pairs ? (float)cm0 : 0.f, (float)cm1, 2.f * ext_m + a.pair_mgf[NOISE ? 1 : 0]);
        bad = false;
        hit = sure;
        const bool wany = __any_sync(FULL, wneed), pany = G > 1 && __any_sync(FULL, near);
        if (wany || pany) {
            double wx = p.x, wy = p.y, mx = p.x, my = p.y;
            Rect rw, rm;
            if (NOISE) {
                float n4[4];
                normal4_cold(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK, (uint32_t)ln.m, n4);
                wx = noisy(p.x, n4[0], a.sigma_p);
                wy = noisy(p.y, n4[1], a.sigma_p);
                mx = noisy(p.x, n4[2], a.sigma_p);
                my = noisy(p.y, n4[3], a.sigma_p);
                float q[4];
                if (wany) {
                    normal4_cold(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK_WQUAT, (uint32_t)ln.m, q);
                    rect_vertices(wx, wy, noisy(1.0, q[0], a.sigma_p), dmul((double)q[1], a.sigma_p), dmul((double)q[2], a.sigma_p),
                                  dmul((double)q[3], a.sigma_p), cw0, cw1, rw);
                }
                if (pany) {
                    normal4_cold(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK_MQUAT, (uint32_t)ln.m, q);
                    rect_vertices(mx, my, noisy(1.0, q[0], a.sigma_p), dmul((double)q[1], a.sigma_p), dmul((double)q[2], a.sigma_p),
                                  dmul((double)q[3], a.sigma_p), cm0, cm1, rm);
                }
            } else {
                rect_vertices_axis(wx, wy, cw0, cw1, rw);
                rect_vertices_axis(mx, my, cm0, cm1, rm);
            }
            if (wneed) bad = !wall_valid<true>(tb, a.L, wx, wy, cw0, rw);
            if (pany) hit = pair_check<G, true>(ln.lane, ln.m, part, mx, my, cm0, cm1, rm, false, 0.0, kmask) || sure;
        }
    }
    // basic:1807 the hook: static obstacles on the wall check's noisy qpos, with the safety offset like that wall check
    bool obad = false;
    if (a.n_obst > 0 && part) {
        const double cs0 = a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 0], cs1 = a.c_mover[(GPR_MAX_MOVERS + mm) * 2 + 1];
        const float ext = BOX ? a.rot_extf * (float)(cs0 + cs1) : 0.f;
        float clear_o;
        const int f = obstacle_screen<BOX>(a, p.x, p.y, (float)cs0 + ext, (float)cs1 + ext, clear_o);
        if (f == 2) {
            double wx = p.x, wy = p.y;
            Rect rm;
            if (NOISE) {
                float w4[4];
                normal4_cold(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK, (uint32_t)ln.m, w4);
                wx = noisy(p.x, w4[0], a.sigma_p);
                wy = noisy(p.y, w4[1], a.sigma_p);
                if (BOX) {
                    float q[4];
                    normal4_cold(a.seed, ln.env_global, event, GPR_RNG_RESET_CHECK_WQUAT, (uint32_t)ln.m, q);
                    rect_vertices(wx, wy, noisy(1.0, q[0], a.sigma_p), dmul((double)q[1], a.sigma_p), dmul((double)q[2], a.sigma_p),
                                  dmul((double)q[3], a.sigma_p), cs0, cs1, rm);
                }
            } else if (BOX) {
                rect_vertices_axis(wx, wy, cs0, cs1, rm);
            }
            obad = obstacle_hit_exact<BOX>(a, wx, wy, cs0, rm);
        } else {
            obad = f == 1;
        }
    }
    const bool wnow = (__ballot_sync(FULL, bad) & ln.gmask) != 0u;
    const bool mnow = (__ballot_sync(FULL, hit) & ln.gmask) != 0u;
    const bool onow = (__ballot_sync(FULL, obad) & ln.gmask) != 0u;
    if (need) {
        wc = wnow;
        mc = mnow;
        oc = onow;
    }
}

// plan:355-418 + basic:1797-1805 for the envs with `need` set; warp-collective (every lane of the warp calls it).
template <int G, bool BOX, bool NOISE>
__device__ __forceinline__ void reset_group(const PlanArgs& a, const Tables& tb, const Lane<G>& ln, bool need,
                                            uint32_t event, const double2* inj_start, const double2* inj_goal,
                                            double2& p, double2& v, double2& acc, double2& goal, bool& mc, bool& wc,
                                            bool& oc, bool& failed) {
    failed = false;
    if (need && inj_start != nullptr && ln.active) p = inj_start[ln.idx];
    if (need && inj_goal != nullptr && ln.active) goal = inj_goal[ln.idx];
    sample_positions<G, BOX, 0>(a, tb, ln, need && inj_start == nullptr, event, p, failed);
    sample_positions<G, BOX, 1>(a, tb, ln, need && inj_goal == nullptr, event, goal, failed);
    failed = (__ballot_sync(FULL, failed) & ln.gmask) != 0u;

    // ---- fresh MjData (plan:336-353): qvel = act = qacc = 0
    if (need) {
        v = make_double2(0.0, 0.0);
        acc = make_double2(0.0, 0.0);
    }
    reset_checks<G, BOX, NOISE>(a, tb, ln, need, event, p, mc, wc, oc);
}

// ---- box shape, step kernel: the exact fallbacks of the float32 screens as ONE out-of-line function each.  Inlined they put
// ~4 Philox + Box-Muller expansions, two quat2mat rectangles, wall_valid<true> and the shuffles of pair_check into the
// 40-cycle loop: the box step kernel was 167 KB of code with "no instruction" as a top stall reason (1.97 per issued
// instruction, profiles/r2_full_planning8box.txt).  Measured on B200 (planning8box, 262,144 envs): out of line 1.96 ms per
// step-kernel launch, inlined 1.91 ms — the calls cost more than the cache misses they save, so inlined stays the default
// (GPR_BOX_OOL=1 builds the out-of-line form).
#ifndef GPR_BOX_OOL
#define GPR_BOX_OOL 0
#endif
#if GPR_BOX_OOL
#define GPR_BOX_COLD static __device__ __noinline__
#else
#define GPR_BOX_COLD __device__ __forceinline__
#endif
// basic:1888-1894 for one box-shaped mover, exact: noisy position + noisy quaternion -> vertices -> wall_valid
template <bool NOISE>
GPR_BOX_COLD bool box_wall_exact(const PlanArgs& a, const Tables& tb, double px, double py, double cw0, double cw1,
                                 uint32_t env_global, uint32_t event, uint32_t s0, uint32_t m) {
    double wx = px, wy = py;
    Rect rw;
    if (NOISE) {
        float n4[4], q[4];
        normal4_cold(a.seed, env_global, event, s0 + GPR_RNG_BLOCK_VEL_WALL, m, n4);
        wx = noisy(px, n4[2], a.sigma_p);
        wy = noisy(py, n4[3], a.sigma_p);
        normal4_cold(a.seed, env_global, event, s0 + GPR_RNG_BLOCK_WALL_QUAT, m, q);
        rect_vertices(wx, wy, noisy(1.0, q[0], a.sigma_p), dmul((double)q[1], a.sigma_p), dmul((double)q[2], a.sigma_p),
                      dmul((double)q[3], a.sigma_p), cw0, cw1, rw);
    } else {
        rect_vertices_axis(wx, wy, cw0, cw1, rw);
    }
    return !wall_valid<true>(tb, a.L, wx, wy, cw0, rw);
}
// basic:1895-1901 for the box shape, exact, warp-collective (EVERY lane of the warp calls it): noisy poses -> vertices ->
// the rectangle test of the pairs in kmask
template <int G, bool NOISE>
GPR_BOX_COLD bool box_pair_exact(const PlanArgs& a, unsigned lane, int m, bool part, double px, double py, double cm0, double cm1,
                                 uint32_t env_global, uint32_t event, uint32_t s0, unsigned kmask) {
    double mx = px, my = py;
    Rect rm;
    if (NOISE) {
        float k4[4], q[4];
        normal4_cold(a.seed, env_global, event, s0 + GPR_RNG_BLOCK_MOVER, (uint32_t)m, k4);
        mx = noisy(px, k4[0], a.sigma_p);
        my = noisy(py, k4[1], a.sigma_p);
        normal4_cold(a.seed, env_global, event, s0 + GPR_RNG_BLOCK_MOVER_QUAT, (uint32_t)m, q);
        rect_vertices(mx, my, noisy(1.0, q[0], a.sigma_p), dmul((double)q[1], a.sigma_p), dmul((double)q[2], a.sigma_p),
                      dmul((double)q[3], a.sigma_p), cm0, cm1, rm);
    } else {
        rect_vertices_axis(mx, my, cm0, cm1, rm);
    }
    return pair_check<G, true>(lane, m, part, mx, my, cm0, cm1, rm, false, 0.0, kmask);
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

#ifndef GPR_STEP_UNROLL
#define GPR_STEP_UNROLL 1
#endif
#define GPR_PRAGMA_(x) _Pragma(#x)
#define GPR_UNROLL(n) GPR_PRAGMA_(unroll n)
#ifndef GPR_STEP_MINB
#define GPR_STEP_MINB 3  // resident threads per SM the step kernel is compiled for, in units of 256 (circle shape)
#endif
#ifndef GPR_STEP_MINB_BOX
#define GPR_STEP_MINB_BOX 2  // same for the box shape (more live state: a tighter register cap spills)
#endif
#ifndef GPR_AR_MINB
#define GPR_AR_MINB 3  // CTAs of 128 threads per SM the auto-reset kernel is compiled for (circle shape)
#endif
#ifndef GPR_AR_MINB_BOX
#define GPR_AR_MINB_BOX 4
#endif
// measured on B200 (planning4 at 65,536 / 1,048,576 envs, planning8box): step kernel 256 threads x 2 CTAs/SM 0.136 / 1.56 ms,
// 128 x 6 (80 registers) 0.118 / 1.22 ms, 64 x 12 0.116 / 1.21 ms, 128 x 8 (64 registers) 0.121 / 1.25 ms; the box shape
// spills under any cap below 128 registers (2.14 -> 2.35 ms).  Auto-reset kernel: 3 / 4 / 6 / 8 CTAs per SM 0.108 / 0.112 /
// 0.112 / 0.122 ms (circle), 0.79 / 0.74 / 0.74 / 0.73 ms (box).
// Tried on top of that and measured slower (planning4, step kernel 0.111 ms): refreshing the wall budget of EVERY mover of a
// warp whenever one of them is due (0.117 ms: more lanes reach the exact fallback), and "blind runs" — k cycles of bare
// integration without per-cycle budget tests, k from a closed-form bound of the travel (0.134 ms: with 32 movers per warp
// one of them is nearly always within a few cycles of its wall budget, so k is 0 or 1 and its computation is overhead).
#ifndef GPR_PAIR_ENV_MARGIN
#define GPR_PAIR_ENV_MARGIN 0  // 1: one pair budget per env (the smallest margin of any of its pairs) instead of per mover
#endif
#ifndef GPR_WALL_LINF
#define GPR_WALL_LINF 1
#endif
#ifndef GPR_STEP_STASH
#define GPR_STEP_STASH 1
#endif
#ifndef GPR_STEP_THREADS
#define GPR_STEP_THREADS 128
#endif
// threads per CTA of planning_step_kernel
template <int G>
struct StepThreads {
    static constexpr int value = GPR_STEP_THREADS;
};

#ifdef GPR_STEP_CTAS  // tuning experiments: resident CTAs per SM of the circle-shape step kernel, given directly
#define GPR_STEP_MINBLOCKS(BOX, T) ((BOX) ? GPR_STEP_MINB_BOX * 256 / (T) : GPR_STEP_CTAS)
#else
#define GPR_STEP_MINBLOCKS(BOX, T) (((BOX) ? GPR_STEP_MINB_BOX : GPR_STEP_MINB) * 256 / (T))
#endif
template <int G, bool BOX, bool NOISE, bool JERK, bool EXTRA>
__global__ void __launch_bounds__(StepThreads<G>::value, GPR_STEP_MINBLOCKS(BOX, StepThreads<G>::value))
    planning_step_kernel(const __grid_constant__ PlanArgs a) {
    // the auto-reset kernel that follows in the stream may become resident as soon as every CTA of this grid has started
    // (programmatic dependent launch): its warps then fill the SM slots the last, partial wave of this grid leaves idle
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ Tables tb;
    __shared__ unsigned s_signed_off;  // warps of this CTA that have dropped their per-env results into shared memory
    load_tables(tb, a.L);
    if (threadIdx.x == 0) s_signed_off = 0u;
    __syncthreads();
    const Lane<G> ln = make_lane<G>(a);
    const int mm = ln.active ? ln.m : 0;
    const double cw0 = a.c_wall[mm * 2 + 0], cw1 = a.c_wall[mm * 2 + 1];  // safety = 0 (basic:1888-1901)
    const double cm0 = a.c_mover[mm * 2 + 0], cm1 = a.c_mover[mm * 2 + 1];
    const float cw0f = (float)cw0;

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
        // desired_goal does not change during a step: store it now, so that (when the output lives in pinned host memory)
        // this part of the result traffic crosses PCIe under the 40-cycle loop instead of in the burst at the end
        if (!pending_reset && a.write_goal && a.out.desired_goal) store_pair(EXTRA && a.out_f64, a.out.desired_goal, ln.idx, goal.x, goal.y);
    }

    // ------------------------------------------------------------------ the 40-cycle loop (basic:1879-1905)
    // TEMPORAL COHERENCE.  A collision check leaves every mover with a certified clearance: the distance it may still
    // travel before the wall check can turn invalid (wall_fast), and the distance the movers of its env may still travel
    // before a pair can start to collide, noise included.  A mover moves dt*|v| <= 2 mm per cycle, so while its accumulated
    // travel stays below a budget, the reference's check is known to report "no collision" and is not evaluated.  The wall
    // check is lane-local and runs for the lanes whose wall budget is used up; the pair check is warp-collective and
    // runs for the warp as soon as one of its envs has used up its pair budget.
    bool alive = ln.env_ok && !pending_reset;
    bool mc = false, wc = false, oc = false;
    int gi = 0, gj = 0;  // tile cell under the mover, tracked across cycles (movement per cycle is ~mm)
    guess_cell(a, p.x, p.y, gi, gj);
    // bounding rectangle / circle used by the float32 screens of the box shape (planning movers never rotate: the only
    // rotation is the sensor noise on the quaternion)
    const float bxf = BOX ? (float)cw0 + a.rot_extf * (float)(cw0 + cw1) : cw0f;
    const float byf = BOX ? (float)cw1 + a.rot_extf * (float)(cw0 + cw1) : cw0f;
    const float pxf = BOX ? (float)cm0 + a.rot_extf * (float)(cm0 + cm1) : 0.f;  // same for the mover-check sizes
    const float pyf = BOX ? (float)cm1 + a.rot_extf * (float)(cm0 + cm1) : 0.f;
    float travel = 0.f;                  // sum over cycles of an upper bound of |v|  (distance / dt)
    // GPR_WALL_LINF: the wall clearance is an L-infinity distance, so the wall budget is charged max(|vx|, |vy|) instead of
    // the Euclidean speed bound (measured: circle shape +2 %, step kernel 0.110 -> 0.106 ms; the register-bound box kernel
    // loses 1 % to the second accumulator and keeps the single one)
    constexpr bool LINF = GPR_WALL_LINF && !BOX;
    float travel_w = 0.f;
    float lim_w = -1.f, lim_p = -1.f;    // `travel` values up to which the wall / pair check is certified negative
    bool any_alive = __any_sync(FULL, alive);
    // `part`: this lane integrates a mover of a live env.  GPR_STEP_STASH: a lane whose env has collided parks its final
    // (p, v, acc) in shared memory and the common path below then runs UNPREDICATED for every lane (a parked lane computes
    // garbage that nothing reads: it takes part in no check and is restored after the loop) — the integration writes
    // straight into the loop-carried registers instead of into temporaries that are copied under a predicate.
    // (measured on B200: circle shape +1 %, box shape -4 % — the box kernel is register-bound — so circle only)
    constexpr bool STASH = GPR_STEP_STASH && !BOX;
    bool part = alive && ln.active;
    bool died = false;
    __shared__ double s_stash[STASH ? 6 : 1][STASH ? StepThreads<G>::value : 1];
    GPR_UNROLL(GPR_STEP_UNROLL)
    for (int cyc = 0; cyc < a.num_cycles && any_alive; ++cyc) {
        const uint32_t s0 = (uint32_t)cyc * 4u;
        float n4[4] = {0.f, 0.f, 0.f, 0.f};
        bool have0 = false;  // block 0 of this cycle generated?
        // ---- plan:420-450 _mujoco_step_callback + mj_step
        // Hot path, decided for the whole warp: every mover "free-runs" — neither ensure_max_dyn_val clip can trigger and
        // the velocity noise (plan:430) cannot matter (|v + dt*a| stays below v_max minus the noise bound), so
        // ensure_max_dyn_val passes its inputs through: qacc = a (acc mode) or act + dt*j (jerk mode, integrator actuator
        // with actearly, plan:305-311), qvel += dt*qacc.  The sums below are the very expressions the general path
        // evaluates (same operands, same two roundings), so both paths produce identical bits.
        const double nax = JERK ? dadd(dmul(a.dt, u.x), acc.x) : u.x;
        const double nay = JERK ? dadd(dmul(a.dt, u.y), acc.y) : u.y;
        const double tx = dadd(dmul(a.dt, nax), v.x), ty = dadd(dmul(a.dt, nay), v.y);
        bool free_run = dadd(dmul(tx, tx), dmul(ty, ty)) < (NOISE ? a.v_lazy2 : a.v_max2_lo);
        if (JERK) free_run = free_run && dadd(dmul(nax, nax), dmul(nay, nay)) < a.a_max2_lo;
        if (!__any_sync(FULL, part && !free_run)) {
            if (STASH || part) {
                acc.x = nax;
                acc.y = nay;
                v.x = tx;
                v.y = ty;
            }
        } else if (part) {
            // general path; d = derivative entering the velocity clip (action or limited acc)
            double dxv = u.x, dyv = u.y, jx = 0.0, jy = 0.0;
            if (JERK) ensure_max(acc.x, acc.y, a.a_max, a.a_max2_lo, u.x, u.y, a.dt, a.inv_dt, dxv, dyv, jx, jy);  // plan:434
            double velx = v.x, vely = v.y;
            if (NOISE) {
                // the velocity noise can only matter if the un-noised |dt*d + v| is within its bound of v_max
                const double sx = dadd(dmul(a.dt, dxv), v.x), sy = dadd(dmul(a.dt, dyv), v.y);
                if (!(dadd(dmul(sx, sx), dmul(sy, sy)) < a.v_lazy2)) {
                    normal4_cold(a.seed, ln.env_global, event, s0 + GPR_RNG_BLOCK_VEL_WALL, (uint32_t)ln.m, n4);
                    have0 = true;
                    velx = noisy(v.x, n4[0], a.sigma_v);
                    vely = noisy(v.y, n4[1], a.sigma_v);
                }
            }
            double t0, t1, ax, ay;
            ensure_max(velx, vely, a.v_max, a.v_max2_lo, dxv, dyv, a.dt, a.inv_dt, t0, t1, ax, ay);  // plan:437 / 442
            if (JERK) {
                if (dxv != ax || dyv != ay) {  // plan:438
                    jx = ddiv_rcp(dsub(ax, acc.x), a.dt, a.inv_dt);
                    jy = ddiv_rcp(dsub(ay, acc.y), a.dt, a.inv_dt);
                }
                acc.x = dadd(acc.x, dmul(a.dt, jx));  // act += dt*ctrl; qacc = act
                acc.y = dadd(acc.y, dmul(a.dt, jy));
            } else {
                acc.x = ax;  // dyntype none, gain = mass (plan:314-320): qacc = ctrl
                acc.y = ay;
            }
            v.x = dadd(v.x, dmul(a.dt, acc.x));  // semi-implicit Euler (MuJoCo): qvel += dt*qacc
            v.y = dadd(v.y, dmul(a.dt, acc.y));
        }
        if (STASH || part) {
            p.x = dadd(p.x, dmul(a.dt, v.x));  // qpos += dt*qvel
            p.y = dadd(p.y, dmul(a.dt, v.y));
            // |v| <= max + min/2 of the absolute components; 1e-4 covers the float roundings of the running sum
            const float avx = fabsf((float)v.x), avy = fabsf((float)v.y);
            // (obst_vmaxf: a moving extra body closes in on its own account; it uses up the shared wall / obstacle budget too)
            travel += (fmaxf(avx, avy) + 0.5f * fminf(avx, avy) + a.obst_vmaxf) * 1.0001f;
            if (LINF) travel_w += (fmaxf(avx, avy) + a.obst_vmaxf) * 1.0001f;
        }
        const bool due_w = part && !((LINF ? travel_w : travel) < lim_w);
        const bool due_p = G > 1 && part && !(travel < lim_p);
        if (!__any_sync(FULL, due_w || due_p)) continue;  // every mover of this warp is certified clear

        // ---- wall check (basic:1888-1894) of the lanes that are due: lane-local
        bool bad = false, obad = false;
        if (due_w) {
            float clear_w;
            if (!BOX) {
                // float32 for the tracked cell, exact float64 (with the lazily generated noise) when too close to call
                bad = wall_bad_circle<NOISE>(a, tb, true, p.x, p.y, cw0, cw0f, gi, gj, ln.env_global, event,
                                             s0 + GPR_RNG_BLOCK_VEL_WALL, ln.m, 2, n4, have0, clear_w);
            } else {
                // float32 screen with a bounding rectangle; the exact vertex / rectangle tests of basic:559-572, 657-783
                // only where the screen cannot certify "valid" (0 is not a verdict: the rectangle is conservative)
                if (wall_fast(a, tb, p.x, p.y, bxf, byf, gi, gj, clear_w) != 1) {
                    bad = box_wall_exact<NOISE>(a, tb, p.x, p.y, cw0, cw1, ln.env_global, event, s0, (uint32_t)ln.m);
                    guess_cell(a, p.x, p.y, gi, gj);  // refresh the tracked cell
                }
            }
            // ---- static obstacles (the hook of basic:1903), checked with the walls on the wall check's noisy qpos
            if (a.n_obst > 0) {
                float clear_o;
                // (an extra body with a prescribed velocity has moved for as long as the movers have been integrated)
                const double t_obst = dmul((double)(elapsed * a.num_cycles + cyc + 1), a.dt);
                const int f = obstacle_screen<BOX>(a, p.x, p.y, BOX ? pxf : (float)cm0, pyf, clear_o, t_obst);
                if (f == 2) {
                    double wx = p.x, wy = p.y;
                    Rect rm;
                    if (NOISE) {
                        float w4[4];
                        normal4_cold(a.seed, ln.env_global, event, s0 + GPR_RNG_BLOCK_VEL_WALL, (uint32_t)ln.m, w4);
                        wx = noisy(p.x, w4[2], a.sigma_p);
                        wy = noisy(p.y, w4[3], a.sigma_p);
                        if (BOX) {
                            float q[4];
                            normal4_cold(a.seed, ln.env_global, event, s0 + GPR_RNG_BLOCK_WALL_QUAT, (uint32_t)ln.m, q);
                            rect_vertices(wx, wy, noisy(1.0, q[0], a.sigma_p), dmul((double)q[1], a.sigma_p), dmul((double)q[2], a.sigma_p),
                                          dmul((double)q[3], a.sigma_p), cm0, cm1, rm);
                        }
                    } else if (BOX) {
                        rect_vertices_axis(wx, wy, cm0, cm1, rm);
                    }
                    obad = obstacle_hit_exact<BOX>(a, wx, wy, cm0, rm, t_obst);
                } else {
                    obad = f == 1;
                }
                // (a circle's clearance is a Euclidean distance: |d|_2 <= sqrt(2) |d|_inf)
                clear_w = fminf(clear_w, LINF ? clear_o * 0.7071f : clear_o);
            }
            lim_w = ((LINF ? travel_w : travel) + clear_w * a.inv_dtf) * 0.999999f;
        }
        // ---- mover check (basic:1895-1901) on an independently noisy qpos: warp-collective
        bool hit = false;
        if (G > 1 && __any_sync(FULL, due_p)) {
            float clear_p;
            if (!BOX) {
                hit = pair_circle_fast<G, NOISE>(a, ln.lane, ln.m, part, p.x, p.y, cm0, 0, ln.env_global, event,
                                                 s0 + GPR_RNG_BLOCK_MOVER, clear_p);
            } else {
                // axis-gap screen; the exact rectangle test of geom:107-138 only for a warp with an uncleared pair
                bool near, sure;
                unsigned kmask;
                pair_box_screen<G>(ln.lane, ln.m, part, p.x, p.y, pxf, pyf, a.pair_mgf[NOISE ? 1 : 0], near, clear_p, kmask, sure,
                                   a.uniform_pairs ? (float)cm0 : 0.f, (float)cm1, 2.f * (pxf - (float)cm0) + a.pair_mgf[NOISE ? 1 : 0]);
                hit = sure;
                if (__any_sync(FULL, near))  // exact test of the pairs the screen could not decide (the others are certain)
                    hit = box_pair_exact<G, NOISE>(a, ln.lane, ln.m, part, p.x, p.y, cm0, cm1, ln.env_global, event, s0, kmask) || sure;
            }
            // each mover may use up half of the smallest margin among its own pairs
#if GPR_PAIR_ENV_MARGIN
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) clear_p = fminf(clear_p, __shfl_xor_sync(FULL, clear_p, o));
#endif
            lim_p = (travel + 0.5f * clear_p * a.inv_dtf) * 0.999999f;
        }
        const unsigned badm = __ballot_sync(FULL, bad), hitm = __ballot_sync(FULL, hit);
        const unsigned obm = a.n_obst > 0 ? __ballot_sync(FULL, obad) : 0u;
        if (badm | hitm | obm) {
            if (alive) {
                wc = (badm & ln.gmask) != 0u;
                mc = (hitm & ln.gmask) != 0u;
                oc = (obm & ln.gmask) != 0u;
                if (wc || mc || oc) {  // basic:1904 break
                    alive = false;
                    if (STASH && part) {
                        const int t = (int)threadIdx.x;
                        s_stash[0][t] = p.x;
                        s_stash[1][t] = p.y;
                        s_stash[2][t] = v.x;
                        s_stash[3][t] = v.y;
                        s_stash[4][t] = acc.x;
                        s_stash[5][t] = acc.y;
                        died = true;
                    }
                    part = false;
                }
            }
            any_alive = __any_sync(FULL, alive);
        }
    }

    if (STASH && died) {
        const int t = (int)threadIdx.x;
        p = make_double2(s_stash[0][t], s_stash[1][t]);
        v = make_double2(s_stash[2][t], s_stash[3][t]);
        acc = make_double2(s_stash[4][t], s_stash[5][t]);
    }

    // ------------------------------------------------------------------ observation, info, reward (basic:1910-1929)
    const bool stepped = ln.env_ok && !pending_reset;
    double2 ag, ov;
    int reached;
    observe<G, NOISE>(a, ln, event, p, v, goal, ag, ov, reached);
    float reward;
    bool term, succ;
    planning_reward(a.N, reached, mc || oc, wc, reward, term, succ);  // (an obstacle hit counts as a collision)
    if (stepped) {
        event += 1u;
        elapsed += 1;
    }
    const bool trunc = stepped && a.max_episode_steps > 0 && elapsed >= a.max_episode_steps;  // gymnasium TimeLimit
    const bool done = stepped && (term || trunc);

    // ------------------------------------------------------------------ auto-reset: hand finished envs to the reset kernel
    // The consumer may already be running (see planning_autoreset_kernel).  An entry carries all the consumer needs of the
    // env (index and RNG event) in one 64-bit store, the env's state is not written back (the reset rewrites all of it),
    // and the one store of this kernel that the consumer overwrites — the early desired_goal row — is fenced before the
    // entry.  Every warp reports exactly once, with the slots it reserves in the same atomic: when all warps of the grid
    // have reported the count is final.  SAME_STEP publishes here, ahead of this kernel's result stores; NEXT_STEP (the
    // consumer also rewrites the per-env flags, which other threads of the CTA store) after them, behind a barrier.
    bool need = false;
    if (a.autoreset == GPR_AUTORESET_SAME_STEP) need = done;
    if (a.autoreset == GPR_AUTORESET_NEXT_STEP) need = ln.env_ok && pending_reset;
    const bool handed = need && a.autoreset == GPR_AUTORESET_SAME_STEP;  // nothing of this env's state is stored below
    unsigned my_slot = 0u;  // this env's position on the work list (every lane of a handed-over env's group)
    auto publish = [&]() {
        const unsigned leaders = __ballot_sync(FULL, need && ln.m == 0);
        if (need && a.write_goal) __threadfence();  // (the early desired_goal store, if there was one)
        unsigned long long t = 0ull;
        if (ln.lane == 0) t = atomicAdd(a.reset_ctl + a.parity, (1ull << 32) | (unsigned long long)__popc(leaders));
        const unsigned slot0 = (unsigned)__shfl_sync(FULL, t, 0);
        const unsigned slot = slot0 + __popc(leaders & ((1u << ln.lane) - 1u));
        GPR_CHECK(a, !(need && ln.m == 0) || slot < (unsigned)a.B, DBG_LIST_SLOT);
        if (need && ln.m == 0) {
            if (EXTRA && a.compact_index) a.compact_index[slot] = ln.env;
            *reinterpret_cast<volatile unsigned long long*>(a.reset_list + slot) =
                ((unsigned long long)event << 32) | (unsigned long long)(uint32_t)ln.env;
        }
        if (EXTRA && a.compact_index) my_slot = __shfl_sync(FULL, slot, (int)(ln.lane & ~(unsigned)(G - 1)));  // from the group's lane m == 0
    };
    if (a.autoreset == GPR_AUTORESET_SAME_STEP) publish();

    // episode statistics (one lane per env accumulates, one atomic per warp and counter)
    {
        const bool lead = ln.env_ok && ln.m == 0;
        float ret = 0.f;
        if (lead && stepped) ret = a.ep_return[ln.env] + reward;
        const bool fin = lead && done;
        if (__any_sync(FULL, fin)) {
            // (planning rewards are whole numbers — -50, +50, -(movers off goal) — so every sum is an exact integer: one
            //  warp-reduce instruction per counter)
            const double s_ep = (double)__reduce_add_sync(FULL, fin ? 1 : 0);
            const double s_ret = (double)__reduce_add_sync(FULL, fin ? (int)ret : 0);
            const double s_len = (double)__reduce_add_sync(FULL, fin ? elapsed : 0);
            const double s_succ = (double)__reduce_add_sync(FULL, (fin && succ) ? 1 : 0);
            const double s_mc = (double)__reduce_add_sync(FULL, (fin && mc) ? 1 : 0);
            const double s_wc = (double)__reduce_add_sync(FULL, (fin && wc) ? 1 : 0);
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

    // per-env scalars of the finished transition: gathered per CTA in shared memory and written as contiguous 8-byte units
    // (one byte per env and flag would otherwise be a 1-byte store per env — harmless in HBM, but when the outputs live in
    // pinned host memory every such fragment is its own PCIe write: measured 47 us per step for 0.3 MB).
    // NO CTA BARRIER on the common path: every warp drops its slice into shared memory and signs off on a counter; the warp
    // that signs off LAST writes the CTA's rows.  A warp whose environments have all collided therefore leaves the SM
    // instead of waiting for the slowest warp of its CTA (the barrier was 1.4 / 2.8 stall cycles per issued instruction of
    // the circle / box kernel, profiles/r1b_full_planning4.txt, r2_full_planning8box.txt).
    // Envs that did not step (NEXT_STEP mode, pending reset) get zeros here and their real values from the auto-reset kernel
    // — which may already be running: in that mode the rows are written behind a real barrier and published afterwards.
    {
        constexpr int EPC = StepThreads<G>::value / G;  // envs per CTA
        constexpr int NW = StepThreads<G>::value / 32;  // warps per CTA
        constexpr int U = EPC >= 8 ? 8 : EPC;           // bytes per flag store unit
        constexpr int FU = EPC / U;                     // units per flag array
        __shared__ __align__(16) uint8_t s_flag[6][EPC];
        __shared__ __align__(16) float s_rew[EPC];
        const int le = (int)threadIdx.x / G;
        const int env0 = (int)blockIdx.x * EPC;
        GPR_CHECK(a, le >= 0 && le < EPC, DBG_SHARED_INDEX);
        if (ln.m == 0) {
            s_rew[le] = stepped ? reward : 0.f;
            s_flag[0][le] = stepped && term;
            s_flag[1][le] = trunc;
            s_flag[2][le] = stepped && succ;
            s_flag[3][le] = stepped && mc;
            s_flag[4][le] = stepped && wc;
            s_flag[5][le] = stepped && oc;
        }
        bool writer;
        if (a.autoreset == GPR_AUTORESET_NEXT_STEP) {
            __syncthreads();
            writer = threadIdx.x < 32u;
        } else {
            __syncwarp();
            unsigned prev = 0u;
            if (ln.lane == 0) {
                __threadfence_block();  // this warp's slice before its signature
                prev = atomicAdd(&s_signed_off, 1u);
            }
            writer = __shfl_sync(FULL, prev, 0) == (unsigned)(NW - 1);
            if (writer) __threadfence_block();  // ... and every other warp's slice before the reads below
        }
        if (writer) {
            uint8_t* const fo[6] = {a.out.terminated, a.out.truncated, a.out.is_success, a.out.mover_collision, a.out.wall_collision,
                                    a.out.other_collision};
            const bool whole = env0 + EPC <= a.B;  // (the last CTA may be partial)
            for (int t = (int)ln.lane; t < 6 * FU; t += 32) {
                const int f = t / FU, k = t % FU;
                uint8_t* dst = fo[f];
                if (!dst) continue;
                GPR_CHECK(a, !whole || env0 + U * k + U <= a.B, DBG_OUTPUT_ROW);
                dst += env0 + U * k;
                const volatile uint8_t* src = &s_flag[f][U * k];
                if (whole && (reinterpret_cast<uintptr_t>(dst) & (uintptr_t)(U - 1)) == 0u) {
                    if constexpr (U == 8) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(const_cast<const uint8_t*>(src));
                    else if constexpr (U == 4) *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(const_cast<const uint8_t*>(src));
                    else if constexpr (U == 2) *reinterpret_cast<uint16_t*>(dst) = *reinterpret_cast<const uint16_t*>(const_cast<const uint8_t*>(src));
                    else *dst = *src;
                } else {
                    for (int i = 0; i < U; ++i)
                        if (env0 + U * k + i < a.B) dst[i] = src[i];
                }
            }
            if (a.out.reward) {
                float* dst = a.out.reward + env0;
                if constexpr (EPC >= 2) {
                    for (int t = (int)ln.lane; t < EPC / 2; t += 32) {
                        if (whole && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0u) {
                            reinterpret_cast<float2*>(dst)[t] = reinterpret_cast<const float2*>(s_rew)[t];
                        } else {
                            for (int i = 2 * t; i < 2 * t + 2; ++i)
                                if (env0 + i < a.B) dst[i] = s_rew[i];
                        }
                    }
                } else {
                    if (ln.lane == 0 && env0 < a.B) dst[0] = s_rew[0];
                }
            }
        }
    }

    if (a.autoreset == GPR_AUTORESET_NEXT_STEP) {
        __threadfence();
        __syncthreads();
        publish();
    }

    // ------------------------------------------------------------------ final observation / observation rows
    if (handed) {
        if (EXTRA && a.compact_index)
            store_obs_row<G>(a, ln, (size_t)my_slot, a.compact_final_obs, a.compact_final_ag, a.compact_final_dg, ov, acc, ag, goal);
        else
            store_obs<G, EXTRA>(a, ln, a.out.final_observation, a.out.final_achieved_goal, a.out.final_desired_goal, ov, acc, ag, goal);
    }
    // (the rows of envs handed to planning_autoreset_kernel are written there: first observation of the new episode)
    if (stepped && !handed) store_obs<G, EXTRA>(a, ln, a.out.observation, a.out.achieved_goal, nullptr, ov, acc, ag, goal);

    // ------------------------------------------------------------------ state write-back
    if (ln.active && stepped && !handed) {
        a.pos[ln.idx] = p;
        a.vel[ln.idx] = v;
        a.acc[ln.idx] = acc;
    }
    if (ln.env_ok && ln.m == 0 && stepped && !handed) {
        a.rng[ln.env] = event;
        a.elapsed[ln.env] = elapsed;
        if (a.autoreset == GPR_AUTORESET_NEXT_STEP) a.needs_reset[ln.env] = done ? 1 : 0;
    }
}

// plan:355-418 + basic:1770-1833 for the envs on the reset list.  A warp claims a BATCH of up to 32/G list entries through an
// atomic cursor (attempt counts are geometric, so work is balanced dynamically; the batch shrinks with the work that is
// left).  The rejection sampling of the batch's envs runs one env after the other with the whole warp working on that
// env (sample_env); lane group g keeps the result of the batch's g-th env.  Everything after the sampling — reset-time
// checks, first observation, state and output stores — then runs ONCE for the whole batch with every lane group acting
// as the movers of its own env (instead of once per env with only G of the 32 lanes active).
//
// STREAMING.  gpr_step launches this kernel as the programmatic dependent of planning_step_kernel: its CTAs become
// resident as soon as the last CTA of the step grid has started and consume list entries while the step kernel's last
// wave is still running (a 65,536-env step is 3.5 waves: the partial wave leaves half of the SM slots idle for a quarter
// of the kernel).  There is no grid-level dependency: an entry is valid once its slot holds an env index (published by
// the step kernel after a fence), and a warp leaves when every warp of the step grid has reported (reset_ctl) and the cursor has
// reached the final count.  Launched the ordinary way (per-kernel timing, profilers) the same code simply finds
// everything published.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int G, bool BOX, bool NOISE, bool EXTRA>
__global__ void __launch_bounds__(128, BOX ? GPR_AR_MINB_BOX : GPR_AR_MINB) planning_autoreset_kernel(const __grid_constant__ PlanArgs a) {
    __shared__ Tables tb;
    load_tables(tb, a.L);
    __syncthreads();
    constexpr unsigned S = 32u / G;  // envs per batch
    const unsigned lane = threadIdx.x & 31u;
    const unsigned grp = lane / G;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // the other buffer belongs to the next step: clear it now
        a.reset_ctl[a.parity ^ 1] = 0ull;
        a.reset_cursor[a.parity ^ 1] = 0u;
    }
    const unsigned warps = gridDim.x * (blockDim.x / 32u);
    const unsigned step_warps = a.step_ctas * (unsigned)(StepThreads<G>::value / 32);
    uint32_t* const cursor = a.reset_cursor + a.parity;
    for (;;) {
        // ---- claim [i0, i0 + nb) once the cursor is behind the slots reserved so far (a plain atomicAdd: a claim that
        //      races past the count simply waits until producers have reserved those slots too, or learns that none will);
        //      nb = 0: the step kernel is done and nothing is left
        uint32_t i0 = 0, nb = 0;
        if (lane == 0) {
            unsigned spins = 0;
            bool claimed = false;
            for (;;) {
                const unsigned long long ctl = ld_acquire_u64(a.reset_ctl + a.parity);
                const bool done = (unsigned)(ctl >> 32) >= step_warps;  // every warp has reported: the count is final
                const uint32_t cnt = (uint32_t)ctl;
                if (!claimed) {
                    const uint32_t cur = *reinterpret_cast<volatile uint32_t*>(cursor);
                    if (cur < cnt) {
                        // as many envs as there are lane groups, but never so many that other warps would stay without work
                        nb = min(S, max(1u, (cnt - cur + warps - 1u) / warps));
                        i0 = atomicAdd(cursor, nb);
                        claimed = true;
                    }
                }
                if (claimed) {
                    if (cnt >= i0 + nb) break;
                    if (done) {
                        nb = cnt > i0 ? cnt - i0 : 0u;
                        break;
                    }
                } else if (done) {
                    break;
                }
                __nanosleep(200);
                if (++spins > (1u << 25)) {  // (~10 s: a step grid that never reports in must not hang the device)
                    atomicAdd(a.fail_count, 1u << 20);
                    nb = 0;
                    break;
                }
            }
        }
        i0 = __shfl_sync(FULL, i0, 0);
        nb = __shfl_sync(FULL, nb, 0);
        if (nb == 0) break;
        GPR_CHECK(a, lane != 0 || ((unsigned long long)i0 + nb <= (unsigned long long)(uint32_t)ld_acquire_u64(a.reset_ctl + a.parity) && i0 + nb <= (uint32_t)a.B), DBG_LIST_CLAIM);
        Lane<G> ln;
        ln.lane = lane;
        ln.gmask = group_mask<G>(lane);
        ln.m = (int)(lane % G);
        ln.env_ok = grp < nb;
        unsigned long long entry = 0ull;
        if (ln.env_ok) {  // the slot is reserved; its entry follows within the producer's next few instructions
            volatile unsigned long long* slot = a.reset_list + i0 + grp;
            do {
                entry = *slot;
            } while (entry == ~0ull);
        }
        __syncwarp();
        if (ln.env_ok && ln.m == 0) a.reset_list[i0 + grp] = ~0ull;  // leave the list empty for the next step
        __threadfence();  // (the producer's early desired_goal store is ordered before the row written below)
        ln.env = (int)(uint32_t)entry;
        GPR_CHECK(a, !ln.env_ok || (uint32_t)entry < (uint32_t)a.B, DBG_LIST_ENTRY);
#ifdef GPR_DEBUG_BOUNDS
        if (ln.env_ok && (uint32_t)entry >= (uint32_t)a.B) ln.env_ok = false;  // (counted above; do not touch memory with it)
#endif
        ln.active = ln.env_ok && ln.m < a.N;
        ln.idx = (size_t)ln.env * (size_t)a.N + (size_t)ln.m;
        ln.env_global = a.env_base + (uint32_t)ln.env;
        const uint32_t event = (uint32_t)(entry >> 32);
        double2 p = make_double2(0, 0), v = p, acc = p, goal = p;
        bool failed = false;
#pragma unroll 1
        for (unsigned b = 0; b < nb; ++b) {  // the whole warp samples for the batch's b-th env
            const uint32_t eg = __shfl_sync(FULL, ln.env_global, (int)(b * G));
            const uint32_t ev = __shfl_sync(FULL, event, (int)(b * G));
            double2 pb = p, gb = p;
            bool f1 = false, f2 = false;
            if constexpr (!BOX && G <= GPR_LANES_MAX_G && GPR_SAMPLER_ONE_COPY) {
#pragma unroll 1
                for (int kind = 0; kind < 2; ++kind) {  // starts, then goals, through one copy of the sampling loop
                    double2 r = p;
                    bool f = false;
                    sample_env_lanes<G, BOX, -1>(a, tb, lane, eg, ev, r, f, kind);
                    if (kind == 0) {
                        pb = r;
                        f1 = f;
                    } else {
                        gb = r;
                        f2 = f;
                    }
                }
            } else {
                sample_env<G, BOX, 0>(a, tb, lane, ln.gmask, eg, ev, pb, f1);  // every lane: position of mover lane % G
                sample_env<G, BOX, 1>(a, tb, lane, ln.gmask, eg, ev, gb, f2);
            }
            if (grp == b) {
                p = pb;
                goal = gb;
                failed = f1 || f2;
            }
        }
        // ---- the rest for all envs of the batch at once: lane group g = the movers of env g
        bool mc = false, wc = false, oc = false;
        reset_checks<G, BOX, NOISE>(a, tb, ln, ln.env_ok, event, p, mc, wc, oc);
        double2 ag, ov;
        int reached;
        observe<G, NOISE>(a, ln, event, p, v, goal, ag, ov, reached);
        if (EXTRA && a.compact_index && a.autoreset == GPR_AUTORESET_SAME_STEP) {
            // compact transport: the new episode's goal goes to this entry's row of the list, not to row `env`
            store_obs<G, EXTRA>(a, ln, a.out.observation, a.out.achieved_goal, nullptr, ov, acc, ag, goal);
            if (ln.active) store_pair(a.out_f64 != 0, a.compact_goal, (size_t)(i0 + grp) * (size_t)a.N + ln.m, goal.x, goal.y);
        } else {
            store_obs<G, EXTRA>(a, ln, a.out.observation, a.out.achieved_goal, a.out.desired_goal, ov, acc, ag, goal);
        }
        if (ln.active) {
            a.pos[ln.idx] = p;
            a.vel[ln.idx] = v;
            a.acc[ln.idx] = acc;
            a.goal[ln.idx] = goal;
        }
        if (ln.env_ok && ln.m == 0) {
            a.rng[ln.env] = event + 1u;
            a.elapsed[ln.env] = 0;
            if (failed) atomicAdd(a.fail_count, 1u);
            if (a.autoreset == GPR_AUTORESET_NEXT_STEP) {
                // gymnasium NEXT_STEP: this call only resets; reward 0, not done; info of the fresh episode
                float r2;
                bool t2, s2;
                planning_reward(a.N, reached, mc || oc, wc, r2, t2, s2);
                a.needs_reset[ln.env] = 0;
                if (a.out.reward) a.out.reward[ln.env] = 0.f;
                if (a.out.terminated) a.out.terminated[ln.env] = 0;
                if (a.out.truncated) a.out.truncated[ln.env] = 0;
                if (a.out.is_success) a.out.is_success[ln.env] = s2;
                if (a.out.mover_collision) a.out.mover_collision[ln.env] = mc;
                if (a.out.wall_collision) a.out.wall_collision[ln.env] = wc;
                if (a.out.other_collision) a.out.other_collision[ln.env] = oc;
            }
        }
    }
    // Launched as the step kernel's programmatic dependent, this grid may run out of work while the step grid's last
    // CTAs are still writing their results (every step warp reports BEFORE its trailing stores).  griddepcontrol.wait
    // returns once the prerequisite grid has completed and its memory operations are visible: whatever follows this
    // kernel in the stream — the next step, a copy, cudaStreamSynchronize — is then ordered after BOTH grids, as PTX
    // requires of a dependent of a grid that executed launch_dependents.  (An ordinary launch: returns at once.)
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <int G, bool BOX, bool NOISE>
__global__ void __launch_bounds__(256) planning_reset_kernel(const __grid_constant__ PlanArgs a) {
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
    bool mc = false, wc = false, oc = false, failed = false;
    reset_group<G, BOX, NOISE>(a, tb, ln, need, event, a.inject_start, a.inject_goal, p, v, acc, goal, mc, wc, oc, failed);
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
        planning_reward(a.N, reached, mc || oc, wc, r, t, s);
        if (a.out.is_success) a.out.is_success[ln.env] = s;
        if (a.out.mover_collision) a.out.mover_collision[ln.env] = mc;
        if (a.out.wall_collision) a.out.wall_collision[ln.env] = wc;
        if (a.out.other_collision) a.out.other_collision[ln.env] = oc;
        a.rng[ln.env] = event + 1u;
        a.elapsed[ln.env] = 0;
        a.ep_return[ln.env] = 0.f;
        if (a.needs_reset) a.needs_reset[ln.env] = 0;
        if (failed) atomicAdd(a.fail_count, 1u);
    }
}

}  // namespace gpr
