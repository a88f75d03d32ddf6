/*
 * gpr_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C float64 restatement of the reference's step path (ubi-coro/gymnasium-planar-robotics v1.1.0a2), one
 * environment at a time, in the reference's own structure and operation order.  It exists to CHECK the CUDA path; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (gymnasium-planar-robotics_b200/) never links, imports or falls back to anything in this directory.
 *
 * Pinning (tests/test_oracle_vs_reference.py live against the mounted reference, tests/test_oracle_golden.py on the committed vectors of tests/golden/):
 *   - gpro_qpos_is_valid, gpro_check_mover_collision, gpro_segments_intersect, gpro_rectangles_intersect,
 *     gpro_ensure_max_dyn_val, gpro_planning_reward are compared against the UNMODIFIED reference functions (imported
 *     from /root/reference with mujoco/gymnasium stubbed, tests/ref_harness.py) on the reference's own 100+34 test
 *     vectors and on random inputs; the resulting vectors are committed under tests/golden/.
 *   - the integrator recurrence is the closed form the reference's tests assert against MuJoCo
 *     (tests/test_benchmark_planning_env.py:86-93, 199-204).  MuJoCo itself is not installable here, so
 *     "MuJoCo == this recurrence" is taken from those tests, not re-verified: PARITY WITH mj_step IS PINNED ONLY THROUGH
 *     THE REFERENCE'S OWN CLOSED FORM.
 *   - pushing contact dynamics: PARITY UNPINNED (no reference test exercises contact).
 *
 * File:line citations are into /root/reference/gymnasium_planar_robotics/ :
 *   basic = envs/basic_envs.py, plan = envs/planning/benchmark_planning_env.py,
 *   push = envs/manipulation/benchmark_pushing_env.py, geom = utils/geometry_2D_utils.py, rot = utils/rotations_utils.py
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off: every product and sum is rounded separately, like NumPy).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gpr.h"
#include "../include/gpr_rng.h"
#include "../include/gpr_push_physics.h"

/* The pushing solve is written with explicit float32 fused multiply-adds (fmaf).  The library is built for the x86-64-v2
 * baseline so that it runs on any host; the two pushing entry points are additionally cloned for FMA3 hardware (resolved at
 * load time by the dynamic linker): fmaf() is then one instruction instead of a call into glibc.  Both clones round
 * identically (fmaf is correctly rounded either way; -ffp-contract=off forbids any other contraction). */
#if defined(__GNUC__) && defined(__x86_64__) && !defined(GPRO_NO_CLONES)
#define GPRO_FMA_CLONES __attribute__((target_clones("fma", "default")))
#else
#define GPRO_FMA_CLONES
#endif

#ifdef _OPENMP
#include <omp.h>
#endif

#define NMAX GPR_MAX_MOVERS

/* number of times a reference `assert` would have fired (basic:514-517 not-above-a-tile, basic:650 mask in {0,1}) */
static int64_t g_assert_trips = 0;
int64_t gpro_assert_trips(int reset) {
    int64_t v = g_assert_trips;
    if (reset) g_assert_trips = 0;
    return v;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* plan:610-645  ensure_max_dyn_val                                                                                     */
/* ------------------------------------------------------------------------------------------------------------------ */
void gpro_ensure_max_dyn_val(const double cur[2], double max_value, const double deriv[2], double dt, double next_val[2],
                             double next_deriv[2]) {
    /* plan:630  next_values_tmp = cycle_time * next_derivs + current_values */
    double tx = dt * deriv[0] + cur[0];
    double ty = dt * deriv[1] + cur[1];
    /* plan:632  np.linalg.norm(.., ord=2, axis=1) == sqrt(add.reduce(x*x)) */
    double nrm = sqrt(tx * tx + ty * ty);
    if (nrm >= max_value) { /* plan:633 mask_norm (>=) */
        /* plan:639-642 */
        next_val[0] = max_value * (tx / nrm);
        next_val[1] = max_value * (ty / nrm);
        next_deriv[0] = (next_val[0] - cur[0]) / dt;
        next_deriv[1] = (next_val[1] - cur[1]) / dt;
    } else {
        next_val[0] = tx;
        next_val[1] = ty;
        next_deriv[0] = deriv[0];
        next_deriv[1] = deriv[1];
    }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* geom:9-69  check_line_segments_intersect (one pair)                                                                  */
/* ------------------------------------------------------------------------------------------------------------------ */
static double orient3(const double a[2], const double b[2], const double c[2]) {
    /* geom:47-60: det([[ax,bx,cx],[ay,by,cy],[1,1,1]]) — the reference calls LAPACK (np.linalg.det); the same
       determinant is evaluated here by cofactor expansion along the last row, which is what an exact LU reduces to. */
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
}

static int pts_equal(const double a[2], const double b[2]) {
    return (fabs(a[0] - b[0]) < 1e-7) && (fabs(a[1] - b[1]) < 1e-7); /* geom:30-35 */
}

int gpro_segments_intersect(const double p1[2], const double p2[2], const double q1[2], const double q2[2]) {
    int points_equal = pts_equal(p1, q1) || pts_equal(p1, q2) || pts_equal(p2, q1) || pts_equal(p2, q2);
    int minmax = 0; /* geom:37-45 */
    for (int k = 0; k < 2; ++k) {
        double min_p = fmin(p1[k], p2[k]), max_p = fmax(p1[k], p2[k]);
        double min_q = fmin(q1[k], q2[k]), max_q = fmax(q1[k], q2[k]);
        int mask_pq = max_p < min_q;
        int mask_qp = max_q < min_p;
        int v = mask_pq * (1 - (fabs(max_p - min_q) < 1e-7)) + mask_qp * (1 - (fabs(max_q - min_p) < 1e-7));
        minmax += v;
    }
    minmax = minmax >= 1;
    double d1 = orient3(p1, p2, q1), d2 = orient3(p1, p2, q2);
    double d3 = orient3(q1, q2, p1), d4 = orient3(q1, q2, p2);
    double pa = d1 * d2, pb = d3 * d4;
    /* geom:62-64:  (sign(pa) <= 0  or |pa| < 1e-7) and (sign(pb) <= 0 or |pb| < 1e-7) */
    int orientation = ((pa <= 0.0) || (fabs(pa) < 1e-7)) && ((pb <= 0.0) || (fabs(pb) < 1e-7));
    int res = orientation; /* geom:66-68, in this order */
    if (minmax) res = 0;
    if (points_equal) res = 1;
    return res;
}

/* geom:72-104 get_2D_rect_vertices (one rectangle); rot:414-461 unit_vector (float32!), rot:248-274 quat2mat          */
void gpro_rect_vertices(const double qpos[7], const double size[2], double vx[4], double vy[4]) {
    /* rot:447  data = np.array(data, dtype=np.float32); length = sqrt(sum(data*data)); data /= length  (all float32)   */
    float qf[4];
    for (int k = 0; k < 4; ++k) qf[k] = (float)qpos[3 + k];
    /* np.sum over the last axis of a (n,4) float32 array: out = d0; out += (0 + d1 + d2 + d3) is NOT what numpy does for
       a contiguous inner reduce of 4 elements: it accumulates left to right, ((d0+d1)+d2)+d3 (verified against the
       reference in tests/test_oracle_golden.py::test_rect_vertices). */
    float s = qf[0] * qf[0];
    s = s + qf[1] * qf[1];
    s = s + qf[2] * qf[2];
    s = s + qf[3] * qf[3];
    float len = sqrtf(s);
    for (int k = 0; k < 4; ++k) qf[k] = qf[k] / len;
    /* rot:253-273 (float64 from here) */
    double w = (double)qf[0], x = (double)qf[1], y = (double)qf[2], z = (double)qf[3];
    double Nq = ((w * w + x * x) + y * y) + z * z;
    double r00, r01, r10, r11;
    if (Nq > 2.220446049250313e-16) {
        double sc = 2.0 / Nq;
        double X = x * sc, Y = y * sc, Z = z * sc;
        double wZ = w * Z;
        double xX = x * X, xY = x * Y;
        double yY = y * Y, zZ = z * Z;
        r00 = 1.0 - (yY + zZ);
        r01 = xY - wZ;
        r10 = xY + wZ;
        r11 = 1.0 - (xX + zZ);
    } else {
        r00 = 1.0;
        r01 = 0.0;
        r10 = 0.0;
        r11 = 1.0;
    }
    /* geom:91-102: local vertices (-sx,-sy), (-sx,sy), (sx,sy), (sx,-sy), z = 0; base = R @ v + pos */
    const double lx[4] = {-size[0], -size[0], size[0], size[0]};
    const double ly[4] = {-size[1], size[1], size[1], -size[1]};
    for (int k = 0; k < 4; ++k) {
        vx[k] = (r00 * lx[k] + r01 * ly[k]) + qpos[0];
        vy[k] = (r10 * lx[k] + r11 * ly[k]) + qpos[1];
    }
}

/* geom:107-138 check_rectangles_intersect (one pair) */
int gpro_rectangles_intersect(const double qpos1[7], const double qpos2[7], const double size1[2],
                              const double size2[2]) {
    double ax[4], ay[4], bx[4], by[4];
    gpro_rect_vertices(qpos1, size1, ax, ay);
    gpro_rect_vertices(qpos2, size2, bx, by);
    int any = 0;
    for (int i = 0; i < 4; ++i) {
        double p1[2] = {ax[i], ay[i]}, p2[2] = {ax[(i + 1) & 3], ay[(i + 1) & 3]}; /* geom:133 np.roll(-1) */
        for (int j = 0; j < 4; ++j) {
            double q1[2] = {bx[j], by[j]}, q2[2] = {bx[(j + 1) & 3], by[(j + 1) & 3]};
            any |= gpro_segments_intersect(p1, p2, q1, q2);
        }
    }
    return any;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* basic:459-788  qpos_is_valid                                                                                         */
/* ------------------------------------------------------------------------------------------------------------------ */
static int lay(const gpr_config* c, int i, int j) {
    /* layout_tiles_wc: padded with one zero row/column (basic:221); python index -1 wraps onto that zero padding */
    if (i < 0 || j < 0 || i >= c->num_tiles_x || j >= c->num_tiles_y) return 0;
    return c->layout[i * c->num_tiles_y + j] != 0;
}

/* basic:205-207, 1313-1339: is (i,j) the centre of a fully populated 3x3 block? */
static int is_3x3_centre(const gpr_config* c, int i, int j) {
    if (i < 1 || j < 1 || i > c->num_tiles_x - 2 || j > c->num_tiles_y - 2) return 0;
    for (int a = -1; a <= 1; ++a)
        for (int b = -1; b <= 1; ++b)
            if (!lay(c, i + a, j + b)) return 0;
    return 1;
}

/* one (qpos, containing cell) row of basic:542-655; px/py = the points tested (1 for circle with +-csz, 4 box vertices) */
static int row_mask_valid(const gpr_config* c, int i, int j, int is_circle, const double* px, const double* py, int np_,
                          double csz) {
    const double hx = c->tile_half[0], hy = c->tile_half[1];
    const double min_x = c->tile_cx[i] - hx, max_x = c->tile_cx[i] + hx; /* basic:519-523 */
    const double min_y = c->tile_cy[j] - hy, max_y = c->tile_cy[j] + hy;
    const int T = lay(c, i, j);
    const int x_lmin = i > 0, y_lmin = j > 0, x_smax = i < c->num_tiles_x - 1, y_smax = j < c->num_tiles_y - 1;
    int all_ok = 1;
    for (int k = 0; k < np_; ++k) {
        int min_x_safe, max_x_safe, min_y_safe, max_y_safe;
        if (is_circle) { /* basic:545-558 */
            min_x_safe = T * (min_x < px[k] - csz);
            max_x_safe = T * (px[k] + csz < max_x);
            min_y_safe = T * (min_y < py[k] - csz);
            max_y_safe = T * (py[k] + csz < max_y);
        } else { /* basic:559-572 */
            min_x_safe = T * (min_x < px[k]);
            max_x_safe = T * (px[k] < max_x);
            min_y_safe = T * (min_y < py[k]);
            max_y_safe = T * (py[k] < max_y);
        }
        int mv = min_x_safe * max_x_safe * min_y_safe * max_y_safe; /* basic:580 */
        /* basic:582-612 */
        int min_x_upd = (1 - min_x_safe) * (x_lmin * lay(c, i, j) * lay(c, i - 1, j));
        mv += min_x_upd * min_y_safe * max_y_safe;
        int mnx_mny = (1 - min_y_safe) * (x_lmin * y_lmin * lay(c, i, j) * lay(c, i, j - 1) * lay(c, i - 1, j - 1));
        mv += min_x_upd * mnx_mny;
        int mnx_mxy = (1 - max_y_safe) * (x_lmin * y_smax * lay(c, i, j) * lay(c, i, j + 1) * lay(c, i - 1, j + 1));
        mv += min_x_upd * mnx_mxy;
        /* basic:614-643 */
        int max_x_upd = (1 - max_x_safe) * (x_smax * lay(c, i, j) * lay(c, i + 1, j));
        mv += max_x_upd * min_y_safe * max_y_safe;
        int mxx_mny = (1 - min_y_safe) * (x_smax * y_lmin * lay(c, i, j) * lay(c, i, j - 1) * lay(c, i + 1, j - 1));
        mv += max_x_upd * mxx_mny;
        int mxx_mxy = (1 - max_y_safe) * (x_smax * y_smax * lay(c, i, j) * lay(c, i, j + 1) * lay(c, i + 1, j + 1));
        mv += max_x_upd * mxx_mxy;
        /* basic:645-657 */
        int min_y_upd = (1 - min_y_safe) * (y_lmin * lay(c, i, j) * lay(c, i, j - 1));
        mv += min_y_upd * min_x_safe * max_x_safe;
        int max_y_upd = (1 - max_y_safe) * (y_smax * lay(c, i, j) * lay(c, i, j + 1));
        mv += max_y_upd * min_x_safe * max_x_safe;
        if (mv != 0 && mv != 1) { /* basic:659 assert */
#pragma omp atomic
            g_assert_trips++;
            mv = 0;
        }
        all_ok &= mv; /* circle: the single value (basic:662); box: sum over 4 vertices == 4 (basic:664) */
    }
    return all_ok;
}

/* csize: n x 2 array that ALREADY includes offset_wall and the optional safety offset (basic:487 is done by the caller,
   i.e. by the host mirror, in Python float64) */
void gpro_qpos_is_valid(const gpr_config* c, int n, const double* qpos, const double* csize, int32_t* valid) {
    const int is_circle = c->c_shape == GPR_SHAPE_CIRCLE;
    const double hx = c->tile_half[0], hy = c->tile_half[1];
    for (int q = 0; q < n; ++q) {
        const double* qp = qpos + 7 * q;
        const double* cs = csize + 2 * q;
        double vx[4], vy[4];
        int np_ = 1;
        if (is_circle) {
            vx[0] = qp[0];
            vy[0] = qp[1];
        } else {
            gpro_rect_vertices(qp, cs, vx, vy); /* basic:494-496 */
            np_ = 4;
        }
        int rows = 0, complete = 0, all_rows = 1;
        for (int i = 0; i < c->num_tiles_x; ++i) {
            for (int j = 0; j < c->num_tiles_y; ++j) {
                /* basic:507-512 mask_above_tile (inclusive on both sides; independent of tile presence) */
                if (!((c->tile_cx[i] - hx <= qp[0]) && (qp[0] <= c->tile_cx[i] + hx) && (c->tile_cy[j] - hy <= qp[1]) &&
                      (qp[1] <= c->tile_cy[j] + hy)))
                    continue;
                rows++;
                if (is_3x3_centre(c, i, j)) complete = 1; /* basic:527-538: valid at once, size and yaw ignored */
                int rv = row_mask_valid(c, i, j, is_circle, vx, vy, np_, cs[0]);
                if (!is_circle && rv) {
                    /* basic:666-783: the four "2x2 block with one missing corner" patterns.  The mover's cell is
                       diagonal to the missing tile and both side tiles exist; reject if the mover rectangle
                       edge-intersects the missing tile's rectangle. */
                    static const int dd[4][2] = {{+1, -1}, {+1, +1}, {-1, -1}, {-1, +1}}; /* bl, br, tl, tr */
                    for (int p = 0; p < 4 && rv; ++p) {
                        int mi = i + dd[p][0], mj = j + dd[p][1];
                        if (mi < 0 || mj < 0 || mi >= c->num_tiles_x || mj >= c->num_tiles_y) continue;
                        if (!(lay(c, i, j) && lay(c, mi, j) && lay(c, i, mj) && !lay(c, mi, mj))) continue;
                        double tq[7] = {c->tile_cx[mi], c->tile_cy[mj], 0, 1, 0, 0, 0};
                        double ts[2] = {hx, hy};
                        if (gpro_rectangles_intersect(qp, tq, cs, ts)) rv = 0;
                    }
                }
                all_rows &= rv;
            }
        }
        if (rows == 0) { /* basic:514-517 assert: "At least one mover is not above a tile" */
#pragma omp atomic
            g_assert_trips++;
        }
        valid[q] = complete || (rows > 0 && all_rows); /* basic:538, 785-786 */
    }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* basic:355-424  check_mover_collision                                                                                 */
/* csize: n x 2, already including the optional safety offset (basic:390)                                               */
/* ------------------------------------------------------------------------------------------------------------------ */
int gpro_check_mover_collision(const gpr_config* c, int n, const double* qpos, const double* csize) {
    if (n < 2) return 0;
    int collision = 0;
    if (c->c_shape == GPR_SHAPE_CIRCLE) {
        if (c->reference_quirks) {
            /* basic:409: (P,) <= (P,1) broadcasts to (P,P): any pair distance <= any pair's radius sum */
            double dmin = INFINITY, rmax = -INFINITY;
            for (int i = 0; i < n - 1; ++i)
                for (int j = i + 1; j < n; ++j) {
                    double dx = qpos[7 * i] - qpos[7 * j], dy = qpos[7 * i + 1] - qpos[7 * j + 1];
                    double d = sqrt(dx * dx + dy * dy);
                    double r = csize[2 * i] + csize[2 * j];
                    if (d < dmin) dmin = d;
                    if (r > rmax) rmax = r;
                }
            return dmin <= rmax;
        }
        for (int i = 0; i < n - 1; ++i)
            for (int j = i + 1; j < n; ++j) {
                double dx = qpos[7 * i] - qpos[7 * j], dy = qpos[7 * i + 1] - qpos[7 * j + 1];
                double d = sqrt(dx * dx + dy * dy);
                collision |= d <= (csize[2 * i] + csize[2 * j]);
            }
        return collision;
    }
    for (int i = 0; i < n - 1; ++i)
        for (int j = i + 1; j < n; ++j) {
            double dx = qpos[7 * i] - qpos[7 * j], dy = qpos[7 * i + 1] - qpos[7 * j + 1];
            double d = sqrt(dx * dx + dy * dy);
            /* basic:411-414 */
            double m = fmax(fmax(csize[2 * i], csize[2 * i + 1]), fmax(csize[2 * j], csize[2 * j + 1]));
            double thr = 2.0 * (fabs(m) + fabs(m));
            if (d <= thr) collision |= gpro_rectangles_intersect(qpos + 7 * i, qpos + 7 * j, csize + 2 * i, csize + 2 * j);
        }
    return collision;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* plan:502-534 compute_reward, plan:459-479 compute_terminated, plan:575-602 _get_info (one env)                       */
/* Static obstacles — the typed form of the reference's extension point _check_for_other_collisions_callback
 * (basic:1976-1986, called at basic:1807 and basic:1903; see gpr_config.num_obstacles in include/gpr.h).  Rules of the
 * mover-mover check (basic:390-424) between every mover and every obstacle: circle  ||p - o|| <= r_mover + r_obstacle
 * (inclusive like basic:409);  box  geom.check_rectangles_intersect (geom:107-138) or the mover's centre inside the
 * obstacle (the edge test alone does not see containment).  qpos: [n][7] (noisy) mover poses; csize: [n][2] sizes. */
int gpro_check_obstacle_collision_at(const gpr_config* c, int n, const double* qpos, const double* csize, double t) {
    for (int k = 0; k < c->num_obstacles; ++k) {
        /* typed extra bodies: prescribed constant velocity, position = p0 + v * t (gpr_config.obstacle_vel) */
        const double ox = c->obstacle_xy[k][0] + c->obstacle_vel[k][0] * t, oy = c->obstacle_xy[k][1] + c->obstacle_vel[k][1] * t;
        for (int m = 0; m < n; ++m) {
            const double* q = qpos + 7 * m;
            if (c->c_shape == GPR_SHAPE_CIRCLE) {
                const double dx = q[0] - ox, dy = q[1] - oy;
                if (sqrt(dx * dx + dy * dy) <= csize[2 * m] + c->obstacle_size[k][0]) return 1;
            } else {
                const double qo[7] = {ox, oy, 0.0, 1.0, 0.0, 0.0, 0.0};
                const double so[2] = {c->obstacle_size[k][0], c->obstacle_size[k][1]};
                if (fabs(q[0] - ox) <= so[0] && fabs(q[1] - oy) <= so[1]) return 1;
                if (gpro_rectangles_intersect(q, qo, csize + 2 * m, so)) return 1;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* the bodies at their initial positions (reset(), start / goal sampling) */
int gpro_check_obstacle_collision(const gpr_config* c, int n, const double* qpos, const double* csize) {
    return gpro_check_obstacle_collision_at(c, n, qpos, csize, 0.0);
}

/* ------------------------------------------------------------------------------------------------------------------ */
void gpro_planning_reward(const gpr_config* c, const double* achieved, const double* desired, int mover_collision,
                          int wall_collision, double* reward, int* terminated, int* is_success) {
    const int N = c->num_movers;
    int reached = 0;
    for (int m = 0; m < N; ++m) {
        double dx = achieved[2 * m] - desired[2 * m], dy = achieved[2 * m + 1] - desired[2 * m + 1];
        double d = sqrt(dx * dx + dy * dy); /* plan:647-664 */
        reached += d <= c->threshold_pos;   /* plan:521 */
    }
    const int coll = mover_collision || wall_collision;
    double r = -50.0 * (double)coll;                   /* plan:526 */
    r += -1.0 * (double)(N - reached) * (double)!coll; /* plan:527 */
    if (reached == N && !coll) r = 50.0;               /* plan:528 */
    *reward = r;
    *terminated = (r == 50.0) || (r == -50.0);                                /* plan:477-478 */
    *is_success = (reached == N) && !mover_collision && !wall_collision;      /* plan:597 */
}

/* push:499-527 compute_reward, push:457-476 compute_terminated, push:578-608 _get_info (one env) */
void gpro_pushing_reward(const gpr_config* c, const double* achieved, const double* desired, int wall_collision,
                         double* reward, int* terminated, int* is_success) {
    double dx = achieved[0] - desired[0], dy = achieved[1] - desired[1];
    double d = sqrt(dx * dx + dy * dy);
    int reached = d <= c->threshold_pos; /* push:519 */
    double r = -50.0 * (double)wall_collision;
    r += -1.0 * (double)!wall_collision;
    if (reached && !wall_collision) r = 0.0; /* push:523 */
    *reward = r;
    *terminated = r == -50.0; /* push:475 */
    *is_success = reached && !wall_collision; /* push:602 */
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Batched environment state (host pointers), same SoA layout as the CUDA handle's gpr_state                            */
/* ------------------------------------------------------------------------------------------------------------------ */
typedef struct gpro_state {
    double* pos;  /* [B, N, 2] */
    double* vel;  /* [B, N, 2] */
    double* acc;  /* [B, N, 2] qacc (== act in planning) */
    double* goal; /* [B, N, 2] */
    int32_t* elapsed_steps;
    uint32_t* rng_counter;
    uint8_t* needs_reset; /* NEXT_STEP autoreset bookkeeping */
    /* pushing only */
    double* act;        /* [B, 2] jerk integrator state */
    double* mover_rot;  /* [B, 3] cos yaw, sin yaw, yaw rate */
    double* object_pos; /* [B, 4] x, y, cos yaw, sin yaw */
    double* object_vel; /* [B, 3] vx, vy, yaw rate */
    float* contact_warm; /* [B, GPR_PUSH_WARM] warm-start state of the contact solve */
} gpro_state;

typedef struct gpro_outputs {
    double* observation;   /* [B, obs_dim] float64 (the product rounds the same values to float32) */
    double* achieved_goal; /* [B, goal_dim] */
    double* desired_goal;
    double* reward; /* [B] */
    uint8_t* terminated;
    uint8_t* truncated;
    uint8_t* is_success;
    uint8_t* mover_collision;
    uint8_t* wall_collision;
    double* final_observation;
    double* final_achieved_goal;
    double* final_desired_goal;
    uint8_t* reset_failed; /* [B] 1 if a rejection loop hit max_reset_attempts */
    uint8_t* other_collision; /* [B] a mover touches a static obstacle */
} gpro_outputs;

static void noisy_qpos(const gpr_config* c, const double* p, int N, const float (*nxy)[2], const float (*nq)[4],
                       double* qpos) {
    /* basic:801-828: qpos + N(0, sigma_p) on all 7 components (z is irrelevant to the planar checks and left 0) */
    const double sp = c->std_noise[0];
    for (int m = 0; m < N; ++m) {
        double* q = qpos + 7 * m;
        q[0] = p[2 * m];
        q[1] = p[2 * m + 1];
        q[2] = 0.0;
        q[3] = 1.0;
        q[4] = 0.0;
        q[5] = 0.0;
        q[6] = 0.0; /* planning movers never rotate: quaternion (1,0,0,0) (plan:364) */
        if (sp != 0.0 && nxy) {
            q[0] = q[0] + (double)nxy[m][0] * sp;
            q[1] = q[1] + (double)nxy[m][1] * sp;
            if (nq) {
                for (int k = 0; k < 4; ++k) q[3 + k] = q[3 + k] + (double)nq[m][k] * sp;
            }
        }
    }
}

static void csize_rows(const gpr_config* c, const double tab[GPR_MAX_MOVERS][2], double* out) {
    for (int m = 0; m < c->num_movers; ++m) {
        out[2 * m] = tab[m][0];
        out[2 * m + 1] = tab[m][1];
    }
}

/* plan:536-573 _get_obs (noisy position and velocity, exact acceleration) */
static void planning_obs(const gpr_config* c, const double* p, const double* v, const double* a, const double* g,
                         uint64_t seed, uint32_t env_global, uint32_t event, double* observation, double* achieved,
                         double* desired) {
    const int N = c->num_movers;
    for (int m = 0; m < N; ++m) {
        double px = p[2 * m], py = p[2 * m + 1], vx = v[2 * m], vy = v[2 * m + 1];
        if (c->std_noise[0] != 0.0 || c->std_noise[1] != 0.0) {
            float n4[4];
            gpr_normal4(seed, env_global, event, GPR_RNG_OBS, (uint32_t)m, n4);
            px = px + (double)n4[0] * c->std_noise[0];
            py = py + (double)n4[1] * c->std_noise[0];
            vx = vx + (double)n4[2] * c->std_noise[1];
            vy = vy + (double)n4[3] * c->std_noise[1];
        }
        achieved[2 * m] = px;
        achieved[2 * m + 1] = py;
        desired[2 * m] = g[2 * m];
        desired[2 * m + 1] = g[2 * m + 1];
        observation[2 * m] = vx;
        observation[2 * m + 1] = vy;
        if (c->learn_jerk) { /* plan:560: velocities of all movers first, then accelerations */
            observation[2 * N + 2 * m] = a[2 * m];
            observation[2 * N + 2 * m + 1] = a[2 * m + 1];
        }
    }
}

/* plan:355-418 + basic:1770-1833 for one env.  Returns 1 if a rejection loop ran out of attempts. */
static int planning_reset_one(const gpr_config* c, uint64_t seed, uint32_t env_global, uint32_t event, double* p,
                              double* v, double* a, double* g, const double* inj_start, const double* inj_goal,
                              int* mover_collision, int* wall_collision, int* other_collision) {
    const int N = c->num_movers;
    double qpos[7 * NMAX], cw[2 * NMAX], cm[2 * NMAX];
    int32_t valid[NMAX];
    int failed = 0;
    const int cap = c->max_reset_attempts > 0 ? c->max_reset_attempts : 1;
    /* plan:369-385 rejection loop A: all starts at once */
    if (inj_start) {
        memcpy(p, inj_start, sizeof(double) * 2 * N);
    } else {
        csize_rows(c, c->c_wall[1], cw);
        csize_rows(c, c->c_mover[1], cm);
        int ok = 0;
        for (int t = 0; t < cap && !ok; ++t) {
            for (int m = 0; m < N; ++m) {
                double ux, uy;
                gpr_sample_xy(seed, env_global, event, GPR_RNG_RESET_SAMPLE, 0u, (uint32_t)t, (uint32_t)m, &ux, &uy);
                /* plan:377 np_random.uniform(low, high): low + (high-low)*u */
                p[2 * m] = c->min_xy_pos[0] + (c->max_xy_pos[0] - c->min_xy_pos[0]) * ux;
                p[2 * m + 1] = c->min_xy_pos[1] + (c->max_xy_pos[1] - c->min_xy_pos[1]) * uy;
            }
            noisy_qpos(c, p, N, NULL, NULL, qpos); /* no noise inside the sampling loop (plan:379-383 pass qpos) */
            for (int m = 0; m < N; ++m) { qpos[7 * m] = p[2 * m]; qpos[7 * m + 1] = p[2 * m + 1]; }
            gpro_qpos_is_valid(c, N, qpos, cw, valid);
            int allv = 1;
            for (int m = 0; m < N; ++m) allv &= valid[m] != 0;
            ok = allv && !gpro_check_mover_collision(c, N, qpos, cm);
            if (ok && c->num_obstacles > 0) ok = !gpro_check_obstacle_collision(c, N, qpos, cm); /* starts clear the obstacles */
        }
        failed |= !ok;
    }
    /* plan:395-413 rejection loop B: all goals at once; strict '<' on min_goal_dist rejects */
    if (inj_goal) {
        memcpy(g, inj_goal, sizeof(double) * 2 * N);
    } else {
        csize_rows(c, c->c_wall[1], cw);
        int ok = 0;
        for (int t = 0; t < cap && !ok; ++t) {
            for (int m = 0; m < N; ++m) {
                double ux, uy;
                gpr_sample_xy(seed, env_global, event, GPR_RNG_RESET_SAMPLE, 1u, (uint32_t)t, (uint32_t)m, &ux, &uy);
                g[2 * m] = c->min_xy_pos[0] + (c->max_xy_pos[0] - c->min_xy_pos[0]) * ux;
                g[2 * m + 1] = c->min_xy_pos[1] + (c->max_xy_pos[1] - c->min_xy_pos[1]) * uy;
            }
            for (int m = 0; m < N; ++m) {
                double* q = qpos + 7 * m;
                q[0] = g[2 * m]; q[1] = g[2 * m + 1]; q[2] = 0; q[3] = 1; q[4] = 0; q[5] = 0; q[6] = 0;
            }
            gpro_qpos_is_valid(c, N, qpos, cw, valid);
            int allv = 1;
            for (int m = 0; m < N; ++m) allv &= valid[m] != 0;
            ok = allv;
            for (int i = 0; i < N && ok; ++i)
                for (int j = i + 1; j < N && ok; ++j) {
                    double dx = g[2 * i] - g[2 * j], dy = g[2 * i + 1] - g[2 * j + 1];
                    if (sqrt(dx * dx + dy * dy) < c->min_goal_dist) ok = 0; /* plan:410 */
                }
            if (ok && c->num_obstacles > 0) { /* a goal must be reachable: the mover's shape there clears the obstacles */
                csize_rows(c, c->c_mover[1], cm);
                ok = !gpro_check_obstacle_collision(c, N, qpos, cm);
            }
        }
        failed |= !ok;
    }
    /* plan:336-353 reload_model: fresh MjData => qvel = 0, act = 0, qacc = 0 */
    for (int k = 0; k < 2 * N; ++k) {
        v[k] = 0.0;
        a[k] = 0.0;
    }
    /* basic:1799-1805: wall check WITH safety offset, mover check WITHOUT, both on freshly noisy qpos */
    float nxy_w[NMAX][2], nxy_m[NMAX][2], nq_w[NMAX][4], nq_m[NMAX][4];
    const int noisy = c->std_noise[0] != 0.0;
    const int box = c->c_shape == GPR_SHAPE_BOX;
    if (noisy) {
        for (int m = 0; m < N; ++m) {
            float n4[4];
            gpr_normal4(seed, env_global, event, GPR_RNG_RESET_CHECK, (uint32_t)m, n4);
            nxy_w[m][0] = n4[0]; nxy_w[m][1] = n4[1]; nxy_m[m][0] = n4[2]; nxy_m[m][1] = n4[3];
            if (box) {
                gpr_normal4(seed, env_global, event, GPR_RNG_RESET_CHECK_WQUAT, (uint32_t)m, nq_w[m]);
                gpr_normal4(seed, env_global, event, GPR_RNG_RESET_CHECK_MQUAT, (uint32_t)m, nq_m[m]);
            }
        }
    }
    csize_rows(c, c->c_wall[1], cw);
    csize_rows(c, c->c_mover[0], cm);
    noisy_qpos(c, p, N, noisy ? nxy_w : NULL, (noisy && box) ? nq_w : NULL, qpos);
    gpro_qpos_is_valid(c, N, qpos, cw, valid);
    int wc = 0;
    for (int m = 0; m < N; ++m) wc |= !valid[m];
    /* basic:1807: the hook, on the wall check's noisy qpos, with the safety offset like the wall check of reset() */
    *other_collision = 0;
    if (c->num_obstacles > 0) {
        double cs[2 * NMAX];
        csize_rows(c, c->c_mover[1], cs);
        *other_collision = gpro_check_obstacle_collision(c, N, qpos, cs);
    }
    noisy_qpos(c, p, N, (noisy && N > 1) ? nxy_m : NULL, (noisy && box && N > 1) ? nq_m : NULL, qpos);
    *wall_collision = wc;
    *mover_collision = gpro_check_mover_collision(c, N, qpos, cm);
    return failed;
}

/* basic:1835-1950 step for one planning env (state updated in place). */
static void planning_step_one(const gpr_config* c, uint64_t seed, uint32_t env_global, uint32_t event, int elapsed,
                              double* p, double* v, double* a, const float* action, int* mover_collision,
                              int* wall_collision, int* other_collision) {
    const int N = c->num_movers;
    const double dt = c->cycle_time;
    const double lim = c->learn_jerk ? c->j_max : c->a_max;
    const int noisy_p = c->std_noise[0] != 0.0, noisy_v = c->std_noise[1] != 0.0;
    const int box = c->c_shape == GPR_SHAPE_BOX;
    double act[2 * NMAX];
    /* basic:1869-1873: clip to the action Box */
    for (int k = 0; k < 2 * N; ++k) {
        double u = (double)action[k];
        act[k] = u < -lim ? -lim : (u > lim ? lim : u);
    }
    double qpos[7 * NMAX], cw[2 * NMAX], cm[2 * NMAX];
    int32_t valid[NMAX];
    csize_rows(c, c->c_wall[0], cw);
    csize_rows(c, c->c_mover[0], cm);
    int mc = 0, wc = 0, oc = 0;
    for (int cyc = 0; cyc < c->num_cycles; ++cyc) { /* basic:1879 */
        float nxy_w[NMAX][2], nxy_m[NMAX][2], nq_w[NMAX][4], nq_m[NMAX][4];
        for (int m = 0; m < N; ++m) {
            float n4[4] = {0.f, 0.f, 0.f, 0.f};
            if (noisy_p || noisy_v)
                gpr_normal4(seed, env_global, event, (uint32_t)cyc * 4u + GPR_RNG_BLOCK_VEL_WALL, (uint32_t)m, n4);
            nxy_w[m][0] = n4[2];
            nxy_w[m][1] = n4[3];
            if (noisy_p && N > 1) {
                float k4[4];
                gpr_normal4(seed, env_global, event, (uint32_t)cyc * 4u + GPR_RNG_BLOCK_MOVER, (uint32_t)m, k4);
                nxy_m[m][0] = k4[0];
                nxy_m[m][1] = k4[1];
            }
            if (noisy_p && box) {
                gpr_normal4(seed, env_global, event, (uint32_t)cyc * 4u + GPR_RNG_BLOCK_WALL_QUAT, (uint32_t)m, nq_w[m]);
                if (N > 1)
                    gpr_normal4(seed, env_global, event, (uint32_t)cyc * 4u + GPR_RNG_BLOCK_MOVER_QUAT, (uint32_t)m,
                                nq_m[m]);
            }
            /* plan:420-450 _mujoco_step_callback */
            double vel[2] = {v[2 * m], v[2 * m + 1]};
            if (noisy_v) { /* plan:430 get_mover_qvel(add_noise=True) */
                vel[0] = vel[0] + (double)n4[0] * c->std_noise[1];
                vel[1] = vel[1] + (double)n4[1] * c->std_noise[1];
            }
            double ctrl[2];
            if (c->learn_jerk) {
                double acc[2] = {a[2 * m], a[2 * m + 1]}; /* plan:433 qacc, no noise */
                double next_acc_tmp[2], next_jerk[2], tmpv[2], next_acc[2];
                gpro_ensure_max_dyn_val(acc, c->a_max, &act[2 * m], dt, next_acc_tmp, next_jerk); /* plan:434 */
                gpro_ensure_max_dyn_val(vel, c->v_max, next_acc_tmp, dt, tmpv, next_acc);          /* plan:437 */
                if (next_acc_tmp[0] != next_acc[0] || next_acc_tmp[1] != next_acc[1]) {            /* plan:438 */
                    next_jerk[0] = (next_acc[0] - acc[0]) / dt;
                    next_jerk[1] = (next_acc[1] - acc[1]) / dt;
                }
                ctrl[0] = next_jerk[0];
                ctrl[1] = next_jerk[1];
                /* mj_step (basic:1882), actuator dyntype=integrator actearly=true (plan:305-311):
                   act <- act + dt*ctrl, force uses the new act, qacc = act */
                a[2 * m] = a[2 * m] + dt * ctrl[0];
                a[2 * m + 1] = a[2 * m + 1] + dt * ctrl[1];
            } else {
                double tmpv[2], next_acc[2];
                gpro_ensure_max_dyn_val(vel, c->v_max, &act[2 * m], dt, tmpv, next_acc); /* plan:442 */
                /* dyntype=none, gain=mass (plan:314-320): qacc = ctrl */
                a[2 * m] = next_acc[0];
                a[2 * m + 1] = next_acc[1];
            }
            /* MuJoCo semi-implicit Euler: qvel += dt*qacc; qpos += dt*qvel (tests/test_benchmark_planning_env.py:86-93) */
            v[2 * m] = v[2 * m] + dt * a[2 * m];
            v[2 * m + 1] = v[2 * m + 1] + dt * a[2 * m + 1];
            p[2 * m] = p[2 * m] + dt * v[2 * m];
            p[2 * m + 1] = p[2 * m + 1] + dt * v[2 * m + 1];
        }
        /* basic:1888-1894 wall check on noisy qpos, no safety offset */
        noisy_qpos(c, p, N, noisy_p ? nxy_w : NULL, (noisy_p && box) ? nq_w : NULL, qpos);
        gpro_qpos_is_valid(c, N, qpos, cw, valid);
        wc = 0;
        for (int m = 0; m < N; ++m) wc |= !valid[m];
        /* basic:1903 the hook: static obstacles, on the wall check's noisy qpos, no safety offset */
        /* (extra bodies with a prescribed velocity have moved for as long as the movers have been integrated) */
        oc = c->num_obstacles > 0
                 ? gpro_check_obstacle_collision_at(c, N, qpos, cm, (double)(elapsed * c->num_cycles + cyc + 1) * dt)
                 : 0;
        /* basic:1895-1901 mover check on an independent noisy qpos */
        noisy_qpos(c, p, N, (noisy_p && N > 1) ? nxy_m : NULL, (noisy_p && box && N > 1) ? nq_m : NULL, qpos);
        mc = gpro_check_mover_collision(c, N, qpos, cm);
        if (mc || wc || oc) break; /* basic:1904 */
    }
    *mover_collision = mc;
    *wall_collision = wc;
    *other_collision = oc;
}

static void write_obs(const gpr_config* c, int obs_dim, int goal_dim, const double* o, const double* ag,
                      const double* dg, int64_t e, double* O, double* AG, double* DG) {
    (void)c;
    if (O) memcpy(O + e * obs_dim, o, sizeof(double) * obs_dim);
    if (AG) memcpy(AG + e * goal_dim, ag, sizeof(double) * goal_dim);
    if (DG) memcpy(DG + e * goal_dim, dg, sizeof(double) * goal_dim);
}

/* Reset the masked envs (mask NULL = all). inject_* NULL = sample. */
void gpro_planning_reset(const gpr_config* c, uint64_t seed, gpro_state* s, const uint8_t* mask,
                         const double* inject_start, const double* inject_goal, gpro_outputs* out, int nthreads) {
    const int N = c->num_movers, B = c->num_envs;
    const int obs_dim = 2 * N * (1 + (c->learn_jerk != 0)), goal_dim = 2 * N;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (int64_t e = 0; e < B; ++e) {
        if (mask && !mask[e]) continue;
        double* p = s->pos + e * 2 * N;
        double* v = s->vel + e * 2 * N;
        double* a = s->acc + e * 2 * N;
        double* g = s->goal + e * 2 * N;
        uint32_t env_global = (uint32_t)(c->env_index_base + e);
        uint32_t event = s->rng_counter[e];
        int mc, wc, oc;
        int failed = planning_reset_one(c, seed, env_global, event, p, v, a, g,
                                        inject_start ? inject_start + e * 2 * N : NULL,
                                        inject_goal ? inject_goal + e * 2 * N : NULL, &mc, &wc, &oc);
        double o[4 * NMAX], ag[2 * NMAX], dg[2 * NMAX];
        planning_obs(c, p, v, a, g, seed, env_global, event, o, ag, dg);
        s->rng_counter[e] = event + 1u;
        s->elapsed_steps[e] = 0;
        if (s->needs_reset) s->needs_reset[e] = 0;
        if (out) {
            write_obs(c, obs_dim, goal_dim, o, ag, dg, e, out->observation, out->achieved_goal, out->desired_goal);
            double r;
            int term, succ;
            gpro_planning_reward(c, ag, dg, mc || oc, wc, &r, &term, &succ); /* an obstacle hit counts as a collision */
            if (out->is_success) out->is_success[e] = (uint8_t)succ;
            if (out->mover_collision) out->mover_collision[e] = (uint8_t)mc;
            if (out->wall_collision) out->wall_collision[e] = (uint8_t)wc;
            if (out->other_collision) out->other_collision[e] = (uint8_t)oc;
            if (out->reset_failed) out->reset_failed[e] = (uint8_t)failed;
        }
    }
    (void)nthreads;
}

/* One env-step for all B envs, including TimeLimit and auto-reset. action: [B, 2N] float32. */
void gpro_planning_step(const gpr_config* c, uint64_t seed, gpro_state* s, const float* action, gpro_outputs* out,
                        int nthreads) {
    const int N = c->num_movers, B = c->num_envs;
    const int obs_dim = 2 * N * (1 + (c->learn_jerk != 0)), goal_dim = 2 * N;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (int64_t e = 0; e < B; ++e) {
        double* p = s->pos + e * 2 * N;
        double* v = s->vel + e * 2 * N;
        double* a = s->acc + e * 2 * N;
        double* g = s->goal + e * 2 * N;
        uint32_t env_global = (uint32_t)(c->env_index_base + e);
        double o[4 * NMAX], ag[2 * NMAX], dg[2 * NMAX];
        int mc = 0, wc = 0, oc = 0;
        if (c->autoreset_mode == GPR_AUTORESET_NEXT_STEP && s->needs_reset && s->needs_reset[e]) {
            /* gymnasium vector NEXT_STEP: this call resets instead of stepping; reward 0, not done */
            uint32_t event = s->rng_counter[e];
            planning_reset_one(c, seed, env_global, event, p, v, a, g, NULL, NULL, &mc, &wc, &oc);
            planning_obs(c, p, v, a, g, seed, env_global, event, o, ag, dg);
            s->rng_counter[e] = event + 1u;
            s->elapsed_steps[e] = 0;
            s->needs_reset[e] = 0;
            write_obs(c, obs_dim, goal_dim, o, ag, dg, e, out->observation, out->achieved_goal, out->desired_goal);
            double r;
            int term, succ;
            gpro_planning_reward(c, ag, dg, mc || oc, wc, &r, &term, &succ);
            if (out->reward) out->reward[e] = 0.0;
            if (out->terminated) out->terminated[e] = 0;
            if (out->truncated) out->truncated[e] = 0;
            if (out->is_success) out->is_success[e] = (uint8_t)succ;
            if (out->mover_collision) out->mover_collision[e] = (uint8_t)mc;
            if (out->wall_collision) out->wall_collision[e] = (uint8_t)wc;
            if (out->other_collision) out->other_collision[e] = (uint8_t)oc;
            continue;
        }
        uint32_t event = s->rng_counter[e];
        planning_step_one(c, seed, env_global, event, s->elapsed_steps[e], p, v, a, action + e * 2 * N, &mc, &wc, &oc);
        planning_obs(c, p, v, a, g, seed, env_global, event, o, ag, dg);
        s->rng_counter[e] = event + 1u;
        double r;
        int term, succ;
        gpro_planning_reward(c, ag, dg, mc || oc, wc, &r, &term, &succ); /* an obstacle hit counts as a collision */
        int steps = s->elapsed_steps[e] + 1;
        s->elapsed_steps[e] = steps;
        int trunc = (c->max_episode_steps > 0) && (steps >= c->max_episode_steps); /* gymnasium TimeLimit */
        if (out->reward) out->reward[e] = r;
        if (out->terminated) out->terminated[e] = (uint8_t)term;
        if (out->truncated) out->truncated[e] = (uint8_t)trunc;
        if (out->is_success) out->is_success[e] = (uint8_t)succ;
        if (out->mover_collision) out->mover_collision[e] = (uint8_t)mc;
        if (out->wall_collision) out->wall_collision[e] = (uint8_t)wc;
        if (out->other_collision) out->other_collision[e] = (uint8_t)oc;
        const int done = term || trunc;
        if (done && c->autoreset_mode == GPR_AUTORESET_SAME_STEP) {
            write_obs(c, obs_dim, goal_dim, o, ag, dg, e, out->final_observation, out->final_achieved_goal,
                      out->final_desired_goal);
            uint32_t ev2 = s->rng_counter[e];
            int mc2, wc2, oc2;
            int failed = planning_reset_one(c, seed, env_global, ev2, p, v, a, g, NULL, NULL, &mc2, &wc2, &oc2);
            planning_obs(c, p, v, a, g, seed, env_global, ev2, o, ag, dg);
            s->rng_counter[e] = ev2 + 1u;
            s->elapsed_steps[e] = 0;
            if (out->reset_failed) out->reset_failed[e] = (uint8_t)failed;
        } else if (done && c->autoreset_mode == GPR_AUTORESET_NEXT_STEP && s->needs_reset) {
            s->needs_reset[e] = 1;
        }
        write_obs(c, obs_dim, goal_dim, o, ag, dg, e, out->observation, out->achieved_goal, out->desired_goal);
    }
    (void)nthreads;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* BenchmarkPushingEnv (push = envs/manipulation/benchmark_pushing_env.py)                                              */
/*                                                                                                                      */
/* Everything AROUND mj_step is a restatement of the reference (control limiting push:419-455, wall check                */
/* basic:1888-1894, observation push:529-576, reward / terminated / info push:457-527, 578-608, reset push:373-417).      */
/* mj_step itself (mover-object contact, object-ground friction, yaw impedance) is the planar specification in          */
/* include/gpr_push_physics.h — MuJoCo is not available, PARITY UNPINNED for that part.                 */
/* ------------------------------------------------------------------------------------------------------------------ */
typedef struct push_env {
    gpr_body2 M, O;
    double acc[2], act[2], goal[2];
    float warm[GPR_PUSH_WARM];
} push_env;

static void push_env_load(const gpro_state* s, int64_t e, push_env* x) {
    x->M.x = s->pos[2 * e];
    x->M.y = s->pos[2 * e + 1];
    x->M.vx = s->vel[2 * e];
    x->M.vy = s->vel[2 * e + 1];
    x->M.c = s->mover_rot[3 * e];
    x->M.s = s->mover_rot[3 * e + 1];
    x->M.w = s->mover_rot[3 * e + 2];
    x->O.x = s->object_pos[4 * e];
    x->O.y = s->object_pos[4 * e + 1];
    x->O.c = s->object_pos[4 * e + 2];
    x->O.s = s->object_pos[4 * e + 3];
    x->O.vx = s->object_vel[3 * e];
    x->O.vy = s->object_vel[3 * e + 1];
    x->O.w = s->object_vel[3 * e + 2];
    for (int k = 0; k < 2; ++k) {
        x->acc[k] = s->acc[2 * e + k];
        x->act[k] = s->act[2 * e + k];
        x->goal[k] = s->goal[2 * e + k];
    }
    for (int i = 0; i < GPR_PUSH_WARM; ++i) x->warm[i] = s->contact_warm ? s->contact_warm[GPR_PUSH_WARM * e + i] : 0.0f;
}

static void push_env_store(gpro_state* s, int64_t e, const push_env* x) {
    s->pos[2 * e] = x->M.x;
    s->pos[2 * e + 1] = x->M.y;
    s->vel[2 * e] = x->M.vx;
    s->vel[2 * e + 1] = x->M.vy;
    s->mover_rot[3 * e] = x->M.c;
    s->mover_rot[3 * e + 1] = x->M.s;
    s->mover_rot[3 * e + 2] = x->M.w;
    s->object_pos[4 * e] = x->O.x;
    s->object_pos[4 * e + 1] = x->O.y;
    s->object_pos[4 * e + 2] = x->O.c;
    s->object_pos[4 * e + 3] = x->O.s;
    s->object_vel[3 * e] = x->O.vx;
    s->object_vel[3 * e + 1] = x->O.vy;
    s->object_vel[3 * e + 2] = x->O.w;
    for (int k = 0; k < 2; ++k) {
        s->acc[2 * e + k] = x->acc[k];
        s->act[2 * e + k] = x->act[k];
        s->goal[2 * e + k] = x->goal[k];
    }
    if (s->contact_warm)
        for (int i = 0; i < GPR_PUSH_WARM; ++i) s->contact_warm[GPR_PUSH_WARM * e + i] = x->warm[i];
}

/* basic:1888-1894 / 1799-1801: check_wall_collision on the noisy qpos of the single mover.
 * nxy: position noise (2 normals) or NULL; qstream: RNG stream of the quaternion noise (box shape only). */
static int push_wall_collision(const gpr_config* c, const gpr_body2* M, int safety, const float* nxy, uint64_t seed,
                               uint32_t env_global, uint32_t event, uint32_t qstream) {
    const double sp = c->std_noise[0];
    const int noisy = (sp != 0.0 || c->std_noise[1] != 0.0) && nxy != NULL;
    double qpos[7] = {M->x, M->y, 0.0, 1.0, 0.0, 0.0, 0.0};
    if (noisy) {
        qpos[0] = M->x + (double)nxy[0] * sp;
        qpos[1] = M->y + (double)nxy[1] * sp;
    }
    if (c->c_shape == GPR_SHAPE_BOX) {
        /* the mover's yaw as MuJoCo's quaternion (cos(yaw/2), 0, 0, sin(yaw/2)), by the half-angle identities */
        const double ch = sqrt(0.5 * (1.0 + M->c));
        const double sh = M->s / (2.0 * ch);
        qpos[3] = ch;
        qpos[6] = sh;
        if (noisy) { /* basic:828: noise on all quaternion components */
            float q[4];
            gpr_normal4(seed, env_global, event, qstream, 0u, q);
            qpos[3] = qpos[3] + (double)q[0] * sp;
            qpos[4] = (double)q[1] * sp;
            qpos[5] = (double)q[2] * sp;
            qpos[6] = qpos[6] + (double)q[3] * sp;
        }
    }
    double cw[2] = {c->c_wall[safety][0][0], c->c_wall[safety][0][1]};
    int32_t valid;
    gpro_qpos_is_valid(c, 1, qpos, cw, &valid);
    return !valid;
}

/* push:529-576 _get_obs: [pos, vel, (acc)] of the mover; achieved = object xy + N(0, 1e-5) (push:565) */
static void pushing_obs(const gpr_config* c, const push_env* x, uint64_t seed, uint32_t env_global, uint32_t event,
                        double* observation, double* achieved, double* desired) {
    double px = x->M.x, py = x->M.y, vx = x->M.vx, vy = x->M.vy;
    if (c->std_noise[0] != 0.0 || c->std_noise[1] != 0.0) {
        float n4[4];
        gpr_normal4(seed, env_global, event, GPR_RNG_OBS, 0u, n4);
        px = px + (double)n4[0] * c->std_noise[0];
        py = py + (double)n4[1] * c->std_noise[0];
        vx = vx + (double)n4[2] * c->std_noise[1];
        vy = vy + (double)n4[3] * c->std_noise[1];
    }
    observation[0] = px;
    observation[1] = py;
    observation[2] = vx;
    observation[3] = vy;
    if (c->learn_jerk) { /* push:556: qacc without noise */
        observation[4] = x->acc[0];
        observation[5] = x->acc[1];
    }
    achieved[0] = x->O.x;
    achieved[1] = x->O.y;
    if (c->object_noise_xy != 0.0) {
        float k4[4];
        gpr_normal4(seed, env_global, event, GPR_RNG_OBJECT, 0u, k4);
        achieved[0] = achieved[0] + (double)k4[0] * c->object_noise_xy;
        achieved[1] = achieved[1] + (double)k4[1] * c->object_noise_xy;
    }
    desired[0] = x->goal[0];
    desired[1] = x->goal[1];
}

/* push:373-417 _reset_callback + push:353-371 reload_model + basic:1797-1801.  Returns 1 if the object loop ran out. */
static int pushing_reset_one(const gpr_config* c, uint64_t seed, uint32_t env_global, uint32_t event, push_env* x,
                             const double* inj_start, const double* inj_goal, const double* inj_object,
                             int* wall_collision) {
    double ux, uy;
    int failed = 0;
    if (inj_start) {
        x->M.x = inj_start[0];
        x->M.y = inj_start[1];
    } else { /* push:387-389 */
        gpr_sample_xy(seed, env_global, event, GPR_RNG_RESET_SAMPLE, 0u, 0u, 0u, &ux, &uy);
        x->M.x = c->min_xy_pos[0] + (c->max_xy_pos[0] - c->min_xy_pos[0]) * ux;
        x->M.y = c->min_xy_pos[1] + (c->max_xy_pos[1] - c->min_xy_pos[1]) * uy;
    }
    if (inj_object) {
        x->O.x = inj_object[0];
        x->O.y = inj_object[1];
    } else { /* push:392-407: redraw the object until it is farther than min_mo_dist from the mover (strict '>') */
        const int cap = c->max_reset_attempts > 0 ? c->max_reset_attempts : 1;
        int ok = 0;
        for (int t = 0; t < cap && !ok; ++t) {
            gpr_sample_xy(seed, env_global, event, GPR_RNG_RESET_OBJECT, 0u, (uint32_t)t, 0u, &ux, &uy);
            x->O.x = c->object_min_xy_pos[0] + (c->object_max_xy_pos[0] - c->object_min_xy_pos[0]) * ux;
            x->O.y = c->object_min_xy_pos[1] + (c->object_max_xy_pos[1] - c->object_min_xy_pos[1]) * uy;
            double dx = x->O.x - x->M.x, dy = x->O.y - x->M.y;
            ok = sqrt(dx * dx + dy * dy) > c->min_mo_dist; /* push:407 */
        }
        failed = !ok;
    }
    if (inj_goal) {
        x->goal[0] = inj_goal[0];
        x->goal[1] = inj_goal[1];
    } else { /* push:411-413 */
        gpr_sample_xy(seed, env_global, event, GPR_RNG_RESET_OBJECT, 1u, 0u, 0u, &ux, &uy);
        x->goal[0] = c->object_min_xy_pos[0] + (c->object_max_xy_pos[0] - c->object_min_xy_pos[0]) * ux;
        x->goal[1] = c->object_min_xy_pos[1] + (c->object_max_xy_pos[1] - c->object_min_xy_pos[1]) * uy;
    }
    /* reload_model (push:353-371): fresh MjData — everything at rest, identity orientations */
    x->M.vx = x->M.vy = x->M.w = 0.0;
    x->M.c = 1.0;
    x->M.s = 0.0;
    x->O.vx = x->O.vy = x->O.w = 0.0;
    x->O.c = 1.0;
    x->O.s = 0.0;
    x->acc[0] = x->acc[1] = x->act[0] = x->act[1] = 0.0;
    for (int i = 0; i < GPR_PUSH_WARM; ++i) x->warm[i] = 0.0f; /* fresh MjData: no warm-start forces */
    /* basic:1799-1801: wall check WITH the safety offset on noisy qpos */
    float n4[4] = {0.f, 0.f, 0.f, 0.f};
    const int noisy = c->std_noise[0] != 0.0 || c->std_noise[1] != 0.0;
    if (noisy) gpr_normal4(seed, env_global, event, GPR_RNG_RESET_CHECK, 0u, n4);
    *wall_collision = push_wall_collision(c, &x->M, 1, noisy ? n4 : NULL, seed, env_global, event, GPR_RNG_RESET_CHECK_WQUAT);
    return failed;
}

/* basic:1835-1950 step for one pushing env. */
static void pushing_step_one(const gpr_config* c, const gpr_push_params* P, uint64_t seed, uint32_t env_global,
                             uint32_t event, push_env* x, const float* action, int* wall_collision) {
    const double dt = c->cycle_time;
    const double lim = c->learn_jerk ? c->j_max : c->a_max;
    const int noisy = c->std_noise[0] != 0.0 || c->std_noise[1] != 0.0;
    double u[2];
    for (int k = 0; k < 2; ++k) { /* basic:1869-1873 */
        double a = (double)action[k];
        u[k] = a < -lim ? -lim : (a > lim ? lim : a);
    }
    int wc = 0;
    for (int cyc = 0; cyc < c->num_cycles; ++cyc) { /* basic:1879 */
        float n4[4] = {0.f, 0.f, 0.f, 0.f};
        if (noisy) gpr_normal4(seed, env_global, event, (uint32_t)cyc * 4u + GPR_RNG_BLOCK_VEL_WALL, 0u, n4);
        /* push:419-455 _mujoco_step_callback */
        double vel[2] = {x->M.vx, x->M.vy};
        if (noisy) { /* push:428 get_mover_qvel(add_noise=True) */
            vel[0] = vel[0] + (double)n4[0] * c->std_noise[1];
            vel[1] = vel[1] + (double)n4[1] * c->std_noise[1];
        }
        double ctrl[2];
        if (c->learn_jerk) {
            double next_acc_tmp[2], next_jerk[2], tmpv[2], next_acc[2];
            gpro_ensure_max_dyn_val(x->acc, c->a_max, u, dt, next_acc_tmp, next_jerk); /* push:432, acc = the real qacc */
            gpro_ensure_max_dyn_val(vel, c->v_max, next_acc_tmp, dt, tmpv, next_acc);  /* push:435 */
            if (next_acc_tmp[0] != next_acc[0] || next_acc_tmp[1] != next_acc[1]) {    /* push:436 */
                next_jerk[0] = (next_acc[0] - x->acc[0]) / dt;
                next_jerk[1] = (next_acc[1] - x->acc[1]) / dt;
            }
            /* integrator actuator, actearly (push:305-311): act += dt*ctrl; the actuator force uses the new act */
            x->act[0] = x->act[0] + dt * next_jerk[0];
            x->act[1] = x->act[1] + dt * next_jerk[1];
            ctrl[0] = x->act[0];
            ctrl[1] = x->act[1];
        } else {
            double tmpv[2];
            gpro_ensure_max_dyn_val(vel, c->v_max, u, dt, tmpv, ctrl); /* push:440 */
        }
        /* mj_step (basic:1882): planar substitute, see include/gpr_push_physics.h */
        double qax, qay;
        gpr_push_substep(P, &x->M, &x->O, ctrl[0], ctrl[1], &qax, &qay, x->warm);
        x->acc[0] = qax;
        x->acc[1] = qay;
        /* basic:1888-1894 wall check; a single mover has no mover-mover check (push:592 asserts no mover collision) */
        wc = push_wall_collision(c, &x->M, 0, noisy ? n4 + 2 : NULL, seed, env_global, event,
                                 (uint32_t)cyc * 4u + GPR_RNG_BLOCK_WALL_QUAT);
        if (wc) break; /* basic:1904 */
    }
    /* an env that ends its step in the free regime holds no warm-start forces (the next free substep would clear them
     * anyway; normalising here keeps the stored state independent of which kernel runs the next cycles) */
    if (gpr_push_is_free(P, &x->M, &x->O))
        for (int i = 0; i < GPR_PUSH_WARM; ++i) x->warm[i] = 0.0f;
    *wall_collision = wc;
}

void gpro_pushing_reset(const gpr_config* c, uint64_t seed, gpro_state* s, const uint8_t* mask,
                        const double* inject_start, const double* inject_goal, const double* inject_object,
                        gpro_outputs* out, int nthreads) {
    const int B = c->num_envs;
    const int obs_dim = 2 * (2 + (c->learn_jerk != 0)), goal_dim = 2;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (int64_t e = 0; e < B; ++e) {
        if (mask && !mask[e]) continue;
        push_env x;
        memset(&x, 0, sizeof(x));
        uint32_t env_global = (uint32_t)(c->env_index_base + e);
        uint32_t event = s->rng_counter[e];
        int wc;
        int failed = pushing_reset_one(c, seed, env_global, event, &x, inject_start ? inject_start + 2 * e : NULL,
                                       inject_goal ? inject_goal + 2 * e : NULL,
                                       inject_object ? inject_object + 2 * e : NULL, &wc);
        double o[6], ag[2], dg[2];
        pushing_obs(c, &x, seed, env_global, event, o, ag, dg);
        push_env_store(s, e, &x);
        s->rng_counter[e] = event + 1u;
        s->elapsed_steps[e] = 0;
        if (s->needs_reset) s->needs_reset[e] = 0;
        if (out) {
            write_obs(c, obs_dim, goal_dim, o, ag, dg, e, out->observation, out->achieved_goal, out->desired_goal);
            double r;
            int term, succ;
            gpro_pushing_reward(c, ag, dg, wc, &r, &term, &succ);
            if (out->is_success) out->is_success[e] = (uint8_t)succ;
            if (out->mover_collision) out->mover_collision[e] = 0;
            if (out->wall_collision) out->wall_collision[e] = (uint8_t)wc;
            if (out->reset_failed) out->reset_failed[e] = (uint8_t)failed;
        }
    }
    (void)nthreads;
}

GPRO_FMA_CLONES void gpro_pushing_step(const gpr_config* c, uint64_t seed, gpro_state* s, const float* action, gpro_outputs* out,
                       int nthreads) {
    const int B = c->num_envs;
    const int obs_dim = 2 * (2 + (c->learn_jerk != 0)), goal_dim = 2;
    gpr_push_params P;
    gpr_push_params_from_config(c, &P);
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (int64_t e = 0; e < B; ++e) {
        push_env x;
        push_env_load(s, e, &x);
        uint32_t env_global = (uint32_t)(c->env_index_base + e);
        double o[6] = {0, 0, 0, 0, 0, 0}, ag[2], dg[2];
        int wc = 0;
        if (c->autoreset_mode == GPR_AUTORESET_NEXT_STEP && s->needs_reset && s->needs_reset[e]) {
            uint32_t event = s->rng_counter[e];
            pushing_reset_one(c, seed, env_global, event, &x, NULL, NULL, NULL, &wc);
            pushing_obs(c, &x, seed, env_global, event, o, ag, dg);
            push_env_store(s, e, &x);
            s->rng_counter[e] = event + 1u;
            s->elapsed_steps[e] = 0;
            s->needs_reset[e] = 0;
            write_obs(c, obs_dim, goal_dim, o, ag, dg, e, out->observation, out->achieved_goal, out->desired_goal);
            double r;
            int term, succ;
            gpro_pushing_reward(c, ag, dg, wc, &r, &term, &succ);
            if (out->reward) out->reward[e] = 0.0;
            if (out->terminated) out->terminated[e] = 0;
            if (out->truncated) out->truncated[e] = 0;
            if (out->is_success) out->is_success[e] = (uint8_t)succ;
            if (out->mover_collision) out->mover_collision[e] = 0;
            if (out->wall_collision) out->wall_collision[e] = (uint8_t)wc;
            continue;
        }
        uint32_t event = s->rng_counter[e];
        pushing_step_one(c, &P, seed, env_global, event, &x, action + 2 * e, &wc);
        pushing_obs(c, &x, seed, env_global, event, o, ag, dg);
        s->rng_counter[e] = event + 1u;
        double r;
        int term, succ;
        gpro_pushing_reward(c, ag, dg, wc, &r, &term, &succ);
        int steps = s->elapsed_steps[e] + 1;
        s->elapsed_steps[e] = steps;
        int trunc = (c->max_episode_steps > 0) && (steps >= c->max_episode_steps);
        if (out->reward) out->reward[e] = r;
        if (out->terminated) out->terminated[e] = (uint8_t)term;
        if (out->truncated) out->truncated[e] = (uint8_t)trunc;
        if (out->is_success) out->is_success[e] = (uint8_t)succ;
        if (out->mover_collision) out->mover_collision[e] = 0;
        if (out->wall_collision) out->wall_collision[e] = (uint8_t)wc;
        const int done = term || trunc;
        if (done && c->autoreset_mode == GPR_AUTORESET_SAME_STEP) {
            write_obs(c, obs_dim, goal_dim, o, ag, dg, e, out->final_observation, out->final_achieved_goal,
                      out->final_desired_goal);
            uint32_t ev2 = s->rng_counter[e];
            int wc2;
            int failed = pushing_reset_one(c, seed, env_global, ev2, &x, NULL, NULL, NULL, &wc2);
            pushing_obs(c, &x, seed, env_global, ev2, o, ag, dg);
            s->rng_counter[e] = ev2 + 1u;
            s->elapsed_steps[e] = 0;
            if (out->reset_failed) out->reset_failed[e] = (uint8_t)failed;
        } else if (done && c->autoreset_mode == GPR_AUTORESET_NEXT_STEP && s->needs_reset) {
            s->needs_reset[e] = 1;
        }
        push_env_store(s, e, &x);
        write_obs(c, obs_dim, goal_dim, o, ag, dg, e, out->observation, out->achieved_goal, out->desired_goal);
    }
    (void)nthreads;
}

/* one substep of the planar push physics on explicit bodies (property tests of the specification) */
GPRO_FMA_CLONES int gpro_push_substep(const gpr_config* c, double* mover7, double* object7, double ux, double uy, double* qacc,
                                      float* warm /* [GPR_PUSH_WARM] in/out or NULL */) {
    gpr_push_params P;
    gpr_push_params_from_config(c, &P);
    gpr_body2 M = {mover7[0], mover7[1], mover7[2], mover7[3], mover7[4], mover7[5], mover7[6]};
    gpr_body2 O = {object7[0], object7[1], object7[2], object7[3], object7[4], object7[5], object7[6]};
    int nc = gpr_push_substep(&P, &M, &O, ux, uy, &qacc[0], &qacc[1], warm);
    double m[7] = {M.x, M.y, M.c, M.s, M.vx, M.vy, M.w}, o[7] = {O.x, O.y, O.c, O.s, O.vx, O.vy, O.w};
    memcpy(mover7, m, sizeof(m));
    memcpy(object7, o, sizeof(o));
    return nc;
}

/* HER relabelling on float32 goals (what gpr_compute_reward sees): goals are promoted to float64 first. */
void gpro_compute_reward(const gpr_config* c, int batch, const float* achieved, const float* desired,
                         const uint8_t* mover_collision, const uint8_t* wall_collision, float* reward,
                         uint8_t* terminated) {
    const int gd = c->env_kind == GPR_ENV_PLANNING ? 2 * c->num_movers : 2;
    for (int b = 0; b < batch; ++b) {
        double ag[2 * NMAX], dg[2 * NMAX];
        for (int k = 0; k < gd; ++k) {
            ag[k] = (double)achieved[(int64_t)b * gd + k];
            dg[k] = (double)desired[(int64_t)b * gd + k];
        }
        int mc = mover_collision ? mover_collision[b] != 0 : 0, wc = wall_collision ? wall_collision[b] != 0 : 0;
        double r;
        int t, s_;
        if (c->env_kind == GPR_ENV_PLANNING)
            gpro_planning_reward(c, ag, dg, mc, wc, &r, &t, &s_);
        else
            gpro_pushing_reward(c, ag, dg, wc, &r, &t, &s_);
        if (reward) reward[b] = (float)r;
        if (terminated) terminated[b] = (uint8_t)t;
    }
}

/* --- small helpers exported for the tests ---------------------------------------------------------------------------- */
void gpro_philox_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, int rounds, uint32_t out[4]) {
    gpr_u32x4 r = gpr_philox4x32_r(c0, c1, c2, c3, k0, k1, rounds);
    memcpy(out, r.v, sizeof(r.v));
}
void gpro_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    gpr_u32x4 r = gpr_philox4x32_10(c0, c1, c2, c3, k0, k1);
    memcpy(out, r.v, sizeof(r.v));
}
void gpro_normals(uint64_t seed, uint32_t env_global, uint32_t event, uint32_t stream, uint32_t lane0, int count4,
                  float* out) {
    for (int i = 0; i < count4; ++i) gpr_normal4(seed, env_global, event, stream, lane0 + (uint32_t)i, out + 4 * i);
}
uint32_t gpro_config_bytes(void) { return (uint32_t)sizeof(gpr_config); }
int gpro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
