/*
 * gpr_push_physics.h — planar ("contact-light") rigid-body substep of BenchmarkPushingEnv: one mover, one box object.
 *
 * STATUS: a SPECIFICATION of this project, not a restatement of reference source.  In the reference every line of this
 * physics lives inside the third-party MuJoCo C library (mj_step: soft-constraint contact with elliptic friction cones,
 * Newton solver), which is neither in /root/reference nor installable offline, and no reference test exercises contact
 * (SURVEY.md §0.5, §8c): PARITY WITH MUJOCO IS UNPINNED for everything in this file.  What IS pinned by the reference's
 * tests (tests/test_benchmark_pushing_env.py: with the object far away the mover follows the planning closed form) holds
 * by construction: without contact the mover integrates exactly like the planning env.
 *
 * Model (SURVEY.md spec 8; citations into gymnasium_planar_robotics/):
 *   bodies   mover  (x, y, yaw):  box 0.155 x 0.155 m, mass 1.24, gravity-compensated, hovering (basic_envs.py:878-879)
 *            object (x, y, yaw):  box 0.07 x 0.07 m, mass 0.01, free joint with damping 0.01 on every DOF, resting on the
 *                                 tiles (manipulation/benchmark_pushing_env.py:172-178, 332-342)
 *   forces   mover x,y  : actuator force m * u (u = limited acceleration / integrated jerk, pushing:305-321)
 *            mover yaw  : impedance controller  tau = k_r (0 - yaw) - 2 sqrt(k_r m) yaw_rate  (utils/impedance_control.py:
 *                         46-47 uses the body MASS for the rotational damping too); z / roll / pitch are dropped
 *            object     : joint damping (implicit, as MuJoCo's Euler integrator does), ground friction at the four bottom
 *                         corners (each carries m g / 4, friction coefficient mu), mover-object box-box contact
 *   precision body states, smooth forces and the integrator are float64.  The CONSTRAINT SOLVE (contact manifold in the
 *            mover's frame, constraint rows, projected Gauss-Seidel sweeps) is float32, like MuJoCo's own GPU port (MJX):
 *            constraint forces come out of a truncated iterative solver with ~1e-3 relative accuracy, float32 rounding
 *            (6e-8) is far below that, and the dependent float32 chain is what bounds the kernel's latency.  Every float32
 *            operation is an explicitly rounded IEEE operation (GPR_FMUL / GPR_FFMA / GPR_FDIV / GPR_FSQRT of gpr_rng.h),
 *            so the CPU oracle and the CUDA kernels still produce bit-identical trajectories.
 *   warm start  like MuJoCo (its default), the solver starts from the constraint forces of the previous substep: the four
 *            ground-friction forces always, the contact forces when the number of contact points is unchanged.  Measured
 *            on 400-substep pushes against the converged solution (tests/test_pushing_independent.py): 3 warm-started
 *            sweeps track it as closely as 8 cold ones (final object position within 2.4e-5 m vs 2.2e-5 m mean), so the
 *            default is contact_iterations = 3 — and the sweeps are the dependent chain that bounds the kernel's latency.
 *            The forces are 13 floats of per-environment state (gpr_state.contact_warm), zero at reset.
 *   contacts MuJoCo-style soft constraints solved by projected Gauss-Seidel in acceleration space:
 *            a_ref = -B v - K d(r) r,  B = 2/(dmax tc),  K = 1/(dmax^2 tc^2 dr^2),  R = (1 - d)/d * A_ii,
 *            solref = (tc, dr) = (0.02, 1), solimp = (0.9, 0.95, 0.001, 0.5, 2) (MuJoCo defaults), friction cone
 *            |f_t| <= mu f_n; planar box-box manifold (separating-axis test + incident-edge clipping, <= 2 points)
 *   integr.  semi-implicit Euler, dt = 1 ms: v += dt a; x += dt v; yaw advanced on the unit circle (cos, sin) with a
 *            second-order step and renormalisation — no transcendental function is called, so the CPU oracle
 *            (gcc -ffp-contract=off) and the CUDA kernels (nvcc -fmad=false) produce bit-identical trajectories.
 *
 * Everything is `static inline`, plain IEEE float64 arithmetic; compiles as C99, C++ and CUDA.
 */
#ifndef GPR_PUSH_PHYSICS_H_
#define GPR_PUSH_PHYSICS_H_

#include <math.h>

#include "gpr.h"
#include "gpr_rng.h" /* the explicitly rounded float32 primitives */

#if defined(__CUDACC__)
#define GPR_PHD __host__ __device__ __forceinline__
#else
#define GPR_PHD static inline
#endif

typedef struct gpr_push_params {
    double dt;
    double mover_mass, mover_inertia, mover_hx, mover_hy;
    double obj_mass, obj_inertia, obj_h, obj_damping;
    double mu, gravity;
    double k_rot, d_rot;          /* yaw impedance: stiffness, damping */
    double sol_B, sol_K;          /* B = 2/(dmax tc), K = 1/(dmax^2 tc^2 dr^2) */
    double imp_d0, imp_dw, imp_width, imp_mid, imp_power;
    double contact_r2; /* (sum of the two circumradii * 1.0001)^2: centres farther apart cannot touch */
    double g_R, g_inv; /* ground-friction rows: regulariser R and 1 / (A + R) — the same for all four corners */
    double g_lim;      /* friction limit per corner: mu * m g / 4 */
    double o_inv_lin, o_inv_rot; /* 1 / (m + dt D), 1 / (I + dt D): implicit joint damping of the object */
    int iterations;
    int warm_start; /* start the sweeps from the previous substep's forces (gpr_config.contact_warm_start) */
    /* float32 copies for the constraint solve (each the float64 value rounded once) */
    float f_hxM, f_hyM, f_hO, f_imM, f_iIM, f_imO, f_iIO, f_mu, f_B, f_K;
    float f_d0, f_dw, f_width, f_mid, f_power, f_gR, f_ginv, f_glim, f_glim2, f_D;
    float f_inv_width, f_inv_mid, f_inv_1m_mid; /* reciprocals of the solimp shape constants (each computed in float64, rounded once) */
} gpr_push_params;

/* Derived physics parameters from the config — ONE definition shared by the CUDA host code and the CPU oracle so both
 * sides see bit-identical constants. */
static inline void gpr_push_params_from_config(const gpr_config* c, gpr_push_params* P) {
    const double lx = 2.0 * c->mover_half[0], ly = 2.0 * c->mover_half[1], lo = 2.0 * c->object_half_xy;
    P->dt = c->cycle_time;
    P->mover_mass = c->mover_mass;
    P->mover_hx = c->mover_half[0];
    P->mover_hy = c->mover_half[1];
    P->mover_inertia = c->mover_mass * (lx * lx + ly * ly) / 12.0; /* box about its vertical axis */
    P->obj_mass = c->object_mass;
    P->obj_h = c->object_half_xy;
    P->obj_inertia = c->object_mass * (lo * lo + lo * lo) / 12.0;
    P->obj_damping = c->object_damping;
    P->mu = c->friction;
    P->gravity = c->gravity;
    P->k_rot = c->imp_k_rot;
    P->d_rot = 2.0 * sqrt(c->imp_k_rot * c->mover_mass); /* utils/impedance_control.py:46-47 */
    P->sol_B = 2.0 / (c->solimp[1] * c->solref[0]);
    P->sol_K = 1.0 / (c->solimp[1] * c->solimp[1] * c->solref[0] * c->solref[0] * c->solref[1] * c->solref[1]);
    P->imp_d0 = c->solimp[0];
    P->imp_dw = c->solimp[1];
    P->imp_width = c->solimp[2];
    P->imp_mid = c->solimp[3];
    P->imp_power = c->solimp[4];
    P->iterations = c->contact_iterations;
    P->warm_start = c->contact_warm_start != 0;
    {
        const double rm = sqrt(c->mover_half[0] * c->mover_half[0] + c->mover_half[1] * c->mover_half[1]);
        const double ro = sqrt(2.0 * c->object_half_xy * c->object_half_xy);
        P->contact_r2 = ((rm + ro) * 1.0001) * ((rm + ro) * 1.0001);
    }
    {
        /* a corner sits at distance sqrt(2) h from the centre; isotropic diagonal approximation of its inverse inertia */
        const double Ag = 1.0 / P->obj_mass + (2.0 * P->obj_h * P->obj_h) * (1.0 / P->obj_inertia) * 0.5;
        P->g_R = (1.0 - P->imp_d0) / P->imp_d0 * Ag;
        P->g_inv = 1.0 / (Ag + P->g_R);
        P->g_lim = P->mu * (P->obj_mass * P->gravity * 0.25);
    }
    P->o_inv_lin = 1.0 / (P->obj_mass + P->dt * P->obj_damping);
    P->o_inv_rot = 1.0 / (P->obj_inertia + P->dt * P->obj_damping);
    P->f_hxM = (float)P->mover_hx;
    P->f_hyM = (float)P->mover_hy;
    P->f_hO = (float)P->obj_h;
    P->f_imM = (float)(1.0 / P->mover_mass);
    P->f_iIM = (float)(1.0 / P->mover_inertia);
    P->f_imO = (float)(1.0 / P->obj_mass);
    P->f_iIO = (float)(1.0 / P->obj_inertia);
    P->f_mu = (float)P->mu;
    P->f_B = (float)P->sol_B;
    P->f_K = (float)P->sol_K;
    P->f_d0 = (float)P->imp_d0;
    P->f_dw = (float)P->imp_dw;
    P->f_width = (float)P->imp_width;
    P->f_mid = (float)P->imp_mid;
    P->f_power = (float)P->imp_power;
    P->f_gR = (float)P->g_R;
    P->f_ginv = (float)P->g_inv;
    P->f_glim = (float)P->g_lim;
    P->f_glim2 = (float)P->g_lim * (float)P->g_lim;
    P->f_D = (float)P->obj_damping;
    P->f_inv_width = (float)(1.0 / P->imp_width);
    P->f_inv_mid = (float)(1.0 / P->imp_mid);
    P->f_inv_1m_mid = (float)(1.0 / (1.0 - P->imp_mid));
}

/* per-environment solver state carried from substep to substep (warm start): contact forces (fn, ft) of up to two points,
 * ground-friction forces (fx, fy) of the four corners, and the number of contact points they belong to */
#define GPR_PUSH_WARM 13

typedef struct gpr_body2 {
    double x, y, c, s;  /* position, (cos yaw, sin yaw) */
    double vx, vy, w;   /* linear velocity, yaw rate */
} gpr_body2;

/* MuJoCo's impedance d(r): d0 -> dw over `width` of penetration, smooth power-law sigmoid (solimp).  float32. */
GPR_PHD float gpr_push_impedance(const gpr_push_params* P, float r) {
    const float x = GPR_FMUL(fabsf(r), P->f_inv_width); /* (multiplications by precomputed reciprocals: no division on the path) */
    if (x >= 1.0f) return P->f_dw;
    if (x <= 0.0f) return P->f_d0;
    float y;
    if (P->f_power == 2.0f) { /* the default: avoids pow() */
        if (x <= P->f_mid) {
            y = GPR_FMUL(GPR_FMUL(x, x), P->f_inv_mid);
        } else {
            const float u = GPR_FSUB(1.0f, x);
            y = GPR_FSUB(1.0f, GPR_FMUL(GPR_FMUL(u, u), P->f_inv_1m_mid));
        }
    } else {
        y = x; /* other powers fall back to a linear ramp (documented deviation; MuJoCo's default power is 2) */
    }
    return GPR_FFMA(y, GPR_FSUB(P->f_dw, P->f_d0), P->f_d0);
}

/* advance (c, s) by angle a = dt * w: second-order rotation, then ONE Newton step of 1/sqrt towards the unit circle
 * (n2 = c^2 + s^2 deviates from 1 by ~a^4/4 <= 1e-12 per substep; inv = 1.5 - 0.5 n2 leaves a deviation of the order of its
 * square, i.e. the iteration holds |(c, s)| = 1 to rounding — without the float64 sqrt and division a literal
 * renormalisation costs on the solve's critical path).  No transcendental function is called. */
GPR_PHD void gpr_push_rotate(double* c, double* s, double a) {
    double ca = 1.0 - 0.5 * (a * a);
    double sa = a - (a * a) * a * (1.0 / 6.0);
    double nc = (*c) * ca - (*s) * sa;
    double ns = (*s) * ca + (*c) * sa;
    double inv = 1.5 - 0.5 * (nc * nc + ns * ns);
    *c = nc * inv;
    *s = ns * inv;
}

/* yaw from (c, s) for small angles: asin series (|yaw| < 0.5 rad: rel. error < 2e-4; the impedance controller keeps the
 * mover's yaw within a few mrad) */
GPR_PHD double gpr_push_small_yaw(double c, double s) {
    (void)c;
    double s2 = s * s;
    return s * (1.0 + s2 * ((1.0 / 6.0) + s2 * (3.0 / 40.0)));
}

/* ---- contact manifold, float32, in the MOVER-CENTRED frame (the mover's centre is the origin; its axes are (cM, sM)) ---- */
typedef struct gpr_contact2 {
    float px, py; /* contact point relative to the mover's centre */
    float nx, ny; /* normal, from mover to object */
    float depth;  /* penetration depth >= 0 */
} gpr_contact2;

/* support radius of a box (half sizes hx, hy, axes (c,s)) along unit direction (nx, ny) */
GPR_PHD float gpr_box_radius(float c, float s, float hx, float hy, float nx, float ny) {
    const float a = fabsf(GPR_FFMA(nx, c, GPR_FMUL(ny, s)));
    const float b = fabsf(GPR_FFMA(ny, c, -GPR_FMUL(nx, s)));
    return GPR_FFMA(hx, a, GPR_FMUL(hy, b));
}

/* Planar box-box manifold (separating-axis test, then clip the incident edge against the reference face's side planes).
 * A = mover at the origin (axes (cA, sA), half sizes ahx, ahy), B = object at (dx, dy) (axes (cB, sB), half size bh).
 * Returns the number of contact points (0..2). */
GPR_PHD int gpr_box_box(float cA, float sA, float ahx, float ahy, float dx, float dy, float cB, float sB, float bh,
                        gpr_contact2 out[2]) {
    /* candidate axes: A's x, A's y, B's x, B's y */
    const float ax[4] = {cA, -sA, cB, -sB};
    const float ay[4] = {sA, cA, sB, cB};
    int best = -1;
    float best_sep = -3.0e38f, bnx = 0.0f, bny = 0.0f;
    for (int k = 0; k < 4; ++k) {
        float nx = ax[k], ny = ay[k];
        float proj = GPR_FFMA(dx, nx, GPR_FMUL(dy, ny));
        if (proj < 0.0f) { /* orient from A to B */
            nx = -nx;
            ny = -ny;
            proj = -proj;
        }
        const float ra = gpr_box_radius(cA, sA, ahx, ahy, nx, ny);
        const float rb = gpr_box_radius(cB, sB, bh, bh, nx, ny);
        const float sep = GPR_FSUB(proj, GPR_FADD(ra, rb));
        if (sep > 0.0f) return 0; /* separated */
        /* prefer the earlier axis on ties (a small bias keeps the choice stable for parallel faces) */
        if (sep > GPR_FADD(best_sep, 1e-7f)) {
            best_sep = sep;
            best = k;
            bnx = nx;
            bny = ny;
        }
    }
    /* reference box = owner of the best axis; incident box = the other one */
    const int ref_is_A = best < 2;
    const float Rx = ref_is_A ? 0.0f : dx, Ry = ref_is_A ? 0.0f : dy, Rc = ref_is_A ? cA : cB, Rs = ref_is_A ? sA : sB;
    const float Ix = ref_is_A ? dx : 0.0f, Iy = ref_is_A ? dy : 0.0f, Ic = ref_is_A ? cB : cA, Is = ref_is_A ? sB : sA;
    const float rhx = ref_is_A ? ahx : bh, rhy = ref_is_A ? ahy : bh;
    const float ihx = ref_is_A ? bh : ahx, ihy = ref_is_A ? bh : ahy;
    /* reference-face normal pointing from the reference box towards the incident box */
    const float rnx = ref_is_A ? bnx : -bnx, rny = ref_is_A ? bny : -bny;
    /* incident face: the face of I whose outward normal is most anti-parallel to (rnx, rny) */
    const float ix = Ic, iy = Is;   /* I's x axis */
    const float jx = -Is, jy = Ic;  /* I's y axis */
    const float dxn = GPR_FFMA(rnx, ix, GPR_FMUL(rny, iy));
    const float dyn = GPR_FFMA(rnx, jx, GPR_FMUL(rny, jy));
    float fcx, fcy, ftx, fty, fext; /* incident face centre, tangent direction, half extent */
    if (fabsf(dxn) >= fabsf(dyn)) {
        const float sg = dxn > 0.0f ? -ihx : ihx;
        fcx = GPR_FFMA(sg, ix, Ix);
        fcy = GPR_FFMA(sg, iy, Iy);
        ftx = jx;
        fty = jy;
        fext = ihy;
    } else {
        const float sg = dyn > 0.0f ? -ihy : ihy;
        fcx = GPR_FFMA(sg, jx, Ix);
        fcy = GPR_FFMA(sg, jy, Iy);
        ftx = ix;
        fty = iy;
        fext = ihx;
    }
    float v0x = GPR_FFMA(-fext, ftx, fcx), v0y = GPR_FFMA(-fext, fty, fcy);
    float v1x = GPR_FFMA(fext, ftx, fcx), v1y = GPR_FFMA(fext, fty, fcy);
    /* reference face: tangent (rtx, rty), half extent rext, offset along the normal */
    const float rtx = -rny, rty = rnx;
    const float rext = gpr_box_radius(Rc, Rs, rhx, rhy, rtx, rty);
    const float roff = gpr_box_radius(Rc, Rs, rhx, rhy, rnx, rny);
    /* clip the incident edge to |t| <= rext in the reference frame */
    float t0 = GPR_FFMA(GPR_FSUB(v0x, Rx), rtx, GPR_FMUL(GPR_FSUB(v0y, Ry), rty));
    float t1 = GPR_FFMA(GPR_FSUB(v1x, Rx), rtx, GPR_FMUL(GPR_FSUB(v1y, Ry), rty));
    if (t0 > t1) { /* order by t */
        float tmp;
        tmp = t0; t0 = t1; t1 = tmp;
        tmp = v0x; v0x = v1x; v1x = tmp;
        tmp = v0y; v0y = v1y; v1y = tmp;
    }
    if (t1 < -rext || t0 > rext) return 0;
    const float span = GPR_FSUB(t1, t0);
    if (t0 < -rext && span > 0.0f) {
        const float a = GPR_FDIV(GPR_FSUB(-rext, t0), span);
        v0x = GPR_FFMA(a, GPR_FSUB(v1x, v0x), v0x);
        v0y = GPR_FFMA(a, GPR_FSUB(v1y, v0y), v0y);
        t0 = -rext;
    }
    if (t1 > rext && span > 0.0f) {
        const float a = GPR_FDIV(GPR_FSUB(rext, t0), GPR_FSUB(t1, t0));
        v1x = GPR_FFMA(a, GPR_FSUB(v1x, v0x), v0x);
        v1y = GPR_FFMA(a, GPR_FSUB(v1y, v0y), v0y);
        t1 = rext;
    }
    int n = 0;
    const float ex[2] = {v0x, v1x}, ey[2] = {v0y, v1y};
    for (int k = 0; k < 2; ++k) {
        /* height above the reference centre */
        const float dn = GPR_FFMA(GPR_FSUB(ex[k], Rx), rnx, GPR_FMUL(GPR_FSUB(ey[k], Ry), rny));
        const float depth = GPR_FSUB(roff, dn);
        if (depth >= 0.0f) {
            /* contact point midway between the two surfaces; normal always from mover (A) to object (B) */
            const float hd = GPR_FMUL(0.5f, depth);
            out[n].px = GPR_FFMA(hd, rnx, ex[k]);
            out[n].py = GPR_FFMA(hd, rny, ey[k]);
            out[n].nx = bnx;
            out[n].ny = bny;
            out[n].depth = depth;
            ++n;
        }
    }
    if (n == 2 && fabsf(GPR_FSUB(t1, t0)) < 1e-7f) n = 1; /* degenerate edge: a single point */
    return n;
}

/* The substep when nothing constrains the bodies: the boxes are out of each other's reach and the object is at rest (no
 * ground friction to solve for).  EXACTLY the arithmetic gpr_push_substep performs in that case (it calls this function), so
 * a kernel that has established the condition may call it directly — the 40-cycle loop of most environments never leaves it. */
GPR_PHD int gpr_push_is_free(const gpr_push_params* P, const gpr_body2* M, const gpr_body2* O) {
    const double cdx = O->x - M->x, cdy = O->y - M->y;
    const int obj_moving = (O->vx != 0.0) || (O->vy != 0.0) || (O->w != 0.0);
    return (cdx * cdx + cdy * cdy > P->contact_r2) && !obj_moving;
}
GPR_PHD void gpr_push_integrate(const gpr_push_params* P, gpr_body2* M, gpr_body2* O, const double aM[3], const double fM[3],
                                const double fO[3], double* qax, double* qay) {
    const double dt = P->dt;
    const double imM = 1.0 / P->mover_mass, iIM = 1.0 / P->mover_inertia;
    /* ---- total accelerations; object damping implicit in velocity like MuJoCo's Euler: (M + dt D) a = f */
    const double axM = aM[0] + fM[0] * imM, ayM = aM[1] + fM[1] * imM, alM = aM[2] + fM[2] * iIM;
    const double axO = (-P->obj_damping * O->vx + fO[0]) * P->o_inv_lin;
    const double ayO = (-P->obj_damping * O->vy + fO[1]) * P->o_inv_lin;
    const double alO = (-P->obj_damping * O->w + fO[2]) * P->o_inv_rot;
    *qax = axM;
    *qay = ayM;
    /* ---- semi-implicit Euler */
    M->vx = M->vx + dt * axM;
    M->vy = M->vy + dt * ayM;
    M->w = M->w + dt * alM;
    M->x = M->x + dt * M->vx;
    M->y = M->y + dt * M->vy;
    if (M->w != 0.0) gpr_push_rotate(&M->c, &M->s, dt * M->w);
    O->vx = O->vx + dt * axO;
    O->vy = O->vy + dt * ayO;
    O->w = O->w + dt * alO;
    O->x = O->x + dt * O->vx;
    O->y = O->y + dt * O->vy;
    if (O->w != 0.0) gpr_push_rotate(&O->c, &O->s, dt * O->w);
}
/* (a free substep has no constraint forces: it also leaves the warm-start state all zero — callers that skip the general
 * function for free environments rely on that state being zero already, see pushing_contact_kernel) */
GPR_PHD void gpr_push_substep_free(const gpr_push_params* P, gpr_body2* M, gpr_body2* O, double ux, double uy, double* qax,
                                   double* qay) {
    const double iIM = 1.0 / P->mover_inertia;
    const double yaw = gpr_push_small_yaw(M->c, M->s);
    const double tau = P->k_rot * (0.0 - yaw) - P->d_rot * M->w; /* impedance_control.py:147 restricted to yaw */
    const double aM[3] = {ux, uy, tau * iIM};
    const double zero[3] = {0.0, 0.0, 0.0};
    gpr_push_integrate(P, M, O, aM, zero, zero, qax, qay);
}

/* 1 / sqrt(x) for the friction-disc projection, float32, x > 0 and normal: the classic integer first guess (relative error
 * < 3.5e-2) refined by three Newton steps y <- y (1.5 - 0.5 x y^2); each step squares the error, the last one leaves < 1e-9,
 * i.e. the result is within an ulp or two of the exact value.  Built from IEEE operations only, so the CPU oracle and the CUDA
 * kernels agree bit for bit — and a dozen dependent cheap operations instead of a correctly rounded sqrt AND a division,
 * which is what the solve's critical path is made of. */
GPR_PHD float gpr_rsqrtf(float x) {
    union {
        float f;
        uint32_t u;
    } cvt;
    cvt.f = x;
    cvt.u = 0x5f3759dfu - (cvt.u >> 1);
    float y = cvt.f;
    const float hx = GPR_FMUL(0.5f, x);
    y = GPR_FMUL(y, GPR_FFMA(-GPR_FMUL(hx, y), y, 1.5f));
    y = GPR_FMUL(y, GPR_FFMA(-GPR_FMUL(hx, y), y, 1.5f));
    y = GPR_FMUL(y, GPR_FFMA(-GPR_FMUL(hx, y), y, 1.5f));
    return y;
}

/* One constraint row of the solve: relative acceleration of the contact point (object minus mover) along a direction,
 * a = c0 + J . acc with acc = (a_Mx, a_My, alpha_M, a_Ox, a_Oy, alpha_O) the acceleration the constraint forces have
 * caused so far, J = (-d, -(rA x d), d, rB x d); a force change df on the row changes acc by df * W, W = M^-1 J^T. */
typedef struct gpr_row {
    float jx, jy, jA, jB;     /* direction d, rA x d, rB x d */
    float wMx, wMy, wMa;      /* W, mover part (already negated) */
    float wOx, wOy, wOa;      /* W, object part */
    float c0;                 /* J . a_smooth - a_ref */
    float R, inv;             /* regulariser, 1 / (A_ii + R) */
    float f;                  /* current force */
} gpr_row;

GPR_PHD void gpr_row_weights(const gpr_push_params* P, gpr_row* r) {
    r->wMx = -GPR_FMUL(r->jx, P->f_imM);
    r->wMy = -GPR_FMUL(r->jy, P->f_imM);
    r->wMa = -GPR_FMUL(r->jA, P->f_iIM);
    r->wOx = GPR_FMUL(r->jx, P->f_imO);
    r->wOy = GPR_FMUL(r->jy, P->f_imO);
    r->wOa = GPR_FMUL(r->jB, P->f_iIO);
}
GPR_PHD float gpr_row_acc(const gpr_row* r, const float acc[6]) {
    /* two independent chains (linear part, angular part), then one add */
    const float lin = GPR_FFMA(r->jy, GPR_FSUB(acc[4], acc[1]), GPR_FFMA(r->jx, GPR_FSUB(acc[3], acc[0]), r->c0));
    const float ang = GPR_FFMA(r->jB, acc[5], -GPR_FMUL(r->jA, acc[2]));
    return GPR_FADD(lin, ang);
}
GPR_PHD void gpr_row_apply(const gpr_row* r, float df, float acc[6]) {
    acc[0] = GPR_FFMA(df, r->wMx, acc[0]);
    acc[1] = GPR_FFMA(df, r->wMy, acc[1]);
    acc[2] = GPR_FFMA(df, r->wMa, acc[2]);
    acc[3] = GPR_FFMA(df, r->wOx, acc[3]);
    acc[4] = GPR_FFMA(df, r->wOy, acc[4]);
    acc[5] = GPR_FFMA(df, r->wOa, acc[5]);
}

/* One 1 ms substep.  (ux, uy): commanded mover acceleration (actuator force / mass).  On return the bodies are advanced
 * and (qax, qay) holds the mover's resulting x/y acceleration (MuJoCo's qacc, which the jerk-mode callback reads back,
 * pushing:431).  Returns the number of mover-object contact points that were active. */
GPR_PHD int gpr_push_substep(const gpr_push_params* P, gpr_body2* M, gpr_body2* O, double ux, double uy, double* qax,
                             double* qay, float* warm /* [GPR_PUSH_WARM] in/out, or NULL: cold start, nothing kept */) {
    if (gpr_push_is_free(P, M, O)) {
        gpr_push_substep_free(P, M, O, ux, uy, qax, qay);
        if (warm)
            for (int i = 0; i < GPR_PUSH_WARM; ++i) warm[i] = 0.0f;
        return 0;
    }
    const int use_warm = warm != 0 && P->warm_start;
    /* ---- smooth accelerations, float64 (no constraint forces; object damping enters as a passive force -D v) */
    const double iIM = 1.0 / P->mover_inertia;
    const double yaw = gpr_push_small_yaw(M->c, M->s);
    const double tau = P->k_rot * (0.0 - yaw) - P->d_rot * M->w; /* impedance_control.py:147 restricted to yaw */
    const double aM[3] = {ux, uy, tau * iIM};

    /* ---- float32 inputs of the constraint solve: relative pose (subtracted in float64, rounded once), orientations,
     *      velocities, smooth accelerations */
    const double cdx = O->x - M->x, cdy = O->y - M->y;
    const float dx = (float)cdx, dy = (float)cdy;
    const float cM = (float)M->c, sM = (float)M->s, cO = (float)O->c, sO = (float)O->s;
    const float vMx = (float)M->vx, vMy = (float)M->vy, wM = (float)M->w;
    const float vOx = (float)O->vx, vOy = (float)O->vy, wO = (float)O->w;
    const float a0[6] = {(float)aM[0], (float)aM[1], (float)aM[2], -GPR_FMUL(GPR_FMUL(P->f_D, vOx), P->f_imO),
                         -GPR_FMUL(GPR_FMUL(P->f_D, vOy), P->f_imO), -GPR_FMUL(GPR_FMUL(P->f_D, wO), P->f_iIO)};

    gpr_contact2 ct[2];
    /* boxes whose centres are farther apart than the sum of their circumradii are separated (the separating-axis test
     * would say so too): skip the manifold computation */
    const int nc = (cdx * cdx + cdy * cdy > P->contact_r2) ? 0 : gpr_box_box(cM, sM, P->f_hxM, P->f_hyM, dx, dy, cO, sO, P->f_hO, ct);

    float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f}; /* acceleration caused by the constraint forces so far */
    /* ---- contact rows: normal and tangent per point */
    gpr_row rn[2], rt[2];
    for (int k = 0; k < 2; ++k) {
        if (k >= nc) continue;
        const float nx = ct[k].nx, ny = ct[k].ny, tx = -ny, ty = nx;
        const float rAx = ct[k].px, rAy = ct[k].py, rBx = GPR_FSUB(ct[k].px, dx), rBy = GPR_FSUB(ct[k].py, dy);
        rn[k].jx = nx;
        rn[k].jy = ny;
        rn[k].jA = GPR_FFMA(rAx, ny, -GPR_FMUL(rAy, nx)); /* r x n */
        rn[k].jB = GPR_FFMA(rBx, ny, -GPR_FMUL(rBy, nx));
        rt[k].jx = tx;
        rt[k].jy = ty;
        rt[k].jA = GPR_FFMA(rAx, ty, -GPR_FMUL(rAy, tx));
        rt[k].jB = GPR_FFMA(rBx, ty, -GPR_FMUL(rBy, tx));
        const float mm = GPR_FADD(P->f_imM, P->f_imO);
        const float Ann = GPR_FFMA(GPR_FMUL(rn[k].jB, rn[k].jB), P->f_iIO, GPR_FFMA(GPR_FMUL(rn[k].jA, rn[k].jA), P->f_iIM, mm));
        const float Att = GPR_FFMA(GPR_FMUL(rt[k].jB, rt[k].jB), P->f_iIO, GPR_FFMA(GPR_FMUL(rt[k].jA, rt[k].jA), P->f_iIM, mm));
        const float d = gpr_push_impedance(P, ct[k].depth);
        const float R = GPR_FMUL(GPR_FDIV(GPR_FSUB(1.0f, d), d), Ann);
        rn[k].R = R;
        rt[k].R = R; /* impratio 1: the friction row shares the normal row's regulariser */
        /* 1 / (Ann + R) and 1 / (Att + R) from ONE division: 1 / (a b), times the other factor */
        const float den_n = GPR_FADD(Ann, R), den_t = GPR_FADD(Att, R);
        const float inv_nt = GPR_FDIV(1.0f, GPR_FMUL(den_n, den_t));
        rn[k].inv = GPR_FMUL(den_t, inv_nt);
        rt[k].inv = GPR_FMUL(den_n, inv_nt);
        /* warm start: the previous substep's forces when it had the same number of contact points */
        const int keep = use_warm && (int)warm[12] == nc;
        rn[k].f = keep ? warm[2 * k] : 0.0f;
        rt[k].f = keep ? warm[2 * k + 1] : 0.0f;
        /* relative velocity of the object w.r.t. the mover at the contact point */
        const float vrx = GPR_FSUB(GPR_FFMA(-wO, rBy, vOx), GPR_FFMA(-wM, rAy, vMx));
        const float vry = GPR_FSUB(GPR_FFMA(wO, rBx, vOy), GPR_FFMA(wM, rAx, vMy));
        const float vn = GPR_FFMA(vrx, nx, GPR_FMUL(vry, ny)), vt = GPR_FFMA(vrx, tx, GPR_FMUL(vry, ty));
        /* a_ref = -B v - K d r with r = -depth; c0 = J . a_smooth - a_ref */
        const float arn = GPR_FFMA(GPR_FMUL(P->f_K, d), ct[k].depth, -GPR_FMUL(P->f_B, vn));
        const float art = -GPR_FMUL(P->f_B, vt);
        /* J . a_smooth: (aO + alO x rB) - (aM + alM x rA) along the direction */
        const float sx = GPR_FSUB(GPR_FFMA(-a0[5], rBy, a0[3]), GPR_FFMA(-a0[2], rAy, a0[0]));
        const float sy = GPR_FSUB(GPR_FFMA(a0[5], rBx, a0[4]), GPR_FFMA(a0[2], rAx, a0[1]));
        rn[k].c0 = GPR_FSUB(GPR_FFMA(sx, nx, GPR_FMUL(sy, ny)), arn);
        rt[k].c0 = GPR_FSUB(GPR_FFMA(sx, tx, GPR_FMUL(sy, ty)), art);
        gpr_row_weights(P, &rn[k]);
        gpr_row_weights(P, &rt[k]);
    }
    /* ---- ground friction points: the four bottom corners of the object; a_ref = -B v at each */
    float gx[4], gy[4], gxI[4], gyI[4], gcx[4], gcy[4], gfx[4] = {0.0f, 0.0f, 0.0f, 0.0f}, gfy[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    {
        const float h = P->f_hO;
        const float lx[4] = {-h, -h, h, h}, ly[4] = {-h, h, h, -h};
        for (int g = 0; g < 4; ++g) {
            gx[g] = GPR_FFMA(cO, lx[g], -GPR_FMUL(sO, ly[g]));
            gy[g] = GPR_FFMA(sO, lx[g], GPR_FMUL(cO, ly[g]));
            /* smooth acceleration of the corner + B * its velocity */
            gcx[g] = GPR_FFMA(P->f_B, GPR_FFMA(-wO, gy[g], vOx), GPR_FFMA(-a0[5], gy[g], a0[3]));
            gcy[g] = GPR_FFMA(P->f_B, GPR_FFMA(wO, gx[g], vOy), GPR_FFMA(a0[5], gx[g], a0[4]));
            gxI[g] = GPR_FMUL(gx[g], P->f_iIO);
            gyI[g] = GPR_FMUL(gy[g], P->f_iIO);
            if (use_warm) {
                gfx[g] = warm[4 + 2 * g];
                gfy[g] = warm[5 + 2 * g];
            }
        }
    }
    if (use_warm) { /* the acceleration the warm forces cause: rows in solve order, then the corners */
        for (int k = 0; k < 2; ++k) {
            if (k >= nc) continue;
            gpr_row_apply(&rn[k], rn[k].f, acc);
            gpr_row_apply(&rt[k], rt[k].f, acc);
        }
        for (int g = 0; g < 4; ++g) {
            acc[3] = GPR_FFMA(gfx[g], P->f_imO, acc[3]);
            acc[4] = GPR_FFMA(gfy[g], P->f_imO, acc[4]);
            acc[5] = GPR_FFMA(gxI[g], gfy[g], GPR_FFMA(-gyI[g], gfx[g], acc[5]));
        }
    }
    /* ---- projected Gauss-Seidel in acceleration space: rows (n_1, t_1, n_2, t_2), then the four corners */
    for (int it = 0; it < P->iterations; ++it) {
        for (int k = 0; k < 2; ++k) {
            if (k >= nc) continue;
            /* normal row */
            float res = GPR_FFMA(rn[k].R, rn[k].f, gpr_row_acc(&rn[k], acc));
            float fn = GPR_FFMA(-res, rn[k].inv, rn[k].f);
            if (fn < 0.0f) fn = 0.0f;
            gpr_row_apply(&rn[k], GPR_FSUB(fn, rn[k].f), acc);
            rn[k].f = fn;
            /* tangent row (friction cone |ft| <= mu fn) */
            res = GPR_FFMA(rt[k].R, rt[k].f, gpr_row_acc(&rt[k], acc));
            float ft = GPR_FFMA(-res, rt[k].inv, rt[k].f);
            const float cone = GPR_FMUL(P->f_mu, fn);
            if (ft > cone) ft = cone;
            if (ft < -cone) ft = -cone;
            gpr_row_apply(&rt[k], GPR_FSUB(ft, rt[k].f), acc);
            rt[k].f = ft;
        }
        for (int g = 0; g < 4; ++g) {
            /* acceleration of the corner caused by the constraint forces so far */
            const float ax_ = GPR_FFMA(-acc[5], gy[g], acc[3]);
            const float ay_ = GPR_FFMA(acc[5], gx[g], acc[4]);
            float fx = GPR_FFMA(-GPR_FFMA(P->f_gR, gfx[g], GPR_FADD(ax_, gcx[g])), P->f_ginv, gfx[g]);
            float fy = GPR_FFMA(-GPR_FFMA(P->f_gR, gfy[g], GPR_FADD(ay_, gcy[g])), P->f_ginv, gfy[g]);
            const float mag2 = GPR_FFMA(fx, fx, GPR_FMUL(fy, fy));
            if (mag2 > P->f_glim2) { /* project onto the friction disc */
                const float sc = GPR_FMUL(P->f_glim, gpr_rsqrtf(mag2));
                fx = GPR_FMUL(fx, sc);
                fy = GPR_FMUL(fy, sc);
            }
            const float dfx = GPR_FSUB(fx, gfx[g]), dfy = GPR_FSUB(fy, gfy[g]);
            gfx[g] = fx;
            gfy[g] = fy;
            acc[3] = GPR_FFMA(dfx, P->f_imO, acc[3]);
            acc[4] = GPR_FFMA(dfy, P->f_imO, acc[4]);
            acc[5] = GPR_FFMA(gxI[g], dfy, GPR_FFMA(-gyI[g], dfx, acc[5]));
        }
    }
    if (warm) {
        for (int k = 0; k < 2; ++k) {
            warm[2 * k] = k < nc ? rn[k].f : 0.0f;
            warm[2 * k + 1] = k < nc ? rt[k].f : 0.0f;
        }
        for (int g = 0; g < 4; ++g) {
            warm[4 + 2 * g] = gfx[g];
            warm[5 + 2 * g] = gfy[g];
        }
        warm[12] = (float)nc;
    }
    /* constraint forces = mass * the acceleration they caused (the integrator adds them to the smooth forces in float64) */
    const double fM[3] = {(double)acc[0] * P->mover_mass, (double)acc[1] * P->mover_mass, (double)acc[2] * P->mover_inertia};
    const double fO[3] = {(double)acc[3] * P->obj_mass, (double)acc[4] * P->obj_mass, (double)acc[5] * P->obj_inertia};
    gpr_push_integrate(P, M, O, aM, fM, fO, qax, qay);
    return nc;
}

#endif /* GPR_PUSH_PHYSICS_H_ */
