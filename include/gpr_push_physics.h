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
}

typedef struct gpr_body2 {
    double x, y, c, s;  /* position, (cos yaw, sin yaw) */
    double vx, vy, w;   /* linear velocity, yaw rate */
} gpr_body2;

/* MuJoCo's impedance d(r): d0 -> dw over `width` of penetration, smooth power-law sigmoid (solimp). */
GPR_PHD double gpr_push_impedance(const gpr_push_params* P, double r) {
    double x = fabs(r) / P->imp_width;
    if (x >= 1.0) return P->imp_dw;
    if (x <= 0.0) return P->imp_d0;
    double y;
    if (P->imp_power == 2.0) { /* the default: avoids pow() */
        if (x <= P->imp_mid) {
            y = (x * x) / P->imp_mid;
        } else {
            double u = 1.0 - x;
            y = 1.0 - (u * u) / (1.0 - P->imp_mid);
        }
    } else {
        y = x; /* other powers fall back to a linear ramp (documented deviation; MuJoCo's default power is 2) */
    }
    return P->imp_d0 + y * (P->imp_dw - P->imp_d0);
}

/* advance (c, s) by angle a = dt * w: second-order rotation + renormalisation (no sin/cos call) */
GPR_PHD void gpr_push_rotate(double* c, double* s, double a) {
    double ca = 1.0 - 0.5 * (a * a);
    double sa = a - (a * a) * a * (1.0 / 6.0);
    double nc = (*c) * ca - (*s) * sa;
    double ns = (*s) * ca + (*c) * sa;
    double inv = 1.0 / sqrt(nc * nc + ns * ns);
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

typedef struct gpr_contact2 {
    double px, py; /* contact point (world) */
    double nx, ny; /* normal, from mover to object */
    double depth;  /* penetration depth >= 0 */
} gpr_contact2;

/* support radius of a box (half sizes hx, hy, axes (c,s)) along unit direction (nx, ny) */
GPR_PHD double gpr_box_radius(double c, double s, double hx, double hy, double nx, double ny) {
    return hx * fabs(nx * c + ny * s) + hy * fabs(-nx * s + ny * c);
}

/* Planar box-box manifold (separating-axis test, then clip the incident edge against the reference face's side planes).
 * A = mover (half sizes ahx, ahy), B = object (half size bh). Returns the number of contact points (0..2). */
GPR_PHD int gpr_box_box(const gpr_body2* A, double ahx, double ahy, const gpr_body2* Bd, double bh, gpr_contact2 out[2]) {
    const double dx = Bd->x - A->x, dy = Bd->y - A->y;
    /* candidate axes: A's x, A's y, B's x, B's y */
    double ax[4], ay[4];
    ax[0] = A->c;
    ay[0] = A->s;
    ax[1] = -A->s;
    ay[1] = A->c;
    ax[2] = Bd->c;
    ay[2] = Bd->s;
    ax[3] = -Bd->s;
    ay[3] = Bd->c;
    int best = -1;
    double best_sep = -1e300, bnx = 0.0, bny = 0.0;
    for (int k = 0; k < 4; ++k) {
        double nx = ax[k], ny = ay[k];
        double proj = dx * nx + dy * ny;
        if (proj < 0.0) { /* orient from A to B */
            nx = -nx;
            ny = -ny;
            proj = -proj;
        }
        double ra = gpr_box_radius(A->c, A->s, ahx, ahy, nx, ny);
        double rb = gpr_box_radius(Bd->c, Bd->s, bh, bh, nx, ny);
        double sep = proj - (ra + rb);
        if (sep > 0.0) return 0; /* separated */
        /* prefer the earlier axis on ties (small bias keeps the choice stable for parallel faces) */
        if (sep > best_sep + 1e-12) {
            best_sep = sep;
            best = k;
            bnx = nx;
            bny = ny;
        }
    }
    /* reference box = owner of the best axis; incident box = the other one */
    const int ref_is_A = best < 2;
    const gpr_body2* R = ref_is_A ? A : Bd;
    const gpr_body2* I = ref_is_A ? Bd : A;
    const double rhx = ref_is_A ? ahx : bh, rhy = ref_is_A ? ahy : bh;
    const double ihx = ref_is_A ? bh : ahx, ihy = ref_is_A ? bh : ahy;
    /* reference-face normal pointing from the reference box towards the incident box */
    const double rnx = ref_is_A ? bnx : -bnx, rny = ref_is_A ? bny : -bny;
    /* incident face: the face of I whose outward normal is most anti-parallel to (rnx, rny) */
    const double ix = I->c, iy = I->s;      /* I's x axis */
    const double jx = -I->s, jy = I->c;     /* I's y axis */
    const double dxn = rnx * ix + rny * iy; /* cos between rn and I's x axis */
    const double dyn = rnx * jx + rny * jy;
    double fcx, fcy, ftx, fty, fext; /* incident face centre, tangent direction, half extent */
    if (fabs(dxn) >= fabs(dyn)) {
        const double sgn = dxn > 0.0 ? -1.0 : 1.0;
        fcx = I->x + sgn * ihx * ix;
        fcy = I->y + sgn * ihx * iy;
        ftx = jx;
        fty = jy;
        fext = ihy;
    } else {
        const double sgn = dyn > 0.0 ? -1.0 : 1.0;
        fcx = I->x + sgn * ihy * jx;
        fcy = I->y + sgn * ihy * jy;
        ftx = ix;
        fty = iy;
        fext = ihx;
    }
    double v0x = fcx - fext * ftx, v0y = fcy - fext * fty;
    double v1x = fcx + fext * ftx, v1y = fcy + fext * fty;
    /* reference face: tangent (rtx, rty), half extent rext, offset along the normal */
    const double rtx = -rny, rty = rnx;
    const double rext = gpr_box_radius(R->c, R->s, rhx, rhy, rtx, rty);
    const double roff = gpr_box_radius(R->c, R->s, rhx, rhy, rnx, rny);
    /* clip the incident edge to |t| <= rext in the reference frame */
    double t0 = (v0x - R->x) * rtx + (v0y - R->y) * rty;
    double t1 = (v1x - R->x) * rtx + (v1y - R->y) * rty;
    if (t0 > t1) { /* order by t */
        double tmp;
        tmp = t0; t0 = t1; t1 = tmp;
        tmp = v0x; v0x = v1x; v1x = tmp;
        tmp = v0y; v0y = v1y; v1y = tmp;
    }
    if (t1 < -rext || t0 > rext) return 0;
    const double span = t1 - t0;
    if (t0 < -rext && span > 0.0) {
        double a = (-rext - t0) / span;
        v0x = v0x + a * (v1x - v0x);
        v0y = v0y + a * (v1y - v0y);
        t0 = -rext;
    }
    if (t1 > rext && span > 0.0) {
        double a = (rext - t0) / (t1 - t0);
        v1x = v0x + a * (v1x - v0x);
        v1y = v0y + a * (v1y - v0y);
        t1 = rext;
    }
    int n = 0;
    const double ex[2] = {v0x, v1x}, ey[2] = {v0y, v1y};
    for (int k = 0; k < 2; ++k) {
        const double dn = (ex[k] - R->x) * rnx + (ey[k] - R->y) * rny; /* height above the reference centre */
        const double depth = roff - dn;
        if (depth >= 0.0) {
            /* contact point midway between the two surfaces; normal always from mover (A) to object (B) */
            out[n].px = ex[k] + 0.5 * depth * rnx;
            out[n].py = ey[k] + 0.5 * depth * rny;
            out[n].nx = bnx;
            out[n].ny = bny;
            out[n].depth = depth;
            ++n;
        }
    }
    if (n == 2 && fabs(t1 - t0) < 1e-9) n = 1; /* degenerate edge: a single point */
    return n;
}

/* One 1 ms substep.  (ux, uy): commanded mover acceleration (actuator force / mass).  On return the bodies are advanced
 * and (qax, qay) holds the mover's resulting x/y acceleration (MuJoCo's qacc, which the jerk-mode callback reads back,
 * pushing:431).  Returns the number of mover-object contact points that were active. */
GPR_PHD int gpr_push_substep(const gpr_push_params* P, gpr_body2* M, gpr_body2* O, double ux, double uy, double* qax,
                             double* qay) {
    const double dt = P->dt;
    const double imM = 1.0 / P->mover_mass, iIM = 1.0 / P->mover_inertia;
    const double imO = 1.0 / P->obj_mass, iIO = 1.0 / P->obj_inertia;
    /* ---- smooth accelerations (no constraint forces; object damping enters as a passive force -D v) */
    const double yaw = gpr_push_small_yaw(M->c, M->s);
    const double tau = P->k_rot * (0.0 - yaw) - P->d_rot * M->w; /* impedance_control.py:147 restricted to yaw */
    double aM[3] = {ux, uy, tau * iIM};
    double aO[3] = {-P->obj_damping * O->vx * imO, -P->obj_damping * O->vy * imO, -P->obj_damping * O->w * iIO};

    gpr_contact2 ct[2];
    /* boxes whose centres are farther apart than the sum of their circumradii are separated (the separating-axis test
     * below would say so too): skip the manifold computation */
    const double cdx = O->x - M->x, cdy = O->y - M->y;
    const int nc = (cdx * cdx + cdy * cdy > P->contact_r2) ? 0 : gpr_box_box(M, P->mover_hx, P->mover_hy, O, P->obj_h, ct);
    const int obj_moving = (O->vx != 0.0) || (O->vy != 0.0) || (O->w != 0.0);
    double fO[3] = {0.0, 0.0, 0.0}, fM[3] = {0.0, 0.0, 0.0}; /* accumulated constraint wrenches */

    if (nc > 0 || obj_moving) {
        /* ---- constraint rows */
        double cn[2] = {0.0, 0.0}, ctg[2] = {0.0, 0.0}; /* contact normal / tangent forces */
        double gfx[4] = {0.0, 0.0, 0.0, 0.0}, gfy[4] = {0.0, 0.0, 0.0, 0.0}; /* ground friction at the corners */
        double rAx[2], rAy[2], rBx[2], rBy[2], Rn[2], invN[2], invT[2], arn[2], art[2];
        for (int k = 0; k < nc; ++k) {
            rAx[k] = ct[k].px - M->x;
            rAy[k] = ct[k].py - M->y;
            rBx[k] = ct[k].px - O->x;
            rBy[k] = ct[k].py - O->y;
            const double nx = ct[k].nx, ny = ct[k].ny, tx = -ny, ty = nx;
            const double rAn = rAx[k] * ny - rAy[k] * nx, rBn = rBx[k] * ny - rBy[k] * nx; /* r x n */
            const double rAt = rAx[k] * ty - rAy[k] * tx, rBt = rBx[k] * ty - rBy[k] * tx;
            const double Ann = imM + imO + rAn * rAn * iIM + rBn * rBn * iIO;
            const double Att = imM + imO + rAt * rAt * iIM + rBt * rBt * iIO;
            const double d = gpr_push_impedance(P, ct[k].depth);
            Rn[k] = (1.0 - d) / d * Ann;
            invN[k] = 1.0 / (Ann + Rn[k]); /* (the sweeps below multiply instead of dividing) */
            invT[k] = 1.0 / (Att + Rn[k]);
            /* relative velocity of the object w.r.t. the mover at the contact point */
            const double vrx = (O->vx - O->w * rBy[k]) - (M->vx - M->w * rAy[k]);
            const double vry = (O->vy + O->w * rBx[k]) - (M->vy + M->w * rAx[k]);
            arn[k] = -P->sol_B * (vrx * nx + vry * ny) + P->sol_K * d * ct[k].depth; /* r = -depth */
            art[k] = -P->sol_B * (vrx * tx + vry * ty);
        }
        /* ground friction points: the four bottom corners of the object; a_ref = -B v at each */
        double gx[4], gy[4], bvx[4], bvy[4];
        const double lim = P->g_lim, lim2 = lim * lim;
        {
            const double h = P->obj_h;
            const double cx[4] = {-h, -h, h, h}, cy[4] = {-h, h, h, -h};
            for (int g = 0; g < 4; ++g) {
                gx[g] = O->c * cx[g] - O->s * cy[g];
                gy[g] = O->s * cx[g] + O->c * cy[g];
                bvx[g] = P->sol_B * (O->vx - O->w * gy[g]);
                bvy[g] = P->sol_B * (O->vy + O->w * gx[g]);
            }
        }
        /* ---- projected Gauss-Seidel in acceleration space */
        for (int it = 0; it < P->iterations; ++it) {
            for (int k = 0; k < nc; ++k) {
                const double nx = ct[k].nx, ny = ct[k].ny, tx = -ny, ty = nx;
                /* current relative acceleration at the contact (object minus mover) */
                double alO = aO[2] + fO[2] * iIO, alM = aM[2] + fM[2] * iIM;
                double arx = (aO[0] + (fO[0] * imO) - alO * rBy[k]) - (aM[0] + (fM[0] * imM) - alM * rAy[k]);
                double ary = (aO[1] + (fO[1] * imO) + alO * rBx[k]) - (aM[1] + (fM[1] * imM) + alM * rAx[k]);
                /* normal row */
                double res = (arx * nx + ary * ny) - arn[k] + Rn[k] * cn[k];
                double fn = cn[k] - res * invN[k];
                if (fn < 0.0) fn = 0.0;
                double df = fn - cn[k];
                cn[k] = fn;
                fO[0] += df * nx;
                fO[1] += df * ny;
                fO[2] += df * (rBx[k] * ny - rBy[k] * nx);
                fM[0] -= df * nx;
                fM[1] -= df * ny;
                fM[2] -= df * (rAx[k] * ny - rAy[k] * nx);
                /* tangent row (friction cone |ft| <= mu fn) */
                alO = aO[2] + fO[2] * iIO;
                alM = aM[2] + fM[2] * iIM;
                arx = (aO[0] + (fO[0] * imO) - alO * rBy[k]) - (aM[0] + (fM[0] * imM) - alM * rAy[k]);
                ary = (aO[1] + (fO[1] * imO) + alO * rBx[k]) - (aM[1] + (fM[1] * imM) + alM * rAx[k]);
                res = (arx * tx + ary * ty) - art[k] + Rn[k] * ctg[k];
                double ft = ctg[k] - res * invT[k];
                const double cone = P->mu * cn[k];
                if (ft > cone) ft = cone;
                if (ft < -cone) ft = -cone;
                df = ft - ctg[k];
                ctg[k] = ft;
                fO[0] += df * tx;
                fO[1] += df * ty;
                fO[2] += df * (rBx[k] * ty - rBy[k] * tx);
                fM[0] -= df * tx;
                fM[1] -= df * ty;
                fM[2] -= df * (rAx[k] * ty - rAy[k] * tx);
            }
            for (int g = 0; g < 4; ++g) {
                /* acceleration of the corner */
                const double alpha = aO[2] + fO[2] * iIO;
                const double ax_ = (aO[0] + fO[0] * imO) - alpha * gy[g];
                const double ay_ = (aO[1] + fO[1] * imO) + alpha * gx[g];
                double fx = gfx[g] - (ax_ + bvx[g] + P->g_R * gfx[g]) * P->g_inv;
                double fy = gfy[g] - (ay_ + bvy[g] + P->g_R * gfy[g]) * P->g_inv;
                const double mag2 = fx * fx + fy * fy;
                if (mag2 > lim2) { /* project onto the friction disc */
                    const double sc = lim / sqrt(mag2);
                    fx = fx * sc;
                    fy = fy * sc;
                }
                const double dfx = fx - gfx[g], dfy = fy - gfy[g];
                gfx[g] = fx;
                gfy[g] = fy;
                fO[0] += dfx;
                fO[1] += dfy;
                fO[2] += gx[g] * dfy - gy[g] * dfx;
            }
        }
    }
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
    return nc;
}

#endif /* GPR_PUSH_PHYSICS_H_ */
