// gpr_device.cuh — device building blocks of the step path (sm_100a).
//
// Thread mapping: one LANE per (environment, mover); the movers of one environment form a lane GROUP of G = 2^k lanes
// (G >= num_movers, G <= 32) inside one warp, so everything per-environment (collision flags, goal counts, the
// break-on-collision of basic_envs.py:1904) is a ballot/shuffle over the group and never touches memory.
//
// Arithmetic: IEEE float64 with explicitly rounded intrinsics (__dadd_rn, __dmul_rn, ...).  nvcc never contracts those
// into FMAs, so every expression below evaluates exactly like the reference's NumPy float64 expression it cites and the
// collision / termination flags are bit-identical to the float64 oracle (DESIGN.md "Precision").
//
// File:line citations are into /root/reference/gymnasium_planar_robotics/ :
//   basic = envs/basic_envs.py, plan = envs/planning/benchmark_planning_env.py, geom = utils/geometry_2D_utils.py,
//   rot = utils/rotations_utils.py
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gpr.h"
#include "../../include/gpr_rng.h"

// Inlining policy of the cold (rarely executed) helpers, overridable per helper for tuning experiments:
//   -DGPR_INL_WALL=0 / GPR_INL_NORMAL=0 / GPR_INL_PAIR=0 / GPR_INL_CONFIRM=0 move a helper out of line.  Measured on B200
//   (planning4, same box, +-0.1%): all inline 0.312 ms/step, all out of line 0.326, mixed 0.320-0.341 -> inline.
#ifndef GPR_INL_WALL
#define GPR_INL_WALL 1
#endif
#ifndef GPR_INL_NORMAL
#define GPR_INL_NORMAL 1
#endif
#ifndef GPR_INL_PAIR
#define GPR_INL_PAIR 1
#endif
#ifndef GPR_INL_CONFIRM
#define GPR_INL_CONFIRM 1
#endif
#define GPR_COLD_1 __device__ __forceinline__
#define GPR_COLD_0 static __device__ __noinline__
#define GPR_COLD_CAT(x) GPR_COLD_##x
#define GPR_COLD(flag) GPR_COLD_CAT(flag)

// GPR_DEBUG_BOUNDS (csrc/Makefile target `debug` -> libgpr_b200_dbg.so): compute-sanitizer is closed on this GPU pool, so a
// debug build range-checks every index the kernels derive for global memory — env / mover indices under ragged batch
// sizes, work-list slots and entries of the streamed auto-reset and of the pushing contact queue, output rows — into a
// device error counter that gpr_debug_errors() reads (tests/test_gpu_debug_bounds.py).  A violating access is SKIPPED in
// the debug build where that is cheap, and always counted.  The release build compiles the checks away.
#ifdef GPR_DEBUG_BOUNDS
#define GPR_CHECK(args, cond, slot)                                   \
    do {                                                              \
        if (!(cond)) atomicAdd((args).debug_errors + (slot), 1u);     \
    } while (0)
#else
#define GPR_CHECK(args, cond, slot) ((void)0)
#endif
// error slots
enum : int {
    DBG_LANE_INDEX = 0,    // a lane's (env, mover) index out of [0, B*N)
    DBG_LIST_SLOT = 1,     // work-list / queue slot out of [0, B)
    DBG_LIST_ENTRY = 2,    // work-list / queue entry names an env out of [0, B) (or a cycle out of range)
    DBG_LIST_CLAIM = 3,    // consumer claimed past the published count
    DBG_OUTPUT_ROW = 4,    // output row out of [0, B)
    DBG_SHARED_INDEX = 5,  // shared-memory staging index out of range
    DBG_NUM_SLOTS = 8
};

namespace gpr {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kMaxTiles = GPR_MAX_TILES_1D;

// ---- exactly rounded float64 primitives ---------------------------------------------------------------------------
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
static __device__ __noinline__ double ddiv_cold(double a, double b) { return __ddiv_rn(a, b); }
// RN(x / y) from z = RN(1 / y), y > 0, without the division sequence (a third of its instructions where one reciprocal serves
// several quotients or y is a constant): q0 = RN(x z) is within 1.5 ulp of x / y; one FMA correction makes it faithful, and for a
// faithful q the residual x - q y is exact and RN(q + r z) is the correctly rounded quotient (Markstein, "Computation of
// elementary functions on the IBM RISC System/6000 processor", 1990, Theorem on FMA-based division).  Zero, tiny x (the
// residual could underflow) and huge y go to the plain division.
__device__ __forceinline__ double ddiv_rcp(double x, double y, double z) {
    if (!(fabs(x) > 1e-200 && y < 1e150)) return ddiv_cold(x, y);
    const double q0 = __dmul_rn(x, z);
    const double r0 = __fma_rn(-q0, y, x);
    const double q1 = __fma_rn(r0, z, q0);
    const double r1 = __fma_rn(-q1, y, x);
    return __fma_rn(r1, z, q1);
}
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }

// sqrt(s) <= t  /  sqrt(s) < t  /  sqrt(s) >= t, decided EXACTLY as the correctly rounded square root would, but the
// square root itself is only taken inside a relative band of 1e-14 around t^2 (rounding of t*t and of sqrt are < 3e-16).
__device__ __forceinline__ bool sqrt_le(double s, double t) {
    const double t2 = dmul(t, t);
    if (s <= dmul(t2, 1.0 - 1e-14)) return true;
    if (s >= dmul(t2, 1.0 + 1e-14)) return false;
    return dsqrt(s) <= t;
}
__device__ __forceinline__ bool sqrt_lt(double s, double t) {
    const double t2 = dmul(t, t);
    if (s <= dmul(t2, 1.0 - 1e-14)) return true;
    if (s >= dmul(t2, 1.0 + 1e-14)) return false;
    return dsqrt(s) < t;
}

// ---- tile layout tables (shared memory) ----------------------------------------------------------------------------
// cell code bits (built on the host from layout_tiles, basic:194-221):
enum : uint32_t {
    CELL_T = 1u << 0,   // tile present at (i,j)
    CELL_W = 1u << 1,   // (i-1,j)
    CELL_E = 1u << 2,   // (i+1,j)
    CELL_S = 1u << 3,   // (i,j-1)
    CELL_N = 1u << 4,   // (i,j+1)
    CELL_SW = 1u << 5,  // (i-1,j-1)
    CELL_NW = 1u << 6,  // (i-1,j+1)
    CELL_SE = 1u << 7,  // (i+1,j-1)
    CELL_NE = 1u << 8,  // (i+1,j+1)
    CELL_3X3 = 1u << 9, // centre of a fully populated 3x3 block (basic:205-207)
    // inner corners: this cell and both side neighbours exist, the diagonal cell is inside the grid and missing
    // (the four "2x2 with one missing corner" patterns, basic:208-219)
    CELL_IC_SE = 1u << 10,  // missing (i+1,j-1)  "bl"
    CELL_IC_NE = 1u << 11,  // missing (i+1,j+1)  "br"
    CELL_IC_SW = 1u << 12,  // missing (i-1,j-1)  "tl"
    CELL_IC_NW = 1u << 13,  // missing (i-1,j+1)  "tr"
};

struct Tables {
    double xlo[kMaxTiles], xhi[kMaxTiles];  // tile_cx[i] -+ half  (basic:519-523)
    double ylo[kMaxTiles], yhi[kMaxTiles];
    uint16_t cell[kMaxTiles * kMaxTiles];   // [i * ny + j]
    float xlof[kMaxTiles], ylof[kMaxTiles]; // float copies of xlo / ylo (float32 screens only, never decisive)
};

struct LayoutArgs {
    int nx, ny;
    double hx, hy;
    double inv_wx, inv_wy;  // 1 / (2*half): only used to GUESS the cell, never to decide
    const double* cx;       // [nx] device
    const double* cy;       // [ny] device
    const uint16_t* cell;   // [nx*ny] device
};

__device__ __forceinline__ void load_tables(Tables& tb, const LayoutArgs& L) {
    for (int i = threadIdx.x; i < L.nx; i += blockDim.x) {
        const double c = L.cx[i];
        tb.xlo[i] = dsub(c, L.hx);
        tb.xhi[i] = dadd(c, L.hx);
        tb.xlof[i] = (float)dsub(c, L.hx);
    }
    for (int j = threadIdx.x; j < L.ny; j += blockDim.x) {
        const double c = L.cy[j];
        tb.ylo[j] = dsub(c, L.hy);
        tb.yhi[j] = dadd(c, L.hy);
        tb.ylof[j] = (float)dsub(c, L.hy);
    }
    for (int k = threadIdx.x; k < L.nx * L.ny; k += blockDim.x) tb.cell[k] = L.cell[k];
}

// unsafe-side bits of one tested point w.r.t. its cell: bit0 min_x, bit1 max_x, bit2 min_y, bit3 max_y (basic:545-572)
// -> is the point forgiven by the neighbouring tiles? (basic:574-657, the integer sum collapses to this)
__device__ __forceinline__ bool sides_ok(uint32_t u, uint32_t code) {
    uint32_t req = (u << 1) & (CELL_W | CELL_E | CELL_S | CELL_N);
    req |= ((u & 5u) == 5u) ? CELL_SW : 0u;
    req |= ((u & 9u) == 9u) ? CELL_NW : 0u;
    req |= ((u & 6u) == 6u) ? CELL_SE : 0u;
    req |= ((u & 10u) == 10u) ? CELL_NE : 0u;
    const bool both = ((u & 3u) == 3u) | ((u & 12u) == 12u);  // reference asserts (basic:650); treated as invalid
    return (code & CELL_T) && ((req & ~code) == 0u) && !both;
}

// ---- 2-D geometry (geom:9-138) --------------------------------------------------------------------------------------
struct Rect {
    double x[4], y[4];  // vertices (-sx,-sy), (-sx,sy), (sx,sy), (sx,-sy) in the base frame
};

__device__ __forceinline__ double orient(double ax, double ay, double bx, double by, double cx, double cy) {
    // det([[ax,bx,cx],[ay,by,cy],[1,1,1]])  (geom:47-60)
    return dsub(dmul(dsub(bx, ax), dsub(cy, ay)), dmul(dsub(by, ay), dsub(cx, ax)));
}
__device__ __forceinline__ bool pts_equal(double ax, double ay, double bx, double by) {
    return (fabs(dsub(ax, bx)) < 1e-7) && (fabs(dsub(ay, by)) < 1e-7);  // geom:30-35
}
__device__ __forceinline__ bool axis_separated(double p1, double p2, double q1, double q2) {
    // geom:37-44 for one coordinate
    const double min_p = fmin(p1, p2), max_p = fmax(p1, p2), min_q = fmin(q1, q2), max_q = fmax(q1, q2);
    const bool a = (max_p < min_q) && !(fabs(dsub(max_p, min_q)) < 1e-7);
    const bool b = (max_q < min_p) && !(fabs(dsub(max_q, min_p)) < 1e-7);
    return a || b;
}
__device__ __forceinline__ bool segments_intersect(double p1x, double p1y, double p2x, double p2y, double q1x,
                                                   double q1y, double q2x, double q2y) {
    if (pts_equal(p1x, p1y, q1x, q1y) || pts_equal(p1x, p1y, q2x, q2y) || pts_equal(p2x, p2y, q1x, q1y) ||
        pts_equal(p2x, p2y, q2x, q2y))
        return true;  // geom:68
    if (axis_separated(p1x, p2x, q1x, q2x) || axis_separated(p1y, p2y, q1y, q2y)) return false;  // geom:67
    const double pa = dmul(orient(p1x, p1y, p2x, p2y, q1x, q1y), orient(p1x, p1y, p2x, p2y, q2x, q2y));
    const double pb = dmul(orient(q1x, q1y, q2x, q2y, p1x, p1y), orient(q1x, q1y, q2x, q2y, p2x, p2y));
    return ((pa <= 0.0) || (fabs(pa) < 1e-7)) && ((pb <= 0.0) || (fabs(pb) < 1e-7));  // geom:62-64
}
static __device__ __noinline__ bool rects_intersect(const Rect& a, const Rect& b) {
    bool any = false;  // geom:132-138: 4 x 4 edge pairs
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const int i2 = (i + 1) & 3;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int j2 = (j + 1) & 3;
            any |= segments_intersect(a.x[i], a.y[i], a.x[i2], a.y[i2], b.x[j], b.y[j], b.x[j2], b.y[j2]);
        }
    }
    return any;
}

// rot:414-461 (float32 normalisation) + rot:248-274 (quat2mat) + geom:91-102, planar part
__device__ __forceinline__ void rect_vertices(double px, double py, double qw, double qx, double qy, double qz,
                                              double sx, double sy, Rect& r) {
    float f0 = (float)qw, f1 = (float)qx, f2 = (float)qy, f3 = (float)qz;
    float s = __fmul_rn(f0, f0);
    s = __fadd_rn(s, __fmul_rn(f1, f1));
    s = __fadd_rn(s, __fmul_rn(f2, f2));
    s = __fadd_rn(s, __fmul_rn(f3, f3));
    const float len = __fsqrt_rn(s);
    f0 = __fdiv_rn(f0, len);
    f1 = __fdiv_rn(f1, len);
    f2 = __fdiv_rn(f2, len);
    f3 = __fdiv_rn(f3, len);
    const double w = (double)f0, x = (double)f1, y = (double)f2, z = (double)f3;
    const double Nq = dadd(dadd(dadd(dmul(w, w), dmul(x, x)), dmul(y, y)), dmul(z, z));
    double r00 = 1.0, r01 = 0.0, r10 = 0.0, r11 = 1.0;
    if (Nq > 2.220446049250313e-16) {
        const double sc = ddiv(2.0, Nq);
        const double X = dmul(x, sc), Y = dmul(y, sc), Z = dmul(z, sc);
        const double wZ = dmul(w, Z), xX = dmul(x, X), xY = dmul(x, Y), yY = dmul(y, Y), zZ = dmul(z, Z);
        r00 = dsub(1.0, dadd(yY, zZ));
        r01 = dsub(xY, wZ);
        r10 = dadd(xY, wZ);
        r11 = dsub(1.0, dadd(xX, zZ));
    }
    const double lx[4] = {-sx, -sx, sx, sx};
    const double ly[4] = {-sy, sy, sy, -sy};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        r.x[k] = dadd(dadd(dmul(r00, lx[k]), dmul(r01, ly[k])), px);
        r.y[k] = dadd(dadd(dmul(r10, lx[k]), dmul(r11, ly[k])), py);
    }
}
// identity orientation (planning, no noise): the general formula reduces to +-s + p exactly
__device__ __forceinline__ void rect_vertices_axis(double px, double py, double sx, double sy, Rect& r) {
    r.x[0] = dadd(-sx, px);
    r.x[1] = r.x[0];
    r.x[2] = dadd(sx, px);
    r.x[3] = r.x[2];
    r.y[0] = dadd(-sy, py);
    r.y[1] = dadd(sy, py);
    r.y[2] = r.y[1];
    r.y[3] = r.y[0];
}

// ---- wall / tile-layout check (basic:459-788, one mover) -------------------------------------------------------------
// Containing cells are found exactly: the guessed index and its two neighbours are each tested with the reference's
// inclusive comparisons against the reference's float64 bounds (basic:507-512); the guess never decides anything.
// (out of line on purpose: it is the rare exact fallback of the float32 screens, called from many places)
template <bool BOX>
GPR_COLD(GPR_INL_WALL) bool wall_valid(const Tables& tb, const LayoutArgs& L, double x, double y, double cs0,
                                               const Rect& rect) {
    int gi = __double2int_rd(dmul(x, L.inv_wx));
    int gj = __double2int_rd(dmul(y, L.inv_wy));
    gi = min(max(gi, 0), L.nx - 1);
    gj = min(max(gj, 0), L.ny - 1);
    uint32_t mx = 0, my = 0;  // bit d+1 set <=> cell index g+d contains the coordinate
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
        const int i = gi + d, j = gj + d;
        if (i >= 0 && i < L.nx && tb.xlo[i] <= x && x <= tb.xhi[i]) mx |= 1u << (d + 1);
        if (j >= 0 && j < L.ny && tb.ylo[j] <= y && y <= tb.yhi[j]) my |= 1u << (d + 1);
    }
    if (mx == 0u || my == 0u) return false;  // not above any cell: reference asserts (basic:514-517) -> wall collision
    bool complete = false, all_rows = true;
#pragma unroll 1
    for (int di = 0; di < 3; ++di) {
        if (!((mx >> di) & 1u)) continue;
        const int i = gi + di - 1;
        const double xl = tb.xlo[i], xh = tb.xhi[i];
#pragma unroll 1
        for (int dj = 0; dj < 3; ++dj) {
            if (!((my >> dj) & 1u)) continue;
            const int j = gj + dj - 1;
            const double yl = tb.ylo[j], yh = tb.yhi[j];
            const uint32_t code = tb.cell[i * L.ny + j];
            complete |= (code & CELL_3X3) != 0u;  // basic:527-538
            bool row;
            if (!BOX) {
                // circle == its axis-aligned bounding square (basic:545-558), strict comparisons
                uint32_t u = 0;
                u |= (xl < dsub(x, cs0)) ? 0u : 1u;
                u |= (dadd(x, cs0) < xh) ? 0u : 2u;
                u |= (yl < dsub(y, cs0)) ? 0u : 4u;
                u |= (dadd(y, cs0) < yh) ? 0u : 8u;
                row = sides_ok(u, code);
            } else {
                row = true;  // basic:559-572 and 655: all four vertices
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t u = 0;
                    u |= (xl < rect.x[k]) ? 0u : 1u;
                    u |= (rect.x[k] < xh) ? 0u : 2u;
                    u |= (yl < rect.y[k]) ? 0u : 4u;
                    u |= (rect.y[k] < yh) ? 0u : 8u;
                    row &= sides_ok(u, code);
                }
                if (row && (code & (CELL_IC_SE | CELL_IC_NE | CELL_IC_SW | CELL_IC_NW))) {
                    // basic:657-783: edge-intersection with the missing tile of an inner corner
#pragma unroll 1
                    for (int p = 0; p < 4 && row; ++p) {
                        if (!(code & (CELL_IC_SE << p))) continue;
                        const int mi = i + ((p < 2) ? 1 : -1);
                        const int mj = j + ((p & 1) ? 1 : -1);
                        Rect t;
                        t.x[0] = tb.xlo[mi];
                        t.x[1] = t.x[0];
                        t.x[2] = tb.xhi[mi];
                        t.x[3] = t.x[2];
                        t.y[0] = tb.ylo[mj];
                        t.y[1] = tb.yhi[mj];
                        t.y[2] = t.y[1];
                        t.y[3] = t.y[0];
                        if (rects_intersect(rect, t)) row = false;
                    }
                }
            }
            all_rows &= row;
        }
    }
    return complete || all_rows;  // basic:538, 785-786
}

// ---- lane-group helpers ----------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ unsigned group_mask(unsigned lane) {
    if constexpr (G == 32) {
        return FULL;
    } else {
        return ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
    }
}

// mover–mover check (basic:355-424), warp-collective: EVERY lane of the warp must call it.
//   part : this lane takes part (its env is being checked and the lane holds a mover)
// returns this lane's OR over the pairs it evaluated; the caller ballots over the group.
//   kmask: bit k set = this lane wants its pair (m, m+k) evaluated (a float32 screen may already have cleared the others)
template <int G, bool BOX>
__device__ __forceinline__ bool pair_check(unsigned lane, int m, bool part, double x, double y, double s0, double s1,
                                           const Rect& rect, bool quirk, double quirk_rsum, unsigned kmask = 0xffffffffu) {
    bool hit = false;
    if (G == 1) return false;
    const unsigned base = lane & ~(unsigned)(G - 1);
#pragma unroll
    for (int k = 1; k <= G / 2; ++k) {
        const int pm = (m + k) & (G - 1);
        const int src = (int)(base | (unsigned)pm);
        const double ox = __shfl_sync(FULL, x, src);
        const double oy = __shfl_sync(FULL, y, src);
        const double os0 = __shfl_sync(FULL, s0, src);
        const bool opart = __shfl_sync(FULL, (int)part, src) != 0;
        // for k == G/2 the pair (m, m+G/2) is seen from both ends: only the lower lane evaluates it
        const bool mine = part && opart && ((kmask >> k) & 1u) && !(k == G / 2 && m >= G / 2);
        const double dx = dsub(x, ox), dy = dsub(y, oy);
        const double d2 = dadd(dmul(dx, dx), dmul(dy, dy));
        if (!BOX) {
            const double t = quirk ? quirk_rsum : dadd(s0, os0);  // basic:409
            if (mine && sqrt_le(d2, t)) hit = true;
        } else {
            const double os1 = __shfl_sync(FULL, s1, src);
            // basic:411-414: dist <= 2 * ||(mx, mx)||_1, mx = max of the four half sizes
            const double mxs = fmax(fmax(s0, s1), fmax(os0, os1));
            const double thr = dmul(2.0, dadd(fabs(mxs), fabs(mxs)));
            const bool near = mine && sqrt_le(d2, thr);
            if (__any_sync(FULL, near)) {
                Rect o;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    o.x[q] = __shfl_sync(FULL, rect.x[q], src);
                    o.y[q] = __shfl_sync(FULL, rect.y[q], src);
                }
                // geom:132-135 pairs the edges of the LOWER-indexed mover (r1) with those of the higher one (r2); the
                // segment test is symmetric in (p,q), so which lane evaluates does not matter.
                if (near && rects_intersect(rect, o)) hit = true;
            }
        }
    }
    return hit;
}

// plan:610-645 ensure_max_dyn_val; the common non-clipping branch never takes the square root
__device__ __forceinline__ void ensure_max(double cx, double cy, double maxv, double max2_lo, double dx, double dy,
                                           double dt, double inv_dt, double& nx, double& ny, double& ndx, double& ndy) {
    const double tx = dadd(dmul(dt, dx), cx);
    const double ty = dadd(dmul(dt, dy), cy);
    const double s = dadd(dmul(tx, tx), dmul(ty, ty));
    nx = tx;
    ny = ty;
    ndx = dx;
    ndy = dy;
    if (!(s < max2_lo)) {  // max2_lo = max^2 * (1 - 1e-14): below it sqrt(s) < max for certain (NaN falls through here)
        const double nrm = dsqrt(s);
        if (nrm >= maxv) {  // plan:633
            const double inv_nrm = __drcp_rn(nrm);  // correctly rounded reciprocal, shared by the two quotients
            nx = dmul(maxv, ddiv_rcp(tx, nrm, inv_nrm));
            ny = dmul(maxv, ddiv_rcp(ty, nrm, inv_nrm));
            ndx = ddiv_rcp(dsub(nx, cx), dt, inv_dt);  // inv_dt = RN(1 / dt), from the host
            ndy = ddiv_rcp(dsub(ny, cy), dt, inv_dt);
        }
    }
}

}  // namespace gpr
