"""TEST-ONLY: a second, independently written implementation of the planar pushing substep (NumPy float64).

Why it exists: the CUDA kernels and the C oracle both compile ``include/gpr_push_physics.h``,
so their bit-exact agreement says nothing about whether that header implements the model it documents.  This module restates
the SAME model from first principles in a different form — generalised coordinates, explicit contact Jacobians, the
constraint-space matrices of MuJoCo's documentation ("Computation" chapter: a1 = A f + a0 with A = J M^-1 J^T,
reference acceleration a_ref = -b (J v) - k r, regulariser R = (1 - d)/d * diag(A), impedance d(r) from solimp,
(b, k) from solref) — and shares no code with the header (it does not read it, include it or call the oracle).
``tests/test_pushing_independent.py`` compares the two.

The model (a specification of this project — MuJoCo itself is not available, parity with it is unpinned):
    q = (x_M, y_M, psi_M, x_O, y_O, psi_O),  M = diag(m_M, m_M, I_M, m_O, m_O, I_O),  I = m (lx^2 + ly^2) / 12
    smooth forces : mover x/y actuator m_M u; mover yaw impedance tau = -k_r psi - 2 sqrt(k_r m_M) psi_dot;
                    object joint damping -D v on its three planar dofs
    constraints   : <= 2 mover-object contact points (normal n from mover to object, tangent t = (-n_y, n_x)) with a
                    friction interval |f_t| <= mu f_n; four ground-friction points at the object's bottom corners, each a
                    2-D force limited to the disc mu m_O g / 4, a_ref = -b v (no position term), R from d0
    solver        : projected Gauss-Seidel, `iterations` sweeps, rows in the order (n_1, t_1, n_2, t_2, corners 1..4), optionally
                    WARM-STARTED from the previous substep's forces (the corner forces always, the contact forces when the
                    number of contact points is unchanged) — MuJoCo's default
    integrator    : qacc = qacc_smooth + M^-1 J^T f; the object's joint damping implicit: (M + dt D) a = f_total (MuJoCo's
                    Euler integrator); v += dt a; x += dt v; orientation kept as (cos, sin), advanced by the second-order
                    rotation (1 - a^2/2, a - a^3/6) and renormalised.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Params:
    dt: float = 0.001
    m_M: float = 1.24
    half_M: tuple = (0.155 / 2, 0.155 / 2)
    m_O: float = 0.01
    half_O: float = 0.035
    D: float = 0.01
    mu: float = 1.0
    g: float = 9.81
    k_rot: float = 0.1
    solref: tuple = (0.02, 1.0)
    solimp: tuple = (0.9, 0.95, 0.001, 0.5, 2.0)
    iterations: int = 3

    @property
    def I_M(self):
        return self.m_M * ((2 * self.half_M[0]) ** 2 + (2 * self.half_M[1]) ** 2) / 12.0

    @property
    def I_O(self):
        return self.m_O * 2 * (2 * self.half_O) ** 2 / 12.0


def impedance(P: Params, r: float) -> float:
    """MuJoCo solimp (d0, dwidth, width, midpoint, power): d(r) rises from d0 to dwidth over `width` of violation."""
    d0, dw, width, mid, power = P.solimp
    x = abs(r) / width
    if x >= 1.0:
        return dw
    if x <= 0.0:
        return d0
    if x <= mid:
        y = x**power / mid ** (power - 1)
    else:
        y = 1.0 - (1.0 - x) ** power / (1.0 - mid) ** (power - 1)
    return d0 + y * (dw - d0)


def _axes(c, s):
    return np.array([c, s]), np.array([-s, c])


def _support(c, s, half, n):
    ex, ey = _axes(c, s)
    return half[0] * abs(n @ ex) + half[1] * abs(n @ ey)


def manifold(pM, rotM, halfM, pO, rotO, halfO):
    """Planar box-box contact: separating-axis test over the four face normals, least penetration picks the reference face
    (ties: the earlier axis, with a 1e-12 bias), the incident edge of the other box is clipped against the reference face's
    side planes; a clipped end point below the reference face is a contact point, placed midway between the two surfaces.
    Returns [(point, normal from mover to object, depth)], at most two."""
    d = pO - pM
    cand = [*_axes(*rotM), *_axes(*rotO)]
    best, best_sep, best_n = -1, -np.inf, None
    for k, ax in enumerate(cand):
        n = ax if d @ ax >= 0 else -ax
        sep = d @ n - (_support(*rotM, halfM, n) + _support(*rotO, halfO, n))
        if sep > 0:
            return []
        if sep > best_sep + 1e-12:
            best, best_sep, best_n = k, sep, n
    ref_is_M = best < 2
    (pR, rotR, halfR), (pI, rotI, halfI) = ((pM, rotM, halfM), (pO, rotO, halfO)) if ref_is_M else ((pO, rotO, halfO), (pM, rotM, halfM))
    rn = best_n if ref_is_M else -best_n  # from the reference box towards the incident box
    ix, iy = _axes(*rotI)
    cx, cy = rn @ ix, rn @ iy
    if abs(cx) >= abs(cy):  # incident face: outward normal most anti-parallel to rn
        fc, ft, fext = pI + (-1.0 if cx > 0 else 1.0) * halfI[0] * ix, iy, halfI[1]
    else:
        fc, ft, fext = pI + (-1.0 if cy > 0 else 1.0) * halfI[1] * iy, ix, halfI[0]
    v0, v1 = fc - fext * ft, fc + fext * ft
    rt = np.array([-rn[1], rn[0]])
    rext, roff = _support(*rotR, halfR, rt), _support(*rotR, halfR, rn)
    t0, t1 = (v0 - pR) @ rt, (v1 - pR) @ rt
    if t0 > t1:
        t0, t1, v0, v1 = t1, t0, v1, v0
    if t1 < -rext or t0 > rext:
        return []
    span = t1 - t0
    if t0 < -rext and span > 0:
        v0 = v0 + (-rext - t0) / span * (v1 - v0)
        t0 = -rext
    if t1 > rext and span > 0:
        v1 = v0 + (rext - t0) / (t1 - t0) * (v1 - v0)
        t1 = rext
    pts = []
    for e in (v0, v1):
        depth = roff - (e - pR) @ rn
        if depth >= 0:
            pts.append((e + 0.5 * depth * rn, best_n.copy(), depth))
    if len(pts) == 2 and abs(t1 - t0) < 1e-9:
        pts = pts[:1]
    return pts


def _cross(r, d):
    return r[0] * d[1] - r[1] * d[0]


def _row(pM, pO, pt, d):
    """Jacobian row of (object point velocity - mover point velocity) . d  w.r.t. (vM, wM, vO, wO)."""
    rA, rB = pt - pM, pt - pO
    return np.array([-d[0], -d[1], -_cross(rA, d), d[0], d[1], _cross(rB, d)])


def substep(P: Params, M: np.ndarray, O: np.ndarray, u: np.ndarray, iterations: int | None = None, warm: dict | None = None):
    """One substep.  M, O: [x, y, cos, sin, vx, vy, w].  Returns (M', O', mover qacc_xy, number of contact points).
    warm: None = cold start; a dict = warm start, read (keys 'n', 'f', 'g'; empty = no forces yet) and updated in place."""
    it = P.iterations if iterations is None else iterations
    dt = P.dt
    Minv = np.array([1 / P.m_M, 1 / P.m_M, 1 / P.I_M, 1 / P.m_O, 1 / P.m_O, 1 / P.I_O])
    pM, pO = M[:2], O[:2]
    vel = np.array([M[4], M[5], M[6], O[4], O[5], O[6]])
    yaw = np.arctan2(M[3], M[2])
    tau = P.k_rot * (0.0 - yaw) - 2.0 * np.sqrt(P.k_rot * P.m_M) * M[6]
    a_smooth = np.array([u[0], u[1], tau / P.I_M, -P.D * O[4] / P.m_O, -P.D * O[5] / P.m_O, -P.D * O[6] / P.I_O])
    dmax, (tc, dr) = P.solimp[1], P.solref
    b, kfac = 2.0 / (dmax * tc), 1.0 / (dmax**2 * tc**2 * dr**2)

    rm = np.hypot(*P.half_M)
    ro = np.hypot(P.half_O, P.half_O)
    near = (pO - pM) @ (pO - pM) <= ((rm + ro) * 1.0001) ** 2
    pts = manifold(pM, (M[2], M[3]), P.half_M, pO, (O[2], O[3]), (P.half_O, P.half_O)) if near else []
    moving = bool(np.any(O[4:7] != 0.0))
    f_gen = np.zeros(6)  # J^T f
    if pts or moving:
        rows, aref, R, diag, kind = [], [], [], [], []
        for pt, n, depth in pts:
            t = np.array([-n[1], n[0]])
            Jn, Jt = _row(pM, pO, pt, n), _row(pM, pO, pt, t)
            d = impedance(P, depth)
            Ann, Att = Jn @ (Minv * Jn), Jt @ (Minv * Jt)
            Rn = (1.0 - d) / d * Ann
            rows += [Jn, Jt]
            aref += [-b * (Jn @ vel) + kfac * d * depth, -b * (Jt @ vel)]
            R += [Rn, Rn]  # impratio = 1, mu = 1: the friction row shares the normal row's regulariser
            diag += [Ann + Rn, Att + Rn]
            kind += ['n', 't']
        # ground friction at the object's bottom corners: diagonal approximated isotropically, 1/m + |r|^2 / (2 I)
        h = P.half_O
        Ag = 1.0 / P.m_O + (2.0 * h * h) / P.I_O * 0.5
        Rg = (1.0 - P.solimp[0]) / P.solimp[0] * Ag
        ex, ey = _axes(O[2], O[3])
        corner_rows = []
        for lx, ly in ((-h, -h), (-h, h), (h, h), (h, -h)):
            r = lx * ex + ly * ey
            Jx = np.array([0, 0, 0, 1.0, 0.0, -r[1]])
            Jy = np.array([0, 0, 0, 0.0, 1.0, r[0]])
            corner_rows.append((Jx, Jy, -b * (Jx @ vel), -b * (Jy @ vel)))
        lim = P.mu * P.m_O * P.g / 4.0
        f = np.zeros(len(rows))
        fg = np.zeros((4, 2))
        if warm:  # previous substep's forces and the generalised force they amount to
            if pts and warm['n'] == len(pts):
                f = warm['f'].copy()
            fg = warm['g'].copy()
            for i, J in enumerate(rows):
                f_gen = f_gen + J * f[i]
            for g_, (Jx, Jy, _, _) in enumerate(corner_rows):
                f_gen = f_gen + Jx * fg[g_, 0] + Jy * fg[g_, 1]
        for _ in range(it):
            for i, J in enumerate(rows):
                acc = J @ (a_smooth + Minv * f_gen)
                new = f[i] - (acc - aref[i] + R[i] * f[i]) / diag[i]
                if kind[i] == 'n':
                    new = max(new, 0.0)
                else:
                    cone = P.mu * f[i - 1]
                    new = min(max(new, -cone), cone)
                f_gen = f_gen + J * (new - f[i])
                f[i] = new
            for g_, (Jx, Jy, arx, ary) in enumerate(corner_rows):
                cur = a_smooth + Minv * f_gen
                nx = fg[g_, 0] - (Jx @ cur - arx + Rg * fg[g_, 0]) / (Ag + Rg)
                ny = fg[g_, 1] - (Jy @ cur - ary + Rg * fg[g_, 1]) / (Ag + Rg)
                mag = np.hypot(nx, ny)
                if mag > lim:
                    nx, ny = nx * lim / mag, ny * lim / mag
                f_gen = f_gen + Jx * (nx - fg[g_, 0]) + Jy * (ny - fg[g_, 1])
                fg[g_] = (nx, ny)
        if warm is not None:
            warm.update(n=len(pts), f=f.copy(), g=fg.copy())
    elif warm is not None:
        warm.clear()  # a free substep has no constraint forces
    qacc = a_smooth + Minv * f_gen
    # the object's joint damping is implicit in MuJoCo's Euler integrator: (M + dt D) a = -D v + J^T f
    qacc[3] = (-P.D * O[4] + f_gen[3]) / (P.m_O + dt * P.D)
    qacc[4] = (-P.D * O[5] + f_gen[4]) / (P.m_O + dt * P.D)
    qacc[5] = (-P.D * O[6] + f_gen[5]) / (P.I_O + dt * P.D)

    def advance(body, a3):
        x, y, c, s, vx, vy, w = body
        vx, vy, w = vx + dt * a3[0], vy + dt * a3[1], w + dt * a3[2]
        x, y = x + dt * vx, y + dt * vy
        ang = dt * w
        if w != 0.0:
            ca, sa = 1.0 - 0.5 * ang * ang, ang - ang**3 / 6.0
            c, s = c * ca - s * sa, s * ca + c * sa
            nrm = np.hypot(c, s)
            c, s = c / nrm, s / nrm
        return np.array([x, y, c, s, vx, vy, w])

    return advance(M, qacc[:3]), advance(O, qacc[3:]), qacc[:2].copy(), len(pts)
