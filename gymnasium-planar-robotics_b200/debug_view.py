"""2-D debug view of ONE environment of a batched simulation.

The reference draws its 2-D view with Matplotlib (``Matplotlib2DViewer``, utils/rendering.py:283-507): silver tiles, one
rectangle per mover in the mover's colour, the outline of the collision shape (circle or box, plus the outline widened by
the safety offset), an arrow for the velocity and a marker for the goal.  Rendering is off the step path and Matplotlib is
not part of this image, so this module rasterises the same picture into a NumPy ``uint8`` RGB array: for looking at what a
chosen env index of a 65,536-env batch is doing.  Only a few hundred bytes of that env's state cross PCIe.

    img = env.debug_view(env_index=123)            # (H, W, 3) uint8, y axis pointing up like the reference's plot
    debug_view.save_ppm('env123.ppm', img)         # viewable everywhere, no imaging library needed

Axes follow the reference: x along the first layout index (``layout_tiles[i, j]``: tile i along x, j along y).
"""

from __future__ import annotations

import numpy as np

SILVER = (192, 192, 192)
BACKGROUND = (255, 255, 255)
BLACK = (0, 0, 0)
# matplotlib's default colour cycle: the reference's docs use named colours per mover
PALETTE = [(31, 119, 180), (255, 127, 14), (44, 160, 44), (214, 39, 40), (148, 103, 189), (140, 86, 75), (227, 119, 194),
           (127, 127, 127), (188, 189, 34), (23, 190, 207)]
OBJECT_COLOUR = (80, 80, 80)
OBSTACLE_COLOUR = (120, 40, 40)


class _Canvas:
    def __init__(self, width_m: float, height_m: float, ppm: float):
        self.ppm = float(ppm)
        self.W = max(int(np.ceil(width_m * ppm)), 1)
        self.H = max(int(np.ceil(height_m * ppm)), 1)
        self.img = np.empty((self.H, self.W, 3), dtype=np.uint8)
        self.img[:] = BACKGROUND
        # pixel centres in metres; row 0 is the TOP of the picture (largest y)
        self.x = (np.arange(self.W) + 0.5) / self.ppm
        self.y = (self.H - np.arange(self.H) - 0.5) / self.ppm

    def _window(self, x0, x1, y0, y1):
        c0 = int(np.clip(np.floor(x0 * self.ppm), 0, self.W))
        c1 = int(np.clip(np.ceil(x1 * self.ppm), 0, self.W))
        r0 = int(np.clip(np.floor(self.H - y1 * self.ppm), 0, self.H))
        r1 = int(np.clip(np.ceil(self.H - y0 * self.ppm), 0, self.H))
        return r0, r1, c0, c1

    def _paint(self, r0, r1, c0, c1, mask, colour):
        if r1 > r0 and c1 > c0:
            self.img[r0:r1, c0:c1][mask] = colour

    def rect(self, cx, cy, hx, hy, colour, cos=1.0, sin=0.0, outline=False, width_px=1.5):
        """Rectangle with half sizes (hx, hy), rotated by the angle given through (cos, sin)."""
        ext = abs(cos) * hx + abs(sin) * hy + 2 / self.ppm, abs(sin) * hx + abs(cos) * hy + 2 / self.ppm
        r0, r1, c0, c1 = self._window(cx - ext[0], cx + ext[0], cy - ext[1], cy + ext[1])
        X, Y = np.meshgrid(self.x[c0:c1] - cx, self.y[r0:r1] - cy)
        u, v = cos * X + sin * Y, -sin * X + cos * Y  # into the rectangle's frame
        if outline:
            w = width_px / self.ppm
            d = np.maximum(np.abs(u) - hx, np.abs(v) - hy)
            mask = np.abs(d) <= w / 2
        else:
            mask = (np.abs(u) <= hx) & (np.abs(v) <= hy)
        self._paint(r0, r1, c0, c1, mask, colour)

    def circle(self, cx, cy, r, colour, outline=False, width_px=1.5):
        e = r + 2 / self.ppm
        r0, r1, c0, c1 = self._window(cx - e, cx + e, cy - e, cy + e)
        X, Y = np.meshgrid(self.x[c0:c1] - cx, self.y[r0:r1] - cy)
        d = np.sqrt(X * X + Y * Y)
        mask = (np.abs(d - r) <= width_px / self.ppm / 2) if outline else (d <= r)
        self._paint(r0, r1, c0, c1, mask, colour)

    def segment(self, x0, y0, x1, y1, colour, width_px=1.5):
        w = width_px / self.ppm
        r0, r1, c0, c1 = self._window(min(x0, x1) - w, max(x0, x1) + w, min(y0, y1) - w, max(y0, y1) + w)
        X, Y = np.meshgrid(self.x[c0:c1], self.y[r0:r1])
        dx, dy = x1 - x0, y1 - y0
        L2 = dx * dx + dy * dy
        t = np.clip(((X - x0) * dx + (Y - y0) * dy) / L2, 0.0, 1.0) if L2 > 0 else np.zeros_like(X)
        d = np.sqrt((X - x0 - t * dx) ** 2 + (Y - y0 - t * dy) ** 2)
        self._paint(r0, r1, c0, c1, d <= w / 2, colour)


def rasterize_scene(layout_tiles, tile_half, mover_pos, mover_half, c_shape, c_size, c_offset=0.0, goals=None, mover_vel=None,
                    mover_yaw=None, object_pose=None, object_half=None, object_goal=None, goal_radius=None, ppm=400.0,
                    colours=None, velocity_scale=0.1, obstacles=None):
    """Draw one environment.

    layout_tiles (nx, ny) 0/1; tile_half (2,) half tile size in m; mover_pos (N, 2); mover_half (N, 2) or (2,) half sizes of
    the mover bodies; c_shape 'circle' | 'box'; c_size scalar / (2,) / per-mover like ``collision_params['size']``;
    c_offset the safety offset (outline drawn when > 0); goals (N, 2) or None; mover_vel (N, 2) or None (arrow of length
    ``velocity_scale`` s * v); mover_yaw (N,) or None; object_pose (x, y, cos, sin) + object_half for the pushing env;
    object_goal (2,); goal_radius: radius of the goal ring (``threshold_pos``); obstacles: the ``obstacles`` kwarg of the
    planning env ((K, 3) circles / (K, 4) boxes).  Returns (H, W, 3) uint8.
    """
    layout = np.asarray(layout_tiles)
    nx, ny = layout.shape
    hx, hy = float(tile_half[0]), float(tile_half[1])
    cv = _Canvas(nx * 2 * hx, ny * 2 * hy, ppm)
    for i in range(nx):
        for j in range(ny):
            if layout[i, j]:
                cx, cy = (i + 0.5) * 2 * hx, (j + 0.5) * 2 * hy
                cv.rect(cx, cy, hx, hy, SILVER)
                cv.rect(cx, cy, hx, hy, BACKGROUND, outline=True, width_px=1.0)  # tile joints
    if obstacles is not None:
        for o in np.asarray(obstacles, dtype=np.float64):  # rows [x, y, size..., (vx, vy)]: already at the time of the view
            if c_shape == 'circle':
                cv.circle(o[0], o[1], o[2], OBSTACLE_COLOUR)
            else:
                cv.rect(o[0], o[1], o[2], o[3], OBSTACLE_COLOUR)
    pos = np.asarray(mover_pos, dtype=np.float64).reshape(-1, 2)
    N = pos.shape[0]
    mh = np.broadcast_to(np.asarray(mover_half, dtype=np.float64).reshape(-1, 2), (N, 2))
    cs = np.asarray(c_size, dtype=np.float64)
    if c_shape == 'circle':
        cs = np.broadcast_to(cs.reshape(-1, 1), (N, 1)) if cs.ndim <= 1 else cs.reshape(N, -1)
    else:
        cs = np.broadcast_to(cs.reshape(-1, 2), (N, 2))
    colours = colours or PALETTE
    if object_goal is not None:
        g = np.asarray(object_goal, dtype=np.float64).reshape(2)
        cv.circle(g[0], g[1], goal_radius or 0.01, OBJECT_COLOUR, outline=True, width_px=2.0)
        cv.circle(g[0], g[1], 2.5 / cv.ppm, OBJECT_COLOUR)
    if goals is not None:
        gl = np.asarray(goals, dtype=np.float64).reshape(-1, 2)
        for m in range(min(N, gl.shape[0])):
            col = colours[m % len(colours)]
            cv.circle(gl[m, 0], gl[m, 1], goal_radius or 0.01, col, outline=True, width_px=2.0)
            cv.circle(gl[m, 0], gl[m, 1], 2.5 / cv.ppm, col)
    if object_pose is not None:
        ox, oy, oc, osn = (float(t) for t in object_pose)
        oh = np.broadcast_to(np.asarray(object_half if object_half is not None else 0.035, dtype=np.float64), (2,))
        cv.rect(ox, oy, oh[0], oh[1], OBJECT_COLOUR, cos=oc, sin=osn)
    for m in range(N):
        col = colours[m % len(colours)]
        yaw = float(mover_yaw[m]) if mover_yaw is not None else 0.0
        c, s = np.cos(yaw), np.sin(yaw)
        cv.rect(pos[m, 0], pos[m, 1], mh[m, 0], mh[m, 1], col, cos=c, sin=s)
        cv.rect(pos[m, 0], pos[m, 1], mh[m, 0], mh[m, 1], BLACK, cos=c, sin=s, outline=True, width_px=1.0)
        # collision shape (black) and, with a safety offset, the shape widened by it (mover colour)
        for off, colr in ((0.0, BLACK), (float(c_offset), col)):
            if colr is not BLACK and off <= 0.0:
                continue
            if c_shape == 'circle':
                cv.circle(pos[m, 0], pos[m, 1], cs[m, 0] + off, colr, outline=True, width_px=2.0)
            else:
                cv.rect(pos[m, 0], pos[m, 1], cs[m, 0] + off, cs[m, 1] + off, colr, cos=c, sin=s, outline=True, width_px=2.0)
        if mover_vel is not None:
            v = np.asarray(mover_vel, dtype=np.float64).reshape(-1, 2)[m]
            cv.segment(pos[m, 0], pos[m, 1], pos[m, 0] + velocity_scale * v[0], pos[m, 1] + velocity_scale * v[1], BLACK, 2.0)
    return cv.img


def save_ppm(path: str, img: np.ndarray) -> None:
    """Binary PPM (P6): no imaging library needed, opens in any viewer."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    with open(path, 'wb') as f:
        f.write(b'P6\n%d %d\n255\n' % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def view_of_env(env, env_index: int = 0, ppm: float = 400.0) -> np.ndarray:
    """Pull the state of ONE env of a ``Benchmark*VecEnv`` / single env / ParallelEnv and draw it."""
    vec = getattr(env, '_vec', env)  # (the single-env and PettingZoo classes wrap a vector env)
    if not 0 <= env_index < vec.num_envs:
        raise IndexError(f'env_index {env_index} out of range (num_envs={vec.num_envs})')
    d = vec.derived
    st = vec.get_state()
    one = {k: v[env_index].detach().cpu().numpy() for k, v in st.items() if k in ('pos', 'vel', 'goal', 'object_pos', 'mover_rot')}
    cfg = vec.cfg
    nx, ny = int(cfg.num_tiles_x), int(cfg.num_tiles_y)
    layout = np.asarray(list(cfg.layout)[: nx * ny], dtype=np.uint8).reshape(nx, ny)
    kw = dict(layout_tiles=layout, tile_half=(cfg.tile_half[0], cfg.tile_half[1]), mover_pos=one['pos'],
              mover_half=np.asarray(d['mover_size'])[:, :2], c_shape=d['c_shape'], c_size=d['c_size'],
              c_offset=float(d['c_size_offset']), mover_vel=one['vel'], goal_radius=float(cfg.threshold_pos), ppm=ppm)
    if 'object_pos' in one:  # pushing: the goal belongs to the object
        rot = one['mover_rot']
        kw.update(object_pose=one['object_pos'], object_half=float(cfg.object_half_xy), object_goal=one['goal'].reshape(-1)[:2],
                  mover_yaw=np.array([np.arctan2(rot[1], rot[0])]))
    else:
        ob = d.get('obstacles')
        if ob is not None and len(ob):
            # extra bodies with a prescribed velocity: where they are after the env-steps of the running episode
            ob = np.array(ob, dtype=np.float64)
            t = float(st['elapsed_steps'][env_index]) * int(cfg.num_cycles) * float(cfg.cycle_time)
            ob[:, :2] += ob[:, -2:] * t
        kw.update(goals=one['goal'], obstacles=ob if ob is not None and len(ob) else None)
    return rasterize_scene(**kw)
