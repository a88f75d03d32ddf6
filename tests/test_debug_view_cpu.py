"""CPU tests of the 2-D debug view: the rasteriser is plain NumPy and needs no GPU."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location('gpr_debug_view', os.path.join(ROOT, 'gymnasium-planar-robotics_b200', 'debug_view.py'))
dv = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(dv)


def px(img, x, y, ppm):
    """Pixel that contains the point (x, y) in metres; y points up."""
    return tuple(int(c) for c in img[img.shape[0] - 1 - int(y * ppm), int(x * ppm)])


def test_scene_geometry_and_colours():
    ppm = 500.0
    layout = np.array([[1, 1, 1], [1, 0, 1], [1, 1, 1]])
    pos = np.array([[0.12, 0.12], [0.60, 0.36]])
    img = dv.rasterize_scene(layout, (0.12, 0.12), pos, (0.0775, 0.0775), 'circle', 0.11, goals=np.array([[0.6, 0.12], [0.12, 0.6]]),
                             mover_vel=np.array([[1.0, 0.0], [0.0, 0.0]]), goal_radius=0.1, ppm=ppm)
    assert img.shape == (360, 360, 3) and img.dtype == np.uint8
    assert px(img, 0.36, 0.36, ppm) == dv.BACKGROUND           # the missing centre tile
    assert px(img, 0.36, 0.12, ppm) == dv.SILVER               # a tile
    assert px(img, 0.12 - 0.05, 0.12 + 0.05, ppm) == dv.PALETTE[0]  # body of mover 0
    assert px(img, 0.60 + 0.05, 0.36 - 0.05, ppm) == dv.PALETTE[1]  # body of mover 1
    assert px(img, 0.12 + 0.05, 0.12, ppm) == dv.BLACK         # velocity arrow of mover 0 (0.1 s * 1 m/s along +x)
    assert px(img, 0.12, 0.12 + 0.11, ppm) == dv.BLACK         # collision circle of mover 0
    assert px(img, 0.6, 0.12, ppm) == dv.PALETTE[0]            # goal dot of mover 0
    assert px(img, 0.6 + 0.1, 0.12, ppm) == dv.PALETTE[0]      # goal ring (threshold_pos)


def test_box_shape_offset_outline_and_object():
    ppm = 400.0
    img = dv.rasterize_scene(np.ones((2, 3)), (0.12, 0.12), np.array([[0.24, 0.36]]), (0.0775, 0.0775), 'box', np.array([0.09, 0.1]),
                             c_offset=0.02, mover_yaw=np.array([0.0]), object_pose=(0.24, 0.6, 1.0, 0.0), object_half=0.035,
                             object_goal=(0.1, 0.1), goal_radius=0.05, ppm=ppm)
    assert img.shape == (288, 192, 3)                          # x: 2 tiles = 0.48 m wide, y: 3 tiles = 0.72 m high
    assert px(img, 0.24 + 0.09, 0.36, ppm) == dv.BLACK         # box outline at +sx
    assert px(img, 0.24, 0.36 + 0.1, ppm) == dv.BLACK          # box outline at +sy
    assert px(img, 0.24 + 0.11, 0.36 + 0.05, ppm) == dv.PALETTE[0]  # the outline widened by the safety offset
    assert px(img, 0.24, 0.6, ppm) == dv.OBJECT_COLOUR         # the pushed object
    assert px(img, 0.1, 0.1, ppm) == dv.OBJECT_COLOUR          # its goal


def test_rotated_object_and_ppm_file(tmp_path):
    c, s = np.cos(np.pi / 4), np.sin(np.pi / 4)
    img = dv.rasterize_scene(np.ones((1, 1)), (0.12, 0.12), np.zeros((0, 2)), (0.07, 0.07), 'circle', 0.11,
                             object_pose=(0.12, 0.12, c, s), object_half=0.05, ppm=1000.0)
    assert px(img, 0.12 + 0.065, 0.12, 1000.0) == dv.OBJECT_COLOUR   # a corner of the 45-degree square reaches 0.0707 along x
    assert px(img, 0.12 + 0.045, 0.12 + 0.045, 1000.0) == dv.SILVER  # where the unrotated square's corner would be
    path = tmp_path / 'v.ppm'
    dv.save_ppm(str(path), img)
    raw = path.read_bytes()
    assert raw.startswith(b'P6\n240 240\n255\n') and len(raw) == len(b'P6\n240 240\n255\n') + 240 * 240 * 3


def test_obstacles_are_drawn():
    img = dv.rasterize_scene(np.ones((2, 2)), (0.12, 0.12), np.array([[0.1, 0.1]]), (0.05, 0.05), 'circle', 0.06,
                             obstacles=[[0.3, 0.3, 0.05]], ppm=500.0)
    assert px(img, 0.3, 0.3, 500.0) == dv.OBSTACLE_COLOUR and px(img, 0.3 + 0.07, 0.3, 500.0) == dv.SILVER
    img = dv.rasterize_scene(np.ones((2, 2)), (0.12, 0.12), np.array([[0.1, 0.1]]), (0.05, 0.05), 'box', (0.06, 0.06),
                             obstacles=[[0.3, 0.3, 0.08, 0.02]], ppm=500.0)
    assert px(img, 0.3 + 0.07, 0.3, 500.0) == dv.OBSTACLE_COLOUR and px(img, 0.3, 0.3 + 0.04, 500.0) == dv.SILVER
