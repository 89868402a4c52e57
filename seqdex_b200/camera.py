"""Pinhole camera of the segmentation features (include/seqdex_b200.h ``sdx_camera_t``).
The reference places one camera sensor per env with ``gym.set_camera_location(handle, env, pos, target)`` and
``CameraProperties(width=128, height=128)`` (SE:755-758, 875; Isaac Gym's default ``horizontal_fov`` is 90 degrees)."""
from __future__ import annotations

import ctypes
import math

import numpy as np


class CameraC(ctypes.Structure):
    _fields_ = [("pos", ctypes.c_float * 3), ("fwd", ctypes.c_float * 3), ("right", ctypes.c_float * 3), ("up", ctypes.c_float * 3),
                ("inv_focal", ctypes.c_float), ("width", ctypes.c_int), ("height", ctypes.c_int)]


def look_at(pos, target, width=128, height=128, horizontal_fov=90.0, world_up=(0.0, 0.0, 1.0)) -> CameraC:
    """camera at ``pos`` looking at ``target`` (env-local coordinates), z-up world: right = fwd x up_world, up = right x fwd"""
    p, t, wu = (np.asarray(v, np.float64) for v in (pos, target, world_up))
    f = t - p
    f /= np.linalg.norm(f)
    r = np.cross(f, wu)
    if np.linalg.norm(r) < 1e-9:
        raise ValueError("camera looks along the world up axis: choose another world_up")
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    c = CameraC()
    for dst, src in ((c.pos, p), (c.fwd, f), (c.right, r), (c.up, u)):
        for i in range(3):
            dst[i] = float(src[i])
    c.inv_focal = math.tan(math.radians(horizontal_fov) / 2.0) / (width / 2.0)
    c.width, c.height = int(width), int(height)
    return c


SEARCH_CAMERA = dict(pos=(0.35, 0.19, 1.0), target=(0.2, 0.19, 0.0))      # SE:875
