"""Search's camera features (SE:1231-1241, 1640-1646) on the CPU oracle: analytic cases that pin the ray caster
(the reference's rasteriser is Isaac Gym's closed binary: no golden images exist), and the pinhole set-up of sdx_camera_t."""
import numpy as np
import pytest

from seqdex_b200.camera import SEARCH_CAMERA, look_at


def _lone_target_env(scene, oracle_lib, n=8):
    """every brick parked under the table except each env's target brick, which lies flat at (0.25, 0.19, 0.75)"""
    o = oracle_lib.OracleEnv(scene, n)
    rows = o.brick_roots()
    rows[:, :, 0:3] = np.array([3.0, 3.0, -2.0], np.float32)       # far away, below the ground slab
    rows[:, :, 3:7] = np.array([0, 0, 0, 1], np.float32)
    rows[:, :, 7:13] = 0
    for e in range(n):
        rows[e, scene.target_brick_index(e), 0:3] = (0.25, 0.19, 0.75)
    o.set_brick_roots(rows)
    # robot folded away: its boxes must not sit between the camera and the brick in these analytic cases
    o.link[:, :, 0:3] = np.array([-3.0, 0.0, 0.3], np.float32)
    return o


def test_look_at_basis():
    c = look_at(**SEARCH_CAMERA)
    f, r, u = (np.array(list(v)) for v in (c.fwd, c.right, c.up))
    assert abs(np.linalg.norm(f) - 1) < 1e-6 and abs(f @ r) < 1e-6 and abs(f @ u) < 1e-6 and abs(r @ u) < 1e-6
    np.testing.assert_allclose(np.cross(r, f), u, atol=1e-6)
    assert f[2] < -0.98 and abs(c.inv_focal - 1.0 / 64.0) < 1e-9 and (c.width, c.height) == (128, 128)     # 90 degrees over 128 px
    with pytest.raises(ValueError):
        look_at((0, 0, 1), (0, 0, 0))


def test_lone_brick_projects_to_its_top_face(scene, oracle_lib):
    o = _lone_target_env(scene, oracle_lib)
    cam = look_at((0.25, 0.19, 1.35), (0.25, 0.19, 0.0), world_up=(0.0, 1.0, 0.0))     # straight down, 128 x 128, 90 degrees
    out = o.segmentation_features(cam)
    half = np.ctypeslib.as_array(scene.c.br_half).reshape(72, 3)
    for e in range(o.n):
        b = scene.target_brick_index(e)
        centre_z = o.brick[e, 2, b]
        D = 1.35 - (centre_z + half[b, 2])                       # camera to the top face
        w_px, h_px = 2 * half[b, 0] * 64 / D, 2 * half[b, 1] * 64 / D
        # the box's sides are visible too (perspective): between the top face and the base footprint seen from D + 2 hz
        lo, hi = (w_px - 1) * (h_px - 1), (w_px + 2) * (h_px + 2)
        assert lo <= out[e, 0] <= hi, (e, out[e], w_px, h_px)
        # centroid at the principal point (the box centre offset br_coff moves it by < 2 px)
        assert abs(out[e, 1] - 63.5) <= 2.5 and abs(out[e, 2] - 63.5) <= 2.5, out[e]
    assert out[:, 0].min() > 20


def test_occlusion_and_emptiness(scene, oracle_lib):
    o = _lone_target_env(scene, oracle_lib, n=8)
    cam = look_at((0.25, 0.19, 1.35), (0.25, 0.19, 0.0), world_up=(0.0, 1.0, 0.0))
    free = o.segmentation_features(cam)
    rows = o.brick_roots()
    tb0 = scene.target_brick_index(0)
    cover = 6 if tb0 != 6 else 5                                   # the 1x4 brick (largest footprint) as the occluder
    # env 0: occluder directly above the target -> nothing of the target is visible
    rows[0, cover, 0:3] = rows[0, tb0, 0:3] + np.array([0, 0, 0.1], np.float32)
    # env 1: occluder above, shifted sideways by the target's half width -> roughly half of it remains
    tb1 = scene.target_brick_index(1)
    half = np.ctypeslib.as_array(scene.c.br_half).reshape(72, 3)
    rows[1, cover, 0:3] = rows[1, tb1, 0:3] + np.array([half[cover, 0] + 0.0, 0, 0.1], np.float32)
    # env 2: the target itself is gone (under the ground): zero pixels, centroid 0, 0 (SE:1237-1239)
    rows[2, scene.target_brick_index(2), 0:3] = (3.0, 3.0, -2.0)
    o.set_brick_roots(rows)
    out = o.segmentation_features(cam)
    assert out[0].tolist() == [0, 0, 0]
    assert 0.2 * free[1, 0] < out[1, 0] < 0.8 * free[1, 0], (free[1], out[1])
    assert out[1, 2] < free[1, 2] or out[1, 1] != free[1, 1]       # centroid moved away from the covered side
    assert out[2].tolist() == [0, 0, 0]
    assert np.array_equal(out[3:], free[3:])                       # untouched envs unchanged


def test_search_camera_sees_the_heap_top(scene, oracle_lib):
    """the reference's camera pose (SE:875, 1 m above the table) over a compact heap: bottom-layer targets are buried under the
    layers above them; lifting one above the heap makes it emerge -- the quantity the emergence reward differentiates
    (SE:1640-1646)"""
    o = oracle_lib.OracleEnv(scene, 8)
    o.link[:, :, 0:3] = np.array([-3.0, 0.0, 0.3], np.float32)
    cam = look_at(**SEARCH_CAMERA)
    rows = o.brick_roots()
    rows[:, :, 2] = np.minimum(rows[:, :, 2], 0.62 + 3 * 0.06)      # the 9-layer drop lattice squeezed below the camera
    o.set_brick_roots(rows)
    buried = o.segmentation_features(cam)
    for e in range(8):
        rows[e, scene.target_brick_index(e), 0:3] = (0.25, 0.19, 0.88 - 0.04 * (e % 2))   # on top of the heap's centre
    o.set_brick_roots(rows)
    top = o.segmentation_features(cam)
    assert (top[:, 0] > buried[:, 0]).all() and (top[:, 0] > 100).all(), (buried[:, 0], top[:, 0])
    assert (buried[:, 0] < 0.5 * top[:, 0]).all()
    assert ((0 <= top[:, 1:]) & (top[:, 1:] < 128)).all()
