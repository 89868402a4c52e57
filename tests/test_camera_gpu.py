"""k_seg_features (csrc/sdx_camera.cuh) through the C-ABI against the oracle's full-image ray caster: integer outputs, BIT-EXACT.
The kernel only visits the pixels of the target's projected bounding rectangle -- the comparison against the oracle's
every-pixel loop is also the proof that the pruning is conservative."""
import numpy as np
import pytest
import torch

from seqdex_b200.camera import SEARCH_CAMERA, look_at

pytestmark = pytest.mark.gpu


def _pair(scene, oracle_lib, n):
    from seqdex_b200.env import SdxEnv
    return SdxEnv(scene, n), oracle_lib.OracleEnv(scene, n)


def _push(g, o):
    for name, v in (("BRICK", o.brick), ("LINK", o.link), ("DOF", o.dof)):
        g.tensor(name).copy_(torch.from_numpy(v))


@pytest.mark.parametrize("cam_kw", [SEARCH_CAMERA, dict(pos=(0.25, 0.19, 1.35), target=(0.25, 0.19, 0.0), world_up=(0.0, 1.0, 0.0)),
                                    dict(pos=(0.9, -0.3, 0.9), target=(0.25, 0.19, 0.65), width=96, height=64, horizontal_fov=60.0)])
def test_segmentation_features_bit_exact(scene, oracle_lib, cam_kw):
    n = 24
    g, o = _pair(scene, oracle_lib, n)
    rng = np.random.default_rng(5)
    rows = o.brick_roots()
    rows[:, :, 2] = np.minimum(rows[:, :, 2], 0.62 + 3 * 0.06)
    rows[:, :, 0:2] += rng.uniform(-0.02, 0.02, size=(n, 72, 2)).astype(np.float32)
    q = rng.normal(size=(n, 72, 4)).astype(np.float32) * np.array([0.3, 0.3, 1.0, 1.0], np.float32)   # tumbled bricks
    rows[:, :, 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    for e in range(0, n, 2):                      # half of the targets on top of the heap, the rest wherever they fell
        rows[e, scene.target_brick_index(e), 0:3] = (0.25 + 0.1 * rng.uniform(-1, 1), 0.19 + 0.1 * rng.uniform(-1, 1), 0.86)
    rows[5, scene.target_brick_index(5), 0:3] = (3.0, 3.0, -2.0)          # a target that is nowhere to be seen
    o.set_brick_roots(rows)
    o.dof[:, 0, 1] += 0.6; o.dof[:, 0, 3] += 0.5                          # the arm reaches over the bin: robot boxes occlude
    o.refresh_links()
    _push(g, o)
    cam = look_at(**cam_kw)
    got = g.segmentation_features(cam).cpu().numpy()
    want = o.segmentation_features(cam)
    assert np.array_equal(got, want), (got[:8], want[:8])
    assert (want[:, 0] > 0).sum() >= 6 and (want[:, 0] == 0).sum() >= 1, want[:, 0]


def test_segmentation_features_after_simulation(scene, oracle_lib):
    """on states the contact step itself produced (bricks settling from the lattice, hand moving)"""
    n = 8
    g, o = _pair(scene, oracle_lib, n)
    cam = look_at(**SEARCH_CAMERA)
    for _ in range(12):
        g.simulate(); o.simulate()
    got = g.segmentation_features(cam).cpu().numpy()
    assert np.array_equal(got, o.segmentation_features(cam))
    assert g.launch_count() > 12
