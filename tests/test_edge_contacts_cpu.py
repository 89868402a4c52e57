"""Edge-edge contacts of the contact step, checked on the CPU oracle (the CUDA kernel is bit-identical to it,
tests/test_parity_gpu.py): the corner-vs-face test alone cannot see two boxes that cross edge over edge."""
import os
import sys

import numpy as np

from seqdex_b200.scene import Scene

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
EDGE_BIT = 1 << 27


def _qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return (aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw,
            aw * bw - ax * bx - ay * by - az * bz)


def _ridges(oracle_lib, edge_contacts, overlap):
    """two 1x2 bricks far above the scene, no gravity, each rolled 45 degrees about its long (x) axis so that one long edge points down /
    up; the lower one is then yawed 90 degrees: its top ridge runs along y, the upper one's bottom ridge along x.  The ridges cross
    mid-way, `overlap` deep; their ends (the corners) are 3 cm away from the crossing and touch nothing."""
    s = Scene(edge_contacts=edge_contacts, sleep_time=0.0)
    s.c.n_bricks = 2
    s.c.gravity_z = 0.0
    half = np.ctypeslib.as_array(s.c.br_half).reshape(-1, 3)
    assert np.allclose(half[0], half[1]) and half[0, 0] >= 0.03
    e = oracle_lib.OracleEnv(s, 1)
    b = e.brick                                     # [1][13][NB]: centre of mass = box centre, quaternion xyzw
    c, sn = np.cos(np.pi / 8), np.sin(np.pi / 8)
    roll = (sn, 0.0, 0.0, c)
    yaw = (0.0, 0.0, np.sin(np.pi / 4), np.cos(np.pi / 4))
    r = (half[0, 1] + half[0, 2]) / np.sqrt(2)      # height of a rolled brick's ridge above / below its centre
    b[0, :, :2] = 0
    b[0, 0:3, 0] = (0.0, 0.0, 3.0)
    b[0, 3:7, 0] = _qmul(yaw, roll)
    b[0, 0:3, 1] = (0.0, 0.0, 3.0 + 2 * r - overlap)
    b[0, 3:7, 1] = roll
    e.slp[:] = 0
    return s, e


def test_crossing_ridges_get_one_edge_contact_and_separate(oracle_lib):
    s, e = _ridges(oracle_lib, True, 0.003)
    e.simulate(dump=True)
    n = int(e.ncontact[0, 0])
    words = e.condump[0, :n, 0].view(np.uint32)
    edge = [i for i in range(n) if words[i] & EDGE_BIT]
    assert n == 1 and len(edge) == 1, (n, [hex(w) for w in words])
    i = edge[0]
    assert (int(words[i]) & 255, (int(words[i]) >> 8) & 255) == (0, 1)              # owner = the lower index, once per unordered pair
    half = np.ctypeslib.as_array(s.c.br_half).reshape(-1, 3)
    off = (half[0, 2] - half[0, 1]) / np.sqrt(2)                                    # a rolled box's ridge is off-centre by (h_z - h_y) / sqrt 2
    cross = [off, off]
    np.testing.assert_allclose(e.condump[0, i, 1:3], cross, atol=2e-4)              # at the crossing of the two ridges
    assert 0.0005 < e.condump[0, i, 4] < 0.0031                                     # depth: the overlap along z (second sub-step: partly pushed out)
    assert e.brick[0, 9, 1] > 0.0 and e.brick[0, 9, 0] < 0.0                        # pushed apart along z
    np.testing.assert_allclose(e.brick[0, 9, 0] / s.c.br_invm[0], -e.brick[0, 9, 1] / s.c.br_invm[1], rtol=1e-4)   # equal and opposite impulses
    for _ in range(20):
        e.simulate(dump=True)
    gap = (e.brick[0, 2, 1] - e.brick[0, 2, 0])
    s2, e2 = _ridges(oracle_lib, True, 0.0)
    assert gap > (e2.brick[0, 2, 1] - e2.brick[0, 2, 0]) - 0.0006                   # back out to the slop


def test_without_edge_contacts_the_ridges_pass_through_each_other(oracle_lib):
    s, e = _ridges(oracle_lib, False, 0.003)
    e.simulate(dump=True)
    assert int(e.ncontact[0, 0]) == 0 and e.brick[0, 9, 1] == 0.0                   # what the corner-vs-face test alone sees: nothing


def test_settled_heap_true_overlaps(oracle_lib):
    """the lattice of 72 bricks dropped and left to settle (sleeping off): TRUE overlaps of all pairs of bricks by the full 15-axis
    separating-axis test in numpy, independent of the contacts generated.  Measured here (2 envs, 150 steps, resting contacts warm-started
    at 0.98): without edge contacts 54 of 206 overlapping pairs deeper than 2 mm, 22 deeper than 5 mm, deepest 18.8 mm; with them 1 of 203,
    none, 2.1 mm (8 envs x 240 steps: DESIGN.md section 3c)."""
    import edge_contact_audit as A
    off = A.audit(False, 2, 150)
    on = A.audit(True, 2, 150)
    print(off, on)
    assert on["deeper_5mm"] <= 1 and off["deeper_5mm"] >= 10
    assert on["deeper_2mm"] <= 6 and on["deeper_2mm"] * 5 <= off["deeper_2mm"]
    assert on["median_mm"] < 0.9 and on["max_mm"] < 6.0 and on["speed_p95"] < 0.05


def test_boxes_with_a_parallel_axis_pair_get_no_edge_contact(oracle_lib):
    """two FLAT bricks (z axes parallel) crossing each other in the plane at 45 degrees, interpenetrating sideways: every edge-pair axis
    coincides with a face axis of one of them, so the pair is skipped before the nine-axis loop (edge_axes_parallel) -- whatever the
    corner-vs-face test generates, no contact carries the edge bit"""
    s = Scene(edge_contacts=True, sleep_time=0.0, substeps=1)      # one sub-step: the dump shows the contacts of the poses set here
    s.c.n_bricks = 2
    s.c.gravity_z = 0.0
    e = oracle_lib.OracleEnv(s, 1)
    b = e.brick
    half = np.ctypeslib.as_array(s.c.br_half).reshape(-1, 3)
    b[0, :, :2] = 0
    b[0, 0:3, 0] = (0.0, 0.0, 3.0)
    b[0, 3:7, 0] = (0.0, 0.0, 0.0, 1.0)
    b[0, 0:3, 1] = (half[0, 0] * 0.9, 0.0, 3.0 + 0.2 * half[0, 2])
    b[0, 3:7, 1] = (0.0, 0.0, np.sin(np.pi / 8), np.cos(np.pi / 8))                  # yawed 45 degrees about the shared z axis
    e.slp[:] = 0
    e.simulate(dump=True)
    n = int(e.ncontact[0, 0])
    words = e.condump[0, :n, 0].view(np.uint32)
    assert n > 0 and not any(int(w) & EDGE_BIT for w in words), [hex(w) for w in words]


def test_resting_contacts_are_warm_started_harder_than_hot_ones(oracle_lib):
    """sdx_scene_t::warm_start (0.98) seeds a persisting contact between bodies at rest, warm_start_hot (0.85) one that involves a hot brick
    (faster than the wake threshold / touched by the robot in the last sub-step): a brick resting on the slab keeps its support impulse
    almost entirely from one sub-step to the next, so the solver's first pass has little left to find"""
    def first_pass_support(hot):
        s = Scene(sleep_time=0.0)
        s.c.n_bricks = 1
        s.c.iters = 0                                     # only the warm-start pass (it = -1) acts: what is carried over is all there is
        e = oracle_lib.OracleEnv(s, 1)
        s16 = Scene(sleep_time=0.0)
        s16.c.n_bricks = 1
        e16 = oracle_lib.OracleEnv(s16, 1)
        rows = e16.brick_roots()
        rows[0, 0, 0:3] = (0.25, 0.2, 0.62)               # over the slab of the bin
        rows[0, 0, 3:7] = (0, 0, 0, 1)
        rows[0, 0, 7:13] = 0
        e16.set_brick_roots(rows)
        for _ in range(60):
            e16.simulate()                                # settle with the full solver: the cache now holds the support impulses
        assert abs(e16.brick[0, 9, 0]) < 5e-3
        # hand the settled state and the impulse cache to the zero-iteration env
        e.brick[:] = e16.brick
        e.ws[:] = e16.ws
        e.wsn[:] = e16.wsn
        e.ws_cur = e16.ws_cur
        e.slp[:] = 0 if hot else 5                        # 0 = hot in the last sub-step
        e.simulate()
        return float(e.brick[0, 9, 0])                    # vertical velocity after one step carried by the cached impulses alone
    vz_rest, vz_hot = first_pass_support(False), first_pass_support(True)
    # gravity adds -g h per sub-step; the cached support takes 98 % / 85 % of it back
    assert vz_hot < vz_rest < 0.0, (vz_rest, vz_hot)
    assert vz_rest > 0.4 * vz_hot, (vz_rest, vz_hot)
