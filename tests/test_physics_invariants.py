"""Physics invariants of the contact step, checked on the CPU oracle (the CUDA kernel is bit-identical to it,
tests/test_parity_gpu.py).  PhysX itself cannot be run here (closed binary) -- these stand in for a reference
trajectory: free fall, momentum exchange, rest stability / bounded penetration, joint limits,
PD tracking, no NaNs from a violent start."""
import numpy as np
import pytest

from seqdex_b200.scene import Scene


def _env(oracle_lib, n_bricks, n=1, **kw):
    s = Scene(**kw)
    s.c.n_bricks = n_bricks
    return s, oracle_lib.OracleEnv(s, n)


def _rows(pos, quat=(0, 0, 0, 1), vel=(0, 0, 0), ang=(0, 0, 0)):
    return list(pos) + list(quat) + list(vel) + list(ang)


def test_free_fall_matches_semi_implicit_euler(oracle_lib):
    s, e = _env(oracle_lib, 1)
    rows = e.brick_roots()
    rows[0, 0] = _rows((0.25, 0.19, 2.0))          # far above everything: no contacts
    e.set_brick_roots(rows)
    h, g = 1.0 / 120.0, -9.81
    z, v = 2.0, 0.0
    for _ in range(10):
        e.simulate()
        for _ in range(2):
            v += h * g
            z += h * v
    r = e.brick_roots()[0, 0]
    assert e.ncontact[0, 0] == 0
    np.testing.assert_allclose(r[2], z, rtol=0, atol=2e-5)
    np.testing.assert_allclose(r[9], v, rtol=0, atol=2e-5)
    np.testing.assert_allclose(r[[0, 1]], [0.25, 0.19], atol=1e-6)


def test_head_on_collision_conserves_momentum(oracle_lib):
    s, e = _env(oracle_lib, 2)
    s.c.gravity_z = 0.0
    s.c.brick_ang_damp = 0.0
    rows = e.brick_roots()
    rows[0, 0] = _rows((0.00, 0.0, 3.0), vel=(0.5, 0, 0))     # brick 0 (1x2) moving +x
    rows[0, 1] = _rows((0.09, 0.0, 3.0), vel=(-0.2, 0, 0))    # brick 1 (1x2_curve) moving -x, 3 cm gap
    e.set_brick_roots(rows)
    m = 1.0 / np.ctypeslib.as_array(s.c.br_invm)[:2]
    p0 = (m[:, None] * e.brick[0, 7:10, :2].T).sum(0)
    touched = 0
    for _ in range(30):
        e.simulate()
        touched = max(touched, int(e.ncontact[0, 0]))
    p1 = (m[:, None] * e.brick[0, 7:10, :2].T).sum(0)
    assert touched > 0, "bricks never met"
    np.testing.assert_allclose(p1, p0, rtol=0, atol=2e-6)      # equal and opposite impulses
    assert e.brick[0, 7, 0] < 0.5 and e.brick[0, 7, 1] > -0.2  # they did exchange momentum
    assert e.brick[0, 7, 1] - e.brick[0, 7, 0] > -1e-4         # and no longer approach each other


def test_bricks_come_to_rest_with_bounded_penetration(oracle_lib):
    s, e = _env(oracle_lib, 8, sleep_time=0.0)       # sleeping off: the solver itself has to hold the bricks still
    rows = e.brick_roots()
    rng = np.random.default_rng(0)
    for b in range(8):        # one layer, just above the fixed floor bricks (top at z = 0.6637), no initial overlap
        rows[0, b] = _rows((0.10 + 0.09 * (b % 4), 0.10 + 0.12 * (b // 4), 0.70), quat=(0, 0, np.sin(0.3 * b), np.cos(0.3 * b)))
    rows[0, :8, 0:2] += rng.uniform(-0.005, 0.005, size=(8, 2))
    e.set_brick_roots(rows.astype(np.float32))
    for _ in range(150):
        e.simulate(dump=True)
    n = e.ncontact[0, 0]
    depth = e.condump[0, :n, 4]
    v = np.linalg.norm(e.brick[0, 7:10, :8], axis=0)
    w = np.linalg.norm(e.brick[0, 10:13, :8], axis=0)
    z = e.brick_roots()[0, :8, 2]
    assert n >= 8 * 3, "every resting brick needs at least 3 support points"
    assert depth.max() < 0.004, f"penetration {depth.max():.4f} m"
    assert v.max() < 0.03 and w.max() < 0.5, (v.max(), w.max())
    assert np.all(np.abs(z - (0.6637 + 0.01875)) < 0.006), z      # resting on the floor bricks' tops
    ke0 = (v ** 2).sum()
    for _ in range(100):
        e.simulate()
    ke1 = (np.linalg.norm(e.brick[0, 7:10, :8], axis=0) ** 2).sum()
    assert ke1 <= ke0 + 1e-3, "resting heap must not gain energy"


def test_joint_limits_and_pd_tracking(oracle_lib):
    s, e = _env(oracle_lib, 0)
    lo, hi = s.dof_lo, s.dof_hi
    e.dof[0, 2, :23] = hi + 1.0                    # targets beyond the upper limits
    e.dof[0, 2, 3] = lo[3] - 1.0                   # and one beyond the lower limit
    for _ in range(240):
        e.simulate()
    q = e.dof[0, 0, :23]
    assert np.all(q <= hi + 1e-6) and np.all(q >= lo - 1e-6)
    assert np.all(np.abs(q[7:] - hi[7:]) < 1e-3), "fingers (k=50, d=1) reach the limit within 4 s"
    assert abs(q[3] - lo[3]) < 0.05
    # PD tracking of an in-range target: critically over-damped arm joint converges monotonically
    s2, e2 = _env(oracle_lib, 0)
    q0 = e2.dof[0, 0, 1]
    e2.dof[0, 2, 1] = q0 - 0.3           # lifts the arm (the other direction pushes the hand into the bin)
    prev, hist = q0, []
    for _ in range(180):
        e2.simulate()
        hist.append(e2.dof[0, 0, 1])
    assert all(b <= a + 1e-6 for a, b in zip(hist, hist[1:])), "no overshoot oscillation"
    assert abs(hist[-1] - (q0 - 0.3)) < 1e-3
    assert e2.ncontact[0, 0] == 0


def test_violent_start_stays_finite(oracle_lib):
    """the reference's own initial condition: 9 layers dropped from up to 1.1 m, first layer overlapping the floor bricks"""
    s, e = _env(oracle_lib, 72, n=2)
    for _ in range(200):
        e.simulate()
    assert np.isfinite(e.brick).all() and np.isfinite(e.dof).all()
    r = e.brick_roots()
    assert r[..., 2].min() > -0.1, "nothing tunnels through the ground plane"
    inside = (np.abs(r[..., 0] - 0.25) < 0.35) & (np.abs(r[..., 1] - 0.19) < 0.26)
    assert inside.mean() > 0.8, "most bricks end up in or next to the bin"


def _layer(e, n_bricks=8, z=0.70):
    rows = e.brick_roots()
    rng = np.random.default_rng(0)
    for b in range(n_bricks):
        rows[0, b] = _rows((0.10 + 0.09 * (b % 4), 0.10 + 0.12 * (b // 4), z), quat=(0, 0, np.sin(0.3 * b), np.cos(0.3 * b)))
    rows[0, :n_bricks, 0:2] += rng.uniform(-0.005, 0.005, size=(n_bricks, 2))
    return rows.astype(np.float32)


def test_resting_bricks_fall_asleep_and_stay_put(oracle_lib):
    """PhysX puts resting actors to sleep (scene.py: sleep_energy / sleep_time); ours: oracle sim_env 'SLEEPING'."""
    s, e = _env(oracle_lib, 8)
    ns = s.c.sleep_substeps
    assert ns == 48 and s.c.sleep_energy > 0
    e.set_brick_roots(_layer(e))
    for _ in range(20):
        e.simulate()
    assert e.ncontact[0, 0] >= 8 * 3 and (e.slp[0, :8] < ns).all()      # landed, supported, still awake (0.4 s not over)
    for _ in range(130):
        e.simulate()
    assert (e.slp[0, :8] >= ns).all(), e.slp[0, :8]
    assert e.ncontact[0, 0] == 0                                       # asleep-vs-static pairs are not even generated
    assert (e.brick[0, 7:13, :8] == 0).all()
    z = e.brick_roots()[0, :8, 2]
    assert np.all(np.abs(z - (0.6637 + 0.01875)) < 0.006), z
    before = e.brick.copy()
    for _ in range(50):
        e.simulate()
    np.testing.assert_array_equal(e.brick, before)                     # bit-stable while asleep


def test_sleeping_brick_wakes_when_hit_and_carries_the_load(oracle_lib):
    s, e = _env(oracle_lib, 9)
    ns = s.c.sleep_substeps
    rows = _layer(e)
    rows[0, 8] = _rows((0.45, 0.10, 0.70))                             # brick 8 parked away from the layer
    e.set_brick_roots(rows)
    for _ in range(150):
        e.simulate()
    assert (e.slp[0, :9] >= ns).all()
    z0 = float(e.brick[0, 2, 0])
    # teleport brick 8 to 12 cm above brick 0 (the facade's indexed root write wakes exactly the actors it touches)
    e.brick[0, 0:3, 8] = e.brick[0, 0:3, 0] + np.array([0, 0, 0.12], np.float32)
    e.brick[0, 7:13, 8] = 0
    e.slp[0, 8] = 0
    woke = False
    for _ in range(60):
        e.simulate()
        woke |= bool(e.slp[0, 0] < ns)
        assert e.brick[0, 2, 0] > z0 - 0.004, "the sleeping brick was pushed into its support"
    assert woke, "the impact did not wake the sleeping brick"
    for _ in range(120):
        e.simulate()
    assert (e.slp[0, :9] >= ns).all(), e.slp[0, :9]                    # everything is quiet again ...
    assert e.brick[0, 2, 8] > z0 + 0.015                               # ... with brick 8 resting ON brick 0, not inside it
    assert np.isfinite(e.brick).all()


def test_sleeping_brick_wakes_when_its_support_is_knocked_away(oracle_lib):
    s, e = _env(oracle_lib, 2)
    ns = s.c.sleep_substeps
    rows = e.brick_roots()
    rows[0, 0] = _rows((0.25, 0.19, 0.70))
    rows[0, 1] = _rows((0.25, 0.19, 0.75))                             # brick 1 stacked on brick 0
    e.set_brick_roots(rows)
    for _ in range(150):
        e.simulate()
    assert (e.slp[0, :2] >= ns).all()
    z1 = float(e.brick[0, 2, 1])
    assert z1 > e.brick[0, 2, 0] + 0.015
    e.brick[0, 7, 0] = 1.5                                             # kick the lower brick out sideways
    e.slp[0, 0] = 0
    for _ in range(120):
        e.simulate()
    assert e.brick[0, 2, 1] < z1 - 0.01, "the upper brick kept floating on a support that is gone"
    assert np.isfinite(e.brick).all()


def test_kept_candidate_lists_cover_what_a_fresh_sweep_finds(oracle_lib):
    """The candidate lists are built once per control step.  Audit over a violent scenario -- the 72 bricks
    dropping from the lattice while the hand flails -- of every pair a fresh per-sub-step sweep would list: what the kept lists
    lack must stay rare, both through KC overflow of the (travel-inflated) lists and through pairs that only came into range
    after the lists were built.  Measured: 0.17 % and 0.026 % of the pairs; the old per-sub-step sweep lost 0.006 % to overflow."""
    import ctypes
    from oracle import oracle
    from tests.util import lattice_bank
    scene, n = Scene(), 16
    L = oracle.lib()
    L.sdxo_reuse_audit(1)
    try:
        o = oracle.OracleEnv(scene, n)
        o.tv = oracle.default_tvalue_weights(1)
        o.set_heap_bank(lattice_bank(scene, 2))
        rng = np.random.default_rng(0)
        for _ in range(60):
            o.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))
        st = (ctypes.c_long * 5)()
        L.sdxo_reuse_stats(st)
        built_pairs, built_missing, kept_pairs, kept_missing, kept_missing_with_room = list(st)
    finally:
        L.sdxo_reuse_audit(0)
    assert built_pairs > 100000 and kept_pairs > 100000
    assert built_missing / built_pairs < 0.005 and kept_missing / kept_pairs < 0.005
    assert kept_missing_with_room / kept_pairs < 0.001
    assert np.isfinite(o.brick).all() and float(o.brick[:, 2, :].min()) > 0.0          # nothing fell through the floor on the way
