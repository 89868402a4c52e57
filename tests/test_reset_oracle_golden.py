"""reset_idx of BlockAssemblyGraspSim: the oracle against golden vectors produced by executing the
reference's own reset_idx (oracle/gen_golden_reset.py; `…grasp_sim.py:1361-1553`).  The reference picks the heap to restore with
Python's `random`; the generator fed it the slots the oracle's Philox stream selects, so everything else is comparable: the grasp
terminal-state banking gate and its ring bookkeeping incl. the wrap at 5000, the restored heap (poses from the pickle row,
velocities zeroed), the hand reset, segmentation_target_init_*, the per-env counters -- and that envs which do not reset are
left alone."""
import ctypes
import os

import numpy as np
import pytest

from seqdex_b200.scene import Scene

G = os.path.join(os.path.dirname(__file__), "golden", "reset_idx.npz")


def test_reset_idx_matches_the_reference(oracle_lib):
    from oracle.oracle import fp, ip, lp
    g = np.load(G)
    scene = Scene()
    n = g["reset"].shape[0]
    per_type = g["bank"].shape[1]
    o = oracle_lib.OracleEnv(scene, n, seed=int(g["seed"][0]))
    o.set_brick_roots(np.ascontiguousarray(g["root_before"][:, 9:81], np.float32))
    o.dof[:, 0, :23] = g["dof_before"][:, :, 0]
    o.dof[:, 1, :23] = g["dof_before"][:, :, 1]
    o.dof[:, 2, :23] = g["targets_before"]
    o.reset[:] = g["reset"]
    o.progress[:] = g["progress_before"]
    o.successes[:] = 1.0
    o.episode[:] = g["episode"]
    o.finger_dist[:] = g["finger_dist"]
    o.tvalue[:] = g["tvalue"]
    o.gb_index[:] = g["index_before"]
    o.slp[:] = 77
    o.wsn[:] = 5
    bank = np.ascontiguousarray(g["bank"], np.float32)
    dof_before = o.dof.copy()
    rows_before = o.brick_roots().copy()
    o.L.sdxo_reset(o.S, n, ctypes.c_uint64(o.seed), fp(bank), per_type, fp(o.brick), fp(o.dof), fp(o.target_init), lp(o.progress),
                   lp(o.reset), fp(o.successes), ip(o.episode), ip(o.wsn), o.slp.ctypes.data_as(ctypes.c_void_p), 1, fp(o.finger_dist),
                   fp(o.tvalue), fp(o.gb_hand), fp(o.gb_obj), ip(o.gb_index))
    did = g["reset"].astype(bool)
    rows = o.brick_roots()
    # the restored heap: the pickle row's poses, zero velocities (root frame <-> COM frame round trip: 1e-6)
    np.testing.assert_allclose(rows[did], g["root_after"][did], rtol=0, atol=2e-6)
    for e in np.nonzero(did)[0]:
        np.testing.assert_allclose(rows[e][:, :7], g["bank"][e % 8, g["slots"][e]][:, :7], rtol=0, atol=2e-6)
        assert np.all(rows[e][:, 7:] == 0)
    assert np.array_equal(rows[~did], rows_before[~did]) and np.array_equal(o.dof[~did], dof_before[~did])     # untouched envs
    # hand reset: prepare pose, scaled finger pose, zero velocities, targets = positions (GS:1524-1536)
    np.testing.assert_allclose(o.dof[did, 0, :23], g["dof_after"][did][:, :, 0], rtol=0, atol=1e-6)
    np.testing.assert_allclose(o.dof[did, 1, :23], g["dof_after"][did][:, :, 1], rtol=0, atol=0)
    np.testing.assert_allclose(o.dof[did, 2, :23], g["cur_targets"][did], rtol=0, atol=1e-6)
    np.testing.assert_allclose(g["prev_targets"][did], g["cur_targets"][did], rtol=0, atol=0)
    # counters and segmentation_target_init_* (GS:1547-1553)
    assert np.array_equal(o.progress, g["progress"]) and np.array_equal(o.reset, g["reset_after"])
    assert np.array_equal(o.successes, g["successes"])
    np.testing.assert_allclose(o.target_init[did, :3], g["target_init_pos"][did], rtol=0, atol=2e-6)
    np.testing.assert_allclose(o.target_init[did, 3:], g["target_init_rot"][did], rtol=0, atol=2e-6)
    # grasp terminal-state rings: which envs were banked, where, and the index after (wrap at > 5000, GS:1441-1443)
    assert np.array_equal(o.gb_index, g["index_after"])
    slots = g["gb_slots"]
    np.testing.assert_allclose(o.gb_hand[:, slots], g["gb_hand"], rtol=0, atol=0)
    np.testing.assert_allclose(o.gb_obj[:, slots], g["gb_obj"], rtol=0, atol=2e-6)
    assert int((np.abs(g["gb_obj"]).sum(axis=-1) > 0).sum()) >= 4                                              # the gate did open
    # ours only: a restored heap wakes its bricks and forgets cached contact impulses
    assert np.all(o.slp[did] == 0) and np.all(o.wsn[did] == 0) and np.all(o.slp[~did] == 77)
    assert np.array_equal(o.episode[did], g["episode"][did] + 1)
