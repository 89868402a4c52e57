"""The on-disk formats of SURVEY.md section 8f.1: heap bank / grasp bank pickles in the reference's layout and the
t-value dataset names; plus the oracle's dataset recorder (the GPU recorder is compared with it in test_parity_gpu.py)."""
import pickle

import numpy as np
import torch

from seqdex_b200 import bank_io
from tests.util import lattice_bank


def test_heap_bank_pickle_round_trip_in_reference_layout(scene, tmp_path):
    bank = lattice_bank(scene, 3)
    p = tmp_path / "saved_searching_ternimal_states_good_mo_tvalue.pkl"
    bank_io.save_heap_bank(p, bank, scene)
    with open(p, "rb") as f:
        ref = pickle.load(f)                                   # what GS:412-413 loads
    assert isinstance(ref, list) and len(ref) == 8
    for t in ref:
        assert isinstance(t, torch.Tensor) and tuple(t.shape) == (3, 132, 13) and t.dtype == torch.float32
        # GS:1511: rows are written to root_state_tensor[lego_indices].view(132, 13): 72 free bricks, then the fixed floor
        np.testing.assert_array_equal(t[:, 72:].numpy(), np.broadcast_to(np.ctypeslib.as_array(scene.c.fixed_root).reshape(60, 13), (3, 60, 13)))
    back = bank_io.load_heap_bank(p)
    np.testing.assert_array_equal(back.numpy(), bank)


def test_reference_written_bank_is_trimmed_to_its_written_rows(scene, tmp_path):
    """the reference preallocates 10000 + 1024 rows per type and fills a prefix (SE:319-331); the zero rows behind it (zero
    quaternions!) must never reach ``set_heap_bank``, whose reset kernel samples ``slot % K`` over every row it is given"""
    bank = lattice_bank(scene, 7)
    ref = []
    for ty in range(8):
        t = torch.zeros(11024, 132, 13)
        k = 7 if ty != 5 else 4                                # type 5 has banked fewer heaps
        t[:k, :72] = torch.from_numpy(bank[ty, :k])
        ref.append(t)
    p = tmp_path / "saved_searching_ternimal_states_good_mo_tvalue.pkl"
    with open(p, "wb") as f:
        pickle.dump(ref, f)
    back = bank_io.load_heap_bank(p)
    assert tuple(back.shape) == (8, 4, 72, 13)                 # K = the fewest leading written rows of any type
    q = back[..., 3:7]
    assert float((q.norm(dim=-1) - 1).abs().max()) < 1e-5      # no zero-quaternion row survives
    ref[2][:] = 0
    with open(p, "wb") as f:
        pickle.dump(ref, f)
    import pytest
    with pytest.raises(ValueError):
        bank_io.load_heap_bank(p)


def test_tvalue_dataset_names_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    s, f = rng.normal(size=(5, 4)).astype(np.float32), rng.normal(size=(12, 4)).astype(np.float32)
    p = tmp_path / "tvalue_rows.npz"
    bank_io.save_tvalue_dataset(p, s, f)
    z = np.load(p)
    assert "success_dataset/0th_success_data" in z and "failure_dataset/11th_failure_data" in z      # GS:1409 / TVT:140 names
    s2, f2 = bank_io.load_tvalue_dataset(p)
    np.testing.assert_array_equal(s2, s)
    np.testing.assert_array_equal(f2, f)


def test_oracle_dataset_recorder_follows_the_reference_gating(scene, oracle_lib):
    n = 16
    o = oracle_lib.OracleEnv(scene, n)
    o.set_heap_bank(lattice_bank(scene, 2))
    o.enable_tvalue_dataset(8)                                 # tiny ring: it must wrap
    rng = np.random.default_rng(1)
    resets = 0
    for t in range(80):
        o.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))
        resets += int(o.reset.sum())
    # every reset after the first step contributes exactly one row (GS:1402-1438: success or failure)
    pending = int(o.reset.sum())                               # flagged by the last step, recorded by the next one
    assert int(o.tvd_counts.sum()) == resets - pending and resets - pending > 8
    assert o.tvd_counts[1] > 0                                 # an untrained hand fails
    assert np.abs(np.linalg.norm(o.tvd_fail[: min(8, o.tvd_counts[1])], axis=1) - 1).max() < 1e-4   # rows are unit quaternions
