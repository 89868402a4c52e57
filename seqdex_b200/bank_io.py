"""Terminal-state banks and t-value datasets in the reference's ON-DISK formats, so the reference's
own tools keep reading what this engine writes and vice versa.

* heap bank the grasp task samples on reset: ``intermediate_state/saved_searching_ternimal_states_*_tvalue.pkl`` = pickle of
  ``list[8]`` (one per target-brick type, ``env % 8``) of ``Tensor[K, 132, 13]`` root rows of all 132 bricks
  (72 free, then the 60 fixed floor bricks; written by Search ``SE:1348-1352``, loaded ``GS:412-413``, sampled ``GS:1508-1511``).
* grasp terminal states: ``saved_grasping_{object,hand}_ternimal_states_*.pkl`` = ``list[8]`` of ``Tensor[11024, 1, 13]`` /
  ``Tensor[11024, 23, 2]`` (``GS:390-405``, dump ``GS:1448-1451``, loaded by InsertSim ``IS:372-375``).
* t-value training rows: HDF5 groups ``success_dataset`` / ``failure_dataset`` with one dataset
  ``"{i}th_success_data"`` / ``"{i}th_failure_data"`` per row (``GS:470-480,1407-1438``; read back ``TVT:132-168``).
  h5py is an optional import: without it the same names are written into an ``.npz`` archive.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch

N_FREE, N_FIXED = 72, 60


def heap_bank_to_reference(bank, scene):
    """ours ``[8, K, 72, 13]`` -> the reference's ``list[8] of Tensor[K, 132, 13]`` (fixed floor bricks appended from the scene)"""
    bank = torch.as_tensor(np.asarray(bank.cpu() if isinstance(bank, torch.Tensor) else bank), dtype=torch.float32)
    fixed = torch.from_numpy(np.ctypeslib.as_array(scene.c.fixed_root).reshape(N_FIXED, 13).astype(np.float32))
    out = []
    for ty in range(8):
        k = bank.shape[1]
        out.append(torch.cat([bank[ty], fixed.unsqueeze(0).expand(k, N_FIXED, 13)], dim=1).contiguous())
    return out


SAMPLE_RANGE = 5000      # GS:1508 samples random.sample(range(0, 5000), 1): rows beyond are never drawn


def heap_bank_from_reference(lst):
    """the reference's list -> ours ``[8, K, 72, 13]`` (velocities zeroed as ``GS:1513`` does on load).

    The reference PREALLOCATES 10000 + 1024 rows per type and fills a prefix (SE:319-331, 1313-1343); the rest stays zero --
    zero positions AND zero quaternions.  ``k_reset`` samples ``slot = r % K`` over all K rows we hand it, so only the leading
    written rows may be kept: K = min over types of (leading non-zero rows, capped at the 5000 the reference samples from)."""
    assert len(lst) == 8, "the bank holds one tensor per target-brick type (env % 8)"
    ts = [torch.as_tensor(t, dtype=torch.float32).reshape(int(t.shape[0]), -1, 13)[:, :N_FREE] for t in lst]
    filled = []
    for ty, t in enumerate(ts):
        written = (t.abs().sum(dim=(1, 2)) > 0).to(torch.int64)
        lead = int(torch.cumprod(written, 0).sum())              # rows before the first all-zero row
        if lead == 0:
            raise ValueError(f"heap bank: brick type {ty} has no written row (the reference fills the rings from slot 0)")
        filled.append(min(lead, SAMPLE_RANGE))
    k = min(filled)
    out = torch.stack([t[:k] for t in ts]).clone()
    out[..., 7:13] = 0
    return out.contiguous()


def save_heap_bank(path, bank, scene):
    with open(path, "wb") as f:
        pickle.dump(heap_bank_to_reference(bank, scene), f)


def load_heap_bank(path):
    with open(path, "rb") as f:
        return heap_bank_from_reference(pickle.load(f))


def search_bank_valid(env):
    """the heaps BlockAssemblySearch banked so far as ``[8, K, 72, 13]`` on the device: what Orient samples on reset (OR:419-420, 1564-1568)"""
    rows, _, index = env.search_bank()
    torch.cuda.synchronize(env.device)
    idx = index.cpu().tolist()
    ring = rows.shape[1]
    filled = [ring if bool(rows[t, ring - 1].abs().sum() > 0) else idx[t] for t in range(8)]
    k = min(filled)
    if k == 0:
        raise RuntimeError(f"BlockAssemblySearch has banked no heap yet for some brick type (per-type counts {filled})")
    out = rows[:, :k].clone()
    out[..., 7:13] = 0
    return out.contiguous()


def search_bank_to_reference(env, scene, rows_per_type=10000 + 1024):
    """(``saved_searching_ternimal_states_list``, ``saved_searching_hand_ternimal_states_list``) as Search pickles them
    (SE:1348-1352): ``list[8]`` of ``Tensor[11024, 132, 13]`` and of ``Tensor[11024, 23, 2]``"""
    rows, hand, _ = env.search_bank()
    torch.cuda.synchronize(env.device)
    rows, hand = rows.cpu(), hand.cpu()
    fixed = torch.from_numpy(np.ctypeslib.as_array(scene.c.fixed_root).reshape(N_FIXED, 13).astype(np.float32))
    heaps, hands = [], []
    for ty in range(8):
        t = torch.zeros(rows_per_type, N_FREE + N_FIXED, 13)
        h = torch.zeros(rows_per_type, 23, 2)
        k = min(rows.shape[1], rows_per_type)
        t[:k, :N_FREE] = rows[ty, :k]
        written = rows[ty, :k].abs().sum(dim=(1, 2)) > 0
        t[:k, N_FREE:][written] = fixed
        h[:k] = hand[ty, :k]
        heaps.append(t); hands.append(h)
    return heaps, hands


def save_search_bank(env, scene, heap_path, hand_path):
    heaps, hands = search_bank_to_reference(env, scene)
    with open(heap_path, "wb") as f:
        pickle.dump(heaps, f)
    with open(hand_path, "wb") as f:
        pickle.dump(hands, f)


def orient_bank_valid(env):
    """the heaps BlockAssemblyOrient banked so far as ``[8, K, 72, 13]`` on the device (K = the fewest any type holds; a ring
    that has wrapped counts as full) -- what the NEXT stage of the chain samples on reset (GS:412-413, 1507-1511)"""
    rows, index = env.orient_heap_bank()
    torch.cuda.synchronize(env.device)
    idx = index.cpu().tolist()
    ring = rows.shape[1]
    filled = [ring if bool(rows[t, ring - 1].abs().sum() > 0) else idx[t] for t in range(8)]     # last slot written <=> wrapped
    k = min(filled)
    if k == 0:
        raise RuntimeError(f"BlockAssemblyOrient has banked no heap yet for some brick type (per-type counts {filled})")
    out = rows[:, :k].clone()
    out[..., 7:13] = 0
    return out.contiguous()


def orient_bank_to_reference(env, scene, rows_per_type=10000 + 1024):
    """``saved_digging_ternimal_states_list`` as the reference pickles it (OR:1510-1512 ->
    ``saved_searching_ternimal_states_good_mo_tvalue.pkl``, loaded by GraspSim GS:412-413): ``list[8]`` of
    ``Tensor[11024, 132, 13]``, rows beyond what has been banked left at zero (OR:404-411 preallocates them)"""
    rows, _ = env.orient_heap_bank()
    torch.cuda.synchronize(env.device)
    rows = rows.cpu()
    fixed = torch.from_numpy(np.ctypeslib.as_array(scene.c.fixed_root).reshape(N_FIXED, 13).astype(np.float32))
    out = []
    for ty in range(8):
        t = torch.zeros(rows_per_type, N_FREE + N_FIXED, 13)
        k = min(rows.shape[1], rows_per_type)
        t[:k, :N_FREE] = rows[ty, :k]
        written = rows[ty, :k].abs().sum(dim=(1, 2)) > 0
        t[:k, N_FREE:][written] = fixed
        out.append(t)
    return out


def save_orient_heap_bank(env, scene, path):
    with open(path, "wb") as f:
        pickle.dump(orient_bank_to_reference(env, scene), f)


def grasp_bank_to_reference(env):
    """device rings of an ``SdxEnv`` -> (hand list[8] of [11024, 23, 2], object list[8] of [11024, 1, 13]) on the CPU"""
    hand, obj, _ = env.grasp_bank()
    torch.cuda.synchronize(env.device)
    return [hand[t].cpu().clone() for t in range(8)], [obj[t].cpu().clone().unsqueeze(1) for t in range(8)]


def save_grasp_bank(env, hand_path, object_path):
    hand, obj = grasp_bank_to_reference(env)
    with open(hand_path, "wb") as f:
        pickle.dump(hand, f)
    with open(object_path, "wb") as f:
        pickle.dump(obj, f)


def save_tvalue_dataset(path, success, failure):
    """``success`` / ``failure``: [n, 4] rows.  ``*.hdf5`` needs h5py (the reference's format); otherwise an ``.npz`` with the same names"""
    success = np.asarray(success.cpu() if isinstance(success, torch.Tensor) else success, np.float32)
    failure = np.asarray(failure.cpu() if isinstance(failure, torch.Tensor) else failure, np.float32)
    if str(path).endswith((".hdf5", ".h5")):
        import h5py   # optional dependency, not in this image
        with h5py.File(path, "w") as f:
            gs, gf = f.create_group("success_dataset"), f.create_group("failure_dataset")
            for i, r in enumerate(success):
                gs.create_dataset(f"{i}th_success_data", data=r)
            for i, r in enumerate(failure):
                gf.create_dataset(f"{i}th_failure_data", data=r)
        return
    arrs = {f"success_dataset/{i}th_success_data": r for i, r in enumerate(success)}
    arrs.update({f"failure_dataset/{i}th_failure_data": r for i, r in enumerate(failure)})
    np.savez(path, **arrs)


def load_tvalue_dataset(path):
    """inverse of ``save_tvalue_dataset`` -> (success [ns, 4], failure [nf, 4]) float32 arrays, rows in index order"""
    def rows(get, keys, tag):
        ks = sorted((k for k in keys if k.endswith(f"_{tag}_data")), key=lambda k: int(k.rsplit("/", 1)[-1].split("th_")[0]))
        return np.stack([np.asarray(get(k), np.float32) for k in ks]) if ks else np.zeros((0, 4), np.float32)
    if str(path).endswith((".hdf5", ".h5")):
        import h5py
        with h5py.File(path, "r") as f:
            keys = [f"{g}/{k}" for g in ("success_dataset", "failure_dataset") for k in f[g].keys()]
            return rows(lambda k: f[k][()], keys, "success"), rows(lambda k: f[k][()], keys, "failure")
    z = np.load(path)
    return rows(lambda k: z[k], list(z.keys()), "success"), rows(lambda k: z[k], list(z.keys()), "failure")
