"""The C-ABI shared library loads without a GPU and exports every symbol include/seqdex_b200.h declares; the scene
struct has the same size on both sides of the ABI (product and oracle).  No compute call is made here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "seqdex_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from seqdex_b200 import _lib
    _lib.build()
    L = _lib.load()
    names = _declared()
    assert len(names) >= 40, names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/seqdex_b200.h but not exported: {missing}"


def test_scene_struct_abi(scene, oracle_lib):
    from seqdex_b200 import _lib
    L = _lib.load()
    assert L.sdx_scene_size() == ctypes.sizeof(scene.c) == oracle_lib.lib().sdxo_scene_size()
    assert L.sdx_sim_smem_bytes() <= 227 * 1024 // 3, "contact-step tile must leave room for 3 CTAs per SM"


def test_create_without_gpu_fails_loudly(scene):
    """no CPU fallback: without a CUDA device the library refuses (and says so)"""
    import torch
    if torch.cuda.is_available():
        return
    from seqdex_b200 import _lib
    L = _lib.load()
    h = ctypes.c_void_p()
    rc = L.sdx_create(ctypes.byref(scene.c), 8, 0, ctypes.c_uint64(1), ctypes.byref(h))
    assert rc != 0 and b"CUDA" in L.sdx_last_error()
    import pytest
    from seqdex_b200.env import SdxEnv
    with pytest.raises(RuntimeError, match="CUDA"):
        SdxEnv(scene, 8)


def test_scene_tables(scene):
    import numpy as np
    from seqdex_b200 import robot_data as RD
    c = scene.c
    assert (c.n_bricks, c.n_fixed, c.n_rshapes) == (72, 60, 26) and c.n_static <= 24
    assert RD.BODY_NAMES[7] == "panda_link7" and [RD.BODY_NAMES[i] for i in (11, 19, 23, 15)] == ["link_3.0", "link_7.0", "link_11.0", "link_15.0"]
    # DoF order is Isaac Gym's (index, thumb, middle, ring): thumb limits sit at slots 11-14
    assert abs(scene.dof_lo[11] - 0.263) < 1e-6 and abs(scene.dof_hi[11] - 1.396) < 1e-6
    masks = np.ctypeslib.as_array(c.link_anc_mask)
    assert masks[7] == 0b1111111 and masks[11] == 0b1111111 | (0b1111 << 7) and masks[0] == 0
    assert [scene.target_brick_index(e) for e in range(8)] == [0, 1, 2, 0, 0, 5, 6, 0]      # GS:962-975
    inertia = np.ctypeslib.as_array(c.dof_inertia)
    assert np.all(inertia > 0) and inertia[:4].min() > 1.0 and inertia[7:].max() < 0.01
    rows = scene.static_actor_roots()
    assert len(rows) == 142 - 72 and np.allclose(rows[141][:3], [0.25, -0.19, 0.618])
