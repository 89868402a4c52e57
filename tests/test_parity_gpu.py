"""GPU parity: every CUDA kernel of the hot path against the CPU oracle on the same seeded inputs,
called through the C-ABI (seqdex_b200.env.SdxEnv is a ctypes shim).  Bar: BIT-EXACT (the contact step and
the task ops share an explicit rounding contract with the oracle: no FMA contraction, own sincos/exp)."""
import numpy as np
import pytest
import torch

from tests.util import lattice_bank

pytestmark = pytest.mark.gpu


def _mk(scene, oracle_lib, n, seed=3, jitter=True):
    from seqdex_b200.env import SdxEnv
    g = SdxEnv(scene, n)
    o = oracle_lib.OracleEnv(scene, n)
    w = oracle_lib.default_tvalue_weights(1)
    g.set_tvalue_weights(w)
    o.tv = w
    if jitter:
        rng = np.random.default_rng(seed)
        j = rng.uniform(-0.01, 0.01, size=(n, 2, 72)).astype(np.float32)
        o.brick[:, 0:2, :] += j
        g.tensor("BRICK")[:, 0:2, :] += torch.from_numpy(j).cuda()
    return g, o


def _cmp(name, a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        raise AssertionError(f"{name}: {len(bad)} of {a.size} differ; max abs diff {d.max():.3e} first at {bad[0]} "
                             f"gpu={a[tuple(bad[0])]!r} oracle={b[tuple(bad[0])]!r}")


def test_initial_state_matches(scene, oracle_lib):
    g, o = _mk(scene, oracle_lib, 4, jitter=False)
    _cmp("brick", g.tensor("BRICK"), o.brick)
    _cmp("dof", g.tensor("DOF"), o.dof)
    _cmp("link", g.tensor("LINK"), o.link)
    _cmp("jac7", g.tensor("JAC7"), o.jac7)


@pytest.mark.parametrize("steps", [1, 4, 12])
def test_contact_step_bit_exact(scene, oracle_lib, steps):
    g, o = _mk(scene, oracle_lib, 6)
    g.tensor("CONTACTS")   # switch the debug dump on
    for _ in range(steps):
        g.simulate()
        o.simulate(dump=True)
    torch.cuda.synchronize()
    _cmp("ncontact", g.tensor("NCONTACT"), o.ncontact)
    nc = o.ncontact[:, 0]
    gc = g.tensor("CONTACTS").cpu().numpy()
    for e in range(o.n):   # contact-pair bookkeeping: same pairs, same order, same points
        _cmp(f"contact words env{e}", gc[e, :nc[e], 0].view(np.uint32), o.condump[e, :nc[e], 0].view(np.uint32))
        _cmp(f"contact rows env{e}", gc[e, :nc[e], 1:6], o.condump[e, :nc[e], 1:6])
    _cmp("brick", g.tensor("BRICK"), o.brick)
    _cmp("dof", g.tensor("DOF"), o.dof)
    _cmp("link", g.tensor("LINK"), o.link)
    _cmp("jac7", g.tensor("JAC7"), o.jac7)
    _cmp("netf", g.tensor("NETF"), o.netf)
    _cmp("sleep counters", g.tensor("SLEEP"), o.slp)
    _cmp("impulse-cache counts", g.tensor("WSN"), o.wsn)
    gws = g.tensor("WS").cpu().numpy()
    for e in range(o.n):     # warm-start cache: same keys, same impulses (latest buffer)
        _cmp(f"impulse cache env{e}", gws[e, o.ws_cur, :nc[e]], o.ws[e, o.ws_cur, :nc[e]])
    assert o.ncontact[:, 0].max() > 100, "test must exercise contacts"
    assert (np.abs(o.ws[:, o.ws_cur, :, 1:]).sum() > 0), "impulses must be cached"


def test_robot_contacts_exercised(scene, oracle_lib):
    """drive the hand down into the heap so robot-brick contacts and joint-space impulses are covered"""
    g, o = _mk(scene, oracle_lib, 4)
    tgt = o.dof[:, 2, :].copy()
    tgt[:, 1] += 0.9; tgt[:, 3] += 0.6
    o.dof[:, 2, :] = tgt
    g.tensor("DOF")[:, 2, :] = torch.from_numpy(tgt).cuda()
    g.tensor("CONTACTS")
    robot_seen = 0
    for _ in range(60):
        g.simulate(); o.simulate(dump=True)
        w = o.condump[0, :o.ncontact[0, 0], 0].view(np.uint32)
        robot_seen = max(robot_seen, int((((w & 255) >= 72) | ((((w >> 8) & 255) >= 72) & (((w >> 8) & 255) < 255))).sum()))
    torch.cuda.synchronize()
    _cmp("brick", g.tensor("BRICK"), o.brick)
    _cmp("dof", g.tensor("DOF"), o.dof)
    _cmp("link", g.tensor("LINK"), o.link)
    _cmp("netf", g.tensor("NETF"), o.netf)
    assert robot_seen > 0, "hand never touched the heap: test does not cover robot contacts"


def test_sleeping_and_waking_bit_exact(oracle_lib):
    """bricks dropped into the bin fall asleep (short timer so it happens within the test), the hand then ploughs into
    the heap and wakes some of them: counters, poses and contact counts must follow the oracle bit for bit"""
    from seqdex_b200.scene import Scene
    sc = Scene(sleep_time=0.1)
    ns = sc.c.sleep_substeps
    assert ns == 12
    g, o = _mk(sc, oracle_lib, 4)
    slept = woken = 0
    for t in range(150):
        if t == 70:                                    # now send the hand down into the heap
            tgt = o.dof[:, 2, :].copy()
            tgt[:, 1] += 0.9; tgt[:, 3] += 0.6
            o.dof[:, 2, :] = tgt
            g.tensor("DOF")[:, 2, :] = torch.from_numpy(tgt).cuda()
        before = o.slp.copy()
        g.simulate(); o.simulate()
        slept = max(slept, int((o.slp >= ns).sum()))
        woken += int(((before >= ns) & (o.slp < ns)).sum())
        if t % 10 == 9 or t in (70, 71, 72):
            torch.cuda.synchronize()
            _cmp(f"sleep@{t}", g.tensor("SLEEP"), o.slp)
            _cmp(f"ncontact@{t}", g.tensor("NCONTACT"), o.ncontact)
            _cmp(f"brick@{t}", g.tensor("BRICK"), o.brick)
            _cmp(f"dof@{t}", g.tensor("DOF"), o.dof)
    assert slept > 4 * 36, f"only {slept} brick-states ever asleep: the test does not cover sleeping"
    assert woken > 0, "nothing was ever woken: the test does not cover waking"


def test_tvalue_dataset_and_grasp_bank_bit_exact(scene, oracle_lib, tmp_path):
    """the device rings of SURVEY 8f.1 (t-value training rows, grasp terminal states) against the oracle's sequential
    loops, and their export in the reference's pickle layout"""
    from seqdex_b200 import bank_io
    n = 64
    g, o = _mk(scene, oracle_lib, n, jitter=False)
    bank = lattice_bank(scene, 4)
    g.set_heap_bank(bank); o.set_heap_bank(bank)
    w = oracle_lib.default_tvalue_weights(1).copy()
    w[-2:] = [-4.0, 4.0]                                   # bias the gate open so that SOME grasps bank (sigmoid(z1) > 0.8)
    g.set_tvalue_weights(w); o.tv = w
    g.enable_tvalue_dataset(32); o.enable_tvalue_dataset(32)   # small rings: they wrap
    rng = np.random.default_rng(9)
    for t in range(160):
        a = rng.uniform(-1.2, 1.2, size=(n, 23)).astype(np.float32)
        g.step(torch.from_numpy(a).cuda()); o.step(a)
    gs, gf, gc = g.tvalue_dataset()
    _cmp("dataset counts", gc, o.tvd_counts)
    assert o.tvd_counts.sum() >= n and o.tvd_counts.min() >= 0
    _cmp("failure rows", gf, o.tvd_fail[: gf.shape[0]])
    _cmp("success rows", gs, o.tvd_succ[: gs.shape[0]])
    hand, obj, idx = g.grasp_bank()
    _cmp("grasp bank index", idx, o.gb_index)
    _cmp("grasp bank hand", hand, o.gb_hand)
    _cmp("grasp bank obj", obj, o.gb_obj)
    bank_io.save_grasp_bank(g, tmp_path / "hand.pkl", tmp_path / "obj.pkl")
    import pickle
    with open(tmp_path / "hand.pkl", "rb") as f:
        hl = pickle.load(f)
    with open(tmp_path / "obj.pkl", "rb") as f:
        ol = pickle.load(f)
    assert len(hl) == len(ol) == 8 and tuple(hl[0].shape) == (11024, 23, 2) and tuple(ol[0].shape) == (11024, 1, 13)   # GS:390-395


def test_full_step_bit_exact(scene, oracle_lib):
    """VecTask.step semantics end to end: reset_idx -> pre_physics -> simulate -> post_physics, 160 steps
    (crosses an episode boundary at progress 149, so resets, banking and the scripted lift are covered)."""
    n = 16
    g, o = _mk(scene, oracle_lib, n, jitter=False)
    bank = lattice_bank(scene, 4)
    g.set_heap_bank(bank); o.set_heap_bank(bank)
    rng = np.random.default_rng(5)
    for t in range(160):
        a = rng.uniform(-1.2, 1.2, size=(n, 23)).astype(np.float32)
        g.step(torch.from_numpy(a).cuda())
        o.step(a)
        if t in (0, 1, 2, 50, 80, 148, 149, 150, 159):
            torch.cuda.synchronize()
            for name, ov in (("OBS", o.obs), ("STATES", o.states), ("REW", o.rew), ("RESET", o.reset), ("PROGRESS", o.progress),
                             ("TVALUE", o.tvalue), ("TARGET_INIT", o.target_init), ("DOF", o.dof), ("BRICK", o.brick),
                             ("EPISODE", o.episode), ("CONSEC", o.consec), ("SLEEP", o.slp)):
                _cmp(f"{name}@{t}", g.tensor(name), ov)
    assert o.episode.min() >= 2


def test_gae_bit_exact(oracle_lib):
    import ctypes
    from seqdex_b200 import _lib
    H, n = 8, 1000
    rng = np.random.default_rng(0)
    r, v = rng.normal(size=(H, n)).astype(np.float32), rng.normal(size=(H, n)).astype(np.float32)
    d = (rng.uniform(size=(H, n)) < 0.1).astype(np.float32)
    lv, ld = rng.normal(size=n).astype(np.float32), (rng.uniform(size=n) < 0.1).astype(np.float32)
    adv, ret = oracle_lib.gae(r, v, d, lv, ld, 0.99, 0.95)
    tr, tv, td, tlv, tld = (torch.from_numpy(x).cuda() for x in (r, v, d, lv, ld))
    ga, gr = torch.empty_like(tr), torch.empty_like(tr)
    L = _lib.load()
    _lib.check(L.sdx_gae(*(ctypes.c_void_p(t.data_ptr()) for t in (tr, tv, td, tlv, tld, ga, gr)), H, n,
                         ctypes.c_float(0.99), ctypes.c_float(0.95), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    _cmp("adv", ga, adv); _cmp("ret", gr, ret)


def test_overflow_of_candidate_lists_and_contact_table_bit_exact(scene, oracle_lib):
    """maximum sizes: all 72 bricks crushed into one small volume -> every owner sees more than KC = 32 candidates and the env
    more than SDX_MAX_CONTACTS = 1024 contacts.  What is kept -- statics first in the candidate lists, then the lowest dynamic
    targets; speculative contacts shed level by level before a touching one is lost -- what is dropped (and counted) and the
    resulting state must be the oracle's, bit for bit: the kept set is defined by the ascending sweep, not by thread timing."""
    n = 5                                               # odd env count: tails of every warp-per-env / 4-envs-per-warp kernel
    g, o = _mk(scene, oracle_lib, n, jitter=False)
    rng = np.random.default_rng(11)
    rows = o.brick_roots()
    rows[:, :, 0] = 0.25 + rng.uniform(-0.05, 0.05, size=(n, 72)).astype(np.float32)
    rows[:, :, 1] = 0.19 + rng.uniform(-0.05, 0.05, size=(n, 72)).astype(np.float32)
    rows[:, :, 2] = 0.66 + rng.uniform(0.0, 0.06, size=(n, 72)).astype(np.float32)
    q = rng.normal(size=(n, 72, 4)).astype(np.float32)
    rows[:, :, 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    rows[:, :, 7:13] = 0
    o.set_brick_roots(rows)
    g.tensor("BRICK").copy_(torch.from_numpy(o.brick))
    g.tensor("CONTACTS")
    seen_drop = seen_shed = seen_cand = 0
    for t in range(6):
        g.simulate(); o.simulate(dump=True)
        torch.cuda.synchronize()
        _cmp(f"ncontact@{t}", g.tensor("NCONTACT"), o.ncontact)
        seen_drop = max(seen_drop, int(o.ncontact[:, 1].max()))
        seen_shed = max(seen_shed, int(o.ncontact[:, 2].max()))
        seen_cand = max(seen_cand, int((o.ncontact[:, 3] & 0xFFFF).max()))
        assert int((o.ncontact[:, 3] >> 16).max()) == 0           # no brick ever loses a pair against a static box
        nc = o.ncontact[:, 0]
        gc = g.tensor("CONTACTS").cpu().numpy()
        for e in range(n):
            _cmp(f"contact words env{e}@{t}", gc[e, :nc[e], 0].view(np.uint32), o.condump[e, :nc[e], 0].view(np.uint32))
        _cmp(f"brick@{t}", g.tensor("BRICK"), o.brick)
        _cmp(f"impulse-cache counts@{t}", g.tensor("WSN"), o.wsn)
    assert seen_cand > 0 and seen_shed > 0, (o.ncontact, "the crush must overflow the candidate lists and make the contact table shed")
    assert np.isfinite(o.brick).all()


def test_scene_without_contacts_bit_exact(scene, oracle_lib):
    """empty input: every brick parked far apart in free fall (no contact at all, zero-length lists everywhere), robot idle"""
    n = 3
    g, o = _mk(scene, oracle_lib, n, jitter=False)
    rows = o.brick_roots()
    k = np.arange(72)
    rows[:, :, 0] = (-2.0 + 0.5 * (k % 9)).astype(np.float32)
    rows[:, :, 1] = (3.0 + 0.5 * (k // 9)).astype(np.float32)
    rows[:, :, 2] = 5.0
    rows[:, :, 7:13] = 0
    o.set_brick_roots(rows)
    g.tensor("BRICK").copy_(torch.from_numpy(o.brick))
    for _ in range(3):
        g.simulate(); o.simulate()
    torch.cuda.synchronize()
    _cmp("ncontact", g.tensor("NCONTACT"), o.ncontact)
    assert o.ncontact[:, 0].max() == 0
    _cmp("brick", g.tensor("BRICK"), o.brick)
    _cmp("dof", g.tensor("DOF"), o.dof)
    assert (o.brick[:, 9, :] < 0).all()                 # everything is falling
