"""BASELINE.json's full size (16 384 envs on one GPU) through size-independent properties (the oracle cannot run that many envs in
seconds): BATCH INVARIANCE -- env e of the 16 384-env run is bit-identical to env e of a 64-env run that the oracle parity tests
cover, so parity at full size follows --, run-to-run determinism, and physical / bookkeeping invariants over every env."""
import numpy as np
import pytest
import torch

from tests.util import lattice_bank

pytestmark = pytest.mark.gpu
FULL = 16384


def _run(scene, n, steps, seed=22, bank=None, act_seed=7, hook=None):
    from seqdex_b200.env import SdxEnv
    from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights
    g = SdxEnv(scene, n, 0, seed)
    g.set_tvalue_weights(default_tvalue_weights(1))
    if bank is not None:
        g.set_heap_bank(bank)
    if hook:
        hook(g)
    gen = torch.Generator(device="cuda").manual_seed(act_seed)
    acts = torch.rand(steps, FULL, 23, device="cuda", generator=gen) * 2.4 - 1.2     # the same action rows whatever n is
    for t in range(steps):
        if t == 3:                                     # a wave of resets in the middle: envs 0, 5, 10, ... time out
            g.tensor("PROGRESS")[::5] = 148
        g.step(acts[t, :n].contiguous())
    torch.cuda.synchronize()
    return g


NAMES = ("BRICK", "DOF", "LINK", "JAC7", "NETF", "OBS", "STATES", "REW", "RESET", "PROGRESS", "TVALUE", "TARGET_INIT", "EPISODE", "SLEEP",
         "NCONTACT", "WSN")


def test_grasp_sim_full_size_batch_invariance_determinism_and_invariants(scene):
    bank = lattice_bank(scene, 4)
    big = _run(scene, FULL, 12, bank=bank)
    small = _run(scene, 64, 12, bank=bank)
    for name in NAMES:
        a, b = big.tensor(name)[:64], small.tensor(name)
        assert torch.equal(a, b), f"{name}: env e of the {FULL}-env run differs from env e of the 64-env run"
    again = _run(scene, FULL, 12, bank=bank)
    for name in NAMES + ("CONSEC",):
        assert torch.equal(big.tensor(name), again.tensor(name)), f"{name}: two identical runs differ (non-deterministic kernel)"
    # invariants over all 16 384 envs
    brick = big.tensor("BRICK")
    assert torch.isfinite(brick).all() and torch.isfinite(big.tensor("OBS")).all() and torch.isfinite(big.tensor("STATES")).all()
    qn = brick[:, 3:7, :].norm(dim=1)
    assert float((qn - 1).abs().max()) < 1e-4                      # unit quaternions after 24 sub-steps of integration
    assert float(brick[:, 2, :].min()) > 0.55                      # nothing sank through the bin floor / table
    assert float(brick[:, 7:10, :].abs().max()) < 20.0             # no exploding velocities
    prog, ep = big.tensor("PROGRESS"), big.tensor("EPISODE")
    assert int(prog.min()) >= 1 and int(prog.max()) <= 150
    assert set(ep.unique().tolist()) == {1, 2}                     # exactly the envs of the reset wave went through a second reset_idx
    assert bool((ep[::5] == 2).all()) and int((ep == 2).sum()) == len(ep[::5])
    nc = big.tensor("NCONTACT")
    assert int(nc[:, 0].max()) <= 1024 and int(nc[:, 0].min()) > 0
    dof = big.tensor("DOF")
    lo, hi = torch.from_numpy(scene.dof_lo).cuda(), torch.from_numpy(scene.dof_hi).cuda()
    assert bool(((dof[:, 0, :23] >= lo - 1e-6) & (dof[:, 0, :23] <= hi + 1e-6)).all())     # joint limits
    assert bool(((dof[:, 2, :23] >= lo) & (dof[:, 2, :23] <= hi)).all())                   # targets are clamped (GS:1636)


def test_orient_and_search_batch_invariance_across_scripted_resets():
    """4096 envs vs 32 envs through the first scripted reset and a few steps (Orient: 53 contact steps inside reset_idx; Search: 60 +
    the ray-cast render): the first 32 envs agree bit for bit"""
    from seqdex_b200.camera import SEARCH_CAMERA, look_at
    from seqdex_b200.tasks.cfg import scene_from_cfg
    for task, extra in (("BlockAssemblyOrient", ()), ("BlockAssemblySearch", ("SEG", "EMERGENCE", "TVOBS"))):
        sc = scene_from_cfg(task)                       # yaml-stated parameters: contact_offset 0.02
        hook = (lambda g: g.set_camera(look_at(**SEARCH_CAMERA))) if task.endswith("Search") else None
        bank = None if task.endswith("Search") else lattice_bank(sc, 2)
        outs = []
        for n in (4096, 32):
            from seqdex_b200.env import SdxEnv
            from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights
            g = SdxEnv(sc, n, 0, 22)
            g.set_tvalue_weights(default_tvalue_weights(1))
            if bank is not None:
                g.set_heap_bank(bank)
            if hook:
                hook(g)
            gen = torch.Generator(device="cuda").manual_seed(3)
            acts = torch.rand(4, 4096, 23, device="cuda", generator=gen) * 2 - 1
            for t in range(4):
                g.step(acts[t, :n].contiguous())
            torch.cuda.synchronize()
            outs.append(g)
        for name in ("BRICK", "DOF", "LINK", "OBS", "STATES", "REW", "RESET", "PROGRESS", "TARGET_INIT", "EPISODE", "SLEEP") + extra:
            assert torch.equal(outs[0].tensor(name)[:32], outs[1].tensor(name)), (task, name)
        # the scripted resets throw about 3 % of the bricks out of the bin; they hit the ground at 3-4 m/s (2.9 cm of travel per sub-step) and
        # the deepest of 295 k bricks is caught a few millimetres late: nothing may TUNNEL through the ground (thinnest half height 1.44 cm)
        assert torch.isfinite(outs[0].tensor("BRICK")).all() and float(outs[0].tensor("BRICK")[:, 2, :].min()) > -0.014


def test_benchmark_mix_overflow_counters_and_settled_penetration(scene):
    """VERDICT r1 item 8 at BASELINE's full size.  (a) 300 steps of the benchmark's episode mix (random actions, staggered resets) at
    16 384 envs: no touching contact is ever dropped (speculative ones are shed first), no brick ever loses a pair against a static
    box.  (b) a heap left alone settles with its touching contacts near the 0.5 mm slop.  MEASURED on 512 settled 9-layer heaps (165 k
    touching contacts): median depth 0.55 mm, 0.6 % deeper than 2 mm, deepest 11.6 mm.  Round 2 started at 0.84 mm / 9.9 % / 29 mm; what
    moved it: edge-edge contacts (two bricks crossing edge over edge used to be invisible until a corner reached a face) and resting
    contacts warm-started at 0.98 instead of 0.85.  The bounds below pin that level."""
    from seqdex_b200.env import SdxEnv, make_heap_bank
    from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights
    bank = make_heap_bank(scene, 8)
    g = SdxEnv(scene, FULL, 0, 22)
    g.set_tvalue_weights(default_tvalue_weights(1))
    g.set_heap_bank(bank)
    gen = torch.Generator(device="cuda").manual_seed(3)
    g.step(torch.rand(FULL, 23, device="cuda", generator=gen) * 2 - 1)
    g.tensor("PROGRESS").copy_(torch.randint(0, 75, (FULL,), device="cuda", generator=gen))
    nc = g.tensor("NCONTACT")
    worst = torch.zeros(4, dtype=torch.int64, device="cuda")
    shed_envs = 0.0
    for t in range(300):
        g.step(torch.rand(FULL, 23, device="cuda", generator=gen) * 2 - 1)
        worst[0] = torch.maximum(worst[0], nc[:, 1].max())                  # contacts dropped beyond the table after shedding
        worst[1] = torch.maximum(worst[1], (nc[:, 3] >> 16).max())          # candidate pairs against statics dropped
        worst[2] = torch.maximum(worst[2], nc[:, 2].max())                  # deepest shedding level used
        worst[3] = torch.maximum(worst[3], (nc[:, 3] & 0xFFFF).max())       # candidate pairs dropped (dynamic targets only)
        shed_envs += float((nc[:, 2] > 0).float().mean())
    torch.cuda.synchronize()
    w = worst.tolist()
    assert w[0] == 0, f"touching contacts were dropped: {w}"
    assert w[1] == 0, f"a brick lost a pair against a static box: {w}"
    assert shed_envs / 300 < 0.02, (shed_envs / 300, w)                    # shedding is a tail event (the hand ploughing through a fresh heap)
    assert torch.isfinite(g.tensor("BRICK")).all()
    # (b) settled heaps: 512 envs restored from the bank, no robot motion, 150 steps; depth = column 4 of the contact dump.  Sleeping is
    #     switched off for this part: a sleeping heap has no contacts left to look at (asleep-vs-asleep / static pairs are not generated)
    from seqdex_b200.tasks.cfg import scene_from_cfg
    h = SdxEnv(scene_from_cfg("BlockAssemblyGraspSim", sleep_time=0.0), 512, 0, 22)
    h.set_heap_bank(bank)
    h.set_tvalue_weights(default_tvalue_weights(1))
    h.step(torch.zeros(512, 23, device="cuda"))                             # first step: every env restores a banked heap
    con = h.tensor("CONTACTS")
    h.simulate(150)
    torch.cuda.synchronize()
    ncon = h.tensor("NCONTACT")[:, 0]
    live = torch.arange(1024, device="cuda")[None, :] < ncon[:, None]
    depth = con[..., 4][live]
    touching = depth[depth > 0]
    assert touching.numel() > 10000
    frac_deep = float((touching > 2e-3).float().mean())
    print(f"settled heaps: {touching.numel()} touching contacts, median depth {float(touching.median()) * 1e3:.3f} mm, "
          f"> 2 mm: {100 * frac_deep:.3f} %, max {float(touching.max()) * 1e3:.2f} mm")
    assert float(touching.median()) < 0.8e-3 and frac_deep < 0.02 and float(touching.max()) < 0.03, (float(touching.median()), frac_deep, float(touching.max()))
