"""scripts/bi_optimization.py's schedule as a function (seqdex_b200/bi_optimization.py): one round at toy sizes -- forward
initialisation of the four BlockAssembly stages with device-resident hand-offs, backward fine-tuning with the t-value trainer."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_one_round_of_the_block_assembly_schedule(tmp_path):
    from seqdex_b200.bi_optimization import STAGES, bi_optimization
    ne = {"BlockAssemblySearch": 32, "BlockAssemblyOrient": 64, "BlockAssemblyGraspSim": 64, "BlockAssemblyInsertSim": 64}
    log = []
    history, state = bi_optimization("BlockAssembly", rounds=1, num_envs=ne, iterations=11, tvalue_rollout=40, work_dir=str(tmp_path),
                                     log=lambda *a: log.append(a[:3]))
    assert [x[1:] for x in log] == [("forward", t) for t in STAGES] + [("backward", t) for t in
                                                                         ("BlockAssemblyInsertSim", "BlockAssemblyGraspSim", "BlockAssemblyOrient")]   # BO:118-128
    rec = history[0]
    for t in STAGES:
        assert rec[f"forward/{t}"] == rec[f"forward/{t}"]                    # finite mean rewards
        assert os.path.exists(os.path.join(tmp_path, t, "nn", t + ".pth"))   # main_rlgames returns this path (BO:107-108)
    assert state.heaps_medium is not None and state.heaps_good is not None and state.grasps is not None
    assert state.heaps_medium.shape[2:] == (72, 13) and state.grasps[0].shape[2:] == (23, 2)
    assert torch.isfinite(state.heaps_good).all()
    # the t-value fit runs only where a stage recorded enough rows of both labels (an untrained policy rarely succeeds): None or an accuracy
    for t in ("BlockAssemblyInsertSim", "BlockAssemblyGraspSim", "BlockAssemblyOrient"):
        v = rec[f"tvalue/{t}"]
        assert v is None or 0.0 <= v <= 1.0
    with pytest.raises(Exception, match="Unrecognized task"):                    # BO:136-138
        bi_optimization("Nonexistent", rounds=1)


def test_one_round_of_the_tool_positioning_schedule(tmp_path):
    """BO:127-134 (BASELINE configs[4] at toy size): Grasp and Orient forward -- Orient starts from what Grasp banked, or from the
    synthetic stand-in grasps where an untrained policy banked none --, Orient backward with its rows recorded, the t-value fit"""
    from seqdex_b200.bi_optimization import TOOL_STAGES, bi_optimization
    ne = {"ToolPositioningGrasp": 64, "ToolPositioningOrient": 64}
    log = []
    history, state = bi_optimization("ToolPositioning", rounds=1, num_envs=ne, iterations=20, tvalue_rollout=40, work_dir=str(tmp_path),
                                     log=lambda *a: log.append(a[:3]))
    assert [x[1:] for x in log] == [("forward", t) for t in TOOL_STAGES] + [("backward", "ToolPositioningOrient")]
    rec = history[0]
    for t in TOOL_STAGES:
        assert rec[f"forward/{t}"] == rec[f"forward/{t}"]
        assert os.path.exists(os.path.join(tmp_path, t, "nn", t + ".pth"))
    hand, obj = state.tool_grasps
    assert hand.shape[0] == 8 and hand.shape[2:] == (23, 2) and obj.shape[2:] == (13,) and torch.isfinite(hand).all() and torch.isfinite(obj).all()
    s, f = state.datasets["ToolPositioningOrient"]                               # 20 iterations x 8 steps > one 125-step episode: every env ended one
    assert s.shape[1] == 4 and f.shape[1] == 4 and len(s) + len(f) >= 64
    n = torch.cat([s, f]).norm(dim=-1)
    assert torch.allclose(n, torch.ones_like(n), atol=1e-4)                      # rows are the start poses' unit quaternions
    v = rec["tvalue/ToolPositioningOrient"]
    assert v is None or 0.0 <= v <= 1.0
