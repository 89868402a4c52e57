"""ctypes front-end of the CPU ORACLE (oracle/sdx_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  ``OracleEnv`` holds the same state arrays, in the same layouts, as the CUDA env
(include/seqdex_b200.h tensor kinds) and runs BaseTask.step's three phases (BT:130-150) on the host.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NB, NL, ND, MAXC = 72, 24, 23, 1024
F32P = ctypes.POINTER(ctypes.c_float)
I32P = ctypes.POINTER(ctypes.c_int)
I64P = ctypes.POINTER(ctypes.c_int64)


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsdx_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.sdxo_tvalue_one.restype = ctypes.c_float
    return _LIB


def fp(a):
    return a.ctypes.data_as(F32P) if a is not None else None


def ip(a):
    return a.ctypes.data_as(I32P)


def lp(a):
    return a.ctypes.data_as(I64P)


def default_tvalue_weights(seed=0):
    """torch-default init of GraspInsertTValue(4, 2) (TVF:30-46) as one flat f32 vector."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = []
    for (o, i) in ((256, 4), (128, 256), (64, 128), (2, 64)):
        bound = 1.0 / np.sqrt(i)
        w = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        b = (torch.rand(o, generator=g) * 2 - 1) * bound
        out += [w.reshape(-1), b]
    return torch.cat(out).numpy().astype(np.float32)


class OracleEnv:
    def __init__(self, scene, num_envs, seed=22, tvalue_weights=None):
        self.L = lib()
        assert self.L.sdxo_scene_size() == ctypes.sizeof(scene.c), "scene struct ABI mismatch"
        self.scene = scene
        self.S = ctypes.byref(scene.c)
        self.n = n = num_envs
        self.seed = seed
        self.brick = np.zeros((n, 13, NB), np.float32)
        self.dof = np.zeros((n, 3, 24), np.float32)
        self.link = np.zeros((n, NL, 13), np.float32)
        self.jac7 = np.zeros((n, 6, 7), np.float32)
        self.netf = np.zeros((n, NL, 3), np.float32)
        self.actions = np.zeros((n, 23), np.float32)
        self.task = int(scene.c.task)               # 0 BlockAssemblyGraspSim, 1 BlockAssemblyOrient
        self.obs = np.zeros((n, {0: 396, 3: 75, 4: 468, 5: 468}.get(self.task, 186)), np.float32)
        if self.task in (4, 5):                     # ToolPositioningGrasp / Orient: plate pose, success_buf, camera-frame tool quaternion, banked grasps
            self.plate = np.zeros((n, 7), np.float32)
            self.success_buf = np.zeros((n, 2), np.float32)
            self.qcam = np.zeros((n, 4), np.float32)
            self.grasp_obj = self.grasp_hand = None
            self.pitch_k = None                     # test hooks: pitch index of the next reset_idx calls, yaw draw / bank slot per env
            self.yaw_u = None
        if self.task == 3:                          # BlockAssemblyInsertSim: base-plate pose, wrist orientation error, success_buf, banked grasps
            self.plate = np.zeros((n, 7), np.float32)   # written by the first reset_idx (every env starts with its reset flag set, BT:63)
            self.rot_err = np.zeros((n, 3), np.float32)
            self.success_buf = np.zeros((n, 2), np.float32)
            self.grasp_obj = self.grasp_hand = None
            self.plate_yaw = None
        if self.task == 2:                          # BlockAssemblySearch: camera features, emergence reward, the gate's 10-frame input
            self.seg = np.zeros((n, 3), np.int32)
            self.emergence = np.zeros(n, np.float32)
            self.last_pixels = np.zeros(n, np.float32)
            self.tvobs = np.zeros((n, 650), np.float32)
            self.cam = None
            self.sb_wrap = 0
        self.states = np.zeros((n, 188 if self.task == 3 else 564), np.float32)
        self.rew = np.zeros(n, np.float32)
        self.reset = np.ones(n, np.int64)           # BT:63
        self.progress = np.zeros(n, np.int64)
        self.tvalue = np.zeros(n, np.float32)
        self.finger_dist = np.zeros(n, np.float32)
        self.target_init = np.zeros((n, 7), np.float32)
        self.successes = np.zeros(n, np.float32)
        self.consec = np.zeros(1, np.float32)
        self.ncontact = np.zeros((n, 4), np.int32)   # contacts | beyond the table | shed level | candidate pairs beyond KC (hi 16: vs statics)
        self.episode = np.zeros(n, np.int32)
        self.condump = np.zeros((n, MAXC, 8), np.float32)
        self.ws = np.zeros((n, 2, MAXC, 4), np.float32)
        self.wsn = np.zeros((n, 2), np.int32)
        self.slp = np.zeros((n, NB), np.uint8)      # sleep counters (sub-steps since last hot)
        self.ws_cur = 0
        self.gb_hand = np.zeros((8, 11024, 23, 2), np.float32)
        self.gb_obj = np.zeros((8, 11024, 13), np.float32)
        self.gb_index = np.zeros(8, np.int32)
        self.total_steps = 0
        self.tv = (tvalue_weights if tvalue_weights is not None else default_tvalue_weights()).astype(np.float32)
        self.bank = None
        self.per_type = 0
        self.reset_all()

    # ---- state helpers
    def reset_all(self):
        c = self.scene.c
        rows = np.ctypeslib.as_array(c.brick_init).reshape(NB, 13).astype(np.float32)
        self.set_brick_roots(np.broadcast_to(rows, (self.n, NB, 13)).copy())
        lo, hi = self.scene.dof_lo, self.scene.dof_hi
        q = np.zeros(24, np.float32)
        q[:7] = np.ctypeslib.as_array(c.prepare_arm)
        fr = np.ctypeslib.as_array(c.finger_reset_unscaled).astype(np.float32)
        q[7:23] = (np.float32(0.5) * (fr + np.float32(1.0)) * (hi[7:] - lo[7:]) + lo[7:]).astype(np.float32)
        self.dof[:, 0, :] = q
        self.dof[:, 1, :] = 0
        self.dof[:, 2, :] = q
        self.progress[:] = 0
        self.reset[:] = 1
        self.wsn[:] = 0
        self.slp[:] = 0
        self.refresh_links()

    def set_brick_roots(self, rows):
        rows = np.ascontiguousarray(rows, np.float32)
        self.L.sdxo_brick_from_root_rows(self.S, self.n, fp(self.brick), fp(rows))
        self.slp[:] = 0      # setting a pose wakes the actor (sdx_set_actor_root_state_indexed does the same)

    def brick_roots(self):
        rows = np.zeros((self.n, NB, 13), np.float32)
        self.L.sdxo_brick_root_rows(self.S, self.n, fp(self.brick), fp(rows))
        return rows

    def refresh_links(self):
        self.L.sdxo_refresh_links(self.S, self.n, fp(self.dof), fp(self.link), fp(self.jac7))

    def set_heap_bank(self, bank):
        self.bank = np.ascontiguousarray(bank, np.float32)
        self.per_type = bank.shape[1]

    def segmentation_features(self, cam):
        """[n, 3] int32: pixels showing the target brick, int(mean row), int(mean column) (SE:1231-1241)"""
        out = np.zeros((self.n, 3), np.int32)
        self.L.sdxo_segmentation_features(self.S, self.n, ctypes.byref(cam), fp(self.brick), fp(self.link), ip(out))
        return out

    # ---- BaseTask.step phases
    def simulate(self, dump=False):
        self.L.sdxo_simulate(self.S, self.n, fp(self.brick), fp(self.dof), fp(self.link), fp(self.jac7), fp(self.netf),
                             ip(self.ncontact), fp(self.condump) if dump else None, fp(self.ws), ip(self.wsn), self.ws_cur,
                             self.slp.ctypes.data_as(ctypes.c_void_p))
        self.ws_cur ^= self.scene.c.substeps & 1

    def enable_tvalue_dataset(self, cap):
        self.tvd_cap = int(cap)
        self.tvd_succ = np.zeros((cap, 4), np.float32)
        self.tvd_fail = np.zeros((cap, 4), np.float32)
        self.tvd_counts = np.zeros(2, np.int64)

    def enable_orient_heap_bank(self, cap):
        self.ob_wrap = int(cap)
        self.ob_rows = np.zeros((8, cap + 1, NB, 13), np.float32)
        self.ob_index = np.zeros(8, np.int32)

    # ---- BlockAssemblySearch (SE:1274-1537)
    def set_camera(self, cam):
        self.cam = cam

    def enable_search_bank(self, cap):
        self.sb_wrap = int(cap)
        self.sb_rows = np.zeros((8, cap + 1, NB, 13), np.float32)
        self.sb_hand = np.zeros((8, cap + 1, 23, 2), np.float32)
        self.sb_index = np.zeros(8, np.int32)

    def _search_render(self, baseline):
        """render_all_camera_sensors + compute_emergence_reward (SE:1446-1455 / 1010-1019)"""
        assert self.cam is not None, "BlockAssemblySearch needs its overview camera (SE:873-878)"
        self.seg[:] = self.segmentation_features(self.cam)
        self.L.sdxo_search_emergence(self.n, ip(self.seg), fp(self.last_pixels), fp(self.emergence), int(baseline))

    def _search_reset_idx(self):
        L, vp = self.L, self.slp.ctypes.data_as(ctypes.c_void_p)
        state = lambda phase: L.sdxo_search_reset(self.S, self.n, ctypes.c_uint64(self.seed), phase, fp(self.brick), fp(self.dof),
                                                  fp(self.target_init), lp(self.progress), lp(self.reset), fp(self.successes),
                                                  ip(self.episode), ip(self.wsn), vp)
        self.last_reset_sim_steps = 0
        if self.total_steps > 0 and self.sb_wrap:
            L.sdxo_search_bank(self.S, self.n, fp(self.brick), fp(self.dof), ip(self.seg), fp(self.sb_rows), fp(self.sb_hand),
                               ip(self.sb_index), self.sb_wrap)
        state(0)
        for _ in range(60):                       # SE:1437-1439: the heap falls into the bin
            self.simulate(); self.last_reset_sim_steps += 1
        self._search_render(True)
        state(1)
        self.refresh_links()                      # the hand was teleported and the next reader (pre_physics) comes before any contact step
        state(2)

    def _search_post(self):
        if self.progress[0] + 1 >= self.scene.c.max_episode_length - 1:      # SE:989: env 0's clock stands for all (lockstep)
            self.L.sdxo_search_hand_pose(self.S, self.n, None, 0, fp(self.dof))
            self.simulate()
            self._search_render(False)
        self.L.sdxo_search_post_physics(self.S, self.n, fp(self.brick), fp(self.dof), fp(self.link), fp(self.netf), fp(self.actions),
                                        fp(self.target_init), ip(self.seg), lp(self.progress), lp(self.reset), fp(self.obs),
                                        fp(self.states), fp(self.tvobs), fp(self.rew), fp(self.finger_dist), fp(self.successes),
                                        fp(self.consec))

    def _orient_reset_idx(self):
        """OR:1390-1695: the scripted reset (lift 50, observe + bank, state reset, settle 2 + 1, approach 50)"""
        L, vp = self.L, self.slp.ctypes.data_as(ctypes.c_void_p)
        script = lambda mode, i: L.sdxo_orient_arm_script(self.S, self.n, lp(self.reset), mode, i, fp(self.dof), fp(self.link),
                                                          fp(self.jac7), fp(self.brick), fp(self.target_init))
        state = lambda phase: L.sdxo_orient_reset(self.S, self.n, ctypes.c_uint64(self.seed), fp(self.bank), self.per_type, phase,
                                                  fp(self.brick), fp(self.dof), fp(self.target_init), lp(self.progress), lp(self.reset),
                                                  fp(self.successes), ip(self.episode), ip(self.wsn), vp)
        self.last_reset_sim_steps = 0
        def sim():
            self.simulate(); self.last_reset_sim_steps += 1
        if self.total_steps > 0:
            for i in range(50):
                script(0, i); sim()
            self._orient_post(0)
            if getattr(self, "ob_wrap", 0):
                L.sdxo_orient_bank(self.S, self.n, fp(self.brick), fp(self.finger_dist), fp(self.tvalue), fp(self.ob_rows),
                                   ip(self.ob_index), self.ob_wrap)
        state(0)
        sim(); sim()                  # OR:1617-1619 (every reader of the link rows below comes after a contact step, which rewrites them)
        state(1)
        sim()                         # OR:1657
        for i in range(50):
            script(1, i); sim()
        state(2)

    def _orient_post(self, count_step):
        self.L.sdxo_orient_post_physics(self.S, self.n, fp(self.tv), fp(self.brick), fp(self.dof), fp(self.link), fp(self.actions),
                                        fp(self.target_init), lp(self.progress), lp(self.reset), fp(self.obs), fp(self.states),
                                        fp(self.rew), fp(self.tvalue), fp(self.finger_dist), fp(self.successes), fp(self.consec),
                                        int(count_step))

    # ---- BlockAssemblyInsertSim (IS:1328-1565)
    def set_grasp_bank(self, hand, obj):
        """hand [8, K, 23, 2], obj [8, K, 13]: saved_grasping_{hand,object}_ternimal_states (IS:372-375)"""
        self.grasp_hand = np.ascontiguousarray(hand, np.float32)
        self.grasp_obj = np.ascontiguousarray(obj, np.float32).reshape(8, -1, 13)

    def plate_yaw_draw(self):
        """random.sample([0, 1], 1), ONE draw per reset_idx call (IS:1435): Philox(seed, total_steps) here"""
        if self.plate_yaw is not None:
            return int(self.plate_yaw)
        return int(dr_philox_bit(self.seed, self.total_steps))

    def _insert_reset_idx(self, slots=None):
        """slots (test hook): the bank slot per ENV, or one per resetting env in env order (what the golden generator recorded)"""
        assert self.grasp_obj is not None, "reset needs the banked grasps (IS:372-375)"
        if slots is None:
            slots = getattr(self, "slot_by_env", None)
        so = None
        if slots is not None:
            so = np.zeros(self.n, np.int32)
            if len(slots) == self.n:
                so[:] = slots
            else:
                so[np.flatnonzero(self.reset)] = slots
        self.L.sdxo_insert_reset(self.S, self.n, ctypes.c_uint64(self.seed), fp(self.grasp_obj), fp(self.grasp_hand), int(self.grasp_obj.shape[1]),
                                 self.plate_yaw_draw(), ip(so) if so is not None else None, int(self.total_steps > 0), fp(self.brick), fp(self.dof),
                                 fp(self.plate), fp(self.target_init), lp(self.progress), lp(self.reset), fp(self.successes), fp(self.success_buf),
                                 ip(self.episode), ip(self.wsn), self.slp.ctypes.data_as(ctypes.c_void_p))
        self.refresh_links()                      # the hand was teleported: pre_physics reads its pose before any contact step

    # ---- ToolPositioningGrasp / ToolPositioningOrient (TG:1412-1675, TO:1265-1509)
    def pitch_draw(self):
        """random.sample(range(4), 1), ONE draw per reset_idx call (TG:1489): Philox(seed, total_steps) here, as in csrc/sdx_env.cu"""
        if self.pitch_k is not None:
            return int(self.pitch_k)
        out = (ctypes.c_uint32 * 4)()
        self.L.sdxo_philox(ctypes.c_uint64(self.seed), ctypes.c_uint32(self.total_steps & 0xFFFFFFFF), ctypes.c_uint32(0xC0FFEE), ctypes.c_uint32(7), out)
        return int(out[0] & 3)

    def _tool_reset_idx(self):
        orient = int(self.task == 5)
        if orient:
            assert self.grasp_obj is not None, "reset needs the banked grasps (TO:365-368)"
        elif self.total_steps > 0:
            self.L.sdxo_tool_bank(self.S, self.n, fp(self.brick), fp(self.dof), lp(self.reset), fp(self.finger_dist), fp(self.plate),
                                  fp(self.gb_hand), fp(self.gb_obj), ip(self.gb_index))
        slots = getattr(self, "slot_by_env", None)
        so = np.ascontiguousarray(slots, np.int32) if slots is not None else None
        yu = np.ascontiguousarray(self.yaw_u, np.float32) if self.yaw_u is not None else None
        self.L.sdxo_tool_reset(self.S, self.n, orient, ctypes.c_uint64(self.seed), fp(self.grasp_obj), fp(self.grasp_hand),
                               int(self.grasp_obj.shape[1]) if orient else 0, self.pitch_draw(), ip(so) if so is not None else None, fp(yu),
                               int(self.total_steps > 0), fp(self.brick), fp(self.dof), fp(self.plate), fp(self.target_init), lp(self.progress),
                               lp(self.reset), fp(self.successes), fp(self.success_buf), ip(self.episode), ip(self.wsn),
                               self.slp.ctypes.data_as(ctypes.c_void_p), fp(self.obs), fp(self.states))
        self.refresh_links()                      # the hand was teleported: pre_physics reads its pose before any contact step

    def tool_insertion_obs(self, ins_actions, ins_progress, ins_obs, ins_max_len=125):
        """ToolPositioningChain.compute_insertion_observations (TC:1404-1440) into the caller's [n, 468] buffer"""
        a = np.ascontiguousarray(ins_actions, np.float32)
        p = np.ascontiguousarray(ins_progress, np.int64)
        self.L.sdxo_tool_insertion_obs(self.n, fp(self.obs), fp(a), lp(p), int(ins_max_len), fp(ins_obs))

    def tool_inner_step(self, actions):
        """one step of ToolPositioningChain's inner loop (TC:1733-1768): fingers from the actions, arm holds, contact step"""
        a = np.ascontiguousarray(np.clip(actions, -1.0, 1.0), np.float32)
        scratch = np.zeros_like(self.actions)
        self.L.sdxo_tool_pre_physics(self.S, self.n, 1, fp(a), fp(scratch), fp(self.dof), fp(self.link), fp(self.jac7), lp(self.progress))
        self.simulate()

    def tool_tvalue_labels(self):
        """TO:1305-1316: success_buf for ALL envs from the current state; returns the class index per env (0 success, 1 failure)"""
        label = np.zeros(self.n, np.int32)
        self.L.sdxo_tool_tvalue_labels(self.S, self.n, fp(self.brick), fp(self.plate), fp(self.success_buf), ip(label))
        return label

    def pre_physics(self, actions):
        if self.task in (4, 5):
            if self.reset.any():
                self._tool_reset_idx()
            a = np.ascontiguousarray(np.clip(actions, -1.0, 1.0), np.float32)   # VR:166
            self.L.sdxo_tool_pre_physics(self.S, self.n, int(self.task == 5), fp(a), fp(self.actions), fp(self.dof), fp(self.link), fp(self.jac7),
                                         lp(self.progress))
            return
        if self.task == 3:
            if self.reset.any():
                self._insert_reset_idx()
            a = np.ascontiguousarray(np.clip(actions, -1.0, 1.0), np.float32)   # VR:166
            self.L.sdxo_insert_pre_physics(self.S, self.n, fp(a), fp(self.actions), fp(self.dof), fp(self.link), fp(self.jac7), fp(self.rot_err))
            return
        if self.task == 2:
            self.last_reset_sim_steps = 0
            if self.reset.any():
                self._search_reset_idx()
            a = np.ascontiguousarray(np.clip(actions, -1.0, 1.0), np.float32)   # VR:166
            self.L.sdxo_search_pre_physics(self.S, self.n, fp(a), fp(self.actions), fp(self.dof), fp(self.link), fp(self.jac7),
                                           fp(self.brick))
            return
        if self.task == 1:
            self.last_reset_sim_steps = 0
            if self.reset.any():
                assert self.bank is not None, "reset needs a heap bank (OR:419-420)"
                self._orient_reset_idx()
            a = np.ascontiguousarray(np.clip(actions, -1.0, 1.0), np.float32)   # VR:166
            self.L.sdxo_orient_pre_physics(self.S, self.n, fp(a), fp(self.actions), fp(self.dof), fp(self.link), fp(self.jac7),
                                           fp(self.brick), lp(self.progress), fp(self.target_init))
            return
        if self.reset.any():
            assert self.bank is not None, "reset needs a heap bank (GS:412-413)"
            if getattr(self, "tvd_cap", 0) and self.total_steps > 0:
                self.L.sdxo_tv_dataset(self.S, self.n, lp(self.reset), fp(self.brick), fp(self.finger_dist), fp(self.tvalue),
                                       fp(self.states), fp(self.tvd_succ), fp(self.tvd_fail), lp(self.tvd_counts), self.tvd_cap)
            self.L.sdxo_reset(self.S, self.n, ctypes.c_uint64(self.seed), fp(self.bank), self.per_type, fp(self.brick),
                              fp(self.dof), fp(self.target_init), lp(self.progress), lp(self.reset), fp(self.successes),
                              ip(self.episode), ip(self.wsn), self.slp.ctypes.data_as(ctypes.c_void_p), int(self.total_steps > 0), fp(self.finger_dist), fp(self.tvalue),
                              fp(self.gb_hand), fp(self.gb_obj), ip(self.gb_index))
        a = np.ascontiguousarray(np.clip(actions, -1.0, 1.0), np.float32)   # VR:166
        self.L.sdxo_pre_physics(self.S, self.n, fp(a), fp(self.actions), fp(self.dof), fp(self.link), fp(self.jac7),
                                lp(self.progress), fp(self.target_init))

    def post_physics(self):
        if self.task in (4, 5):
            self.L.sdxo_tool_post_physics(self.S, self.n, int(self.task == 5), fp(self.brick), fp(self.dof), fp(self.link), fp(self.actions),
                                          fp(self.target_init), fp(self.plate), lp(self.progress), lp(self.reset), fp(self.obs), fp(self.states),
                                          fp(self.rew), fp(self.qcam), fp(self.finger_dist), fp(self.successes), fp(self.consec))
            self.total_steps += 1
            return
        if self.task == 3:
            self.L.sdxo_insert_post_physics(self.S, self.n, fp(self.brick), fp(self.dof), fp(self.link), fp(self.actions), fp(self.target_init),
                                            fp(self.plate), fp(self.rot_err), lp(self.progress), lp(self.reset), fp(self.obs), fp(self.states),
                                            fp(self.rew), fp(self.finger_dist), fp(self.successes), fp(self.consec))
            self.total_steps += 1
            return
        if self.task == 2:
            self._search_post()
            self.total_steps += 1
            return
        if self.task == 1:
            self._orient_post(1)
            self.total_steps += 1
            return
        self.L.sdxo_post_physics(self.S, self.n, fp(self.tv), fp(self.brick), fp(self.dof), fp(self.link), fp(self.actions),
                                 fp(self.target_init), lp(self.progress), lp(self.reset), fp(self.obs), fp(self.states),
                                 fp(self.rew), fp(self.tvalue), fp(self.finger_dist), fp(self.successes), fp(self.consec))
        self.total_steps += 1

    def step(self, actions):
        self.pre_physics(actions)
        self.simulate()
        self.post_physics()
        return (np.clip(self.obs, -5, 5), np.clip(self.states, -5, 5), self.rew, self.reset)   # VR:171-177


def dr_philox_bit(seed, counter):
    """low bit of Philox4x32-10(seed; counter, 0xC0FFEE, 7): the base-plate yaw draw of InsertSim's reset_idx (same stream in csrc/sdx_env.cu)"""
    out = (ctypes.c_uint32 * 4)()
    lib().sdxo_philox(ctypes.c_uint64(seed), ctypes.c_uint32(counter & 0xFFFFFFFF), ctypes.c_uint32(0xC0FFEE), ctypes.c_uint32(7), out)
    return out[0] & 1


def gae(rewards, values, dones, last_values, last_dones, gamma, tau):
    H, n = rewards.shape
    adv = np.zeros((H, n), np.float32)
    ret = np.zeros((H, n), np.float32)
    lib().sdxo_gae(fp(np.ascontiguousarray(rewards, np.float32)), fp(np.ascontiguousarray(values, np.float32)),
                   fp(np.ascontiguousarray(dones, np.float32)), fp(np.ascontiguousarray(last_values, np.float32)),
                   fp(np.ascontiguousarray(last_dones, np.float32)), fp(adv), fp(ret), H, n,
                   ctypes.c_float(gamma), ctypes.c_float(tau))
    return adv, ret
