"""Thin Python owner of one ``sdx_env_t`` (include/seqdex_b200.h): device buffers exposed as
zero-copy torch views, the way ``gymtorch.wrap_tensor`` exposes PhysX buffers (GS:237-246)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .scene import Scene

T = dict(BRICK=0, DOF=1, LINK=2, JAC7=3, NETF=4, ACTIONS=5, OBS=6, STATES=7, REW=8, RESET=9, PROGRESS=10, TVALUE=11,
         TARGET_INIT=12, SUCCESSES=13, CONSEC=14, NCONTACT=15, ROOT=16, RB=17, DOF_STATE=18, JACOBIAN=19, EPISODE=20,
         CONTACTS=21, WS=22, WSN=23, SLEEP=24, SEG=25, EMERGENCE=26, TVOBS=27, PLATE=28, ROT_ERR=29, SUCCESS=30)
_DT = {0: (torch.float32, "<f4"), 1: (torch.int64, "<i8"), 2: (torch.int32, "<i4"), 3: (torch.uint8, "|u1")}


class _DevView:
    """``__cuda_array_interface__`` carrier so torch can alias a raw device pointer without copying."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class SdxEnv:
    def __init__(self, scene: Scene, num_envs: int, device: int = 0, seed: int = 22):
        if not torch.cuda.is_available():
            raise RuntimeError("seqdex_b200 needs a CUDA device: there is no CPU path")
        self.L = _lib.load()
        assert self.L.sdx_scene_size() == ctypes.sizeof(scene.c), "scene struct ABI mismatch"
        self.scene, self.n, self.device_index = scene, num_envs, device
        self.device = torch.device("cuda", device)
        self.h = ctypes.c_void_p()
        _lib.check(self.L.sdx_create(ctypes.byref(scene.c), num_envs, device, ctypes.c_uint64(seed), ctypes.byref(self.h)))
        rows = np.zeros((142, 13), np.float32)
        for k, v in scene.static_actor_roots().items():
            rows[k] = v
        _lib.check(self.L.sdx_set_static_rows(self.h, rows.ctypes.data_as(ctypes.c_void_p)))
        self._views = {}
        self.set_stream(torch.cuda.current_stream(self.device))

    def close(self):
        if self.h:
            self.L.sdx_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        _lib.check(self.L.sdx_set_stream(self.h, ctypes.c_void_p(stream.cuda_stream)))

    def tensor(self, name: str) -> torch.Tensor:
        if name in self._views:
            return self._views[name]
        ptr = ctypes.c_void_p()
        shape = (ctypes.c_int64 * 4)()
        nd, dt = ctypes.c_int(), ctypes.c_int()
        _lib.check(self.L.sdx_tensor(self.h, T[name], ctypes.byref(ptr), shape, ctypes.byref(nd), ctypes.byref(dt)))
        shp = [int(shape[i]) for i in range(nd.value)]
        with torch.cuda.device(self.device):
            t = torch.as_tensor(_DevView(ptr.value, shp, _DT[dt.value][1]), device=self.device)
        self._views[name] = t
        return t

    # ---- setup
    def set_heap_bank(self, bank):
        """bank: [8, per_type, 72, 13] root-frame rows (numpy or torch)."""
        if isinstance(bank, torch.Tensor) and bank.is_cuda:
            b = bank.contiguous().float()
            _lib.check(self.L.sdx_set_heap_bank_dev(self.h, ctypes.c_void_p(b.data_ptr()), int(b.shape[1])))
        else:
            b = np.ascontiguousarray(bank.cpu().numpy() if isinstance(bank, torch.Tensor) else bank, np.float32)
            _lib.check(self.L.sdx_set_heap_bank(self.h, b.ctypes.data_as(ctypes.c_void_p), int(b.shape[1])))

    def set_tvalue_weights(self, w):
        w = np.ascontiguousarray(w, np.float32)
        assert w.size == 42562
        _lib.check(self.L.sdx_set_tvalue_weights(self.h, w.ctypes.data_as(ctypes.c_void_p)))

    def reset_all(self):
        _lib.check(self.L.sdx_reset_all(self.h))

    # ---- stepping (BT:130-150)
    def pre_physics(self, actions: torch.Tensor):
        a = actions.contiguous().float()
        _lib.check(self.L.sdx_pre_physics(self.h, ctypes.c_void_p(a.data_ptr())))

    def simulate(self, n=1):
        _lib.check(self.L.sdx_simulate_n(self.h, n))

    def post_physics(self):
        _lib.check(self.L.sdx_post_physics(self.h))

    def step(self, actions: torch.Tensor):
        a = actions.contiguous().float()
        _lib.check(self.L.sdx_step(self.h, ctypes.c_void_p(a.data_ptr())))

    def step_host(self, actions, obs, states, rew, reset):
        """pinned host tensors in/out (VecTask.step through the C-ABI with host buffers)."""
        _lib.check(self.L.sdx_step_host(self.h, ctypes.c_void_p(actions.data_ptr()), ctypes.c_void_p(obs.data_ptr()),
                                        ctypes.c_void_p(states.data_ptr()), ctypes.c_void_p(rew.data_ptr()),
                                        ctypes.c_void_p(reset.data_ptr())))

    def clamped_copy(self, name, dst, lim=5.0):
        """dst <- clamp(tensor(name), -lim, lim) in one kernel (VecTask's clip_obs, VR:174-175)"""
        _lib.check(self.L.sdx_clamped_copy(self.h, T[name], ctypes.c_void_p(dst.data_ptr()), ctypes.c_float(lim)))

    def refresh(self, name):
        _lib.check(self.L.sdx_refresh(self.h, T[name]))

    def launch_count(self):
        return int(self.L.sdx_launch_count(self.h))

    def _view(self, ptr, shape, typestr="<f4"):
        with torch.cuda.device(self.device):
            return torch.as_tensor(_DevView(ptr, shape, typestr), device=self.device)

    def grasp_bank(self):
        """the grasp terminal-state rings reset_idx fills (GS:1399-1445): (hand [8,11024,23,2], obj [8,11024,13], index [8])"""
        h, o, i = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(self.L.sdx_grasp_bank(self.h, ctypes.byref(h), ctypes.byref(o), ctypes.byref(i)))
        return self._view(h.value, (8, 11024, 23, 2)), self._view(o.value, (8, 11024, 13)), self._view(i.value, (8,), "<i4")

    def aux(self):
        """(camera-frame target quaternion [N, 4], arm_hand_finger_dist [N]) of the last compute_observations, as device views"""
        q, f = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(self.L.sdx_aux(self.h, ctypes.byref(q), ctypes.byref(f)))
        return self._view(q.value, (self.n, 4)), self._view(f.value, (self.n,))

    def enable_tvalue_dataset(self, capacity=65536):
        """record the t-value training rows on the device (the reference's save_hdf5 branch, GS:1402-1438)"""
        _lib.check(self.L.sdx_tvalue_dataset(self.h, int(capacity), None, None, None))
        self._tvd_cap = int(capacity)

    def tvalue_dataset(self):
        """(success [ns,4], failure [nf,4], counts [2] i64): rows recorded so far (ring order once a ring has wrapped)"""
        s, f, c = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(self.L.sdx_tvalue_dataset(self.h, self._tvd_cap, ctypes.byref(s), ctypes.byref(f), ctypes.byref(c)))
        counts = self._view(c.value, (2,), "<i8")
        torch.cuda.synchronize(self.device)
        ns, nf = (min(int(x), self._tvd_cap) for x in counts.tolist())
        return self._view(s.value, (self._tvd_cap, 4))[:ns], self._view(f.value, (self._tvd_cap, 4))[:nf], counts

    def enable_orient_heap_bank(self, capacity=10000):
        """record the heaps BlockAssemblyOrient leaves face up (saved_digging_ternimal_states_list, OR:1465-1481); capacity is the
        slot after which the ring index returns to 0 (10000 in the reference)"""
        _lib.check(self.L.sdx_orient_heap_bank(self.h, int(capacity), None, None))
        self._ob_wrap = int(capacity)

    def orient_heap_bank(self):
        """(rows [8, capacity + 1, 72, 13], index [8] i32) of the re-oriented heap rings"""
        r, i = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(self.L.sdx_orient_heap_bank(self.h, self._ob_wrap, ctypes.byref(r), ctypes.byref(i)))
        return self._view(r.value, (8, self._ob_wrap + 1, 72, 13)), self._view(i.value, (8,), "<i4")

    # ---- BlockAssemblySearch
    def set_camera(self, cam):
        """the overview camera Search renders after every reset and at the end of every episode (SE:873-878)"""
        _lib.check(self.L.sdx_set_camera(self.h, ctypes.byref(cam)))

    def enable_search_bank(self, capacity=10000):
        """record the heaps (and hand states) Search digs the target out of (saved_searching_*_ternimal_states_list, SE:1305-1352)"""
        _lib.check(self.L.sdx_search_bank(self.h, int(capacity), None, None, None))
        self._sb_wrap = int(capacity)

    def search_bank(self):
        """(rows [8, capacity + 1, 72, 13], hand [8, capacity + 1, 23, 2], index [8] i32)"""
        r, h, i = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(self.L.sdx_search_bank(self.h, self._sb_wrap, ctypes.byref(r), ctypes.byref(h), ctypes.byref(i)))
        return (self._view(r.value, (8, self._sb_wrap + 1, 72, 13)), self._view(h.value, (8, self._sb_wrap + 1, 23, 2)),
                self._view(i.value, (8,), "<i4"))

    # ---- BlockAssemblyInsertSim
    def set_grasp_bank(self, hand, obj):
        """the banked grasps InsertSim's reset_idx restores (IS:372-375, 1449-1453): hand [8, K, 23, 2], obj [8, K, 13] (or [8, K, 1, 13]);
        numpy or torch, host or device (GraspSim's rings hand over on the device)"""
        if isinstance(hand, torch.Tensor) and hand.is_cuda:
            h, o = hand.contiguous().float(), obj.contiguous().float().reshape(8, -1, 13)
            _lib.check(self.L.sdx_set_grasp_bank(self.h, ctypes.c_void_p(h.data_ptr()), ctypes.c_void_p(o.data_ptr()), int(h.shape[1]), 1))
        else:
            h = np.ascontiguousarray(hand.cpu().numpy() if isinstance(hand, torch.Tensor) else hand, np.float32)
            o = np.ascontiguousarray(obj.cpu().numpy() if isinstance(obj, torch.Tensor) else obj, np.float32).reshape(8, -1, 13)
            _lib.check(self.L.sdx_set_grasp_bank(self.h, h.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p), int(h.shape[1]), 0))

    def insert_test_hooks(self, slot_by_env=None, plate_yaw=-1):
        s = None if slot_by_env is None else np.ascontiguousarray(slot_by_env, np.int32)
        _lib.check(self.L.sdx_insert_test_hooks(self.h, s.ctypes.data_as(ctypes.c_void_p) if s is not None else None, int(plate_yaw)))

    def tool_test_hooks(self, slot_by_env=None, pitch_k=-1, yaw_u=None):
        """parity-test hook of the ToolPositioning tasks (sdx_tool_test_hooks): bank slot per env, pitch index, yaw draw per env"""
        s = None if slot_by_env is None else np.ascontiguousarray(slot_by_env, np.int32)
        u = None if yaw_u is None else np.ascontiguousarray(yaw_u, np.float32)
        _lib.check(self.L.sdx_tool_test_hooks(self.h, s.ctypes.data_as(ctypes.c_void_p) if s is not None else None, int(pitch_k),
                                              u.ctypes.data_as(ctypes.c_void_p) if u is not None else None))

    def tool_inner_step(self, actions):
        """one step of ToolPositioningChain's inner loop (sdx_tool_inner_step; TC:1733-1768)"""
        _lib.check(self.L.sdx_tool_inner_step(self.h, ctypes.c_void_p(actions.data_ptr())))

    def tool_insertion_obs(self, ins_actions, ins_progress, ins_obs, ins_max_len=125):
        """ToolPositioningChain.compute_insertion_observations (sdx_tool_insertion_obs; TC:1404-1440) into ``ins_obs`` [N, 468]"""
        _lib.check(self.L.sdx_tool_insertion_obs(self.h, ctypes.c_void_p(ins_actions.data_ptr()), ctypes.c_void_p(ins_progress.data_ptr()),
                                                 int(ins_max_len), ctypes.c_void_p(ins_obs.data_ptr())))

    def tool_tvalue_labels(self, out=None):
        """ToolPositioningOrient's online t-value labels (TO:1305-1316): int32 [N] on the device, 0 success / 1 failure; also writes SUCCESS"""
        if out is None:
            out = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        _lib.check(self.L.sdx_tool_tvalue_labels(self.h, ctypes.c_void_p(out.data_ptr())))
        return out

    def last_reset_sim_steps(self):
        return int(self.L.sdx_last_reset_sim_steps(self.h))

    def segmentation_features(self, cam, out=None):
        """Search's camera features (SE:1231-1241, 1640-1646) for the camera ``cam`` (seqdex_b200.camera.look_at):
        int32 [N, 3] = pixels that show the target brick, int(mean row), int(mean column)"""
        if out is None:
            out = torch.zeros(self.n, 3, dtype=torch.int32, device=self.device)
        _lib.check(self.L.sdx_segmentation_features(self.h, ctypes.byref(cam), ctypes.c_void_p(out.data_ptr())))
        return out

    def brick_roots(self):
        """[N, 72, 13] Isaac-Gym root rows of the free bricks (actors 9..80 of the root tensor)."""
        self.refresh("ROOT")
        return self.tensor("ROOT").view(self.n, 142, 13)[:, 9:81]


def make_heap_bank(scene: Scene, per_type: int, device: int = 0, settle_steps: int = 240, seed: int = 22):
    """Synthesise the terminal-state heap bank the task samples on reset -- the stand-in for the
    unshipped ``saved_searching_ternimal_states_*.pkl`` (GS:412-413): drop the 72 free
    bricks from the staggered lattice (GS:737-742) under gravity and let them settle; per-heap variety comes
    from a small Philox-free jitter of the lattice.  Runs on the GPU with the product kernel."""
    n = 8 * per_type
    g = torch.Generator(device="cpu").manual_seed(seed)
    env = SdxEnv(scene, n, device, seed)
    brick = env.tensor("BRICK")
    jit = (torch.rand(n, 3, 72, generator=g) * 2 - 1) * torch.tensor([0.01, 0.01, 0.0]).view(1, 3, 1)
    brick[:, 0:3, :] += jit.to(brick.device)
    old = scene.c.brick_lin_damp
    # gentle settle: linear damping while the 9 layers come down, then free
    import ctypes as _c
    scene.c.brick_lin_damp = 10.0
    env2 = env  # the device copy of the scene was taken at creation: recreate with damping
    env.close()
    env = SdxEnv(scene, n, device, seed)
    env.tensor("BRICK")[:, 0:3, :] += jit.to(env.device)
    env.simulate(settle_steps * 2 // 3)
    rows_damped = env.brick_roots().clone()
    scene.c.brick_lin_damp = old
    env3 = SdxEnv(scene, n, device, seed)
    idx = torch.arange(n * 142, dtype=torch.int32, device=env3.device)
    root = env3.tensor("ROOT")
    env3.refresh("ROOT")
    root.view(n, 142, 13)[:, 9:81] = rows_damped
    _lib.check(env3.L.sdx_set_actor_root_state_indexed(env3.h, ctypes.c_void_p(root.data_ptr()), ctypes.c_void_p(idx.data_ptr()), n * 142))
    env3.simulate(settle_steps // 3)
    rows = env3.brick_roots().clone()
    rows[:, :, 7:13] = 0
    env.close(); env3.close()
    return rows.view(8, per_type, 72, 13).contiguous()
