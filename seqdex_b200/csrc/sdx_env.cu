// sdx_env.cu -- C-ABI (include/seqdex_b200.h) over the contact-step and task kernels.
// One sdx_env_t = one Isaac Gym "sim" with N envs on one GPU (BT:122-126).  No CPU fallback: every
// entry point either launches CUDA work or fails with a message.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "sdx_sim.cuh"
#include "sdx_task.cuh"
#include "sdx_task_orient.cuh"
#include "sdx_task_insert.cuh"
#include "sdx_task_tool.cuh"
#include "sdx_camera.cuh"
#include "sdx_dr.cuh"
#include "sdx_task_search.cuh"

static thread_local std::string g_err;
extern "C" const char* sdx_last_error(void) { return g_err.c_str(); }
void sdx_set_error(const char* msg) { g_err = msg; }
static int fail(const char* what, cudaError_t e, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s:%d %s: %s", file, line, what, cudaGetErrorString(e));
  g_err = buf;
  return -1;
}
#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return fail(#x, _e, __FILE__, __LINE__); } while (0)
#define CKL() do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return fail("kernel launch", _e, __FILE__, __LINE__); } while (0)

struct sdx_env {
  int n = 0, device = 0;
  uint64_t seed = 0;
  cudaStream_t stream = 0;
  sdx_scene_t host_scene;
  sdx_scene_t* scene = nullptr;
  void* buf[SDX_T_COUNT] = {nullptr};
  float *qcam = nullptr, *finger_dist = nullptr, *static_rows = nullptr, *bank = nullptr, *tvw = nullptr;
  float *gb_hand = nullptr, *gb_obj = nullptr; int* gb_index = nullptr;
  float *tvd_succ = nullptr, *tvd_fail = nullptr; long long* tvd_counts = nullptr; int tvd_cap = 0;   // t-value dataset rings
  int* red_count = nullptr; float* red_sum = nullptr;
  float *stage_obs = nullptr, *stage_states = nullptr, *stage_actions = nullptr;
  int per_type = 0;
  long long total_steps = 0, launches = 0;
  int ws_cur = 0;          // which impulse-cache buffer holds the latest contact list
  bool dump_contacts = false;
  // BlockAssemblyOrient
  int task = 0;
  int* flag_count = nullptr;       // device: number of reset flags set
  int* flag_count_host = nullptr;  // pinned mirror
  int* ob_slot = nullptr;          // [n] ring slot of each env in the current banking call
  float* ob_rows = nullptr; int* ob_index = nullptr; int ob_wrap = 0;   // re-oriented heap rings (sdx_orient_heap_bank)
  int last_reset_sim_steps = 0;
  // BlockAssemblySearch
  sdx_camera_t cam; bool has_cam = false;
  float* last_pixels = nullptr;
  float *sb_rows = nullptr, *sb_hand = nullptr; int* sb_index = nullptr; int sb_wrap = 0;
  int64_t* progress0_host = nullptr;   // pinned: progress_buf[0] (SE:989 reads it on the host every step)
  float4* cscratch = nullptr;      // [n][4][MAXC] contact records / impulses / edge-contact normals of k_simulate
  // BlockAssemblyInsertSim
  float *ib_obj = nullptr, *ib_hand = nullptr; int ib_per_type = 0;   // the banked grasps reset_idx restores (sdx_set_grasp_bank)
  int* slot_by_env = nullptr;      // test hook: reset slots given per env instead of drawn
  int plate_yaw_override = -1;
  // ToolPositioningGrasp / Orient
  float* yaw_u = nullptr;          // test hook: the yaw draw per env
  int pitch_override = -1;
};
static bool is_tool(int task) { return task == SDX_TASK_TOOL_GRASP || task == SDX_TASK_TOOL_ORIENT; }
static int obs_frame(const sdx_env* E) {
  return E->task == SDX_TASK_GRASP_SIM ? SDX_OBS_FRAME : E->task == SDX_TASK_INSERT_SIM ? SDX_INSERT_OBS_FRAME : is_tool(E->task) ? SDX_TOOL_OBS_FRAME : SDX_ORIENT_OBS_FRAME;
}
static int obs_stack(const sdx_env* E) { return E->task == SDX_TASK_INSERT_SIM ? 1 : SDX_STACK; }   // IS:171 stack_obs = 1

static size_t kind_elems(const sdx_env* E, int kind, int64_t shape[4], int* ndim, int* dtype) {
  int64_t n = E->n;
  int64_t s[4] = {1, 1, 1, 1}; int nd = 1, dt = 0;
  switch (kind) {
    case SDX_T_BRICK: s[0] = n; s[1] = 13; s[2] = NB; nd = 3; break;
    case SDX_T_DOF: s[0] = n; s[1] = 3; s[2] = 24; nd = 3; break;
    case SDX_T_LINK: s[0] = n; s[1] = SDX_NL; s[2] = 13; nd = 3; break;
    case SDX_T_JAC7: s[0] = n; s[1] = 6; s[2] = 7; nd = 3; break;
    case SDX_T_NETF: s[0] = n; s[1] = SDX_NL; s[2] = 3; nd = 3; break;
    case SDX_T_ACTIONS: s[0] = n; s[1] = 23; nd = 2; break;
    case SDX_T_OBS: s[0] = n; s[1] = obs_stack(E) * obs_frame(E); nd = 2; break;
    case SDX_T_STATES: s[0] = n; s[1] = obs_stack(E) * SDX_STATE_FRAME; nd = 2; break;
    case SDX_T_REW: case SDX_T_TVALUE: case SDX_T_SUCCESSES: s[0] = n; break;
    case SDX_T_RESET: case SDX_T_PROGRESS: s[0] = n; dt = 1; break;
    case SDX_T_TARGET_INIT: s[0] = n; s[1] = 7; nd = 2; break;
    case SDX_T_CONSEC: s[0] = 1; break;
    case SDX_T_NCONTACT: s[0] = n; s[1] = 4; nd = 2; dt = 2; break;
    case SDX_T_ROOT: s[0] = n * SDX_ACTORS_PER_ENV; s[1] = 13; nd = 2; break;
    case SDX_T_RB: s[0] = n * SDX_RB_PER_ENV; s[1] = 13; nd = 2; break;
    case SDX_T_DOF_STATE: s[0] = n * SDX_ND; s[1] = 2; nd = 2; break;
    case SDX_T_JACOBIAN: s[0] = n; s[1] = SDX_ND; s[2] = 6; s[3] = SDX_ND; nd = 4; break;
    case SDX_T_EPISODE: s[0] = n; dt = 2; break;
    case SDX_T_CONTACTS: s[0] = n; s[1] = SDX_MAX_CONTACTS; s[2] = 8; nd = 3; break;
    case SDX_T_WS: s[0] = n; s[1] = 2; s[2] = SDX_MAX_CONTACTS; s[3] = 4; nd = 4; break;
    case SDX_T_WSN: s[0] = n; s[1] = 2; nd = 2; dt = 2; break;
    case SDX_T_SLEEP: s[0] = n; s[1] = NB; nd = 2; dt = 3; break;
    case SDX_T_SEG: s[0] = n; s[1] = 3; nd = 2; dt = 2; break;
    case SDX_T_EMERGENCE: s[0] = n; break;
    case SDX_T_TVOBS: s[0] = E->task == SDX_TASK_SEARCH ? n : 1; s[1] = SEARCH_TVOBS; nd = 2; break;
    case SDX_T_PLATE: s[0] = n; s[1] = 7; nd = 2; break;
    case SDX_T_ROT_ERR: s[0] = n; s[1] = 3; nd = 2; break;
    case SDX_T_SUCCESS: s[0] = n; s[1] = 2; nd = 2; break;
    default: return 0;
  }
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = s[i];
  if (ndim) *ndim = nd;
  if (dtype) *dtype = dt;
  return (size_t)(s[0] * s[1] * s[2] * s[3]);
}
static size_t dtype_size(int dt) { return dt == 1 ? 8 : (dt == 3 ? 1 : 4); }

extern "C" int sdx_create(const sdx_scene_t* scene, int num_envs, int device, uint64_t seed, sdx_env_t** out) {
  if (!scene || !out || num_envs <= 0) { g_err = "sdx_create: bad arguments"; return -1; }
  if (scene->n_static > KSTAT || scene->n_rshapes > SDX_MAX_RSHAPES || scene->n_bricks > NB) { g_err = "sdx_create: scene exceeds kernel tables"; return -1; }
  if (scene->n_bshapes < 0 || scene->n_bshapes > NB) { g_err = "sdx_create: n_bshapes out of range"; return -1; }
  for (int a = 0; a < scene->n_bshapes; ++a)
    if (scene->bs_body[a] < 0 || scene->bs_body[a] >= scene->n_bricks || (a > 0 && scene->bs_body[a] < scene->bs_body[a - 1])) {
      g_err = "sdx_create: bs_body must map every box to a body, the boxes of one body consecutive"; return -1;
    }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "sdx_create: no CUDA device (this library has no CPU path)"; return -1; }
  CK(cudaSetDevice(device));
  sdx_env* E = new sdx_env();
  E->n = num_envs; E->device = device; E->seed = seed; E->host_scene = *scene;
  if (scene->task < SDX_TASK_GRASP_SIM || scene->task > SDX_TASK_TOOL_ORIENT) { g_err = "sdx_create: unknown scene.task"; delete E; return -1; }
  E->task = scene->task;
  CK(cudaMalloc(&E->scene, sizeof(sdx_scene_t)));
  CK(cudaMemcpy(E->scene, scene, sizeof(sdx_scene_t), cudaMemcpyHostToDevice));
  for (int k = 0; k < SDX_T_COUNT; ++k) {
    int dt = 0;
    size_t ne = kind_elems(E, k, nullptr, nullptr, &dt);
    if (k == SDX_T_CONTACTS) continue;   // debug dump: allocated on demand
    CK(cudaMalloc(&E->buf[k], ne * dtype_size(dt)));
    CK(cudaMemset(E->buf[k], 0, ne * dtype_size(dt)));
  }
  size_t n = num_envs;
  CK(cudaMalloc(&E->qcam, n * 4 * 4)); CK(cudaMemset(E->qcam, 0, n * 16));
  CK(cudaMalloc(&E->finger_dist, n * 4)); CK(cudaMemset(E->finger_dist, 0, n * 4));
  CK(cudaMalloc(&E->static_rows, SDX_ACTORS_PER_ENV * 13 * 4)); CK(cudaMemset(E->static_rows, 0, SDX_ACTORS_PER_ENV * 13 * 4));
  CK(cudaMalloc(&E->tvw, SDX_TVALUE_PARAMS * 4)); CK(cudaMemset(E->tvw, 0, SDX_TVALUE_PARAMS * 4));
  CK(cudaMalloc(&E->gb_hand, (size_t)8 * SDX_GRASP_BANK * 46 * 4)); CK(cudaMemset(E->gb_hand, 0, (size_t)8 * SDX_GRASP_BANK * 46 * 4));
  CK(cudaMalloc(&E->gb_obj, (size_t)8 * SDX_GRASP_BANK * 13 * 4)); CK(cudaMemset(E->gb_obj, 0, (size_t)8 * SDX_GRASP_BANK * 13 * 4));
  CK(cudaMalloc(&E->gb_index, 8 * 4)); CK(cudaMemset(E->gb_index, 0, 32));
  CK(cudaMalloc(&E->red_count, 4)); CK(cudaMemset(E->red_count, 0, 4));
  CK(cudaMalloc(&E->red_sum, 4)); CK(cudaMemset(E->red_sum, 0, 4));
  CK(cudaMalloc(&E->flag_count, 4)); CK(cudaMemset(E->flag_count, 0, 4));
  CK(cudaMallocHost(&E->flag_count_host, 4)); *E->flag_count_host = 0;
  CK(cudaMalloc(&E->ob_slot, n * 4));
  CK(cudaMalloc(&E->last_pixels, n * 4)); CK(cudaMemset(E->last_pixels, 0, n * 4));
  CK(cudaMallocHost(&E->progress0_host, 8)); *E->progress0_host = 0;
  CK(cudaMalloc(&E->stage_obs, n * obs_stack(E) * obs_frame(E) * 4));
  CK(cudaMalloc(&E->stage_states, n * obs_stack(E) * SDX_STATE_FRAME * 4));
  CK(cudaMalloc(&E->stage_actions, n * 23 * 4));
#if SIM_GLOBAL_CONTACTS
  CK(cudaMalloc(&E->cscratch, n * 4 * MAXC * sizeof(float4)));
#endif
  CK(cudaFuncSetAttribute(k_simulate<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SimSmem)));
  CK(cudaFuncSetAttribute(k_simulate<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SimSmem)));
  CK(cudaFuncSetAttribute(k_simulate<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SimSmem)));
  CK(cudaFuncSetAttribute(k_simulate<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SimSmem)));
  *out = E;
  return sdx_reset_all(E);
}

extern "C" void sdx_destroy(sdx_env_t* E) {
  if (!E) return;
  cudaSetDevice(E->device);
  for (int k = 0; k < SDX_T_COUNT; ++k) cudaFree(E->buf[k]);
  cudaFree(E->tvd_succ); cudaFree(E->tvd_fail); cudaFree(E->tvd_counts);
  cudaFree(E->scene); cudaFree(E->qcam); cudaFree(E->finger_dist); cudaFree(E->static_rows); cudaFree(E->bank); cudaFree(E->tvw);
  cudaFree(E->gb_hand); cudaFree(E->gb_obj); cudaFree(E->gb_index); cudaFree(E->red_count); cudaFree(E->red_sum);
  cudaFree(E->stage_obs); cudaFree(E->stage_states); cudaFree(E->stage_actions);
  cudaFree(E->cscratch);
  cudaFree(E->last_pixels); cudaFree(E->sb_rows); cudaFree(E->sb_hand); cudaFree(E->sb_index); cudaFreeHost(E->progress0_host);
  cudaFree(E->ib_obj); cudaFree(E->ib_hand); cudaFree(E->slot_by_env); cudaFree(E->yaw_u);
  cudaFree(E->flag_count); cudaFreeHost(E->flag_count_host); cudaFree(E->ob_slot); cudaFree(E->ob_rows); cudaFree(E->ob_index);
  delete E;
}
extern "C" int sdx_set_stream(sdx_env_t* E, void* stream) { E->stream = (cudaStream_t)stream; return 0; }
extern "C" int sdx_num_envs(const sdx_env_t* E) { return E->n; }
extern "C" int64_t sdx_launch_count(const sdx_env_t* E) { return E->launches; }
extern "C" int sdx_scene_size(void) { return (int)sizeof(sdx_scene_t); }
extern "C" int sdx_sim_smem_bytes(void) { return (int)sizeof(SimSmem); }

extern "C" int sdx_tensor(sdx_env_t* E, int kind, void** dev_ptr, int64_t shape[4], int* ndim, int* dtype) {
  if (kind < 0 || kind >= SDX_T_COUNT) { g_err = "sdx_tensor: bad kind"; return -1; }
  int dt = 0;
  size_t ne = kind_elems(E, kind, shape, ndim, &dt);
  if (dtype) *dtype = dt;
  if (kind == SDX_T_CONTACTS && !E->buf[kind]) {
    CK(cudaSetDevice(E->device));
#ifdef SIM_PROFILE
    const size_t extra = (size_t)E->n * 2 * SIM_PROF_REC * 8;   // per-warp cycle counters behind the contact rows (tools/sim_phase_cycles.py)
#else
    const size_t extra = 0;
#endif
    CK(cudaMalloc(&E->buf[kind], ne * 4 + extra));
    CK(cudaMemset(E->buf[kind], 0, ne * 4 + extra));
    E->dump_contacts = true;
  }
  *dev_ptr = E->buf[kind];
  return 0;
}

#define F(k) ((float*)E->buf[k])
#define I64(k) ((int64_t*)E->buf[k])
#define I32(k) ((int*)E->buf[k])

extern "C" int sdx_set_static_rows(sdx_env_t* E, const float* rows_host /*[142][13]*/) {
  CK(cudaSetDevice(E->device));
  CK(cudaMemcpy(E->static_rows, rows_host, SDX_ACTORS_PER_ENV * 13 * 4, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int sdx_refresh(sdx_env_t* E, int kind) {
  CK(cudaSetDevice(E->device));
  const int n = E->n, T = 256;
  switch (kind) {
    case SDX_T_ROOT: k_refresh_root<<<(n * SDX_ACTORS_PER_ENV + T - 1) / T, T, 0, E->stream>>>(E->scene, n, F(SDX_T_BRICK), E->static_rows, F(SDX_T_ROOT)); break;
    case SDX_T_RB: k_refresh_rb<<<(n * SDX_RB_PER_ENV + T - 1) / T, T, 0, E->stream>>>(E->scene, n, F(SDX_T_BRICK), F(SDX_T_LINK), E->static_rows, F(SDX_T_RB)); break;
    case SDX_T_DOF_STATE: k_refresh_dof_state<<<(n * SDX_ND + T - 1) / T, T, 0, E->stream>>>(n, F(SDX_T_DOF), F(SDX_T_DOF_STATE)); break;
    case SDX_T_JACOBIAN: k_refresh_jacobian<<<(n * SDX_ND * SDX_ND + T - 1) / T, T, 0, E->stream>>>(E->scene, n, F(SDX_T_LINK), F(SDX_T_JACOBIAN)); break;
    case SDX_T_LINK: case SDX_T_JAC7: k_refresh_links<<<(n + 63) / 64, 64, 0, E->stream>>>(E->scene, F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7), n); break;
    case SDX_T_NETF: return 0;   // net contact forces are written by the step itself
    default: g_err = "sdx_refresh: kind is not a refreshable tensor"; return -1;
  }
  E->launches++;
  CKL();
  return 0;
}

extern "C" int sdx_set_actor_root_state_indexed(sdx_env_t* E, const float* root_dev, const int32_t* idx_dev, int n) {
  CK(cudaSetDevice(E->device));
  if (n <= 0) return 0;
  k_set_root_indexed<<<(n + 255) / 256, 256, 0, E->stream>>>(E->scene, E->n, F(SDX_T_BRICK), root_dev, idx_dev, n, (unsigned char*)E->buf[SDX_T_SLEEP]);
  E->launches++; CKL(); return 0;
}
extern "C" int sdx_set_dof_state_indexed(sdx_env_t* E, const float* src, const int32_t* idx_dev, int n) {
  CK(cudaSetDevice(E->device));
  if (n <= 0) return 0;
  k_set_dof_indexed<<<(n * SDX_ND + 255) / 256, 256, 0, E->stream>>>(E->n, F(SDX_T_DOF), src, idx_dev, n, 0);
  E->launches++; CKL(); return 0;
}
extern "C" int sdx_set_dof_target_indexed(sdx_env_t* E, const float* src, const int32_t* idx_dev, int n) {
  CK(cudaSetDevice(E->device));
  if (n <= 0) return 0;
  k_set_dof_indexed<<<(n * SDX_ND + 255) / 256, 256, 0, E->stream>>>(E->n, F(SDX_T_DOF), src, idx_dev, n, 1);
  E->launches++; CKL(); return 0;
}
extern "C" int sdx_set_dof_targets(sdx_env_t* E, const float* src) {
  CK(cudaSetDevice(E->device));
  k_set_dof_targets<<<(E->n * SDX_ND + 255) / 256, 256, 0, E->stream>>>(E->n, F(SDX_T_DOF), src);
  E->launches++; CKL(); return 0;
}

extern "C" int sdx_set_heap_bank(sdx_env_t* E, const float* bank_host, int per_type) {
  CK(cudaSetDevice(E->device));
  if (per_type <= 0) { g_err = "sdx_set_heap_bank: per_type must be positive"; return -1; }
  CK(cudaStreamSynchronize(E->stream));
  cudaFree(E->bank);
  size_t bytes = (size_t)8 * per_type * NB * 13 * 4;
  CK(cudaMalloc(&E->bank, bytes));
  CK(cudaMemcpy(E->bank, bank_host, bytes, cudaMemcpyHostToDevice));
  E->per_type = per_type;
  return 0;
}
// same, from a device buffer (bank generated on the GPU by settling heaps)
extern "C" int sdx_set_heap_bank_dev(sdx_env_t* E, const float* bank_dev, int per_type) {
  CK(cudaSetDevice(E->device));
  CK(cudaStreamSynchronize(E->stream));
  cudaFree(E->bank);
  size_t bytes = (size_t)8 * per_type * NB * 13 * 4;
  CK(cudaMalloc(&E->bank, bytes));
  CK(cudaMemcpy(E->bank, bank_dev, bytes, cudaMemcpyDeviceToDevice));
  E->per_type = per_type;
  return 0;
}

extern "C" int sdx_set_tvalue_weights(sdx_env_t* E, const float* w) {
  CK(cudaSetDevice(E->device));
  std::vector<float> d(SDX_TVALUE_PARAMS);
  const float* W1 = w; const float* b1 = W1 + 1024; const float* W2 = b1 + 256; const float* b2 = W2 + 128 * 256;
  const float* W3 = b2 + 128; const float* b3 = W3 + 64 * 128; const float* W4 = b3 + 64; const float* b4 = W4 + 128;
  float* o = d.data();
  memcpy(o, W1, 1024 * 4); o += 1024; memcpy(o, b1, 256 * 4); o += 256;
  for (int k = 0; k < 256; ++k) for (int q = 0; q < 128; ++q) o[k * 128 + q] = W2[q * 256 + k];
  o += 256 * 128; memcpy(o, b2, 128 * 4); o += 128;
  for (int k = 0; k < 128; ++k) for (int q = 0; q < 64; ++q) o[k * 64 + q] = W3[q * 128 + k];
  o += 128 * 64; memcpy(o, b3, 64 * 4); o += 64;
  memcpy(o, W4, 128 * 4); o += 128; memcpy(o, b4, 2 * 4);
  CK(cudaMemcpy(E->tvw, d.data(), SDX_TVALUE_PARAMS * 4, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int sdx_reset_all(sdx_env_t* E) {
  CK(cudaSetDevice(E->device));
  CK(cudaMemsetAsync(E->buf[SDX_T_WSN], 0, (size_t)E->n * 2 * 4, E->stream));
  CK(cudaMemsetAsync(E->buf[SDX_T_SLEEP], 0, (size_t)E->n * NB, E->stream));
  k_reset_all<<<E->n, 128, 0, E->stream>>>(E->scene, E->n, F(SDX_T_BRICK), F(SDX_T_DOF), I64(SDX_T_PROGRESS), I64(SDX_T_RESET));
  CKL();
  k_refresh_links<<<(E->n + 63) / 64, 64, 0, E->stream>>>(E->scene, F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7), E->n);
  CKL();
  E->launches += 2;
  return 0;
}

static int orient_pre_physics(sdx_env_t* E, const float* actions_dev);
static int search_pre_physics(sdx_env_t* E, const float* actions_dev);
/* the base-plate yaw of a reset_idx call: random.sample([0, 1], 1), ONE draw for all envs that reset (IS:1435).  Philox(seed; step) bit */
static uint32_t reset_call_draw(const sdx_env_t* E) {
  uint32_t k0 = (uint32_t)E->seed, k1 = (uint32_t)(E->seed >> 32), c[4] = {(uint32_t)E->total_steps, 0xC0FFEEu, 7u, 0u};
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c[0];
}
static int insert_plate_yaw(const sdx_env_t* E) { return E->plate_yaw_override >= 0 ? E->plate_yaw_override : (int)(reset_call_draw(E) & 1u); }
/* the tool's pitch index of a reset_idx call: random.sample(range(4), 1), ONE draw for all envs that reset (TG:1489) */
static int tool_pitch_k(const sdx_env_t* E) { return E->pitch_override >= 0 ? E->pitch_override : (int)(reset_call_draw(E) & 3u); }
static int tool_pre_physics(sdx_env_t* E, const float* actions_dev) {
  const int n = E->n, orient = E->task == SDX_TASK_TOOL_ORIENT;
  if (orient && !E->ib_obj) { g_err = "sdx_pre_physics: ToolPositioningOrient needs the banked grasps (sdx_set_grasp_bank; TO:365-368)"; return -1; }
  if (!orient && E->total_steps > 0) {
    k_tool_bank<<<8, 256, 0, E->stream>>>(E->scene, n, F(SDX_T_BRICK), F(SDX_T_DOF), I64(SDX_T_RESET), E->finger_dist, F(SDX_T_PLATE), E->gb_hand,
                                          E->gb_obj, E->gb_index);
    E->launches++;
  }
  k_tool_reset<<<(n + 127) / 128, 128, 0, E->stream>>>(E->scene, n, orient, E->seed, E->ib_obj, E->ib_hand, E->ib_per_type, tool_pitch_k(E), E->slot_by_env,
                                                     E->yaw_u, E->total_steps > 0 ? 1 : 0, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_PLATE),
                                                     F(SDX_T_TARGET_INIT), I64(SDX_T_PROGRESS), I64(SDX_T_RESET), F(SDX_T_SUCCESSES), F(SDX_T_SUCCESS),
                                                     I32(SDX_T_EPISODE), I32(SDX_T_WSN), (unsigned char*)E->buf[SDX_T_SLEEP], F(SDX_T_OBS), F(SDX_T_STATES));
  // the hand has been teleported: the link rows / Jacobian pre_physics reads must be those of the new joint angles
  k_refresh_links<<<(n + 63) / 64, 64, 0, E->stream>>>(E->scene, F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7), n);
  k_tool_pre_physics<<<(n + 127) / 128, 128, 0, E->stream>>>(E->scene, n, orient, actions_dev, F(SDX_T_ACTIONS), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7),
                                                           I64(SDX_T_PROGRESS));
  E->launches += 3;
  CKL();
  return 0;
}
/* ToolPositioningChain (TC:1733-1768): one step of the inner loop -- the frozen policy's actions drive the fingers, the arm holds its previous
 * target (= ToolPositioningOrient's pre-physics, TO:1455-1473), then the contact step.  No reset, no observation, no reward. */
extern "C" int sdx_tool_inner_step(sdx_env_t* E, const float* actions_dev) {
  if (!E || !actions_dev || !is_tool(E->task)) { g_err = "sdx_tool_inner_step: needs a ToolPositioning env and device actions"; return -1; }
  CK(cudaSetDevice(E->device));
  const int n = E->n;
  k_tool_pre_physics<<<(n + 127) / 128, 128, 0, E->stream>>>(E->scene, n, 1, actions_dev, E->stage_actions, F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7),
                                                           I64(SDX_T_PROGRESS));
  E->launches++;
  CKL();
  return sdx_simulate(E);
}
extern "C" int sdx_tool_insertion_obs(sdx_env_t* E, const float* ins_actions_dev, const int64_t* ins_progress_dev, int ins_max_len, float* ins_obs_dev) {
  if (!E || !ins_actions_dev || !ins_progress_dev || !ins_obs_dev || !is_tool(E->task) || ins_max_len <= 0) { g_err = "sdx_tool_insertion_obs: bad arguments"; return -1; }
  CK(cudaSetDevice(E->device));
  k_tool_insertion_obs<<<(E->n + POST_WARPS - 1) / POST_WARPS, 32 * POST_WARPS, 0, E->stream>>>(E->n, F(SDX_T_OBS), ins_actions_dev, ins_progress_dev, ins_max_len,
                                                                                              ins_obs_dev);
  E->launches++;
  CKL();
  return 0;
}
extern "C" int sdx_tool_tvalue_labels(sdx_env_t* E, int* label_dev) {
  if (!E || !label_dev || E->task != SDX_TASK_TOOL_ORIENT) { g_err = "sdx_tool_tvalue_labels: needs a ToolPositioningOrient env and a device label buffer"; return -1; }
  CK(cudaSetDevice(E->device));
  k_tool_tvalue_labels<<<(E->n + 127) / 128, 128, 0, E->stream>>>(E->scene, E->n, F(SDX_T_BRICK), F(SDX_T_PLATE), F(SDX_T_SUCCESS), label_dev);
  E->launches++;
  CKL();
  return 0;
}
extern "C" int sdx_tool_test_hooks(sdx_env_t* E, const int* slot_by_env_host, int pitch_k, const float* yaw_u_host) {
  if (!E || !is_tool(E->task)) { g_err = "sdx_tool_test_hooks: the env does not run a ToolPositioning task"; return -1; }
  if (pitch_k > 3) { g_err = "sdx_tool_test_hooks: pitch index out of range"; return -1; }
  CK(cudaSetDevice(E->device));
  cudaFree(E->slot_by_env); E->slot_by_env = nullptr;
  cudaFree(E->yaw_u); E->yaw_u = nullptr;
  if (slot_by_env_host) {
    CK(cudaMalloc(&E->slot_by_env, (size_t)E->n * 4));
    CK(cudaMemcpy(E->slot_by_env, slot_by_env_host, (size_t)E->n * 4, cudaMemcpyHostToDevice));
  }
  if (yaw_u_host) {
    CK(cudaMalloc(&E->yaw_u, (size_t)E->n * 4));
    CK(cudaMemcpy(E->yaw_u, yaw_u_host, (size_t)E->n * 4, cudaMemcpyHostToDevice));
  }
  E->pitch_override = pitch_k;
  return 0;
}
static int insert_pre_physics(sdx_env_t* E, const float* actions_dev) {
  const int n = E->n;
  if (!E->ib_obj) { g_err = "sdx_pre_physics: InsertSim needs the banked grasps (sdx_set_grasp_bank; IS:372-375)"; return -1; }
  k_insert_reset<<<(n + 127) / 128, 128, 0, E->stream>>>(E->scene, n, E->seed, E->ib_obj, E->ib_hand, E->ib_per_type, insert_plate_yaw(E), E->slot_by_env,
                                                       E->total_steps > 0 ? 1 : 0, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_PLATE), F(SDX_T_TARGET_INIT),
                                                       I64(SDX_T_PROGRESS), I64(SDX_T_RESET), F(SDX_T_SUCCESSES), F(SDX_T_SUCCESS), I32(SDX_T_EPISODE),
                                                       I32(SDX_T_WSN), (unsigned char*)E->buf[SDX_T_SLEEP]);
  // the hand may have been teleported: the link rows / Jacobian pre_physics reads must be those of the restored joint angles
  k_refresh_links<<<(n + 63) / 64, 64, 0, E->stream>>>(E->scene, F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7), n);
  k_insert_pre_physics<<<(n + 127) / 128, 128, 0, E->stream>>>(E->scene, n, actions_dev, F(SDX_T_ACTIONS), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7),
                                                             F(SDX_T_ROT_ERR));
  E->launches += 3;
  CKL();
  return 0;
}
extern "C" int sdx_set_grasp_bank(sdx_env_t* E, const float* hand, const float* obj, int per_type, int is_device) {
  if (!E || !hand || !obj || per_type <= 0) { g_err = "sdx_set_grasp_bank: bad arguments"; return -1; }
  if (E->task != SDX_TASK_INSERT_SIM && E->task != SDX_TASK_TOOL_ORIENT) { g_err = "sdx_set_grasp_bank: the env runs neither BlockAssemblyInsertSim nor ToolPositioningOrient"; return -1; }
  CK(cudaSetDevice(E->device));
  CK(cudaStreamSynchronize(E->stream));
  cudaFree(E->ib_obj); cudaFree(E->ib_hand);
  CK(cudaMalloc(&E->ib_obj, (size_t)8 * per_type * 13 * 4));
  CK(cudaMalloc(&E->ib_hand, (size_t)8 * per_type * 46 * 4));
  const cudaMemcpyKind kind = is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CK(cudaMemcpy(E->ib_obj, obj, (size_t)8 * per_type * 13 * 4, kind));
  CK(cudaMemcpy(E->ib_hand, hand, (size_t)8 * per_type * 46 * 4, kind));
  E->ib_per_type = per_type;
  return 0;
}
/* test hook: the slot every env restores on its next resets (nullptr: drawn from Philox) and the plate yaw index (-1: drawn) */
extern "C" int sdx_insert_test_hooks(sdx_env_t* E, const int* slot_by_env_host, int plate_yaw) {
  CK(cudaSetDevice(E->device));
  cudaFree(E->slot_by_env); E->slot_by_env = nullptr;
  if (slot_by_env_host) {
    CK(cudaMalloc(&E->slot_by_env, (size_t)E->n * 4));
    CK(cudaMemcpy(E->slot_by_env, slot_by_env_host, (size_t)E->n * 4, cudaMemcpyHostToDevice));
  }
  E->plate_yaw_override = plate_yaw;
  return 0;
}
extern "C" int sdx_pre_physics(sdx_env_t* E, const float* actions_dev) {
  CK(cudaSetDevice(E->device));
  const int n = E->n;
  if (E->task == SDX_TASK_SEARCH) return search_pre_physics(E, actions_dev);     // Search resets from the drop lattice: no bank
  if (E->task == SDX_TASK_INSERT_SIM) return insert_pre_physics(E, actions_dev);
  if (is_tool(E->task)) return tool_pre_physics(E, actions_dev);
  if (!E->bank) { g_err = "sdx_pre_physics: no heap bank set (reset_idx samples it, GS:1507-1511)"; return -1; }
  if (E->task == SDX_TASK_ORIENT) return orient_pre_physics(E, actions_dev);
  if (E->total_steps > 0) {
    k_bank_terminal<<<8, 256, 0, E->stream>>>(E->scene, n, F(SDX_T_BRICK), F(SDX_T_DOF), I64(SDX_T_RESET), E->finger_dist,
                                              F(SDX_T_TVALUE), E->gb_hand, E->gb_obj, E->gb_index);
    E->launches++;
    if (E->tvd_cap > 0) {
      k_tv_dataset<<<1, 1024, 0, E->stream>>>(E->scene, n, F(SDX_T_BRICK), I64(SDX_T_RESET), E->finger_dist, F(SDX_T_TVALUE), E->qcam,
                                              E->tvd_succ, E->tvd_fail, E->tvd_counts, E->tvd_cap);
      E->launches++;
    }
  }
  k_reset<<<n, 128, 0, E->stream>>>(E->scene, n, E->seed, E->bank, E->per_type, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_TARGET_INIT),
                                    I64(SDX_T_PROGRESS), I64(SDX_T_RESET), F(SDX_T_SUCCESSES), I32(SDX_T_EPISODE), I32(SDX_T_WSN),
                                    (unsigned char*)E->buf[SDX_T_SLEEP]);
  k_pre_physics<<<(n + 127) / 128, 128, 0, E->stream>>>(E->scene, n, actions_dev, F(SDX_T_ACTIONS), F(SDX_T_DOF), F(SDX_T_LINK),
                                                        F(SDX_T_JAC7), I64(SDX_T_PROGRESS), F(SDX_T_TARGET_INIT));
  E->launches += 2;
  CKL();
  return 0;
}

extern "C" int sdx_simulate(sdx_env_t* E) {
  CK(cudaSetDevice(E->device));
  const bool cmp = E->host_scene.n_bshapes > 0, edge = E->host_scene.edge_contacts > 0.5f;   // compound free bodies or one box per body | edge-edge contacts
  auto kern = cmp ? (edge ? k_simulate<true, true> : k_simulate<true, false>) : (edge ? k_simulate<false, true> : k_simulate<false, false>);
  kern<<<E->n, SIM_THREADS, sizeof(SimSmem), E->stream>>>(E->scene, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7),
                                                              F(SDX_T_NETF), I32(SDX_T_NCONTACT),
                                                              E->dump_contacts ? F(SDX_T_CONTACTS) : nullptr, F(SDX_T_WS), I32(SDX_T_WSN),
                                                              E->ws_cur, (unsigned char*)E->buf[SDX_T_SLEEP], E->n, E->cscratch);
  E->ws_cur ^= (E->host_scene.substeps & 1);
  E->launches++;
  CKL();
  return 0;
}
extern "C" int sdx_simulate_n(sdx_env_t* E, int steps) {
  for (int i = 0; i < steps; ++i) if (sdx_simulate(E)) return -1;
  return 0;
}

static int orient_observe(sdx_env_t* E, int count_step) {
  const int n = E->n;
  k_orient_post_physics<<<(n + POST_WARPS - 1) / POST_WARPS, 32 * POST_WARPS, 0, E->stream>>>(
      E->scene, n, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_ACTIONS), F(SDX_T_TARGET_INIT), I64(SDX_T_PROGRESS),
      I64(SDX_T_RESET), F(SDX_T_OBS), F(SDX_T_STATES), F(SDX_T_REW), E->qcam, E->finger_dist, F(SDX_T_SUCCESSES), E->red_count, E->red_sum,
      count_step);
  k_tvalue<<<(n + TV_ENVS * TV_WARPS - 1) / (TV_ENVS * TV_WARPS), 32 * TV_WARPS, 0, E->stream>>>(E->tvw, n, E->qcam, F(SDX_T_TVALUE), 0.99f);
  E->launches += 2;
  CKL();
  return 0;
}

// BlockAssemblyOrient.pre_physics_step (OR:1697-1778).  reset_idx (OR:1390-1695) is a script over the whole sim, run here when
// any reset flag is set: [lift 50 steps, observe, bank]* (* skipped before the first step), state reset, 2 + 1 settle steps,
// 50 approach steps, flags cleared.
static int orient_pre_physics(sdx_env_t* E, const float* actions_dev) {
  const int n = E->n, T = 128, G = (n + T - 1) / T;
  CK(cudaMemsetAsync(E->flag_count, 0, 4, E->stream));
  k_count_flags<<<(n + 255) / 256, 256, 0, E->stream>>>(I64(SDX_T_RESET), n, E->flag_count);
  E->launches++;
  CK(cudaMemcpyAsync(E->flag_count_host, E->flag_count, 4, cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));                       // the reference's reset_buf.nonzero() is the same host sync
  E->last_reset_sim_steps = 0;
  if (*E->flag_count_host > 0) {
    auto script = [&](int mode, int it) {
      k_orient_arm_script<<<G, T, 0, E->stream>>>(E->scene, n, I64(SDX_T_RESET), mode, it, F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7),
                                                  F(SDX_T_BRICK), F(SDX_T_TARGET_INIT));
      E->launches++;
    };
    auto state = [&](int phase) {
      k_orient_reset<<<n, 128, 0, E->stream>>>(E->scene, n, E->seed, E->bank, E->per_type, phase, F(SDX_T_BRICK), F(SDX_T_DOF),
                                               F(SDX_T_TARGET_INIT), I64(SDX_T_PROGRESS), I64(SDX_T_RESET), F(SDX_T_SUCCESSES),
                                               I32(SDX_T_EPISODE), I32(SDX_T_WSN), (unsigned char*)E->buf[SDX_T_SLEEP]);
      E->launches++;
    };
    auto sim = [&]() -> int { E->last_reset_sim_steps++; return sdx_simulate(E); };
    if (E->total_steps > 0) {
      for (int i = 0; i < 50; ++i) { script(0, i); if (sim()) return -1; }
      if (orient_observe(E, 0)) return -1;
      if (E->ob_wrap > 0) {
        k_orient_bank_slots<<<8, 256, 0, E->stream>>>(E->scene, n, F(SDX_T_BRICK), E->finger_dist, F(SDX_T_TVALUE), E->ob_index, E->ob_wrap, E->ob_slot);
        k_orient_bank_write<<<n, 96, 0, E->stream>>>(E->scene, n, F(SDX_T_BRICK), E->ob_slot, E->ob_rows, E->ob_wrap);
        E->launches += 2;
      }
    }
    state(0);
    if (sim() || sim()) return -1;
    state(1);
    if (sim()) return -1;
    for (int i = 0; i < 50; ++i) { script(1, i); if (sim()) return -1; }
    state(2);
  }
  k_orient_pre_physics<<<G, T, 0, E->stream>>>(E->scene, n, actions_dev, F(SDX_T_ACTIONS), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7),
                                               F(SDX_T_BRICK), I64(SDX_T_PROGRESS), F(SDX_T_TARGET_INIT));
  E->launches++;
  CKL();
  return 0;
}

// ---- BlockAssemblySearch
static int search_render(sdx_env_t* E, int baseline) {     // render_all_camera_sensors + compute_emergence_reward (SE:1446-1455 / 1010-1019)
  if (!E->has_cam) { g_err = "BlockAssemblySearch needs its overview camera: call sdx_set_camera (SE:873-878)"; return -1; }
  if (sdx_segmentation_features(E, &E->cam, I32(SDX_T_SEG))) return -1;
  k_search_emergence<<<(E->n + 255) / 256, 256, 0, E->stream>>>(E->n, I32(SDX_T_SEG), E->last_pixels, F(SDX_T_EMERGENCE), baseline);
  E->launches++;
  CKL();
  return 0;
}
static int search_pre_physics(sdx_env_t* E, const float* actions_dev) {
  const int n = E->n, T = 128, G = (n + T - 1) / T;
  CK(cudaMemsetAsync(E->flag_count, 0, 4, E->stream));
  k_count_flags<<<(n + 255) / 256, 256, 0, E->stream>>>(I64(SDX_T_RESET), n, E->flag_count);
  E->launches++;
  CK(cudaMemcpyAsync(E->flag_count_host, E->flag_count, 4, cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));                       // reset_buf.nonzero() (SE:1540)
  E->last_reset_sim_steps = 0;
  if (*E->flag_count_host > 0) {
    auto state = [&](int phase) {
      k_search_reset<<<n, 128, 0, E->stream>>>(E->scene, n, E->seed, phase, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_TARGET_INIT),
                                               I64(SDX_T_PROGRESS), I64(SDX_T_RESET), F(SDX_T_SUCCESSES), I32(SDX_T_EPISODE),
                                               I32(SDX_T_WSN), (unsigned char*)E->buf[SDX_T_SLEEP]);
      E->launches++;
    };
    if (E->total_steps > 0 && E->sb_wrap > 0) {
      k_search_bank_slots<<<8, 256, 0, E->stream>>>(n, I32(SDX_T_SEG), E->sb_index, E->sb_wrap, E->ob_slot);
      k_search_bank_write<<<n, 96, 0, E->stream>>>(E->scene, n, F(SDX_T_BRICK), F(SDX_T_DOF), E->ob_slot, E->sb_rows, E->sb_hand, E->sb_wrap);
      E->launches += 2;
    }
    state(0);
    for (int i = 0; i < 60; ++i) { E->last_reset_sim_steps++; if (sdx_simulate(E)) return -1; }     // SE:1437-1439
    if (search_render(E, 1)) return -1;
    state(1);
    k_refresh_links<<<(n + 63) / 64, 64, 0, E->stream>>>(E->scene, F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7), n);   // teleported hand
    E->launches++;
    state(2);
  }
  k_search_pre_physics<<<G, T, 0, E->stream>>>(E->scene, n, actions_dev, F(SDX_T_ACTIONS), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_JAC7),
                                               F(SDX_T_BRICK));
  E->launches++;
  CKL();
  return 0;
}
static int search_post_physics(sdx_env_t* E) {
  const int n = E->n;
  CK(cudaMemcpyAsync(E->progress0_host, I64(SDX_T_PROGRESS), 8, cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));                       // `if ... self.progress_buf[0] >= self.max_episode_length - 1` (SE:989)
  if (*E->progress0_host + 1 >= E->host_scene.max_episode_length - 1) {
    k_search_hand_pose<<<(n * SDX_ND + 255) / 256, 256, 0, E->stream>>>(E->scene, n, nullptr, 0, F(SDX_T_DOF));
    E->launches++;
    if (sdx_simulate(E)) return -1;
    if (search_render(E, 0)) return -1;
  }
  k_search_post_physics<<<(n + POST_WARPS - 1) / POST_WARPS, 32 * POST_WARPS, 0, E->stream>>>(
      E->scene, n, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_NETF), F(SDX_T_ACTIONS), F(SDX_T_TARGET_INIT), I32(SDX_T_SEG),
      I64(SDX_T_PROGRESS), I64(SDX_T_RESET), F(SDX_T_OBS), F(SDX_T_STATES), F(SDX_T_TVOBS), F(SDX_T_REW), E->finger_dist,
      F(SDX_T_SUCCESSES), E->red_count, E->red_sum);
  k_finalize<<<1, 1, 0, E->stream>>>(E->scene, E->red_count, E->red_sum, F(SDX_T_CONSEC));
  E->launches += 2;
  E->total_steps++;
  CKL();
  return 0;
}
extern "C" int sdx_set_camera(sdx_env_t* E, const sdx_camera_t* cam) {
  if (!cam || cam->width <= 0 || cam->height <= 0) { g_err = "sdx_set_camera: bad camera"; return -1; }
  E->cam = *cam; E->has_cam = true;
  return 0;
}
extern "C" int sdx_search_bank(sdx_env_t* E, int capacity, void** rows_dev, void** hand_dev, void** index_dev) {
  CK(cudaSetDevice(E->device));
  if (E->task != SDX_TASK_SEARCH) { g_err = "sdx_search_bank: the env does not run BlockAssemblySearch"; return -1; }
  if (capacity > 0 && capacity != E->sb_wrap) {
    CK(cudaStreamSynchronize(E->stream));
    cudaFree(E->sb_rows); cudaFree(E->sb_hand); cudaFree(E->sb_index);
    size_t slots = (size_t)8 * (capacity + 1);
    CK(cudaMalloc(&E->sb_rows, slots * NB * 13 * 4)); CK(cudaMemset(E->sb_rows, 0, slots * NB * 13 * 4));
    CK(cudaMalloc(&E->sb_hand, slots * 46 * 4)); CK(cudaMemset(E->sb_hand, 0, slots * 46 * 4));
    CK(cudaMalloc(&E->sb_index, 32)); CK(cudaMemset(E->sb_index, 0, 32));
    E->sb_wrap = capacity;
  }
  if (capacity == 0) E->sb_wrap = 0;
  if (rows_dev) *rows_dev = E->sb_rows;
  if (hand_dev) *hand_dev = E->sb_hand;
  if (index_dev) *index_dev = E->sb_index;
  return 0;
}

extern "C" int sdx_orient_heap_bank(sdx_env_t* E, int capacity, void** rows_dev, void** index_dev) {
  CK(cudaSetDevice(E->device));
  if (E->task != SDX_TASK_ORIENT) { g_err = "sdx_orient_heap_bank: the env does not run BlockAssemblyOrient"; return -1; }
  if (capacity > 0 && capacity != E->ob_wrap) {
    CK(cudaStreamSynchronize(E->stream));
    cudaFree(E->ob_rows); cudaFree(E->ob_index);
    size_t bytes = (size_t)8 * (capacity + 1) * NB * 13 * 4;
    CK(cudaMalloc(&E->ob_rows, bytes)); CK(cudaMemset(E->ob_rows, 0, bytes));
    CK(cudaMalloc(&E->ob_index, 32)); CK(cudaMemset(E->ob_index, 0, 32));
    E->ob_wrap = capacity;
  }
  if (capacity == 0) E->ob_wrap = 0;
  if (rows_dev) *rows_dev = E->ob_rows;
  if (index_dev) *index_dev = E->ob_index;
  return 0;
}
extern "C" int sdx_last_reset_sim_steps(const sdx_env_t* E) { return E->last_reset_sim_steps; }

extern "C" int sdx_post_physics(sdx_env_t* E) {
  CK(cudaSetDevice(E->device));
  const int n = E->n;
  if (E->task == SDX_TASK_SEARCH) return search_post_physics(E);
  if (E->task == SDX_TASK_INSERT_SIM) {
    k_insert_post_physics<<<(n + 127) / 128, 128, 0, E->stream>>>(E->scene, n, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_ACTIONS),
                                                              F(SDX_T_TARGET_INIT), F(SDX_T_PLATE), F(SDX_T_ROT_ERR), I64(SDX_T_PROGRESS), I64(SDX_T_RESET),
                                                              F(SDX_T_OBS), F(SDX_T_STATES), F(SDX_T_REW), E->finger_dist, F(SDX_T_SUCCESSES), E->red_count,
                                                              E->red_sum);
    k_finalize<<<1, 1, 0, E->stream>>>(E->scene, E->red_count, E->red_sum, F(SDX_T_CONSEC));
    E->launches += 2;
    E->total_steps++;
    CKL();
    return 0;
  }
  if (is_tool(E->task)) {
    k_tool_post_physics<<<(n + POST_WARPS - 1) / POST_WARPS, 32 * POST_WARPS, 0, E->stream>>>(
        E->scene, n, E->task == SDX_TASK_TOOL_ORIENT, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_ACTIONS), F(SDX_T_TARGET_INIT), F(SDX_T_PLATE),
        I64(SDX_T_PROGRESS), I64(SDX_T_RESET), F(SDX_T_OBS), F(SDX_T_STATES), F(SDX_T_REW), E->qcam, E->finger_dist, F(SDX_T_SUCCESSES), E->red_count,
        E->red_sum);
    k_finalize<<<1, 1, 0, E->stream>>>(E->scene, E->red_count, E->red_sum, F(SDX_T_CONSEC));
    E->launches += 2;
    E->total_steps++;
    CKL();
    return 0;
  }
  if (E->task == SDX_TASK_ORIENT) {
    if (orient_observe(E, 1)) return -1;
    k_finalize<<<1, 1, 0, E->stream>>>(E->scene, E->red_count, E->red_sum, F(SDX_T_CONSEC));
    E->launches++;
    E->total_steps++;
    CKL();
    return 0;
  }
  k_post_physics<<<(n + POST_WARPS - 1) / POST_WARPS, 32 * POST_WARPS, 0, E->stream>>>(
      E->scene, n, F(SDX_T_BRICK), F(SDX_T_DOF), F(SDX_T_LINK), F(SDX_T_ACTIONS), F(SDX_T_TARGET_INIT), I64(SDX_T_PROGRESS),
      I64(SDX_T_RESET), F(SDX_T_OBS), F(SDX_T_STATES), F(SDX_T_REW), E->qcam, E->finger_dist, F(SDX_T_SUCCESSES), E->red_count, E->red_sum);
  k_tvalue<<<(n + TV_ENVS * TV_WARPS - 1) / (TV_ENVS * TV_WARPS), 32 * TV_WARPS, 0, E->stream>>>(E->tvw, n, E->qcam, F(SDX_T_TVALUE), 0.0f);
  k_finalize<<<1, 1, 0, E->stream>>>(E->scene, E->red_count, E->red_sum, F(SDX_T_CONSEC));
  E->launches += 3;
  E->total_steps++;
  CKL();
  return 0;
}

extern "C" int sdx_step(sdx_env_t* E, const float* actions_dev) {
  if (sdx_pre_physics(E, actions_dev)) return -1;
  if (sdx_simulate(E)) return -1;
  return sdx_post_physics(E);
}

extern "C" int sdx_step_host(sdx_env_t* E, const float* actions_host, float* obs_host, float* states_host, float* rew_host,
                             int64_t* reset_host) {
  CK(cudaSetDevice(E->device));
  const size_t n = E->n;
  CK(cudaMemcpyAsync(E->stage_actions, actions_host, n * 23 * 4, cudaMemcpyHostToDevice, E->stream));
  if (sdx_step(E, E->stage_actions)) return -1;
  const size_t no = n * obs_stack(E) * obs_frame(E), ns = n * obs_stack(E) * SDX_STATE_FRAME;
  k_clamp_copy<<<(unsigned)((no + 255) / 256), 256, 0, E->stream>>>(F(SDX_T_OBS), E->stage_obs, no, 5.0f);
  k_clamp_copy<<<(unsigned)((ns + 255) / 256), 256, 0, E->stream>>>(F(SDX_T_STATES), E->stage_states, ns, 5.0f);
  E->launches += 2;
  CKL();
  if (obs_host) CK(cudaMemcpyAsync(obs_host, E->stage_obs, no * 4, cudaMemcpyDeviceToHost, E->stream));
  if (states_host) CK(cudaMemcpyAsync(states_host, E->stage_states, ns * 4, cudaMemcpyDeviceToHost, E->stream));
  if (rew_host) CK(cudaMemcpyAsync(rew_host, F(SDX_T_REW), n * 4, cudaMemcpyDeviceToHost, E->stream));
  if (reset_host) CK(cudaMemcpyAsync(reset_host, I64(SDX_T_RESET), n * 8, cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  return 0;
}

extern "C" int sdx_gae(const float* rewards, const float* values, const float* dones, const float* last_values,
                       const float* last_dones, float* adv, float* returns, int horizon, int n, float gamma, float tau, void* stream) {
  k_gae<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rewards, values, dones, last_values, last_dones, adv, returns, horizon, n, gamma, tau);
  CKL();
  return 0;
}

/* VecTask clamp of a task buffer into caller memory (VR:174-175): dst = clamp(tensor(kind), -lim, lim) */
extern "C" int sdx_clamped_copy(sdx_env_t* E, int kind, float* dst_dev, float lim) {
  CK(cudaSetDevice(E->device));
  int dt = 0;
  size_t ne = kind_elems(E, kind, nullptr, nullptr, &dt);
  if (dt != 0 || !E->buf[kind]) { g_err = "sdx_clamped_copy: not a float tensor"; return -1; }
  k_clamp_copy<<<(unsigned)((ne + 255) / 256), 256, 0, E->stream>>>((const float*)E->buf[kind], dst_dev, ne, lim);
  E->launches++;
  CKL();
  return 0;
}

/* domain randomisation: correlated-noise tensor and the noise hook of BaseTask.step (BT:131-132, 149-150, 263-340) */
extern "C" int sdx_dr_randn(sdx_env_t* E, float* dst_dev, int64_t n, uint64_t seed, uint32_t counter) {
  if (!E || !dst_dev || n <= 0) { g_err = "sdx_dr_randn: bad arguments"; return -1; }
  CK(cudaSetDevice(E->device));
  k_dr_randn<<<(unsigned)((n + 1023) / 1024), 256, 0, E->stream>>>(dst_dev, n, seed, counter);
  E->launches++;
  CKL();
  return 0;
}
extern "C" int sdx_dr_noise(sdx_env_t* E, float* dst_dev, const float* src_dev, const float* corr_dev, int64_t n, float a_corr,
                            float b_corr, float a, float b, int distribution, int operation, uint64_t seed, uint32_t counter) {
  if (!E || !dst_dev || !src_dev || !corr_dev || n <= 0) { g_err = "sdx_dr_noise: bad arguments"; return -1; }
  if (distribution < 0 || distribution > 1 || operation < 0 || operation > 1) { g_err = "sdx_dr_noise: unknown distribution / operation"; return -1; }
  CK(cudaSetDevice(E->device));
  k_dr_noise<<<(unsigned)((n + 1023) / 1024), 256, 0, E->stream>>>(dst_dev, src_dev, corr_dev, n, a_corr, b_corr, a, b, distribution, operation, seed, counter);
  E->launches++;
  CKL();
  return 0;
}
/* sim_params.gravity randomisation (BT:342-355): the z component of gravity for every env from the next step on */
extern "C" int sdx_set_gravity(sdx_env_t* E, float gravity_z) {
  if (!E) { g_err = "sdx_set_gravity: bad arguments"; return -1; }
  CK(cudaSetDevice(E->device));
  E->host_scene.gravity_z = gravity_z;
  CK(cudaMemcpyAsync(&E->scene->gravity_z, &E->host_scene.gravity_z, sizeof(float), cudaMemcpyHostToDevice, E->stream));
  return 0;
}

/* grasp terminal-state banks: device pointers for export */
extern "C" int sdx_grasp_bank(sdx_env_t* E, void** hand_dev, void** obj_dev, void** index_dev) {
  *hand_dev = E->gb_hand; *obj_dev = E->gb_obj; *index_dev = E->gb_index;
  return 0;
}
/* t-value training data rings (GS:1402-1438 save_hdf5 / TVT:132-168): capacity > 0 (re)allocates and switches recording on */
extern "C" int sdx_tvalue_dataset(sdx_env_t* E, int capacity, void** succ_dev, void** fail_dev, void** counts_dev) {
  CK(cudaSetDevice(E->device));
  if (capacity > 0 && capacity != E->tvd_cap) {
    CK(cudaStreamSynchronize(E->stream));
    cudaFree(E->tvd_succ); cudaFree(E->tvd_fail); cudaFree(E->tvd_counts);
    CK(cudaMalloc(&E->tvd_succ, (size_t)capacity * 16)); CK(cudaMemset(E->tvd_succ, 0, (size_t)capacity * 16));
    CK(cudaMalloc(&E->tvd_fail, (size_t)capacity * 16)); CK(cudaMemset(E->tvd_fail, 0, (size_t)capacity * 16));
    CK(cudaMalloc(&E->tvd_counts, 16)); CK(cudaMemset(E->tvd_counts, 0, 16));
    E->tvd_cap = capacity;
  }
  if (capacity == 0) E->tvd_cap = 0;   // recording off (buffers kept until destroy)
  if (succ_dev) *succ_dev = E->tvd_succ;
  if (fail_dev) *fail_dev = E->tvd_fail;
  if (counts_dev) *counts_dev = E->tvd_counts;
  return 0;
}
extern "C" int sdx_segmentation_features(sdx_env_t* E, const sdx_camera_t* cam, int32_t* out_dev) {
  CK(cudaSetDevice(E->device));
  if (!cam || !out_dev || cam->width <= 0 || cam->height <= 0 || cam->width > 4096 || cam->height > 4096) { g_err = "sdx_segmentation_features: bad camera"; return -1; }
  if (E->host_scene.n_bricks + E->host_scene.n_rshapes + E->host_scene.n_static > CAM_MAX_SHAPES) { g_err = "sdx_segmentation_features: scene exceeds the shape table"; return -1; }
  k_seg_features<<<E->n, 128, 0, E->stream>>>(E->scene, E->n, *cam, F(SDX_T_BRICK), F(SDX_T_LINK), out_dev);
  E->launches++;
  CKL();
  return 0;
}
extern "C" int sdx_aux(sdx_env_t* E, void** qcam_dev, void** finger_dist_dev) { *qcam_dev = E->qcam; *finger_dist_dev = E->finger_dist; return 0; }
