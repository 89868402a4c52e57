// sdx_task_tool.cuh -- ToolPositioningGrasp / ToolPositioningOrient (SDX_TASK_TOOL_GRASP / SDX_TASK_TOOL_ORIENT; BASELINE configs[4]) as
// fused kernels.  TG = tasks/tool_positioning/allegro_hand_tool_positioning_grasp.py, TO = ..._orient.py.  One free body per env: the
// tool (body 0, a compound of boxes).
//   k_tool_bank         : TG reset_idx's banking of good grasps into per-type rings, in env order                (TG:1436-1457)
//   k_tool_reset        : reset_idx -- TG: tool to its start pose with a drawn pitch / yaw, hand to its start pose, history zeroed
//                         (TG:1412-1578); TO: a banked grasp restored, velocities included (TO:1265-1436)
//   k_tool_pre_physics  : TG: finger EMA + arm IK with the scripted lift / park (TG:1580-1675); TO: fingers only (TO:1438-1509)
//   k_tool_post_physics : 156-slot observation x 3, 188-slot privileged state x 3, reward, reset flags, one warp per env
//                         (TG:1137-1368, 1741-1893; TO:1018-1236, 1574-1626)
// Same arithmetic, operation for operation, as oracle/sdx_oracle.c "ToolPositioning" (bit-exact parity); the oracle is pinned to the
// reference's own Python (tests/golden/tool_*.npz).
#pragma once
#include "sdx_task_insert.cuh"

#define TOOL_OBS SDX_TOOL_OBS_FRAME
#define TOOL_BODY 0

__device__ __forceinline__ float tool_rot_dist(q4 tq, q4 eq) {     // TG:1871-1872, TO:1587-1588
  const q4 d = qmul(tq, qconj(eq));
  const float nn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
  return 2.0f * sdx_asin(nn > 1.0f ? 1.0f : nn);
}
__device__ __forceinline__ float tool_signed_sq(float d) { return (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * (d * d); }   // sign(d) * d ** 2
// pytorch3d.transforms.quaternion_to_matrix reads (r, i, j, k) = q[0..3]; the reference hands it xyzw quaternions (TG:1853-1854), so
// "r" is the x component.  Rows of the matrix it returns.
__device__ __forceinline__ void tool_p3d_matrix(q4 q, float* M) {
  const float r = q.x, i = q.y, j = q.z, k = q.w;
  const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  M[0] = 1.0f - two_s * (j * j + k * k); M[1] = two_s * (i * j - k * r); M[2] = two_s * (i * k + j * r);
  M[3] = two_s * (i * j + k * r); M[4] = 1.0f - two_s * (i * i + k * k); M[5] = two_s * (j * k - i * r);
  M[6] = two_s * (i * k - j * r); M[7] = two_s * (j * k + i * r); M[8] = 1.0f - two_s * (i * i + j * j);
}

// Banking (TG:1436-1457): a resetting env whose tool is above 0.8 m, within 0.4 of the fingertips and within 1 rad of the plate's
// orientation appends (hand DoF state, tool root row) to the ring of its type (env % 8) in ENV ORDER; the index returns to 0
// after slot SDX_TOOL_BANK_WRAP.  (The reference's eight lists alias ONE tensor, TG:441-442; here the rings are separate.)
__global__ void __launch_bounds__(256)
k_tool_bank(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ dof,
            const int64_t* __restrict__ reset, const float* __restrict__ finger_dist, const float* __restrict__ plate,
            float* __restrict__ gb_hand, float* __restrict__ gb_obj, int* __restrict__ gb_index) {
  __shared__ int cnt[256];
  const int ty = blockIdx.x, tid = threadIdx.x;
  const int m = (n - ty + 7) / 8;
  const int per = (m + 255) / 256;
  const int i0 = tid * per, i1 = min(m, i0 + per);
  int c = 0;
  for (int i = i0; i < i1; ++i) {
    const int e = ty + 8 * i;
    if (!reset[e]) continue;
    float row[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, TOOL_BODY, row);
    const float* pl = plate + 7 * e;
    if (row[2] > 0.8f && finger_dist[e] < 0.4f && tool_rot_dist(Q4(row[3], row[4], row[5], row[6]), Q4(pl[3], pl[4], pl[5], pl[6])) < 1.0f) c++;
  }
  cnt[tid] = c;
  __syncthreads();
  __shared__ int base, total;
  if (tid == 0) {
    int o = 0;
    for (int t = 0; t < 256; ++t) { const int v = cnt[t]; cnt[t] = o; o += v; }
    base = gb_index[ty]; total = o;
  }
  __syncthreads();
  int k = cnt[tid];
  for (int i = i0; i < i1; ++i) {
    const int e = ty + 8 * i;
    if (!reset[e]) continue;
    float row[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, TOOL_BODY, row);
    const float* pl = plate + 7 * e;
    if (row[2] > 0.8f && finger_dist[e] < 0.4f && tool_rot_dist(Q4(row[3], row[4], row[5], row[6]), Q4(pl[3], pl[4], pl[5], pl[6])) < 1.0f) {
      const int slot = (base + k) % (SDX_TOOL_BANK_WRAP + 1);
      float* hd = gb_hand + ((size_t)ty * SDX_GRASP_BANK + slot) * 46;
      const float* d = dof + (size_t)e * 72;
      for (int j = 0; j < SDX_ND; ++j) { hd[2 * j] = d[j]; hd[2 * j + 1] = d[24 + j]; }
      float* ob = gb_obj + ((size_t)ty * SDX_GRASP_BANK + slot) * 13;
      for (int q = 0; q < 13; ++q) ob[q] = row[q];
      k++;
    }
  }
  __syncthreads();
  if (tid == 0) gb_index[ty] = (base + total) % (SDX_TOOL_BANK_WRAP + 1);
}

// Labels of ToolPositioningOrient's online t-value update (TO:1305-1316), for ALL envs from their CURRENT state: success = the tool within
// 1 cm of the plate's position and within 0.1 rad of its orientation or that orientation turned by pi about z.  Writes
// success_buf = [success, not success] and label = the column that is 1 (0 success, 1 failure: the target layout of sdx_tvalue_bce).
__global__ void k_tool_tvalue_labels(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ plate,
                                     float* __restrict__ success_buf, int* __restrict__ label) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float tg[13];
  brick_root_row(S, brick + (size_t)e * 13 * NB, TOOL_BODY, tg);
  const float* pl = plate + 7 * e;
  const float rd = rot_dist_sym(Q4(tg[3], tg[4], tg[5], tg[6]), Q4(pl[3], pl[4], pl[5], pl[6]));
  const v3 dp = vsub(V3(pl[0], pl[1], pl[2]), V3(tg[0], tg[1], tg[2]));
  const float ok = (sqrtf(vdot(dp, dp)) < 0.01f && rd < 0.1f) ? 1.0f : 0.0f;
  success_buf[2 * e] = ok; success_buf[2 * e + 1] = ok <= 0.5f ? 1.0f : 0.0f;
  label[e] = ok > 0.5f ? 0 : 1;
}

// ToolPositioningChain's second observation buffer (TC = tasks/tool_positioning/allegro_hand_tool_positioning_chain.py:1404-1440): the frame
// compute_contact_observations has just written (TC:1306 calls this right after it) with the INNER policy's last actions in 23:46 and the
// inner episode clock in slot 60, over its own two history frames.  One warp per env.
__global__ void __launch_bounds__(32 * POST_WARPS)
k_tool_insertion_obs(int n, const float* __restrict__ obs, const float* __restrict__ ins_actions, const int64_t* __restrict__ ins_progress,
                     int ins_max_len, float* __restrict__ ins_obs) {
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * POST_WARPS + wid;
  if (e >= n) return;
  const float* o = obs + (size_t)e * 3 * TOOL_OBS;
  float* io = ins_obs + (size_t)e * 3 * TOOL_OBS;
  float ho[(2 * TOOL_OBS + 31) / 32];
#pragma unroll
  for (int i = 0; i < (2 * TOOL_OBS + 31) / 32; ++i) { const int k = lane + 32 * i; ho[i] = k < 2 * TOOL_OBS ? io[k] : 0.0f; }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < (2 * TOOL_OBS + 31) / 32; ++i) { const int k = lane + 32 * i; if (k < 2 * TOOL_OBS) io[TOOL_OBS + k] = ho[i]; }
  for (int k = lane; k < TOOL_OBS; k += 32) {
    float v = o[k];
    if (k >= 23 && k < 46) v = ins_actions[23 * (size_t)e + (k - 23)];
    if (k == 60) v = (float)ins_progress[e] / (float)ins_max_len;
    io[k] = v;
  }
}

// orient = 0: TG reset_idx; orient = 1: TO reset_idx.  slot_by_env / yaw_u (nullable) are the parity tests' hooks.
__global__ void __launch_bounds__(128)
k_tool_reset(const sdx_scene_t* __restrict__ S, int n, int orient, uint64_t seed, const float* __restrict__ bank_obj,
             const float* __restrict__ bank_hand, int per_type, int pitch_k, const int* __restrict__ slot_by_env,
             const float* __restrict__ yaw_u, int do_success, float* __restrict__ brick, float* __restrict__ dof, float* __restrict__ plate,
             float* __restrict__ target_init, int64_t* __restrict__ progress, int64_t* __restrict__ reset, float* __restrict__ successes,
             float* __restrict__ success_buf, int* __restrict__ episode, int* __restrict__ wsn, unsigned char* __restrict__ slp,
             float* __restrict__ obs, float* __restrict__ states) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !reset[e]) return;
  float* B = brick + (size_t)e * 13 * NB;
  float* d = dof + (size_t)e * 72;
  float* pl = plate + 7 * e;
  if (do_success) {                                                  // TG:1425-1428 (< 0.3), TO:1279-1282 (< 0.5): from the state the episode ended in
    float tg[13];
    brick_root_row(S, B, TOOL_BODY, tg);
    const float rd = tool_rot_dist(Q4(tg[3], tg[4], tg[5], tg[6]), Q4(pl[3], pl[4], pl[5], pl[6]));
    success_buf[2 * e] = rd < (orient ? 0.5f : 0.3f) ? 1.0f : 0.0f;
  }
  uint32_t r[4];
  philox(seed, (uint32_t)e, (uint32_t)episode[e], 1u, r);
  float row[13];
  if (orient) {
    const int slot = slot_by_env ? slot_by_env[e] : (int)(r[0] % (uint32_t)per_type);
    const float* ob = bank_obj + (((size_t)(e % 8)) * per_type + slot) * 13;
    const float* hd = bank_hand + (((size_t)(e % 8)) * per_type + slot) * 46;
    for (int k = 0; k < 13; ++k) row[k] = ob[k];                     // TO:1397: the whole root row, velocities included
    for (int j = 0; j < SDX_ND; ++j) { d[j] = hd[2 * j]; d[24 + j] = hd[2 * j + 1]; d[48 + j] = hd[2 * j]; }   // TO:1398, 1421-1422
  } else {
    const float u = yaw_u ? yaw_u[e] : (float)(r[1] >> 8) * (2.0f / 16777216.0f) - 1.0f;
    float sy, cy;
    sdx_sincos((u * 3.14f) * 0.5f, &sy, &cy);                        // quat_from_euler_xyz(0, k * 1.571, u * 3.14), TG:1495
    const float sp = S->tool_pitch_sc[2 * pitch_k], cp = S->tool_pitch_sc[2 * pitch_k + 1];
    row[0] = S->tool_reset_pos[0]; row[1] = S->tool_reset_pos[1]; row[2] = S->tool_reset_pos[2];
    row[3] = 0.0f - sy * sp; row[4] = cy * sp; row[5] = sy * cp; row[6] = cy * cp;
    for (int k = 7; k < 13; ++k) row[k] = 0.0f;
    for (int j = 0; j < 7; ++j) { d[j] = S->prepare_arm[j]; d[24 + j] = 0.0f; d[48 + j] = S->prepare_arm[j]; }        // TG:1536-1541
    for (int i = 0; i < 16; ++i) {                                                                                   // TG:1543-1548
      const float v = scalef(S->finger_reset_unscaled[i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
      d[7 + i] = v; d[24 + 7 + i] = 0.0f; d[48 + 7 + i] = v;
    }
    float* o = obs + (size_t)e * 3 * TOOL_OBS;                       // TG:1563-1568: obs_buf and both stacks of history frames zeroed
    for (int k = 0; k < 3 * TOOL_OBS; ++k) o[k] = 0.0f;
    float* s = states + (size_t)e * 3 * STATE_FRAME;
    for (int k = 0; k < 3 * STATE_FRAME; ++k) s[k] = 0.0f;
  }
  brick_from_root_row(S, B, TOOL_BODY, row);
  for (int k = 0; k < 7; ++k) { pl[k] = S->tool_plate_pose[k]; target_init[7 * e + k] = row[k]; }
  progress[e] = 0; reset[e] = 0; successes[e] = 0.0f;
  wsn[2 * e] = 0; wsn[2 * e + 1] = 0;
  for (int b = 0; b < NB; ++b) slp[(size_t)e * NB + b] = 0;
  episode[e] += 1;
}

__global__ void __launch_bounds__(128)
k_tool_pre_physics(const sdx_scene_t* __restrict__ S, int n, int orient, const float* __restrict__ actions_in, float* __restrict__ actions,
                   float* __restrict__ dof, const float* __restrict__ link, const float* __restrict__ jac7,
                   const int64_t* __restrict__ progress) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float* a = actions_in + 23 * (size_t)e;
  float* d = dof + (size_t)e * 72;
  float cur[23];
  for (int k = 0; k < 23; ++k) actions[23 * (size_t)e + k] = a[k];
  for (int i = 0; i < 16; ++i) {
    const float t = scalef(a[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
    cur[7 + i] = S->act_moving_average * t + (1.0f - S->act_moving_average) * d[48 + 7 + i];
  }
  if (orient) {
    for (int j = 0; j < 7; ++j) cur[j] = clampf(d[48 + j], S->dof_lo[j], S->dof_hi[j]);      // TO:1471-1473: the arm holds its previous target
  } else {
    const int64_t pg = progress[e];
    float dpose[6] = {a[0] * 0.2f, a[1] * 0.2f, a[2] * 0.2f, 0.0f, 0.0f, 0.0f};
    if (pg >= 60) { dpose[2] = 0.1f; dpose[0] = 0.0f; dpose[1] = 0.0f; }                        // TG:1622-1626: lift
    const float* hb = link + ((size_t)e * SDX_NL + 7) * 13;
    const q4 want = Q4(S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]);
    const v3 re = orientation_error(want, Q4(hb[3], hb[4], hb[5], hb[6]));
    dpose[3] = re.x * 5.0f; dpose[4] = re.y * 5.0f; dpose[5] = re.z * 5.0f;                    // TG:1629
    float u[7];
    control_ik(jac7 + 42 * (size_t)e, dpose, u);
    for (int j = 0; j < 7; ++j) cur[j] = d[j] + u[j];
    if (pg > 90) {                                                                              // TG:1635-1636: park, fingers hold
      for (int j = 0; j < 7; ++j) cur[j] = S->insert_prep0[j];
      for (int i = 7; i < 23; ++i) cur[i] = d[48 + i];
    }
  }
  for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
}

__global__ void __launch_bounds__(32 * POST_WARPS)
k_tool_post_physics(const sdx_scene_t* __restrict__ S, int n, int orient, const float* __restrict__ brick, const float* __restrict__ dof,
                    const float* __restrict__ link, const float* __restrict__ actions, const float* __restrict__ target_init,
                    const float* __restrict__ plate, int64_t* __restrict__ progress, int64_t* __restrict__ reset, float* __restrict__ obs,
                    float* __restrict__ states, float* __restrict__ rew, float* __restrict__ qcam, float* __restrict__ finger_dist_out,
                    float* __restrict__ successes, int* __restrict__ red_count, float* __restrict__ red_sum) {
  __shared__ float fo[POST_WARPS][TOOL_OBS];
  __shared__ float fs[POST_WARPS][STATE_FRAME];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * POST_WARPS + wid;
  if (e >= n) return;
  float* o = obs + (size_t)e * 3 * TOOL_OBS;
  float* s = states + (size_t)e * 3 * STATE_FRAME;
  // history shift (TG:1334-1336, 1366-1368): the two newest frames move one slot older
  float ho[(2 * TOOL_OBS + 31) / 32], hs[(2 * STATE_FRAME + 31) / 32];
#pragma unroll
  for (int i = 0; i < (2 * TOOL_OBS + 31) / 32; ++i) { const int k = lane + 32 * i; ho[i] = k < 2 * TOOL_OBS ? o[k] : 0.0f; }
#pragma unroll
  for (int i = 0; i < (2 * STATE_FRAME + 31) / 32; ++i) { const int k = lane + 32 * i; hs[i] = k < 2 * STATE_FRAME ? s[k] : 0.0f; }
  if (lane == 0) {
    float* f = fo[wid]; float* g = fs[wid];
    const int64_t pg = progress[e] + 1;
    progress[e] = pg;
    const float* L = link + (size_t)e * SDX_NL * 13;
    const float* d = dof + (size_t)e * 72;
    const float* hb = L + 7 * 13;
    const float* ff = L + 11 * 13; const float* mf = L + 19 * 13; const float* rf = L + 23 * 13; const float* th = L + 15 * 13;
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, TOOL_BODY, tg);
    const v3 tp = V3(tg[0], tg[1], tg[2]); const q4 tq = Q4(tg[3], tg[4], tg[5], tg[6]);
    v3 tip[4]; const float* fsr[4] = {ff, mf, rf, th};
    float nrm[4];
    for (int i = 0; i < 4; ++i) {
      const q4 fq = Q4(fsr[i][3], fsr[i][4], fsr[i][5], fsr[i][6]);
      tip[i] = vadd(V3(fsr[i][0], fsr[i][1], fsr[i][2]), qrot(fq, V3(0.0f, 0.0f, 1.0f * 0.04f)));
      const v3 dd = vsub(tp, tip[i]); nrm[i] = sqrtf(vdot(dd, dd));
    }
    const float fdist = nrm[0] + nrm[1] + nrm[2] + nrm[3];
    finger_dist_out[e] = fdist;
    const q4 hq = Q4(hb[3], hb[4], hb[5], hb[6]); const v3 hp = V3(hb[0], hb[1], hb[2]);
    const q4 cq0 = Q4(S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]);
    const q4 cq = qmul(hq, cq0); const v3 cp = vadd(qrot(hq, V3(S->cam_off_pos[0], S->cam_off_pos[1], S->cam_off_pos[2])), hp);
    const q4 cqi = qconj(cq); const v3 cpi = vneg(qrot(cqi, cp));
    const q4 cvq = qmul(cqi, tq); const v3 cvp = vadd(qrot(cqi, tp), cpi);
    qcam[4 * e] = cvq.x; qcam[4 * e + 1] = cvq.y; qcam[4 * e + 2] = cvq.z; qcam[4 * e + 3] = cvq.w;
    const float* ti = target_init + 7 * e;
    const float* pl = plate + 7 * e;
    const v3 ep = V3(pl[0], pl[1], pl[2]); const q4 eq = Q4(pl[3], pl[4], pl[5], pl[6]);
    // ---- observation frame (TG:1338-1364 = TO:1201-1232)
    for (int j = 0; j < 23; ++j) { f[j] = unscalef(d[j], S->dof_lo[j], S->dof_hi[j]); f[23 + j] = actions[23 * (size_t)e + j]; }
    for (int k = 0; k < 7; ++k) { f[46 + k] = hb[k]; f[53 + k] = tg[k]; f[61 + k] = pl[k]; }
    f[60] = (float)pg / (float)S->max_episode_length;
    f[68] = tp.x - ep.x; f[69] = tp.y - ep.y; f[70] = tp.z - ep.z;
    { const q4 r = qmul(tq, qconj(eq)); f[71] = r.x; f[72] = r.y; f[73] = r.z; f[74] = r.w; }
    for (int k = 0; k < 13; ++k) { f[75 + k] = ff[k]; f[88 + k] = rf[k]; f[101 + k] = mf[k]; f[114 + k] = th[k]; }
    for (int j = 0; j < 23; ++j) f[127 + j] = S->vel_obs_scale * d[24 + j];
    for (int k = 0; k < 6; ++k) f[150 + k] = tg[7 + k];
    // ---- privileged frame (TG:1274-1332; TO:1137-1195 differs in 181:188)
    for (int j = 0; j < 23; ++j) { g[j] = unscalef(d[j], S->dof_lo[j], S->dof_hi[j]); g[23 + j] = S->vel_obs_scale * d[24 + j]; }
    g[46] = tip[0].x; g[47] = tip[0].y; g[48] = tip[0].z;
    g[49] = tip[2].x; g[50] = tip[2].y; g[51] = tip[2].z;
    g[52] = tip[1].x; g[53] = tip[1].y; g[54] = tip[1].z;
    g[55] = tip[3].x; g[56] = tip[3].y; g[57] = tip[3].z;
    for (int k = 0; k < 23; ++k) g[58 + k] = actions[23 * (size_t)e + k];
    for (int k = 0; k < 7; ++k) { g[81 + k] = hb[k]; g[88 + k] = tg[k]; }
    for (int k = 0; k < 6; ++k) g[95 + k] = hb[7 + k];
    for (int k = 0; k < 4; ++k) { g[101 + k] = ff[3 + k]; g[111 + k] = mf[3 + k]; g[121 + k] = rf[3 + k]; g[131 + k] = th[3 + k]; }
    for (int k = 0; k < 6; ++k) { g[105 + k] = ff[7 + k]; g[115 + k] = mf[7 + k]; g[125 + k] = rf[7 + k]; g[135 + k] = th[7 + k]; }
    for (int k = 0; k < 6; ++k) g[142 + k] = tg[7 + k];
    g[148] = ti[0]; g[149] = ti[1]; g[150] = ti[2];
    g[151] = tp.x - ti[0]; g[152] = tp.y - ti[1]; g[153] = tp.z - ti[2];
    g[154] = hp.x - tp.x; g[155] = hp.y - tp.y; g[156] = hp.z - tp.z;
    { const q4 rel = qmul(hq, qconj(tq)); g[157] = rel.x; g[158] = rel.y; g[159] = rel.z; g[160] = rel.w; }
    { const v3 a = vsub(tp, tip[0]), b = vsub(tp, tip[2]), c = vsub(tp, tip[1]), dd = vsub(tp, tip[3]);
      g[161] = a.x; g[162] = a.y; g[163] = a.z; g[164] = b.x; g[165] = b.y; g[166] = b.z;
      g[167] = c.x; g[168] = c.y; g[169] = c.z; g[170] = dd.x; g[171] = dd.y; g[172] = dd.z; }
    g[173] = fdist;
    g[174] = cvp.x; g[175] = cvp.y; g[176] = cvp.z; g[177] = cvq.x; g[178] = cvq.y; g[179] = cvq.z; g[180] = cvq.w;
    if (orient) { for (int k = 0; k < 7; ++k) g[181 + k] = pl[k]; }
    else { g[181] = cvp.x; g[182] = cvp.y; g[183] = cvp.z; g[184] = cvq.x; g[185] = cvq.y; g[186] = cvq.z; g[187] = cvq.w; }
    // ---- reward / resets
    const float dist = nrm[0] + nrm[1] + nrm[2] + 3.0f * nrm[3];
    const float rd = tool_rot_dist(tq, eq);
    int64_t rs = reset[e];
    float sc = successes[e];
    if (orient) {                                                    // TO:1574-1626
      if (dist >= 20.0f) rs = 1;
      if ((float)pg >= (float)S->max_episode_length - 1.0f) rs = 1;
      rew[e] = (rd < 0.2f ? 1.0f : 0.0f) + sdx_exp(-1.0f * rd);
    } else {                                                         // TG:1741-1893
      const float zal = tool_signed_sq(qrot(tq, V3(0.0f, 0.0f, 1.0f)).z);
      if (dist <= -1.0f) rs = 1;
      if (pg >= 150 && zal <= 0.75f) rs = 1;
      if (pg >= 150 && dist >= 0.4f) rs = 1;
      const float ay = tp.y - ti[1], ax = tp.x - ti[0];
      if (pg <= 90 && (ay < 0.0f ? -ay : ay) >= 0.08f) rs = 1;
      if (pg <= 90 && (ax < 0.0f ? -ax : ax) >= 0.08f) rs = 1;
      if ((float)pg >= (float)S->max_episode_length - 1.0f) rs = 1;
      // successes = [1 - cos(angle between the tool's y axis now and at reset) >= 1.95] with both rotation matrices read real-first (TG:1848-1868)
      float Mi[9], Mc[9];
      tool_p3d_matrix(Q4(ti[3], ti[4], ti[5], ti[6]), Mi);
      tool_p3d_matrix(tq, Mc);
      const float i0 = 0.0f * Mi[0] + 1.0f * Mi[3] + 0.0f * Mi[6], i1 = 0.0f * Mi[1] + 1.0f * Mi[4] + 0.0f * Mi[7], i2 = 0.0f * Mi[2] + 1.0f * Mi[5] + 0.0f * Mi[8];
      const float c0 = Mc[0] * i0 + Mc[1] * i1 + Mc[2] * i2, c1 = Mc[3] * i0 + Mc[4] * i1 + Mc[5] * i2, c2 = Mc[6] * i0 + Mc[7] * i1 + Mc[8] * i2;
      const float angle_difference = (0.0f * c0 + 1.0f * c1 + 0.0f * c2) - 1.0f;
      sc = -angle_difference >= 1.95f ? 1.0f : 0.0f;
      float cl = dist - 0.4f; if (cl < 0.0f) cl = 0.0f;
      float cr = rd - 0.5f; if (cr < 0.0f) cr = 0.0f;
      const float up = clampf(tp.z - 0.6f, 0.0f, 0.2f);
      rew[e] = sdx_exp(-1.0f * (5.0f * cl + cr)) * (1.0f + 10.0f * up) + (rd < 0.5f ? 1.0f : 0.0f);
      successes[e] = sc;
    }
    reset[e] = rs;
    if (rs) { atomicAdd(red_count, 1); if (sc != 0.0f) atomicAdd(red_sum, sc); }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < (2 * TOOL_OBS + 31) / 32; ++i) { const int k = lane + 32 * i; if (k < 2 * TOOL_OBS) o[TOOL_OBS + k] = ho[i]; }
#pragma unroll
  for (int i = 0; i < (2 * STATE_FRAME + 31) / 32; ++i) { const int k = lane + 32 * i; if (k < 2 * STATE_FRAME) s[STATE_FRAME + k] = hs[i]; }
  // slot 141 of the privileged frame is never written by the reference (TG:1308, TO:1172): left untouched
  for (int k = lane; k < TOOL_OBS; k += 32) o[k] = fo[wid][k];
  for (int k = lane; k < STATE_FRAME; k += 32) if (k != 141) s[k] = fs[wid][k];
}
