// sdx_task_insert.cuh -- BlockAssemblyInsertSim (SDX_TASK_INSERT_SIM; the last link of BASELINE configs[3]) as fused kernels, one
// thread per env.  IS = tasks/block_assembly/allegro_hand_block_assembly_insert_sim.py:
//   k_insert_reset        : reset_idx -- success_buf, bricks parked, base-plate pose, banked grasp restored   (IS:1328-1493)
//   k_insert_pre_physics  : finger EMA, arm IK for (0.64 a[0:3], wrist orientation error)                      (IS:1516-1565)
//   k_insert_post_physics : 75-slot observation, 188-slot privileged state, reward, reset flags                (IS:1090-1298, 1640-1694)
// Same arithmetic, operation for operation, as oracle/sdx_oracle.c "BlockAssemblyInsertSim" (bit-exact parity); the oracle is
// pinned to the reference's own Python (tests/golden/insert_*.npz).
#pragma once
#include "sdx_task.cuh"
#include "sdx_task_orient.cuh"

#define INSERT_OBS SDX_INSERT_OBS_FRAME

// asin on [-1, 1] (cephes asinf: minimax polynomial on |x| <= 0.5, pi/2 - 2 asin(sqrt((1 - x) / 2)) beyond)
__device__ __forceinline__ float sdx_asin(float x) {
  float a = fabsf(x), z, w;
  const bool big = a > 0.5f;
  if (big) { z = 0.5f * (1.0f - a); w = sqrtf(z); } else { w = a; z = a * a; }
  float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * w + w;
  if (big) p = 1.5707963267948966f - (p + p);
  return x < 0.0f ? -p : p;
}
// the pose the held brick has to reach (IS:1124-1132): plate position, one brick height per plate level up, half a stud along y
// (and x for the 1x1), every offset rotated with the plate and added in the reference's order
__device__ __forceinline__ void insert_target(int e, const float* __restrict__ plate, v3* pos, q4* rot) {
  const q4 q = Q4(plate[3], plate[4], plate[5], plate[6]);
  v3 p = V3(plate[0], plate[1], plate[2]);
  const float lvl = (float)(1 + e % 3);
  p = vadd(p, qrot(q, V3(0.0f * (0.0375f * lvl), 0.0f * (0.0375f * lvl), 1.0f * (0.0375f * lvl))));
  if (e % 8 == 5) {
    p = vadd(p, qrot(q, V3(1.0f * 0.015f, 0.0f * 0.015f, 0.0f * 0.015f)));
    p = vadd(p, qrot(q, V3(0.0f * 0.015f, 1.0f * 0.015f, 0.0f * 0.015f)));
  } else p = vadd(p, qrot(q, V3(0.0f * 0.015f, 1.0f * 0.015f, 0.0f * 0.015f)));
  *pos = p; *rot = q;
}
__device__ __forceinline__ float rot_dist_sym(q4 tq, q4 eq) {     // IS:1656-1660
  const q4 d1 = qmul(tq, qconj(eq));
  const q4 sym = qmul(eq, Q4(0.0f, 0.0f, 1.0f, 0.0f));
  const q4 d2 = qmul(tq, qconj(sym));
  const float n1 = sqrtf(d1.x * d1.x + d1.y * d1.y + d1.z * d1.z), n2 = sqrtf(d2.x * d2.x + d2.y * d2.y + d2.z * d2.z);
  const float r1 = 2.0f * sdx_asin(n1 > 1.0f ? 1.0f : n1), r2 = 2.0f * sdx_asin(n2 > 1.0f ? 1.0f : n2);
  return r1 < r2 ? r1 : r2;
}

__global__ void k_insert_pre_physics(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ actions_in, float* __restrict__ actions,
                                     float* __restrict__ dof, const float* __restrict__ link, const float* __restrict__ jac7,
                                     float* __restrict__ rot_err) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float* a = actions_in + 23 * e;
  float* d = dof + (size_t)e * 72;
  float cur[23];
  for (int k = 0; k < 23; ++k) actions[23 * e + k] = a[k];
  for (int i = 0; i < 16; ++i) {
    const float t = scalef(a[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
    cur[7 + i] = S->act_moving_average * t + (1.0f - S->act_moving_average) * d[48 + 7 + i];
  }
  const float* hb = link + ((size_t)e * SDX_NL + 7) * 13;
  const q4 want = Q4(S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]);
  const v3 re = orientation_error(want, Q4(hb[3], hb[4], hb[5], hb[6]));
  rot_err[3 * e] = re.x; rot_err[3 * e + 1] = re.y; rot_err[3 * e + 2] = re.z;
  const float dpose[6] = {a[0] * 0.64f, a[1] * 0.64f, a[2] * 0.64f, re.x, re.y, re.z};
  float u[7];
  control_ik(jac7 + 42 * (size_t)e, dpose, u);
  for (int j = 0; j < 7; ++j) cur[j] = clampf(d[j] + u[j], S->dof_lo[j], S->dof_hi[j]);
  for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
}

__global__ void k_insert_post_physics(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ dof,
                                      const float* __restrict__ link, const float* __restrict__ actions, const float* __restrict__ target_init,
                                      const float* __restrict__ plate, const float* __restrict__ rot_err, int64_t* __restrict__ progress,
                                      int64_t* __restrict__ reset, float* __restrict__ obs, float* __restrict__ states, float* __restrict__ rew,
                                      float* __restrict__ finger_dist_out, const float* __restrict__ successes, int* __restrict__ red_count,
                                      float* __restrict__ red_sum) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int64_t pg = progress[e] + 1;
  progress[e] = pg;
  const float* L = link + (size_t)e * SDX_NL * 13;
  const float* d = dof + (size_t)e * 72;
  const float* hb = L + 7 * 13;
  const float* ff = L + 11 * 13; const float* mf = L + 19 * 13; const float* rf = L + 23 * 13; const float* th = L + 15 * 13;   // IS:166-169
  float tg[13];
  brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
  const v3 tp = V3(tg[0], tg[1], tg[2]); const q4 tq = Q4(tg[3], tg[4], tg[5], tg[6]);
  v3 tip[4]; const float* fs[4] = {ff, mf, rf, th};
  for (int i = 0; i < 4; ++i)
    tip[i] = vadd(V3(fs[i][0], fs[i][1], fs[i][2]), qrot(Q4(fs[i][3], fs[i][4], fs[i][5], fs[i][6]), V3(0.0f, 0.0f, 1.0f * 0.04f)));
  float nrm[4];
  for (int i = 0; i < 4; ++i) { const v3 dd = vsub(tp, tip[i]); nrm[i] = sqrtf(vdot(dd, dd)); }
  const float fdist = nrm[0] + nrm[1] + nrm[2] + nrm[3];
  finger_dist_out[e] = fdist;
  const q4 hq = Q4(hb[3], hb[4], hb[5], hb[6]); const v3 hp = V3(hb[0], hb[1], hb[2]);
  const q4 cq0 = Q4(S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]);
  const q4 cq = qmul(hq, cq0); const v3 cp = vadd(qrot(hq, V3(S->cam_off_pos[0], S->cam_off_pos[1], S->cam_off_pos[2])), hp);
  const q4 cqi = qconj(cq); const v3 cpi = vneg(qrot(cqi, cp));
  const q4 cvq = qmul(cqi, tq); const v3 cvp = vadd(qrot(cqi, tp), cpi);
  v3 ep; q4 eq;
  insert_target(e, plate + 7 * e, &ep, &eq);
  const float* ti = target_init + 7 * e;
  // ---- observation frame (IS:1280-1298); slots 16:23 and 60 are never written
  float* o = obs + (size_t)e * INSERT_OBS;
  for (int i = 0; i < 16; ++i) o[i] = unscalef(d[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
  for (int k = 0; k < 23; ++k) o[23 + k] = actions[23 * e + k];
  o[46] = hp.x - ep.x; o[47] = hp.y - ep.y; o[48] = hp.z - ep.z;
  { const q4 r = qmul(hq, qconj(eq)); o[49] = r.x; o[50] = r.y; o[51] = r.z; o[52] = r.w; }
  o[53] = hp.x - tp.x; o[54] = hp.y - tp.y; o[55] = hp.z - tp.z;
  { const q4 r = qmul(hq, qconj(tq)); o[56] = r.x; o[57] = r.y; o[58] = r.z; o[59] = r.w; }
  o[61] = ep.x; o[62] = ep.y; o[63] = ep.z; o[64] = eq.x; o[65] = eq.y; o[66] = eq.z; o[67] = eq.w;
  o[68] = tp.x - ep.x; o[69] = tp.y - ep.y; o[70] = tp.z - ep.z;
  { const q4 r = qmul(tq, qconj(eq)); o[71] = r.x; o[72] = r.y; o[73] = r.z; o[74] = r.w; }
  // ---- privileged frame (IS:1222-1278)
  float* s = states + (size_t)e * STATE_FRAME;
  for (int j = 0; j < 23; ++j) { s[j] = unscalef(d[j], S->dof_lo[j], S->dof_hi[j]); s[23 + j] = S->vel_obs_scale * d[24 + j]; }
  s[46] = tip[0].x; s[47] = tip[0].y; s[48] = tip[0].z;
  s[49] = tip[2].x; s[50] = tip[2].y; s[51] = tip[2].z;
  s[52] = tip[1].x; s[53] = tip[1].y; s[54] = tip[1].z;
  s[55] = tip[3].x; s[56] = tip[3].y; s[57] = tip[3].z;
  for (int k = 0; k < 23; ++k) s[58 + k] = actions[23 * e + k];
  for (int k = 0; k < 7; ++k) { s[81 + k] = hb[k]; s[88 + k] = tg[k]; }
  for (int k = 0; k < 6; ++k) s[95 + k] = hb[7 + k];
  for (int k = 0; k < 4; ++k) { s[101 + k] = ff[3 + k]; s[111 + k] = mf[3 + k]; s[121 + k] = rf[3 + k]; s[131 + k] = th[3 + k]; }
  for (int k = 0; k < 6; ++k) { s[105 + k] = ff[7 + k]; s[115 + k] = mf[7 + k]; s[125 + k] = rf[7 + k]; s[135 + k] = th[7 + k]; }
  s[141] = (float)pg / (float)S->max_episode_length;
  for (int k = 0; k < 6; ++k) s[142 + k] = tg[7 + k];
  s[148] = ti[0]; s[149] = ti[1]; s[150] = ti[2];
  s[151] = tp.x - ti[0]; s[152] = tp.y - ti[1]; s[153] = tp.z - ti[2];
  s[154] = hp.x - tp.x; s[155] = hp.y - tp.y; s[156] = hp.z - tp.z;
  { const q4 rel = qmul(hq, qconj(tq)); s[157] = rel.x; s[158] = rel.y; s[159] = rel.z; s[160] = rel.w; }
  { const v3 a = vsub(tp, tip[0]), b = vsub(tp, tip[2]), c = vsub(tp, tip[1]), dd = vsub(tp, tip[3]);
    s[161] = a.x; s[162] = a.y; s[163] = a.z; s[164] = b.x; s[165] = b.y; s[166] = b.z;
    s[167] = c.x; s[168] = c.y; s[169] = c.z; s[170] = dd.x; s[171] = dd.y; s[172] = dd.z; }
  s[173] = fdist;
  s[174] = cvp.x; s[175] = cvp.y; s[176] = cvp.z; s[177] = cvq.x; s[178] = cvq.y; s[179] = cvq.z; s[180] = cvq.w;
  s[181] = ep.x; s[182] = ep.y; s[183] = ep.z; s[184] = eq.x; s[185] = eq.y; s[186] = eq.z; s[187] = eq.w;
  // ---- reward / resets (IS:1650-1693)
  const float dist = nrm[0] + nrm[1] + nrm[2] + 3.0f * nrm[3];
  const float rd = rot_dist_sym(tq, eq);
  const v3 dp = vsub(tp, ep);
  const float pd = sqrtf(vdot(dp, dp));
  const float insert_reward = sdx_exp(-1.0f * rd - 20.0f * (pd < 0.0f ? 0.0f : pd));
  const v3 dp2 = vsub(ep, tp);
  const float bonus = (sqrtf(vdot(dp2, dp2)) < 0.02f && rd < 0.2f) ? 1.0f : 0.0f;
  int64_t rs = reset[e];
  if (dist >= 0.6f) rs = 1;
  const float* re = rot_err + 3 * e;
  if (re[0] * re[0] + re[1] * re[1] + re[2] * re[2] >= 0.03f) rs = 1;
  if ((float)pg >= (float)S->max_episode_length - 1.0f) rs = 1;
  rew[e] = bonus + insert_reward;
  reset[e] = rs;
  if (rs) { atomicAdd(red_count, 1); atomicAdd(red_sum, successes[e]); }
}

// slot_override (nullable, test hook): slot of the k-th resetting env in env order; needs a single-block launch to be ordered,
// so the tests pass slots for ALL envs indexed by env instead
__global__ void k_insert_reset(const sdx_scene_t* __restrict__ S, int n, uint64_t seed, const float* __restrict__ bank_obj,
                               const float* __restrict__ bank_hand, int per_type, int plate_yaw_idx, const int* __restrict__ slot_by_env,
                               int do_success, float* __restrict__ brick, float* __restrict__ dof, float* __restrict__ plate,
                               float* __restrict__ target_init, int64_t* __restrict__ progress, int64_t* __restrict__ reset,
                               float* __restrict__ successes, float* __restrict__ success_buf, int* __restrict__ episode, int* __restrict__ wsn,
                               unsigned char* __restrict__ slp) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !reset[e]) return;
  float* B = brick + (size_t)e * 13 * NB;
  float* d = dof + (size_t)e * 72;
  const int tb = target_brick(e);
  if (do_success) {
    float tg[13];
    brick_root_row(S, B, tb, tg);
    v3 ep; q4 eq;
    insert_target(e, plate + 7 * e, &ep, &eq);
    const float rd = rot_dist_sym(Q4(tg[3], tg[4], tg[5], tg[6]), eq);
    const v3 dp = vsub(ep, V3(tg[0], tg[1], tg[2]));
    const float ok = (sqrtf(vdot(dp, dp)) < 0.02f && rd < 0.2f) ? 1.0f : 0.0f;
    success_buf[2 * e] = ok; success_buf[2 * e + 1] = ok <= 0.5f ? 1.0f : 0.0f;
  }
  for (int b = 0; b < S->n_bricks; ++b) {
    float row[13];
    for (int k = 0; k < 7; ++k) row[k] = S->brick_init[13 * b + k];
    for (int k = 7; k < 13; ++k) row[k] = 0.0f;
    brick_from_root_row(S, B, b, row);
  }
  float* pl = plate + 7 * e;
  pl[0] = 0.25f; pl[1] = -0.2f; pl[2] = 0.618f;
  pl[3] = 0.0f; pl[4] = 0.0f; pl[5] = plate_yaw_idx ? S->insert_plate_zw[0] : 0.0f; pl[6] = plate_yaw_idx ? S->insert_plate_zw[1] : 1.0f;
  uint32_t r[4];
  philox(seed, (uint32_t)e, (uint32_t)episode[e], 1u, r);
  const int slot = slot_by_env ? slot_by_env[e] : (int)(r[0] % (uint32_t)per_type);
  const float* ob = bank_obj + (((size_t)(e % 8)) * per_type + slot) * 13;
  const float* hd = bank_hand + (((size_t)(e % 8)) * per_type + slot) * 46;
  float row[13];
  for (int k = 0; k < 7; ++k) row[k] = ob[k];
  for (int k = 7; k < 13; ++k) row[k] = 0.0f;
  brick_from_root_row(S, B, tb, row);
  for (int j = 0; j < SDX_ND; ++j) { d[j] = hd[2 * j]; d[24 + j] = 0.0f; d[48 + j] = hd[2 * j]; }
  for (int k = 0; k < 7; ++k) target_init[7 * e + k] = ob[k];
  progress[e] = 0; reset[e] = 0; successes[e] = 0.0f;
  wsn[2 * e] = 0; wsn[2 * e + 1] = 0;
  for (int b = 0; b < NB; ++b) slp[(size_t)e * NB + b] = 0;
  episode[e] += 1;
}
