// sdx_task_orient.cuh -- the per-env task ops of BlockAssemblyOrient (SDX_TASK_ORIENT) as fused kernels.
// OR = tasks/block_assembly/allegro_hand_block_assembly_orient.py.  Scene, contact step, t-value gate and the privileged
// state frame are GraspSim's (sdx_task.cuh); these are Orient's own:
//   k_orient_pre_physics   : finger EMA + object-centric arm IK                         (OR:1711-1778, 1922-1934)
//   k_orient_post_physics  : 62-slot observation frame, state frame, reward, reset flags (OR:1087-1326, 1843-1907)
//   k_orient_arm_script    : the two scripted arm motions of reset_idx / post_reset      (OR:1430-1455, 1659-1690)
//   k_orient_bank_slots/_write : banking of the re-oriented heaps, env order, wrap-around (OR:1465-1481)
//   k_orient_reset         : state writes of reset_idx / post_reset                      (OR:1523-1610, 1623-1645)
// The arithmetic is the oracle's (oracle/sdx_oracle.c "BlockAssemblyOrient"), operation for operation.
#pragma once
#include "sdx_task.cuh"

#define OR_OBS_FRAME SDX_ORIENT_OBS_FRAME

__device__ __forceinline__ float sgnf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }   // torch.sign
__device__ __forceinline__ v3 orientation_error(q4 desired, q4 current) {                                    // OR:1922-1925
  q4 r = qmul(desired, qconj(current));
  float sg = sgnf(r.w);
  return V3(r.x * sg, r.y * sg, r.z * sg);
}
__device__ __forceinline__ float z_align(q4 q) {                                                             // OR:1857-1860
  float d = qrot(q, V3(0.0f, 0.0f, 1.0f)).z;
  return sgnf(d) * (d * d);
}

// how many envs have their reset flag set (the reference's reset_buf.nonzero(), OR:1698)
__global__ void k_count_flags(const int64_t* __restrict__ reset, int n, int* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  int f = (e < n && reset[e]) ? 1 : 0;
  unsigned b = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, __popc(b));
}

__global__ void __launch_bounds__(128)
k_orient_pre_physics(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ actions_in, float* __restrict__ actions,
                     float* __restrict__ dof, const float* __restrict__ link, const float* __restrict__ jac7,
                     const float* __restrict__ brick, const int64_t* __restrict__ progress, const float* __restrict__ target_init) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float a[23], cur[23], Jl[42];
  float* d = dof + (size_t)e * 72;
  for (int k = 0; k < 23; ++k) { a[k] = clampf(actions_in[23 * (size_t)e + k], -1.0f, 1.0f); actions[23 * (size_t)e + k] = a[k]; }   // VR:166
  for (int k = 0; k < 42; ++k) Jl[k] = jac7[42 * (size_t)e + k];
  for (int i = 0; i < 16; ++i) {
    float t = scalef(a[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
    cur[7 + i] = S->act_moving_average * t + (1.0f - S->act_moving_average) * d[48 + 7 + i];
  }
  float tg[13];
  brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
  const float* hb = link + ((size_t)e * SDX_NL + 7) * 13;
  float dpose[6];
  dpose[0] = (tg[0] - hb[0]) - 0.18f; dpose[1] = tg[1] - hb[1]; dpose[2] = (tg[2] - hb[2]) + 0.22f;
  int64_t pg = progress[e];
  if (pg > 75) dpose[2] = ((target_init[7 * e + 2] - hb[2]) + 0.15f) + 0.24f;                  // OR:1735
  q4 want = Q4(S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]);
  v3 re = orientation_error(want, Q4(hb[3], hb[4], hb[5], hb[6]));
  dpose[3] = re.x; dpose[4] = re.y; dpose[5] = re.z;
  float u[7];
  control_ik(Jl, dpose, u);
  for (int j = 0; j < 7; ++j) cur[j] = d[j] + u[j];
  if (pg > 75) for (int i = 7; i < 23; ++i) cur[i] = d[48 + i];                                // OR:1743
  for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
}

// one warp per env: lane 0 evaluates the frames into shared memory, all lanes move the state history and write the rows.
// count_step == 0: the bare compute_observations() inside reset_idx (OR:1461) -- no progress increment, no reward.
__global__ void __launch_bounds__(32 * POST_WARPS)
k_orient_post_physics(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ dof,
                      const float* __restrict__ link, const float* __restrict__ actions, const float* __restrict__ target_init,
                      int64_t* __restrict__ progress, int64_t* __restrict__ reset, float* __restrict__ obs, float* __restrict__ states,
                      float* __restrict__ rew, float* __restrict__ qcam, float* __restrict__ finger_dist_out,
                      const float* __restrict__ successes, int* __restrict__ red_count, float* __restrict__ red_sum, int count_step) {
  __shared__ float fo[POST_WARPS][48];
  __shared__ float fs[POST_WARPS][STATE_FRAME];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * POST_WARPS + wid;
  if (e >= n) return;
  float* o = obs + (size_t)e * 3 * OR_OBS_FRAME;
  float* s = states + (size_t)e * 3 * STATE_FRAME;
  float hs[(2 * STATE_FRAME + 31) / 32];
#pragma unroll
  for (int i = 0; i < (2 * STATE_FRAME + 31) / 32; ++i) { int k = lane + 32 * i; hs[i] = k < 2 * STATE_FRAME ? s[k] : 0.0f; }
  if (lane == 0) {
    float* f = fo[wid]; float* g = fs[wid];
    int64_t pg = progress[e];
    if (count_step) { pg = pg + 1; progress[e] = pg; }
    const float* L = link + (size_t)e * SDX_NL * 13;
    const float* d = dof + (size_t)e * 72;
    const float* hb = L + 7 * 13;
    const float* ff = L + 11 * 13; const float* mf = L + 19 * 13; const float* rf = L + 23 * 13; const float* th = L + 15 * 13;
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
    v3 tp = V3(tg[0], tg[1], tg[2]); q4 tq = Q4(tg[3], tg[4], tg[5], tg[6]);
    v3 tip[4]; const float* fsr[4] = {ff, mf, rf, th};
    float nrm[4];
    for (int i = 0; i < 4; ++i) {
      q4 fq = Q4(fsr[i][3], fsr[i][4], fsr[i][5], fsr[i][6]);
      tip[i] = vadd(V3(fsr[i][0], fsr[i][1], fsr[i][2]), qrot(fq, V3(0.0f, 0.0f, 1.0f * 0.04f)));
      v3 dd = vsub(tp, tip[i]); nrm[i] = sqrtf(vdot(dd, dd));
    }
    float fdist = nrm[0] + nrm[1] + nrm[2] + nrm[3];
    finger_dist_out[e] = fdist;
    q4 hq = Q4(hb[3], hb[4], hb[5], hb[6]); v3 hp = V3(hb[0], hb[1], hb[2]);
    q4 cq0 = Q4(S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]);
    q4 cq = qmul(hq, cq0); v3 cp = vadd(qrot(hq, V3(S->cam_off_pos[0], S->cam_off_pos[1], S->cam_off_pos[2])), hp);
    q4 cqi = qconj(cq); v3 cpi = vneg(qrot(cqi, cp));
    q4 cvq = qmul(cqi, tq); v3 cvp = vadd(qrot(cqi, tp), cpi);
    qcam[4 * e] = cvq.x; qcam[4 * e + 1] = cvq.y; qcam[4 * e + 2] = cvq.z; qcam[4 * e + 3] = cvq.w;
    const float* ti = target_init + 7 * e;
    // obs frame 0 (OR:1308-1326): slots 0-15, 30-45, 46-61 -> f[0..47]
    for (int i = 0; i < 16; ++i) {
      float us = unscalef(d[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
      float ac = actions[23 * (size_t)e + 7 + i];
      f[i] = us; f[16 + i] = ac - us; f[32 + i] = ac;
    }
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
    q4 rel = qmul(hq, qconj(tq));
    g[157] = rel.x; g[158] = rel.y; g[159] = rel.z; g[160] = rel.w;
    { v3 a = vsub(tp, tip[0]), b = vsub(tp, tip[2]), c = vsub(tp, tip[1]), dd = vsub(tp, tip[3]);
      g[161] = a.x; g[162] = a.y; g[163] = a.z; g[164] = b.x; g[165] = b.y; g[166] = b.z;
      g[167] = c.x; g[168] = c.y; g[169] = c.z; g[170] = dd.x; g[171] = dd.y; g[172] = dd.z; }
    g[173] = fdist;
    g[174] = cvp.x; g[175] = cvp.y; g[176] = cvp.z; g[177] = cvq.x; g[178] = cvq.y; g[179] = cvq.z; g[180] = cvq.w;
    g[181] = cvp.x; g[182] = cvp.y; g[183] = cvp.z; g[184] = cvq.x; g[185] = cvq.y; g[186] = cvq.z; g[187] = cvq.w;
    if (count_step) {                                                                       // OR:1852-1907
      float dist = nrm[0] + nrm[1] + nrm[2] + 3.0f * nrm[3];
      int64_t rs = reset[e];
      if (dist <= -1.0f) rs = 1;
      if ((float)pg >= (float)S->max_episode_length - 1.0f) rs = 1;
      float drew = dist - 0.4f; if (drew < 0.0f) drew = 0.0f;
      if (pg > 175) drew = 0.0f;
      float zrew = 1.0f - ((z_align(tq) + 1.0f) / 2.0f);
      rew[e] = sdx_exp(-(5.0f * zrew + 5.0f * drew));
      reset[e] = rs;
      if (rs) { atomicAdd(red_count, 1); float sc = successes[e]; if (sc != 0.0f) atomicAdd(red_sum, sc); }
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < (2 * STATE_FRAME + 31) / 32; ++i) { int k = lane + 32 * i; if (k < 2 * STATE_FRAME) s[STATE_FRAME + k] = hs[i]; }
  for (int k = lane; k < STATE_FRAME; k += 32) if (k != 141) s[k] = fs[wid][k];            // slot 141 is never written (OR:1279)
  if (lane < 16) { o[lane] = fo[wid][lane]; o[30 + lane] = fo[wid][16 + lane]; o[46 + lane] = fo[wid][32 + lane]; }
}

// scripted arm motions of the reset, envs with the reset flag set (oracle: sdxo_orient_arm_script)
__global__ void __launch_bounds__(128)
k_orient_arm_script(const sdx_scene_t* __restrict__ S, int n, const int64_t* __restrict__ reset, int mode, int iter,
                    float* __restrict__ dof, const float* __restrict__ link, const float* __restrict__ jac7,
                    const float* __restrict__ brick, const float* __restrict__ target_init) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !reset[e]) return;
  float* d = dof + (size_t)e * 72;
  const float* hb = link + ((size_t)e * SDX_NL + 7) * 13;
  float dpose[6], Jl[42];
  for (int k = 0; k < 42; ++k) Jl[k] = jac7[42 * (size_t)e + k];
  if (mode == 0) {
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
    dpose[0] = (tg[0] - hb[0]) - 0.18f; dpose[1] = tg[1] - hb[1]; dpose[2] = (tg[2] - hb[2]) + 0.42f;
  } else {
    const float* ti = target_init + 7 * e;
    float z = (ti[2] - hb[2]) + 0.22f;
    if (iter < 20) z = z + 0.2f;
    dpose[0] = (ti[0] - hb[0]) - 0.18f; dpose[1] = ti[1] - hb[1]; dpose[2] = z;
  }
  q4 want = Q4(S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]);
  v3 re = orientation_error(want, Q4(hb[3], hb[4], hb[5], hb[6]));
  dpose[3] = re.x; dpose[4] = re.y; dpose[5] = re.z;
  float u[7];
  control_ik(Jl, dpose, u);
  if (mode == 0) {
    for (int j = 0; j < 7; ++j) d[48 + j] = clampf(d[j] + u[j], S->dof_lo[j], S->dof_hi[j]);
    if (iter == 0) for (int i = 7; i < 23; ++i) d[48 + i] = d[48 + i] - 0.01f;
  } else {
    for (int j = 0; j < 7; ++j) { float t = d[j] + u[j]; d[j] = t; d[48 + j] = t; d[24 + j] = 0.0f; }
    for (int i = 0; i < 16; ++i) {
      float v = scalef(S->finger_reset_unscaled[i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
      d[7 + i] = v; d[24 + 7 + i] = 0.0f; d[48 + 7 + i] = v;
    }
  }
}

// banking, step 1: one block per brick type ranks its envs that pass the gate (env order) and advances the ring index the
// way the sequential loop does (index += 1; if index > wrap: index = 0  =>  slot_k = (base + k) mod (wrap + 1)).
// slot[e] = ring slot to write, or -1 (not banked, or overwritten later in this very call: the last writer wins).
__global__ void __launch_bounds__(256)
k_orient_bank_slots(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ finger_dist,
                    const float* __restrict__ tvalue, int* __restrict__ index, int wrap, int* __restrict__ slot) {
  __shared__ int cnt[256];
  __shared__ int base, total;
  const int ty = blockIdx.x, tid = threadIdx.x;
  const int m = (n - ty + 7) / 8;
  const int per = (m + 255) / 256;
  const int i0 = tid * per, i1 = min(m, i0 + per);
  int c = 0;
  for (int i = i0; i < i1; ++i) {
    int e = ty + 8 * i;
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
    bool ok = finger_dist[e] > 0.3f && 0.5f > tg[1] && tg[1] > 0.0f && tvalue[e] > 0.6f;
    slot[e] = ok ? 0 : -1;
    c += ok ? 1 : 0;
  }
  cnt[tid] = c;
  __syncthreads();
  if (tid == 0) {
    int o = 0;
    for (int t = 0; t < 256; ++t) { int v = cnt[t]; cnt[t] = o; o += v; }
    base = index[ty]; total = o;
  }
  __syncthreads();
  int k = cnt[tid];
  const int ring = wrap + 1;
  for (int i = i0; i < i1; ++i) {
    int e = ty + 8 * i;
    if (slot[e] < 0) continue;
    slot[e] = (k < total - ring) ? -1 : (base + k) % ring;
    k++;
  }
  __syncthreads();
  if (tid == 0) index[ty] = (base + total) % ring;
}
// step 2: one block per env writes the 72 free-brick root rows of a banked env
__global__ void __launch_bounds__(96)
k_orient_bank_write(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const int* __restrict__ slot,
                    float* __restrict__ rows_out, int wrap) {
  const int e = blockIdx.x, b = threadIdx.x;
  if (e >= n || b >= NB) return;
  const int sl = slot[e];
  if (sl < 0) return;
  float row[13];
  brick_root_row(S, brick + (size_t)e * 13 * NB, b, row);
  float* dst = rows_out + ((((size_t)(e % 8)) * (wrap + 1) + sl) * NB + b) * 13;
  for (int k = 0; k < 13; ++k) dst[k] = row[k];
}

// state writes of reset_idx / post_reset (oracle: sdxo_orient_reset), one block per env
__global__ void __launch_bounds__(128)
k_orient_reset(const sdx_scene_t* __restrict__ S, int n, uint64_t seed, const float* __restrict__ bank, int per_type, int phase,
               float* __restrict__ brick, float* __restrict__ dof, float* __restrict__ target_init, int64_t* __restrict__ progress,
               int64_t* __restrict__ reset, float* __restrict__ successes, int* __restrict__ episode, int* __restrict__ wsn,
               unsigned char* __restrict__ slp) {
  const int e = blockIdx.x, tid = threadIdx.x;
  if (e >= n || !reset[e]) return;
  float* B = brick + (size_t)e * 13 * NB;
  float* d = dof + (size_t)e * 72;
  const int ep = episode[e];
  __syncthreads();   // everyone has read reset[e] / episode[e] before thread 127 rewrites them
  if (phase == 0 && tid < NB) {
    uint32_t r[4];
    philox(seed, (uint32_t)e, (uint32_t)ep, 1u, r);
    const int range = per_type < S->bank_sample_range ? per_type : S->bank_sample_range;
    const int slot = (int)(r[0] % (uint32_t)range);
    const float* rows = bank + (((size_t)(e % 8)) * per_type + slot) * NB * 13;
    float row[13];
    for (int k = 0; k < 7; ++k) row[k] = rows[tid * 13 + k];
    for (int k = 7; k < 13; ++k) row[k] = 0.0f;                                   // OR:1570
    brick_from_root_row(S, B, tid, row);
    slp[(size_t)e * NB + tid] = 0;
  }
  if (phase <= 1) {
    if (tid >= 96 && tid < 96 + 7) {
      int j = tid - 96;
      d[j] = S->prepare_arm[j]; d[24 + j] = 0.0f; d[48 + j] = S->prepare_arm[j];   // OR:1583-1586
    } else if (tid >= 96 + 7 && tid < 96 + 23) {
      int i = tid - 96 - 7;
      float v = scalef(S->finger_reset_unscaled[i], S->dof_lo[7 + i], S->dof_hi[7 + i]);   // OR:1588-1593
      d[7 + i] = v; d[24 + 7 + i] = 0.0f; d[48 + 7 + i] = v;
    }
  }
  if (tid == 127) {
    if (phase == 0) { wsn[2 * e] = 0; wsn[2 * e + 1] = 0; episode[e] = ep + 1; }
    if (phase == 1) {
      float tg[13];
      brick_root_row(S, B, target_brick(e), tg);
      for (int k = 0; k < 7; ++k) target_init[7 * e + k] = tg[k];                  // OR:1623-1624
    }
    if (phase == 2) { progress[e] = 0; reset[e] = 0; successes[e] = 0.0f; }        // OR:1607-1609
  }
}
