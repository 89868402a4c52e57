// sdx_task_search.cuh -- the per-env task ops of BlockAssemblySearch (SDX_TASK_SEARCH; BASELINE configs[0]) as fused kernels.
// SE = tasks/block_assembly/allegro_hand_block_assembly_search.py.  Scene / contact step / finger drives: see scene.py; camera
// features: sdx_camera.cuh.  These are Search's own:
//   k_search_pre_physics  : clamped finger EMA + arm IK 24 cm above the target                        (SE:1546-1596)
//   k_search_post_physics : 62-slot observation frame, 175-slot privileged frame, the gate's 10 x 65 input, reward, reset flags
//                                                                                                     (SE:1036-1245, 1660-1712)
//   k_search_hand_pose    : hand teleported to the default / prepare pose                             (SE:990-998, 1405-1410, 1483-1493)
//   k_search_emergence    : 5 x (pixels now - pixels at the last render)                              (SE:1640-1646)
//   k_search_bank_slots/_write : banking of the dug-out heaps + hand states, env order, wrap-around   (SE:1305-1340)
//   k_search_reset        : state writes of reset_idx / post_reset                                    (SE:1380-1433, 1483-1496)
// The arithmetic is the oracle's (oracle/sdx_oracle.c "BlockAssemblySearch"), operation for operation.
#pragma once
#include "sdx_task_orient.cuh"

#define SEARCH_TVOBS 650
__device__ __forceinline__ float u11(uint32_t r) { return (float)(r >> 8) * (2.0f / 16777216.0f) - 1.0f; }
__device__ __constant__ int SEARCH_PIXEL_THRESHOLD[8] = {20, 20, 15, 20, 20, 30, 30, 20};   // SE:1290

__global__ void __launch_bounds__(128)
k_search_pre_physics(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ actions_in, float* __restrict__ actions,
                     float* __restrict__ dof, const float* __restrict__ link, const float* __restrict__ jac7,
                     const float* __restrict__ brick) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float a[23], cur[23], Jl[42];
  float* d = dof + (size_t)e * 72;
  for (int k = 0; k < 23; ++k) { a[k] = clampf(actions_in[23 * (size_t)e + k], -1.0f, 1.0f); actions[23 * (size_t)e + k] = a[k]; }   // VR:166
  for (int k = 0; k < 42; ++k) Jl[k] = jac7[42 * (size_t)e + k];
  for (int i = 0; i < 16; ++i) {
    float t = scalef(a[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
    float c = S->act_moving_average * t + (1.0f - S->act_moving_average) * d[48 + 7 + i];
    cur[7 + i] = clampf(c, S->dof_lo[7 + i], S->dof_hi[7 + i]);
  }
  float tg[13];
  brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
  const float* hb = link + ((size_t)e * SDX_NL + 7) * 13;
  float dpose[6];
  dpose[0] = (tg[0] - hb[0]) - 0.18f; dpose[1] = tg[1] - hb[1]; dpose[2] = (tg[2] - hb[2]) + 0.24f;
  q4 want = Q4(S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]);
  v3 re = orientation_error(want, Q4(hb[3], hb[4], hb[5], hb[6]));
  dpose[3] = re.x; dpose[4] = re.y; dpose[5] = re.z;
  float u[7];
  control_ik(Jl, dpose, u);
  for (int j = 0; j < 7; ++j) cur[j] = d[j] + u[j];
  for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
}

// one warp per env: lane 0 evaluates the frames into shared memory, all lanes shift the gate input and write the rows
__global__ void __launch_bounds__(32 * POST_WARPS)
k_search_post_physics(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ dof,
                      const float* __restrict__ link, const float* __restrict__ netf, const float* __restrict__ actions,
                      const float* __restrict__ target_init, const int* __restrict__ seg, int64_t* __restrict__ progress,
                      int64_t* __restrict__ reset, float* __restrict__ obs, float* __restrict__ states, float* __restrict__ tvobs,
                      float* __restrict__ rew, float* __restrict__ finger_dist_out, const float* __restrict__ successes,
                      int* __restrict__ red_count, float* __restrict__ red_sum) {
  __shared__ float fo[POST_WARPS][48];
  __shared__ float fs[POST_WARPS][176];
  __shared__ float ft[POST_WARPS][8];     // camera-frame target quaternion, centre / 128 x2, pixels / 100
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * POST_WARPS + wid;
  if (e >= n) return;
  float* o = obs + (size_t)e * 3 * OR_OBS_FRAME;
  float* s = states + (size_t)e * 3 * STATE_FRAME;
  float* tvo = tvobs + (size_t)e * SEARCH_TVOBS;
  float ht[(9 * 65 + 31) / 32];           // the nine newer frames of the gate input, moved one frame older below
#pragma unroll
  for (int i = 0; i < (9 * 65 + 31) / 32; ++i) { int k = lane + 32 * i; ht[i] = k < 9 * 65 ? tvo[65 + k] : 0.0f; }
  if (lane == 0) {
    float* f = fo[wid]; float* g = fs[wid];
    int64_t pg = progress[e] + 1;
    progress[e] = pg;
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
    q4 hq = Q4(hb[3], hb[4], hb[5], hb[6]);
    q4 cq0 = Q4(S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]);
    q4 cvq = qmul(qconj(qmul(hq, cq0)), tq);
    float contacts = 0.0f;
    for (int k = 0; k < 7; ++k) {
      const float* fk = netf + ((size_t)e * SDX_NL + k) * 3;
      float nf = sqrtf(vdot(V3(fk[0], fk[1], fk[2]), V3(fk[0], fk[1], fk[2])));
      contacts = contacts + (nf >= 0.1f ? 1.0f : 0.0f);
    }
    const float* ti = target_init + 7 * e;
    const float sx = (float)seg[3 * e + 1] / 128.0f, sy = (float)seg[3 * e + 2] / 128.0f, sn = (float)seg[3 * e] / 100.0f;
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
    for (int k = 96; k < 120; ++k) g[k] = 0.0f;
    g[120] = sx; g[121] = sy; g[122] = sn;
    for (int k = 0; k < 6; ++k) g[123 + k] = hb[7 + k];
    for (int k = 0; k < 4; ++k) { g[129 + k] = ff[3 + k]; g[139 + k] = mf[3 + k]; g[149 + k] = rf[3 + k]; g[159 + k] = th[3 + k]; }
    for (int k = 0; k < 6; ++k) { g[133 + k] = ff[7 + k]; g[143 + k] = mf[7 + k]; g[153 + k] = rf[7 + k]; g[163 + k] = th[7 + k]; }
    for (int k = 0; k < 6; ++k) g[169 + k] = tg[7 + k];
    ft[wid][0] = cvq.x; ft[wid][1] = cvq.y; ft[wid][2] = cvq.z; ft[wid][3] = cvq.w; ft[wid][4] = sx; ft[wid][5] = sy; ft[wid][6] = sn;
    // reward / reset flags (SE:1668-1712)
    float dist_rew = -0.2f * fdist; if (dist_rew > -0.06f) dist_rew = -0.06f;
    float asq = 0.0f;
    for (int k = 0; k < 23; ++k) { float ak = actions[23 * (size_t)e + k]; asq = asq + ak * ak; }
    float action_penalty = asq * 0.005f;
    float up = clampf(tp.z - ti[2], 0.0f, 0.1f) * 1000.0f - clampf(tp.x - ti[0], 0.0f, 0.1f) * 1000.0f - clampf(tp.y - ti[1], 0.0f, 0.1f) * 1000.0f;
    rew[e] = (((dist_rew - contacts) + 0.0f) - action_penalty) + up;
    int64_t rs = reset[e];
    if (fdist <= -1.0f) rs = 1;
    if ((float)pg >= (float)S->max_episode_length - 1.0f) rs = 1;
    reset[e] = rs;
    if (rs) { atomicAdd(red_count, 1); float sc = successes[e]; if (sc != 0.0f) atomicAdd(red_sum, sc); }
  }
  __syncwarp();
  if (lane < 16) { o[lane] = fo[wid][lane]; o[30 + lane] = fo[wid][16 + lane]; o[46 + lane] = fo[wid][32 + lane]; }
  for (int k = lane; k < 175; k += 32) if (k != 95) s[k] = fs[wid][k];                      // slot 95 is never written (SE:1182-1185)
#pragma unroll
  for (int i = 0; i < (9 * 65 + 31) / 32; ++i) { int k = lane + 32 * i; if (k < 9 * 65) tvo[k] = ht[i]; }
  // newest frame = this step's obs[0:62] with slots 26..29 = camera-frame target quaternion, then the three camera features
  __syncwarp();
  for (int k = lane; k < 65; k += 32) {
    float v;
    if (k >= 62) v = ft[wid][4 + (k - 62)];
    else if (k >= 26 && k < 30) v = ft[wid][k - 26];
    else if (k < 16) v = fo[wid][k];
    else if (k >= 30 && k < 46) v = fo[wid][16 + (k - 30)];
    else if (k >= 46) v = fo[wid][32 + (k - 46)];
    else v = o[k];                                                                       // slots 16..25: whatever obs_buf holds
    tvo[9 * 65 + k] = v;
  }
}

__global__ void k_search_hand_pose(const sdx_scene_t* __restrict__ S, int n, const int64_t* __restrict__ mask, int which, float* __restrict__ dof) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * SDX_ND) return;
  int e = i / SDX_ND, j = i % SDX_ND;
  if (mask && !mask[e]) return;
  const float v = which ? S->prepare_dof[j] : S->default_dof[j];
  float* d = dof + (size_t)e * 72;
  d[j] = v; d[24 + j] = 0.0f; d[48 + j] = v;
}

__global__ void k_search_emergence(int n, const int* __restrict__ seg, float* __restrict__ last_pixels, float* __restrict__ emergence, int baseline) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float pix = (float)seg[3 * e];
  if (!baseline) emergence[e] = (pix - last_pixels[e]) * 5.0f;
  last_pixels[e] = pix;
}

// banking, step 1 (see k_orient_bank_slots): one block per brick type ranks its envs whose target shows enough pixels
__global__ void __launch_bounds__(256)
k_search_bank_slots(int n, const int* __restrict__ seg, int* __restrict__ index, int wrap, int* __restrict__ slot) {
  __shared__ int cnt[256];
  __shared__ int base, total;
  const int ty = blockIdx.x, tid = threadIdx.x;
  const int m = (n - ty + 7) / 8;
  const int per = (m + 255) / 256;
  const int i0 = tid * per, i1 = min(m, i0 + per);
  int c = 0;
  for (int i = i0; i < i1; ++i) {
    int e = ty + 8 * i;
    bool ok = seg[3 * e] > SEARCH_PIXEL_THRESHOLD[ty];
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
// step 2: one block per banked env writes its 72 free-brick root rows and the hand's DoF state
__global__ void __launch_bounds__(96)
k_search_bank_write(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ dof,
                    const int* __restrict__ slot, float* __restrict__ rows_out, float* __restrict__ hand_out, int wrap) {
  const int e = blockIdx.x, b = threadIdx.x;
  if (e >= n) return;
  const int sl = slot[e];
  if (sl < 0) return;
  const size_t at = ((size_t)(e % 8)) * (wrap + 1) + sl;
  if (b < NB) {
    float row[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, b, row);
    float* dst = rows_out + (at * NB + b) * 13;
    for (int k = 0; k < 13; ++k) dst[k] = row[k];
  } else if (b - NB < SDX_ND) {
    const int j = b - NB;
    const float* d = dof + (size_t)e * 72;
    hand_out[at * 46 + 2 * j] = d[j]; hand_out[at * 46 + 2 * j + 1] = d[24 + j];
  }
}

__global__ void __launch_bounds__(128)
k_search_reset(const sdx_scene_t* __restrict__ S, int n, uint64_t seed, int phase, float* __restrict__ brick, float* __restrict__ dof,
               float* __restrict__ target_init, int64_t* __restrict__ progress, int64_t* __restrict__ reset,
               float* __restrict__ successes, int* __restrict__ episode, int* __restrict__ wsn, unsigned char* __restrict__ slp) {
  const int e = blockIdx.x, tid = threadIdx.x;
  if (e >= n || !reset[e]) return;
  float* B = brick + (size_t)e * 13 * NB;
  float* d = dof + (size_t)e * 72;
  const int ep = episode[e];
  __syncthreads();   // everyone has read reset[e] / episode[e] before thread 127 rewrites them
  if (phase == 0 && tid < NB) {
    uint32_t r[4];
    philox(seed, (uint32_t)e, (uint32_t)ep, 16u + (uint32_t)tid, r);
    float row[13];
    for (int k = 0; k < 7; ++k) row[k] = S->brick_init[tid * 13 + k];
    for (int k = 7; k < 13; ++k) row[k] = 0.0f;
    row[0] = row[0] + u11(r[0]) * 0.02f;
    row[1] = row[1] + u11(r[1]) * 0.02f;
    if (tid == target_brick(e)) {
      uint32_t q[4];
      philox(seed, (uint32_t)e, (uint32_t)ep, 2u, q);
      float rr = u11(q[0]);
      row[0] = 0.25f + rr * 0.2f; row[1] = 0.19f + rr * 0.15f; row[2] = 0.9f;
    }
    brick_from_root_row(S, B, tid, row);
    slp[(size_t)e * NB + tid] = 0;
  }
  if (phase <= 1 && tid >= 96 && tid < 96 + SDX_ND) {
    const int j = tid - 96;
    const float v = phase == 0 ? S->default_dof[j] : S->prepare_dof[j];
    d[j] = v; d[24 + j] = 0.0f; d[48 + j] = v;
  }
  if (tid == 127) {
    if (phase == 0) { wsn[2 * e] = 0; wsn[2 * e + 1] = 0; episode[e] = ep + 1; }
    if (phase == 1) {
      float tg[13];
      brick_root_row(S, B, target_brick(e), tg);
      for (int k = 0; k < 7; ++k) target_init[7 * e + k] = tg[k];
    }
    if (phase == 2) { progress[e] = 0; reset[e] = 0; successes[e] = 0.0f; }
  }
}
