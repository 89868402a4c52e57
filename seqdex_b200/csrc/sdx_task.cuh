// sdx_task.cuh -- the per-env task ops of BlockAssemblyGraspSim as fused kernels:
//   k_bank_terminal : reset_idx's terminal-state banking        (GS:1399-1445)
//   k_reset         : reset_idx                                  (GS:1460-1553)
//   k_pre_physics   : actions -> DoF targets, 6x7 DLS IK         (GS:1570-1638, 1796-1804)
//   k_post_physics  : observations + privileged states + reward + reset flags
//                     (GS:1090-1332, 1706-1776), one warp per env, coalesced row writes
//   k_tvalue        : GraspInsertTValue MLP + sigmoid, one fused multiply-add per weight (TVF:30-46, GS:1200-1201)
//   facade kernels  : Isaac-Gym-shaped tensors (refresh_* / set_*_indexed, GS:1091-1095, 1514-1545)
#pragma once
#include "sdx_math.cuh"
#include "../../include/seqdex_b200.h"

#ifndef NB
#define NB SDX_MAX_BRICKS
#endif
#define OBS_FRAME SDX_OBS_FRAME
#define STATE_FRAME SDX_STATE_FRAME

__device__ __forceinline__ int target_brick(int env) { int s = env % 8; return (s == 3 || s == 4 || s == 7) ? 0 : s; }  // GS:962-975

// COM-frame brick block -> Isaac Gym root row of brick b
__device__ __forceinline__ void brick_root_row(const sdx_scene_t* __restrict__ S, const float* __restrict__ B, int b, float* row) {
  v3 x = V3(B[0 * NB + b], B[1 * NB + b], B[2 * NB + b]);
  q4 q = Q4(B[3 * NB + b], B[4 * NB + b], B[5 * NB + b], B[6 * NB + b]);
  v3 v = V3(B[7 * NB + b], B[8 * NB + b], B[9 * NB + b]);
  v3 w = V3(B[10 * NB + b], B[11 * NB + b], B[12 * NB + b]);
  v3 off = qrot(q, V3(S->br_coff[3 * b], S->br_coff[3 * b + 1], S->br_coff[3 * b + 2]));
  v3 p = vsub(x, off);
  v3 vr = vsub(v, vcross(w, off));
  row[0] = p.x; row[1] = p.y; row[2] = p.z; row[3] = q.x; row[4] = q.y; row[5] = q.z; row[6] = q.w;
  row[7] = vr.x; row[8] = vr.y; row[9] = vr.z; row[10] = w.x; row[11] = w.y; row[12] = w.z;
}
__device__ __forceinline__ void brick_from_root_row(const sdx_scene_t* __restrict__ S, float* __restrict__ B, int b, const float* row) {
  q4 q = Q4(row[3], row[4], row[5], row[6]);
  v3 off = qrot(q, V3(S->br_coff[3 * b], S->br_coff[3 * b + 1], S->br_coff[3 * b + 2]));
  v3 w = V3(row[10], row[11], row[12]);
  v3 x = vadd(V3(row[0], row[1], row[2]), off);
  v3 v = vadd(V3(row[7], row[8], row[9]), vcross(w, off));
  B[0 * NB + b] = x.x; B[1 * NB + b] = x.y; B[2 * NB + b] = x.z;
  B[3 * NB + b] = q.x; B[4 * NB + b] = q.y; B[5 * NB + b] = q.z; B[6 * NB + b] = q.w;
  B[7 * NB + b] = v.x; B[8 * NB + b] = v.y; B[9 * NB + b] = v.z;
  B[10 * NB + b] = w.x; B[11 * NB + b] = w.y; B[12 * NB + b] = w.z;
}

// ---------------------------------------------------------------- FK-only refresh (creation, reset_all)
__global__ void k_refresh_links(const sdx_scene_t* __restrict__ S, const float* __restrict__ dof, float* __restrict__ link,
                                float* __restrict__ jac7, int n) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float* d = dof + (size_t)e * 72;
  v3 lx[SDX_NL]; q4 lq[SDX_NL]; v3 ja[SDX_ND], jo[SDX_ND];
  lx[0] = V3(S->base_pos[0], S->base_pos[1], S->base_pos[2]);
  lq[0] = Q4(S->base_quat[0], S->base_quat[1], S->base_quat[2], S->base_quat[3]);
  for (int j = 0; j < SDX_ND; ++j) {
    int L = j + 1, P = S->body_parent[L];
    q4 qf = Q4(S->joint_quat[4 * j], S->joint_quat[4 * j + 1], S->joint_quat[4 * j + 2], S->joint_quat[4 * j + 3]);
    v3 ax = V3(S->joint_axis[3 * j], S->joint_axis[3 * j + 1], S->joint_axis[3 * j + 2]);
    q4 qj = qmul(lq[P], qf);
    v3 x = vadd(lx[P], qrot(lq[P], V3(S->joint_xyz[3 * j], S->joint_xyz[3 * j + 1], S->joint_xyz[3 * j + 2])));
    float s, c;
    sdx_sincos(0.5f * d[j], &s, &c);
    lq[L] = qmul(qj, Q4(ax.x * s, ax.y * s, ax.z * s, c));
    lx[L] = x; ja[j] = qrot(qj, ax); jo[j] = x;
  }
  for (int L = 0; L < SDX_NL; ++L) {
    v3 w = V3(0.0f, 0.0f, 0.0f), v = V3(0.0f, 0.0f, 0.0f);
    unsigned m = S->link_anc_mask[L];
    for (int j = 0; j < SDX_ND; ++j)
      if (m & (1u << j)) {
        w = vadd(w, vscale(ja[j], d[24 + j]));
        v = vadd(v, vscale(vcross(ja[j], vsub(lx[L], jo[j])), d[24 + j]));
      }
    float* o = link + ((size_t)e * SDX_NL + L) * 13;
    o[0] = lx[L].x; o[1] = lx[L].y; o[2] = lx[L].z; o[3] = lq[L].x; o[4] = lq[L].y; o[5] = lq[L].z; o[6] = lq[L].w;
    o[7] = v.x; o[8] = v.y; o[9] = v.z; o[10] = w.x; o[11] = w.y; o[12] = w.z;
  }
  float* J = jac7 + (size_t)e * 42;
  for (int j = 0; j < 7; ++j) {
    v3 lin = vcross(ja[j], vsub(lx[7], jo[j]));
    J[0 * 7 + j] = lin.x; J[1 * 7 + j] = lin.y; J[2 * 7 + j] = lin.z;
    J[3 * 7 + j] = ja[j].x; J[4 * 7 + j] = ja[j].y; J[5 * 7 + j] = ja[j].z;
  }
}

// ---------------------------------------------------------------- reset_idx
// Terminal-state banking: per brick type (env % 8), the resetting envs that pass the gate
// (target y < 0, finger_dist < 0.6, tvalue > 0.8) append (hand DoF state, target root row) to a ring
// of 5001 slots in ENV ORDER -- slot_k = (index + k) mod 5001, exactly the reference's sequential loop.
__global__ void __launch_bounds__(256)
k_bank_terminal(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ dof,
                const int64_t* __restrict__ reset, const float* __restrict__ finger_dist, const float* __restrict__ tvalue,
                float* __restrict__ gb_hand, float* __restrict__ gb_obj, int* __restrict__ gb_index) {
  __shared__ int cnt[256];
  const int ty = blockIdx.x, tid = threadIdx.x;
  const int m = (n - ty + 7) / 8;                  // envs of this type: e = ty + 8 i
  const int per = (m + 255) / 256;
  const int i0 = tid * per, i1 = min(m, i0 + per);
  int c = 0;
  for (int i = i0; i < i1; ++i) {
    int e = ty + 8 * i;
    if (!reset[e]) continue;
    float row[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), row);
    if (row[1] < 0.0f && finger_dist[e] < 0.6f && tvalue[e] > 0.8f) c++;
  }
  cnt[tid] = c;
  __syncthreads();
  __shared__ int base, total;
  if (tid == 0) {
    int o = 0;
    for (int t = 0; t < 256; ++t) { int v = cnt[t]; cnt[t] = o; o += v; }
    base = gb_index[ty]; total = o;
  }
  __syncthreads();
  int k = cnt[tid];
  for (int i = i0; i < i1; ++i) {
    int e = ty + 8 * i;
    if (!reset[e]) continue;
    float row[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), row);
    if (row[1] < 0.0f && finger_dist[e] < 0.6f && tvalue[e] > 0.8f) {
      int slot = (base + k) % 5001;
      float* hd = gb_hand + ((size_t)ty * SDX_GRASP_BANK + slot) * 46;
      const float* d = dof + (size_t)e * 72;
      for (int j = 0; j < SDX_ND; ++j) { hd[2 * j] = d[j]; hd[2 * j + 1] = d[24 + j]; }
      float* ob = gb_obj + ((size_t)ty * SDX_GRASP_BANK + slot) * 13;
      for (int q = 0; q < 13; ++q) ob[q] = row[q];
      k++;
    }
  }
  __syncthreads();
  if (tid == 0) gb_index[ty] = (base + total) % 5001;
}

// t-value training data (GS:1402-1438 with save_hdf5): every resetting env appends its gate input (camera-frame target
// quaternion) to the success ring when the grasp is banked, to the failure ring otherwise -- in ENV ORDER (block scan).
__global__ void __launch_bounds__(1024)
k_tv_dataset(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const int64_t* __restrict__ reset,
             const float* __restrict__ finger_dist, const float* __restrict__ tvalue, const float* __restrict__ qcam,
             float* __restrict__ succ, float* __restrict__ fail, long long* __restrict__ counts, int cap) {
  __shared__ int cs[1024], cf[1024];
  __shared__ long long base[2], endc[2];
  const int tid = threadIdx.x, per = (n + 1023) / 1024;
  const int i0 = tid * per, i1 = min(n, i0 + per);
  int ns = 0, nf = 0;
  for (int e = i0; e < i1; ++e) {
    if (!reset[e]) continue;
    float row[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), row);
    if (row[1] < 0.0f && finger_dist[e] < 0.6f && tvalue[e] > 0.8f) ns++; else nf++;
  }
  cs[tid] = ns; cf[tid] = nf;
  __syncthreads();
  if (tid == 0) {
    int os = 0, of = 0;
    for (int t = 0; t < 1024; ++t) { int a = cs[t], b = cf[t]; cs[t] = os; cf[t] = of; os += a; of += b; }
    base[0] = counts[0]; base[1] = counts[1];
    counts[0] = endc[0] = base[0] + os; counts[1] = endc[1] = base[1] + of;
  }
  __syncthreads();
  long long ks = base[0] + cs[tid], kf = base[1] + cf[tid];
  for (int e = i0; e < i1; ++e) {
    if (!reset[e]) continue;
    float row[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), row);
    const bool ok = row[1] < 0.0f && finger_dist[e] < 0.6f && tvalue[e] > 0.8f;
    const long long idx = ok ? ks++ : kf++;
    if (idx < endc[ok ? 0 : 1] - cap) continue;            // overwritten by a later row of this very call: the sequential loop's last writer wins
    float* dst = (ok ? succ : fail) + 4 * (size_t)(idx % cap);
    for (int k = 0; k < 4; ++k) dst[k] = qcam[4 * e + k];
  }
}

__global__ void __launch_bounds__(128)
k_reset(const sdx_scene_t* __restrict__ S, int n, uint64_t seed, const float* __restrict__ bank, int per_type,
        float* __restrict__ brick, float* __restrict__ dof, float* __restrict__ target_init, int64_t* __restrict__ progress,
        int64_t* __restrict__ reset, float* __restrict__ successes, int* __restrict__ episode, int* __restrict__ wsn,
        unsigned char* __restrict__ slp) {
  const int e = blockIdx.x, tid = threadIdx.x;
  if (e >= n || !reset[e]) return;
  const int ep = episode[e];
  uint32_t r[4];
  philox(seed, (uint32_t)e, (uint32_t)ep, 1u, r);
  const int slot = (int)(r[0] % (uint32_t)per_type);
  const float* rows = bank + (((size_t)(e % 8)) * per_type + slot) * NB * 13;
  float* B = brick + (size_t)e * 13 * NB;
  float* d = dof + (size_t)e * 72;
  __syncthreads();   // everyone has read reset[e] / episode[e] before thread 0 rewrites them
  if (tid < NB) {
    float row[13];
    for (int k = 0; k < 7; ++k) row[k] = rows[tid * 13 + k];
    for (int k = 7; k < 13; ++k) row[k] = 0.0f;                                   // GS:1513
    brick_from_root_row(S, B, tid, row);
    slp[(size_t)e * NB + tid] = 0;                                                // setting a pose wakes the actor
  } else if (tid >= 96 && tid < 96 + 7) {
    int j = tid - 96;
    d[j] = S->prepare_arm[j]; d[24 + j] = 0.0f; d[48 + j] = S->prepare_arm[j];     // GS:1526-1529
  } else if (tid >= 96 + 7 && tid < 96 + 23) {
    int i = tid - 96 - 7;
    float v = scalef(S->finger_reset_unscaled[i], S->dof_lo[7 + i], S->dof_hi[7 + i]);   // GS:1531-1536
    d[7 + i] = v; d[24 + 7 + i] = 0.0f; d[48 + 7 + i] = v;
  }
  if (tid == 127) {
    int tb = target_brick(e);
    for (int k = 0; k < 7; ++k) target_init[7 * e + k] = rows[tb * 13 + k];         // GS:1547-1548
    progress[e] = 0; reset[e] = 0; successes[e] = 0.0f; episode[e] = ep + 1;        // GS:1550-1552
    wsn[2 * e] = 0; wsn[2 * e + 1] = 0;                                             // a new heap: no contact persists
  }
}

// ---------------------------------------------------------------- pre_physics_step
__device__ __forceinline__ void control_ik(const float* __restrict__ J, const float* dpose, float* u) {
  float A[6][6], y[6];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) {
      float s = 0.0f;
      for (int k = 0; k < 7; ++k) s = s + J[r * 7 + k] * J[c * 7 + k];
      if (r == c) s = s + 0.05f * 0.05f;
      A[r][c] = s;
    }
  for (int c = 0; c < 6; ++c) {
    float d = A[c][c];
    for (int k = 0; k < c; ++k) d = d - A[c][k] * A[c][k];
    d = sqrtf(d);
    A[c][c] = d;
    for (int r = c + 1; r < 6; ++r) {
      float s = A[r][c];
      for (int k = 0; k < c; ++k) s = s - A[r][k] * A[c][k];
      A[r][c] = s / d;
    }
  }
  for (int r = 0; r < 6; ++r) { float s = dpose[r]; for (int k = 0; k < r; ++k) s = s - A[r][k] * y[k]; y[r] = s / A[r][r]; }
  for (int r = 5; r >= 0; --r) { float s = y[r]; for (int k = r + 1; k < 6; ++k) s = s - A[k][r] * y[k]; y[r] = s / A[r][r]; }
  for (int k = 0; k < 7; ++k) { float s = 0.0f; for (int r = 0; r < 6; ++r) s = s + J[r * 7 + k] * y[r]; u[k] = s; }
}

__global__ void __launch_bounds__(128)
k_pre_physics(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ actions_in, float* __restrict__ actions,
              float* __restrict__ dof, const float* __restrict__ link, const float* __restrict__ jac7,
              const int64_t* __restrict__ progress, const float* __restrict__ target_init) {
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
  float dpose[6] = {a[0] * 0.64f, a[1] * 0.64f, a[2] * 0.64f, a[3] * 0.2f, a[4] * 0.2f, a[5] * 0.2f};
  int64_t pg = progress[e];
  if (pg > 75) {
    dpose[2] = 0.2f + 0.22f + (target_init[7 * e + 2] - link[((size_t)e * SDX_NL + 7) * 13 + 2]);
    dpose[0] = 0.0f; dpose[1] = 0.0f;
  }
  float u[7];
  control_ik(Jl, dpose, u);
  for (int j = 0; j < 7; ++j) cur[j] = d[j] + u[j];
  if (pg > 100) for (int j = 0; j < 7; ++j) cur[j] = S->insert_prep0[j];
  if (pg > 125) for (int j = 0; j < 7; ++j) cur[j] = S->insert_prep1[j];
  if (pg > 75) for (int i = 7; i < 23; ++i) cur[i] = d[48 + i];
  for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
}

// ---------------------------------------------------------------- post_physics_step
#define POST_WARPS 4
__global__ void __launch_bounds__(32 * POST_WARPS)
k_post_physics(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ dof,
               const float* __restrict__ link, const float* __restrict__ actions, const float* __restrict__ target_init,
               int64_t* __restrict__ progress, int64_t* __restrict__ reset, float* __restrict__ obs, float* __restrict__ states,
               float* __restrict__ rew, float* __restrict__ qcam, float* __restrict__ finger_dist_out,
               const float* __restrict__ successes, int* __restrict__ red_count, float* __restrict__ red_sum) {
  __shared__ float fo[POST_WARPS][OBS_FRAME];
  __shared__ float fs[POST_WARPS][STATE_FRAME];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * POST_WARPS + wid;
  if (e >= n) return;
  float* o = obs + (size_t)e * 3 * OBS_FRAME;
  float* s = states + (size_t)e * 3 * STATE_FRAME;
  // history shift (GS:1330-1332, 1278-1280): read the two newest frames, then write them one slot older
  float ho[(2 * OBS_FRAME + 31) / 32], hs[(2 * STATE_FRAME + 31) / 32];
#pragma unroll
  for (int i = 0; i < (2 * OBS_FRAME + 31) / 32; ++i) { int k = lane + 32 * i; ho[i] = k < 2 * OBS_FRAME ? o[k] : 0.0f; }
#pragma unroll
  for (int i = 0; i < (2 * STATE_FRAME + 31) / 32; ++i) { int k = lane + 32 * i; hs[i] = k < 2 * STATE_FRAME ? s[k] : 0.0f; }
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
    q4 bq = Q4(S->base_quat[0], S->base_quat[1], S->base_quat[2], S->base_quat[3]);
    q4 bqi = qconj(bq); v3 bpi = vneg(qrot(bqi, V3(S->base_pos[0], S->base_pos[1], S->base_pos[2])));
    q4 hq = Q4(hb[3], hb[4], hb[5], hb[6]); v3 hp = V3(hb[0], hb[1], hb[2]);
    q4 hvq = qmul(bqi, hq); v3 hvp = vadd(qrot(bqi, hp), bpi);
    q4 cq0 = Q4(S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]);
    q4 cq = qmul(hq, cq0); v3 cp = vadd(qrot(hq, V3(S->cam_off_pos[0], S->cam_off_pos[1], S->cam_off_pos[2])), hp);
    q4 cqi = qconj(cq); v3 cpi = vneg(qrot(cqi, cp));
    q4 cvq = qmul(cqi, tq); v3 cvp = vadd(qrot(cqi, tp), cpi);
    qcam[4 * e] = cvq.x; qcam[4 * e + 1] = cvq.y; qcam[4 * e + 2] = cvq.z; qcam[4 * e + 3] = cvq.w;
    const float* ti = target_init + 7 * e;
    for (int i = 0; i < 16; ++i) f[i] = unscalef(d[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
    f[16] = hvp.x; f[17] = hvp.y; f[18] = hvp.z; f[19] = hvq.x; f[20] = hvq.y; f[21] = hvq.z; f[22] = hvq.w;
    f[23] = cvp.x; f[24] = cvp.y; f[25] = cvp.z; f[26] = cvq.x; f[27] = cvq.y; f[28] = cvq.z; f[29] = cvq.w;
    for (int i = 0; i < 16; ++i) f[30 + i] = S->vel_obs_scale * d[24 + 7 + i];
    for (int k = 0; k < 13; ++k) { f[46 + k] = ff[k]; f[59 + k] = rf[k]; f[72 + k] = mf[k]; f[85 + k] = th[k]; f[98 + k] = tg[k]; }
    for (int k = 0; k < 7; ++k) f[111 + k] = hb[k];
    for (int k = 0; k < 7; ++k) f[118 + k] = ti[k];
    f[125] = tp.x - ti[0]; f[126] = tp.y - ti[1]; f[127] = tp.z - ti[2];
    f[128] = hp.x - tp.x; f[129] = hp.y - tp.y; f[130] = hp.z - tp.z;
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
    // reward / reset (GS:1719-1755)
    float dist = nrm[0] + nrm[1] + nrm[2] + 3.0f * nrm[3];
    int64_t rs = reset[e];
    if (dist <= -1.0f) rs = 1;
    if ((float)pg >= (float)S->max_episode_length - 1.0f) rs = 1;
    float cl = dist - 0.5f; if (cl < 0.0f) cl = 0.0f;
    float dist_rew = sdx_exp(-2.0f * cl) * 0.1f;
    float up = clampf(tp.z - ti[2], 0.0f, 0.2f) * 100.0f;
    if (!(dist < 0.5f)) up = 0.0f;
    if (up > 20.0f) up = 20.0f;
    rew[e] = dist_rew + up;
    if (pg >= 75 && dist >= 0.6f) rs = 1;
    reset[e] = rs;
    if (rs) { atomicAdd(red_count, 1); float sc = successes[e]; if (sc != 0.0f) atomicAdd(red_sum, sc); }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < (2 * OBS_FRAME + 31) / 32; ++i) { int k = lane + 32 * i; if (k < 2 * OBS_FRAME) o[OBS_FRAME + k] = ho[i]; }
#pragma unroll
  for (int i = 0; i < (2 * STATE_FRAME + 31) / 32; ++i) { int k = lane + 32 * i; if (k < 2 * STATE_FRAME) s[STATE_FRAME + k] = hs[i]; }
  // slots 131 (obs) and 141 (states) are never written by the reference (GS:1328, 1253-1255): left untouched
  for (int k = lane; k < OBS_FRAME; k += 32) if (k != 131) o[k] = fo[wid][k];
  for (int k = lane; k < STATE_FRAME; k += 32) if (k != 141) s[k] = fs[wid][k];
}

// consecutive_successes EMA (GS:1771-1774) from the reductions of k_post_physics
__global__ void k_finalize(const sdx_scene_t* __restrict__ S, int* __restrict__ red_count, float* __restrict__ red_sum, float* __restrict__ consec) {
  int c = *red_count;
  if (c > 0) consec[0] = S->av_factor * (*red_sum) / (float)c + (1.0f - S->av_factor) * consec[0];
  *red_count = 0; *red_sum = 0.0f;
}

// ---------------------------------------------------------------- t-value gate
// weights on device: W1[256][4] b1[256] W2t[256][128] b2[128] W3t[128][64] b3[64] W4[2][64] b4[2]
#define TV_ENVS 4
#define TV_WARPS 4
__global__ void __launch_bounds__(32 * TV_WARPS)
k_tvalue(const float* __restrict__ wts, int n, const float* __restrict__ qcam, float* __restrict__ tvalue, float thresh) {
  __shared__ float h1[TV_WARPS][TV_ENVS][256];
  __shared__ float h2[TV_WARPS][TV_ENVS][128];
  __shared__ float h3[TV_WARPS][TV_ENVS][64];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e0 = (blockIdx.x * TV_WARPS + wid) * TV_ENVS;
  if (e0 >= n) return;
  const float* W1 = wts; const float* b1 = W1 + 1024; const float* W2t = b1 + 256; const float* b2 = W2t + 256 * 128;
  const float* W3t = b2 + 128; const float* b3 = W3t + 128 * 64; const float* W4 = b3 + 64; const float* b4 = W4 + 128;
  float x[TV_ENVS][4];
#pragma unroll
  for (int v = 0; v < TV_ENVS; ++v) {
    int e = min(e0 + v, n - 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) x[v][k] = qcam[4 * e + k];
  }
  for (int r = 0; r < 8; ++r) {
    int o = lane + 32 * r;
    float w0 = W1[o * 4], w1 = W1[o * 4 + 1], w2 = W1[o * 4 + 2], w3 = W1[o * 4 + 3], bb = b1[o];
#pragma unroll
    for (int v = 0; v < TV_ENVS; ++v) {
      float a = bb; a = fmaf(w0, x[v][0], a); a = fmaf(w1, x[v][1], a); a = fmaf(w2, x[v][2], a); a = fmaf(w3, x[v][3], a);
      h1[wid][v][o] = sdx_elu(a);
    }
  }
  __syncwarp();
  {
    float acc[4][TV_ENVS];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int v = 0; v < TV_ENVS; ++v) acc[r][v] = b2[lane + 32 * r];
    for (int k = 0; k < 256; ++k) {
      float hv[TV_ENVS];
#pragma unroll
      for (int v = 0; v < TV_ENVS; ++v) hv[v] = h1[wid][v][k];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float w = W2t[k * 128 + lane + 32 * r];
#pragma unroll
        for (int v = 0; v < TV_ENVS; ++v) acc[r][v] = fmaf(w, hv[v], acc[r][v]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int v = 0; v < TV_ENVS; ++v) h2[wid][v][lane + 32 * r] = sdx_elu(acc[r][v]);
  }
  __syncwarp();
  {
    float acc[2][TV_ENVS];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int v = 0; v < TV_ENVS; ++v) acc[r][v] = b3[lane + 32 * r];
    for (int k = 0; k < 128; ++k) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float w = W3t[k * 64 + lane + 32 * r];
#pragma unroll
        for (int v = 0; v < TV_ENVS; ++v) acc[r][v] = fmaf(w, h2[wid][v][k], acc[r][v]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int v = 0; v < TV_ENVS; ++v) h3[wid][v][lane + 32 * r] = sdx_elu(acc[r][v]);
  }
  __syncwarp();
  if (lane < TV_ENVS && e0 + lane < n) {
    float a = b4[1];
    for (int k = 0; k < 64; ++k) a = fmaf(W4[64 + k], h3[wid][lane][k], a);
    a = sdx_elu(a);
    const float v = 1.0f / (1.0f + sdx_exp(-a));
    tvalue[e0 + lane] = thresh > 0.0f ? (v > thresh ? 1.0f : 0.0f) : v;   // Orient keeps only the thresholded gate (OR:1203-1205)
  }
}

// ---------------------------------------------------------------- Isaac-Gym-shaped facade
// actor order per env (GS:907-1000): 0 hand, 1 object, 2 goal, 3 table, 4-8 bin, 9..140 legos (72 free + 60 fixed), 141 base-plate
__global__ void k_refresh_root(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick,
                               const float* __restrict__ static_rows /*[142][13]*/, float* __restrict__ root) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * SDX_ACTORS_PER_ENV) return;
  int e = i / SDX_ACTORS_PER_ENV, a = i % SDX_ACTORS_PER_ENV;
  float* o = root + (size_t)i * 13;
  if (a >= 9 && a < 9 + NB) { float row[13]; brick_root_row(S, brick + (size_t)e * 13 * NB, a - 9, row); for (int k = 0; k < 13; ++k) o[k] = row[k]; }
  else for (int k = 0; k < 13; ++k) o[k] = static_rows[a * 13 + k];
}
// rigid bodies per env: 0-23 robot links, then one body per remaining actor (actors 1..141 -> bodies 24..164)
__global__ void k_refresh_rb(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ brick, const float* __restrict__ link,
                             const float* __restrict__ static_rows, float* __restrict__ rb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * SDX_RB_PER_ENV) return;
  int e = i / SDX_RB_PER_ENV, b = i % SDX_RB_PER_ENV;
  float* o = rb + (size_t)i * 13;
  if (b < SDX_NL) { for (int k = 0; k < 13; ++k) o[k] = link[((size_t)e * SDX_NL + b) * 13 + k]; return; }
  int a = b - SDX_NL + 1;
  if (a >= 9 && a < 9 + NB) { float row[13]; brick_root_row(S, brick + (size_t)e * 13 * NB, a - 9, row); for (int k = 0; k < 13; ++k) o[k] = row[k]; }
  else for (int k = 0; k < 13; ++k) o[k] = static_rows[a * 13 + k];
}
__global__ void k_refresh_dof_state(int n, const float* __restrict__ dof, float* __restrict__ ds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * SDX_ND) return;
  int e = i / SDX_ND, j = i % SDX_ND;
  ds[2 * i] = dof[(size_t)e * 72 + j]; ds[2 * i + 1] = dof[(size_t)e * 72 + 24 + j];
}
// full jacobian [N][23][6][23] from the link rows: column j of link L is [a_j x (x_L - o_j); a_j] for ancestors
__global__ void k_refresh_jacobian(const sdx_scene_t* __restrict__ S, int n, const float* __restrict__ link, float* __restrict__ J) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * SDX_ND * SDX_ND) return;
  int e = i / (SDX_ND * SDX_ND), r = i % (SDX_ND * SDX_ND), Lm1 = r / SDX_ND, j = r % SDX_ND;
  int L = Lm1 + 1;
  float* o = J + (((size_t)e * SDX_ND + Lm1) * 6) * SDX_ND + j;
  v3 lin = V3(0.0f, 0.0f, 0.0f), ang = V3(0.0f, 0.0f, 0.0f);
  if (S->link_anc_mask[L] & (1u << j)) {
    const float* cj = link + ((size_t)e * SDX_NL + j + 1) * 13;   // child link of joint j: its frame origin = joint origin
    const float* cl = link + ((size_t)e * SDX_NL + L) * 13;
    v3 a = qrot(Q4(cj[3], cj[4], cj[5], cj[6]), V3(S->joint_axis[3 * j], S->joint_axis[3 * j + 1], S->joint_axis[3 * j + 2]));
    lin = vcross(a, vsub(V3(cl[0], cl[1], cl[2]), V3(cj[0], cj[1], cj[2])));
    ang = a;
  }
  o[0 * SDX_ND] = lin.x; o[1 * SDX_ND] = lin.y; o[2 * SDX_ND] = lin.z; o[3 * SDX_ND] = ang.x; o[4 * SDX_ND] = ang.y; o[5 * SDX_ND] = ang.z;
}
__global__ void k_set_root_indexed(const sdx_scene_t* __restrict__ S, int n_envs, float* __restrict__ brick, const float* __restrict__ root,
                                   const int32_t* __restrict__ idx, int n, unsigned char* __restrict__ slp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ai = idx[i], e = ai / SDX_ACTORS_PER_ENV, a = ai % SDX_ACTORS_PER_ENV;
  if (e >= n_envs || a < 9 || a >= 9 + NB) return;   // only free bricks carry simulation state; the rest are fixed actors
  brick_from_root_row(S, brick + (size_t)e * 13 * NB, a - 9, root + (size_t)ai * 13);
  slp[(size_t)e * NB + (a - 9)] = 0;                  // setting a pose wakes the actor (PhysX does the same)
}
__global__ void k_set_dof_indexed(int n_envs, float* __restrict__ dof, const float* __restrict__ src, const int32_t* __restrict__ idx, int n, int mode) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * SDX_ND) return;
  int e = idx[i / SDX_ND] / SDX_ACTORS_PER_ENV, j = i % SDX_ND;
  if (e >= n_envs) return;
  float* d = dof + (size_t)e * 72;
  if (mode == 0) { d[j] = src[2 * ((size_t)e * SDX_ND + j)]; d[24 + j] = src[2 * ((size_t)e * SDX_ND + j) + 1]; }   // dof_state rows
  else d[48 + j] = src[(size_t)e * SDX_ND + j];                                                                        // targets
}
__global__ void k_set_dof_targets(int n_envs, float* __restrict__ dof, const float* __restrict__ src) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_envs * SDX_ND) return;
  dof[(size_t)(i / SDX_ND) * 72 + 48 + i % SDX_ND] = src[i];
}
__global__ void k_reset_all(const sdx_scene_t* __restrict__ S, int n, float* __restrict__ brick, float* __restrict__ dof,
                            int64_t* __restrict__ progress, int64_t* __restrict__ reset) {
  const int e = blockIdx.x, tid = threadIdx.x;
  if (e >= n) return;
  float* B = brick + (size_t)e * 13 * NB;
  float* d = dof + (size_t)e * 72;
  if (tid < NB) { float row[13]; for (int k = 0; k < 13; ++k) row[k] = S->brick_init[tid * 13 + k]; brick_from_root_row(S, B, tid, row); }
  else if (tid >= 96 && tid < 96 + 7) { int j = tid - 96; d[j] = S->prepare_arm[j]; d[24 + j] = 0.0f; d[48 + j] = S->prepare_arm[j]; }
  else if (tid >= 96 + 7 && tid < 96 + 23) {
    int i = tid - 96 - 7;
    float v = scalef(S->finger_reset_unscaled[i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
    d[7 + i] = v; d[24 + 7 + i] = 0.0f; d[48 + 7 + i] = v;
  } else if (tid == 127) { progress[e] = 0; reset[e] = 1; d[23] = 0.0f; d[47] = 0.0f; d[71] = 0.0f; }
}
// VecTask clamp of obs / states into the D2H staging buffers (VR:171-175)
__global__ void k_clamp_copy(const float* __restrict__ src, float* __restrict__ dst, size_t n, float lim) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = clampf(src[i], -lim, lim);
}
// GAE sweep (rl_games discount_values; call sites RGC:1473-1478): one thread per env, reverse scan in registers
__global__ void k_gae(const float* __restrict__ rewards, const float* __restrict__ values, const float* __restrict__ dones,
                      const float* __restrict__ last_values, const float* __restrict__ last_dones, float* __restrict__ adv,
                      float* __restrict__ returns, int H, int n, float gamma, float tau) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float lastgaelam = 0.0f;
  float nnt = 1.0f - last_dones[e], nv = last_values[e];
  for (int t = H - 1; t >= 0; --t) {
    float v = values[(size_t)t * n + e];
    float delta = rewards[(size_t)t * n + e] + gamma * nv * nnt - v;
    lastgaelam = delta + gamma * tau * nnt * lastgaelam;
    adv[(size_t)t * n + e] = lastgaelam;
    returns[(size_t)t * n + e] = lastgaelam + v;
    nnt = 1.0f - dones[(size_t)t * n + e]; nv = v;
  }
}
