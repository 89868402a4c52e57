// sdx_camera.cuh -- BlockAssemblySearch's camera features without a renderer.
// SE = tasks/block_assembly/allegro_hand_block_assembly_search.py.  The reference renders a 128 x 128 SEGMENTATION image
// per env with Isaac Gym's camera sensor (SE:755-758, 873-878) and reduces it to three numbers: how many pixels show the
// target brick and the centroid (row, column) of those pixels (SE:1231-1241; the pixel count also drives the emergence
// reward, SE:1640-1646).  Only "is the nearest surface along this pixel's ray the target brick?" matters, so the image is
// never materialised: per env, the pixels inside the target's projected bounding rectangle cast one ray each against
// the oriented boxes of the scene (72 free bricks, robot shapes, statics) and the three integers are reduced in shared
// memory.  Integer outputs => bit-exact against the oracle (sdxo_segmentation_features), same arithmetic text.
#pragma once
#include "sdx_task.cuh"

#define CAM_MAX_SHAPES (SDX_MAX_BRICKS + SDX_MAX_RSHAPES + 24)

// slab test in the box frame; *t_entry = ray parameter where the ray enters the box (negative if the origin is inside)
__device__ __forceinline__ bool ray_box(v3 o, v3 d, v3 c, const float* R, v3 h, float* t_entry) {
  v3 ol = mtmul(R, vsub(o, c));
  v3 dl = mtmul(R, d);
  float tmin = -3.0e38f, tmax = 3.0e38f;
  const float oa[3] = {ol.x, ol.y, ol.z}, da[3] = {dl.x, dl.y, dl.z}, ha[3] = {h.x, h.y, h.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (da[a] == 0.0f) { if (oa[a] < -ha[a] || oa[a] > ha[a]) return false; }
    else {
      float inv = 1.0f / da[a];
      float t1 = (-ha[a] - oa[a]) * inv, t2 = (ha[a] - oa[a]) * inv;
      if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
      if (t1 > tmin) tmin = t1;
      if (t2 < tmax) tmax = t2;
    }
  }
  if (tmax < tmin || tmax < 0.0f) return false;
  *t_entry = tmin;
  return true;
}

__global__ void __launch_bounds__(128)
k_seg_features(const sdx_scene_t* __restrict__ S, int n, sdx_camera_t cam, const float* __restrict__ brick,
               const float* __restrict__ link, int* __restrict__ out) {
  __shared__ float sc[CAM_MAX_SHAPES][3], sR[CAM_MAX_SHAPES][9], sh[CAM_MAX_SHAPES][3];
  __shared__ int acc[3], rect[4], nshape;
  const int e = blockIdx.x, tid = threadIdx.x;
  if (e >= n) return;
  const int nbr = S->n_bricks, nrs = S->n_rshapes, nst = S->n_static;
  const int tb = target_brick(e);
  const float* B = brick + (size_t)e * 13 * NB;
  if (tid < 3) acc[tid] = 0;
  // shape table: [0, nbr) bricks (box centre = COM), then robot boxes, then statics
  if (tid < nbr) {
    sc[tid][0] = B[0 * NB + tid]; sc[tid][1] = B[1 * NB + tid]; sc[tid][2] = B[2 * NB + tid];
    qmat(Q4(B[3 * NB + tid], B[4 * NB + tid], B[5 * NB + tid], B[6 * NB + tid]), sR[tid]);
    sh[tid][0] = S->br_half[3 * tid]; sh[tid][1] = S->br_half[3 * tid + 1]; sh[tid][2] = S->br_half[3 * tid + 2];
  }
  if (tid < nrs) {
    const int t = nbr + tid, L = S->rs_body[tid];
    const float* lr = link + ((size_t)e * SDX_NL + L) * 13;
    q4 qL = Q4(lr[3], lr[4], lr[5], lr[6]);
    v3 x = vadd(V3(lr[0], lr[1], lr[2]), qrot(qL, V3(S->rs_c[3 * tid], S->rs_c[3 * tid + 1], S->rs_c[3 * tid + 2])));
    sc[t][0] = x.x; sc[t][1] = x.y; sc[t][2] = x.z;
    qmat(qmul(qL, Q4(S->rs_quat[4 * tid], S->rs_quat[4 * tid + 1], S->rs_quat[4 * tid + 2], S->rs_quat[4 * tid + 3])), sR[t]);
    sh[t][0] = S->rs_h[3 * tid]; sh[t][1] = S->rs_h[3 * tid + 1]; sh[t][2] = S->rs_h[3 * tid + 2];
  }
  for (int s2 = tid; s2 < nst; s2 += 128) {
    const int t = nbr + nrs + s2;
    sc[t][0] = S->st_c[3 * s2]; sc[t][1] = S->st_c[3 * s2 + 1]; sc[t][2] = S->st_c[3 * s2 + 2];
#pragma unroll
    for (int i = 0; i < 9; ++i) sR[t][i] = (i % 4 == 0) ? 1.0f : 0.0f;
    sh[t][0] = S->st_h[3 * s2]; sh[t][1] = S->st_h[3 * s2 + 1]; sh[t][2] = S->st_h[3 * s2 + 2];
  }
  __syncthreads();
  const v3 o = V3(cam.pos[0], cam.pos[1], cam.pos[2]), fw = V3(cam.fwd[0], cam.fwd[1], cam.fwd[2]);
  const v3 rt = V3(cam.right[0], cam.right[1], cam.right[2]), up = V3(cam.up[0], cam.up[1], cam.up[2]);
  const int W = cam.width, H = cam.height;
  if (tid == 0) {   // pixel rectangle that contains the target's projection (conservative; whole image if a corner is behind the camera)
    int c0 = W, c1 = -1, r0 = H, r1 = -1;
    bool whole = false;
    for (int k = 0; k < 8; ++k) {
      v3 pl = V3((k & 1) ? sh[tb][0] : -sh[tb][0], (k & 2) ? sh[tb][1] : -sh[tb][1], (k & 4) ? sh[tb][2] : -sh[tb][2]);
      v3 pw = vadd(V3(sc[tb][0], sc[tb][1], sc[tb][2]), mmul(sR[tb], pl));
      v3 dv = vsub(pw, o);
      float depth = vdot(dv, fw);
      if (!(depth > 1.0e-4f)) { whole = true; break; }
      float px = (vdot(dv, rt) / depth) / cam.inv_focal + 0.5f * (float)W;
      float py = -(vdot(dv, up) / depth) / cam.inv_focal + 0.5f * (float)H;
      int ci = (int)floorf(px), ri = (int)floorf(py);
      c0 = min(c0, ci); c1 = max(c1, ci); r0 = min(r0, ri); r1 = max(r1, ri);
    }
    if (whole) { c0 = 0; c1 = W - 1; r0 = 0; r1 = H - 1; }
    rect[0] = max(c0 - 2, 0); rect[1] = min(c1 + 2, W - 1); rect[2] = max(r0 - 2, 0); rect[3] = min(r1 + 2, H - 1);
    nshape = nbr + nrs + nst;
  }
  __syncthreads();
  const int c0 = rect[0], c1 = rect[1], r0 = rect[2], r1 = rect[3];
  const int rw = c1 - c0 + 1, rh = r1 - r0 + 1;
  int cnt = 0, sr = 0, scol = 0;
  if (rw > 0 && rh > 0) {
    const v3 tc = V3(sc[tb][0], sc[tb][1], sc[tb][2]), th = V3(sh[tb][0], sh[tb][1], sh[tb][2]);
    for (int p = tid; p < rw * rh; p += 128) {
      const int r = r0 + p / rw, c = c0 + p % rw;
      const float sx = (((float)c + 0.5f) - 0.5f * (float)W) * cam.inv_focal;
      const float sy = -((((float)r + 0.5f) - 0.5f * (float)H) * cam.inv_focal);
      const v3 d = vadd(vadd(fw, vscale(rt, sx)), vscale(up, sy));
      float tt;
      if (!ray_box(o, d, tc, sR[tb], th, &tt)) continue;
      bool occluded = false;
      for (int s2 = 0; s2 < nshape && !occluded; ++s2) {
        if (s2 == tb) continue;
        float ts;
        if (ray_box(o, d, V3(sc[s2][0], sc[s2][1], sc[s2][2]), sR[s2], V3(sh[s2][0], sh[s2][1], sh[s2][2]), &ts) && ts < tt) occluded = true;
      }
      if (!occluded) { cnt++; sr += r; scol += c; }
    }
  }
  if (cnt) { atomicAdd(&acc[0], cnt); atomicAdd(&acc[1], sr); atomicAdd(&acc[2], scol); }
  __syncthreads();
  if (tid == 0) {
    const int c = acc[0];
    out[3 * e] = c;                                                              // segmentation_object_point_num (SE:1241)
    out[3 * e + 1] = c > 0 ? (int)((float)acc[1] / (float)c) : 0;                // int(mean(row index))    (SE:1235)
    out[3 * e + 2] = c > 0 ? (int)((float)acc[2] / (float)c) : 0;                // int(mean(column index)) (SE:1236)
  }
}
