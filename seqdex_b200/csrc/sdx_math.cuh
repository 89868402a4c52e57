// sdx_math.cuh -- fp32 vector / quaternion helpers of the contact-step and task kernels.
//
// Rounding contract (DESIGN.md "numerics"): this translation unit is compiled with -fmad=false, so
// every expression below rounds exactly as written (mul then add), sqrt and division are the IEEE
// correctly-rounded ones, and the only transcendental functions are the polynomial kernels defined
// here.  That makes the kernels reproducible bit for bit on any device -- and checkable against a
// scalar CPU restatement.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct v3 { float x, y, z; };
struct q4 { float x, y, z, w; };

__device__ __forceinline__ v3 V3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 vscale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
// explicit single-rounding FMAs: with -fmad=false these are the ONLY fused operations in the kernels
__device__ __forceinline__ float vdot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ v3 vcross(v3 a, v3 b) {
  return V3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ v3 vmad(v3 a, float s, v3 b) { return V3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }  // a*s + b
__device__ __forceinline__ v3 vneg(v3 a) { return V3(-a.x, -a.y, -a.z); }

// quaternion product, xyzw (same operation order as isaacgym.torch_utils.quat_mul)
__device__ __forceinline__ q4 qmul(q4 a, q4 b) {
  float x1 = a.x, y1 = a.y, z1 = a.z, w1 = a.w, x2 = b.x, y2 = b.y, z2 = b.z, w2 = b.w;
  float ww = (z1 + x1) * (x2 + y2);
  float yy = (w1 - y1) * (w2 + z2);
  float zz = (w1 + y1) * (w2 - z2);
  float xx = ww + yy + zz;
  float qq = 0.5f * (xx + (z1 - x1) * (x2 - y2));
  q4 r;
  r.w = qq - ww + (z1 - y1) * (y2 - z2);
  r.x = qq - xx + (x1 + w1) * (x2 + w2);
  r.y = qq - yy + (w1 - x1) * (y2 + z2);
  r.z = qq - zz + (z1 + y1) * (w2 - x2);
  return r;
}
__device__ __forceinline__ q4 qconj(q4 a) { q4 r; r.x = -a.x; r.y = -a.y; r.z = -a.z; r.w = a.w; return r; }
__device__ __forceinline__ q4 Q4(float x, float y, float z, float w) { q4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
// quat_apply
__device__ __forceinline__ v3 qrot(q4 q, v3 b) {
  v3 xyz = V3(q.x, q.y, q.z);
  v3 t = vscale(vcross(xyz, b), 2.0f);
  return vadd(vmad(t, q.w, b), vcross(xyz, t));
}
__device__ __forceinline__ void qmat(q4 q, float* R) {
  float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z, xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z,
        xw = q.x * q.w, yw = q.y * q.w, zw = q.z * q.w;
  R[0] = 1.0f - 2.0f * (yy + zz); R[1] = 2.0f * (xy - zw); R[2] = 2.0f * (xz + yw);
  R[3] = 2.0f * (xy + zw); R[4] = 1.0f - 2.0f * (xx + zz); R[5] = 2.0f * (yz - xw);
  R[6] = 2.0f * (xz - yw); R[7] = 2.0f * (yz + xw); R[8] = 1.0f - 2.0f * (xx + yy);
}
__device__ __forceinline__ v3 mcol(const float* R, int k) { return V3(R[k], R[3 + k], R[6 + k]); }
__device__ __forceinline__ v3 mmul(const float* R, v3 a) {
  return V3(fmaf(R[2], a.z, fmaf(R[1], a.y, R[0] * a.x)), fmaf(R[5], a.z, fmaf(R[4], a.y, R[3] * a.x)),
            fmaf(R[8], a.z, fmaf(R[7], a.y, R[6] * a.x)));
}
__device__ __forceinline__ v3 mtmul(const float* R, v3 a) {
  return V3(fmaf(R[6], a.z, fmaf(R[3], a.y, R[0] * a.x)), fmaf(R[7], a.z, fmaf(R[4], a.y, R[1] * a.x)),
            fmaf(R[8], a.z, fmaf(R[5], a.y, R[2] * a.x)));
}

// sin/cos: Cody-Waite reduction by pi/2 + cephes minimax kernels (|x| < ~100)
__device__ __forceinline__ void sdx_sincos(float x, float* s, float* c) {
  float k = rintf(x * 0.63661977236758134f);
  float r = x - k * 1.5703125f;
  r = r - k * 4.837512969970703125e-4f;
  r = r - k * 7.54978995489188e-8f;
  float z = r * r;
  float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
  float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
  int n = ((int)k) & 3;
  float ss = (n & 1) ? cp : sp;
  float cc = (n & 1) ? sp : cp;
  if (n == 1 || n == 2) cc = -cc;
  if (n >= 2) ss = -ss;
  *s = ss; *c = cc;
}
// exp: cephes expf kernel with exact 2^n scaling
__device__ __forceinline__ float sdx_exp(float x) {
  if (x > 88.0f) x = 88.0f;
  if (x < -87.0f) x = -87.0f;
  float n = rintf(x * 1.44269504088896341f);
  float r = x - n * 0.693359375f;
  r = r - n * -2.12194440e-4f;
  float z = r * r;
  float p = ((((1.9875691500e-4f * r + 1.3981999507e-3f) * r + 8.3334519073e-3f) * r + 4.1665795894e-2f) * r +
             1.6666665459e-1f) * r + 5.0000001201e-1f;
  float y = p * z + r + 1.0f;
  return y * __uint_as_float((uint32_t)((int)n + 127) << 23);
}
__device__ __forceinline__ float sdx_elu(float x) { return x > 0.0f ? x : sdx_exp(x) - 1.0f; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ float scalef(float x, float lo, float hi) { return 0.5f * (x + 1.0f) * (hi - lo) + lo; }
__device__ __forceinline__ float unscalef(float x, float lo, float hi) { return (2.0f * x - hi - lo) / (hi - lo); }

// Philox4x32-10, key = seed, counter = (c0, c1, c2, 0)
__device__ __forceinline__ void philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c[4] = {c0, c1, c2, 0u};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1,
             n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
