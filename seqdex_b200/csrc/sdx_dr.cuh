// sdx_dr.cuh -- domain randomisation, the non-physical half: the observation / action noise that
// BaseTask.step applies around the env step (BT:131-132, 149-150) with the parameters apply_randomizations prepares
// (BT:263-340).  The reference draws torch.randn_like / torch.rand_like; here the white noise is an own Philox stream
// (seed, element quad, call counter) -> Box-Muller, as for the reset sampling.
//   correlated' = corr * a_corr + b_corr                       (gaussian: var_corr, mu_corr | uniform: hi_corr - lo_corr, lo_corr)
//   noise       = (correlated' + white * a) + b                (gaussian: white ~ N(0,1), var, mu | uniform: white ~ U[0,1), hi - lo, lo)
//   dst         = src + noise  (additive)   |   src * noise  (scaling)
#pragma once
#include "sdx_math.cuh"

#define SDX_DR_STREAM 0x44520000u   /* third Philox counter word of this stream ("DR") */

__device__ __forceinline__ void dr_white4(uint64_t seed, uint32_t quad, uint32_t counter, int uniform, float w[4]) {
  uint32_t r[4];
  philox(seed, quad, counter, SDX_DR_STREAM, r);
  if (uniform) {
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = (float)(r[j] >> 8) * (1.0f / 16777216.0f);
  } else {
    const float u0 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u1 = (float)(r[1] >> 8) * (1.0f / 16777216.0f);
    const float u2 = ((float)(r[2] >> 8) + 0.5f) * (1.0f / 16777216.0f), u3 = (float)(r[3] >> 8) * (1.0f / 16777216.0f);
    const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    w[0] = ra * cosf(6.283185307179586f * u1); w[1] = ra * sinf(6.283185307179586f * u1);
    w[2] = rb * cosf(6.283185307179586f * u3); w[3] = rb * sinf(6.283185307179586f * u3);
  }
}

// dst[i] ~ N(0,1): the correlated-noise tensor, drawn once per refresh of the randomisation parameters (BT:293-296)
__global__ void k_dr_randn(float* __restrict__ dst, int64_t n, uint64_t seed, uint32_t counter) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (4 * q >= n) return;
  float w[4];
  dr_white4(seed, (uint32_t)q, counter, 0, w);
#pragma unroll
  for (int j = 0; j < 4; ++j) if (4 * q + j < n) dst[4 * q + j] = w[j];
}

__global__ void k_dr_noise(float* __restrict__ dst, const float* __restrict__ src, const float* __restrict__ corr, int64_t n,
                           float a_corr, float b_corr, float a, float b, int uniform, int scaling, uint64_t seed, uint32_t counter) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (4 * q >= n) return;
  float w[4];
  dr_white4(seed, (uint32_t)q, counter, uniform, w);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = 4 * q + j;
    if (i < n) {
      const float c = corr[i] * a_corr + b_corr;
      const float noise = (c + w[j] * a) + b;
      const float x = src[i];
      dst[i] = scaling ? x * noise : x + noise;
    }
  }
}
