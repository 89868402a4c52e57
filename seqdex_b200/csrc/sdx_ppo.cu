// sdx_ppo.cu -- C-ABI of the PPO tensor path: tcgen05 GEMM launcher (sdx_gemm.cuh) + the fused PPO kernels.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <string.h>
#include <math.h>

#include "sdx_gemm.cuh"
#include "../../include/seqdex_b200.h"

extern "C" const char* sdx_last_error(void);
void sdx_set_error(const char* msg);   // defined in sdx_env.cu

#define PCK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { char b[256]; snprintf(b, sizeof b, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(_e)); sdx_set_error(b); return -1; } } while (0)

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn g_encode = nullptr;
static long long g_ppo_launches = 0;   // kernels launched by this translation unit
extern "C" long long sdx_ppo_launch_count(void) { return g_ppo_launches; }

static int get_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  PCK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) { sdx_set_error("cuTensorMapEncodeTiled not available from the driver"); return -1; }
  g_encode = (encode_tiled_fn)fn;
  return 0;
}

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows x box_cols], 128B or 64B swizzle
static int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols = GEMM_BK,
                    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
  if (get_encode()) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { char b[160]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld); sdx_set_error(b); return -1; }
  return 0;
}

#define GEMM_BN 128
#define GEMM_STAGES 4   /* 4 x 32 KB TMA stages + 64 KB epilogue staging: one persistent CTA per SM */
typedef GemmSmem<GEMM_BN, GEMM_STAGES> GSm;
typedef GemmSmem<256, 3> GSmW;
typedef GemmSmem<256, 4, 128> GSmP;   // CTA pair: 256 x 256 tile per pair, each CTA stages A 128 x 64 + half of B 128 x 64 per stage   // wide tiles: 128 x 256, 3 x 48 KB stages + 64 KB staging (higher flop/byte against the L2 bound)
static bool g_attr_set = false;
static int set_attrs() {
  if (g_attr_set) return 0;
  int bytes = (int)sizeof(GSm) + 1024;
  PCK(cudaFuncSetAttribute(k_gemm_tn<GEMM_BN, GEMM_STAGES, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PCK(cudaFuncSetAttribute(k_gemm_tn<GEMM_BN, GEMM_STAGES, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PCK(cudaFuncSetAttribute(k_gemm_tn<GEMM_BN, GEMM_STAGES, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PCK(cudaFuncSetAttribute(k_gemm_tn<GEMM_BN, GEMM_STAGES, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PCK(cudaFuncSetAttribute(k_gemm_tn<GEMM_BN, GEMM_STAGES, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PCK(cudaFuncSetAttribute(k_gemm_tn<256, 3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmW) + 1024));
  PCK(cudaFuncSetAttribute(k_gemm_tn<256, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmW) + 1024));
  PCK(cudaFuncSetAttribute(k_gemm_tn2<256, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmP) + 1024));
  PCK(cudaFuncSetAttribute(k_gemm_tn2<256, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmP) + 1024));
  PCK(cudaFuncSetAttribute(k_gemm_tn2<256, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmP) + 1024));
  g_attr_set = true;
  return 0;
}

// D[M,N] = A[M,K] . B[N,K]^T, bf16 operands (row-major, K contiguous, lda/ldb in elements, multiples of 8).
// mode: see sdx_gemm.cuh.  splits: split-K factor (mode 2 only; output must be zeroed by the caller).
extern "C" int sdx_gemm_bf16_tn(int mode, const void* A, int M, int K, int lda, const void* B, int N, int ldb, const float* bias,
                                const void* h, int ldh, void* out, int ldo, void* out_t, int ldt, float* outf, int ldf,
                                int splits, void* stream) {
  if (set_attrs()) return -1;
  if ((lda % 8) || (ldb % 8) || M <= 0 || N <= 0 || K <= 0) { sdx_set_error("sdx_gemm_bf16_tn: bad shape (ld must be a multiple of 8)"); return -1; }
  CUtensorMap ma, mb, mo, mt, mh;
  static int wide_ok = -1;
  if (wide_ok < 0) { const char* e = getenv("SDX_GEMM_WIDE"); wide_ok = e ? atoi(e) : 1; }
  // 128 x 256 tiles for the large bf16-output GEMMs: twice the flops per byte pulled from L2 into the SM
  static int pair_ok = -1;
  if (pair_ok < 0) { const char* e = getenv("SDX_GEMM_PAIR"); pair_ok = e ? atoi(e) : 1; }
  // CTA pairs (cta_group::2, 256 x 256 tile per pair) for the large bf16-output GEMMs: half the operand bytes per SM
  // ... and for the split-K dW GEMMs (mode 2): 128 x 128 tiles pull (128 + 128) x 64 x 2 B out of L2 per 128 x 128 x 64 MACs and are
  // L2-bandwidth-bound (537 MB through L2 for dW L1 at M = 32768: 65-80 us); a 256 x 256 tile per CTA pair halves the bytes per MAC
  static int pair_dw = -1;
  if (pair_dw < 0) { const char* e = getenv("SDX_GEMM_PAIR_DW"); pair_dw = e ? atoi(e) : 1; }
  const bool pair = pair_ok && (((mode == 0 && N >= 256) || (mode == 1 && N >= 1024)) && (long long)M * N >= (long long)256 * 256 * 74 ||
                                (mode == 2 && pair_dw && M >= 256 && N >= 256 && K >= 8192));
  const bool wide = !pair && wide_ok && mode == 0 && N >= 256 && (long long)M * N >= (long long)GEMM_BM * 256 * 148;
  const int BNsel = (wide || pair) ? 256 : GEMM_BN;
  if (make_map(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM)) return -1;
  if (make_map(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, pair ? 128 : BNsel)) return -1;
  mo = ma; mt = ma; mh = ma;   // placeholders when the mode has no bf16 output / no h input
  if (mode == 0 || mode == 1) {
    if (!out || (ldo % 8) || (out_t && (ldt % 8))) { sdx_set_error("sdx_gemm_bf16_tn: bf16 outputs need ld % 8 == 0"); return -1; }
    if (make_map(&mo, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;                   // [32 rows x 64 cols] boxes
    if (mode == 1) {
      if (!h || (ldh % 8)) { sdx_set_error("sdx_gemm_bf16_tn: mode 1 needs h with ldh % 8 == 0"); return -1; }
      if (make_map(&mh, h, (uint64_t)M, (uint64_t)N, (uint64_t)ldh, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    }
    if (out_t && make_map(&mt, out_t, (uint64_t)N, (uint64_t)M, (uint64_t)ldt, 64, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return -1;          // [64 n x 32 m] boxes
  }
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
  if (mode != 2 || splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  g.kblocks_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + g.kblocks_per_split - 1) / g.kblocks_per_split;
  g.bias = bias; g.h = (const __nv_bfloat16*)h; g.ldh = ldh; g.out = (__nv_bfloat16*)out; g.ldo = ldo;
  g.out_t = (__nv_bfloat16*)out_t; g.ldt = ldt; g.outf = outf; g.ldf = ldf;
  const int tile_m = pair ? 2 * GEMM_BM : GEMM_BM;
  int n_tiles = ((N + BNsel - 1) / BNsel) * ((M + tile_m - 1) / tile_m) * splits;
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  dim3 grid(n_tiles < n_sm ? n_tiles : n_sm);   // persistent: one CTA per SM walks the tiles
  size_t smem = sizeof(GSm) + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  if (pair) {
    const size_t smp = sizeof(GSmP) + 1024;
    const int pairs = n_tiles < n_sm / 2 ? n_tiles : n_sm / 2;
    if (mode == 0) k_gemm_tn2<256, 4, 0><<<2 * pairs, GEMM_THREADS, smp, st>>>(ma, mb, mo, mt, mh, g);
    else if (mode == 1) k_gemm_tn2<256, 4, 1><<<2 * pairs, GEMM_THREADS, smp, st>>>(ma, mb, mo, mt, mh, g);
    else k_gemm_tn2<256, 4, 2><<<2 * pairs, GEMM_THREADS, smp, st>>>(ma, mb, mo, mt, mh, g);
  } else if (wide) {
    const size_t smw = sizeof(GSmW) + 1024;
    if (mode == 0) k_gemm_tn<256, 3, 0><<<grid, GEMM_THREADS, smw, st>>>(ma, mb, mo, mt, mh, g);
    else k_gemm_tn<256, 3, 1><<<grid, GEMM_THREADS, smw, st>>>(ma, mb, mo, mt, mh, g);
  } else switch (mode) {
    case 0: k_gemm_tn<GEMM_BN, GEMM_STAGES, 0><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, mo, mt, mh, g); break;
    case 1: k_gemm_tn<GEMM_BN, GEMM_STAGES, 1><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, mo, mt, mh, g); break;
    case 2: k_gemm_tn<GEMM_BN, GEMM_STAGES, 2><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, mo, mt, mh, g); break;
    case 3: k_gemm_tn<GEMM_BN, GEMM_STAGES, 3><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, mo, mt, mh, g); break;
    case 4: k_gemm_tn<GEMM_BN, GEMM_STAGES, 4><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, mo, mt, mh, g); break;
    default: sdx_set_error("sdx_gemm_bf16_tn: bad mode"); return -1;
  }
  g_ppo_launches++;
  PCK(cudaGetLastError());
  return 0;
}

// =====================================================================================================
// MLP object: fp32 master parameters + Adam state, bf16 compute copies (and their transposes), activation
// and gradient staging buffers laid out so that EVERY contraction of forward and backward is the K-major
// tcgen05 GEMM above:
//   forward   a_{l+1} = ELU(a_l W_l^T + b_l)            A = a_l [M,d_l]          B = W_l  [d_{l+1}, d_l]
//   head      out     = a_3 W_3^T + b_3   (fp32)        A = a_3 [M,256]          B = W_3  [out, 256]
//   dX        dz_l    = (dz_{l+1} W_l) * ELU'(a_l)      A = dz_{l+1} [M,d_{l+1}] B = W_l^T [d_l, d_{l+1}]
//   dW,db     [dW_l | db_l] = dz_{l+1}^T [a_l | 1]      A = dz_{l+1}^T [d_{l+1},M] B = [a_l | 1]^T [d_l+16, M]
// (the "ones" row appended to the transposed activations makes the bias gradient one more output column).
// Parameter vector order = torch state_dict of rl_games' network: W0,b0,W1,b1,W2,b2,W3(mu / value),b3[,sigma].
// =====================================================================================================
struct sdx_mlp {
  int in_dim, in_pad, out_dim, max_rows, has_sigma;
  int d[5];
  size_t nparams, w_off[4], b_off[4], sigma_off;
  float *params, *grads, *adam_m, *adam_v, *out, *scal, *gW[4];
  __nv_bfloat16 *W[4], *Wt[4], *A[4], *At[4], *dZ[5], *dZt[5];
  const __nv_bfloat16 *a0, *at0; int ldt0;   // layer-0 input of the last forward (own staging buffers or a caller-converted batch)
  long long* adam_t;   // DEVICE: optimiser step counter, followed by float bc[2] = 1 - beta^t (k_adam_tick) -- on the device so that a captured
                       // CUDA graph of the update replays with the right bias corrections
  cudaEvent_t layer_done[4];   // recorded when layer l's gradient slice is complete in `grads` (pipelined all-reduce, sdx_mlp_backward_pipelined)
  int pipelined;                // last backward ran pipelined: per-layer unpack already done
};
static inline int pad64(int x) { return (x + 63) / 64 * 64; }

__global__ void k_fill_bf16(__nv_bfloat16* p, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = __float2bfloat16_rn(v);
}
// fp32 [R, C] (ld) -> bf16 row-major [R, ldo] (cols [C, Cpad) zero-filled) and/or transposed bf16 [Cpad.., ldt];
// optional per-column normalisation (x - mean) / sqrt(var + 1e-5) clamped to +-5 (rl_games RunningMeanStd)
// perm_h > 0: output row r = n * perm_h + t reads source row t * (R / perm_h) + n -- rl_games' swap_and_flatten01 (time-major
// rollout buffers [H][N] -> env-major batch rows, RGC:1480-1481) fused into the conversion
__global__ void k_cvt_2way(const float* __restrict__ src, int R, int C, int ld, int Cpad, __nv_bfloat16* __restrict__ dst, int ldo,
                           __nv_bfloat16* __restrict__ dst_t, int ldt, const float* __restrict__ mean, const float* __restrict__ var,
                           int perm_h = 0) {
  __shared__ float tile[32][33];
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.0f;
    if (r < R && c < C) {
      const int rs = perm_h > 0 ? (r % perm_h) * (R / perm_h) + r / perm_h : r;
      v = src[(size_t)rs * ld + c];
      if (mean) { v = (v - mean[c]) / sqrtf(var[c] + 1e-5f); v = fminf(fmaxf(v, -5.0f), 5.0f); }
    }
    tile[i][threadIdx.x] = v;
    if (dst && r < R && c < Cpad) dst[(size_t)r * ldo + c] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  if (dst_t)
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int c = c0 + i, r = r0 + threadIdx.x;
      if (c < Cpad && r < R) dst_t[(size_t)c * ldt + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
}
__global__ void k_unpack_grads(const float* __restrict__ gW, int rows, int kreal, int kones, int ldg, float* __restrict__ gw, float* __restrict__ gb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (kreal + 1)) return;
  int r = i / (kreal + 1), c = i % (kreal + 1);
  if (c < kreal) gw[(size_t)r * kreal + c] = gW[(size_t)r * ldg + c];
  else gb[r] = gW[(size_t)r * ldg + kones];
}
// sum of squares in TWO deterministic stages (per-block partials, then one block in fixed order): data-parallel replicas
// must compute bit-identical clip factors from their bit-identical all-reduced gradients, or they drift apart
__global__ void k_sumsq(const float* __restrict__ g, size_t n, float* __restrict__ partial) {
  __shared__ float sh[32];
  float s = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += g[i] * g[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0f;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}
__global__ void k_sumsq_final(const float* __restrict__ partial, int nb, float* __restrict__ out) {
  __shared__ float sh[32];
  float s = 0.0f;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0f;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *out = s;
  }
}
// step += 1; bias corrections 1 - beta^step for the k_adam that follows (one thread)
__global__ void k_adam_tick(long long* __restrict__ t, float b1, float b2) {
  const long long s = *t + 1;
  *t = s;
  float* bc = (float*)(t + 1);
  bc[0] = 1.0f - powf(b1, (float)s); bc[1] = 1.0f - powf(b2, (float)s);
}
// torch.optim.Adam (eps 1e-8, no weight decay; RGC:1102) with rl_games' global grad-norm clip folded in (RGC:1866-1872)
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                       float lr, float b1, float b2, float eps, const long long* __restrict__ tick, float max_norm, const float* __restrict__ sumsq,
                       const float* __restrict__ lr_dev = nullptr) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float bc1 = ((const float*)(tick + 1))[0], bc2 = ((const float*)(tick + 1))[1];
  if (lr_dev) lr = *lr_dev;
  float scale = 1.0f;
  if (max_norm > 0.0f) { float nrm = sqrtf(*sumsq); scale = fminf(1.0f, max_norm / (nrm + 1e-6f)); }
  float gi = g[i] * scale;
  float mi = b1 * m[i] + (1.0f - b1) * gi;
  float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  p[i] -= (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
}
__global__ void k_sync_w(const float* __restrict__ w, int N, int Kreal, __nv_bfloat16* __restrict__ W, int ldw, int Kpad,
                         __nv_bfloat16* __restrict__ Wt, int ldwt, int Npad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int tot = max(N * Kpad, Wt ? Kreal * Npad : 0);
  if (i >= tot) return;
  if (i < N * Kpad) { int n = i / Kpad, k = i % Kpad; W[(size_t)n * ldw + k] = __float2bfloat16_rn(k < Kreal ? w[(size_t)n * Kreal + k] : 0.0f); }
  if (Wt && i < Kreal * Npad) { int k = i / Npad, n = i % Npad; Wt[(size_t)k * ldwt + n] = __float2bfloat16_rn(n < N ? w[(size_t)n * Kreal + k] : 0.0f); }
}

struct LayerTab {
  const float* w[4]; __nv_bfloat16* W[4]; __nv_bfloat16* Wt[4]; int N[4], Kreal[4], Kpad[4], Npad[4];
  const float* gW[4]; float* gw[4]; float* gb[4]; int kones[4], ldg[4];
};
// all four layers in one launch (blockIdx.y = layer): fp32 master -> bf16 W [N][Kpad] and W^T [Kreal][Npad]
__global__ void k_sync_all(LayerTab t) {
  int l = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  int N = t.N[l], Kreal = t.Kreal[l], Kpad = t.Kpad[l], Npad = t.Npad[l];
  const float* w = t.w[l];
  if (i < N * Kpad) { int n = i / Kpad, k = i % Kpad; t.W[l][(size_t)n * Kpad + k] = __float2bfloat16_rn(k < Kreal ? w[(size_t)n * Kreal + k] : 0.0f); }
  if (t.Wt[l] && i < Kreal * Npad) { int k = i / Npad, n = i % Npad; t.Wt[l][(size_t)k * Npad + n] = __float2bfloat16_rn(n < N ? w[(size_t)n * Kreal + k] : 0.0f); }
}
// all four layers in one launch: split-K scratch [N][K+16] -> flat gradient vector (weights, then the bias column)
__global__ void k_unpack_all(LayerTab t) {
  int l = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  int rows = t.N[l], kreal = t.Kreal[l];
  if (i >= rows * (kreal + 1)) return;
  int r = i / (kreal + 1), c = i % (kreal + 1);
  if (c < kreal) t.gw[l][(size_t)r * kreal + c] = t.gW[l][(size_t)r * t.ldg[l] + c];
  else t.gb[l][r] = t.gW[l][(size_t)r * t.ldg[l] + t.kones[l]];
}
static LayerTab make_tab(sdx_mlp* m);

// hidden widths h1,h2,h3 must be multiples of 64 (K dimensions of the tensor-core tiles)
extern "C" int sdx_mlp_create_ex(int in_dim, int out_dim, int h1, int h2, int h3, int max_rows, int has_sigma, sdx_mlp** out) {
  if ((h1 % 64) || (h2 % 64) || (h3 % 64) || out_dim < 1 || out_dim > 64) { sdx_set_error("sdx_mlp_create_ex: hidden widths must be multiples of 64, out_dim <= 64"); return -1; }
  sdx_mlp* m = new sdx_mlp();
  memset(m, 0, sizeof(*m));
  m->in_dim = in_dim; m->in_pad = pad64(in_dim); m->out_dim = out_dim; m->max_rows = max_rows; m->has_sigma = has_sigma;
  m->d[0] = m->in_pad; m->d[1] = h1; m->d[2] = h2; m->d[3] = h3; m->d[4] = out_dim;
  if (max_rows % 8) { sdx_set_error("sdx_mlp_create: max_rows must be a multiple of 8"); return -1; }
  size_t off = 0;
  for (int l = 0; l < 4; ++l) {
    int kreal = l == 0 ? in_dim : m->d[l];
    m->w_off[l] = off; off += (size_t)m->d[l + 1] * kreal;
    m->b_off[l] = off; off += m->d[l + 1];
  }
  m->sigma_off = off; if (has_sigma) off += out_dim;
  m->nparams = off;
  PCK(cudaMalloc(&m->params, off * 4)); PCK(cudaMemset(m->params, 0, off * 4));
  PCK(cudaMalloc(&m->grads, (off + SDX_GRAD_TAIL) * 4)); PCK(cudaMemset(m->grads, 0, (off + SDX_GRAD_TAIL) * 4));   // tail: loss statistics that ride along with the gradient all-reduce
  PCK(cudaMalloc(&m->adam_m, off * 4)); PCK(cudaMemset(m->adam_m, 0, off * 4));
  PCK(cudaMalloc(&m->adam_v, off * 4)); PCK(cudaMemset(m->adam_v, 0, off * 4));
  PCK(cudaMalloc(&m->out, (size_t)max_rows * out_dim * 4));
  PCK(cudaMalloc(&m->scal, 2048)); PCK(cudaMemset(m->scal, 0, 2048));   // [0] grad norm^2, [16..16+296) per-block partials
  PCK(cudaMalloc(&m->adam_t, 16)); PCK(cudaMemset(m->adam_t, 0, 16));
  size_t R = max_rows;
  for (int l = 0; l < 4; ++l) {
    int K = m->d[l], N = m->d[l + 1], Npad = pad64(N);
    PCK(cudaMalloc(&m->W[l], (size_t)N * K * 2)); PCK(cudaMemset(m->W[l], 0, (size_t)N * K * 2));
    if (l > 0) { PCK(cudaMalloc(&m->Wt[l], (size_t)K * Npad * 2)); PCK(cudaMemset(m->Wt[l], 0, (size_t)K * Npad * 2)); }
    PCK(cudaMalloc(&m->A[l], R * K * 2)); PCK(cudaMemset(m->A[l], 0, R * K * 2));
    PCK(cudaMalloc(&m->At[l], (size_t)(K + 16) * R * 2)); PCK(cudaMemset(m->At[l], 0, (size_t)(K + 16) * R * 2));
    k_fill_bf16<<<(unsigned)((R + 255) / 256), 256>>>(m->At[l] + (size_t)K * R, R, 1.0f);   // the "ones" row -> bias gradients
    PCK(cudaMalloc(&m->gW[l], (size_t)N * (K + 16) * 4));
    PCK(cudaMalloc(&m->dZ[l + 1], R * Npad * 2)); PCK(cudaMemset(m->dZ[l + 1], 0, R * Npad * 2));
    PCK(cudaMalloc(&m->dZt[l + 1], (size_t)Npad * R * 2)); PCK(cudaMemset(m->dZt[l + 1], 0, (size_t)Npad * R * 2));
  }
  for (int l = 0; l < 4; ++l) PCK(cudaEventCreateWithFlags(&m->layer_done[l], cudaEventDisableTiming));
  PCK(cudaDeviceSynchronize());
  *out = m;
  return 0;
}
extern "C" int sdx_mlp_create(int in_dim, int out_dim, int max_rows, int has_sigma, sdx_mlp** out) {
  return sdx_mlp_create_ex(in_dim, out_dim, 1024, 512, 256, max_rows, has_sigma, out);   // cfg/lego/ppo_continuous_grasp.yaml:21-23
}
extern "C" void sdx_mlp_destroy(sdx_mlp* m) {
  if (!m) return;
  cudaFree(m->params); cudaFree(m->grads); cudaFree(m->adam_m); cudaFree(m->adam_v); cudaFree(m->out); cudaFree(m->scal); cudaFree(m->adam_t);
  for (int l = 0; l < 4; ++l) cudaEventDestroy(m->layer_done[l]);
  for (int l = 0; l < 4; ++l) { cudaFree(m->W[l]); cudaFree(m->Wt[l]); cudaFree(m->A[l]); cudaFree(m->At[l]); cudaFree(m->gW[l]); cudaFree(m->dZ[l + 1]); cudaFree(m->dZt[l + 1]); }
  delete m;
}
extern "C" int sdx_mlp_info(sdx_mlp* m, int64_t* nparams, void** params, void** grads, void** out, void** adam_m, void** adam_v) {
  *nparams = (int64_t)m->nparams; *params = m->params; *grads = m->grads; *out = m->out;
  if (adam_m) *adam_m = m->adam_m;
  if (adam_v) *adam_v = m->adam_v;
  return 0;
}
static LayerTab make_tab(sdx_mlp* m) {
  LayerTab t;
  for (int l = 0; l < 4; ++l) {
    t.w[l] = m->params + m->w_off[l]; t.W[l] = m->W[l]; t.Wt[l] = l > 0 ? m->Wt[l] : nullptr;
    t.N[l] = m->d[l + 1]; t.Kreal[l] = l == 0 ? m->in_dim : m->d[l]; t.Kpad[l] = m->d[l]; t.Npad[l] = pad64(m->d[l + 1]);
    t.gW[l] = m->gW[l]; t.gw[l] = m->grads + m->w_off[l]; t.gb[l] = m->grads + m->b_off[l]; t.kones[l] = m->d[l]; t.ldg[l] = m->d[l] + 16;
  }
  return t;
}
// refresh the bf16 compute copies from the fp32 master parameters (after load / optimiser step)
extern "C" int sdx_mlp_sync(sdx_mlp* m, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  {
    int mx = 0;
    for (int l = 0; l < 4; ++l) { int a = m->d[l + 1] * m->d[l], b = (l == 0 ? m->in_dim : m->d[l]) * pad64(m->d[l + 1]); mx = a > mx ? a : mx; mx = b > mx ? b : mx; }
    dim3 grd((mx + 255) / 256, 4);
    k_sync_all<<<grd, 256, 0, st>>>(make_tab(m));
    g_ppo_launches++;
    PCK(cudaGetLastError());
    return 0;
  }
}
// x: fp32 [M, in_dim] (device).  mean/var (nullable): RunningMeanStd input normalisation.  train != 0 keeps the
// transposed activations the backward pass needs.  Result: m->out fp32 [M, out_dim].
static int mlp_forward_core(sdx_mlp* m, int M, int train, void* stream) {
  for (int l = 0; l < 3; ++l)
    if (sdx_gemm_bf16_tn(0, l == 0 ? (const void*)m->a0 : (const void*)m->A[l], M, m->d[l], m->d[l], m->W[l], m->d[l + 1], m->d[l], m->params + m->b_off[l], nullptr, 0,
                         m->A[l + 1], m->d[l + 1], train ? m->At[l + 1] : nullptr, m->max_rows, nullptr, 0, 1, stream)) return -1;
  return sdx_gemm_bf16_tn(4, m->A[3], M, m->d[3], m->d[3], m->W[3], m->out_dim, m->d[3], m->params + m->b_off[3], nullptr, 0, nullptr, 0, nullptr, 0,
                          m->out, m->out_dim, 1, stream);
}
extern "C" int sdx_mlp_forward(sdx_mlp* m, const float* x, int M, const float* mean, const float* var, int train, void* stream) {
  if (M > m->max_rows || M <= 0) { sdx_set_error("sdx_mlp_forward: M exceeds max_rows"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 blk(32, 8), grd((m->in_pad + 31) / 32, (M + 31) / 32);
  k_cvt_2way<<<grd, blk, 0, st>>>(x, M, m->in_dim, m->in_dim, m->in_pad, m->A[0], m->in_pad, train ? m->At[0] : nullptr, m->max_rows, mean, var);
  g_ppo_launches++;
  PCK(cudaGetLastError());
  m->a0 = m->A[0]; m->at0 = m->At[0]; m->ldt0 = m->max_rows;
  return mlp_forward_core(m, M, train, stream);
}
// Convert a whole rollout batch ONCE per iteration: x fp32 [B, in_dim] -> xb bf16 [B, in_pad] and xt bf16 [in_pad + 16, B]
// (row in_pad = ones, for the bias gradient).  Minibatches are then row ranges of xb / column ranges of xt.
static int convert_batch(sdx_mlp* m, const float* x, int B, int horizon, const float* mean, const float* var, void* xb, void* xt, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B % 8 || (horizon > 0 && B % horizon)) { sdx_set_error("sdx_mlp_convert_batch: B must be a multiple of 8 (and of the horizon)"); return -1; }
  dim3 blk(32, 8), grd((m->in_pad + 31) / 32, (B + 31) / 32);
  k_cvt_2way<<<grd, blk, 0, st>>>(x, B, m->in_dim, m->in_dim, m->in_pad, (__nv_bfloat16*)xb, m->in_pad, (__nv_bfloat16*)xt, B, mean, var, horizon);
  PCK(cudaMemsetAsync((__nv_bfloat16*)xt + (size_t)m->in_pad * B, 0, (size_t)16 * B * 2, st));
  k_fill_bf16<<<(unsigned)((B + 255) / 256), 256, 0, st>>>((__nv_bfloat16*)xt + (size_t)m->in_pad * B, (size_t)B, 1.0f);
  g_ppo_launches += 2;
  PCK(cudaGetLastError());
  return 0;
}
extern "C" int sdx_mlp_convert_batch(sdx_mlp* m, const float* x, int B, const float* mean, const float* var, void* xb, void* xt, void* stream) {
  return convert_batch(m, x, B, 0, mean, var, xb, xt, stream);
}
/* x is a time-major rollout buffer [horizon][B / horizon][in_dim]; the converted batch is env-major (swap_and_flatten01) */
extern "C" int sdx_mlp_convert_batch_env_major(sdx_mlp* m, const float* x, int B, int horizon, const float* mean, const float* var, void* xb, void* xt,
                                               void* stream) {
  return convert_batch(m, x, B, horizon, mean, var, xb, xt, stream);
}
// forward on rows [row0, row0 + M) of a converted batch (xb [B, in_pad], xt [in_pad + 16, B])
extern "C" int sdx_mlp_forward_pre(sdx_mlp* m, const void* xb, const void* xt, int B, int row0, int M, int train, void* stream) {
  if (M > m->max_rows || M <= 0 || row0 % 8) { sdx_set_error("sdx_mlp_forward_pre: bad row range"); return -1; }
  m->a0 = (const __nv_bfloat16*)xb + (size_t)row0 * m->in_pad;
  m->at0 = (const __nv_bfloat16*)xt + row0;
  m->ldt0 = B;
  return mlp_forward_core(m, M, train, stream);
}
// dout: fp32 [M, out_dim] = dLoss/d(out).  Fills m->grads (W and b of every layer; sigma is the loss kernel's job).
static int mlp_backward(sdx_mlp* m, const float* dout, int M, void* stream, int pipelined) {
  if (M > m->max_rows || M <= 0) { sdx_set_error("sdx_mlp_backward: M exceeds max_rows"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  int opad = pad64(m->out_dim);
  dim3 blk(32, 8), grd((opad + 31) / 32, (M + 31) / 32);
  k_cvt_2way<<<grd, blk, 0, st>>>(dout, M, m->out_dim, m->out_dim, opad, m->dZ[4], opad, m->dZt[4], m->max_rows, nullptr, nullptr);
  g_ppo_launches++;
  PCK(cudaGetLastError());
  for (int l = 3; l >= 0; --l) {
    int N = m->d[l + 1], K = m->d[l], ldg = K + 16;
    PCK(cudaMemsetAsync(m->gW[l], 0, (size_t)N * ldg * 4, st));
    int tiles = ((N + 127) / 128) * ((K + 16 + 127) / 128);
    int splits = 296 / tiles; if (splits < 1) splits = 1;   // two tiles per persistent CTA: the second one's main loop hides the first one's reduction epilogue
    static int dw_items = -1;                                // CTA-pair dW (256 x 256 tiles, sdx_gemm_bf16_tn): work items per pair (1: 47 / 52 / 29 us, 2: 49 / 53 / 33 us)
    if (dw_items < 0) { const char* e = getenv("SDX_GEMM_DW_ITEMS"); dw_items = e ? atoi(e) : 1; }
    { const char* e = getenv("SDX_GEMM_PAIR_DW");
      if ((!e || atoi(e)) && N >= 256 && K + 16 >= 256 && M >= 8192) {
        const int t2 = ((N + 255) / 256) * ((K + 16 + 255) / 256);
        splits = (74 * dw_items) / t2; if (splits < 1) splits = 1;
      } }
    if (sdx_gemm_bf16_tn(2, m->dZt[l + 1], N, M, m->max_rows, l == 0 ? (const void*)m->at0 : (const void*)m->At[l], K + 16, l == 0 ? m->ldt0 : m->max_rows, nullptr, nullptr, 0,
                         nullptr, 0, nullptr, 0, m->gW[l], ldg, splits, stream)) return -1;
    if (pipelined) {   // this layer's slice of the flat gradient vector is final NOW: publish it before the layers below are differentiated
      const int kreal = l == 0 ? m->in_dim : m->d[l];
      const int tot = N * (kreal + 1);
      k_unpack_grads<<<(tot + 255) / 256, 256, 0, st>>>(m->gW[l], N, kreal, m->d[l], ldg, m->grads + m->w_off[l], m->grads + m->b_off[l]);
      g_ppo_launches++;
      PCK(cudaEventRecord(m->layer_done[l], st));
    }
    if (l > 0) {
      int Kd = pad64(N);
      if (sdx_gemm_bf16_tn(1, m->dZ[l + 1], M, Kd, Kd, m->Wt[l], K, Kd, nullptr, m->A[l], K, m->dZ[l], K, m->dZt[l], m->max_rows, nullptr, 0, 1, stream)) return -1;
    }
  }
  if (!pipelined) {
    int mx = 0;
    for (int l = 0; l < 4; ++l) { int a = m->d[l + 1] * ((l == 0 ? m->in_dim : m->d[l]) + 1); mx = a > mx ? a : mx; }
    dim3 grd2((mx + 255) / 256, 4);
    k_unpack_all<<<grd2, 256, 0, st>>>(make_tab(m));
    g_ppo_launches++;
  }
  PCK(cudaGetLastError());
  return 0;
}
extern "C" int sdx_mlp_backward(sdx_mlp* m, const float* dout, int M, void* stream) { return mlp_backward(m, dout, M, stream, 0); }
/* same gradients, published layer by layer: layer l's slice of the flat gradient vector [w_off[l], b_off[l] + rows) is written right
 * after its dW GEMM and an event is recorded, so a data-parallel caller can all-reduce layer l while layers l-1 .. 0 are still being
 * differentiated (sdx_mlp_wait_layer makes another stream wait for that event) */
extern "C" int sdx_mlp_backward_pipelined(sdx_mlp* m, const float* dout, int M, void* stream) { return mlp_backward(m, dout, M, stream, 1); }
extern "C" int sdx_mlp_wait_layer(sdx_mlp* m, int layer, void* waiting_stream) {
  if (layer < 0 || layer > 3) { sdx_set_error("sdx_mlp_wait_layer: layer out of range"); return -1; }
  PCK(cudaStreamWaitEvent((cudaStream_t)waiting_stream, m->layer_done[layer], 0));
  return 0;
}
/* element range [begin, end) of layer l's gradients (weights then bias) in the flat vector */
extern "C" int sdx_mlp_layer_range(sdx_mlp* m, int layer, int64_t* begin, int64_t* end) {
  if (layer < 0 || layer > 3) { sdx_set_error("sdx_mlp_layer_range: layer out of range"); return -1; }
  *begin = (int64_t)m->w_off[layer]; *end = (int64_t)(m->b_off[layer] + m->d[layer + 1]);
  return 0;
}
/* optimiser step counter (the `step` of torch.optim.Adam's state): read with set < 0, overwritten otherwise (checkpoint restore) */
extern "C" long long sdx_mlp_adam_step(sdx_mlp* m, long long set) {
  long long t = 0;
  cudaDeviceSynchronize();
  if (set >= 0) { cudaMemcpy(m->adam_t, &set, 8, cudaMemcpyHostToDevice); return set; }
  cudaMemcpy(&t, m->adam_t, 8, cudaMemcpyDeviceToHost);
  return t;
}
/* kernels a replayed CUDA graph launched on behalf of this library (the graph was captured from these entry points; the counter only
 * sees the capture) */
extern "C" void sdx_ppo_add_launches(long long n) { g_ppo_launches += n; }
static int mlp_adam(sdx_mlp* m, float lr, const float* lr_dev, float b1, float b2, float eps, float max_norm, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  k_adam_tick<<<1, 1, 0, st>>>(m->adam_t, b1, b2);
  g_ppo_launches++;
  if (max_norm > 0.0f) {
    k_sumsq<<<296, 256, 0, st>>>(m->grads, m->nparams, m->scal + 16);
    k_sumsq_final<<<1, 256, 0, st>>>(m->scal + 16, 296, m->scal);
    g_ppo_launches++;
  }
  g_ppo_launches++;
  k_adam<<<(unsigned)((m->nparams + 255) / 256), 256, 0, st>>>(m->params, m->grads, m->adam_m, m->adam_v, m->nparams, lr, b1, b2, eps, m->adam_t, max_norm, m->scal, lr_dev);
  g_ppo_launches++;
  PCK(cudaGetLastError());
  return sdx_mlp_sync(m, stream);
}
extern "C" int sdx_mlp_adam(sdx_mlp* m, float lr, float b1, float b2, float eps, float max_norm, void* stream) {
  return mlp_adam(m, lr, nullptr, b1, b2, eps, max_norm, stream);
}
/* the learning rate is read from device memory when the step runs: the adaptive-KL schedule is applied on the device after every
 * minibatch (sdx_ppo_adaptive_lr) without a host round trip */
extern "C" int sdx_mlp_adam_dev(sdx_mlp* m, const float* lr_dev, float b1, float b2, float eps, float max_norm, void* stream) {
  return mlp_adam(m, 0.0f, lr_dev, b1, b2, eps, max_norm, stream);
}
// rl_games AdaptiveScheduler.update on the device (schedule_type 'legacy': after every minibatch, RGC:1360-1365): kl = stats[2] * inv_count;
// kl > 2 thr -> lr = max(lr / 1.5, lr_min); kl < 0.5 thr -> lr = min(lr * 1.5, lr_max).  The minibatch statistics are then added to
// accum[0..4) (+ accum[4] = minibatches seen, accum[5] = last kl) and cleared for the next minibatch.
__global__ void k_adaptive_lr(float* __restrict__ stats, float inv_count, float thr, float lr_min, float lr_max, float* __restrict__ lr,
                              float* __restrict__ accum, int adaptive) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float kl = stats[2] * inv_count;
  if (adaptive) {
    float l = *lr;
    if (kl > 2.0f * thr) l = fmaxf(l / 1.5f, lr_min);
    if (kl < 0.5f * thr) l = fminf(l * 1.5f, lr_max);
    *lr = l;
  }
  for (int i = 0; i < 4; ++i) { accum[i] += stats[i]; stats[i] = 0.0f; }
  accum[4] += 1.0f; accum[5] = kl;
}
extern "C" int sdx_ppo_adaptive_lr(float* stats_dev, float inv_count, float kl_threshold, float lr_min, float lr_max, float* lr_dev,
                                   float* accum_dev, int adaptive, void* stream) {
  k_adaptive_lr<<<1, 32, 0, (cudaStream_t)stream>>>(stats_dev, inv_count, kl_threshold, lr_min, lr_max, lr_dev, accum_dev, adaptive);
  g_ppo_launches++;
  PCK(cudaGetLastError());
  return 0;
}

// =====================================================================================================
// PPO elementwise kernels (rl_games 1.5.2 semantics)
// =====================================================================================================
__device__ __forceinline__ void philox4(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c[4] = {c0, c1, c2, 0u};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// a = mu + exp(logstd) * N(0,1);  neglogp = 0.5 sum((a-mu)/sigma)^2 + 0.5 A ln(2 pi) + sum(logstd)   (RGC:2115-2127)
__global__ void k_ppo_sample(const float* __restrict__ mu, const float* __restrict__ logstd, int M, int A, uint64_t seed, uint32_t counter,
                             float* __restrict__ actions, float* __restrict__ neglogp) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M) return;
  float nlp = 0.0f, sls = 0.0f;
  for (int i = 0; i < A; i += 4) {
    uint32_t r[4];
    philox4(seed, (uint32_t)e, counter, (uint32_t)(i >> 2), r);
    float u0 = ((r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u1 = (r[1] >> 8) * (1.0f / 16777216.0f);
    float u2 = ((r[2] >> 8) + 0.5f) * (1.0f / 16777216.0f), u3 = (r[3] >> 8) * (1.0f / 16777216.0f);
    float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    float z[4] = {ra * cosf(6.283185307f * u1), ra * sinf(6.283185307f * u1), rb * cosf(6.283185307f * u3), rb * sinf(6.283185307f * u3)};
    for (int j = 0; j < 4 && i + j < A; ++j) {
      float ls = logstd[i + j];
      actions[(size_t)e * A + i + j] = mu[(size_t)e * A + i + j] + expf(ls) * z[j];
      nlp += 0.5f * z[j] * z[j];
      sls += ls;
    }
  }
  neglogp[e] = nlp + 0.5f * (float)A * 1.8378770664093453f + sls;
}
// stats: [0] sum a_loss, [1] sum b_loss, [2] sum kl, [3] sum clipped indicator
__global__ void k_ppo_actor_loss(const float* __restrict__ mu, const float* __restrict__ logstd, const float* __restrict__ actions,
                                 const float* __restrict__ old_mu, const float* __restrict__ old_logstd, const float* __restrict__ old_neglogp,
                                 const float* __restrict__ adv, int M, int A, float e_clip, float bounds_coef, float inv_batch,
                                 float* __restrict__ dmu, float* __restrict__ dlogstd, float* __restrict__ stats) {
  __shared__ float s_dl[32], s_st[4];
  if (threadIdx.x < 32) s_dl[threadIdx.x] = 0.0f;
  if (threadIdx.x < 4) s_st[threadIdx.x] = 0.0f;
  __syncthreads();
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  float gk = 0.0f, st0 = 0.0f, st1 = 0.0f, st2 = 0.0f, st3 = 0.0f;
  if (e < M) {
    float nlp = 0.0f, sls = 0.0f, bl = 0.0f, kl = 0.0f;
    for (int i = 0; i < A; ++i) {
      float ls = logstd[i], sg = expf(ls), m_ = mu[(size_t)e * A + i];
      float z = (actions[(size_t)e * A + i] - m_) / sg;
      nlp += 0.5f * z * z; sls += ls;
      float hi = fmaxf(m_ - 1.1f, 0.0f), lo = fminf(m_ + 1.1f, 0.0f);
      bl += hi * hi + lo * lo;
      float so = expf(old_logstd[i]), dm = old_mu[(size_t)e * A + i] - m_;
      kl += logf(so / sg + 1e-5f) + (sg * sg + dm * dm) / (2.0f * (so * so + 1e-5f)) - 0.5f;   // torch_ext.policy_kl(mu, sigma, old_mu, old_sigma), RGC:1903
    }
    nlp += 0.5f * (float)A * 1.8378770664093453f + sls;
    float ratio = expf(old_neglogp[e] - nlp), a = adv[e];
    float s1 = -a * ratio, s2 = -a * fminf(fmaxf(ratio, 1.0f - e_clip), 1.0f + e_clip);
    bool inside = ratio > 1.0f - e_clip && ratio < 1.0f + e_clip;
    float al = fmaxf(s1, s2);
    // d a_loss / d neglogp: branch s1 -> a * ratio; branch s2 -> a * ratio inside the clip range, else 0
    float g = (s1 >= s2 || inside) ? a * ratio : 0.0f;
    gk = g; st0 = al; st1 = bl; st2 = kl; st3 = inside ? 0.0f : 1.0f;
  }
  // per-action pass: every lane takes part in the reductions (inactive rows contribute zeros)
  for (int i = 0; i < A; ++i) {
    float contrib = 0.0f;
    if (e < M) {
      float ls = logstd[i], sg = expf(ls), m_ = mu[(size_t)e * A + i];
      float z = (actions[(size_t)e * A + i] - m_) / sg;
      float db = 2.0f * fmaxf(m_ - 1.1f, 0.0f) + 2.0f * fminf(m_ + 1.1f, 0.0f);
      dmu[(size_t)e * A + i] = (gk * (-z / sg) + bounds_coef * db) * inv_batch;
      contrib = gk * (1.0f - z * z) * inv_batch;
    }
    for (int o = 16; o; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_dl[i], contrib);
  }
  for (int o = 16; o; o >>= 1) {
    st0 += __shfl_xor_sync(0xffffffffu, st0, o); st1 += __shfl_xor_sync(0xffffffffu, st1, o);
    st2 += __shfl_xor_sync(0xffffffffu, st2, o); st3 += __shfl_xor_sync(0xffffffffu, st3, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_st[0], st0); atomicAdd(&s_st[1], st1); atomicAdd(&s_st[2], st2); atomicAdd(&s_st[3], st3); }
  __syncthreads();
  if (threadIdx.x < A) atomicAdd(&dlogstd[threadIdx.x], s_dl[threadIdx.x]);
  if (threadIdx.x < 4) atomicAdd(&stats[threadIdx.x], s_st[threadIdx.x]);
}
// clipped value loss: c = max((v-R)^2, (v_old + clip(v - v_old, +-e) - R)^2); dv = scale * dc/dv.  stats[0] += sum c
__global__ void k_ppo_value_loss(const float* __restrict__ v, const float* __restrict__ v_old, const float* __restrict__ ret, int M,
                                 float e_clip, int clip_value, float scale, float* __restrict__ dv, float* __restrict__ stats) {
  __shared__ float s;
  if (threadIdx.x == 0) s = 0.0f;
  __syncthreads();
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < M) {
    float vn = v[e], vo = v_old[e], r = ret[e];
    float l1 = (vn - r) * (vn - r), g = 2.0f * (vn - r), c = l1;
    if (clip_value) {
      float d = vn - vo, dc = fminf(fmaxf(d, -e_clip), e_clip);
      float vc = vo + dc, l2 = (vc - r) * (vc - r);
      if (l2 > l1) { c = l2; g = (d > -e_clip && d < e_clip) ? 2.0f * (vc - r) : 0.0f; }
    }
    dv[e] = g * scale;
    atomicAdd(&s, c);
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&stats[0], s);
}
__global__ void k_moments(const float* __restrict__ x, size_t n, double* __restrict__ out) {
  __shared__ double sh[2][32];
  double s = 0.0, q = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { double v = x[i]; s += v; q += v * v; }
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[0][threadIdx.x] : 0.0; q = threadIdx.x < (blockDim.x >> 5) ? sh[1][threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (threadIdx.x == 0) { atomicAdd(&out[0], s); atomicAdd(&out[1], q); }
  }
}
// (x - mean) / (std + 1e-8) with mean/std from sums over `count` samples (unbiased std like torch.std; RGC:1651)
__global__ void k_normalize(float* __restrict__ x, size_t n, const double* __restrict__ mom, double count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double mean = mom[0] / count, var = (mom[1] - count * mean * mean) / (count - 1.0);
  float sd = (float)sqrt(var > 0.0 ? var : 0.0);
  x[i] = (x[i] - (float)mean) / (sd + 1e-8f);
}
// per-column sum / sum of squares of x [B, D] into colmom [2][D] (double)
__global__ void k_col_moments(const float* __restrict__ x, int B, int D, double* __restrict__ colmom) {
  int c = blockIdx.x * 32 + (threadIdx.x & 31);
  int lane_r = threadIdx.x >> 5, nr = blockDim.x >> 5;
  double s = 0.0, q = 0.0;
  if (c < D)
    for (int r = blockIdx.y * nr + lane_r; r < B; r += gridDim.y * nr) { double v = x[(size_t)r * D + c]; s += v; q += v * v; }
  if (c < D) { atomicAdd(&colmom[c], s); atomicAdd(&colmom[D + c], q); }
}
// rl_games RunningMeanStd merge (parallel-variance): running (mean, var, count) with the batch moments
__global__ void k_rms_merge(float* __restrict__ mean, float* __restrict__ var, double* __restrict__ count, const double* __restrict__ colmom,
                            int D, double bcount) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  double bm = colmom[c] / bcount, bv = (colmom[D + c] - bcount * bm * bm) / (bcount - 1.0);
  double cnt = *count, tot = cnt + bcount, delta = bm - (double)mean[c];
  double m2 = (double)var[c] * cnt + bv * bcount + delta * delta * cnt * bcount / tot;
  mean[c] = (float)((double)mean[c] + delta * bcount / tot);
  var[c] = (float)(m2 / tot);
}
__global__ void k_add_count(double* count, double b) { *count += b; }

// t-value trainer loss (TVT:199-201,226): y = ELU(z) (the network ends in an ELU, TVF:44), BCEWithLogits(y, onehot(label)) with
// mean reduction over M x 2 elements.  dz = dL/dz (through the ELU), stats[0] += sum of the element losses.
__global__ void k_tvalue_bce(const float* __restrict__ z, const int* __restrict__ label, int M, float* __restrict__ dz, float* __restrict__ stats) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.0f;
  if (e < M) {
    for (int c = 0; c < 2; ++c) {
      float zz = z[2 * e + c], y = zz > 0.0f ? zz : (expf(zz) - 1.0f), t = (label[e] == c) ? 1.0f : 0.0f;
      float sg = 1.0f / (1.0f + expf(-y));
      l += fmaxf(y, 0.0f) - y * t + log1pf(expf(-fabsf(y)));
      dz[2 * e + c] = (sg - t) * (zz > 0.0f ? 1.0f : y + 1.0f) / (2.0f * (float)M);
    }
  }
  for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(stats, l);
}
extern "C" int sdx_tvalue_bce(const float* z, const int* label, int M, float* dz, float* stats, void* stream) {
  k_tvalue_bce<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(z, label, M, dz, stats);
  g_ppo_launches++;
  PCK(cudaGetLastError()); return 0;
}
extern "C" int sdx_ppo_sample(const float* mu, const float* logstd, int M, int A, uint64_t seed, uint32_t counter, float* actions, float* neglogp, void* stream) {
  k_ppo_sample<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mu, logstd, M, A, seed, counter, actions, neglogp);
  g_ppo_launches++;
  PCK(cudaGetLastError()); return 0;
}
extern "C" int sdx_ppo_actor_loss(const float* mu, const float* logstd, const float* actions, const float* old_mu, const float* old_logstd,
                                  const float* old_neglogp, const float* adv, int M, int A, float e_clip, float bounds_coef, float inv_batch,
                                  float* dmu, float* dlogstd, float* stats, void* stream) {
  if (A > 32) { sdx_set_error("sdx_ppo_actor_loss: at most 32 actions"); return -1; }
  k_ppo_actor_loss<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mu, logstd, actions, old_mu, old_logstd, old_neglogp, adv, M, A, e_clip, bounds_coef,
                                                                         inv_batch, dmu, dlogstd, stats);
  PCK(cudaGetLastError()); return 0;
}
extern "C" int sdx_ppo_value_loss(const float* v, const float* v_old, const float* ret, int M, float e_clip, int clip_value, float scale, float* dv,
                                  float* stats, void* stream) {
  k_ppo_value_loss<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(v, v_old, ret, M, e_clip, clip_value, scale, dv, stats);
  g_ppo_launches++;
  PCK(cudaGetLastError()); return 0;
}
// mom: double[2] device scratch (zeroed here).  After the call x is normalised in place (count = n unless n_total given for DP).
extern "C" int sdx_moments(const float* x, int64_t n, double* mom, void* stream) {
  PCK(cudaMemsetAsync(mom, 0, 16, (cudaStream_t)stream));
  k_moments<<<296, 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, mom);
  g_ppo_launches++;
  PCK(cudaGetLastError()); return 0;
}
extern "C" int sdx_normalize(float* x, int64_t n, const double* mom, double count, void* stream) {
  k_normalize<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, mom, count);
  g_ppo_launches++;
  PCK(cudaGetLastError()); return 0;
}
extern "C" int sdx_col_moments(const float* x, int B, int D, double* colmom, void* stream) {
  PCK(cudaMemsetAsync(colmom, 0, (size_t)2 * D * 8, (cudaStream_t)stream));
  dim3 grd((D + 31) / 32, 64);
  k_col_moments<<<grd, 256, 0, (cudaStream_t)stream>>>(x, B, D, colmom);
  g_ppo_launches++;
  PCK(cudaGetLastError()); return 0;
}
extern "C" int sdx_rms_merge(float* mean, float* var, double* count, const double* colmom, int D, double bcount, void* stream) {
  k_rms_merge<<<(D + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean, var, count, colmom, D, bcount);
  g_ppo_launches++;
  k_add_count<<<1, 1, 0, (cudaStream_t)stream>>>(count, bcount);
  g_ppo_launches++;
  PCK(cudaGetLastError()); return 0;
}
