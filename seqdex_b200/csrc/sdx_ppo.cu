// sdx_ppo.cu -- C-ABI of the PPO tensor path: tcgen05 GEMM launcher (sdx_gemm.cuh) + the fused PPO kernels.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string>

#include "sdx_gemm.cuh"

extern "C" const char* sdx_last_error(void);
void sdx_set_error(const char* msg);   // defined in sdx_env.cu

#define PCK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { char b[256]; snprintf(b, sizeof b, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(_e)); sdx_set_error(b); return -1; } } while (0)

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn g_encode = nullptr;

static int get_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  PCK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) { sdx_set_error("cuTensorMapEncodeTiled not available from the driver"); return -1; }
  g_encode = (encode_tiled_fn)fn;
  return 0;
}

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows x 64 cols], 128B swizzle
static int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  if (get_encode()) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {GEMM_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { char b[160]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld); sdx_set_error(b); return -1; }
  return 0;
}

#define GEMM_BN 128
#define GEMM_STAGES 5
typedef GemmSmem<GEMM_BN, GEMM_STAGES> GSm;
static bool g_attr_set = false;
static int set_attrs() {
  if (g_attr_set) return 0;
  int bytes = (int)sizeof(GSm) + 1024;
  PCK(cudaFuncSetAttribute(k_gemm_tn<GEMM_BN, GEMM_STAGES, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PCK(cudaFuncSetAttribute(k_gemm_tn<GEMM_BN, GEMM_STAGES, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PCK(cudaFuncSetAttribute(k_gemm_tn<GEMM_BN, GEMM_STAGES, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PCK(cudaFuncSetAttribute(k_gemm_tn<GEMM_BN, GEMM_STAGES, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
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
  CUtensorMap ma, mb;
  if (make_map(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM)) return -1;
  if (make_map(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, GEMM_BN)) return -1;
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
  if (mode != 2 || splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  g.kblocks_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + g.kblocks_per_split - 1) / g.kblocks_per_split;
  g.bias = bias; g.h = (const __nv_bfloat16*)h; g.ldh = ldh; g.out = (__nv_bfloat16*)out; g.ldo = ldo;
  g.out_t = (__nv_bfloat16*)out_t; g.ldt = ldt; g.outf = outf; g.ldf = ldf;
  dim3 grid((N + GEMM_BN - 1) / GEMM_BN, (M + GEMM_BM - 1) / GEMM_BM, splits);
  size_t smem = sizeof(GSm) + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case 0: k_gemm_tn<GEMM_BN, GEMM_STAGES, 0><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, g); break;
    case 1: k_gemm_tn<GEMM_BN, GEMM_STAGES, 1><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, g); break;
    case 2: k_gemm_tn<GEMM_BN, GEMM_STAGES, 2><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, g); break;
    case 3: k_gemm_tn<GEMM_BN, GEMM_STAGES, 3><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, g); break;
    default: sdx_set_error("sdx_gemm_bf16_tn: bad mode"); return -1;
  }
  PCK(cudaGetLastError());
  return 0;
}
