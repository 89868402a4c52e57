// sdx_gemm.cuh -- bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), the dense
// contraction of the PPO MLPs:   D[M,N] = A[M,K] . B[N,K]^T   (both K-major).
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d, 128B-swizzled 64-column K slabs, mbarrier tx)
//   warp 1      : TMEM allocator + MMA issuer (one thread issues tcgen05.mma.cta_group::1.kind::f16,
//                 UMMA 128 x BN x 16, accumulator in TMEM; tcgen05.commit frees smem stages)
//   warps 2..9  : epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> smem boxes -> TMA store);
//                 two warps per TMEM lane group, each owning a 64-column slice (latency-bound otherwise)
//
// Fused epilogues (MODE):
//   0  forward      out = ELU(acc + bias[n])               -> bf16 [M,ldo] (+ optional transposed bf16 [N,ldt])
//   1  backward dX  out = acc * ELU'(h[m,n])  (h = the layer's own ELU output; ELU'(h) = h > 0 ? 1 : h + 1)
//                                                          -> bf16 [M,ldo] (+ optional transposed bf16 [N,ldt])
//   2  backward dW  out += acc   (fp32 red.global.add, split-K over blockIdx.z)  -> fp32 [M,ldf]
//   3  plain        out = acc                                                   -> fp32 [M,ldf]
//   4  head         out = acc + bias[n]   (linear output layer: mu / value)      -> fp32 [M,ldf]
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define GEMM_BM 128
#define GEMM_BK 64
#define GEMM_THREADS 320   // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue

struct GemmArgs {
  int M, N, K;              // problem (K = total reduction length; split-K slices it by gridDim.z)
  int kblocks_per_split;    // BK-blocks each z-slice reduces
  const float* bias;        // MODE 0
  const __nv_bfloat16* h; int ldh;          // MODE 1
  __nv_bfloat16* out; int ldo;              // MODE 0/1 row-major
  __nv_bfloat16* out_t; int ldt;            // MODE 0/1 transposed copy (may be null)
  float* outf; int ldf;                     // MODE 2/3
};

namespace gemm {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t a = smem_u32(b), done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, 128B-swizzled operand tile: rows 128 B apart, 8-row groups 1024 B apart (SBO), LBO = 1 (16 B), version 1
__device__ __forceinline__ uint64_t umma_desc(const void* smem_tile) {
  uint64_t d = (uint64_t)((smem_u32(smem_tile) & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the leader CTA (rank 0) issues one M = 256 MMA over both SMs' shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the LEADER's mbarrier (peer bit of the barrier address cleared)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// MMA-completion arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
               ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// ELU without branches or denormal fix-ups: exp only ever sees min(a, 0), so ex2.approx.ftz is exact enough for a bf16 result
__device__ __forceinline__ float elu_fast(float a) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(a, 0.0f) * 1.4426950408889634f));
  return a > 0.0f ? a : e - 1.0f;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
}  // namespace gemm

template <int BN, int STAGES, int BROWS = BN>
struct GemmSmem {
  __nv_bfloat16 a[STAGES][GEMM_BM * GEMM_BK];
  __nv_bfloat16 b[STAGES][BROWS * GEMM_BK];   // BROWS = BN, or BN / 2 when a CTA pair shares the B tile
  unsigned char stage_rm[8 * 4096];      // epilogue staging, row-major box per warp: [32 rows][64 cols], 128B-swizzled
  unsigned char stage_t[8 * 4096];       // epilogue staging, transposed box per warp: [64 n][32 m], 64B-swizzled
  float bias[8][BN / 2];                 // per-warp copy of its bias slices (MODE 0)
  uint64_t full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], hbar[8];
  uint32_t tmem_base;
};

extern __shared__ unsigned char gsm_raw[];

// PERSISTENT: each CTA (CTA2: each CTA PAIR) walks tiles t = first, first + stride, ...  The TMA ring and the MMA issuer run
// ahead into the next tile while the epilogue warps drain the previous accumulator (two TMEM accumulators of BN columns).
// CTA2: a cluster of two CTAs owns a 256 x BN tile.  Each CTA loads its own 128 rows of A and HALF of the B tile; the leader
// issues tcgen05.mma.cta_group::2 (M = 256), which reads both halves of B across the pair -- per output element each SM
// pulls half as many operand bytes out of L2 (the bound of the single-CTA kernel); each CTA drains its own 128 TMEM lanes.
template <int BN, int STAGES, int MODE, bool CTA2>
__device__ __forceinline__ void gemm_body(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapO, const CUtensorMap& mapT,
                                          const CUtensorMap& mapH, const GemmArgs& g) {
  using namespace gemm;
  constexpr int BROWS = CTA2 ? BN / 2 : BN;
  constexpr int TILE_M = CTA2 ? 2 * GEMM_BM : GEMM_BM;
  // 128B-swizzled TMA/UMMA tiles need 1024 B alignment: align by hand (the launcher over-allocates 1 KB)
  auto& S = *reinterpret_cast<GemmSmem<BN, STAGES, BROWS>*>(gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;           // 0 = leader
  const int first = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, stride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tiles_n = (g.N + BN - 1) / BN, tiles_m = (g.M + TILE_M - 1) / TILE_M;
  const int total_kb = (g.K + GEMM_BK - 1) / GEMM_BK;
  const int splits = (total_kb + g.kblocks_per_split - 1) / g.kblocks_per_split;
  const int n_tiles = tiles_n * tiles_m * splits;
  constexpr uint32_t STAGE_BYTES = (CTA2 ? 2u : 1u) * (GEMM_BM + BROWS) * GEMM_BK * 2;      // bytes landing per stage (both CTAs)
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&S.tmem_full[a], 1); mbar_init(&S.tmem_empty[a], CTA2 ? 16 : 8); }
    for (int a = 0; a < 8; ++a) mbar_init(&S.hbar[a], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CTA2) { __syncthreads(); cluster_sync_all(); }             // barriers of both CTAs exist before anything remote touches them
  if (warp == 1) {
    if (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(2 * BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(2 * BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kbg = 0;                                  // k-block counter across all tiles of this CTA (stage ring position)
      for (int t = first; t < n_tiles; t += stride) {
        const int nx = t % tiles_n, my = (t / tiles_n) % tiles_m, z = t / (tiles_n * tiles_m);
        const int m0 = my * TILE_M + (int)rank * GEMM_BM, n0 = nx * BN + (int)rank * BROWS, kb0 = z * g.kblocks_per_split;
        const int nkb = min(g.kblocks_per_split, total_kb - kb0);
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          const int s = kbg % STAGES;
          const uint32_t ph = (kbg / STAGES) & 1;
          mbar_wait(&S.empty[s], ph ^ 1);
          if (CTA2) {                                      // both CTAs load; the bytes of both are expected on the leader's barrier
            if (rank == 0) mbar_expect_tx(&S.full[s], STAGE_BYTES);
            tma_load_2d_pair(&mapA, S.a[s], &S.full[s], (kb0 + kb) * GEMM_BK, m0);
            tma_load_2d_pair(&mapB, S.b[s], &S.full[s], (kb0 + kb) * GEMM_BK, n0);
          } else {
            mbar_expect_tx(&S.full[s], STAGE_BYTES);
            tma_load_2d(&mapA, S.a[s], &S.full[s], (kb0 + kb) * GEMM_BK, m0);
            tma_load_2d(&mapB, S.b[s], &S.full[s], (kb0 + kb) * GEMM_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      uint32_t kbg = 0, it = 0;
      for (int t = first; t < n_tiles; t += stride, ++it) {
        const int z = t / (tiles_n * tiles_m), kb0 = z * g.kblocks_per_split;
        const int nkb = min(g.kblocks_per_split, total_kb - kb0);
        const uint32_t acc = it & 1;
        mbar_wait(&S.tmem_empty[acc], ((it >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          const int s = kbg % STAGES;
          const uint32_t ph = (kbg / STAGES) & 1;
          mbar_wait(&S.full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = umma_desc(S.a[s]), db = umma_desc(S.b[s]);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)   // +32 B (= 2 x 16 B) along K inside the 128 B swizzle row per UMMA_K
            if (CTA2) umma_f16_pair(tmem + acc * BN, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0 ? 1u : 0u);
            else umma_f16(tmem + acc * BN, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0 ? 1u : 0u);
          if (CTA2) umma_commit_pair(&S.empty[s]); else umma_commit(&S.empty[s]);
        }
        if (CTA2) umma_commit_pair(&S.tmem_full[acc]); else umma_commit(&S.tmem_full[acc]);
      }
    }
  } else {
    // ---------------- epilogue: 8 warps.  Warp w reads TMEM lanes [32 (w % 4), +32) (= 32 output rows of the tile) and,
    // of every 128-column span of the tile, the 64-column slice ch = (w - 2) / 4: two 32-column TMEM loads per slice.
    const int lg = warp & 3, ch = (warp - 2) >> 2, ew = ch * 4 + lg;
    unsigned char* const rm = S.stage_rm + ew * 4096;      // [32 rows][64 cols] box, 128B-swizzled
    unsigned char* const tb = S.stage_t + ew * 4096;       // [64 n][32 m] box, 64B-swizzled
    uint32_t it = 0;
    [[maybe_unused]] uint32_t hph = 0;                     // uses of this warp's h barrier so far
    for (int t = first; t < n_tiles; t += stride, ++it) {
      const int nx = t % tiles_n, my = (t / tiles_n) % tiles_m;
      const int m0 = my * TILE_M + (int)rank * GEMM_BM, n0 = nx * BN;
      const uint32_t acc = it & 1;
      if (MODE == 0) {                                     // this warp's copy of its bias slices
#pragma unroll
        for (int q = lane; q < BN / 2; q += 32) {
          const int col = n0 + (q >> 6) * 128 + ch * 64 + (q & 63);
          S.bias[ew][q] = col < g.N ? g.bias[col] : 0.0f;
        }
        __syncwarp();
      }
      // MODE 1 reads the layer's own activations h through the SAME staging box the results leave by: one TMA load
      // (full, coalesced lines), each thread then reads and overwrites exactly its own 16-byte units.
      auto stage_slice = [&](int ns) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // earlier TMA stores have read the staging area
          if (MODE == 1) {
            mbar_expect_tx(&S.hbar[ew], 4096u);
            tma_load_2d(&mapH, rm, &S.hbar[ew], ns, m0 + lg * 32);
          }
        }
        __syncwarp();
      };
      if ((MODE == 0 || MODE == 1) && n0 + ch * 64 < g.N) stage_slice(n0 + ch * 64);
      mbar_wait(&S.tmem_full[acc], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + lg * 32 + lane;
      const bool row_ok = row < g.M;
#pragma unroll 1
      for (int sp = 0; sp < BN / 128; ++sp) {
        const int ns = n0 + sp * 128 + ch * 64;             // first column of this warp's slice
        if (ns >= g.N) break;
        if (MODE == 0 || MODE == 1) {
          if (sp > 0) stage_slice(ns);
          if (MODE == 1) { mbar_wait(&S.hbar[ew], hph & 1); ++hph; }
        }
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int col0 = ns + cc * 32;
          if (col0 >= g.N) break;
          uint32_t r[32];
          tmem_ld32(tmem + acc * BN + ((uint32_t)(lg * 32) << 16) + (uint32_t)(sp * 128 + ch * 64 + cc * 32), r);
          if (MODE == 0 || MODE == 1) {
            float v[32];
            if (MODE == 0) {
              const float4* bq = reinterpret_cast<const float4*>(&S.bias[ew][sp * 64 + cc * 32]);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = bq[j >> 2];                // broadcast read
                v[j] = elu_fast(__uint_as_float(r[j]) + b4.x); v[j + 1] = elu_fast(__uint_as_float(r[j + 1]) + b4.y);
                v[j + 2] = elu_fast(__uint_as_float(r[j + 2]) + b4.z); v[j + 3] = elu_fast(__uint_as_float(r[j + 3]) + b4.w);
              }
            } else {
              const unsigned char* hb = rm + lane * 128;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u = *reinterpret_cast<const uint4*>(hb + (((cc * 4 + q) ^ (lane & 7)) << 4));
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  float2 hv = __bfloat1622float2(h2[e2]);
                  int j = 8 * q + 2 * e2;
                  v[j] = __uint_as_float(r[j]) * (hv.x > 0.0f ? 1.0f : hv.x + 1.0f);
                  v[j + 1] = __uint_as_float(r[j + 1]) * (hv.y > 0.0f ? 1.0f : hv.y + 1.0f);
                }
              }
            }
            // stage the bf16 results in shared memory in the layouts of 128B- / 64B-swizzled TMA boxes;
            // one TMA store per box then writes full, coalesced lines
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]), p1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]), p3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
              uint4 u;
              u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
              u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(rm + lane * 128 + (((cc * 4 + q) ^ (lane & 7)) << 4)) = u;
            }
            if (g.out_t) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int nl = cc * 32 + j;
                *reinterpret_cast<__nv_bfloat16*>(tb + nl * 64 + ((((lane >> 3) ^ ((nl >> 1) & 3))) << 4) + (lane & 7) * 2) = __float2bfloat16_rn(v[j]);
              }
            }
          } else if (row_ok) {
            float* dst = g.outf + (size_t)row * g.ldf + col0;
            if (MODE == 2 && col0 + 32 <= g.N && (g.ldf & 3) == 0) {      // split-K accumulation: 16-byte vector reductions
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                atomicAdd(reinterpret_cast<float4*>(dst + j),
                          make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < g.N) {
                  if (MODE == 2) atomicAdd(dst + j, __uint_as_float(r[j]));
                  else if (MODE == 4) dst[j] = __uint_as_float(r[j]) + g.bias[col0 + j];
                  else dst[j] = __uint_as_float(r[j]);
                }
            }
          }
        }
        if (MODE == 0 || MODE == 1) {                        // slice complete: one TMA store per staged box
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapO, rm, ns, m0 + lg * 32);
            if (g.out_t) tma_store_2d(&mapT, tb, m0 + lg * 32, ns);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      // this warp has read its TMEM lanes of the accumulator: hand it back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_cluster(&S.tmem_empty[acc], 0u);   // the leader issues the MMAs of both CTAs
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.tmem_empty[acc])) : "memory");
      }
    }
    if ((MODE == 0 || MODE == 1) && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CTA2) cluster_sync_all();                                  // the peer may still be reading this CTA's half of B / signalling its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN) : "memory");
  }
}

template <int BN, int STAGES, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tn(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapO,
          const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapH, const GemmArgs g) {
  gemm_body<BN, STAGES, MODE, false>(mapA, mapB, mapO, mapT, mapH, g);
}
// CTA-pair launch: grid = 2 x pairs, cluster (2,1,1)
template <int BN, int STAGES, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tn2(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapO,
           const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapH, const GemmArgs g) {
  gemm_body<BN, STAGES, MODE, true>(mapA, mapB, mapO, mapT, mapH, g);
}
