// sdx_sim.cuh -- the contact step (replaces gym.simulate(), BT:140) as ONE kernel launch per control
// step: one CTA per environment, the whole per-env working set resident in shared memory for both
// sub-steps and all 2 x 16 solver iterations; HBM is touched only to load / store the env's state tile.
//
//   load   : free-brick tile [13][72] f32 (3744 B, contiguous per env) by TMA bulk copy
//            (cp.async.bulk -> mbarrier), DoF block [3][24] by plain loads
//   per sub-step (h = dt/substeps):
//     forward kinematics (robot warp in lockstep, chain by shuffles), shape poses, free velocities (gravity, implicit PD),
//     world-frame inverse inertia of every brick, world AABBs -- ONE block barrier: the robot warp runs the articulation's whole chain
//     broad phase   : FIRST sub-step of a step (or after a sleeping brick was woken): every unordered pair of moving shapes is box-tested
//                     ONCE (ring pairing, two threads per owner) with the travel bounds of all sub-steps left, hits entered in both shapes'
//                     128-bit rows; the candidate lists are read off the rows in ascending order; later sub-steps keep the lists
//     narrow phase  : ordered pairs -> SAT reference face -> sample points -> contacts (two-pass,
//                     deterministic compaction: owner, candidate, point order) + warm-start lookup (galloping from the same slot);
//                     EDGE instantiation: pairs without a parallel axis pair are queued and the nine edge-pair axes tested one pair per
//                     thread (pass 1b); a pair's edge-edge contact is constructed one contact per thread (pass 2e)
//     -- a sub-step without a single contact skips the three stages below --
//     CSR incidence : per body, contacts in index order (fixed summation order => reproducible); 16-byte work items for phase B
//     solver        : mass-splitting Jacobi on total impulses; phase A = 1 thread / contact,
//                     phase B = 2 lanes / awake touched brick gather its incident impulses; the articulation sees
//                     contacts through per-link wrenches -> joint-space impulses (diag. inertia), links in contact only
//     integrate
//   store  : brick tile by TMA bulk store, DoF block, and ONLY the rows the task reads
//            (24 link rows, 6x7 Jacobian of link7, net contact force per link)
#pragma once
#include "sdx_math.cuh"
#include "../../include/seqdex_b200.h"

#define NB SDX_MAX_BRICKS
#define NBODY (NB + SDX_NL)
#define KSTAT 24                       /* static boxes the kernel keeps in shared memory */
#define NOWN (NB + SDX_MAX_RSHAPES)    /* owner shapes: bricks + robot boxes */
#define NT (NOWN + KSTAT)              /* target boxes */
#define KC 32
#define MAXC SDX_MAX_CONTACTS
#define STATIC_BODY 255
#ifndef SIM_THREADS
#define SIM_THREADS 256
#endif
// SIM_GLOBAL_CONTACTS: the two contact-record arrays that are READ-ONLY during the solver iterations (point | bias and
// inverse masses | packed word, 32 KB per env at MAXC = 1024) live in global memory behind L1 instead of shared memory:
// 42.6 KB of shared memory per CTA instead of 74.6 KB and a 64-register budget put FOUR CTAs on an SM instead of three.
// (line-level ncu: 45 % of the warp stall samples are block-barrier waits, the issue slots are busy 42 % of the time --
// resident warps are the lever)  Same arithmetic, same results.
#ifndef SIM_GLOBAL_CONTACTS
#define SIM_GLOBAL_CONTACTS 1
#endif
#ifndef SIM_MIN_CTAS
#define SIM_MIN_CTAS (SIM_GLOBAL_CONTACTS ? 4 : 3)   /* CTAs per SM the register budget is sized for */
#endif
// SIM_GLOBAL_CF: the impulse array (and the scratch tables that share its storage) also live in the per-env global scratch behind L1:
// 30.5 KB of shared memory per CTA instead of 46.5 KB.  Together with 128-thread CTAs that puts SIX envs on an SM instead of four
// (the SM's 256 KB of shared memory + L1 holds about six envs' working sets).
#ifndef SIM_GLOBAL_CF
#define SIM_GLOBAL_CF (SIM_THREADS == 128)
#endif
#define ROBOT_TID0 (SIM_THREADS - 32)  /* the LAST warp owns the articulation: lane j = DoF j, lane L = link L */
#define NBW (ROBOT_TID0 / 32)          /* brick warps */
#define SIM_PROF_REC (18 * 2 * 8 + 16)  /* SIM_PROFILE builds: int64 counters per env and sub-step */
#ifndef SIM_DEAL_RR
#define SIM_DEAL_RR 0                  /* phase B: 16 consecutive work items per warp (1: dealt round-robin over the brick warps -- measured SLOWER,
                                          5.40 vs 5.18 ms per launch: every warp then runs the loop for about the same maximal trip count) */
#endif
#ifndef SIM_SMALL_CODE
#define SIM_SMALL_CODE 1               /* out-of-line FK, rolled joint / axis loops (see robot_fk) */
#endif
#if SIM_SMALL_CODE
#define SIM_FK_INLINE __noinline__
#else
#define SIM_FK_INLINE
#endif
#ifndef SIM_BROAD_ROLLED
#define SIM_BROAD_ROLLED 1             /* broad-phase target loops kept rolled (code size; A/B profiles/r01_ab_code_size.txt) */
#endif
#if SIM_BROAD_ROLLED
#define SIM_BROAD_UNROLL _Pragma("unroll 1")
#else
#define SIM_BROAD_UNROLL
#endif
#ifndef SIM_SCAN_OUTLINE
#define SIM_SCAN_OUTLINE 1             /* one out-of-line copy of the warp scan */
#endif
#if SIM_SCAN_OUTLINE
#define SIM_SCAN_INLINE __noinline__
#else
#define SIM_SCAN_INLINE __forceinline__
#endif
#ifndef SIM_FK_WARP
#define SIM_FK_WARP 1                  /* forward kinematics by the whole robot warp in lockstep (0: one lane per chain) */
#endif
#ifndef SIM_GATHER_U
#define SIM_GATHER_U 1                 /* incidences whose loads phase B issues together per lane (2 / 4 measured slower: 5.56 / 5.87 vs 5.18 ms) */
#endif
#define PB_LANES_MAX 1152              /* sum of phase-B lane groups: < NB + (2 MAXC) / 2, rounded up to whole warps */
#ifndef SIM_KIN_FUSED
#define SIM_KIN_FUSED 1                  /* kinematics .. world AABBs behind one block barrier: the robot warp runs the articulation's chain on its own */
#endif
#ifndef SIM_BROAD_SYM
#define SIM_BROAD_SYM 1                  /* broad phase: every unordered pair of moving shapes tested once, hits entered in both shapes' bit rows */
#endif
#ifndef SIM_PASS1_STRIDED
#define SIM_PASS1_STRIDED 1
#endif
#define PPMAX 26                       /* pairs per thread in the narrow phase (SIM_THREADS*PPMAX >= pairs) */

struct SimSmem {
  union {                              // lifetimes do not overlap: tile = kernel prologue / epilogue, candidates = broad..narrow phase
    float tile[13 * NB];               // TMA landing / staging zone for the brick tile
    struct { unsigned char cand[NOWN][KC]; int ncand[NOWN]; };
  };
  unsigned long long mbar;
  float sc[NT][3], sR[NT][9], sh[NT][3], srad[NOWN];
  union {                              // world AABBs + travel bounds live until the narrow phase; the target-side lists after it
    float4 sab[NT];                    // world-AABB half extents .xyz | per-sub-step travel bound .w
    unsigned short blist[MAXC];
  };
  unsigned char sbody[NT];
  float4 bx[NBODY], bv[NBODY], bw[NBODY];      // body origin, linear, angular velocity (16-byte records)
  // per-brick records of phase B's tail, written once per sub-step (16-byte vectors: one LDS.128 each)
  float4 bfv[NB];   // free linear velocity .xyz | inverse mass (0 while asleep)
  float4 bfw[NB];   // free angular velocity .xyz
  float4 bI0[NB];   // world-frame inverse inertia R diag(1/I) R^T: xx xy xz yy
  float4 bI1[NB];   //                                              yz zz
  float4 bq[NB];    // COMPOUND scenes only: body orientation, read by the boxes that ride on the body
  int4 irec[NB];    // phase-B record of brick b (valid while it is awake and touched): a0 | na | b0 | b + (ntot << 8) + (log2 lanes << 20)
  float lq[SDX_NL][4], ja[SDX_ND][3], jo[SDX_ND][3];
  float q[SDX_ND + 1], qd[SDX_ND + 1], tgt[SDX_ND + 1], qdfree[SDX_ND + 1], ieff[SDX_ND + 1];
  float linkF[SDX_NL][3], linkM[SDX_NL][3];
  int nb[NBODY], nj[SDX_ND + 1];
  // incidence of body b in summation order: owned contacts [astart, aend) then the target-side list (ascending)
  int astart[NBODY], aend[NBODY], boff[NBODY + 1], bcur[NBODY];
  int poff[NOWN + 1];
  int scan[SIM_THREADS];
  int ncon, ndropped;                  // contacts of the sub-step | contacts beyond MAXC (after shedding the speculative ones)
  int ndrop_cand, ndrop_static;        // candidate pairs beyond KC when the lists were last built | of those, pairs against statics (statics claim their slots first: 0)
  int nelist;                          // EDGE: pairs queued for the edge-axis test of this sub-step (narrow phase, pass 1b)
  int nact;                            // lanes phase B occupies: sum over awake, touched bricks of their lane-group size
  unsigned char lmap[PB_LANES_MAX];    // phase-B lane -> brick
  unsigned char sflag[NB], touch[NB];  // sleeping: sflag bit0 = asleep this sub-step, bit1 = hot at its start; touch bit0 = robot, bit1 = hot brick
  // contact records as three 16-byte vectors (one LDS.128 / STS.128 each)
#if !SIM_GLOBAL_CONTACTS
  float4 ca[MAXC];    // contact point w.xyz | bias
  float4 cb[MAXC];    // 1/den along n, t1, t2 | packed word (bodies, target shape, face axis, sign)
#endif
#if !SIM_GLOBAL_CF
  float4 cf4[MAXC];   // total impulse f.xyz | unused          (also scratch: candidate overflow lists, pair tables)
#endif
};

static_assert(SIM_THREADS == 256 || SIM_THREADS == 128, "the broad phase deals (owner, half) pairs over 256 or 128 threads");
static_assert(NT <= 128 && NOWN <= 128, "the broad phase keeps one 128-bit hit row per owner over the target index space");
static_assert(NOWN * KC * 2 <= 8192 && 8192 + NOWN * KC * 2 <= MAXC * 16 && NOWN * KC + NOWN * 16 + NOWN * KSTAT <= MAXC * 16 && KSTAT <= KC,
              "the pair tables / candidate scratch lists are laid out inside the impulse array (cf4)");

// exclusive prefix sum of arr[0..n) (n <= 256) by ONE warp, in place; returns the total to every lane
__device__ SIM_SCAN_INLINE int warp_excl_scan(int* arr, int n, int lane) {
  constexpr int PER = 8;
  int v[PER], s = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) { int k = lane * PER + i; v[i] = k < n ? arr[k] : 0; s += v[i]; }
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  int run = incl - s;
#pragma unroll
  for (int i = 0; i < PER; ++i) { int k = lane * PER + i; if (k < n) arr[k] = run; run += v[i]; }
  return __shfl_sync(0xffffffffu, incl, 31);
}

__device__ __forceinline__ v3 ld3(const float* p) { return V3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ v3 ldv(const float4& p) { float4 t = p; return V3(t.x, t.y, t.z); }          // one LDS.128
__device__ __forceinline__ v3 ld3(const float4& p) { return ldv(p); }
__device__ __forceinline__ void st3(float4& p, v3 a) { p = make_float4(a.x, a.y, a.z, 0.0f); }   // one STS.128

// forward kinematics of the collapsed Panda+Allegro tree by ONE WARP: lane 0 walks the 7 arm joints, then lanes 0-3
// walk the four finger chains (DoFs 7+4f .. 10+4f) concurrently -> dependency depth 11 instead of 23
__device__ __forceinline__ void fk_joint(const sdx_scene_t* __restrict__ S, SimSmem& M, int j) {
  int L = j + 1, P = S->body_parent[L];
  q4 qP = Q4(M.lq[P][0], M.lq[P][1], M.lq[P][2], M.lq[P][3]);
  q4 qf = Q4(S->joint_quat[4 * j], S->joint_quat[4 * j + 1], S->joint_quat[4 * j + 2], S->joint_quat[4 * j + 3]);
  v3 ax = V3(S->joint_axis[3 * j], S->joint_axis[3 * j + 1], S->joint_axis[3 * j + 2]);
  q4 qj = qmul(qP, qf);
  v3 x = vadd(ld3(M.bx[NB + P]), qrot(qP, V3(S->joint_xyz[3 * j], S->joint_xyz[3 * j + 1], S->joint_xyz[3 * j + 2])));
  float s, c;
  sdx_sincos(0.5f * M.q[j], &s, &c);
  q4 qr = Q4(ax.x * s, ax.y * s, ax.z * s, c);
  q4 ql = qmul(qj, qr);
  M.lq[L][0] = ql.x; M.lq[L][1] = ql.y; M.lq[L][2] = ql.z; M.lq[L][3] = ql.w;
  st3(M.bx[NB + L], x);
  st3(M.ja[j], qrot(qj, ax));
  st3(M.jo[j], x);
}
// Code size is a performance number in this kernel (128 KB of SASS against the SM's instruction cache, four CTAs in different
// phases: sm__icc_request_hit_rate 83 %): ONE out-of-line copy, joint loops kept rolled -- the chain is serial anyway.
#if SIM_FK_WARP
// The whole warp walks the tree in lockstep.  What a joint needs that does NOT depend on its parent -- its constants, the sine
// and cosine of its angle -- is prepared by lane j for all 23 joints at once and handed to the chain by shuffle; the chain
// itself carries only what is serial: two quaternion products per joint for the link frames (7 arm steps, then the four
// finger chains side by side in lanes 0-3), and later one vector add per joint for the origins.  Joint axes and offsets are
// again one joint per lane.  Every product is the one fk_joint forms, operand for operand: same bits, a third of the
// instructions on the critical path (a lone lane used to issue all ~150 instructions of each of 11 chained joints).
__device__ __forceinline__ q4 shfl_q4(q4 a, int src) {
  return Q4(__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src), __shfl_sync(0xffffffffu, a.z, src),
            __shfl_sync(0xffffffffu, a.w, src));
}
__device__ SIM_FK_INLINE void robot_fk(const sdx_scene_t* __restrict__ S, SimSmem& M, int lane) {
  const int j = lane < SDX_ND ? lane : 0;
  const q4 qf = Q4(S->joint_quat[4 * j], S->joint_quat[4 * j + 1], S->joint_quat[4 * j + 2], S->joint_quat[4 * j + 3]);
  const v3 ax = V3(S->joint_axis[3 * j], S->joint_axis[3 * j + 1], S->joint_axis[3 * j + 2]);
  float sn, cs;
  sdx_sincos(0.5f * M.q[j], &sn, &cs);
  const q4 qr = Q4(ax.x * sn, ax.y * sn, ax.z * sn, cs);
  q4 qP = Q4(S->base_quat[0], S->base_quat[1], S->base_quat[2], S->base_quat[3]);
  if (lane == 0) { M.lq[0][0] = qP.x; M.lq[0][1] = qP.y; M.lq[0][2] = qP.z; M.lq[0][3] = qP.w; }
#pragma unroll 1
  for (int i = 0; i < 11; ++i) {                               // steps 0-6: the arm (all lanes alike); 7-10: finger (lane & 3)
    const int src = i < 7 ? i : 4 * (lane & 3) + i;            // = 7 + 4 f + (i - 7)
    const q4 ql = qmul(qmul(qP, shfl_q4(qf, src)), shfl_q4(qr, src));
    if (lane == 0 || (i >= 7 && lane < 4)) { M.lq[src + 1][0] = ql.x; M.lq[src + 1][1] = ql.y; M.lq[src + 1][2] = ql.z; M.lq[src + 1][3] = ql.w; }
    qP = ql;
  }
  __syncwarp();
  const int P = S->body_parent[j + 1];
  const q4 qPj = Q4(M.lq[P][0], M.lq[P][1], M.lq[P][2], M.lq[P][3]);
  const v3 d = qrot(qPj, V3(S->joint_xyz[3 * j], S->joint_xyz[3 * j + 1], S->joint_xyz[3 * j + 2]));
  if (lane < SDX_ND) st3(M.ja[j], qrot(qmul(qPj, qf), ax));
  v3 x = V3(S->base_pos[0], S->base_pos[1], S->base_pos[2]);
  if (lane == 0) st3(M.bx[NB], x);
#pragma unroll 1
  for (int i = 0; i < 11; ++i) {
    const int src = i < 7 ? i : 4 * (lane & 3) + i;
    x = vadd(x, V3(__shfl_sync(0xffffffffu, d.x, src), __shfl_sync(0xffffffffu, d.y, src), __shfl_sync(0xffffffffu, d.z, src)));
    if (lane == 0 || (i >= 7 && lane < 4)) { st3(M.bx[NB + src + 1], x); st3(M.jo[src], x); }
  }
  __syncwarp();
}
#else
__device__ SIM_FK_INLINE void robot_fk(const sdx_scene_t* __restrict__ S, SimSmem& M, int lane) {
  if (lane == 0) {
    st3(M.bx[NB], V3(S->base_pos[0], S->base_pos[1], S->base_pos[2]));
    M.lq[0][0] = S->base_quat[0]; M.lq[0][1] = S->base_quat[1]; M.lq[0][2] = S->base_quat[2]; M.lq[0][3] = S->base_quat[3];
#pragma unroll 1
    for (int j = 0; j < 7; ++j) fk_joint(S, M, j);
  }
  __syncwarp();
  if (lane < 4) {
#pragma unroll 1
    for (int j = 7 + 4 * lane; j < 11 + 4 * lane; ++j) fk_joint(S, M, j);
  }
  __syncwarp();
}

#endif

__device__ __forceinline__ void link_twist(const sdx_scene_t* __restrict__ S, SimSmem& M, int L, unsigned m) {
  v3 w = V3(0.0f, 0.0f, 0.0f), v = V3(0.0f, 0.0f, 0.0f);
  v3 xl = ld3(M.bx[NB + L]);
  while (m) {                       // ancestors in ascending DoF order
    int j = __ffs(m) - 1; m &= m - 1;
    v3 a = ld3(M.ja[j]);
    float qd = M.qd[j];
    w = vmad(a, qd, w);
    v = vmad(vcross(a, vsub(xl, ld3(M.jo[j]))), qd, v);
  }
  st3(M.bv[NB + L], v); st3(M.bw[NB + L], w);
}

// world-frame inverse inertia of a brick, W = R diag(invI) R^T, as six numbers (xx xy xz yy | yz zz) computed ONCE per sub-step
// (the oracle's brick_world_invI, operation for operation), and its product with a vector
__device__ __forceinline__ void brick_world_invI(const float* R, v3 invI, float4* w0, float4* w1) {
  const float a0x = R[0] * invI.x, a0y = R[1] * invI.y, a0z = R[2] * invI.z;
  const float a1x = R[3] * invI.x, a1y = R[4] * invI.y, a1z = R[5] * invI.z;
  const float a2x = R[6] * invI.x, a2y = R[7] * invI.y, a2z = R[8] * invI.z;
  *w0 = make_float4(fmaf(a0z, R[2], fmaf(a0y, R[1], a0x * R[0])), fmaf(a0z, R[5], fmaf(a0y, R[4], a0x * R[3])),
                    fmaf(a0z, R[8], fmaf(a0y, R[7], a0x * R[6])), fmaf(a1z, R[5], fmaf(a1y, R[4], a1x * R[3])));
  *w1 = make_float4(fmaf(a1z, R[8], fmaf(a1y, R[7], a1x * R[6])), fmaf(a2z, R[8], fmaf(a2y, R[7], a2x * R[6])), 0.0f, 0.0f);
}
__device__ __forceinline__ v3 brick_Iinv_mul(const float4 w0, const float4 w1, v3 u) {
  return V3(fmaf(w0.z, u.z, fmaf(w0.y, u.y, w0.x * u.x)), fmaf(w1.x, u.z, fmaf(w0.w, u.y, w0.y * u.x)),
            fmaf(w1.y, u.z, fmaf(w1.x, u.y, w0.z * u.x)));
}

// phase B's inner loop: lane k of a body's group of L lanes sums the impulses (and their moments about xb) of the incidences
// e = k, k + L, k + 2 L, ... of that body -- owned contacts [a0, a0 + na) first, then the ascending target-side list at b0 --
// strictly in that order (the oracle's Fk[k] / Tk[k]).
__device__ __forceinline__ void gather_incident(const SimSmem& M, const float4* CA, const float4* CF, int a0, int na, int b0, int ntot,
                                                int k, int L, v3 xb, v3* Fo, v3* To) {
  v3 F = V3(0.0f, 0.0f, 0.0f), T = V3(0.0f, 0.0f, 0.0f);
#pragma unroll 1
  for (int ee = k; ee < ntot; ee += L) {
    const int idx = ee < na ? a0 + ee : (int)M.blist[b0 + (ee - na)];
    const float4 F4 = CF[idx], A4 = CA[idx];
    v3 f = V3(F4.x, F4.y, F4.z);
    if (ee >= na) f = vneg(f);
    F = vadd(F, f);
    T = vadd(T, vcross(vsub(V3(A4.x, A4.y, A4.z), xb), f));
  }
  *Fo = F; *To = T;
}
// lanes that share a brick's incidences in phase B: the smallest power of two that leaves each lane at most 4 (capped at a warp)
__device__ __forceinline__ int phaseb_lg(int ninc) {
  int lg = 0;
  while (lg < 5 && (4 << lg) < ninc) ++lg;
  return lg;
}

__device__ __forceinline__ void contact_axes(const SimSmem& M, uint32_t word, v3* n, v3* t1, v3* t2) {
  int sh = (word >> 16) & 255, k = (word >> 24) & 3;
  float sg = ((word >> 26) & 1) ? -1.0f : 1.0f;
  *n = vscale(mcol(M.sR[sh], k), sg);
  *t1 = mcol(M.sR[sh], (k + 1) % 3);
  *t2 = mcol(M.sR[sh], (k + 2) % 3);
}
// edge-edge contacts carry their normal explicitly (CN); t1 = the target's edge direction (axis k of its box, perpendicular to n by
// construction), t2 = n x t1
template <bool EDGE>
__device__ __forceinline__ void contact_axes_e(const SimSmem& M, uint32_t word, const float4* CN, int i, v3* n, v3* t1, v3* t2) {
  if (EDGE && (word & (1u << 27))) {
    const float4 N4 = CN[i];
    *n = V3(N4.x, N4.y, N4.z);
    *t1 = mcol(M.sR[(word >> 16) & 255], (word >> 24) & 3);
    *t2 = vcross(*n, *t1);
  } else contact_axes(M, word, n, t1, t2);
}

// the same axes WITHOUT a branch, for the solver's phase A (a warp that holds one edge contact would otherwise run both paths in every
// pass): the explicit normal N4 = CN[i] is loaded for every contact (garbage for the others: selected away, never computed with)
template <bool EDGE>
__device__ __forceinline__ void contact_axes_sel(const SimSmem& M, uint32_t word, const float4 N4, v3* n, v3* t1, v3* t2) {
  if (!EDGE) { contact_axes(M, word, n, t1, t2); return; }
  const int sh = (word >> 16) & 255, k = (word >> 24) & 3;
  const float sg = ((word >> 26) & 1) ? -1.0f : 1.0f;
  const bool e = (word & (1u << 27)) != 0u;
  const v3 c0 = mcol(M.sR[sh], k), c1 = mcol(M.sR[sh], (k + 1) % 3), c2 = mcol(M.sR[sh], (k + 2) % 3);
  const v3 nf = vscale(c0, sg);
  *n = e ? V3(N4.x, N4.y, N4.z) : nf;
  *t1 = e ? c0 : c1;
  const v3 tx = vcross(*n, c0);
  *t2 = e ? tx : c2;
}

__device__ float body_k(const sdx_scene_t* __restrict__ S, const SimSmem& M, int body, v3 wpt, v3 d) {
  if (body == STATIC_BODY) return 0.0f;
  if (body < NB) {
    v3 rxd = vcross(vsub(wpt, ld3(M.bx[body])), d);
    float k = M.bfv[body].w + vdot(rxd, brick_Iinv_mul(M.bI0[body], M.bI1[body], rxd));
    return (float)M.nb[body] * k;
  }
  unsigned m = S->link_anc_mask[body - NB];
  float k = 0.0f;
  while (m) {
    int j = __ffs(m) - 1; m &= m - 1;
    float g = vdot(ld3(M.ja[j]), vcross(vsub(wpt, ld3(M.jo[j])), d));
    k = k + (float)M.nj[j] * (g * g) / M.ieff[j];
  }
  return k;
}

// pair-level SAT: reference face of target t for owner a.  returns false if separated beyond m.
struct PairGeom { v3 lc; float C[9]; int k; float sgf; uint32_t sg; float htk; v3 ha, ht; };
__device__ __forceinline__ bool pair_geom(const SimSmem& M, int a, int t, float m, PairGeom& G, bool t_static) {
  G.lc = mtmul(M.sR[t], vsub(ld3(M.sc[a]), ld3(M.sc[t])));
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c2 = 0; c2 < 3; ++c2)
      G.C[3 * r + c2] = M.sR[t][r] * M.sR[a][c2] + M.sR[t][3 + r] * M.sR[a][3 + c2] + M.sR[t][6 + r] * M.sR[a][6 + c2];
  G.ha = ld3(M.sh[a]); G.ht = ld3(M.sh[t]);
  float o0 = G.ht.x + (fabsf(G.C[0]) * G.ha.x + fabsf(G.C[1]) * G.ha.y + fabsf(G.C[2]) * G.ha.z) - fabsf(G.lc.x);
  float o1 = G.ht.y + (fabsf(G.C[3]) * G.ha.x + fabsf(G.C[4]) * G.ha.y + fabsf(G.C[5]) * G.ha.z) - fabsf(G.lc.y);
  float o2 = G.ht.z + (fabsf(G.C[6]) * G.ha.x + fabsf(G.C[7]) * G.ha.y + fabsf(G.C[8]) * G.ha.z) - fabsf(G.lc.z);
  int k = 0; float ov = o0;
  if (o1 < ov) { k = 1; ov = o1; }
  if (o2 < ov) { k = 2; ov = o2; }
  if (ov < -m) return false;
  float lck = k == 0 ? G.lc.x : (k == 1 ? G.lc.y : G.lc.z);
  G.k = k;
  G.sgf = lck >= 0.0f ? 1.0f : -1.0f;
  G.sg = lck >= 0.0f ? 0u : 1u;
  if (t_static && k == 2) { G.sgf = 1.0f; G.sg = 0u; }   // statics rest on each other: their z faces only push UP
  G.htk = k == 0 ? G.ht.x : (k == 1 ? G.ht.y : G.ht.z);
  return true;
}
__device__ __forceinline__ v3 sample_point(v3 ha, int p) {
  if (p < 8) return V3((p & 1) ? ha.x : -ha.x, (p & 2) ? ha.y : -ha.y, (p & 4) ? ha.z : -ha.z);
  return V3(0.0f, (p & 1) ? ha.y : -ha.y, (p & 2) ? ha.z : -ha.z);
}
// does sample point p of the owner touch the reference face?  depth returned through *depth.
// The coordinate along the face axis is evaluated first (most points fail the depth test); the values are the same
// as evaluating the full point l = lc + C pl.
__device__ __forceinline__ bool point_hit(const PairGeom& G, int p, float m, float margin, float* depth) {
  v3 pl = sample_point(G.ha, p);
  const int k = G.k;
  const float c0 = k == 0 ? G.C[0] : (k == 1 ? G.C[3] : G.C[6]), c1 = k == 0 ? G.C[1] : (k == 1 ? G.C[4] : G.C[7]),
              c2 = k == 0 ? G.C[2] : (k == 1 ? G.C[5] : G.C[8]);
  const float lck = k == 0 ? G.lc.x : (k == 1 ? G.lc.y : G.lc.z);
  float lk = lck + fmaf(c2, pl.z, fmaf(c1, pl.y, c0 * pl.x));
  float d = G.htk - G.sgf * lk;
  if (!(d > -m)) return false;
  v3 l = vadd(G.lc, mmul(G.C, pl));
  bool inface = (k == 0 || fabsf(l.x) <= G.ht.x + margin) && (k == 1 || fabsf(l.y) <= G.ht.y + margin) &&
                (k == 2 || fabsf(l.z) <= G.ht.z + margin);
  *depth = d;
  return inface;
}

// EDGE-EDGE contact of the pair (owner a, target t), in t's frame (oracle: edge_contact, operation for operation).  The corner-vs-face
// test above cannot see two boxes that cross edge over edge: no corner of either lies over a face of the other until they have sunk
// centimetres into each other (84 % of the overlaps deeper than 2 mm in a settled heap were of this kind).  This is the rest of the
// separating-axis test: the owner's three face axes and the nine edge-pair axes t_r x a_c.  When every axis overlaps by more than
// -m and the axis of LEAST overlap is an edge pair -- by more than `pref` over every face axis -- one contact is generated at the
// closest points of the two edges, normal = that axis (pointing from t to a), depth = the overlap along it.
#define EDGE_POINT 12                  /* its bit in the pair mask / its "sample point" number in the warm-start key */
#define EDGE_BIT (1u << 27)            /* contact word: the normal is explicit (CN), t1 = axis k of the target shape */
struct EdgeGeom { v3 p, n; float depth; int r; };
__device__ __forceinline__ float sel3(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }
// A pair of boxes with one axis of each (nearly) parallel has no edge-edge contact: with t_r || a_c every edge-pair axis t_r' x a_c'
// coincides with a face axis of one of the boxes (or vanishes), so the face axes already hold the least overlap -- every brick that
// lies flat on the slab or on another flat brick.  Same threshold as the per-axis test below (1 - cc^2 < 1e-3: within 2 degrees).
__device__ __forceinline__ bool edge_axes_parallel(const float* C) {
  float mx = C[0] * C[0];
#pragma unroll
  for (int i = 1; i < 9; ++i) { const float c2 = C[i] * C[i]; mx = c2 > mx ? c2 : mx; }
  return 1.0f - mx < 1e-3f;
}
// The SAT part, fully unrolled so that C / |C| stay in registers.  An edge axis whose UN-normalised overlap already is no smaller than the
// least face overlap cannot be the axis of least overlap (the normalised overlap is ov / s with s <= 1): its square root and division are
// skipped -- the result is what the plain loop of the oracle computes.  Returns the edge pair (r, c) in *rc (r * 3 + c) or -1.
__device__ __forceinline__ bool edge_sat(const PairGeom& G, float m, float pref, int* rc_out, float* be_out) {
  const float* C = G.C;
  float A[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = fabsf(C[i]);
  const float hav[3] = {G.ha.x, G.ha.y, G.ha.z}, htv[3] = {G.ht.x, G.ht.y, G.ht.z}, lcv[3] = {G.lc.x, G.lc.y, G.lc.z};
  float of = htv[0] + (A[0] * hav[0] + A[1] * hav[1] + A[2] * hav[2]) - fabsf(lcv[0]);
  { const float o1 = htv[1] + (A[3] * hav[0] + A[4] * hav[1] + A[5] * hav[2]) - fabsf(lcv[1]); if (o1 < of) of = o1; }
  { const float o2 = htv[2] + (A[6] * hav[0] + A[7] * hav[1] + A[8] * hav[2]) - fabsf(lcv[2]); if (o2 < of) of = o2; }
#pragma unroll
  for (int c = 0; c < 3; ++c) {                                      // the owner's face axes
    const float la = lcv[0] * C[c] + lcv[1] * C[3 + c] + lcv[2] * C[6 + c];
    const float oa = hav[c] + (A[c] * htv[0] + A[3 + c] * htv[1] + A[6 + c] * htv[2]) - fabsf(la);
    if (oa < -m) return false;
    if (oa < of) of = oa;
  }
  float be = 1e30f; int brc = -1;
#pragma unroll
  for (int rc = 0; rc < 9; ++rc) {                                   // edge pairs: axis t_r x a_c
    const int r = rc / 3, c = rc - 3 * r;
    const float cc = C[3 * r + c], s2 = 1.0f - cc * cc;
    if (s2 < 1e-3f) continue;                                        // edges within 2 degrees of parallel: the face axes cover it
    const int r1 = r == 2 ? 0 : r + 1, r2 = r == 0 ? 2 : r - 1, c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
    const float ra = hav[c1] * A[3 * r + c2] + hav[c2] * A[3 * r + c1];
    const float rb = htv[r1] * A[3 * r2 + c] + htv[r2] * A[3 * r1 + c];
    const float dist = fabsf(lcv[r2] * C[3 * r1 + c] - lcv[r1] * C[3 * r2 + c]);
    const float raw = ra + rb - dist;
    if (raw >= 0.0f && raw >= of - pref) continue;                   // ov = raw / s >= raw: not below the face axes, and not separating
    const float ov = raw / sqrtf(s2);
    if (ov < -m) return false;
    if (ov < be) { be = ov; brc = rc; }
  }
  if (brc < 0 || !(be < of - pref)) return false;                    // a face axis is (about) the axis of least overlap: the corner-face contacts have it
  *rc_out = brc; *be_out = be;
  return true;
}
// the contact of the edge pair (r, c) pass 1 found (it travels in the pair mask): closest points of the two edges, normal, and the
// depth = the overlap along that axis, evaluated by the expression edge_sat (and the oracle) used when it chose the pair
__device__ __forceinline__ void edge_point(const PairGeom& G, int rc, EdgeGeom& E) {
  // r, c are run-time values: every indexed access is a select over registers (an array indexed at run time would live in local memory)
  const float* C = G.C;
  const int r = rc / 3, c = rc - 3 * r;
  const int r1 = r == 2 ? 0 : r + 1, r2 = r == 0 ? 2 : r - 1, c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
  const v3 ha = G.ha, ht = G.ht, lc = G.lc;
  const v3 Cr = r == 0 ? V3(C[0], C[1], C[2]) : (r == 1 ? V3(C[3], C[4], C[5]) : V3(C[6], C[7], C[8]));        // row r
  const v3 Cr1 = r1 == 0 ? V3(C[0], C[1], C[2]) : (r1 == 1 ? V3(C[3], C[4], C[5]) : V3(C[6], C[7], C[8]));
  const v3 Cr2 = r2 == 0 ? V3(C[0], C[1], C[2]) : (r2 == 1 ? V3(C[3], C[4], C[5]) : V3(C[6], C[7], C[8]));
  const float cc = sel3(Cr.x, Cr.y, Cr.z, c), s2 = 1.0f - cc * cc, s = sqrtf(s2);
  float be;
  {
    const float ra = sel3(ha.x, ha.y, ha.z, c1) * fabsf(sel3(Cr.x, Cr.y, Cr.z, c2)) + sel3(ha.x, ha.y, ha.z, c2) * fabsf(sel3(Cr.x, Cr.y, Cr.z, c1));
    const float rb = sel3(ht.x, ht.y, ht.z, r1) * fabsf(sel3(Cr2.x, Cr2.y, Cr2.z, c)) + sel3(ht.x, ht.y, ht.z, r2) * fabsf(sel3(Cr1.x, Cr1.y, Cr1.z, c));
    const float dist = fabsf(sel3(lc.x, lc.y, lc.z, r2) * sel3(Cr1.x, Cr1.y, Cr1.z, c) - sel3(lc.x, lc.y, lc.z, r1) * sel3(Cr2.x, Cr2.y, Cr2.z, c));
    be = (ra + rb - dist) / s;
  }
  const v3 ac = V3(sel3(C[0], C[1], C[2], c), sel3(C[3], C[4], C[5], c), sel3(C[6], C[7], C[8], c));   // the owner's axis c in t's frame
  v3 n = r == 0 ? V3(0.0f, -ac.z, ac.y) : (r == 1 ? V3(ac.z, 0.0f, -ac.x) : V3(-ac.y, ac.x, 0.0f));   // e_r x ac
  n = V3(n.x / s, n.y / s, n.z / s);
  if (n.x * lc.x + n.y * lc.y + n.z * lc.z < 0.0f) n = vneg(n);                                         // from t towards a
  v3 pt = V3(r == 0 ? 0.0f : (n.x >= 0.0f ? ht.x : -ht.x), r == 1 ? 0.0f : (n.y >= 0.0f ? ht.y : -ht.y),
             r == 2 ? 0.0f : (n.z >= 0.0f ? ht.z : -ht.z));                                             // t's edge: its support towards a
  v3 pa = lc;
#pragma unroll
  for (int k = 0; k < 3; ++k) {                                                                        // a's edge: its support towards t
    if (k == c) continue;
    const float nk = n.x * C[k] + n.y * C[3 + k] + n.z * C[6 + k];
    const float hk = k == 0 ? ha.x : (k == 1 ? ha.y : ha.z);
    const float co = nk >= 0.0f ? -hk : hk;
    pa.x = pa.x + co * C[k]; pa.y = pa.y + co * C[3 + k]; pa.z = pa.z + co * C[6 + k];
  }
  const v3 d0 = V3(pa.x - pt.x, pa.y - pt.y, pa.z - pt.z);
  const float de = sel3(d0.x, d0.y, d0.z, r), da = d0.x * ac.x + d0.y * ac.y + d0.z * ac.z;
  float u = (de - cc * da) / s2, v = (cc * de - da) / s2;            // closest points of the two edge LINES, clamped to the edges
  const float htr = sel3(ht.x, ht.y, ht.z, r), hac = sel3(ha.x, ha.y, ha.z, c);
  u = clampf(u, -htr, htr); v = clampf(v, -hac, hac);
  const v3 qt = V3(r == 0 ? u : pt.x, r == 1 ? u : pt.y, r == 2 ? u : pt.z);
  const v3 qa = V3(pa.x + v * ac.x, pa.y + v * ac.y, pa.z + v * ac.z);
  E.p = V3(0.5f * (qt.x + qa.x), 0.5f * (qt.y + qa.y), 0.5f * (qt.z + qa.z));
  E.n = n; E.depth = be; E.r = r;
}

// contacts of a pair: the sample points that hit (bits 0-11) + its edge-edge contact (bits 12-15 non-zero)
__device__ __forceinline__ int mask_count(unsigned mk) { return __popc(mk & 0xfffu) + ((mk >> EDGE_POINT) ? 1 : 0); }
__device__ __forceinline__ void touch_or(unsigned char* flags, int i, unsigned bits) {
  atomicOr(reinterpret_cast<unsigned*>(flags) + (i >> 2), bits << (8 * (i & 3)));
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}

// CMP: the scene has COMPOUND free bodies (a body = several boxes, sdx_scene_t::n_bshapes > 0).  The one-box-per-body instantiation
// (every BlockAssembly scene) is the code it always was: box index == body index.
// EDGE: edge-edge contacts (sdx_scene_t::edge_contacts > 0.5); the instantiation without them carries none of their code
// (code size is a performance number in this kernel).
template <bool CMP, bool EDGE>
__global__ void __launch_bounds__(SIM_THREADS, SIM_MIN_CTAS)
k_simulate(const sdx_scene_t* __restrict__ S, float* __restrict__ brick, float* __restrict__ dof,
           float* __restrict__ link_out, float* __restrict__ jac7, float* __restrict__ netf,
           int* __restrict__ ncontact, float* __restrict__ condump, float* ws, int* wsn, int ws_cur,
           unsigned char* __restrict__ slp, int n_envs, float4* cscratch /* [n_envs][4][MAXC]: contact records (SIM_GLOBAL_CONTACTS), impulses (SIM_GLOBAL_CF), explicit normals of edge contacts */) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SimSmem& M = *reinterpret_cast<SimSmem*>(smem_raw);
  const int e = blockIdx.x, tid = threadIdx.x;
  if (e >= n_envs) return;
  const int nbr = S->n_bricks, nrs = S->n_rshapes, nst = S->n_static;
  const int nbs = CMP ? S->n_bshapes : nbr;          // collision boxes of the free bodies (CMP: box a rides on body bs_body[a])
  const int n_owner = NB + nrs;
#if !SIM_BROAD_SYM
  const int n_target = NB + nrs + nst;
#endif
  const int substeps = S->substeps, iters = S->iters;
  const float h = S->dt / (float)substeps;
  const float margin = S->contact_offset;
  const float fmargin = S->face_margin;          // in-face tolerance (NOT the speculative range)
  int shed_max = 0;                              // block-uniform: most shedding any sub-step needed
  float* gbrick = brick + (size_t)e * 13 * NB;
  float* gdof = dof + (size_t)e * 72;
#if SIM_GLOBAL_CONTACTS
  float4* const CA = cscratch + (size_t)e * 4 * MAXC;   // plain (coherent, L1-cached) loads / stores: written and read by this CTA only
  float4* const CB = CA + MAXC;
#else
  float4* const CA = M.ca;
  float4* const CB = M.cb;
#endif
#if SIM_GLOBAL_CF
  float4* const CF = cscratch + (size_t)e * 4 * MAXC + 2 * MAXC;
#else
  float4* const CF = M.cf4;
#endif
  float4* const CN = cscratch + (size_t)e * 4 * MAXC + 3 * MAXC;         // explicit normals (edge-edge contacts only)
  const float epref = S->edge_pref;
  unsigned char* const cf_bytes = reinterpret_cast<unsigned char*>(CF);   // 16 KB of scratch while the impulses are not live

  // ---- TMA bulk load of the env's brick tile into shared memory
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&M.mbar);
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(M.tile);
  if (tid == 0) {
    M.nelist = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(13 * NB * 4) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tile_s), "l"(gbrick), "r"(13 * NB * 4), "r"(bar) : "memory");
  }
  // robot state + static tables while the tile is in flight
  if (tid >= ROBOT_TID0 && tid < ROBOT_TID0 + SDX_ND) {
    int j = tid - ROBOT_TID0;
    M.q[j] = gdof[j]; M.qd[j] = gdof[24 + j]; M.tgt[j] = gdof[48 + j];
  }
  for (int s = tid; s < nst; s += SIM_THREADS) {
    int t = NB + nrs + s;
    v3 stc = V3(S->st_c[3 * s], S->st_c[3 * s + 1], S->st_c[3 * s + 2]);
    if (S->st_mod[s] > 0 && e % S->st_mod[s] != S->st_rem[s]) stc.z = -1000.0f;   // not part of this env's scene (InsertSim's base-plate by env % 3, IS:971-977)
    st3(M.sc[t], stc);
    st3(M.sh[t], V3(S->st_h[3 * s], S->st_h[3 * s + 1], S->st_h[3 * s + 2]));
#pragma unroll
    for (int i = 0; i < 9; ++i) M.sR[t][i] = (i % 4 == 0) ? 1.0f : 0.0f;
    M.sbody[t] = STATIC_BODY;
  }
  mbar_wait(&M.mbar, 0);

  // brick threads keep their body in registers across the step
  v3 bxr = V3(0, 0, 0), bvr = V3(0, 0, 0), bwr = V3(0, 0, 0);
  q4 bqr = Q4(0, 0, 0, 1);
  v3 halfb = V3(0, 0, 0);
  float brad = 0.0f;                                             // bounding radius of this thread's body about its COM
  bool built_asleep = false;                                     // was this brick asleep when the candidate lists were built?
  int slpc = (tid < NB) ? (int)slp[(size_t)e * NB + tid] : 0;   // sub-steps since this brick was last hot (oracle: sim_env SLEEPING)
  const int sleep_n = S->sleep_substeps;
  const unsigned my_anc = (tid >= ROBOT_TID0 && tid < ROBOT_TID0 + SDX_NL) ? S->link_anc_mask[tid - ROBOT_TID0] : 0u;
  unsigned my_desc = 0u;   // DoF threads: links moved by this DoF
  if (tid >= ROBOT_TID0 && tid < ROBOT_TID0 + SDX_ND)
    for (int L = 0; L < SDX_NL; ++L) if (S->link_anc_mask[L] & (1u << (tid - ROBOT_TID0))) my_desc |= 1u << L;
  if (tid < NB) {
    bxr = V3(M.tile[0 * NB + tid], M.tile[1 * NB + tid], M.tile[2 * NB + tid]);
    bqr = Q4(M.tile[3 * NB + tid], M.tile[4 * NB + tid], M.tile[5 * NB + tid], M.tile[6 * NB + tid]);
    bvr = V3(M.tile[7 * NB + tid], M.tile[8 * NB + tid], M.tile[9 * NB + tid]);
    bwr = V3(M.tile[10 * NB + tid], M.tile[11 * NB + tid], M.tile[12 * NB + tid]);
    halfb = V3(S->br_half[3 * tid], S->br_half[3 * tid + 1], S->br_half[3 * tid + 2]);
    st3(M.sh[tid], halfb);
    M.srad[tid] = sqrtf(vdot(halfb, halfb));
    M.sbody[tid] = (unsigned char)(CMP && tid < nbs ? S->bs_body[tid] : tid);
    brad = M.srad[tid];
    if (CMP) {                                                   // bounding radius of body tid: max over its boxes of |centre| + half diagonal
      brad = 0.0f;
      for (int a = 0; a < nbs; ++a)
        if (S->bs_body[a] == tid) {
          const v3 c = V3(S->bs_c[3 * a], S->bs_c[3 * a + 1], S->bs_c[3 * a + 2]), hh = V3(S->br_half[3 * a], S->br_half[3 * a + 1], S->br_half[3 * a + 2]);
          const float cand = sqrtf(vdot(c, c)) + sqrtf(vdot(hh, hh));
          if (cand > brad) brad = cand;
        }
    }
  }
  if (tid < nrs) {
    int t = NB + tid;
    v3 hh = V3(S->rs_h[3 * tid], S->rs_h[3 * tid + 1], S->rs_h[3 * tid + 2]);
    st3(M.sh[t], hh);
    M.srad[t] = sqrtf(vdot(hh, hh));
    M.sbody[t] = (unsigned char)(NB + S->rs_body[tid]);
  }
  __syncthreads();

  float* gws = ws + (size_t)e * 2 * MAXC * 4;
  int* gwsn = wsn + 2 * e;
  const float warm = S->warm_start;
  int rb = ws_cur;
  for (int sub = 0; sub < substeps; ++sub) {
    const float* wsr = gws + (size_t)rb * MAXC * 4;        // impulse cache of the previous sub-step / step
    float* wsw = gws + (size_t)(1 - rb) * MAXC * 4;
    const int nprev = gwsn[rb];
#ifdef SIM_PROFILE
    // per pass and warp: clock64 when the warp ARRIVES at the barrier after phase A and at the one after phase B (arrival times
    // are exact; a read placed after a BAR.SYNC.DEFER_BLOCKING is not); PMARK(k): thread 0's clock after the barrier that ends
    // stage k of the sub-step.  Sink: behind the contact rows of the debug dump (tools/sim_phase_cycles.py).
    long long* const prof = condump ? reinterpret_cast<long long*>(condump + ((size_t)n_envs * MAXC) * 8) + ((size_t)e * 2 + sub) * SIM_PROF_REC : nullptr;
#define PMARK(k) do { if (prof && tid == 0) prof[18 * 2 * 8 + 4 + (k)] = clock64(); } while (0)
#else
#define PMARK(k) do { } while (0)
#endif
    PMARK(0);
#if SIM_KIN_FUSED
    // 1.-4. kinematics, poses, free velocities, world AABBs behind ONE block barrier (the vote below; compound scenes: two).  Everything
    //    about the articulation is the robot warp's own business -- FK, then a lane per DoF (implicit PD) and per robot box (pose), then a
    //    lane per link (twist), then a lane per box again (AABB + travel bound), __syncwarp between -- and a brick thread needs nothing but
    //    its own brick: it used to wait at three block barriers for the robot's chain.  Same arithmetic, stage for stage.
    auto shape_aabb = [&](int t) {
      const float* R = M.sR[t];
      v3 hh = ld3(M.sh[t]);
      const v3 ext = V3(fabsf(R[0]) * hh.x + fabsf(R[1]) * hh.y + fabsf(R[2]) * hh.z,
                        fabsf(R[3]) * hh.x + fabsf(R[4]) * hh.y + fabsf(R[5]) * hh.z,
                        fabsf(R[6]) * hh.x + fabsf(R[7]) * hh.y + fabsf(R[8]) * hh.z);
      int bd = M.sbody[t];
      v3 dc = vsub(ld3(M.sc[t]), ld3(M.bx[bd]));
      float reach = sqrtf(vdot(dc, dc)) + M.srad[t];
      v3 bv = ld3(M.bv[bd]), bw = ld3(M.bw[bd]);
      M.sab[t] = make_float4(ext.x, ext.y, ext.z, h * (sqrtf(vdot(bv, bv)) + sqrtf(vdot(bw, bw)) * reach));
    };
    bool asleep = false;
    if (tid >= ROBOT_TID0) {
      const int lane = tid - ROBOT_TID0;
      robot_fk(S, M, lane);                                      // ends in __syncwarp
      if (CMP) __syncthreads();                                  // (pairs with the brick threads' barrier below)
      if (lane < nrs) {                                          // robot box poses
        int t = NB + lane, L = S->rs_body[lane];
        q4 qL = Q4(M.lq[L][0], M.lq[L][1], M.lq[L][2], M.lq[L][3]);
        q4 ql = Q4(S->rs_quat[4 * lane], S->rs_quat[4 * lane + 1], S->rs_quat[4 * lane + 2], S->rs_quat[4 * lane + 3]);
        st3(M.sc[t], vadd(ld3(M.bx[NB + L]), qrot(qL, V3(S->rs_c[3 * lane], S->rs_c[3 * lane + 1], S->rs_c[3 * lane + 2]))));
        qmat(qmul(qL, ql), M.sR[t]);
      }
      if (lane < SDX_ND) {                                       // implicit PD free joint velocities
        int j = lane;
        float I = S->dof_inertia[j], kp = S->dof_kp[j], kd = S->dof_kd[j];
        float ieff = I + h * kd + (h * h) * kp;
        float qdo = M.qd[j];
        float qn = (I * qdo + h * kp * (M.tgt[j] - M.q[j])) / ieff;
        float tau = I * (qn - qdo) / h;
        if (tau > S->dof_effort[j]) qn = qdo + S->dof_effort[j] * h / I;
        if (tau < -S->dof_effort[j]) qn = qdo - S->dof_effort[j] * h / I;
        qn = clampf(qn, -S->dof_vmax[j], S->dof_vmax[j]);
        M.ieff[j] = ieff; M.qdfree[j] = qn; M.qd[j] = qn;
      }
      __syncwarp();
      if (lane < SDX_NL) link_twist(S, M, lane, my_anc);
      __syncwarp();
      if (lane < nrs) shape_aabb(NB + lane);
    } else {
      if (tid < NB) {
        asleep = tid < nbr && sleep_n > 0 && slpc >= sleep_n;
        M.sflag[tid] = (unsigned char)((asleep ? 1 : 0) | ((tid < nbr && slpc == 0) ? 2 : 0));
        M.touch[tid] = 0;
        const float invm = asleep ? 0.0f : S->br_invm[tid];    // a sleeping brick is immovable for this sub-step
        const v3 invI = asleep ? V3(0.0f, 0.0f, 0.0f) : V3(S->br_invI[3 * tid], S->br_invI[3 * tid + 1], S->br_invI[3 * tid + 2]);
        float Rb[9];
        qmat(bqr, Rb);
        if (!CMP) {
#pragma unroll
          for (int i = 0; i < 9; ++i) M.sR[tid][i] = Rb[i];
          st3(M.sc[tid], bxr);
        } else M.bq[tid] = make_float4(bqr.x, bqr.y, bqr.z, bqr.w);
        st3(M.bx[tid], bxr);
        float damp = 1.0f - h * S->brick_ang_damp;
        float ldamp = 1.0f - h * S->brick_lin_damp;
        v3 vfree = vscale(V3(bvr.x, bvr.y, bvr.z + h * S->gravity_z), ldamp);
        v3 wfree = vscale(bwr, damp);
        if (tid >= nbr || asleep) { vfree = V3(0.0f, 0.0f, 0.0f); wfree = V3(0.0f, 0.0f, 0.0f); }
        st3(M.bv[tid], vfree); st3(M.bw[tid], wfree);
        M.bfv[tid] = make_float4(vfree.x, vfree.y, vfree.z, invm);
        M.bfw[tid] = make_float4(wfree.x, wfree.y, wfree.z, 0.0f);
        brick_world_invI(Rb, invI, &M.bI0[tid], &M.bI1[tid]);
      }
      if (CMP) {
        __syncthreads();                                         // the boxes of a compound body read their body's frame
        if (tid < NB) {
          const int b = M.sbody[tid];
          const float4 q4b = M.bq[b];
          qmat(Q4(q4b.x, q4b.y, q4b.z, q4b.w), M.sR[tid]);
          const v3 xb = ld3(M.bx[b]);
          st3(M.sc[tid], tid < nbs ? vadd(xb, mmul(M.sR[tid], V3(S->bs_c[3 * tid], S->bs_c[3 * tid + 1], S->bs_c[3 * tid + 2]))) : xb);
        }
      }
      if (tid < NB) shape_aabb(tid);
      for (int s2 = tid; s2 < nst; s2 += ROBOT_TID0) {           // statics: AABB = the box itself, no travel (rewritten each sub-step: aliased storage)
        int t = NB + nrs + s2;
        const v3 hs = ld3(M.sh[t]);
        M.sab[t] = make_float4(hs.x, hs.y, hs.z, 0.0f);
      }
    }
    PMARK(1);
#else
    // 1. kinematics (one thread walks the chain) || brick poses + free velocities
    if (tid >= ROBOT_TID0) robot_fk(S, M, tid - ROBOT_TID0);
    bool asleep = false;
    if (tid < NB) {
      asleep = tid < nbr && sleep_n > 0 && slpc >= sleep_n;
      M.sflag[tid] = (unsigned char)((asleep ? 1 : 0) | ((tid < nbr && slpc == 0) ? 2 : 0));
      M.touch[tid] = 0;
      const float invm = asleep ? 0.0f : S->br_invm[tid];    // a sleeping brick is immovable for this sub-step
      const v3 invI = asleep ? V3(0.0f, 0.0f, 0.0f) : V3(S->br_invI[3 * tid], S->br_invI[3 * tid + 1], S->br_invI[3 * tid + 2]);
      float Rb[9];
      qmat(bqr, Rb);
      if (!CMP) {
#pragma unroll
        for (int i = 0; i < 9; ++i) M.sR[tid][i] = Rb[i];
        st3(M.sc[tid], bxr);
      } else M.bq[tid] = make_float4(bqr.x, bqr.y, bqr.z, bqr.w);
      st3(M.bx[tid], bxr);
      float damp = 1.0f - h * S->brick_ang_damp;
      float ldamp = 1.0f - h * S->brick_lin_damp;
      v3 vfree = vscale(V3(bvr.x, bvr.y, bvr.z + h * S->gravity_z), ldamp);
      v3 wfree = vscale(bwr, damp);
      if (tid >= nbr || asleep) { vfree = V3(0.0f, 0.0f, 0.0f); wfree = V3(0.0f, 0.0f, 0.0f); }
      st3(M.bv[tid], vfree); st3(M.bw[tid], wfree);
      M.bfv[tid] = make_float4(vfree.x, vfree.y, vfree.z, invm);
      M.bfw[tid] = make_float4(wfree.x, wfree.y, wfree.z, 0.0f);
      brick_world_invI(Rb, invI, &M.bI0[tid], &M.bI1[tid]);
    }
    __syncthreads();
    // 2. robot shape poses || implicit PD free joint velocities (CMP: || the boxes of the compound bodies)
    if (CMP && tid < NB) {
      const int b = M.sbody[tid];
      const float4 q4b = M.bq[b];
      qmat(Q4(q4b.x, q4b.y, q4b.z, q4b.w), M.sR[tid]);
      const v3 xb = ld3(M.bx[b]);
      st3(M.sc[tid], tid < nbs ? vadd(xb, mmul(M.sR[tid], V3(S->bs_c[3 * tid], S->bs_c[3 * tid + 1], S->bs_c[3 * tid + 2]))) : xb);
    }
    if (tid < nrs) {
      int t = NB + tid, L = S->rs_body[tid];
      q4 qL = Q4(M.lq[L][0], M.lq[L][1], M.lq[L][2], M.lq[L][3]);
      q4 ql = Q4(S->rs_quat[4 * tid], S->rs_quat[4 * tid + 1], S->rs_quat[4 * tid + 2], S->rs_quat[4 * tid + 3]);
      st3(M.sc[t], vadd(ld3(M.bx[NB + L]), qrot(qL, V3(S->rs_c[3 * tid], S->rs_c[3 * tid + 1], S->rs_c[3 * tid + 2]))));
      qmat(qmul(qL, ql), M.sR[t]);
    }
    if (tid >= ROBOT_TID0 && tid < ROBOT_TID0 + SDX_ND) {
      int j = tid - ROBOT_TID0;
      float I = S->dof_inertia[j], kp = S->dof_kp[j], kd = S->dof_kd[j];
      float ieff = I + h * kd + (h * h) * kp;
      float qdo = M.qd[j];
      float qn = (I * qdo + h * kp * (M.tgt[j] - M.q[j])) / ieff;
      float tau = I * (qn - qdo) / h;
      if (tau > S->dof_effort[j]) qn = qdo + S->dof_effort[j] * h / I;
      if (tau < -S->dof_effort[j]) qn = qdo - S->dof_effort[j] * h / I;
      qn = clampf(qn, -S->dof_vmax[j], S->dof_vmax[j]);
      M.ieff[j] = ieff; M.qdfree[j] = qn; M.qd[j] = qn;
    }
    __syncthreads();
    // 3. link twists
    if (tid >= ROBOT_TID0 && tid < ROBOT_TID0 + SDX_NL) link_twist(S, M, tid - ROBOT_TID0, my_anc);
    __syncthreads();
    PMARK(1);
    // 4. world AABBs + per-sub-step travel bound of every moving shape
    if (tid < n_owner) {
      int t = tid;
      const float* R = M.sR[t];
      v3 hh = ld3(M.sh[t]);
      const v3 ext = V3(fabsf(R[0]) * hh.x + fabsf(R[1]) * hh.y + fabsf(R[2]) * hh.z,
                        fabsf(R[3]) * hh.x + fabsf(R[4]) * hh.y + fabsf(R[5]) * hh.z,
                        fabsf(R[6]) * hh.x + fabsf(R[7]) * hh.y + fabsf(R[8]) * hh.z);
      int bd = M.sbody[t];
      v3 dc = vsub(ld3(M.sc[t]), ld3(M.bx[bd]));
      float reach = sqrtf(vdot(dc, dc)) + M.srad[t];
      v3 bv = ld3(M.bv[bd]), bw = ld3(M.bw[bd]);
      M.sab[t] = make_float4(ext.x, ext.y, ext.z, h * (sqrtf(vdot(bv, bv)) + sqrtf(vdot(bw, bw)) * reach));
    }
    for (int s2 = tid; s2 < nst; s2 += SIM_THREADS) {          // statics: AABB = the box itself, no travel (rewritten each sub-step: aliased storage)
      int t = NB + nrs + s2;
      const v3 hs = ld3(M.sh[t]);
      M.sab[t] = make_float4(hs.x, hs.y, hs.z, 0.0f);
    }
#endif
    // The candidate lists are built in the FIRST sub-step of a step for ALL its sub-steps (travel bounds scaled by the number of
    // sub-steps left, plus the speed gravity adds in between) and rebuilt later only if a brick that was asleep when they were
    // built has been woken since: its pairs with sleeping bricks and statics were filtered (oracle: sim_env 3.)
#if SIM_BROAD_SYM
    for (int w = tid; w < NOWN * 4; w += SIM_THREADS) reinterpret_cast<unsigned*>(cf_bytes)[w] = 0u;   // hit rows of the broad phase (the impulses that lived here are in the cache by now)
#endif
    const int rebuild = __syncthreads_or(sub == 0 || (tid < NB && built_asleep && !asleep));
    PMARK(2);
    // 5. broad phase: TWO threads per owner shape (thread tid: owner tid & 127, target-range half tid >> 7).  Dynamic targets
    //    (bricks, robot boxes) and statics are listed separately; at the merge the STATICS CLAIM THEIR SLOTS FIRST -- a brick must
    //    never lose its floor / wall pair to a crowd of neighbours -- and the dynamic targets fill what is left in ascending order.
    //    The list stays ascending (dynamic part, then statics) and equals what the oracle's single sweep produces.
    if (rebuild) {
      const int left = substeps - sub;
      const float infl = (float)left, slack = (float)(left - 1) * ((h * h) * fabsf(S->gravity_z));
      built_asleep = asleep;
#if SIM_BROAD_SYM
      // The box test is symmetric in (owner, target) -- |c_a - c_t| <= e_a + e_t + m with m built from sums of both shapes' travel bounds,
      // bit for bit the same either way round -- so every unordered pair of moving shapes is tested ONCE and the hit is entered in both
      // shapes' 128-bit rows (bit t of row a, bit a of row t; atomicOr: the result does not depend on the order).  Pairing: owner a tests
      // t = a + d (mod n_owner) for d = 1 .. (n_owner - 1) / 2 (+ d = n_owner / 2 for the lower half of the owners when n_owner is even): each
      // pair exactly once, the same number of tests for every owner; two threads per owner split the d range and the statics.  6.8 k tests
      // per rebuild instead of 11.3 k; the lists read off the rows in ascending order are the ones the one-sided sweep built.
      unsigned* rows = reinterpret_cast<unsigned*>(cf_bytes);                // [NOWN][4] hit masks over the target index space (zeroed before the barrier above)
      if (tid == 0) { M.ndrop_cand = 0; M.ndrop_static = 0; }
      const int a = tid & 127;
      for (int half = tid >> 7; half < 2; half += SIM_THREADS >> 7)   // 256 threads: one (owner, half) each; 128 threads: both halves in turn
      if (a < n_owner && !(a < NB && a >= nbs)) {
        const v3 ca = ld3(M.sc[a]);
        const float4 A4 = M.sab[a];
        const int sba = M.sbody[a];
        const bool a_sl = a < NB && (M.sflag[sba] & 1);
        unsigned own[4] = {0u, 0u, 0u, 0u};
        auto hit = [&](int t) {
          const v3 d = vsub(ca, ld3(M.sc[t]));
          const float4 T4 = M.sab[t];
          const float m = margin + infl * (A4.w + T4.w) + slack;
          return fabsf(d.x) <= A4.x + T4.x + m && fabsf(d.y) <= A4.y + T4.y + m && fabsf(d.z) <= A4.z + T4.z + m;
        };
        auto mark = [&](int t) {
          if (t < 32) own[0] |= 1u << t; else if (t < 64) own[1] |= 1u << (t - 32); else if (t < 96) own[2] |= 1u << (t - 64); else own[3] |= 1u << (t - 96);
        };
        const int n = n_owner;
        const int dmax = ((n - 1) >> 1) + (((n & 1) == 0 && a < (n >> 1)) ? 1 : 0);
        const int dsplit = (dmax + 1) >> 1;
        const int dlo = half ? dsplit + 1 : 1, dhi = half ? dmax : dsplit;
SIM_BROAD_UNROLL
        for (int d = dlo; d <= dhi; ++d) {
          int t = a + d;
          if (t >= n) t -= n;
          if (t < NB && t >= nbs) continue;                                  // unused box slot
          if (a >= NB && t >= NB) continue;                                  // robot-robot pairs are filtered (GS:906)
          if (CMP && (int)M.sbody[t] == sba) continue;                       // boxes of one body
          if (a_sl && t < NB && (M.sflag[CMP ? (int)M.sbody[t] : t] & 1)) continue;   // two sleeping bricks: neither box can move
          if (hit(t)) { mark(t); atomicOr(&rows[4 * t + (a >> 5)], 1u << (a & 31)); }
        }
        if (!a_sl) {                                                         // a sleeping brick against a static: neither box can move
          const int smid = nst >> 1;
SIM_BROAD_UNROLL
          for (int s2 = half ? smid : 0; s2 < (half ? nst : smid); ++s2) { const int t = NB + nrs + s2; if (hit(t)) mark(t); }
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) if (own[w]) atomicOr(&rows[4 * a + w], own[w]);
      }
      __syncthreads();
      if (tid < n_owner) {
        // row -> candidate list: dynamic targets ascending, capped so that the statics (which claim their slots first) all fit
        unsigned r[4] = {rows[4 * tid], rows[4 * tid + 1], rows[4 * tid + 2], rows[4 * tid + 3]};
        const int s0 = NB + nrs;                                             // first static target: bits below are moving shapes
        int nd_all = 0, ns_all = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const unsigned dynm = s0 >= 32 * (w + 1) ? 0xffffffffu : (s0 <= 32 * w ? 0u : (1u << (s0 - 32 * w)) - 1u);
          nd_all += __popc(r[w] & dynm); ns_all += __popc(r[w] & ~dynm);
        }
        const int ns = min(ns_all, KSTAT), kd = min(nd_all, KC - ns);
        int k = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const unsigned dynm = s0 >= 32 * (w + 1) ? 0xffffffffu : (s0 <= 32 * w ? 0u : (1u << (s0 - 32 * w)) - 1u);
          unsigned m = r[w] & dynm;
          while (m && k < kd) { const int bpos = __ffs(m) - 1; m &= m - 1; M.cand[tid][k++] = (unsigned char)(32 * w + bpos); }
        }
        k = kd;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const unsigned dynm = s0 >= 32 * (w + 1) ? 0xffffffffu : (s0 <= 32 * w ? 0u : (1u << (s0 - 32 * w)) - 1u);
          unsigned m = r[w] & ~dynm;
          while (m && k < kd + ns) { const int bpos = __ffs(m) - 1; m &= m - 1; M.cand[tid][k++] = (unsigned char)(32 * w + bpos); }
        }
        M.ncand[tid] = kd + ns;
        const int dropped = (nd_all - kd) + (ns_all - ns);
        if (dropped) atomicAdd(&M.ndrop_cand, dropped);
        if (ns_all > ns) atomicAdd(&M.ndrop_static, ns_all - ns);
      }
#else
      unsigned char* tmpc = cf_bytes;                                       // [NOWN][KC]    dynamic hits of the second half
      int* tmpn = reinterpret_cast<int*>(cf_bytes + NOWN * KC);             // [NOWN][4]     all dynamic hits of half 0 | half 1 | statics kept | statics seen
      unsigned char* tmps = cf_bytes + NOWN * KC + NOWN * 16;               // [NOWN][KSTAT] static hits
      if (tid == 0) { M.ndrop_cand = 0; M.ndrop_static = 0; }
      const int a = tid & 127;
      for (int half = tid >> 7; half < 2; half += SIM_THREADS >> 7)   // 256 threads: one (owner, half) each; 128 threads: both halves in turn
      if (a < n_owner) {
        int k = 0, kall = 0, ks = 0, ksall = 0;
        if (!(a < NB && a >= nbs)) {
          const v3 ca = ld3(M.sc[a]);
          const float4 A4 = M.sab[a];
          const int sba = M.sbody[a];
          const bool a_sl = a < NB && (M.sflag[sba] & 1);
          const int tmid = n_target >> 1;                        // < NB + nrs: the statics all fall into the second half
          const int tlo = half ? tmid : 0, thi = half ? n_target : tmid;
          unsigned char* dst = half ? tmpc + a * KC : M.cand[a];
          unsigned char* sdst = tmps + a * KSTAT;
          auto hit = [&](int t) {
            const v3 d = vsub(ca, ld3(M.sc[t]));
            const float4 T4 = M.sab[t];
            const float m = margin + infl * (A4.w + T4.w) + slack;
            return fabsf(d.x) <= A4.x + T4.x + m && fabsf(d.y) <= A4.y + T4.y + m && fabsf(d.z) <= A4.z + T4.z + m;
          };
          // the target index space is bricks [0, nbr) | robot shapes [NB, NB+nrs) | statics [NB+nrs, n_target): one loop per
          // class with the class-level filters hoisted (same ascending order as one sweep over t)
SIM_BROAD_UNROLL
          for (int t = max(tlo, 0); t < min(thi, nbs); ++t) {
            if (CMP ? (int)M.sbody[t] == sba : t == a) continue;   // itself (CMP: any box of the same body)
            if (a_sl && (M.sflag[CMP ? (int)M.sbody[t] : t] & 1)) continue;              // neither box can move
            if (hit(t)) { if (k < KC) dst[k++] = (unsigned char)t; kall++; }
          }
          if (a < NB) {                                          // robot-robot pairs are filtered (GS:906)
SIM_BROAD_UNROLL
            for (int t = max(tlo, NB); t < min(thi, NB + nrs); ++t)
              if (hit(t)) { if (k < KC) dst[k++] = (unsigned char)t; kall++; }
          }
          if (!a_sl) {                                           // a sleeping brick against a static: neither box can move
SIM_BROAD_UNROLL
            for (int t = max(tlo, NB + nrs); t < thi; ++t)
              if (hit(t)) { if (ks < KSTAT) sdst[ks++] = (unsigned char)t; ksall++; }
          }
        }
        if (half) { tmpn[4 * a + 1] = kall; tmpn[4 * a + 2] = ks; tmpn[4 * a + 3] = ksall; } else tmpn[4 * a] = kall;
      }
      __syncthreads();
      if (tid < n_owner) {
        const int nd0_all = tmpn[4 * tid], nd1_all = tmpn[4 * tid + 1], ns = tmpn[4 * tid + 2], ns_all = tmpn[4 * tid + 3];
        const int nd_all = nd0_all + nd1_all;
        const int kd = min(nd_all, KC - ns);
        const int nd0 = min(nd0_all, KC);
        for (int i = nd0; i < kd; ++i) M.cand[tid][i] = tmpc[tid * KC + (i - nd0)];
        for (int i = 0; i < ns; ++i) M.cand[tid][kd + i] = tmps[tid * KSTAT + i];
        M.ncand[tid] = kd + ns;
        const int dropped = (nd_all - kd) + (ns_all - ns);
        if (dropped) atomicAdd(&M.ndrop_cand, dropped);
        if (ns_all > ns) atomicAdd(&M.ndrop_static, ns_all - ns);
      }
#endif
      __syncthreads();
      PMARK(3);
      if (tid < 32) {
        for (int a = tid; a < n_owner; a += 32) M.poff[a] = M.ncand[a];
        __syncwarp();
        int tot = warp_excl_scan(M.poff, n_owner, tid);
        if (tid == 0) M.poff[n_owner] = tot;
      }
      __syncthreads();
    } else {
      PMARK(3);
    }
    PMARK(4);
    // 6. narrow phase, pass 1 (pair-parallel): per-pair hit masks + running contact offsets.  The pair tables live in
    //    the (still unused) impulse / inverse-mass arrays.  Contiguous pair chunks per thread => offsets are in pair order.
    const int npairs = M.poff[n_owner];
    const int PP = (npairs + SIM_THREADS - 1) / SIM_THREADS;
    const int p0 = tid * PP, p1 = min(npairs, p0 + PP);
    unsigned short* pmask = reinterpret_cast<unsigned short*>(cf_bytes);          // [npairs] <= 3328 * 2 B < 8 KB
    unsigned short* pstart = reinterpret_cast<unsigned short*>(cf_bytes + 8192);  // [npairs]
    // SHEDDING: contacts with a positive gap are speculative.  If the table would overflow, the speculative range is halved
    // (twice) and then dropped altogether -- the contacts that go first are those that cannot act in this sub-step anyway --
    // before a touching contact is lost (oracle: sim_env 4.).  Block-uniform loop; level 0 in all but the most crowded sub-steps.
    int mycount, incl, total, level = 0;
    float gs = 1.0f;
    unsigned short* const elist = reinterpret_cast<unsigned short*>(CN);           // EDGE: pairs queued for the edge-axis test (CN is not live before pass 2)
    for (;;) {
      mycount = 0;
      int a = 0;
      // EDGE: pairs are dealt to the threads round-robin (a warp's range of the pair list is one owner class -- bricks with many hits, robot
      // boxes with none -- so contiguous chunks leave the warps unevenly loaded; the barrier this needs is there anyway for pass 1b: 5.11 ->
      // 5.08 ms per launch); the contact COUNTS below are still taken per contiguous chunk.  Without edge contacts: contiguous chunks, no barrier.
      constexpr bool strided = EDGE && SIM_PASS1_STRIDED;
      for (int i = strided ? tid : p0; i < (strided ? npairs : p1); i += strided ? SIM_THREADS : 1) {
        while (M.poff[a + 1] <= i) ++a;
        int t = M.cand[a][i - M.poff[a]];
        float m = (margin + M.sab[a].w + M.sab[t].w) * gs;
        unsigned short mk = 0;
        PairGeom G;
        const bool dead = !rebuild && a < NB && (M.sflag[CMP ? (int)M.sbody[a] : a] & 1) && (t >= NB + nrs || (t < NB && (M.sflag[CMP ? (int)M.sbody[t] : t] & 1)));   // kept list, both asleep by now
        if (!dead && pair_geom(M, a, t, m, G, t >= NB + nrs)) {
          int npts = (a < NB && G.ha.x > 0.04f) ? 12 : 8;
          for (int p = 0; p < npts; ++p) { float d; if (point_hit(G, p, m, fmargin, &d)) mk |= (unsigned short)(1u << p); }
          // edge-edge contact, once per unordered pair: the pair is only QUEUED here (a few pairs per warp qualify; testing them in
          // place would cost every warp the whole nine-axis loop) -- pass 1b below tests the queue with one pair per thread
          if (EDGE && a < t && !edge_axes_parallel(G.C)) elist[atomicAdd(&M.nelist, 1)] = (unsigned short)i;
        }
        pmask[i] = mk;
        mycount += mask_count(mk);
      }
      if (EDGE) {
        __syncthreads();
        // pass 1b (EDGE): the rest of the separating-axis test for the queued pairs; bits 12-15 of the pair mask: edge pair + 1
        const int nel = M.nelist;
        for (int j = tid; j < nel; j += SIM_THREADS) {
          const int i = elist[j];
          int lo = 0, hi = n_owner - 1;                  // owner a with poff[a] <= i < poff[a+1]
          while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (M.poff[mid] <= i) lo = mid; else hi = mid - 1; }
          const int ea = lo, et = M.cand[ea][i - M.poff[ea]];
          const float m = (margin + M.sab[ea].w + M.sab[et].w) * gs;
          PairGeom G;
          pair_geom(M, ea, et, m, G, false);
          int erc; float ebe;
          // speculative range of an edge-edge contact: the travel bounds plus at most the geometric tolerance of the contact offset (oracle: me)
          const float me = ((margin < fmargin ? margin : fmargin) + M.sab[ea].w + M.sab[et].w) * gs;
          if (edge_sat(G, me, epref, &erc, &ebe)) pmask[i] = (unsigned short)(pmask[i] | ((erc + 1) << EDGE_POINT));
        }
        __syncthreads();
        if (nel > 0 || strided) { mycount = 0; for (int i = p0; i < p1; ++i) mycount += mask_count(pmask[i]); }
      }
      // running contact offsets: warp-level inclusive scan of the per-thread counts + the totals of the warps before (one barrier)
      incl = mycount;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
      if ((tid & 31) == 31) M.scan[tid >> 5] = incl;
      if (tid < NBODY) { M.astart[tid] = 0; M.aend[tid] = 0; M.nb[tid] = 0; }
      if (EDGE && tid == 0) M.nelist = 0;                        // for the next level / the next sub-step (everyone has read it)
      __syncthreads();
      total = 0;
#pragma unroll
      for (int w = 0; w < SIM_THREADS / 32; ++w) total += M.scan[w];
      if (total <= MAXC || level == 3) break;
      ++level;
      gs = level == 3 ? 0.0f : gs * 0.5f;
      __syncthreads();                                           // everyone has read the warp totals before they are rewritten
    }
    shed_max = max(shed_max, level);
    PMARK(5);
    {
      int run = incl - mycount;
#pragma unroll
      for (int w = 0; w < SIM_THREADS / 32; ++w) { const int v = M.scan[w]; if (w < (tid >> 5)) run += v; }
      if (tid == 0) { M.ncon = total < MAXC ? total : MAXC; M.ndropped = total > MAXC ? total - MAXC : 0; }
      for (int i = p0; i < p1; ++i) { pstart[i] = (unsigned short)min(run, 65535); run += mask_count(pmask[i]); }
    }
    __syncthreads();
    PMARK(6);
    // pass 2 (contact-parallel): slot -> (pair, point) by binary search over the pair offsets, then regenerate the hit
    {
      const int ncon_w = M.ncon;
      unsigned short stash_mask[MAXC / SIM_THREADS], stash_pair[MAXC / SIM_THREADS];   // read the tables BEFORE they are overwritten
      unsigned char stash_j[MAXC / SIM_THREADS];
      uint32_t* const equeue = reinterpret_cast<uint32_t*>(CB);                      // EDGE: queue of pass 2e (CB is written by the loop after it)
#pragma unroll
      for (int r = 0; r < MAXC / SIM_THREADS; ++r) {
        int slot = r * SIM_THREADS + tid;
        stash_pair[r] = 0; stash_mask[r] = 0; stash_j[r] = 0;
        if (slot < ncon_w) {
          int lo = 0, hi = npairs - 1;                 // largest i with pstart[i] <= slot (that pair is non-empty)
          while (lo < hi) { int mid = (lo + hi + 1) >> 1; if ((int)pstart[mid] <= slot) lo = mid; else hi = mid - 1; }
          stash_pair[r] = (unsigned short)lo; stash_mask[r] = pmask[lo]; stash_j[r] = (unsigned char)(slot - (int)pstart[lo]);
          // EDGE: the pair's last contact is its edge-edge contact -- queued (slot | pair) for pass 2e below: one thread per edge contact
          // instead of every warp that holds one walking through the closest-point construction
          if (EDGE && (stash_mask[r] >> EDGE_POINT) && (int)stash_j[r] >= __popc(stash_mask[r] & 0xfffu))
            equeue[atomicAdd(&M.nelist, 1)] = (uint32_t)slot | ((uint32_t)lo << 16);
        }
      }
      __syncthreads();
      if (EDGE) {
        // pass 2e (EDGE): closest points of the two edges; world point | depth -> CA[slot], world normal | target edge axis -> CN[slot]
        const int nel = M.nelist;
        for (int q = tid; q < nel; q += SIM_THREADS) {
          const uint32_t ent = equeue[q];
          const int slot = ent & 0xffff, i = ent >> 16;
          int lo = 0, hi = n_owner - 1;
          while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (M.poff[mid] <= i) lo = mid; else hi = mid - 1; }
          const int a = lo, t = M.cand[a][i - M.poff[a]];
          PairGeom G;
          pair_geom(M, a, t, 1e30f, G, false);
          EdgeGeom E;
          edge_point(G, (int)(pmask[i] >> EDGE_POINT) - 1, E);
          const v3 wpt = vadd(ld3(M.sc[t]), mmul(M.sR[t], E.p));
          const v3 nw = mmul(M.sR[t], E.n);
          CA[slot] = make_float4(wpt.x, wpt.y, wpt.z, E.depth);
          CN[slot] = make_float4(nw.x, nw.y, nw.z, __int_as_float(E.r));
        }
        __syncthreads();
        if (tid == 0) M.nelist = 0;
      }
#pragma unroll 1
      for (int r = 0; r < MAXC / SIM_THREADS; ++r) {
        int slot = r * SIM_THREADS + tid;
        if (slot >= ncon_w) break;
        // warm start, first probe: the cached record of the SAME slot (key | impulse: one 16-byte load), issued here so that its trip to L2 /
        // HBM runs under the geometry below instead of after it
        float4 w4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        int mid = -1;
        if (warm > 0.0f && nprev > 0) { mid = slot < nprev - 1 ? slot : nprev - 1; w4 = *reinterpret_cast<const float4*>(wsr + 4 * mid); }
        int i = stash_pair[r];
        unsigned mk = stash_mask[r];
        int j = stash_j[r], p = 0;
        if (EDGE && j >= __popc(mk & 0xfffu)) p = EDGE_POINT;             // the pair's last contact: its edge-edge contact
        else for (;; ++p) { if (mk & (1u << p)) { if (j == 0) break; --j; } }
        int lo = 0, hi = n_owner - 1;                  // owner a with poff[a] <= i < poff[a+1]
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (M.poff[mid] <= i) lo = mid; else hi = mid - 1; }
        const int a = lo, t = M.cand[a][i - M.poff[a]];
        float m = (margin + M.sab[a].w + M.sab[t].w) * gs;
        float depth;
        v3 wpt;
        uint32_t wdn;
        if (EDGE && p == EDGE_POINT) {                                    // geometry from pass 2e
          const float4 g = CA[slot];
          wpt = V3(g.x, g.y, g.z); depth = g.w;
          wdn = (uint32_t)M.sbody[a] | ((uint32_t)M.sbody[t] << 8) | ((uint32_t)t << 16) | ((uint32_t)__float_as_int(CN[slot].w) << 24) | EDGE_BIT;
        } else {
          PairGeom G;
          pair_geom(M, a, t, m, G, t >= NB + nrs);
          point_hit(G, p, m, fmargin, &depth);
          wpt = vadd(ld3(M.sc[a]), mmul(M.sR[a], sample_point(G.ha, p)));
          wdn = (uint32_t)M.sbody[a] | ((uint32_t)M.sbody[t] << 8) | ((uint32_t)t << 16) | ((uint32_t)G.k << 24) | (G.sg << 26);
        }
        float bias = 0.0f;
        if (depth > S->slop) { bias = S->baumgarte * (depth - S->slop) / h; if (bias > S->max_depen_vel) bias = S->max_depen_vel; }
        else if (depth < 0.0f) bias = depth / h;
        CA[slot] = make_float4(wpt.x, wpt.y, wpt.z, bias);
        CB[slot] = make_float4(depth, 0.0f, 0.0f, __uint_as_float(wdn));   // .x carries the depth until the inverse masses are computed
        {   // warm start from the cached impulse of the same (owner shape, target shape, sample point), if it persisted
          const uint32_t key = ((uint32_t)a << 12) | ((uint32_t)t << 4) | (uint32_t)p;
          wsw[4 * slot] = __uint_as_float(key);
          v3 f0 = V3(0.0f, 0.0f, 0.0f);
          if (warm > 0.0f) {
            // the keys of both sub-steps ascend with the contact order and most contacts persist in place: look at the same
            // slot first, gallop away from it, then bisect the bracket (finds what a plain bisection finds, in 1-3 probes
            // for a settled heap instead of 10 dependent global loads)
            int lo2 = 0, hi2 = nprev - 1, step = 1, dir = 0;
            bool first = true;
            while (lo2 <= hi2) {
              if (!first) w4 = *reinterpret_cast<const float4*>(wsr + 4 * mid);
              first = false;
              const uint32_t kv = __float_as_uint(w4.x);
              if (kv == key) {
                // a contact that involves a robot link or a HOT brick (hit by the robot / faster than the wake threshold in the last sub-step)
                // moves too fast for its last impulse to be trusted as far as a resting contact's (oracle: hotc)
                const bool hotc = a >= NB || (M.sflag[M.sbody[a]] & 2) || (t < NB ? (M.sflag[M.sbody[t]] & 2) != 0 : t < NB + nrs);
                const float wf = hotc ? S->warm_start_hot : warm;
                f0 = V3(wf * w4.y, wf * w4.z, wf * w4.w); break;
              }
              const int d = kv < key ? 1 : -1;
              if (d > 0) lo2 = mid + 1; else hi2 = mid - 1;
              if (dir == 0) dir = d;
              if (dir == d) { mid += d * step; step <<= 1; if (mid < lo2 || mid > hi2) { dir = 2; mid = (lo2 + hi2) >> 1; } }
              else { dir = 2; mid = (lo2 + hi2) >> 1; }
            }
          }
          CF[slot] = make_float4(f0.x, f0.y, f0.z, 0.0f);
        }
      }
    }
    __syncthreads();
    PMARK(7);
    const int ncon = M.ncon;
    unsigned ract = 0u;                                        // robot warp: links with at least one contact
    // An env without a single contact in this sub-step (a heap that sleeps, the hand in the air) has nothing to solve: bodies
    // keep their free velocities, which is exactly what 17 empty passes would leave -- 40 barriers and three scans are skipped.
    if (ncon > 0) {
    // 7. incidence: owned contacts of a body are one contiguous range (contacts are generated owner-major);
    //    target-side contacts go to a per-body list, filled with atomics and then sorted ascending
    //    (=> the summation order of phase B is fixed, whatever the fill order was)
    for (int i = tid; i < ncon; i += SIM_THREADS) {
      uint32_t wd = __float_as_uint(CB[i].w);
      int a = wd & 255, b = (wd >> 8) & 255;
      if (i == 0 || (int)(__float_as_uint(CB[i - 1].w) & 255) != a) M.astart[a] = i;
      if (i == ncon - 1 || (int)(__float_as_uint(CB[i + 1].w) & 255) != a) M.aend[a] = i + 1;
      if (b != STATIC_BODY) atomicAdd(&M.nb[b], 1);
      // sleeping: who touched whom (all writers of a flag byte OR in bits => atomicOr on the containing word)
      if (a < NB && b != STATIC_BODY) { if (b >= NB) touch_or(M.touch, a, 1u); else if (M.sflag[b] & 2) touch_or(M.touch, a, 2u); }
      if (b < NB) { if (a >= NB) touch_or(M.touch, b, 1u); else if (M.sflag[a] & 2) touch_or(M.touch, b, 2u); }
    }
    __syncthreads();
    if (tid < 32) {
      for (int b = tid; b < NBODY; b += 32) M.boff[b] = M.nb[b];
      __syncwarp();
      int tot = warp_excl_scan(M.boff, NBODY, tid);
      if (tid == 0) M.boff[NBODY] = tot;
      __syncwarp();
      for (int b = tid; b < NBODY; b += 32) M.bcur[b] = M.boff[b];
    }
    __syncthreads();
    for (int i = tid; i < ncon; i += SIM_THREADS) {
      int b = (__float_as_uint(CB[i].w) >> 8) & 255;
      if (b != STATIC_BODY) M.blist[atomicAdd(&M.bcur[b], 1)] = (unsigned short)i;
    }
    __syncthreads();
    if (tid < NBODY) {
      const int o0 = M.boff[tid], o1 = M.boff[tid + 1];
      for (int i = o0 + 1; i < o1; ++i) {      // insertion sort (lists are short)
        unsigned short key = M.blist[i];
        int j = i - 1;
        while (j >= o0 && M.blist[j] > key) { M.blist[j + 1] = M.blist[j]; --j; }
        M.blist[j + 1] = key;
      }
      M.nb[tid] = (M.aend[tid] - M.astart[tid]) + (o1 - o0);
    }
    __syncthreads();
    if (tid < 32) {                                            // warp 0: lay out phase B's lane groups
      // A brick that is awake and touched gets a group of L = 2^lg lanes (phaseb_lg: at most 4 incidences per lane).  Groups are
      // placed class by class in DESCENDING size, so every group is aligned to its own size and never straddles a warp; where a
      // group sits does not matter for the result (its lanes only talk to each other, by xor-shuffles).
      bool act[(NB + 31) / 32]; int lgs[(NB + 31) / 32], pos[(NB + 31) / 32];
#pragma unroll
      for (int r = 0; r < (NB + 31) / 32; ++r) {
        const int b = r * 32 + tid;
        act[r] = b < NB && !(M.sflag[b] & 1) && M.nb[b] > 0;
        lgs[r] = act[r] ? phaseb_lg(M.nb[b]) : -1;
        pos[r] = -1;
      }
      int total = 0;
#pragma unroll 1
      for (int c = 5; c >= 0; --c) {
        int run = 0;
#pragma unroll
        for (int r = 0; r < (NB + 31) / 32; ++r) {
          const unsigned bal = __ballot_sync(0xffffffffu, lgs[r] == c);
          if (lgs[r] == c) pos[r] = total + ((run + __popc(bal & ((1u << tid) - 1u))) << c);
          run += __popc(bal);
        }
        total += run << c;
      }
#pragma unroll
      for (int r = 0; r < (NB + 31) / 32; ++r) {
        const int b = r * 32 + tid;
        if (b < NB) M.bcur[b] = pos[r];
        if (act[r]) {                                            // everything a lane needs in one LDS.128
          const int a0 = M.astart[b];
          M.irec[b] = make_int4(a0, M.aend[b] - a0, M.boff[b], b | (M.nb[b] << 8) | (lgs[r] << 20));
        }
      }
      if (tid == 0) M.nact = total;
    }
    if (tid >= ROBOT_TID0) {
      const int L = tid - ROBOT_TID0;
      const bool has = L < SDX_NL && M.nb[NB + L] > 0;
      ract = __ballot_sync(0xffffffffu, has);
      if (L < SDX_NL && !has) { st3(M.linkF[L], V3(0.0f, 0.0f, 0.0f)); st3(M.linkM[L], V3(0.0f, 0.0f, 0.0f)); }
      if (L < SDX_ND) {
        int nn = 0;
        unsigned m = my_desc & ract;                           // descendants in contact (the others count zero)
        while (m) { int L2 = __ffs(m) - 1; m &= m - 1; nn += M.nb[NB + L2]; }
        M.nj[L] = nn;
      }
    }
    if (condump && sub == substeps - 1)
      for (int i = tid; i < ncon; i += SIM_THREADS) {
        float* o = condump + ((size_t)e * MAXC + i) * 8;
        float4 A4 = CA[i], B4 = CB[i];
        o[0] = B4.w; o[1] = A4.x; o[2] = A4.y; o[3] = A4.z; o[4] = B4.x;
        o[5] = A4.w; o[6] = 0.0f; o[7] = 0.0f;
      }
    __syncthreads();
    PMARK(8);
    // 8. inverse mass-split effective masses along n, t1, t2  (|| the lane -> brick map of phase B, group by group)
    if (tid < NB) {
      const int pos = M.bcur[tid];
      if (pos >= 0) { const int L = 1 << ((M.irec[tid].w >> 20) & 7); for (int k = 0; k < L; ++k) M.lmap[pos + k] = (unsigned char)tid; }
    }
    for (int i = tid; i < ncon; i += SIM_THREADS) {
      const float4 A4 = CA[i];
      uint32_t wd = __float_as_uint(CB[i].w);
      int a = wd & 255, b = (wd >> 8) & 255;
      v3 wpt = V3(A4.x, A4.y, A4.z);
#if SIM_SMALL_CODE
      float inv[3];
      const int sh = (wd >> 16) & 255, k = (wd >> 24) & 3;
      v3 en = V3(0.0f, 0.0f, 0.0f), et1 = en, et2 = en;
      if (EDGE && (wd & EDGE_BIT)) contact_axes_e<EDGE>(M, wd, CN, i, &en, &et1, &et2);
#pragma unroll 1
      for (int ax = 0; ax < 3; ++ax) {                        // n, t1, t2 in turn (contact_axes), one copy of body_k x 2
        const int col = k + ax >= 3 ? k + ax - 3 : k + ax;
        v3 d = mcol(M.sR[sh], col);
        if (ax == 0) d = vscale(d, ((wd >> 26) & 1) ? -1.0f : 1.0f);
        if (EDGE && (wd & EDGE_BIT)) d = ax == 0 ? en : (ax == 1 ? et1 : et2);
        const float iv = 1.0f / (body_k(S, M, a, wpt, d) + body_k(S, M, b, wpt, d));
        if (ax == 0) inv[0] = iv; else if (ax == 1) inv[1] = iv; else inv[2] = iv;
      }
      CB[i] = make_float4(inv[0], inv[1], inv[2], __uint_as_float(wd));
#else
      v3 n, t1, t2; contact_axes_e<EDGE>(M, wd, CN, i, &n, &t1, &t2);
      float i0 = 1.0f / (body_k(S, M, a, wpt, n) + body_k(S, M, b, wpt, n));
      float i1 = 1.0f / (body_k(S, M, a, wpt, t1) + body_k(S, M, b, wpt, t1));
      float i2 = 1.0f / (body_k(S, M, a, wpt, t2) + body_k(S, M, b, wpt, t2));
      CB[i] = make_float4(i0, i1, i2, __uint_as_float(wd));
#endif
    }
    __syncthreads();
    // 9. Jacobi iterations on total impulses
    const float mu = S->friction;
    PMARK(9);
#ifdef SIM_PROFILE
    if (prof && tid == 0) { prof[18 * 2 * 8] = clock64(); prof[18 * 2 * 8 + 1] = ncon; prof[18 * 2 * 8 + 2] = M.nact; }
    if (prof && tid == ROBOT_TID0) prof[18 * 2 * 8 + 3] = (long long)__popc(ract);
#endif
    for (int it = -1; it < iters; ++it) {                      // it = -1: phase B only = apply the warm-start impulses
      if (it >= 0)
      for (int i = tid; i < ncon; i += SIM_THREADS) {          // phase A: one thread per contact
        const float4 A4 = CA[i], B4 = CB[i], F4 = CF[i];
        uint32_t wd = __float_as_uint(B4.w);
        int a = wd & 255, b = (wd >> 8) & 255;
        const float4 N4 = EDGE ? CN[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        v3 n, t1, t2; contact_axes_sel<EDGE>(M, wd, N4, &n, &t1, &t2);
        v3 wpt = V3(A4.x, A4.y, A4.z);
        v3 f = V3(F4.x, F4.y, F4.z);
        v3 vrel = vadd(ldv(M.bv[a]), vcross(ldv(M.bw[a]), vsub(wpt, ldv(M.bx[a]))));
        if (b != STATIC_BODY) vrel = vsub(vrel, vadd(ldv(M.bv[b]), vcross(ldv(M.bw[b]), vsub(wpt, ldv(M.bx[b])))));
        float ln = fmaf(A4.w - vdot(vrel, n), B4.x, vdot(f, n));
        ln = ln > 0.0f ? ln : 0.0f;
        float lim = mu * ln;
        float l1 = clampf(fmaf(-vdot(vrel, t1), B4.y, vdot(f, t1)), -lim, lim);
        float l2 = clampf(fmaf(-vdot(vrel, t2), B4.z, vdot(f, t2)), -lim, lim);
        f = vmad(t2, l2, vmad(t1, l1, vscale(n, ln)));
        CF[i] = make_float4(f.x, f.y, f.z, 0.0f);
      }
#ifdef SIM_PROFILE
      if (prof && (tid & 31) == 0) prof[((it + 1) * 2 + 0) * 8 + (tid >> 5)] = clock64();
#endif
      __syncthreads();
      // phase B: TWO lanes per body, lane k sums incidences e = k (mod 2); partials combined 0+1.
      // brick warps (all but the last): only bricks that are awake and in contact (irec, built once per sub-step: one LDS.128
      // per lane and pass instead of a chain of five), 16 per warp; the other bricks keep their free velocity (zero when asleep).
      // last warp: the articulation -- only links that HAVE contacts are gathered (their wrenches are zero otherwise, set
      // once per sub-step), and the joint-space update is skipped entirely while the robot touches nothing.
      const bool robot_warp = tid >= ROBOT_TID0;
      if (!robot_warp) {
        const int nl = M.nact;
#pragma unroll 1
        for (int base = tid & ~31; base < nl; base += 32 * NBW) {          // warp-uniform: this warp's 32-lane window
          const int p = base + (tid & 31);
          const bool live = p < nl;
          const int4 r = M.irec[live ? (int)M.lmap[p] : 0];
          const int body = live ? r.w & 255 : 0, lg = live ? (r.w >> 20) & 7 : 0, L = 1 << lg, k = p & (L - 1);
          v3 F, T;
          gather_incident(M, CA, CF, r.x, r.y, r.z, live ? (r.w >> 8) & 4095 : 0, k, L, ld3(M.bx[body]), &F, &T);
          const int steps = __reduce_max_sync(0xffffffffu, lg);             // butterfly over the group: p[k] += p[k ^ o], o = 1, 2, 4 ...
#pragma unroll 1
          for (int o = 1; o < (1 << steps); o <<= 1) {
            const float fx = __shfl_xor_sync(0xffffffffu, F.x, o), fy = __shfl_xor_sync(0xffffffffu, F.y, o), fz = __shfl_xor_sync(0xffffffffu, F.z, o);
            const float tx = __shfl_xor_sync(0xffffffffu, T.x, o), ty = __shfl_xor_sync(0xffffffffu, T.y, o), tz = __shfl_xor_sync(0xffffffffu, T.z, o);
            if (o < L) { F.x += fx; F.y += fy; F.z += fz; T.x += tx; T.y += ty; T.z += tz; }
          }
          if (k == 0 && live) {
            const float4 fv = M.bfv[body], fw = M.bfw[body];
            st3(M.bv[body], vmad(F, fv.w, V3(fv.x, fv.y, fv.z)));
            st3(M.bw[body], vadd(V3(fw.x, fw.y, fw.z), brick_Iinv_mul(M.bI0[body], M.bI1[body], T)));
          }
        }
      } else {
        const int n_items = 2 * __popc(ract);
#pragma unroll 1
        for (int base = 0; base < n_items; base += 32) {
          const int item = base + tid - ROBOT_TID0;
          const bool live = item < n_items;
          const int body = NB + (live ? (int)__fns(ract, 0, (item >> 1) + 1) : 0);
          const int a0 = M.astart[body], na = M.aend[body] - a0, b0 = M.boff[body], ntot = live ? na + (M.boff[body + 1] - b0) : 0;
          v3 F, T;
          gather_incident(M, CA, CF, a0, na, b0, ntot, item & 1, 2, V3(0.0f, 0.0f, 0.0f), &F, &T);
          F.x += __shfl_xor_sync(0xffffffffu, F.x, 1); F.y += __shfl_xor_sync(0xffffffffu, F.y, 1); F.z += __shfl_xor_sync(0xffffffffu, F.z, 1);
          T.x += __shfl_xor_sync(0xffffffffu, T.x, 1); T.y += __shfl_xor_sync(0xffffffffu, T.y, 1); T.z += __shfl_xor_sync(0xffffffffu, T.z, 1);
          if (!(item & 1) && live) { st3(M.linkF[body - NB], F); st3(M.linkM[body - NB], T); }
        }
      }
      if (robot_warp && ract != 0u) {                          // joint-space impulse + link twists (one warp)
        __syncwarp();
        if (tid < ROBOT_TID0 + SDX_ND) {
          int j = tid - ROBOT_TID0;
          v3 Fd = V3(0.0f, 0.0f, 0.0f), Md = V3(0.0f, 0.0f, 0.0f);
          unsigned m = my_desc & ract;                         // descendant links IN CONTACT (the others carry no wrench)
          while (m) { int L = __ffs(m) - 1; m &= m - 1; Fd = vadd(Fd, ld3(M.linkF[L])); Md = vadd(Md, ld3(M.linkM[L])); }
          v3 aj = ld3(M.ja[j]);
          float g = vdot(aj, vsub(Md, vcross(ld3(M.jo[j]), Fd)));
          M.qd[j] = M.qdfree[j] + g / M.ieff[j];
        }
        __syncwarp();
        if ((ract >> (tid - ROBOT_TID0)) & 1u) link_twist(S, M, tid - ROBOT_TID0, my_anc);   // phase A reads only the twists of links in contact
      }
#ifdef SIM_PROFILE
      if (prof && (tid & 31) == 0) prof[((it + 1) * 2 + 1) * 8 + (tid >> 5)] = clock64();
#endif
      __syncthreads();
    }
    } else {
      if (tid >= ROBOT_TID0 && tid < ROBOT_TID0 + SDX_NL) { st3(M.linkF[tid - ROBOT_TID0], V3(0.0f, 0.0f, 0.0f)); st3(M.linkM[tid - ROBOT_TID0], V3(0.0f, 0.0f, 0.0f)); }
      PMARK(8); PMARK(9);
    }
    PMARK(10);
    for (int i = tid; i < ncon; i += SIM_THREADS) { const float4 F4 = CF[i]; wsw[4 * i + 1] = F4.x; wsw[4 * i + 2] = F4.y; wsw[4 * i + 3] = F4.z; }
    if (tid == 0) gwsn[1 - rb] = ncon;
    rb = 1 - rb;
    if (iters == 0 && tid < SDX_NL) { st3(M.linkF[tid], V3(0, 0, 0)); st3(M.linkM[tid], V3(0, 0, 0)); }
    // 10. integrate
    if (tid < nbr && asleep) {                                 // pose unchanged; only a touch restarts the counter
      const unsigned tc = M.touch[tid];
      if (tc & 1) slpc = 0; else if (tc & 2) slpc = 1;
      bvr = V3(0.0f, 0.0f, 0.0f); bwr = V3(0.0f, 0.0f, 0.0f);
    } else if (tid < nbr) {
      v3 w = ld3(M.bw[tid]), v = ld3(M.bv[tid]);
      float w2 = vdot(w, w), mw = S->max_ang_vel;
      if (w2 > mw * mw) { w = vscale(w, mw / sqrtf(w2)); }
      float v2 = vdot(v, v), mv = S->max_lin_vel;
      if (v2 > mv * mv) { v = vscale(v, mv / sqrtf(v2)); }
      {
        const float E = 0.5f * (vdot(v, v) + vdot(w, w) * (brad * brad * (1.0f / 3.0f)));
        const unsigned tc = M.touch[tid];
        if ((tc & 1) || E >= S->wake_energy) slpc = 0;
        else if ((tc & 2) || E >= S->sleep_energy) slpc = 1;
        else slpc = slpc + 1 > 255 ? 255 : slpc + 1;
      }
      bvr = v; bwr = w;
      bxr = vmad(v, h, bxr);
      q4 q = bqr;
      float hh = 0.5f * h;
      q4 dq;
      dq.x = hh * (w.x * q.w + w.y * q.z - w.z * q.y);
      dq.y = hh * (w.y * q.w + w.z * q.x - w.x * q.z);
      dq.z = hh * (w.z * q.w + w.x * q.y - w.y * q.x);
      dq.w = hh * (-(w.x * q.x + w.y * q.y + w.z * q.z));
      q.x = q.x + dq.x; q.y = q.y + dq.y; q.z = q.z + dq.z; q.w = q.w + dq.w;
      float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
      q.x = q.x * inv; q.y = q.y * inv; q.z = q.z * inv; q.w = q.w * inv;
      bqr = q;
    } else if (tid < NB) { bvr = ld3(M.bv[tid]); bwr = ld3(M.bw[tid]); }
    if (tid >= ROBOT_TID0 && tid < ROBOT_TID0 + SDX_ND) {
      int j = tid - ROBOT_TID0;
      float qdj = M.qd[j];
      float qn = M.q[j] + h * qdj;
      if (qn < S->dof_lo[j]) { qn = S->dof_lo[j]; if (qdj < 0.0f) qdj = 0.0f; }
      if (qn > S->dof_hi[j]) { qn = S->dof_hi[j]; if (qdj > 0.0f) qdj = 0.0f; }
      M.q[j] = qn; M.qd[j] = qdj;
    }
    __syncthreads();
    PMARK(11);
  }

  // ---- epilogue: brick tile back through shared memory + TMA bulk store; task-visible robot rows
  if (tid < NB) {
    M.tile[0 * NB + tid] = bxr.x; M.tile[1 * NB + tid] = bxr.y; M.tile[2 * NB + tid] = bxr.z;
    M.tile[3 * NB + tid] = bqr.x; M.tile[4 * NB + tid] = bqr.y; M.tile[5 * NB + tid] = bqr.z; M.tile[6 * NB + tid] = bqr.w;
    M.tile[7 * NB + tid] = bvr.x; M.tile[8 * NB + tid] = bvr.y; M.tile[9 * NB + tid] = bvr.z;
    M.tile[10 * NB + tid] = bwr.x; M.tile[11 * NB + tid] = bwr.y; M.tile[12 * NB + tid] = bwr.z;
  }
  if (tid >= ROBOT_TID0) robot_fk(S, M, tid - ROBOT_TID0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gbrick), "r"(tile_s), "r"(13 * NB * 4) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  if (tid >= ROBOT_TID0 && tid < ROBOT_TID0 + SDX_NL) link_twist(S, M, tid - ROBOT_TID0, my_anc);
  if (tid >= ROBOT_TID0 && tid < ROBOT_TID0 + SDX_ND) { int j = tid - ROBOT_TID0; gdof[j] = M.q[j]; gdof[24 + j] = M.qd[j]; }
  __syncthreads();
  if (tid < SDX_NL) {
    int L = tid;
    float* o = link_out + ((size_t)e * SDX_NL + L) * 13;
    o[0] = M.bx[NB + L].x; o[1] = M.bx[NB + L].y; o[2] = M.bx[NB + L].z;
    o[3] = M.lq[L][0]; o[4] = M.lq[L][1]; o[5] = M.lq[L][2]; o[6] = M.lq[L][3];
    o[7] = M.bv[NB + L].x; o[8] = M.bv[NB + L].y; o[9] = M.bv[NB + L].z;
    o[10] = M.bw[NB + L].x; o[11] = M.bw[NB + L].y; o[12] = M.bw[NB + L].z;
    float invh = 1.0f / h;
    float* f = netf + ((size_t)e * SDX_NL + L) * 3;
    f[0] = M.linkF[L][0] * invh; f[1] = M.linkF[L][1] * invh; f[2] = M.linkF[L][2] * invh;
  }
  if (tid >= 32 && tid < 39) {
    int j = tid - 32;
    v3 aj = ld3(M.ja[j]);
    v3 lin = vcross(aj, vsub(ld3(M.bx[NB + 7]), ld3(M.jo[j])));
    float* J = jac7 + (size_t)e * 42;
    J[0 * 7 + j] = lin.x; J[1 * 7 + j] = lin.y; J[2 * 7 + j] = lin.z;
    J[3 * 7 + j] = aj.x; J[4 * 7 + j] = aj.y; J[5 * 7 + j] = aj.z;
  }
  if (tid == 64) { ncontact[4 * e] = M.ncon; ncontact[4 * e + 1] = M.ndropped; ncontact[4 * e + 2] = shed_max; ncontact[4 * e + 3] = M.ndrop_cand | (M.ndrop_static << 16); }
  if (tid < NB) slp[(size_t)e * NB + tid] = (unsigned char)slpc;
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
