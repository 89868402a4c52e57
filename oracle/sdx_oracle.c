/* sdx_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the SeqDex hot path for BlockAssemblyGraspSim.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (seqdex_b200/) never does.
 *
 * PARITY STATUS:
 *   * task ops (pre_physics / control_ik / observations / reward / reset / t-value,
 *     GAE): restated from the reference Python cited per function below, and
 *     PINNED against golden vectors produced by importing the reference's own
 *     functions with Isaac Gym stubbed (oracle/gen_golden.py -> tests/golden/).
 *   * contact step (gym.simulate): the reference executes NVIDIA PhysX, a closed
 *     binary that is absent here -> "PARITY UNPINNED".  This file defines the
 *     algorithm (box-SDF contacts, mass-splitting Jacobi, implicit PD joints); the
 *     CUDA kernel must match it BIT FOR BIT, and physics-invariant tests stand in
 *     for a reference trajectory.
 *
 * Bit-exactness contract with the CUDA path: IEEE fp32, no FMA contraction
 * (gcc -ffp-contract=off  <->  nvcc -fmad=false), correctly-rounded sqrt and
 * division, and our own sincos/exp polynomials (no libm transcendentals), every
 * sum in the order written here.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define SDX_MAX_BRICKS 72
#define SDX_MAX_FIXED 60
#define SDX_NL 24
#define SDX_ND 23
#define SDX_MAX_RSHAPES 32
#define SDX_MAX_STATIC 80
#define SDX_MAX_CONTACTS 1024
#define NB SDX_MAX_BRICKS
#define NBODY (NB + SDX_NL)
#define NSHAPE (NB + SDX_MAX_RSHAPES + SDX_MAX_STATIC)
#define KC 32 /* broad-phase candidates kept per owner shape */
#define STATIC_BODY 255
#define OBS_FRAME 132
#define STATE_FRAME 188

typedef struct sdx_scene_t { /* must mirror include/seqdex_b200.h */
  int n_bricks, n_fixed, n_rshapes, n_static;
  int substeps, iters, max_episode_length, sleep_substeps;
  float dt, gravity_z, contact_offset, friction, baumgarte, slop, max_depen_vel, brick_ang_damp, max_ang_vel,
      max_lin_vel, brick_lin_damp, sleep_energy;
  float base_pos[3], base_quat[4];
  float face_margin;   /* a sample point counts as over the reference face up to this far beyond its edge (NOT the speculative contact_offset) */
  int body_parent[SDX_NL];
  unsigned link_anc_mask[SDX_NL];
  float joint_xyz[SDX_ND * 3], joint_quat[SDX_ND * 4], joint_axis[SDX_ND * 3];
  float dof_lo[SDX_ND], dof_hi[SDX_ND], dof_kp[SDX_ND], dof_kd[SDX_ND], dof_effort[SDX_ND], dof_vmax[SDX_ND],
      dof_inertia[SDX_ND];
  int rs_body[SDX_MAX_RSHAPES];
  float rs_c[SDX_MAX_RSHAPES * 3], rs_quat[SDX_MAX_RSHAPES * 4], rs_h[SDX_MAX_RSHAPES * 3];
  float br_half[SDX_MAX_BRICKS * 3], br_coff[SDX_MAX_BRICKS * 3], br_invm[SDX_MAX_BRICKS],
      br_invI[SDX_MAX_BRICKS * 3];
  float st_c[SDX_MAX_STATIC * 3], st_h[SDX_MAX_STATIC * 3];
  float fixed_root[SDX_MAX_FIXED * 13];
  float brick_init[SDX_MAX_BRICKS * 13];
  float prepare_arm[7], insert_prep0[7], insert_prep1[7], finger_reset_unscaled[16];
  float cam_off_pos[3], cam_off_quat[4];
  float act_moving_average, av_factor, vel_obs_scale, warm_start, wake_energy;
  int task;
  float hand_target_quat[4];
  int bank_sample_range;
  int pad3[2];
  float default_dof[SDX_ND];  /* Search: arm_hand_default_dof_pos, the pose that parks the hand beside the bin (SE:207-211) */
  float prepare_dof[SDX_ND];  /* Search: arm_hand_prepare_dof_pos_list[0], where an episode starts (SE:220-223, 316) */
  float insert_plate_zw[2];   /* InsertSim: (z, w) of gymapi.Quat.from_euler_zyx(0, 0, 1.57), the base-plate's second yaw (IS:1436-1437) */
  int st_mod[SDX_MAX_STATIC], st_rem[SDX_MAX_STATIC];   /* static s exists only in envs with env % st_mod == st_rem (st_mod 0: everywhere) */
  int n_bshapes;                       /* collision boxes of the free bodies; 0 = every body is one box (n_bricks boxes, box a = body a) */
  int bs_body[SDX_MAX_BRICKS];         /* box -> body; boxes of one body are consecutive */
  float bs_c[SDX_MAX_BRICKS * 3];      /* box centre in its body's COM frame (axes = the body's) */
  float tool_reset_pos[3];             /* tool tasks: where reset_idx puts the tool (TG:1496-1498) */
  float tool_pitch_sc[8];              /* (sin, cos) of k * 1.571 / 2, k = 0..3 (TG:1493-1495) */
  float tool_plate_pose[7];            /* the "extra lego" pose after reset_idx (TG:1505-1512) */
  float edge_contacts, edge_pref;      /* edge-edge contacts on (> 0.5) | how much smaller than every face overlap the edge overlap must be [m] */
  float warm_start_hot;                /* warm start of contacts that involve a robot link or a hot brick */
} sdx_scene_t;
#define ORIENT_OBS_FRAME 62
#define ORIENT_BANK_WRAP 10000

/* ------------------------------------------------------------------ math */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } q4;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vscale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
/* explicit single-rounding FMAs (the CUDA side uses the same fmaf() calls; nothing else is contracted) */
static inline float vdot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline v3 vcross(v3 a, v3 b) {
  return V3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
static inline v3 vmad(v3 a, float s, v3 b) { return V3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); } /* a*s + b */
static inline v3 vneg(v3 a) { return V3(-a.x, -a.y, -a.z); }

/* isaacgym.torch_utils.quat_mul (xyzw), same operation order as the public
 * IsaacGymEnvs torch_jit_utils restatement. */
static inline q4 qmul(q4 a, q4 b) {
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
static inline q4 qconj(q4 a) { q4 r = {-a.x, -a.y, -a.z, a.w}; return r; }
/* quat_apply: b + w*t + xyz x t, t = 2 (xyz x b) */
static inline v3 qrot(q4 q, v3 b) {
  v3 xyz = V3(q.x, q.y, q.z);
  v3 t = vscale(vcross(xyz, b), 2.0f);
  return vadd(vmad(t, q.w, b), vcross(xyz, t));
}
/* rotation matrix (row-major) of a unit quaternion */
static inline void qmat(q4 q, float* R) {
  float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z, xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z,
        xw = q.x * q.w, yw = q.y * q.w, zw = q.z * q.w;
  R[0] = 1.0f - 2.0f * (yy + zz); R[1] = 2.0f * (xy - zw); R[2] = 2.0f * (xz + yw);
  R[3] = 2.0f * (xy + zw); R[4] = 1.0f - 2.0f * (xx + zz); R[5] = 2.0f * (yz - xw);
  R[6] = 2.0f * (xz - yw); R[7] = 2.0f * (yz + xw); R[8] = 1.0f - 2.0f * (xx + yy);
}
static inline v3 mcol(const float* R, int k) { return V3(R[k], R[3 + k], R[6 + k]); }
static inline v3 mmul(const float* R, v3 a) {
  return V3(fmaf(R[2], a.z, fmaf(R[1], a.y, R[0] * a.x)), fmaf(R[5], a.z, fmaf(R[4], a.y, R[3] * a.x)),
            fmaf(R[8], a.z, fmaf(R[7], a.y, R[6] * a.x)));
}
static inline v3 mtmul(const float* R, v3 a) {
  return V3(fmaf(R[6], a.z, fmaf(R[3], a.y, R[0] * a.x)), fmaf(R[7], a.z, fmaf(R[4], a.y, R[1] * a.x)),
            fmaf(R[8], a.z, fmaf(R[5], a.y, R[2] * a.x)));
}

/* sin/cos for |x| < ~100: Cody-Waite reduction by pi/2, cephes sinf/cosf minimax kernels */
static inline void sdx_sincos(float x, float* s, float* c) {
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
/* exp: cephes expf kernel, exact 2^n scaling */
static inline float sdx_exp(float x) {
  if (x > 88.0f) x = 88.0f;
  if (x < -87.0f) x = -87.0f;
  float n = rintf(x * 1.44269504088896341f);
  float r = x - n * 0.693359375f;
  r = r - n * -2.12194440e-4f;
  float z = r * r;
  float p = ((((1.9875691500e-4f * r + 1.3981999507e-3f) * r + 8.3334519073e-3f) * r + 4.1665795894e-2f) * r +
             1.6666665459e-1f) * r + 5.0000001201e-1f;
  float y = p * z + r + 1.0f;
  union { uint32_t u; float f; } sc;
  sc.u = (uint32_t)((int)n + 127) << 23;
  return y * sc.f;
}
static inline float sdx_elu(float x) { return x > 0.0f ? x : sdx_exp(x) - 1.0f; }
static inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
/* torch_utils.scale / unscale */
static inline float scalef(float x, float lo, float hi) { return 0.5f * (x + 1.0f) * (hi - lo) + lo; }
static inline float unscalef(float x, float lo, float hi) { return (2.0f * x - hi - lo) / (hi - lo); }

/* Philox4x32-10, key = seed, counter = (env, episode, stream, 0) */
static inline void philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c[4] = {c0, c1, c2, 0u};
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1,
             n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* ------------------------------------------------------------------ kinematics */
typedef struct {
  v3 lx[SDX_NL];   /* link origin (world, env-local) */
  q4 lq[SDX_NL];
  v3 ja[SDX_ND];   /* joint axis world */
  v3 jo[SDX_ND];   /* joint origin world */
} fk_t;

/* forward kinematics of the collapsed Panda+Allegro tree */
static void robot_fk(const sdx_scene_t* S, const float* q, fk_t* K) {
  K->lx[0] = V3(S->base_pos[0], S->base_pos[1], S->base_pos[2]);
  K->lq[0].x = S->base_quat[0]; K->lq[0].y = S->base_quat[1]; K->lq[0].z = S->base_quat[2]; K->lq[0].w = S->base_quat[3];
  for (int j = 0; j < SDX_ND; ++j) {
    int L = j + 1, P = S->body_parent[L];
    q4 qf = {S->joint_quat[4 * j], S->joint_quat[4 * j + 1], S->joint_quat[4 * j + 2], S->joint_quat[4 * j + 3]};
    v3 ax = V3(S->joint_axis[3 * j], S->joint_axis[3 * j + 1], S->joint_axis[3 * j + 2]);
    q4 qj = qmul(K->lq[P], qf);
    v3 x = vadd(K->lx[P], qrot(K->lq[P], V3(S->joint_xyz[3 * j], S->joint_xyz[3 * j + 1], S->joint_xyz[3 * j + 2])));
    float s, c;
    sdx_sincos(0.5f * q[j], &s, &c);
    q4 qr = {ax.x * s, ax.y * s, ax.z * s, c};
    K->lq[L] = qmul(qj, qr);
    K->lx[L] = x;
    K->ja[j] = qrot(qj, ax);
    K->jo[j] = x;
  }
}

/* ------------------------------------------------------------------ contact step */
typedef struct {
  uint32_t word;   /* a_body | b_body<<8 | b_shape<<16 | axis<<24 | sign<<26 */
  float w[3];      /* contact point, world */
  float bias;      /* target normal velocity */
  float inv[3];    /* 1 / (mass-split effective inverse mass) along n, t1, t2 */
  float f[3];      /* total impulse of the contact, world frame: lam_n n + lam_1 t1 + lam_2 t2 */
  float n[3];      /* edge-edge contacts (word & EDGE_BIT): the normal, world frame */
} contact_t;

typedef struct {
  /* bodies: bricks 0..71 (origin = COM), links 72..95 (origin = link frame origin) */
  v3 bx[NBODY]; q4 bq[NBODY]; float bR[NBODY][9];
  v3 bv[NBODY], bw[NBODY];
  v3 vfree[NB], wfree[NB];
  float wI[NB][6];   /* world-frame inverse inertia R diag(1/I) R^T of a brick: xx xy xz yy yz zz, once per sub-step */
  float q[SDX_ND], qd[SDX_ND], tgt[SDX_ND], qdfree[SDX_ND], ieff[SDX_ND];
  fk_t K;
  /* target boxes: bricks, robot shapes, statics */
  v3 sc[NSHAPE]; float sR[NSHAPE][9]; v3 sh[NSHAPE]; float srad[NSHAPE]; int sbody[NSHAPE];
  v3 sa[NSHAPE]; float spd[NSHAPE];
  unsigned char cand[NB + SDX_MAX_RSHAPES][KC]; int ncand[NB + SDX_MAX_RSHAPES];
  contact_t con[SDX_MAX_CONTACTS]; int ncon, ndropped;
  int nb[NBODY]; int nj[SDX_ND];
  /* incidence of body b, in summation order: owner-side contacts [astart,aend) then target-side list (ascending) */
  int astart[NBODY], aend[NBODY], boff[NBODY + 1]; unsigned short blist[SDX_MAX_CONTACTS];
  v3 linkF[SDX_NL], linkM[SDX_NL];
  uint32_t ckey[SDX_MAX_CONTACTS];
  unsigned char asleep[NB], hot[NB], touch[NB];   /* sleeping (see sim_env): touch bit0 = robot, bit1 = hot brick */
  float brad[NB];   /* bounding radius of a free body about its COM: max over its boxes of |centre| + half diagonal */
  unsigned char built_asleep[NB]; int cand_dropped, cand_dropped_static; /* candidate lists are kept over the sub-steps of a step (see sim_env 3.) */
  int shed_level;   /* most speculative-contact shedding any sub-step of the step needed (0 = none) */
} work_t;

/* W = R diag(1/I) R^T (symmetric, six numbers), formed ONCE per sub-step and brick -- the kernel keeps it as two 16-byte
   shared-memory records -- so that every later product with a vector is nine multiply-adds and no rotation matrix */
static inline void brick_world_invI(const float* R, const float* invI, float* w) {
  const float a0x = R[0] * invI[0], a0y = R[1] * invI[1], a0z = R[2] * invI[2];
  const float a1x = R[3] * invI[0], a1y = R[4] * invI[1], a1z = R[5] * invI[2];
  const float a2x = R[6] * invI[0], a2y = R[7] * invI[1], a2z = R[8] * invI[2];
  w[0] = fmaf(a0z, R[2], fmaf(a0y, R[1], a0x * R[0]));
  w[1] = fmaf(a0z, R[5], fmaf(a0y, R[4], a0x * R[3]));
  w[2] = fmaf(a0z, R[8], fmaf(a0y, R[7], a0x * R[6]));
  w[3] = fmaf(a1z, R[5], fmaf(a1y, R[4], a1x * R[3]));
  w[4] = fmaf(a1z, R[8], fmaf(a1y, R[7], a1x * R[6]));
  w[5] = fmaf(a2z, R[8], fmaf(a2y, R[7], a2x * R[6]));
}
static inline v3 brick_Iinv_mul(const work_t* W, int b, v3 u) {
  const float* w = W->wI[b];
  return V3(fmaf(w[2], u.z, fmaf(w[1], u.y, w[0] * u.x)), fmaf(w[4], u.z, fmaf(w[3], u.y, w[1] * u.x)),
            fmaf(w[5], u.z, fmaf(w[4], u.y, w[2] * u.x)));
}

/* lanes that share a brick's incidences in phase B: the smallest power of two that leaves each lane at most 4 (capped at a warp) */
static inline int phaseb_lanes(int ninc) {
  int L = 1;
  while (L < 32 && 4 * L < ninc) L <<= 1;
  return L;
}

static void link_twists(const sdx_scene_t* S, work_t* W) {
  for (int L = 0; L < SDX_NL; ++L) {
    v3 w = V3(0, 0, 0), v = V3(0, 0, 0);
    unsigned m = S->link_anc_mask[L];
    for (int j = 0; j < SDX_ND; ++j)
      if (m & (1u << j)) {
        w = vmad(W->K.ja[j], W->qd[j], w);
        v = vmad(vcross(W->K.ja[j], vsub(W->K.lx[L], W->K.jo[j])), W->qd[j], v);
      }
    W->bv[NB + L] = v; W->bw[NB + L] = w;
  }
}

#define EDGE_POINT 12
#define EDGE_BIT (1u << 27)
static void contact_axes(const work_t* W, const contact_t* c, v3* n, v3* t1, v3* t2) {
  int sh = (c->word >> 16) & 255, k = (c->word >> 24) & 3;
  if (c->word & EDGE_BIT) { /* explicit normal; t1 = the target's edge direction (perpendicular to n by construction), t2 = n x t1 */
    *n = V3(c->n[0], c->n[1], c->n[2]);
    *t1 = mcol(W->sR[sh], k);
    *t2 = vcross(*n, *t1);
    return;
  }
  float sg = ((c->word >> 26) & 1) ? -1.0f : 1.0f;
  *n = vscale(mcol(W->sR[sh], k), sg);
  *t1 = mcol(W->sR[sh], (k + 1) % 3);
  *t2 = mcol(W->sR[sh], (k + 2) % 3);
}

static float body_k(const sdx_scene_t* S, const work_t* W, int body, v3 wpt, v3 d) {
  if (body == STATIC_BODY) return 0.0f;
  if (body < NB) {
    if (W->asleep[body]) return 0.0f;   /* a sleeping brick is immovable for this sub-step */
    v3 rxd = vcross(vsub(wpt, W->bx[body]), d);
    float k = S->br_invm[body] + vdot(rxd, brick_Iinv_mul(W, body, rxd));
    return (float)W->nb[body] * k;
  }
  int L = body - NB;
  unsigned m = S->link_anc_mask[L];
  float k = 0.0f;
  for (int j = 0; j < SDX_ND; ++j)
    if (m & (1u << j)) {
      float g = vdot(W->K.ja[j], vcross(vsub(wpt, W->K.jo[j]), d));
      k = k + (float)W->nj[j] * (g * g) / W->ieff[j];
    }
  return k;
}

/* one env, one control step = `substeps` sub-steps (gym.simulate, BT:140; yaml sim: substeps 2,
 * 16 position iterations).  brick: [13][72]; dof: [3][24]; link_out: [24][13]; jac7: [6][7];
 * netf: [24][3]; ncontact: [4] = contacts of the last sub-step | contacts beyond the table (after shedding) | shed level | candidate
 * pairs beyond KC (low 16 bits; of which against statics: high 16 bits); condump: [MAXC][8] or NULL */
/* ws: [2][MAXC][4] impulse cache of this env (key bits, f.xyz), wsn: [2] entry counts, ws_cur: buffer holding the latest list
 *
 * SLEEPING (what PhysX does to resting actors; its defaults apply to the reference because the yaml sets none):
 * slp[b] counts the sub-steps since brick b was last energetic; 0 = HOT.  A brick with slp >= sleep_substeps is ASLEEP
 * for the sub-step: zero velocity, no gravity, infinite mass, not integrated, and pairs of two non-awake boxes (asleep
 * brick vs asleep brick / static) are dropped in the broad phase.  With E = (|v|^2 + |w|^2 r^2 / 3) / 2 (r = half diagonal
 * of the box) of the solved velocities, at the end of the sub-step
 *     asleep: touched by a robot link -> 0 (hot); touched by a brick that was hot at the sub-step start -> 1 (woken);
 *             otherwise unchanged
 *     awake : touched by a robot link or E >= wake_energy  -> 0   (hot: it wakes / keeps awake what it touches)
 *             else E >= sleep_energy or touched by a hot brick -> 1   (awake, timer restarted)
 *             else                                            -> slp + 1 (saturating at 255)
 * sleep_substeps = 0 disables the mechanism. */
/* test hook (tests/test_physics_invariants.py): 0 = a fresh broad phase in every sub-step with this sub-step's travel bounds
 * only -- what the kept lists must be equivalent to as long as no candidate is missed */
static int g_broad_reuse = 1;
void sdxo_set_broad_reuse(int on) { g_broad_reuse = on; }
/* audit of the candidate lists: [0] pairs a fresh sweep finds in sub-steps that (re)built the lists, [1] of those not in the lists
 * (KC overflow), [2] / [3] the same in sub-steps that kept the lists, [4] missing although the owner's list had room (a pair
 * that came into range after the lists were built) */
static int g_reuse_audit = 0;
static long g_reuse_stats[5];
void sdxo_reuse_audit(int on) { g_reuse_audit = on; for (int i = 0; i < 5; ++i) g_reuse_stats[i] = 0; }
void sdxo_reuse_stats(long out[5]) { for (int i = 0; i < 5; ++i) out[i] = g_reuse_stats[i]; }

/* EDGE-EDGE contact of the pair (owner a, target t), in t's frame.  The corner-vs-face test cannot see two boxes that cross edge over
 * edge: no corner of either lies over a face of the other until they have sunk centimetres into each other.  This is the rest of the
 * separating-axis test: the owner's three face axes and the nine edge-pair axes t_r x a_c.  When every axis overlaps by more than -m and
 * the axis of LEAST overlap is an edge pair -- by more than `pref` over every face axis -- one contact is generated at the closest points
 * of the two edges, normal = that axis (pointing from t to a), depth = the overlap along it.  (csrc/sdx_sim.cuh: edge_contact, same text.) */
static int edge_contact(const float* C, v3 lc, v3 ha, v3 ht, float m, float pref, v3* p_out, v3* n_out, float* depth_out, int* r_out) {
  /* one axis of each box (nearly) parallel, t_r || a_c: every edge-pair axis t_r' x a_c' then coincides with a face axis of one of the
   * boxes (or vanishes), so a face axis holds the least overlap -- no edge-edge contact (csrc/sdx_sim.cuh: edge_axes_parallel) */
  {
    float mx = C[0] * C[0];
    for (int i = 1; i < 9; ++i) { const float c2 = C[i] * C[i]; mx = c2 > mx ? c2 : mx; }
    if (1.0f - mx < 1e-3f) return 0;
  }
  float A[9];
  for (int i = 0; i < 9; ++i) A[i] = fabsf(C[i]);
  const float hav[3] = {ha.x, ha.y, ha.z}, htv[3] = {ht.x, ht.y, ht.z}, lcv[3] = {lc.x, lc.y, lc.z};
  float of = htv[0] + (A[0] * hav[0] + A[1] * hav[1] + A[2] * hav[2]) - fabsf(lcv[0]);
  { const float o1 = htv[1] + (A[3] * hav[0] + A[4] * hav[1] + A[5] * hav[2]) - fabsf(lcv[1]); if (o1 < of) of = o1; }
  { const float o2 = htv[2] + (A[6] * hav[0] + A[7] * hav[1] + A[8] * hav[2]) - fabsf(lcv[2]); if (o2 < of) of = o2; }
  for (int c = 0; c < 3; ++c) { /* the owner's face axes */
    const float la = lcv[0] * C[c] + lcv[1] * C[3 + c] + lcv[2] * C[6 + c];
    const float oa = hav[c] + (A[c] * htv[0] + A[3 + c] * htv[1] + A[6 + c] * htv[2]) - fabsf(la);
    if (oa < -m) return 0;
    if (oa < of) of = oa;
  }
  float be = 1e30f; int br = -1, bc = -1;
  for (int rc = 0; rc < 9; ++rc) { /* edge pairs: axis t_r x a_c */
    const int r = rc / 3, c = rc - 3 * r;
    const float cc = C[3 * r + c], s2 = 1.0f - cc * cc;
    if (s2 < 1e-3f) continue; /* edges within 2 degrees of parallel: the face axes cover it */
    const int r1 = r == 2 ? 0 : r + 1, r2 = r == 0 ? 2 : r - 1, c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
    const float ra = hav[c1] * A[3 * r + c2] + hav[c2] * A[3 * r + c1];
    const float rb = htv[r1] * A[3 * r2 + c] + htv[r2] * A[3 * r1 + c];
    const float dist = fabsf(lcv[r2] * C[3 * r1 + c] - lcv[r1] * C[3 * r2 + c]);
    const float ov = (ra + rb - dist) / sqrtf(s2);
    if (ov < -m) return 0;
    if (ov < be) { be = ov; br = r; bc = c; }
  }
  if (br < 0 || !(be < of - pref)) return 0; /* a face axis is (about) the axis of least overlap: the corner-face contacts have it */
  const int r = br, c = bc;
  const float cc = C[3 * r + c], s2 = 1.0f - cc * cc, s = sqrtf(s2);
  const v3 ac = V3(C[c], C[3 + c], C[6 + c]); /* the owner's axis c in t's frame */
  v3 n = r == 0 ? V3(0.0f, -ac.z, ac.y) : (r == 1 ? V3(ac.z, 0.0f, -ac.x) : V3(-ac.y, ac.x, 0.0f)); /* e_r x ac */
  n = V3(n.x / s, n.y / s, n.z / s);
  if (n.x * lcv[0] + n.y * lcv[1] + n.z * lcv[2] < 0.0f) n = vneg(n); /* from t towards a */
  const float nv[3] = {n.x, n.y, n.z};
  float pt[3], pa[3] = {lcv[0], lcv[1], lcv[2]};
  for (int k = 0; k < 3; ++k) pt[k] = k == r ? 0.0f : (nv[k] >= 0.0f ? htv[k] : -htv[k]); /* t's edge: its support towards a */
  for (int k = 0; k < 3; ++k) {
    if (k == c) continue;
    const float nk = nv[0] * C[k] + nv[1] * C[3 + k] + nv[2] * C[6 + k];
    const float co = nk >= 0.0f ? -hav[k] : hav[k]; /* a's edge: its support towards t */
    pa[0] = pa[0] + co * C[k]; pa[1] = pa[1] + co * C[3 + k]; pa[2] = pa[2] + co * C[6 + k];
  }
  const float d0[3] = {pa[0] - pt[0], pa[1] - pt[1], pa[2] - pt[2]};
  const float de = d0[r], da = d0[0] * ac.x + d0[1] * ac.y + d0[2] * ac.z;
  float u = (de - cc * da) / s2, v = (cc * de - da) / s2; /* closest points of the two edge LINES, clamped to the edges */
  u = clampf(u, -htv[r], htv[r]); v = clampf(v, -hav[c], hav[c]);
  float qt[3] = {pt[0], pt[1], pt[2]};
  qt[r] = u;
  const v3 qa = V3(pa[0] + v * ac.x, pa[1] + v * ac.y, pa[2] + v * ac.z);
  *p_out = V3(0.5f * (qt[0] + qa.x), 0.5f * (qt[1] + qa.y), 0.5f * (qt[2] + qa.z));
  *n_out = n; *depth_out = be; *r_out = r;
  return 1;
}

static void sim_env(const sdx_scene_t* S, float* brick, float* dof, float* link_out, float* jac7, float* netf,
                    int* ncontact, float* condump, float* ws, int* wsn, int ws_cur, unsigned char* slp, work_t* W, int env) {
  const int nbr = S->n_bricks, nbs = S->n_bshapes > 0 ? S->n_bshapes : S->n_bricks, nrs = S->n_rshapes, nst = S->n_static;   /* bodies | their boxes (a body may be a compound of boxes) */
  const int n_owner = NB + nrs, n_target = NB + nrs + nst;
  const float h = S->dt / (float)S->substeps;
  const float margin = S->contact_offset;
  const float fmargin = S->face_margin;   /* how far beyond the edge of the reference face a sample point still counts as over it */
  W->shed_level = 0;
  for (int b = 0; b < NB; ++b) {
    float r = 0.0f;
    for (int a = 0; a < nbs; ++a)
      if (S->bs_body[a] == b) {
        v3 c = V3(S->bs_c[3 * a], S->bs_c[3 * a + 1], S->bs_c[3 * a + 2]), hh = V3(S->br_half[3 * a], S->br_half[3 * a + 1], S->br_half[3 * a + 2]);
        float cand = sqrtf(vdot(c, c)) + sqrtf(vdot(hh, hh));
        if (cand > r) r = cand;
      }
    W->brad[b] = r;
  }
  for (int b = 0; b < NB; ++b) {
    W->bx[b] = V3(brick[0 * NB + b], brick[1 * NB + b], brick[2 * NB + b]);
    W->bq[b].x = brick[3 * NB + b]; W->bq[b].y = brick[4 * NB + b]; W->bq[b].z = brick[5 * NB + b]; W->bq[b].w = brick[6 * NB + b];
    W->bv[b] = V3(brick[7 * NB + b], brick[8 * NB + b], brick[9 * NB + b]);
    W->bw[b] = V3(brick[10 * NB + b], brick[11 * NB + b], brick[12 * NB + b]);
  }
  for (int j = 0; j < SDX_ND; ++j) { W->q[j] = dof[j]; W->qd[j] = dof[24 + j]; W->tgt[j] = dof[48 + j]; }
  /* static target boxes never change */
  for (int s = 0; s < nst; ++s) {
    int t = NB + nrs + s;
    W->sc[t] = V3(S->st_c[3 * s], S->st_c[3 * s + 1], S->st_c[3 * s + 2]);
    if (S->st_mod[s] > 0 && env % S->st_mod[s] != S->st_rem[s]) W->sc[t].z = -1000.0f;   /* not part of this env's scene (InsertSim's base-plate by env % 3, IS:971-977) */
    W->sh[t] = V3(S->st_h[3 * s], S->st_h[3 * s + 1], S->st_h[3 * s + 2]);
    for (int i = 0; i < 9; ++i) W->sR[t][i] = (i % 4 == 0) ? 1.0f : 0.0f;
    W->sbody[t] = STATIC_BODY;
    W->srad[t] = 0.0f;
  }
  int rb = ws_cur;
  for (int sub = 0; sub < S->substeps; ++sub) {
    const float* wsr = ws + (size_t)rb * SDX_MAX_CONTACTS * 4;
    float* wsw = ws + (size_t)(1 - rb) * SDX_MAX_CONTACTS * 4;
    const int nprev = wsn[rb];
    for (int b = 0; b < NB; ++b) {
      W->asleep[b] = (unsigned char)(b < nbr && S->sleep_substeps > 0 && (int)slp[b] >= S->sleep_substeps);
      W->hot[b] = (unsigned char)(b < nbr && slp[b] == 0);
      W->touch[b] = 0;
    }
    /* 1. kinematics + shape poses */
    robot_fk(S, W->q, &W->K);
    for (int L = 0; L < SDX_NL; ++L) { W->bx[NB + L] = W->K.lx[L]; W->bq[NB + L] = W->K.lq[L]; qmat(W->K.lq[L], W->bR[NB + L]); }
    for (int b = 0; b < NB; ++b) qmat(W->bq[b], W->bR[b]);
    for (int a = 0; a < NB; ++a) {          /* boxes of the free bodies: box a rides on body bs_body[a] at bs_c[a] in that body's COM frame */
      const int b = a < nbs ? S->bs_body[a] : a;
      W->sc[a] = a < nbs ? vadd(W->bx[b], mmul(W->bR[b], V3(S->bs_c[3 * a], S->bs_c[3 * a + 1], S->bs_c[3 * a + 2]))) : W->bx[b];
      for (int i = 0; i < 9; ++i) W->sR[a][i] = W->bR[b][i];
      W->sh[a] = V3(S->br_half[3 * a], S->br_half[3 * a + 1], S->br_half[3 * a + 2]);
      W->srad[a] = sqrtf(vdot(W->sh[a], W->sh[a]));
      W->sbody[a] = b;
    }
    for (int r = 0; r < nrs; ++r) {
      int t = NB + r, L = S->rs_body[r];
      q4 ql = {S->rs_quat[4 * r], S->rs_quat[4 * r + 1], S->rs_quat[4 * r + 2], S->rs_quat[4 * r + 3]};
      W->sc[t] = vadd(W->K.lx[L], qrot(W->K.lq[L], V3(S->rs_c[3 * r], S->rs_c[3 * r + 1], S->rs_c[3 * r + 2])));
      qmat(qmul(W->K.lq[L], ql), W->sR[t]);
      W->sh[t] = V3(S->rs_h[3 * r], S->rs_h[3 * r + 1], S->rs_h[3 * r + 2]);
      W->srad[t] = sqrtf(vdot(W->sh[t], W->sh[t]));
      W->sbody[t] = NB + L;
    }
    /* 2. free velocities: gravity + angular damping on bricks, implicit PD on joints */
    {
      float damp = 1.0f - h * S->brick_ang_damp;
      float ldamp = 1.0f - h * S->brick_lin_damp;
      for (int b = 0; b < NB; ++b) {
        W->vfree[b] = vscale(V3(W->bv[b].x, W->bv[b].y, W->bv[b].z + h * S->gravity_z), ldamp);
        W->wfree[b] = vscale(W->bw[b], damp);
        if (b >= nbr || W->asleep[b]) { W->vfree[b] = V3(0, 0, 0); W->wfree[b] = V3(0, 0, 0); }
        W->bv[b] = W->vfree[b]; W->bw[b] = W->wfree[b];
        brick_world_invI(W->bR[b], &S->br_invI[3 * b], W->wI[b]);
      }
      for (int j = 0; j < SDX_ND; ++j) {
        float I = S->dof_inertia[j], kp = S->dof_kp[j], kd = S->dof_kd[j];
        float ieff = I + h * kd + (h * h) * kp;
        float qn = (I * W->qd[j] + h * kp * (W->tgt[j] - W->q[j])) / ieff;
        float tau = I * (qn - W->qd[j]) / h;
        if (tau > S->dof_effort[j]) qn = W->qd[j] + S->dof_effort[j] * h / I;
        if (tau < -S->dof_effort[j]) qn = W->qd[j] - S->dof_effort[j] * h / I;
        qn = clampf(qn, -S->dof_vmax[j], S->dof_vmax[j]);
        W->ieff[j] = ieff; W->qdfree[j] = qn; W->qd[j] = qn;
      }
      link_twists(S, W);
    }
    /* 3. broad phase: per owner shape, candidate target boxes in index order (world-AABB overlap,
     *    inflated by the contact offset plus the distance both shapes can travel in this sub-step) */
    for (int t = 0; t < n_target; ++t) {
      const float* R = W->sR[t];
      W->sa[t] = V3(fabsf(R[0]) * W->sh[t].x + fabsf(R[1]) * W->sh[t].y + fabsf(R[2]) * W->sh[t].z,
                    fabsf(R[3]) * W->sh[t].x + fabsf(R[4]) * W->sh[t].y + fabsf(R[5]) * W->sh[t].z,
                    fabsf(R[6]) * W->sh[t].x + fabsf(R[7]) * W->sh[t].y + fabsf(R[8]) * W->sh[t].z);
      float spd = 0.0f;
      if (t < NB + nrs) {
        int bd = W->sbody[t];
        v3 dc = vsub(W->sc[t], W->bx[bd]);
        float reach = sqrtf(vdot(dc, dc)) + W->srad[t];
        spd = h * (sqrtf(vdot(W->bv[bd], W->bv[bd])) + sqrtf(vdot(W->bw[bd], W->bw[bd])) * reach);
      }
      W->spd[t] = spd;
    }
    /* The candidate lists are built in the FIRST sub-step of a step for ALL its sub-steps -- the travel bounds scaled by the
     * number of sub-steps left, plus the speed gravity adds in between -- and rebuilt in a later sub-step only if a brick that
     * was asleep when they were built has been woken since (its pairs with sleeping bricks and statics were filtered). */
    int rebuild = sub == 0 || !g_broad_reuse;
    if (sub > 0) for (int b = 0; b < NB; ++b) if (W->built_asleep[b] && !W->asleep[b]) rebuild = 1;
    if (rebuild) {
      const int left = g_broad_reuse ? S->substeps - sub : 1;
      const float infl = (float)left, slack = (float)(left - 1) * ((h * h) * fabsf(S->gravity_z));
      for (int b = 0; b < NB; ++b) W->built_asleep[b] = W->asleep[b];
      W->cand_dropped = 0; W->cand_dropped_static = 0;
      /* per owner: dynamic targets (bricks, then robot boxes) and statics are swept separately; the statics claim their slots
       * FIRST (a brick must never lose its floor / wall pair to a crowd of neighbours), the dynamic targets fill what is left in
       * ascending order; the list itself stays ascending: dynamic part, then statics */
      for (int a = 0; a < n_owner; ++a) {
        unsigned char dyn[KC], sta[KC];
        int nd = 0, nd_all = 0, ns = 0, ns_all = 0;
        if (a < NB && a >= nbs) { W->ncand[a] = 0; continue; }
        for (int t = 0; t < n_target; ++t) {
          if (t < NB && (t >= nbs || (a < NB && W->sbody[t] == W->sbody[a]))) continue;   /* no box, or a box of the same body */
          if (a >= NB && t >= NB && t < NB + nrs) continue; /* robot-robot filtered (GS:906 filter -1) */
          if (a < NB && W->asleep[W->sbody[a]] && (t >= NB + nrs || (t < NB && W->asleep[W->sbody[t]]))) continue; /* neither box can move */
          v3 d = vsub(W->sc[a], W->sc[t]);
          float m = margin + infl * (W->spd[a] + W->spd[t]) + slack;
          int hit = fabsf(d.x) <= W->sa[a].x + W->sa[t].x + m && fabsf(d.y) <= W->sa[a].y + W->sa[t].y + m &&
                    fabsf(d.z) <= W->sa[a].z + W->sa[t].z + m;
          if (!hit) continue;
          if (t >= NB + nrs) { if (ns < KC) sta[ns++] = (unsigned char)t; ns_all++; }
          else { if (nd < KC) dyn[nd++] = (unsigned char)t; nd_all++; }
        }
        const int kd = nd_all < KC - ns ? nd_all : KC - ns;
        for (int i = 0; i < kd; ++i) W->cand[a][i] = dyn[i];
        for (int i = 0; i < ns; ++i) W->cand[a][kd + i] = sta[i];
        W->ncand[a] = kd + ns;
        W->cand_dropped += (nd_all - kd) + (ns_all - ns);
        W->cand_dropped_static += ns_all - ns;
      }
    }
    if (g_reuse_audit) {   /* test hook: what would a fresh broad phase of THIS sub-step list, and is it in the lists in use? */
      for (int a = 0; a < n_owner; ++a) {
        if (a < NB && a >= nbs) continue;
        for (int t = 0; t < n_target; ++t) {
          if ((t < NB && (t >= nbs || (a < NB && W->sbody[t] == W->sbody[a]))) || (a >= NB && t >= NB && t < NB + nrs)) continue;
          if (a < NB && W->asleep[W->sbody[a]] && (t >= NB + nrs || (t < NB && W->asleep[W->sbody[t]]))) continue;
          v3 d = vsub(W->sc[a], W->sc[t]);
          float m = margin + W->spd[a] + W->spd[t];
          if (!(fabsf(d.x) <= W->sa[a].x + W->sa[t].x + m && fabsf(d.y) <= W->sa[a].y + W->sa[t].y + m &&
                fabsf(d.z) <= W->sa[a].z + W->sa[t].z + m)) continue;
          int found = 0;
          for (int ci = 0; ci < W->ncand[a]; ++ci) if (W->cand[a][ci] == t) found = 1;
          __sync_fetch_and_add(&g_reuse_stats[rebuild ? 0 : 2], 1);                                   /* pairs a fresh sweep finds */
          if (!found) __sync_fetch_and_add(&g_reuse_stats[(rebuild ? 0 : 2) + 1], 1);                 /* ... missing from the lists */
          if (!found && W->ncand[a] < KC) __sync_fetch_and_add(&g_reuse_stats[4], 1);                 /* missing though the owner's list had room */
        }
      }
    }
    /* 4. narrow phase, per ordered pair (owner a, target t): reference face of t = its axis of least
     *    overlap with a (SAT over t's three face axes); a's sample points that lie over that face and
     *    within the speculative margin below/above it become contacts.  Order: owner, candidate, point. */
    /* SHEDDING: contacts with a positive gap are speculative; if the table would overflow, the speculative range is halved (twice),
     * then dropped altogether -- the contacts that go first are the ones that cannot act in this sub-step anyway -- before any
     * touching contact is lost.  Level 0 = the full range m, 1 = m / 2, 2 = m / 4, 3 = touching contacts only. */
    int level = 0; float gs = 1.0f;
  regenerate:
    W->ncon = 0; W->ndropped = 0;
    for (int a = 0; a < n_owner; ++a) {
      int npts = (a < NB && W->sh[a].x > 0.04f) ? 12 : 8; /* long bricks add 4 mid-edge points */
      for (int ci = 0; ci < W->ncand[a]; ++ci) {
        int t = W->cand[a][ci];
        if (a < NB && W->asleep[W->sbody[a]] && (t >= NB + nrs || (t < NB && W->asleep[W->sbody[t]]))) continue; /* kept list, both asleep by now */
        float m = (margin + W->spd[a] + W->spd[t]) * gs;
        v3 lc = mtmul(W->sR[t], vsub(W->sc[a], W->sc[t]));
        float C[9]; /* C = R_t^T R_a */
        for (int r = 0; r < 3; ++r)
          for (int c2 = 0; c2 < 3; ++c2)
            C[3 * r + c2] = W->sR[t][r] * W->sR[a][c2] + W->sR[t][3 + r] * W->sR[a][3 + c2] + W->sR[t][6 + r] * W->sR[a][6 + c2];
        v3 ha = W->sh[a], ht = W->sh[t];
        float o0 = ht.x + (fabsf(C[0]) * ha.x + fabsf(C[1]) * ha.y + fabsf(C[2]) * ha.z) - fabsf(lc.x);
        float o1 = ht.y + (fabsf(C[3]) * ha.x + fabsf(C[4]) * ha.y + fabsf(C[5]) * ha.z) - fabsf(lc.y);
        float o2 = ht.z + (fabsf(C[6]) * ha.x + fabsf(C[7]) * ha.y + fabsf(C[8]) * ha.z) - fabsf(lc.z);
        int k = 0; float ov = o0;
        if (o1 < ov) { k = 1; ov = o1; }
        if (o2 < ov) { k = 2; ov = o2; }
        if (ov < -m) continue;
        float lck = k == 0 ? lc.x : (k == 1 ? lc.y : lc.z);
        float sgf = lck >= 0.0f ? 1.0f : -1.0f;
        uint32_t sg = lck >= 0.0f ? 0u : 1u;
        if (t >= NB + nrs && k == 2) { sgf = 1.0f; sg = 0u; } /* statics rest on each other: their z faces only push UP */
        float htk = k == 0 ? ht.x : (k == 1 ? ht.y : ht.z);
        for (int p = 0; p <= EDGE_POINT; ++p) {
          if (p >= npts && p < EDGE_POINT) continue;
          float depth; v3 wpt; uint32_t word; v3 en = V3(0.0f, 0.0f, 0.0f);
          if (p == EDGE_POINT) { /* the pair's edge-edge contact, once per unordered pair (owner index below target index) */
            if (!(S->edge_contacts > 0.5f && a < t)) continue;
            v3 ep; int er;
            /* speculative range of an edge-edge contact: the travel bounds, plus at most the geometric tolerance (2 mm) of the contact offset --
             * the contact exists to resolve crossings the corner test cannot see, not to anticipate them from 2 cm away (Orient / Search) */
            const float me = ((margin < fmargin ? margin : fmargin) + W->spd[a] + W->spd[t]) * gs;
            if (!edge_contact(C, lc, ha, ht, me, S->edge_pref, &ep, &en, &depth, &er)) continue;
            wpt = vadd(W->sc[t], mmul(W->sR[t], ep));
            en = mmul(W->sR[t], en);
            word = (uint32_t)W->sbody[a] | ((uint32_t)W->sbody[t] << 8) | ((uint32_t)t << 16) | ((uint32_t)er << 24) | EDGE_BIT;
          } else {
          v3 pl;
          if (p < 8) pl = V3((p & 1) ? ha.x : -ha.x, (p & 2) ? ha.y : -ha.y, (p & 4) ? ha.z : -ha.z);
          else pl = V3(0.0f, (p & 1) ? ha.y : -ha.y, (p & 2) ? ha.z : -ha.z);
          v3 l = vadd(lc, mmul(C, pl));
          float lk = k == 0 ? l.x : (k == 1 ? l.y : l.z);
          depth = htk - sgf * lk;
          if (!(depth > -m)) continue;
          int inface = (k == 0 || fabsf(l.x) <= ht.x + fmargin) && (k == 1 || fabsf(l.y) <= ht.y + fmargin) &&
                       (k == 2 || fabsf(l.z) <= ht.z + fmargin);
          if (!inface) continue;
          wpt = vadd(W->sc[a], mmul(W->sR[a], pl));
          word = (uint32_t)W->sbody[a] | ((uint32_t)W->sbody[t] << 8) | ((uint32_t)t << 16) | ((uint32_t)k << 24) | (sg << 26);
          }
          if (W->ncon >= SDX_MAX_CONTACTS) { W->ndropped++; continue; }
          contact_t* c = &W->con[W->ncon++];
          c->word = word;
          c->n[0] = en.x; c->n[1] = en.y; c->n[2] = en.z;
          c->w[0] = wpt.x; c->w[1] = wpt.y; c->w[2] = wpt.z;
          float bias = 0.0f;
          if (depth > S->slop) { bias = S->baumgarte * (depth - S->slop) / h; if (bias > S->max_depen_vel) bias = S->max_depen_vel; }
          else if (depth < 0.0f) bias = depth / h;
          c->bias = bias;
          c->f[0] = c->f[1] = c->f[2] = 0.0f;
          {   /* warm start: a contact is identified by (owner shape, target shape, sample point); keys ascend with the contact order */
            uint32_t key = ((uint32_t)a << 12) | ((uint32_t)t << 4) | (uint32_t)p;
            W->ckey[W->ncon - 1] = key;
            if (S->warm_start > 0.0f) {
              int lo = 0, hi = nprev - 1;
              while (lo <= hi) {
                int mid = (lo + hi) >> 1;
                union { float f; uint32_t u; } kv; kv.f = wsr[4 * mid];
                if (kv.u < key) lo = mid + 1;
                else if (kv.u > key) hi = mid - 1;
                else {
                  /* a contact that involves a robot link or a HOT brick (hit by the robot / faster than the wake threshold in the last
                   * sub-step) moves too fast for its last impulse to be trusted as far as a resting contact's */
                  const int hotc = a >= NB || W->hot[W->sbody[a]] || (t < NB ? W->hot[W->sbody[t]] : (t < NB + nrs));
                  const float wf = hotc ? S->warm_start_hot : S->warm_start;
                  c->f[0] = wf * wsr[4 * mid + 1]; c->f[1] = wf * wsr[4 * mid + 2]; c->f[2] = wf * wsr[4 * mid + 3]; break;
                }
              }
            }
          }
          c->inv[0] = depth; /* scratch: kept for the debug dump, overwritten below */
        }
      }
    }
    if (W->ndropped > 0 && level < 3) { ++level; gs = level == 3 ? 0.0f : gs * 0.5f; goto regenerate; }
    if (level > W->shed_level) W->shed_level = level;
    /* 5. incidence + mass-splitting counts.  Contacts are generated owner-major, so the contacts a body OWNS
     *    are one contiguous range; the contacts in which it is the TARGET are listed in ascending order. */
    for (int b = 0; b < NBODY; ++b) { W->astart[b] = 0; W->aend[b] = 0; W->nb[b] = 0; }
    for (int i = 0; i < W->ncon; ++i) {
      int a = W->con[i].word & 255, b = (W->con[i].word >> 8) & 255;
      if (i == 0 || (int)(W->con[i - 1].word & 255) != a) W->astart[a] = i;
      W->aend[a] = i + 1;
      if (b != STATIC_BODY) W->nb[b]++;
      if (a < NB && b != STATIC_BODY) { if (b >= NB) W->touch[a] |= 1; else if (W->hot[b]) W->touch[a] |= 2; }
      if (b < NB) { if (a >= NB) W->touch[b] |= 1; else if (W->hot[a]) W->touch[b] |= 2; }
    }
    W->boff[0] = 0;
    for (int b = 0; b < NBODY; ++b) W->boff[b + 1] = W->boff[b] + W->nb[b];
    for (int b = 0; b < NBODY; ++b) {
      int o = W->boff[b];
      for (int i = 0; i < W->ncon; ++i) if ((int)((W->con[i].word >> 8) & 255) == b) W->blist[o++] = (unsigned short)i;
      W->nb[b] = (W->aend[b] - W->astart[b]) + (W->boff[b + 1] - W->boff[b]);
    }
    for (int j = 0; j < SDX_ND; ++j) {
      int n = 0;
      for (int L = 0; L < SDX_NL; ++L) if (S->link_anc_mask[L] & (1u << j)) n += W->nb[NB + L];
      W->nj[j] = n;
    }
    if (condump && sub == S->substeps - 1)
      for (int i = 0; i < W->ncon; ++i) {
        float* o = condump + 8 * i;
        union { uint32_t u; float f; } cv; cv.u = W->con[i].word;
        o[0] = cv.f; o[1] = W->con[i].w[0]; o[2] = W->con[i].w[1]; o[3] = W->con[i].w[2]; o[4] = W->con[i].inv[0];
        o[5] = W->con[i].bias; o[6] = 0.0f; o[7] = 0.0f;
      }
    for (int i = 0; i < W->ncon; ++i) {
      contact_t* c = &W->con[i];
      int a = c->word & 255, b = (c->word >> 8) & 255;
      v3 n, t1, t2; contact_axes(W, c, &n, &t1, &t2);
      v3 wpt = V3(c->w[0], c->w[1], c->w[2]);
      c->inv[0] = 1.0f / (body_k(S, W, a, wpt, n) + body_k(S, W, b, wpt, n));
      c->inv[1] = 1.0f / (body_k(S, W, a, wpt, t1) + body_k(S, W, b, wpt, t1));
      c->inv[2] = 1.0f / (body_k(S, W, a, wpt, t2) + body_k(S, W, b, wpt, t2));
    }
    /* 6. mass-splitting Jacobi iterations on the total impulses */
    for (int it = -1; it < S->iters; ++it) { /* it = -1: only phase B, i.e. apply the warm-start impulses */
      if (it >= 0) for (int i = 0; i < W->ncon; ++i) { /* phase A: one contact each */
        contact_t* c = &W->con[i];
        int a = c->word & 255, b = (c->word >> 8) & 255;
        v3 n, t1, t2; contact_axes(W, c, &n, &t1, &t2);
        v3 wpt = V3(c->w[0], c->w[1], c->w[2]);
        v3 f = V3(c->f[0], c->f[1], c->f[2]);
        v3 vrel = vadd(W->bv[a], vcross(W->bw[a], vsub(wpt, W->bx[a])));
        if (b != STATIC_BODY) vrel = vsub(vrel, vadd(W->bv[b], vcross(W->bw[b], vsub(wpt, W->bx[b]))));
        float ln = fmaf(c->bias - vdot(vrel, n), c->inv[0], vdot(f, n));
        ln = ln > 0.0f ? ln : 0.0f;
        float lim = S->friction * ln;
        float l1 = clampf(fmaf(-vdot(vrel, t1), c->inv[1], vdot(f, t1)), -lim, lim);
        float l2 = clampf(fmaf(-vdot(vrel, t2), c->inv[2], vdot(f, t2)), -lim, lim);
        f = vmad(t2, l2, vmad(t1, l1, vscale(n, ln)));
        c->f[0] = f.x; c->f[1] = f.y; c->f[2] = f.z;
      }
      for (int b = 0; b < NBODY; ++b) { /* phase B: one body each.  The incidences of a body are dealt round-robin to L partial sums
                                         * (L = phaseb_lanes: a power of two, so that no partial holds more than 4 of them up to 128; links: 2)
                                         * which are then combined by a butterfly: p[k] += p[k ^ o] for o = 1, 2, 4 ...  -- the order a group
                                         * of L lanes produces with xor-shuffles, whatever lanes of whatever warp the group sits on */
        int na = W->aend[b] - W->astart[b], nbl = W->boff[b + 1] - W->boff[b];
        if (b < NB && (W->asleep[b] || na + nbl == 0)) continue; /* asleep: stays at rest; untouched: keeps its free velocity */
        const int L = b < NB ? phaseb_lanes(na + nbl) : 2;
        v3 Fk[32], Tk[32];
        for (int k = 0; k < L; ++k) { Fk[k] = V3(0, 0, 0); Tk[k] = V3(0, 0, 0); }
        v3 xb = b < NB ? W->bx[b] : V3(0, 0, 0);
        for (int e = 0; e < na + nbl; ++e) {
          int i = e < na ? W->astart[b] + e : W->blist[W->boff[b] + (e - na)];
          const contact_t* c = &W->con[i];
          v3 f = V3(c->f[0], c->f[1], c->f[2]);
          if (e >= na) f = vneg(f);
          v3 wpt = V3(c->w[0], c->w[1], c->w[2]);
          int k = e & (L - 1);
          Fk[k] = vadd(Fk[k], f);
          Tk[k] = vadd(Tk[k], vcross(vsub(wpt, xb), f));
        }
        for (int o = 1; o < L; o <<= 1) {
          v3 Fn[32], Tn[32];
          for (int k = 0; k < L; ++k) { Fn[k] = vadd(Fk[k], Fk[k ^ o]); Tn[k] = vadd(Tk[k], Tk[k ^ o]); }
          for (int k = 0; k < L; ++k) { Fk[k] = Fn[k]; Tk[k] = Tn[k]; }
        }
        v3 F = Fk[0];
        v3 T = Tk[0];
        if (b < NB) {
          W->bv[b] = vmad(F, S->br_invm[b], W->vfree[b]);
          W->bw[b] = vadd(W->wfree[b], brick_Iinv_mul(W, b, T));
        } else { W->linkF[b - NB] = F; W->linkM[b - NB] = T; }
      }
      for (int j = 0; j < SDX_ND; ++j) {
        v3 Fd = V3(0, 0, 0), Md = V3(0, 0, 0);
        for (int L = 0; L < SDX_NL; ++L)
          if ((S->link_anc_mask[L] & (1u << j)) && W->nb[NB + L] > 0) { Fd = vadd(Fd, W->linkF[L]); Md = vadd(Md, W->linkM[L]); } /* links in contact only */
        float g = vdot(W->K.ja[j], vsub(Md, vcross(W->K.jo[j], Fd)));
        W->qd[j] = W->qdfree[j] + g / W->ieff[j];
      }
      link_twists(S, W);
    }
    for (int i = 0; i < W->ncon; ++i) {
      union { float f; uint32_t u; } kv; kv.u = W->ckey[i];
      wsw[4 * i] = kv.f; wsw[4 * i + 1] = W->con[i].f[0]; wsw[4 * i + 2] = W->con[i].f[1]; wsw[4 * i + 3] = W->con[i].f[2];
    }
    wsn[1 - rb] = W->ncon;
    rb = 1 - rb;
    if (S->iters == 0) for (int L = 0; L < SDX_NL; ++L) { W->linkF[L] = V3(0, 0, 0); W->linkM[L] = V3(0, 0, 0); }
    /* 7. integrate */
    for (int b = 0; b < nbr; ++b) {
      if (W->asleep[b]) {   /* pose and (zero) velocity unchanged; only a touch can restart the counter */
        if (W->touch[b] & 1) slp[b] = 0; else if (W->touch[b] & 2) slp[b] = 1;
        continue;
      }
      v3 w = W->bw[b], v = W->bv[b];
      float w2 = vdot(w, w), mw = S->max_ang_vel;
      if (w2 > mw * mw) { w = vscale(w, mw / sqrtf(w2)); W->bw[b] = w; }
      float v2 = vdot(v, v), mv = S->max_lin_vel;
      if (v2 > mv * mv) { v = vscale(v, mv / sqrtf(v2)); W->bv[b] = v; }
      {
        float E = 0.5f * (vdot(v, v) + vdot(w, w) * (W->brad[b] * W->brad[b] * (1.0f / 3.0f)));
        int c = slp[b];
        if ((W->touch[b] & 1) || E >= S->wake_energy) c = 0;
        else if ((W->touch[b] & 2) || E >= S->sleep_energy) c = 1;
        else c = c + 1 > 255 ? 255 : c + 1;
        slp[b] = (unsigned char)c;
      }
      W->bx[b] = vmad(v, h, W->bx[b]);
      q4 q = W->bq[b];
      float hh = 0.5f * h;
      q4 dq;
      dq.x = hh * (w.x * q.w + w.y * q.z - w.z * q.y);
      dq.y = hh * (w.y * q.w + w.z * q.x - w.x * q.z);
      dq.z = hh * (w.z * q.w + w.x * q.y - w.y * q.x);
      dq.w = hh * (-(w.x * q.x + w.y * q.y + w.z * q.z));
      q.x = q.x + dq.x; q.y = q.y + dq.y; q.z = q.z + dq.z; q.w = q.w + dq.w;
      float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
      q.x = q.x * inv; q.y = q.y * inv; q.z = q.z * inv; q.w = q.w * inv;
      W->bq[b] = q;
    }
    for (int j = 0; j < SDX_ND; ++j) {
      float qn = W->q[j] + h * W->qd[j];
      if (qn < S->dof_lo[j]) { qn = S->dof_lo[j]; if (W->qd[j] < 0.0f) W->qd[j] = 0.0f; }
      if (qn > S->dof_hi[j]) { qn = S->dof_hi[j]; if (W->qd[j] > 0.0f) W->qd[j] = 0.0f; }
      W->q[j] = qn;
    }
  }
  /* epilogue: state write-back + the rows the task reads */
  for (int b = 0; b < NB; ++b) {
    brick[0 * NB + b] = W->bx[b].x; brick[1 * NB + b] = W->bx[b].y; brick[2 * NB + b] = W->bx[b].z;
    brick[3 * NB + b] = W->bq[b].x; brick[4 * NB + b] = W->bq[b].y; brick[5 * NB + b] = W->bq[b].z; brick[6 * NB + b] = W->bq[b].w;
    brick[7 * NB + b] = W->bv[b].x; brick[8 * NB + b] = W->bv[b].y; brick[9 * NB + b] = W->bv[b].z;
    brick[10 * NB + b] = W->bw[b].x; brick[11 * NB + b] = W->bw[b].y; brick[12 * NB + b] = W->bw[b].z;
  }
  for (int j = 0; j < SDX_ND; ++j) { dof[j] = W->q[j]; dof[24 + j] = W->qd[j]; }
  robot_fk(S, W->q, &W->K);
  link_twists(S, W);
  for (int L = 0; L < SDX_NL; ++L) {
    float* o = link_out + 13 * L;
    o[0] = W->K.lx[L].x; o[1] = W->K.lx[L].y; o[2] = W->K.lx[L].z;
    o[3] = W->K.lq[L].x; o[4] = W->K.lq[L].y; o[5] = W->K.lq[L].z; o[6] = W->K.lq[L].w;
    o[7] = W->bv[NB + L].x; o[8] = W->bv[NB + L].y; o[9] = W->bv[NB + L].z;
    o[10] = W->bw[NB + L].x; o[11] = W->bw[NB + L].y; o[12] = W->bw[NB + L].z;
    float invh = 1.0f / h;
    netf[3 * L] = W->linkF[L].x * invh; netf[3 * L + 1] = W->linkF[L].y * invh; netf[3 * L + 2] = W->linkF[L].z * invh;
  }
  for (int j = 0; j < 7; ++j) {
    v3 lin = vcross(W->K.ja[j], vsub(W->K.lx[7], W->K.jo[j]));
    jac7[0 * 7 + j] = lin.x; jac7[1 * 7 + j] = lin.y; jac7[2 * 7 + j] = lin.z;
    jac7[3 * 7 + j] = W->K.ja[j].x; jac7[4 * 7 + j] = W->K.ja[j].y; jac7[5 * 7 + j] = W->K.ja[j].z;
  }
  ncontact[0] = W->ncon; ncontact[1] = W->ndropped; ncontact[2] = W->shed_level; ncontact[3] = W->cand_dropped | (W->cand_dropped_static << 16);
}

typedef struct {
  const sdx_scene_t* S; int n; float *brick, *dof, *link, *jac7, *netf; int* ncontact; float* condump; float* ws; int* wsn; int ws_cur; unsigned char* slp;
  int tid, nthreads;
} sim_job_t;
static void* sim_worker(void* arg) {
  sim_job_t* J = (sim_job_t*)arg;
  work_t* W = (work_t*)malloc(sizeof(work_t));
  for (int e = J->tid; e < J->n; e += J->nthreads)
    sim_env(J->S, J->brick + (size_t)e * 13 * NB, J->dof + (size_t)e * 72, J->link + (size_t)e * SDX_NL * 13,
            J->jac7 + (size_t)e * 42, J->netf + (size_t)e * SDX_NL * 3, J->ncontact + 4 * e,
            J->condump ? J->condump + (size_t)e * SDX_MAX_CONTACTS * 8 : 0, J->ws + (size_t)e * 2 * SDX_MAX_CONTACTS * 4, J->wsn + 2 * e,
            J->ws_cur, J->slp + (size_t)e * NB, W, e);
  free(W);
  return 0;
}
static int g_threads = 0;
void sdxo_set_threads(int t) { g_threads = t; }
int sdxo_get_threads(void) {
  if (g_threads > 0) return g_threads;
  long c = sysconf(_SC_NPROCESSORS_ONLN);
  return c > 0 ? (int)c : 1;
}
/* envs are independent: static round-robin over host threads (results do not depend on the thread count) */
void sdxo_simulate(const sdx_scene_t* S, int n, float* brick, float* dof, float* link, float* jac7, float* netf,
                   int* ncontact, float* condump, float* ws, int* wsn, int ws_cur, unsigned char* slp) {
  int nt = sdxo_get_threads();
  if (nt > n) nt = n;
  if (nt < 1) nt = 1;
  if (nt > 256) nt = 256;
  pthread_t th[256]; sim_job_t jobs[256];
  for (int t = 0; t < nt; ++t) {
    sim_job_t j = {S, n, brick, dof, link, jac7, netf, ncontact, condump, ws, wsn, ws_cur, slp, t, nt};
    jobs[t] = j;
    if (t > 0) pthread_create(&th[t], 0, sim_worker, &jobs[t]);
  }
  sim_worker(&jobs[0]);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], 0);
}

/* t-value training data (GS:1402-1438, save_hdf5): every env that resets (after the first step) contributes its
 * camera-frame target quaternion -- states[177:181] of the newest frame, the gate's input (GS:1200) -- to the SUCCESS set
 * when the grasp is banked (target y < 0, finger_dist < 0.6, tvalue > 0.8) and to the FAILURE set otherwise; env order,
 * rings of `cap` rows, counts[0/1] = rows ever written (the reference's success_v_count / failure_v_count). */
void sdxo_tv_dataset(const sdx_scene_t* S, int n, const int64_t* reset, const float* brick, const float* finger_dist,
                     const float* tvalue, const float* states, float* succ, float* fail, int64_t* counts, int cap);

/* FK-only refresh of link rows / jacobian (used after resets and at creation) */
void sdxo_refresh_links(const sdx_scene_t* S, int n, const float* dof, float* link, float* jac7) {
  for (int e = 0; e < n; ++e) {
    work_t* W = (work_t*)malloc(sizeof(work_t));
    const float* d = dof + (size_t)e * 72;
    for (int j = 0; j < SDX_ND; ++j) { W->q[j] = d[j]; W->qd[j] = d[24 + j]; }
    robot_fk(S, W->q, &W->K);
    link_twists(S, W);
    for (int L = 0; L < SDX_NL; ++L) {
      float* o = link + (size_t)e * SDX_NL * 13 + 13 * L;
      o[0] = W->K.lx[L].x; o[1] = W->K.lx[L].y; o[2] = W->K.lx[L].z;
      o[3] = W->K.lq[L].x; o[4] = W->K.lq[L].y; o[5] = W->K.lq[L].z; o[6] = W->K.lq[L].w;
      o[7] = W->bv[NB + L].x; o[8] = W->bv[NB + L].y; o[9] = W->bv[NB + L].z;
      o[10] = W->bw[NB + L].x; o[11] = W->bw[NB + L].y; o[12] = W->bw[NB + L].z;
    }
    float* J = jac7 + (size_t)e * 42;
    for (int j = 0; j < 7; ++j) {
      v3 lin = vcross(W->K.ja[j], vsub(W->K.lx[7], W->K.jo[j]));
      J[0 * 7 + j] = lin.x; J[1 * 7 + j] = lin.y; J[2 * 7 + j] = lin.z;
      J[3 * 7 + j] = W->K.ja[j].x; J[4 * 7 + j] = W->K.ja[j].y; J[5 * 7 + j] = W->K.ja[j].z;
    }
    free(W);
  }
}

/* ------------------------------------------------------------------ task ops */
static inline int target_brick(int env) { int s = env % 8; return (s == 3 || s == 4 || s == 7) ? 0 : s; } /* GS:962-975 */

/* COM-frame brick block -> Isaac-Gym root row of brick b (pos = COM - R*coff, v_root = v + w x (root-COM)) */
static void brick_root_row(const sdx_scene_t* S, const float* brick, int b, float* row) {
  v3 x = V3(brick[0 * NB + b], brick[1 * NB + b], brick[2 * NB + b]);
  q4 q = {brick[3 * NB + b], brick[4 * NB + b], brick[5 * NB + b], brick[6 * NB + b]};
  v3 v = V3(brick[7 * NB + b], brick[8 * NB + b], brick[9 * NB + b]);
  v3 w = V3(brick[10 * NB + b], brick[11 * NB + b], brick[12 * NB + b]);
  v3 off = qrot(q, V3(S->br_coff[3 * b], S->br_coff[3 * b + 1], S->br_coff[3 * b + 2]));
  v3 p = vsub(x, off);
  v3 vr = vsub(v, vcross(w, off));
  row[0] = p.x; row[1] = p.y; row[2] = p.z; row[3] = q.x; row[4] = q.y; row[5] = q.z; row[6] = q.w;
  row[7] = vr.x; row[8] = vr.y; row[9] = vr.z; row[10] = w.x; row[11] = w.y; row[12] = w.z;
}
static void brick_from_root_row(const sdx_scene_t* S, float* brick, int b, const float* row) {
  q4 q = {row[3], row[4], row[5], row[6]};
  v3 off = qrot(q, V3(S->br_coff[3 * b], S->br_coff[3 * b + 1], S->br_coff[3 * b + 2]));
  v3 w = V3(row[10], row[11], row[12]);
  v3 x = vadd(V3(row[0], row[1], row[2]), off);
  v3 v = vadd(V3(row[7], row[8], row[9]), vcross(w, off));
  brick[0 * NB + b] = x.x; brick[1 * NB + b] = x.y; brick[2 * NB + b] = x.z;
  brick[3 * NB + b] = q.x; brick[4 * NB + b] = q.y; brick[5 * NB + b] = q.z; brick[6 * NB + b] = q.w;
  brick[7 * NB + b] = v.x; brick[8 * NB + b] = v.y; brick[9 * NB + b] = v.z;
  brick[10 * NB + b] = w.x; brick[11 * NB + b] = w.y; brick[12 * NB + b] = w.z;
}
void sdxo_brick_root_rows(const sdx_scene_t* S, int n, const float* brick, float* rows /* [n][72][13] */) {
  for (int e = 0; e < n; ++e)
    for (int b = 0; b < NB; ++b) brick_root_row(S, brick + (size_t)e * 13 * NB, b, rows + ((size_t)e * NB + b) * 13);
}
void sdxo_brick_from_root_rows(const sdx_scene_t* S, int n, float* brick, const float* rows) {
  for (int e = 0; e < n; ++e)
    for (int b = 0; b < NB; ++b) brick_from_root_row(S, brick + (size_t)e * 13 * NB, b, rows + ((size_t)e * NB + b) * 13);
}

/* GraspInsertTValue forward (TVF:30-46) + sigmoid(.)[1] (GS:1200-1201).
 * weights: W1[256][4] b1[256] W2[128][256] b2 W3[64][128] b3 W4[2][64] b4 */
float sdxo_tvalue_one(const float* wts, const float* qin) {
  const float* W1 = wts; const float* b1 = W1 + 256 * 4; const float* W2 = b1 + 256; const float* b2 = W2 + 128 * 256;
  const float* W3 = b2 + 128; const float* b3 = W3 + 64 * 128; const float* W4 = b3 + 64; const float* b4 = W4 + 2 * 64;
  float h1[256], h2[128], h3[64];
  for (int o = 0; o < 256; ++o) { float a = b1[o]; for (int k = 0; k < 4; ++k) a = fmaf(W1[o * 4 + k], qin[k], a); h1[o] = sdx_elu(a); }   /* one fused multiply-add per weight, as the kernel */
  for (int o = 0; o < 128; ++o) { float a = b2[o]; for (int k = 0; k < 256; ++k) a = fmaf(W2[o * 256 + k], h1[k], a); h2[o] = sdx_elu(a); }
  for (int o = 0; o < 64; ++o) { float a = b3[o]; for (int k = 0; k < 128; ++k) a = fmaf(W3[o * 128 + k], h2[k], a); h3[o] = sdx_elu(a); }
  float a = b4[1]; for (int k = 0; k < 64; ++k) a = fmaf(W4[64 + k], h3[k], a);
  a = sdx_elu(a);
  return 1.0f / (1.0f + sdx_exp(-a));
}
void sdxo_tvalue(const float* wts, int n, const float* qin, float* out) { for (int e = 0; e < n; ++e) out[e] = sdxo_tvalue_one(wts, qin + 4 * e); }

/* control_ik (GS:1796-1804): u = J^T (J J^T + 0.05^2 I)^-1 dpose, 6x7 J, solved by Cholesky */
static void control_ik(const float* J /*[6][7]*/, const float* dpose, float* u /*[7]*/) {
  float A[6][6], y[6];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) {
      float s = 0.0f;
      for (int k = 0; k < 7; ++k) s = s + J[r * 7 + k] * J[c * 7 + k];
      if (r == c) s = s + 0.05f * 0.05f;
      A[r][c] = s;
    }
  for (int c = 0; c < 6; ++c) { /* A = L L^T, in place (lower) */
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
void sdxo_control_ik(int n, const float* J, const float* dpose, float* u) { for (int e = 0; e < n; ++e) control_ik(J + 42 * e, dpose + 6 * e, u + 7 * e); }

/* reset_idx (GS:1361-1553) for the envs whose reset_buf is set.  Dead writes of the reference are
 * not restated (DESIGN.md "reset"): the randomised target pose (GS:1488-1499) is overwritten by the
 * banked heap row (GS:1508-1511) and perturb_* (GS:1460-1461) feed a disabled branch. */
void sdxo_reset(const sdx_scene_t* S, int n, uint64_t seed, const float* bank, int per_type, float* brick, float* dof,
                float* target_init, int64_t* progress, int64_t* reset, float* successes, int* episode, int* wsn, unsigned char* slp,
                /* grasp terminal-state banking (GS:1399-1445) */
                int do_bank, const float* finger_dist, const float* tvalue, float* gb_hand, float* gb_obj, int* gb_index) {
  if (do_bank) {
    for (int e = 0; e < n; ++e) {
      if (!reset[e]) continue;
      int ty = e % 8, tb = target_brick(e);
      float row[13];
      float* B = brick + (size_t)e * 13 * NB;
      brick_root_row(S, B, tb, row);
      if (row[1] < 0.0f && finger_dist[e] < 0.6f && tvalue[e] > 0.8f) {
        int slot = gb_index[ty];
        float* hd = gb_hand + ((size_t)ty * 11024 + slot) * 46;
        const float* d = dof + (size_t)e * 72;
        for (int j = 0; j < SDX_ND; ++j) { hd[2 * j] = d[j]; hd[2 * j + 1] = d[24 + j]; }
        float* ob = gb_obj + ((size_t)ty * 11024 + slot) * 13;
        for (int k = 0; k < 13; ++k) ob[k] = row[k];
        gb_index[ty] = slot + 1;
      }
      if (gb_index[ty] > 5000) gb_index[ty] = 0;
    }
  }
  for (int e = 0; e < n; ++e) {
    if (!reset[e]) continue;
    float* B = brick + (size_t)e * 13 * NB;
    float* d = dof + (size_t)e * 72;
    uint32_t r[4];
    philox(seed, (uint32_t)e, (uint32_t)episode[e], 1u, r);
    int slot = (int)(r[0] % (uint32_t)per_type);
    const float* rows = bank + (((size_t)(e % 8)) * per_type + slot) * NB * 13;
    for (int b = 0; b < NB; ++b) {
      float row[13];
      for (int k = 0; k < 7; ++k) row[k] = rows[b * 13 + k];
      for (int k = 7; k < 13; ++k) row[k] = 0.0f; /* GS:1513 */
      brick_from_root_row(S, B, b, row);
    }
    for (int j = 0; j < 7; ++j) { d[j] = S->prepare_arm[j]; d[24 + j] = 0.0f; d[48 + j] = S->prepare_arm[j]; }
    for (int i = 0; i < 16; ++i) {
      float v = scalef(S->finger_reset_unscaled[i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
      d[7 + i] = v; d[24 + 7 + i] = 0.0f; d[48 + 7 + i] = v;
    }
    int tb = target_brick(e);
    for (int k = 0; k < 7; ++k) target_init[7 * e + k] = rows[tb * 13 + k]; /* GS:1547-1548 */
    progress[e] = 0; reset[e] = 0; successes[e] = 0.0f; /* GS:1550-1552 */
    wsn[2 * e] = 0; wsn[2 * e + 1] = 0; /* a new heap: no contact persists */
    for (int b = 0; b < NB; ++b) slp[(size_t)e * NB + b] = 0; /* setting a pose wakes the actor */
    episode[e] += 1;
  }
}

/* pre_physics_step after resets (GS:1570-1638): actions -> DoF position targets */
void sdxo_pre_physics(const sdx_scene_t* S, int n, const float* actions_in, float* actions, float* dof, const float* link,
                      const float* jac7, const int64_t* progress, const float* target_init) {
  for (int e = 0; e < n; ++e) {
    const float* a = actions_in + 23 * e;
    float* d = dof + (size_t)e * 72;
    float cur[23];
    for (int k = 0; k < 23; ++k) actions[23 * e + k] = a[k];
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
    control_ik(jac7 + 42 * (size_t)e, dpose, u);
    for (int j = 0; j < 7; ++j) cur[j] = d[j] + u[j];
    if (pg > 100) for (int j = 0; j < 7; ++j) cur[j] = S->insert_prep0[j];
    if (pg > 125) for (int j = 0; j < 7; ++j) cur[j] = S->insert_prep1[j];
    if (pg > 75) for (int i = 7; i < 23; ++i) cur[i] = d[48 + i];
    for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
  }
}

/* post_physics_step (GS:1640-1645): progress += 1, compute_observations (GS:1090-1332),
 * compute_hand_reward (GS:1706-1776).  obs [n][396], states [n][564] (UNCLAMPED task buffers). */
void sdxo_post_physics(const sdx_scene_t* S, int n, const float* tv_wts, const float* brick, const float* dof,
                       const float* link, const float* actions, const float* target_init, int64_t* progress,
                       int64_t* reset, float* obs, float* states, float* rew, float* tvalue, float* finger_dist_out,
                       const float* successes, float* consec) {
  int64_t num_resets = 0; float finished = 0.0f;
  for (int e = 0; e < n; ++e) {
    progress[e] += 1;
    const float* L = link + (size_t)e * SDX_NL * 13;
    const float* d = dof + (size_t)e * 72;
    const float* hb = L + 7 * 13; /* hand base = panda_link7 (GS:355,1110) */
    const float* ff = L + 11 * 13; const float* mf = L + 19 * 13; const float* rf = L + 23 * 13; const float* th = L + 15 * 13; /* GS:183-186,1130-1152 */
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
    v3 tp = V3(tg[0], tg[1], tg[2]); q4 tq = {tg[3], tg[4], tg[5], tg[6]};
    v3 tip[4]; const float* fs[4] = {ff, mf, rf, th};
    for (int i = 0; i < 4; ++i) { /* GS:1154-1157 */
      q4 fq = {fs[i][3], fs[i][4], fs[i][5], fs[i][6]};
      tip[i] = vadd(V3(fs[i][0], fs[i][1], fs[i][2]), qrot(fq, V3(0.0f, 0.0f, 1.0f * 0.04f)));
    }
    float nrm[4];
    for (int i = 0; i < 4; ++i) { v3 dd = vsub(tp, tip[i]); nrm[i] = sqrtf(vdot(dd, dd)); }
    float fdist = nrm[0] + nrm[1] + nrm[2] + nrm[3]; /* GS:1164-1165 (ff, mf, rf, th) */
    finger_dist_out[e] = fdist;
    /* hand in robot-base frame (GS:1172-1173) */
    q4 bq = {S->base_quat[0], S->base_quat[1], S->base_quat[2], S->base_quat[3]};
    q4 bqi = qconj(bq); v3 bpi = vneg(qrot(bqi, V3(S->base_pos[0], S->base_pos[1], S->base_pos[2])));
    q4 hq = {hb[3], hb[4], hb[5], hb[6]}; v3 hp = V3(hb[0], hb[1], hb[2]);
    q4 hvq = qmul(bqi, hq); v3 hvp = vadd(qrot(bqi, hp), bpi);
    /* target in wrist-camera frame (GS:1176-1182) */
    q4 cq0 = {S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]};
    q4 cq = qmul(hq, cq0); v3 cp = vadd(qrot(hq, V3(S->cam_off_pos[0], S->cam_off_pos[1], S->cam_off_pos[2])), hp);
    q4 cqi = qconj(cq); v3 cpi = vneg(qrot(cqi, cp));
    q4 cvq = qmul(cqi, tq); v3 cvp = vadd(qrot(cqi, tp), cpi);
    float qin[4] = {cvq.x, cvq.y, cvq.z, cvq.w};
    float tv = sdxo_tvalue_one(tv_wts, qin);
    tvalue[e] = tv;
    const float* ti = target_init + 7 * e;
    /* ---- obs frame (GS:1299-1332) */
    float* o = obs + (size_t)e * 3 * OBS_FRAME;
    for (int k = 2 * OBS_FRAME - 1; k >= 0; --k) o[OBS_FRAME + k] = o[k]; /* history shift GS:1330-1332 */
    for (int i = 0; i < 16; ++i) o[i] = unscalef(d[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
    o[16] = hvp.x; o[17] = hvp.y; o[18] = hvp.z; o[19] = hvq.x; o[20] = hvq.y; o[21] = hvq.z; o[22] = hvq.w;
    o[23] = cvp.x; o[24] = cvp.y; o[25] = cvp.z; o[26] = cvq.x; o[27] = cvq.y; o[28] = cvq.z; o[29] = cvq.w;
    for (int i = 0; i < 16; ++i) o[30 + i] = S->vel_obs_scale * d[24 + 7 + i];
    for (int k = 0; k < 13; ++k) { o[46 + k] = ff[k]; o[59 + k] = rf[k]; o[72 + k] = mf[k]; o[85 + k] = th[k]; o[98 + k] = tg[k]; }
    for (int k = 0; k < 7; ++k) o[111 + k] = hb[k];
    for (int k = 0; k < 7; ++k) o[118 + k] = ti[k];
    o[125] = tp.x - ti[0]; o[126] = tp.y - ti[1]; o[127] = tp.z - ti[2];
    o[128] = hp.x - tp.x; o[129] = hp.y - tp.y; o[130] = hp.z - tp.z;
    /* o[131] is never written by the reference (GS:1328 commented out): it keeps its value */
    /* ---- privileged state frame (GS:1220-1280) */
    float* s = states + (size_t)e * 3 * STATE_FRAME;
    for (int k = 2 * STATE_FRAME - 1; k >= 0; --k) s[STATE_FRAME + k] = s[k];
    for (int j = 0; j < 23; ++j) { s[j] = unscalef(d[j], S->dof_lo[j], S->dof_hi[j]); s[23 + j] = S->vel_obs_scale * d[24 + j]; }
    s[46] = tip[0].x; s[47] = tip[0].y; s[48] = tip[0].z; /* ff */
    s[49] = tip[2].x; s[50] = tip[2].y; s[51] = tip[2].z; /* rf */
    s[52] = tip[1].x; s[53] = tip[1].y; s[54] = tip[1].z; /* mf */
    s[55] = tip[3].x; s[56] = tip[3].y; s[57] = tip[3].z; /* th */
    for (int k = 0; k < 23; ++k) s[58 + k] = actions[23 * e + k];
    for (int k = 0; k < 7; ++k) { s[81 + k] = hb[k]; s[88 + k] = tg[k]; }
    for (int k = 0; k < 6; ++k) s[95 + k] = hb[7 + k];
    for (int k = 0; k < 4; ++k) { s[101 + k] = ff[3 + k]; s[111 + k] = mf[3 + k]; s[121 + k] = rf[3 + k]; s[131 + k] = th[3 + k]; }
    for (int k = 0; k < 6; ++k) { s[105 + k] = ff[7 + k]; s[115 + k] = mf[7 + k]; s[125 + k] = rf[7 + k]; s[135 + k] = th[7 + k]; }
    /* s[141] likewise untouched (GS:1253-1255) */
    for (int k = 0; k < 6; ++k) s[142 + k] = tg[7 + k];
    s[148] = ti[0]; s[149] = ti[1]; s[150] = ti[2];
    s[151] = tp.x - ti[0]; s[152] = tp.y - ti[1]; s[153] = tp.z - ti[2];
    s[154] = hp.x - tp.x; s[155] = hp.y - tp.y; s[156] = hp.z - tp.z;
    q4 rel = qmul(hq, qconj(tq));
    s[157] = rel.x; s[158] = rel.y; s[159] = rel.z; s[160] = rel.w;
    { v3 a = vsub(tp, tip[0]), b = vsub(tp, tip[2]), c = vsub(tp, tip[1]), dd = vsub(tp, tip[3]);
      s[161] = a.x; s[162] = a.y; s[163] = a.z; s[164] = b.x; s[165] = b.y; s[166] = b.z;
      s[167] = c.x; s[168] = c.y; s[169] = c.z; s[170] = dd.x; s[171] = dd.y; s[172] = dd.z; }
    s[173] = fdist;
    s[174] = cvp.x; s[175] = cvp.y; s[176] = cvp.z; s[177] = cvq.x; s[178] = cvq.y; s[179] = cvq.z; s[180] = cvq.w;
    s[181] = cvp.x; s[182] = cvp.y; s[183] = cvp.z; s[184] = cvq.x; s[185] = cvq.y; s[186] = cvq.z; s[187] = cvq.w;
    /* ---- reward / reset (GS:1719-1755) */
    float dist = nrm[0] + nrm[1] + nrm[2] + 3.0f * nrm[3];
    int64_t rs = reset[e];
    if (dist <= -1.0f) rs = 1;
    if ((float)progress[e] >= (float)S->max_episode_length - 1.0f) rs = 1;
    float cl = dist - 0.5f; if (cl < 0.0f) cl = 0.0f;
    float dist_rew = sdx_exp(-2.0f * cl) * 0.1f;
    float up = clampf(tp.z - ti[2], 0.0f, 0.2f) * 100.0f;
    if (!(dist < 0.5f)) up = 0.0f;
    if (up > 20.0f) up = 20.0f;
    rew[e] = dist_rew + up;
    if (progress[e] >= 75 && dist >= 0.6f) rs = 1;
    reset[e] = rs;
    num_resets += rs; finished = finished + successes[e] * (float)rs;
  }
  if (num_resets > 0) consec[0] = S->av_factor * finished / (float)num_resets + (1.0f - S->av_factor) * consec[0]; /* GS:1771-1774 */
}

/* rl_games discount_values (call sites RGC:1473-1478): [H][N] arrays */
void sdxo_gae(const float* rewards, const float* values, const float* dones, const float* last_values,
              const float* last_dones, float* adv, float* returns, int H, int n, float gamma, float tau) {
  for (int e = 0; e < n; ++e) {
    float lastgaelam = 0.0f;
    for (int t = H - 1; t >= 0; --t) {
      float nnt, nv;
      if (t == H - 1) { nnt = 1.0f - last_dones[e]; nv = last_values[e]; }
      else { nnt = 1.0f - dones[(size_t)(t + 1) * n + e]; nv = values[(size_t)(t + 1) * n + e]; }
      float delta = rewards[(size_t)t * n + e] + gamma * nv * nnt - values[(size_t)t * n + e];
      lastgaelam = delta + gamma * tau * nnt * lastgaelam;
      adv[(size_t)t * n + e] = lastgaelam;
      returns[(size_t)t * n + e] = lastgaelam + values[(size_t)t * n + e];
    }
  }
}

void sdxo_tv_dataset(const sdx_scene_t* S, int n, const int64_t* reset, const float* brick, const float* finger_dist,
                     const float* tvalue, const float* states, float* succ, float* fail, int64_t* counts, int cap) {
  for (int e = 0; e < n; ++e) {
    if (!reset[e]) continue;
    float row[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), row);
    int ok = row[1] < 0.0f && finger_dist[e] < 0.6f && tvalue[e] > 0.8f;
    float* dst = (ok ? succ : fail) + 4 * (size_t)(counts[ok ? 0 : 1] % cap);
    for (int k = 0; k < 4; ++k) dst[k] = states[(size_t)e * 3 * STATE_FRAME + 177 + k];
    counts[ok ? 0 : 1] += 1;
  }
}


/* ================================================================== BlockAssemblyOrient (SDX_TASK_ORIENT)
 * OR = tasks/block_assembly/allegro_hand_block_assembly_orient.py.  Same scene, contact step, t-value gate and privileged
 * state frame as GraspSim; its own action mapping (object-centric arm IK, OR:1697-1778), observation frame (62 slots, of
 * which compute_real_observations writes 48, OR:1308-1326), reward (OR:1843-1907) and a scripted reset (OR:1390-1695).
 * PARITY: pre-physics, observations, reward / reset flags PINNED to the reference's own Python (oracle/gen_golden_orient.py
 * -> tests/golden/orient_*.npz); the reset script is sequencing of those pieces around the (unpinned) contact step. */
static inline float sgnf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); } /* torch.sign */
/* orientation_error (OR:1922-1925): xyz of desired * conj(current), flipped to the short way round */
static inline v3 orientation_error(q4 desired, q4 current) {
  q4 r = qmul(desired, qconj(current));
  float sg = sgnf(r.w);
  return V3(r.x * sg, r.y * sg, r.z * sg);
}
static inline float z_align(q4 q) { /* sign(d) d^2 with d = (R(q) z) . z  (OR:1198-1201, 1857-1860) */
  float d = qrot(q, V3(0.0f, 0.0f, 1.0f)).z;
  return sgnf(d) * (d * d);
}

/* pre_physics_step, after any reset (OR:1711-1778): fingers = EMA of the scaled actions; arm = IK towards the pose 22 cm
 * above / 18 cm behind the target brick with the fixed wrist orientation hand_target_quat */
void sdxo_orient_pre_physics(const sdx_scene_t* S, int n, const float* actions_in, float* actions, float* dof, const float* link,
                             const float* jac7, const float* brick, const int64_t* progress, const float* target_init) {
  for (int e = 0; e < n; ++e) {
    const float* a = actions_in + 23 * e;
    float* d = dof + (size_t)e * 72;
    float cur[23];
    for (int k = 0; k < 23; ++k) actions[23 * e + k] = a[k];
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
    if (pg > 75) dpose[2] = ((target_init[7 * e + 2] - hb[2]) + 0.15f) + 0.24f; /* OR:1735 */
    q4 want = {S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]};
    q4 hq = {hb[3], hb[4], hb[5], hb[6]};
    v3 re = orientation_error(want, hq);
    dpose[3] = re.x; dpose[4] = re.y; dpose[5] = re.z;
    float u[7];
    control_ik(jac7 + 42 * (size_t)e, dpose, u);
    for (int j = 0; j < 7; ++j) cur[j] = d[j] + u[j];
    if (pg > 75) for (int i = 7; i < 23; ++i) cur[i] = d[48 + i]; /* OR:1743 */
    for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
  }
}

/* post_physics_step (OR:1780-1784) when count_step != 0: progress += 1, compute_observations (OR:1087-1242 ->
 * compute_real_observations OR:1308-1326 + compute_contact_asymmetric_observations OR:1244-1306), compute_hand_reward
 * (OR:1843-1907).  count_step == 0 is the bare compute_observations() call inside reset_idx (OR:1461): observations,
 * finger distance and gate value only.  obs [n][186], states [n][564]. */
void sdxo_orient_post_physics(const sdx_scene_t* S, int n, const float* tv_wts, const float* brick, const float* dof,
                              const float* link, const float* actions, const float* target_init, int64_t* progress,
                              int64_t* reset, float* obs, float* states, float* rew, float* tvalue, float* finger_dist_out,
                              const float* successes, float* consec, int count_step) {
  int64_t num_resets = 0; float finished = 0.0f;
  for (int e = 0; e < n; ++e) {
    if (count_step) progress[e] += 1;
    const float* L = link + (size_t)e * SDX_NL * 13;
    const float* d = dof + (size_t)e * 72;
    const float* hb = L + 7 * 13;
    const float* ff = L + 11 * 13; const float* mf = L + 19 * 13; const float* rf = L + 23 * 13; const float* th = L + 15 * 13; /* OR:183-186 */
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
    v3 tp = V3(tg[0], tg[1], tg[2]); q4 tq = {tg[3], tg[4], tg[5], tg[6]};
    v3 tip[4]; const float* fs[4] = {ff, mf, rf, th};
    float nrm[4];
    for (int i = 0; i < 4; ++i) { /* OR:1158-1161 */
      q4 fq = {fs[i][3], fs[i][4], fs[i][5], fs[i][6]};
      tip[i] = vadd(V3(fs[i][0], fs[i][1], fs[i][2]), qrot(fq, V3(0.0f, 0.0f, 1.0f * 0.04f)));
      v3 dd = vsub(tp, tip[i]); nrm[i] = sqrtf(vdot(dd, dd));
    }
    float fdist = nrm[0] + nrm[1] + nrm[2] + nrm[3]; /* OR:1174-1175 */
    finger_dist_out[e] = fdist;
    q4 hq = {hb[3], hb[4], hb[5], hb[6]}; v3 hp = V3(hb[0], hb[1], hb[2]);
    q4 cq0 = {S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]};
    q4 cq = qmul(hq, cq0); v3 cp = vadd(qrot(hq, V3(S->cam_off_pos[0], S->cam_off_pos[1], S->cam_off_pos[2])), hp); /* OR:1181-1187 */
    q4 cqi = qconj(cq); v3 cpi = vneg(qrot(cqi, cp));
    q4 cvq = qmul(cqi, tq); v3 cvp = vadd(qrot(cqi, tp), cpi);
    float qin[4] = {cvq.x, cvq.y, cvq.z, cvq.w};
    float tv = sdxo_tvalue_one(tv_wts, qin);
    tvalue[e] = tv > 0.99f ? 1.0f : 0.0f; /* OR:1203-1205 */
    const float* ti = target_init + 7 * e;
    /* ---- obs frame 0 (OR:1308-1326); slots 16..29 and frames 1, 2 are never written by the reference */
    float* o = obs + (size_t)e * 3 * ORIENT_OBS_FRAME;
    for (int i = 0; i < 16; ++i) {
      float us = unscalef(d[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
      o[i] = us;
      o[30 + i] = actions[23 * e + 7 + i] - us;
      o[46 + i] = actions[23 * e + 7 + i];
    }
    /* ---- privileged state frame (OR:1244-1306): the GraspSim layout */
    float* s = states + (size_t)e * 3 * STATE_FRAME;
    for (int k = 2 * STATE_FRAME - 1; k >= 0; --k) s[STATE_FRAME + k] = s[k];
    for (int j = 0; j < 23; ++j) { s[j] = unscalef(d[j], S->dof_lo[j], S->dof_hi[j]); s[23 + j] = S->vel_obs_scale * d[24 + j]; }
    s[46] = tip[0].x; s[47] = tip[0].y; s[48] = tip[0].z;
    s[49] = tip[2].x; s[50] = tip[2].y; s[51] = tip[2].z;
    s[52] = tip[1].x; s[53] = tip[1].y; s[54] = tip[1].z;
    s[55] = tip[3].x; s[56] = tip[3].y; s[57] = tip[3].z;
    for (int k = 0; k < 23; ++k) s[58 + k] = actions[23 * e + k];
    for (int k = 0; k < 7; ++k) { s[81 + k] = hb[k]; s[88 + k] = tg[k]; }
    for (int k = 0; k < 6; ++k) s[95 + k] = hb[7 + k];
    for (int k = 0; k < 4; ++k) { s[101 + k] = ff[3 + k]; s[111 + k] = mf[3 + k]; s[121 + k] = rf[3 + k]; s[131 + k] = th[3 + k]; }
    for (int k = 0; k < 6; ++k) { s[105 + k] = ff[7 + k]; s[115 + k] = mf[7 + k]; s[125 + k] = rf[7 + k]; s[135 + k] = th[7 + k]; }
    for (int k = 0; k < 6; ++k) s[142 + k] = tg[7 + k];
    s[148] = ti[0]; s[149] = ti[1]; s[150] = ti[2];
    s[151] = tp.x - ti[0]; s[152] = tp.y - ti[1]; s[153] = tp.z - ti[2];
    s[154] = hp.x - tp.x; s[155] = hp.y - tp.y; s[156] = hp.z - tp.z;
    q4 rel = qmul(hq, qconj(tq));
    s[157] = rel.x; s[158] = rel.y; s[159] = rel.z; s[160] = rel.w;
    { v3 a = vsub(tp, tip[0]), b = vsub(tp, tip[2]), c = vsub(tp, tip[1]), dd = vsub(tp, tip[3]);
      s[161] = a.x; s[162] = a.y; s[163] = a.z; s[164] = b.x; s[165] = b.y; s[166] = b.z;
      s[167] = c.x; s[168] = c.y; s[169] = c.z; s[170] = dd.x; s[171] = dd.y; s[172] = dd.z; }
    s[173] = fdist;
    s[174] = cvp.x; s[175] = cvp.y; s[176] = cvp.z; s[177] = cvq.x; s[178] = cvq.y; s[179] = cvq.z; s[180] = cvq.w;
    s[181] = cvp.x; s[182] = cvp.y; s[183] = cvp.z; s[184] = cvq.x; s[185] = cvq.y; s[186] = cvq.z; s[187] = cvq.w;
    if (!count_step) continue;
    /* ---- reward / reset flags (OR:1852-1907) */
    float dist = nrm[0] + nrm[1] + nrm[2] + 3.0f * nrm[3];
    int64_t rs = reset[e];
    if (dist <= -1.0f) rs = 1;
    if ((float)progress[e] >= (float)S->max_episode_length - 1.0f) rs = 1;
    float drew = dist - 0.4f; if (drew < 0.0f) drew = 0.0f;
    if (progress[e] > 175) drew = 0.0f;
    float zrew = 1.0f - ((z_align(tq) + 1.0f) / 2.0f);
    rew[e] = sdx_exp(-(5.0f * zrew + 5.0f * drew));
    reset[e] = rs;
    num_resets += rs; finished = finished + successes[e] * (float)rs;
  }
  if (count_step && num_resets > 0) consec[0] = S->av_factor * finished / (float)num_resets + (1.0f - S->av_factor) * consec[0];
}

/* the two scripted arm motions of reset_idx, for the envs whose reset flag is set:
 * mode 0 (OR:1430-1455, before the banking): lift the hand to 42 cm above / 18 cm behind the CURRENT target pose; the arm
 *   target is clamped; the finger targets open by 0.01 rad ONCE (cur_targets_clone - 0.01 is the same every iteration)
 * mode 1 (OR:1659-1690, post_reset): teleport the arm along the IK solution towards 22 cm (+20 cm for the first 20
 *   iterations) above the INITIAL target pose -- joint positions, targets (unclamped) and zero velocities are written into
 *   the sim; the fingers are held at their reset pose */
void sdxo_orient_arm_script(const sdx_scene_t* S, int n, const int64_t* reset, int mode, int iter, float* dof, const float* link,
                            const float* jac7, const float* brick, const float* target_init) {
  for (int e = 0; e < n; ++e) {
    if (!reset[e]) continue;
    float* d = dof + (size_t)e * 72;
    const float* hb = link + ((size_t)e * SDX_NL + 7) * 13;
    float dpose[6];
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
    q4 want = {S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]};
    q4 hq = {hb[3], hb[4], hb[5], hb[6]};
    v3 re = orientation_error(want, hq);
    dpose[3] = re.x; dpose[4] = re.y; dpose[5] = re.z;
    float u[7];
    control_ik(jac7 + 42 * (size_t)e, dpose, u);
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
}

/* banking of the re-oriented heaps (OR:1465-1481), EVERY env in env order (the reference loops over range(num_envs), not
 * env_ids): fingers away from the brick (> 0.3), brick inside the bin's near half (0 < y < 0.5), gate passed (tvalue, already
 * thresholded to {0, 1}, > 0.6) -> the env's 72 free-brick root rows go to ring[type][index]; the index returns to 0 after
 * slot `wrap`.  rows_out [8][wrap + 1][72][13].  (The reference's .view(num_envs, 108, 13) at OR:1465 cannot hold the 132
 * bricks it indexes; the intent -- the whole heap, as Search banks it, SE:1348-1352 -- is what is restated.) */
void sdxo_orient_bank(const sdx_scene_t* S, int n, const float* brick, const float* finger_dist, const float* tvalue,
                      float* rows_out, int* index, int wrap) {
  for (int e = 0; e < n; ++e) {
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
    if (!(finger_dist[e] > 0.3f)) continue;
    if (!(0.5f > tg[1] && tg[1] > 0.0f)) continue;
    if (!(tvalue[e] > 0.6f)) continue;
    int ty = e % 8;
    float* dst = rows_out + (((size_t)ty * (wrap + 1)) + index[ty]) * NB * 13;
    for (int b = 0; b < NB; ++b) brick_root_row(S, brick + (size_t)e * 13 * NB, b, dst + b * 13);
    index[ty] += 1;
    if (index[ty] > wrap) index[ty] = 0;
  }
}

/* state writes of reset_idx / post_reset for the envs whose reset flag is set.
 * phase 0 (OR:1541-1605): heap <- bank row Philox(seed, env, episode) % min(per_type, bank_sample_range), velocities 0;
 *          hand at the prepare pose with the finger reset pose, velocities 0, targets = positions
 * phase 1 (OR:1623-1645, after 2 settle steps): remember the target's pose as its initial pose; hand written again
 * phase 2 (OR:1607-1610): progress, reset flag, successes cleared */
void sdxo_orient_reset(const sdx_scene_t* S, int n, uint64_t seed, const float* bank, int per_type, int phase, float* brick,
                       float* dof, float* target_init, int64_t* progress, int64_t* reset, float* successes, int* episode,
                       int* wsn, unsigned char* slp) {
  for (int e = 0; e < n; ++e) {
    if (!reset[e]) continue;
    float* B = brick + (size_t)e * 13 * NB;
    float* d = dof + (size_t)e * 72;
    if (phase == 0) {
      uint32_t r[4];
      philox(seed, (uint32_t)e, (uint32_t)episode[e], 1u, r);
      int range = per_type < S->bank_sample_range ? per_type : S->bank_sample_range;
      int slot = (int)(r[0] % (uint32_t)range);
      const float* rows = bank + (((size_t)(e % 8)) * per_type + slot) * NB * 13;
      for (int b = 0; b < NB; ++b) {
        float row[13];
        for (int k = 0; k < 7; ++k) row[k] = rows[b * 13 + k];
        for (int k = 7; k < 13; ++k) row[k] = 0.0f; /* OR:1570 */
        brick_from_root_row(S, B, b, row);
        slp[(size_t)e * NB + b] = 0;
      }
      wsn[2 * e] = 0; wsn[2 * e + 1] = 0;
      episode[e] += 1;
    }
    if (phase == 1) {
      float tg[13];
      brick_root_row(S, B, target_brick(e), tg);
      for (int k = 0; k < 7; ++k) target_init[7 * e + k] = tg[k]; /* OR:1623-1624 */
    }
    if (phase == 0 || phase == 1) {
      for (int j = 0; j < 7; ++j) { d[j] = S->prepare_arm[j]; d[24 + j] = 0.0f; d[48 + j] = S->prepare_arm[j]; }
      for (int i = 0; i < 16; ++i) {
        float v = scalef(S->finger_reset_unscaled[i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
        d[7 + i] = v; d[24 + 7 + i] = 0.0f; d[48 + 7 + i] = v;
      }
    }
    if (phase == 2) { progress[e] = 0; reset[e] = 0; successes[e] = 0.0f; }
  }
}


/* ================================================================== Search's camera features
 * SE = tasks/block_assembly/allegro_hand_block_assembly_search.py.  The reference renders a 128 x 128 segmentation image per env
 * (Isaac Gym camera sensor, SE:755-758, 873-878) and keeps the number of pixels showing the target brick and the centroid
 * (row, column) of those pixels (SE:1231-1241, 1640-1646).  Restated as ray casting against the scene's oriented boxes:
 * a pixel shows the target iff its ray hits the target box and no other box earlier.  PARITY UNPINNED against the reference
 * (Isaac Gym's rasteriser is a closed binary and renders the brick MESHES, we render their bounding boxes); pinned by the
 * analytic cases in tests/test_camera_cpu.py.  out [n][3] = pixels, int(mean row), int(mean column). */
typedef struct sdx_camera_t { float pos[3], fwd[3], right[3], up[3]; float inv_focal; int width, height; } sdx_camera_t;
static int ray_box(v3 o, v3 d, v3 c, const float* R, v3 h, float* t_entry) {
  v3 ol = mtmul(R, vsub(o, c));
  v3 dl = mtmul(R, d);
  float tmin = -3.0e38f, tmax = 3.0e38f;
  const float oa[3] = {ol.x, ol.y, ol.z}, da[3] = {dl.x, dl.y, dl.z}, ha[3] = {h.x, h.y, h.z};
  for (int a = 0; a < 3; ++a) {
    if (da[a] == 0.0f) { if (oa[a] < -ha[a] || oa[a] > ha[a]) return 0; }
    else {
      float inv = 1.0f / da[a];
      float t1 = (-ha[a] - oa[a]) * inv, t2 = (ha[a] - oa[a]) * inv;
      if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
      if (t1 > tmin) tmin = t1;
      if (t2 < tmax) tmax = t2;
    }
  }
  if (tmax < tmin || tmax < 0.0f) return 0;
  *t_entry = tmin;
  return 1;
}
void sdxo_segmentation_features(const sdx_scene_t* S, int n, const sdx_camera_t* cam, const float* brick, const float* link, int* out) {
  enum { MAXS = SDX_MAX_BRICKS + SDX_MAX_RSHAPES + SDX_MAX_STATIC };
  static __thread float sc[MAXS][3], sR[MAXS][9], sh[MAXS][3];
  const int nbr = S->n_bricks, nrs = S->n_rshapes, nst = S->n_static, nshape = nbr + nrs + nst;
  const v3 o = V3(cam->pos[0], cam->pos[1], cam->pos[2]), fw = V3(cam->fwd[0], cam->fwd[1], cam->fwd[2]);
  const v3 rt = V3(cam->right[0], cam->right[1], cam->right[2]), up = V3(cam->up[0], cam->up[1], cam->up[2]);
  const int W = cam->width, H = cam->height;
  for (int e = 0; e < n; ++e) {
    const float* B = brick + (size_t)e * 13 * NB;
    const int tb = target_brick(e);
    for (int b = 0; b < nbr; ++b) {
      sc[b][0] = B[0 * NB + b]; sc[b][1] = B[1 * NB + b]; sc[b][2] = B[2 * NB + b];
      q4 q = {B[3 * NB + b], B[4 * NB + b], B[5 * NB + b], B[6 * NB + b]};
      qmat(q, sR[b]);
      for (int k = 0; k < 3; ++k) sh[b][k] = S->br_half[3 * b + k];
    }
    for (int i = 0; i < nrs; ++i) {
      int t = nbr + i, L = S->rs_body[i];
      const float* lr = link + ((size_t)e * SDX_NL + L) * 13;
      q4 qL = {lr[3], lr[4], lr[5], lr[6]};
      v3 x = vadd(V3(lr[0], lr[1], lr[2]), qrot(qL, V3(S->rs_c[3 * i], S->rs_c[3 * i + 1], S->rs_c[3 * i + 2])));
      sc[t][0] = x.x; sc[t][1] = x.y; sc[t][2] = x.z;
      q4 ql = {S->rs_quat[4 * i], S->rs_quat[4 * i + 1], S->rs_quat[4 * i + 2], S->rs_quat[4 * i + 3]};
      qmat(qmul(qL, ql), sR[t]);
      for (int k = 0; k < 3; ++k) sh[t][k] = S->rs_h[3 * i + k];
    }
    for (int i = 0; i < nst; ++i) {
      int t = nbr + nrs + i;
      for (int k = 0; k < 3; ++k) { sc[t][k] = S->st_c[3 * i + k]; sh[t][k] = S->st_h[3 * i + k]; }
      for (int k = 0; k < 9; ++k) sR[t][k] = (k % 4 == 0) ? 1.0f : 0.0f;
    }
    int cnt = 0, sr = 0, scol = 0;
    const v3 tc = V3(sc[tb][0], sc[tb][1], sc[tb][2]), th = V3(sh[tb][0], sh[tb][1], sh[tb][2]);
    for (int r = 0; r < H; ++r)      /* every pixel: the kernel's bounding-rectangle pruning cannot change the result */
      for (int c = 0; c < W; ++c) {
        const float sx = (((float)c + 0.5f) - 0.5f * (float)W) * cam->inv_focal;
        const float sy = -((((float)r + 0.5f) - 0.5f * (float)H) * cam->inv_focal);
        const v3 d = vadd(vadd(fw, vscale(rt, sx)), vscale(up, sy));
        float tt;
        if (!ray_box(o, d, tc, sR[tb], th, &tt)) continue;
        int occluded = 0;
        for (int s2 = 0; s2 < nshape && !occluded; ++s2) {
          if (s2 == tb) continue;
          float ts;
          if (ray_box(o, d, V3(sc[s2][0], sc[s2][1], sc[s2][2]), sR[s2], V3(sh[s2][0], sh[s2][1], sh[s2][2]), &ts) && ts < tt) occluded = 1;
        }
        if (!occluded) { cnt++; sr += r; scol += c; }
      }
    out[3 * e] = cnt;
    out[3 * e + 1] = cnt > 0 ? (int)((float)sr / (float)cnt) : 0;
    out[3 * e + 2] = cnt > 0 ? (int)((float)scol / (float)cnt) : 0;
  }
}


/* ================================================================== BlockAssemblySearch (SDX_TASK_SEARCH; BASELINE configs[0])
 * SE = tasks/block_assembly/allegro_hand_block_assembly_search.py.  Scene = GraspSim's with the drop lattice 6 cm higher, the
 * floor bricks 5 mm higher and a 12x12 base-plate (seqdex_b200/scene.py); Orient's finger drives.  The hand digs through the
 * heap until the overview camera sees enough of the target brick.
 * PARITY: pre-physics, observations, privileged state, reward / reset flags PINNED to the reference's own Python
 * (oracle/gen_golden_search.py -> tests/golden/search_*.npz).  The camera features come from sdxo_segmentation_features
 * (unpinned: closed rasteriser); the reset is sequencing around the (unpinned) contact step. */
#define SEARCH_TVOBS 650 /* 10 frames x 65 (SE:400) */
static inline float u11(uint32_t r) { return (float)(r >> 8) * (2.0f / 16777216.0f) - 1.0f; } /* U[-1, 1) from 24 random bits */
static const int SEARCH_PIXEL_THRESHOLD[8] = {20, 20, 15, 20, 20, 30, 30, 20}; /* SE:1290 */

/* pre_physics_step after any reset (SE:1546-1596): fingers = clamped EMA of the scaled actions, arm = IK towards 24 cm above /
 * 18 cm behind the target brick with the wrist orientation quat_from_euler_xyz(0, 3.14, 1.57).  (The reference tracks the
 * target position cached by the last compute_observations; ours reads the current one -- they differ only in the first
 * step after a reset.) */
void sdxo_search_pre_physics(const sdx_scene_t* S, int n, const float* actions_in, float* actions, float* dof, const float* link,
                             const float* jac7, const float* brick) {
  for (int e = 0; e < n; ++e) {
    const float* a = actions_in + 23 * e;
    float* d = dof + (size_t)e * 72;
    float cur[23];
    for (int k = 0; k < 23; ++k) actions[23 * e + k] = a[k];
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
    q4 want = {S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]};
    q4 hq = {hb[3], hb[4], hb[5], hb[6]};
    v3 re = orientation_error(want, hq);
    dpose[3] = re.x; dpose[4] = re.y; dpose[5] = re.z;
    float u[7];
    control_ik(jac7 + 42 * (size_t)e, dpose, u);
    for (int j = 0; j < 7; ++j) cur[j] = d[j] + u[j];
    for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
  }
}

/* post_physics_step (SE:1598-1602) after the end-of-episode camera branch (handled by the caller): progress += 1,
 * compute_observations (SE:1036-1166 -> compute_contact_observations SE:1220-1245, compute_contact_asymmetric_observations
 * SE:1168-1218, the 10-frame gate input SE:1154-1166), compute_hand_reward (SE:1660-1712).
 * seg [n][3] = pixels / centre row / centre column of the LAST rendered segmentation image.  obs [n][186], states [n][564],
 * tvobs [n][650]. */
void sdxo_search_post_physics(const sdx_scene_t* S, int n, const float* brick, const float* dof, const float* link, const float* netf,
                              const float* actions, const float* target_init, const int* seg, int64_t* progress, int64_t* reset,
                              float* obs, float* states, float* tvobs, float* rew, float* finger_dist_out, const float* successes,
                              float* consec) {
  int64_t num_resets = 0; float finished = 0.0f;
  for (int e = 0; e < n; ++e) {
    progress[e] += 1;
    const float* L = link + (size_t)e * SDX_NL * 13;
    const float* d = dof + (size_t)e * 72;
    const float* hb = L + 7 * 13;
    const float* ff = L + 11 * 13; const float* mf = L + 19 * 13; const float* rf = L + 23 * 13; const float* th = L + 15 * 13;
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
    v3 tp = V3(tg[0], tg[1], tg[2]); q4 tq = {tg[3], tg[4], tg[5], tg[6]};
    v3 tip[4]; const float* fs[4] = {ff, mf, rf, th};
    float nrm[4];
    for (int i = 0; i < 4; ++i) { /* SE:1086-1089 */
      q4 fq = {fs[i][3], fs[i][4], fs[i][5], fs[i][6]};
      tip[i] = vadd(V3(fs[i][0], fs[i][1], fs[i][2]), qrot(fq, V3(0.0f, 0.0f, 1.0f * 0.04f)));
      v3 dd = vsub(tp, tip[i]); nrm[i] = sqrtf(vdot(dd, dd));
    }
    float fdist = nrm[0] + nrm[1] + nrm[2] + nrm[3];
    finger_dist_out[e] = fdist;
    q4 hq = {hb[3], hb[4], hb[5], hb[6]}; v3 hp = V3(hb[0], hb[1], hb[2]);
    q4 cq0 = {S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]};
    q4 cq = qmul(hq, cq0); v3 cp = vadd(qrot(hq, V3(S->cam_off_pos[0], S->cam_off_pos[1], S->cam_off_pos[2])), hp); /* SE:1104-1108 */
    q4 cqi = qconj(cq);
    q4 cvq = qmul(cqi, tq);
    (void)cp;
    float contacts = 0.0f; /* arm links 0..6 whose net contact force reaches 0.1 N (SE:919, 1111-1115) */
    for (int k = 0; k < 7; ++k) {
      const float* f = netf + ((size_t)e * SDX_NL + k) * 3;
      float nf = sqrtf(vdot(V3(f[0], f[1], f[2]), V3(f[0], f[1], f[2])));
      contacts = contacts + (nf >= 0.1f ? 1.0f : 0.0f);
    }
    const float* ti = target_init + 7 * e;
    const float sx = (float)seg[3 * e + 1] / 128.0f, sy = (float)seg[3 * e + 2] / 128.0f, sn = (float)seg[3 * e] / 100.0f;
    /* ---- obs frame 0 (SE:1220-1229) */
    float* o = obs + (size_t)e * 3 * ORIENT_OBS_FRAME;
    for (int i = 0; i < 16; ++i) {
      float us = unscalef(d[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
      o[i] = us;
      o[30 + i] = actions[23 * e + 7 + i] - us;
      o[46 + i] = actions[23 * e + 7 + i];
    }
    /* ---- privileged state, frame 0 only: Search never shifts a history (SE:1168-1218) */
    float* s = states + (size_t)e * 3 * STATE_FRAME;
    for (int j = 0; j < 23; ++j) { s[j] = unscalef(d[j], S->dof_lo[j], S->dof_hi[j]); s[23 + j] = S->vel_obs_scale * d[24 + j]; }
    s[46] = tip[0].x; s[47] = tip[0].y; s[48] = tip[0].z;
    s[49] = tip[2].x; s[50] = tip[2].y; s[51] = tip[2].z;
    s[52] = tip[1].x; s[53] = tip[1].y; s[54] = tip[1].z;
    s[55] = tip[3].x; s[56] = tip[3].y; s[57] = tip[3].z;
    for (int k = 0; k < 23; ++k) s[58 + k] = actions[23 * e + k];
    for (int k = 0; k < 7; ++k) { s[81 + k] = hb[k]; s[88 + k] = tg[k]; }
    for (int k = 96; k < 120; ++k) s[k] = 0.0f; /* hand_pos_history_0..7 are means of a zeroed buffer (SE:1457-1465) */
    s[120] = sx; s[121] = sy; s[122] = sn;
    for (int k = 0; k < 6; ++k) s[123 + k] = hb[7 + k];
    for (int k = 0; k < 4; ++k) { s[129 + k] = ff[3 + k]; s[139 + k] = mf[3 + k]; s[149 + k] = rf[3 + k]; s[159 + k] = th[3 + k]; }
    for (int k = 0; k < 6; ++k) { s[133 + k] = ff[7 + k]; s[143 + k] = mf[7 + k]; s[153 + k] = rf[7 + k]; s[163 + k] = th[7 + k]; }
    for (int k = 0; k < 6; ++k) s[169 + k] = tg[7 + k];
    /* ---- the gate's 10-frame input (SE:1154-1166): oldest frame out, newest = obs[0:62] with slots 26..29 = camera-frame
     * target quaternion, then centre / 128, centre / 128, pixels / 100 */
    float* tvo = tvobs + (size_t)e * SEARCH_TVOBS;
    for (int k = 0; k < 9 * 65; ++k) tvo[k] = tvo[k + 65];
    float* fr = tvo + 9 * 65;
    for (int k = 0; k < 62; ++k) fr[k] = o[k];
    fr[26] = cvq.x; fr[27] = cvq.y; fr[28] = cvq.z; fr[29] = cvq.w;
    fr[62] = sx; fr[63] = sy; fr[64] = sn;
    /* ---- reward / reset flags (SE:1668-1712) */
    float dist_rew = -0.2f * fdist; if (dist_rew > -0.06f) dist_rew = -0.06f;
    float asq = 0.0f;
    for (int k = 0; k < 23; ++k) asq = asq + actions[23 * e + k] * actions[23 * e + k];
    float action_penalty = asq * 0.005f;
    float up = clampf(tp.z - ti[2], 0.0f, 0.1f) * 1000.0f - clampf(tp.x - ti[0], 0.0f, 0.1f) * 1000.0f - clampf(tp.y - ti[1], 0.0f, 0.1f) * 1000.0f;
    rew[e] = (((dist_rew - contacts) + 0.0f) - action_penalty) + up;
    int64_t rs = reset[e];
    if (fdist <= -1.0f) rs = 1;
    if ((float)progress[e] >= (float)S->max_episode_length - 1.0f) rs = 1;
    reset[e] = rs;
    num_resets += rs; finished = finished + successes[e] * (float)rs;
  }
  if (num_resets > 0) consec[0] = S->av_factor * finished / (float)num_resets + (1.0f - S->av_factor) * consec[0];
}

/* every env's hand (or the flagged ones) teleported to a stored pose: which = 0 arm_hand_default_dof_pos (SE:990-998, 1405-1410),
 * 1 arm_hand_prepare_dof_poses (SE:1483-1493); positions = targets = pose, velocities 0 */
void sdxo_search_hand_pose(const sdx_scene_t* S, int n, const int64_t* mask, int which, float* dof) {
  const float* pose = which ? S->prepare_dof : S->default_dof;
  for (int e = 0; e < n; ++e) {
    if (mask && !mask[e]) continue;
    float* d = dof + (size_t)e * 72;
    for (int j = 0; j < SDX_ND; ++j) { d[j] = pose[j]; d[24 + j] = 0.0f; d[48 + j] = pose[j]; }
  }
}

/* compute_emergence_reward (SE:1640-1646): 5 x (pixels now - pixels at the last render); baseline != 0 only records */
void sdxo_search_emergence(int n, const int* seg, float* last_pixels, float* emergence, int baseline) {
  for (int e = 0; e < n; ++e) {
    float pix = (float)seg[3 * e];
    if (!baseline) emergence[e] = (pix - last_pixels[e]) * 5.0f;
    last_pixels[e] = pix;
  }
}

/* banking of the dug-out heaps (SE:1305-1340), EVERY env in env order: enough pixels of the target visible -> the 72 free-brick
 * root rows and the hand's DoF state go to ring[type][index]; the index returns to 0 after slot `wrap` (10000) */
void sdxo_search_bank(const sdx_scene_t* S, int n, const float* brick, const float* dof, const int* seg, float* rows_out,
                      float* hand_out, int* index, int wrap) {
  for (int e = 0; e < n; ++e) {
    int ty = e % 8;
    if (!(seg[3 * e] > SEARCH_PIXEL_THRESHOLD[ty])) continue;
    float* dst = rows_out + (((size_t)ty * (wrap + 1)) + index[ty]) * NB * 13;
    for (int b = 0; b < NB; ++b) brick_root_row(S, brick + (size_t)e * 13 * NB, b, dst + b * 13);
    float* hd = hand_out + (((size_t)ty * (wrap + 1)) + index[ty]) * 46;
    const float* d = dof + (size_t)e * 72;
    for (int j = 0; j < SDX_ND; ++j) { hd[2 * j] = d[j]; hd[2 * j + 1] = d[24 + j]; }
    index[ty] += 1;
    if (index[ty] > wrap) index[ty] = 0;
  }
}

/* state writes of reset_idx / post_reset for the flagged envs.
 * phase 0 (SE:1388-1411): free bricks back on the drop lattice with 2 cm xy jitter, velocities 0; the target brick at
 *          (0.25 + 0.2 r, 0.19 + 0.15 r, 0.9) with ONE r per env (SE:1392-1394 index the same random column twice); hand at the
 *          default pose.  Own Philox stream instead of torch's generator.
 * phase 1 (SE:1483-1496, after the 60 settle steps): hand at the prepare pose; the target's pose becomes its initial pose
 * phase 2 (SE:1430-1433): progress, reset flag, successes cleared */
void sdxo_search_reset(const sdx_scene_t* S, int n, uint64_t seed, int phase, float* brick, float* dof, float* target_init,
                       int64_t* progress, int64_t* reset, float* successes, int* episode, int* wsn, unsigned char* slp) {
  for (int e = 0; e < n; ++e) {
    if (!reset[e]) continue;
    float* B = brick + (size_t)e * 13 * NB;
    float* d = dof + (size_t)e * 72;
    if (phase == 0) {
      const int tb = target_brick(e);
      for (int b = 0; b < NB; ++b) {
        uint32_t r[4];
        philox(seed, (uint32_t)e, (uint32_t)episode[e], 16u + (uint32_t)b, r);
        float row[13];
        for (int k = 0; k < 7; ++k) row[k] = S->brick_init[b * 13 + k];
        for (int k = 7; k < 13; ++k) row[k] = 0.0f;
        row[0] = row[0] + u11(r[0]) * 0.02f;
        row[1] = row[1] + u11(r[1]) * 0.02f;
        if (b == tb) {
          uint32_t q[4];
          philox(seed, (uint32_t)e, (uint32_t)episode[e], 2u, q);
          float rr = u11(q[0]);
          row[0] = 0.25f + rr * 0.2f; row[1] = 0.19f + rr * 0.15f; row[2] = 0.9f;
        }
        brick_from_root_row(S, B, b, row);
        slp[(size_t)e * NB + b] = 0;
      }
      wsn[2 * e] = 0; wsn[2 * e + 1] = 0;
      episode[e] += 1;
      for (int j = 0; j < SDX_ND; ++j) { d[j] = S->default_dof[j]; d[24 + j] = 0.0f; d[48 + j] = S->default_dof[j]; }
    }
    if (phase == 1) {
      for (int j = 0; j < SDX_ND; ++j) { d[j] = S->prepare_dof[j]; d[24 + j] = 0.0f; d[48 + j] = S->prepare_dof[j]; }
      float tg[13];
      brick_root_row(S, B, target_brick(e), tg);
      for (int k = 0; k < 7; ++k) target_init[7 * e + k] = tg[k]; /* SE:1495-1496 */
    }
    if (phase == 2) { progress[e] = 0; reset[e] = 0; successes[e] = 0.0f; }
  }
}

/* the reduction the reference applies to a rendered segmentation image (SE:1231-1241): mask [h][w] of the pixels that carry the
 * target's id -> pixels, int(mean row), int(mean column).  Exposed so the golden test can pin the arithmetic on real images. */
void sdxo_mask_features(const unsigned char* mask, int h, int w, int* out) {
  int cnt = 0, sr = 0, sc2 = 0;
  for (int r = 0; r < h; ++r) for (int c = 0; c < w; ++c) if (mask[r * w + c]) { cnt++; sr += r; sc2 += c; }
  out[0] = cnt;
  out[1] = cnt > 0 ? (int)((float)sr / (float)cnt) : 0;
  out[2] = cnt > 0 ? (int)((float)sc2 / (float)cnt) : 0;
}

void sdxo_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t out[4]) { philox(seed, c0, c1, c2, out); }
int sdxo_scene_size(void) { return (int)sizeof(sdx_scene_t); }
int sdxo_work_size(void) { return (int)sizeof(work_t); }

/* ================================================================== BlockAssemblyInsertSim (SDX_TASK_INSERT_SIM; last link of BASELINE configs[3])
 * IS = tasks/block_assembly/allegro_hand_block_assembly_insert_sim.py.  One observation frame of 75 (IS:1280-1298), one privileged
 * frame of 188 (IS:1222-1278), reward / resets IS:1640-1694, reset from the banked grasps IS:1328-1493.
 * PARITY: pre-physics, observations, reward / reset flags and reset_idx PINNED to the reference's own Python
 * (oracle/gen_golden_insert.py -> tests/golden/insert_*.npz). */
#define INSERT_OBS 75
/* asin on [-1, 1] (cephes asinf: minimax polynomial on |x| <= 0.5, pi/2 - 2 asin(sqrt((1 - x) / 2)) beyond); same text in csrc */
static inline float sdx_asin(float x) {
  float a = fabsf(x), z, w;
  int big = a > 0.5f;
  if (big) { z = 0.5f * (1.0f - a); w = sqrtf(z); } else { w = a; z = a * a; }
  float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * w + w;
  if (big) p = 1.5707963267948966f - (p + p);
  return x < 0.0f ? -p : p;
}
/* the pose the held brick has to reach: plate position, one brick height per plate level up, half a stud along y (and x for the
 * 1x1), every offset rotated with the plate and added in the reference's order (IS:1124-1132) */
static void insert_target(int e, const float* plate, v3* pos, q4* rot) {
  q4 q = {plate[3], plate[4], plate[5], plate[6]};
  v3 p = V3(plate[0], plate[1], plate[2]);
  const float lvl = (float)(1 + e % 3);
  p = vadd(p, qrot(q, V3(0.0f * (0.0375f * lvl), 0.0f * (0.0375f * lvl), 1.0f * (0.0375f * lvl))));
  if (e % 8 == 5) {
    p = vadd(p, qrot(q, V3(1.0f * 0.015f, 0.0f * 0.015f, 0.0f * 0.015f)));
    p = vadd(p, qrot(q, V3(0.0f * 0.015f, 1.0f * 0.015f, 0.0f * 0.015f)));
  } else p = vadd(p, qrot(q, V3(0.0f * 0.015f, 1.0f * 0.015f, 0.0f * 0.015f)));
  *pos = p; *rot = q;
}
static inline float rot_dist_sym(q4 tq, q4 eq) {   /* IS:1656-1660: the plate's pose or that pose turned by pi about z */
  q4 d1 = qmul(tq, qconj(eq));
  q4 sym = qmul(eq, (q4){0.0f, 0.0f, 1.0f, 0.0f});
  q4 d2 = qmul(tq, qconj(sym));
  float n1 = sqrtf(d1.x * d1.x + d1.y * d1.y + d1.z * d1.z), n2 = sqrtf(d2.x * d2.x + d2.y * d2.y + d2.z * d2.z);
  float r1 = 2.0f * sdx_asin(n1 > 1.0f ? 1.0f : n1), r2 = 2.0f * sdx_asin(n2 > 1.0f ? 1.0f : n2);
  return r1 < r2 ? r1 : r2;
}

/* pre_physics_step after resets (IS:1516-1565): fingers = EMA of the scaled actions; arm = IK for (0.64 a[0:3], orientation error
 * to the fixed wrist orientation hand_target_quat), which is also kept for the reward's reset test */
void sdxo_insert_pre_physics(const sdx_scene_t* S, int n, const float* actions_in, float* actions, float* dof, const float* link,
                             const float* jac7, float* rot_err) {
  for (int e = 0; e < n; ++e) {
    const float* a = actions_in + 23 * e;
    float* d = dof + (size_t)e * 72;
    float cur[23];
    for (int k = 0; k < 23; ++k) actions[23 * e + k] = a[k];
    for (int i = 0; i < 16; ++i) {
      float t = scalef(a[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
      cur[7 + i] = S->act_moving_average * t + (1.0f - S->act_moving_average) * d[48 + 7 + i];
    }
    const float* hb = link + ((size_t)e * SDX_NL + 7) * 13;
    q4 want = {S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]};
    q4 hq = {hb[3], hb[4], hb[5], hb[6]};
    v3 re = orientation_error(want, hq);
    rot_err[3 * e] = re.x; rot_err[3 * e + 1] = re.y; rot_err[3 * e + 2] = re.z;
    float dpose[6] = {a[0] * 0.64f, a[1] * 0.64f, a[2] * 0.64f, re.x, re.y, re.z};
    float u[7];
    control_ik(jac7 + 42 * (size_t)e, dpose, u);
    for (int j = 0; j < 7; ++j) cur[j] = clampf(d[j] + u[j], S->dof_lo[j], S->dof_hi[j]);
    for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
  }
}

/* post_physics_step (IS:1567-1572): progress += 1, observations, reward, reset flags.  obs [n][75], states [n][188] */
void sdxo_insert_post_physics(const sdx_scene_t* S, int n, const float* brick, const float* dof, const float* link, const float* actions,
                              const float* target_init, const float* plate, const float* rot_err, int64_t* progress, int64_t* reset,
                              float* obs, float* states, float* rew, float* finger_dist_out, const float* successes, float* consec) {
  int64_t num_resets = 0; float finished = 0.0f;
  for (int e = 0; e < n; ++e) {
    progress[e] += 1;
    const float* L = link + (size_t)e * SDX_NL * 13;
    const float* d = dof + (size_t)e * 72;
    const float* hb = L + 7 * 13;
    const float* ff = L + 11 * 13; const float* mf = L + 19 * 13; const float* rf = L + 23 * 13; const float* th = L + 15 * 13; /* IS:166-169 */
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, target_brick(e), tg);
    v3 tp = V3(tg[0], tg[1], tg[2]); q4 tq = {tg[3], tg[4], tg[5], tg[6]};
    v3 tip[4]; const float* fs[4] = {ff, mf, rf, th};
    for (int i = 0; i < 4; ++i) { /* IS:1166-1169 */
      q4 fq = {fs[i][3], fs[i][4], fs[i][5], fs[i][6]};
      tip[i] = vadd(V3(fs[i][0], fs[i][1], fs[i][2]), qrot(fq, V3(0.0f, 0.0f, 1.0f * 0.04f)));
    }
    float nrm[4];
    for (int i = 0; i < 4; ++i) { v3 dd = vsub(tp, tip[i]); nrm[i] = sqrtf(vdot(dd, dd)); }
    float fdist = nrm[0] + nrm[1] + nrm[2] + nrm[3]; /* IS:1183-1184 */
    finger_dist_out[e] = fdist;
    q4 hq = {hb[3], hb[4], hb[5], hb[6]}; v3 hp = V3(hb[0], hb[1], hb[2]);
    q4 cq0 = {S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]};
    q4 cq = qmul(hq, cq0); v3 cp = vadd(qrot(hq, V3(S->cam_off_pos[0], S->cam_off_pos[1], S->cam_off_pos[2])), hp);
    q4 cqi = qconj(cq); v3 cpi = vneg(qrot(cqi, cp));
    q4 cvq = qmul(cqi, tq); v3 cvp = vadd(qrot(cqi, tp), cpi);
    v3 ep; q4 eq;
    insert_target(e, plate + 7 * e, &ep, &eq);
    const float* ti = target_init + 7 * e;
    /* ---- observation frame (IS:1280-1298); slots 16:23 and 60 are never written */
    float* o = obs + (size_t)e * INSERT_OBS;
    for (int i = 0; i < 16; ++i) o[i] = unscalef(d[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
    for (int k = 0; k < 23; ++k) o[23 + k] = actions[23 * e + k];
    o[46] = hp.x - ep.x; o[47] = hp.y - ep.y; o[48] = hp.z - ep.z;
    { q4 r = qmul(hq, qconj(eq)); o[49] = r.x; o[50] = r.y; o[51] = r.z; o[52] = r.w; }
    o[53] = hp.x - tp.x; o[54] = hp.y - tp.y; o[55] = hp.z - tp.z;
    { q4 r = qmul(hq, qconj(tq)); o[56] = r.x; o[57] = r.y; o[58] = r.z; o[59] = r.w; }
    o[61] = ep.x; o[62] = ep.y; o[63] = ep.z; o[64] = eq.x; o[65] = eq.y; o[66] = eq.z; o[67] = eq.w;
    o[68] = tp.x - ep.x; o[69] = tp.y - ep.y; o[70] = tp.z - ep.z;
    { q4 r = qmul(tq, qconj(eq)); o[71] = r.x; o[72] = r.y; o[73] = r.z; o[74] = r.w; }
    /* ---- privileged frame (IS:1222-1278): GraspSim's, with the episode clock in slot 141 and the insertion pose in 181:188 */
    float* s = states + (size_t)e * STATE_FRAME;
    for (int j = 0; j < 23; ++j) { s[j] = unscalef(d[j], S->dof_lo[j], S->dof_hi[j]); s[23 + j] = S->vel_obs_scale * d[24 + j]; }
    s[46] = tip[0].x; s[47] = tip[0].y; s[48] = tip[0].z;
    s[49] = tip[2].x; s[50] = tip[2].y; s[51] = tip[2].z;
    s[52] = tip[1].x; s[53] = tip[1].y; s[54] = tip[1].z;
    s[55] = tip[3].x; s[56] = tip[3].y; s[57] = tip[3].z;
    for (int k = 0; k < 23; ++k) s[58 + k] = actions[23 * e + k];
    for (int k = 0; k < 7; ++k) { s[81 + k] = hb[k]; s[88 + k] = tg[k]; }
    for (int k = 0; k < 6; ++k) s[95 + k] = hb[7 + k];
    for (int k = 0; k < 4; ++k) { s[101 + k] = ff[3 + k]; s[111 + k] = mf[3 + k]; s[121 + k] = rf[3 + k]; s[131 + k] = th[3 + k]; }
    for (int k = 0; k < 6; ++k) { s[105 + k] = ff[7 + k]; s[115 + k] = mf[7 + k]; s[125 + k] = rf[7 + k]; s[135 + k] = th[7 + k]; }
    s[141] = (float)progress[e] / (float)S->max_episode_length;
    for (int k = 0; k < 6; ++k) s[142 + k] = tg[7 + k];
    s[148] = ti[0]; s[149] = ti[1]; s[150] = ti[2];
    s[151] = tp.x - ti[0]; s[152] = tp.y - ti[1]; s[153] = tp.z - ti[2];
    s[154] = hp.x - tp.x; s[155] = hp.y - tp.y; s[156] = hp.z - tp.z;
    { q4 rel = qmul(hq, qconj(tq)); s[157] = rel.x; s[158] = rel.y; s[159] = rel.z; s[160] = rel.w; }
    { v3 a = vsub(tp, tip[0]), b = vsub(tp, tip[2]), c = vsub(tp, tip[1]), dd = vsub(tp, tip[3]);
      s[161] = a.x; s[162] = a.y; s[163] = a.z; s[164] = b.x; s[165] = b.y; s[166] = b.z;
      s[167] = c.x; s[168] = c.y; s[169] = c.z; s[170] = dd.x; s[171] = dd.y; s[172] = dd.z; }
    s[173] = fdist;
    s[174] = cvp.x; s[175] = cvp.y; s[176] = cvp.z; s[177] = cvq.x; s[178] = cvq.y; s[179] = cvq.z; s[180] = cvq.w;
    s[181] = ep.x; s[182] = ep.y; s[183] = ep.z; s[184] = eq.x; s[185] = eq.y; s[186] = eq.z; s[187] = eq.w;
    /* ---- reward / resets (IS:1650-1693) */
    float dist = nrm[0] + nrm[1] + nrm[2] + 3.0f * nrm[3];
    float rd = rot_dist_sym(tq, eq);
    v3 dp = vsub(tp, ep);
    float pd = sqrtf(vdot(dp, dp));
    float insert_reward = sdx_exp(-1.0f * rd - 20.0f * (pd < 0.0f ? 0.0f : pd));
    v3 dp2 = vsub(ep, tp);
    float bonus = (sqrtf(vdot(dp2, dp2)) < 0.02f && rd < 0.2f) ? 1.0f : 0.0f;
    int64_t rs = reset[e];
    if (dist >= 0.6f) rs = 1;
    const float* re = rot_err + 3 * e;
    if (re[0] * re[0] + re[1] * re[1] + re[2] * re[2] >= 0.03f) rs = 1;
    if ((float)progress[e] >= (float)S->max_episode_length - 1.0f) rs = 1;
    rew[e] = bonus + insert_reward;
    reset[e] = rs;
    num_resets += rs; finished += successes[e] * (float)rs;
  }
  if (num_resets > 0) consec[0] = S->av_factor * finished / (float)num_resets + (1.0f - S->av_factor) * consec[0];
}

/* reset_idx (IS:1328-1493) for the envs whose reset flag is set.  success_buf first (IS:1341-1350, from the state the episode
 * ended in), then: every brick back to where it is parked (IS:1432-1433), the base-plate to (0.25, -0.2, 0.618) with the yaw
 * `plate_yaw_idx` x 1.57 -- ONE draw per call for all envs that reset (IS:1435-1446; torch_rand_int(0, 1) is always 0, so the
 * plate never shifts) --, the target brick and the hand's DoF positions from a banked grasp (slot = Philox % per_type; the
 * reference draws random.sample(range(5000)); slot_by_env, a test hook, names the slot per env instead), velocities zero, targets = positions. */
void sdxo_insert_reset(const sdx_scene_t* S, int n, uint64_t seed, const float* bank_obj, const float* bank_hand, int per_type, int plate_yaw_idx,
                       const int* slot_by_env, int do_success, float* brick, float* dof, float* plate, float* target_init, int64_t* progress,
                       int64_t* reset, float* successes, float* success_buf, int* episode, int* wsn, unsigned char* slp) {
  for (int e = 0; e < n; ++e) {
    if (!reset[e]) continue;
    float* B = brick + (size_t)e * 13 * NB;
    float* d = dof + (size_t)e * 72;
    const int tb = target_brick(e);
    if (do_success) {
      float tg[13];
      brick_root_row(S, B, tb, tg);
      v3 ep; q4 eq;
      insert_target(e, plate + 7 * e, &ep, &eq);
      q4 tq = {tg[3], tg[4], tg[5], tg[6]};
      float rd = rot_dist_sym(tq, eq);
      v3 dp = vsub(ep, V3(tg[0], tg[1], tg[2]));
      float ok = (sqrtf(vdot(dp, dp)) < 0.02f && rd < 0.2f) ? 1.0f : 0.0f;
      success_buf[2 * e] = ok; success_buf[2 * e + 1] = ok <= 0.5f ? 1.0f : 0.0f;
    }
    for (int b = 0; b < S->n_bricks; ++b) {
      float row[13];
      for (int k = 0; k < 7; ++k) row[k] = S->brick_init[13 * b + k];
      for (int k = 7; k < 13; ++k) row[k] = 0.0f;
      brick_from_root_row(S, B, b, row);
    }
    float* pl = plate + 7 * e;
    pl[0] = 0.25f; pl[1] = -0.2f; pl[2] = 0.618f;
    pl[3] = 0.0f; pl[4] = 0.0f; pl[5] = plate_yaw_idx ? S->insert_plate_zw[0] : 0.0f; pl[6] = plate_yaw_idx ? S->insert_plate_zw[1] : 1.0f;
    uint32_t r[4];
    philox(seed, (uint32_t)e, (uint32_t)episode[e], 1u, r);
    int slot = slot_by_env ? slot_by_env[e] : (int)(r[0] % (uint32_t)per_type);
    const float* ob = bank_obj + (((size_t)(e % 8)) * per_type + slot) * 13;
    const float* hd = bank_hand + (((size_t)(e % 8)) * per_type + slot) * 46;
    float row[13];
    for (int k = 0; k < 7; ++k) row[k] = ob[k];
    for (int k = 7; k < 13; ++k) row[k] = 0.0f;
    brick_from_root_row(S, B, tb, row);
    for (int j = 0; j < SDX_ND; ++j) { d[j] = hd[2 * j]; d[24 + j] = 0.0f; d[48 + j] = hd[2 * j]; }
    for (int k = 0; k < 7; ++k) target_init[7 * e + k] = ob[k];
    progress[e] = 0; reset[e] = 0; successes[e] = 0.0f;
    wsn[2 * e] = 0; wsn[2 * e + 1] = 0;
    for (int b = 0; b < NB; ++b) slp[(size_t)e * NB + b] = 0;
    episode[e] += 1;
  }
}


/* ================================================================== ToolPositioningGrasp / ToolPositioningOrient (BASELINE configs[4])
 * TG = tasks/tool_positioning/allegro_hand_tool_positioning_grasp.py, TO = tasks/tool_positioning/allegro_hand_tool_positioning_orient.py.
 * One free body per env: the tool (body 0, a compound of boxes).  Observation frame 156 x 3 (TG:1338-1368 = TO:1201-1236), privileged
 * frame 188 x 3 (TG:1274-1336; TO:1137-1199 puts the plate pose in 181:188), rewards / resets TG:1741-1893 and TO:1574-1626,
 * reset_idx TG:1412-1578 and TO:1265-1436, pre_physics_step TG:1580-1675 and TO:1438-1509.
 * PARITY: pre-physics, observations, reward / reset flags and reset_idx of BOTH tasks PINNED to the reference's own Python
 * (oracle/gen_golden_tool.py -> tests/golden/tool_*.npz). */
#define TOOL_OBS 156
#define TOOL_BODY 0
#define TOOL_BANK_WRAP 10000
static inline float tool_rot_dist(q4 tq, q4 eq) {   /* TG:1871-1872, TO:1587-1588 */
  q4 d = qmul(tq, qconj(eq));
  float nn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
  return 2.0f * sdx_asin(nn > 1.0f ? 1.0f : nn);
}
static inline float tool_signed_sq(float d) { return (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * (d * d); }   /* torch.sign(d) * d ** 2 */
/* pytorch3d.transforms.quaternion_to_matrix (third party, not in /root/reference; published algorithm: (r, i, j, k) = q[0..3],
 * two_s = 2 / |q|^2, the usual nine entries).  The reference hands it xyzw quaternions (TG:1853-1854), so "r" is the x component. */
static void tool_p3d_matrix(q4 q, float* M) {
  const float r = q.x, i = q.y, j = q.z, k = q.w;
  const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  M[0] = 1.0f - two_s * (j * j + k * k); M[1] = two_s * (i * j - k * r); M[2] = two_s * (i * k + j * r);
  M[3] = two_s * (i * j + k * r); M[4] = 1.0f - two_s * (i * i + k * k); M[5] = two_s * (j * k - i * r);
  M[6] = two_s * (i * k - j * r); M[7] = two_s * (j * k + i * r); M[8] = 1.0f - two_s * (i * i + j * j);
}

/* TG reset_idx's banking (TG:1436-1457), a sequential loop over the resetting envs in env order: the tool above 0.8 m, within 0.4 of
 * the fingertips (the unweighted sum of compute_observations) and within 1 rad of the plate's orientation -> (hand DoF state, tool
 * root row) into the ring of type env % 8; the index returns to 0 after slot 10000.  (The reference's eight lists alias ONE tensor,
 * TG:441-442; the rings here are separate.) */
void sdxo_tool_bank(const sdx_scene_t* S, int n, const float* brick, const float* dof, const int64_t* reset, const float* finger_dist,
                    const float* plate, float* gb_hand, float* gb_obj, int* gb_index) {
  for (int e = 0; e < n; ++e) {
    if (!reset[e]) continue;
    const int ty = e % 8;
    float row[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, TOOL_BODY, row);
    const float* pl = plate + 7 * e;
    q4 tq = {row[3], row[4], row[5], row[6]}, eq = {pl[3], pl[4], pl[5], pl[6]};
    if (row[2] > 0.8f && finger_dist[e] < 0.4f && tool_rot_dist(tq, eq) < 1.0f) {
      const int slot = gb_index[ty];
      float* hd = gb_hand + ((size_t)ty * 11024 + slot) * 46;
      const float* d = dof + (size_t)e * 72;
      for (int j = 0; j < SDX_ND; ++j) { hd[2 * j] = d[j]; hd[2 * j + 1] = d[24 + j]; }
      float* ob = gb_obj + ((size_t)ty * 11024 + slot) * 13;
      for (int q = 0; q < 13; ++q) ob[q] = row[q];
      gb_index[ty] += 1;
    }
    if (gb_index[ty] > TOOL_BANK_WRAP) gb_index[ty] = 0;
  }
}

/* reset_idx for the envs whose reset flag is set.  success_buf[:, 0] first (TG:1425-1428 rot_dist < 0.3; TO:1279-1282 < 0.5).
 * orient = 0 (TG:1459-1578): the tool to tool_reset_pos with quat_from_euler_xyz(0, k * 1.571, u * 3.14) -- k = pitch_k, ONE
 * random.sample(range(4)) per call; u per env in [-1, 1) (yaw_u, a test hook, names it; else Philox) --, velocities zero; the hand to
 * prepare_arm (= arm_hand_default_dof_pos[:7], TG:284) and the scaled finger_reset_unscaled, targets = positions; obs_buf and every
 * history frame of the env zeroed (TG:1563-1568; states_buf itself is not, but its only unrefreshed slot, 141, is never written).
 * orient = 1 (TO:1380-1436): the tool's WHOLE root row (velocities too) and the hand's DoF positions AND velocities from a banked
 * grasp (slot = Philox % per_type; slot_by_env names it), targets = positions; history kept.
 * Both: the plate to tool_plate_pose, target_init = the tool's new pose. */
void sdxo_tool_reset(const sdx_scene_t* S, int n, int orient, uint64_t seed, const float* bank_obj, const float* bank_hand, int per_type,
                     int pitch_k, const int* slot_by_env, const float* yaw_u, int do_success, float* brick, float* dof, float* plate,
                     float* target_init, int64_t* progress, int64_t* reset, float* successes, float* success_buf, int* episode, int* wsn,
                     unsigned char* slp, float* obs, float* states) {
  for (int e = 0; e < n; ++e) {
    if (!reset[e]) continue;
    float* B = brick + (size_t)e * 13 * NB;
    float* d = dof + (size_t)e * 72;
    float* pl = plate + 7 * e;
    if (do_success) {
      float tg[13];
      brick_root_row(S, B, TOOL_BODY, tg);
      q4 tq = {tg[3], tg[4], tg[5], tg[6]}, eq = {pl[3], pl[4], pl[5], pl[6]};
      float rd = tool_rot_dist(tq, eq);
      success_buf[2 * e] = rd < (orient ? 0.5f : 0.3f) ? 1.0f : 0.0f;
    }
    uint32_t r[4];
    philox(seed, (uint32_t)e, (uint32_t)episode[e], 1u, r);
    float row[13];
    if (orient) {
      int slot = slot_by_env ? slot_by_env[e] : (int)(r[0] % (uint32_t)per_type);
      const float* ob = bank_obj + (((size_t)(e % 8)) * per_type + slot) * 13;
      const float* hd = bank_hand + (((size_t)(e % 8)) * per_type + slot) * 46;
      for (int k = 0; k < 13; ++k) row[k] = ob[k];
      for (int j = 0; j < SDX_ND; ++j) { d[j] = hd[2 * j]; d[24 + j] = hd[2 * j + 1]; d[48 + j] = hd[2 * j]; }
    } else {
      float u = yaw_u ? yaw_u[e] : (float)(r[1] >> 8) * (2.0f / 16777216.0f) - 1.0f;
      float sy, cy;
      sdx_sincos((u * 3.14f) * 0.5f, &sy, &cy);
      float sp = S->tool_pitch_sc[2 * pitch_k], cp = S->tool_pitch_sc[2 * pitch_k + 1];
      row[0] = S->tool_reset_pos[0]; row[1] = S->tool_reset_pos[1]; row[2] = S->tool_reset_pos[2];
      row[3] = 0.0f - sy * sp; row[4] = cy * sp; row[5] = sy * cp; row[6] = cy * cp;
      for (int k = 7; k < 13; ++k) row[k] = 0.0f;
      for (int j = 0; j < 7; ++j) { d[j] = S->prepare_arm[j]; d[24 + j] = 0.0f; d[48 + j] = S->prepare_arm[j]; }
      for (int i = 0; i < 16; ++i) {
        float v = scalef(S->finger_reset_unscaled[i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
        d[7 + i] = v; d[24 + 7 + i] = 0.0f; d[48 + 7 + i] = v;
      }
      float* o = obs + (size_t)e * 3 * TOOL_OBS;
      for (int k = 0; k < 3 * TOOL_OBS; ++k) o[k] = 0.0f;
      float* s = states + (size_t)e * 3 * STATE_FRAME;
      for (int k = 0; k < 3 * STATE_FRAME; ++k) s[k] = 0.0f;
    }
    brick_from_root_row(S, B, TOOL_BODY, row);
    for (int k = 0; k < 7; ++k) { pl[k] = S->tool_plate_pose[k]; target_init[7 * e + k] = row[k]; }
    progress[e] = 0; reset[e] = 0; successes[e] = 0.0f;
    wsn[2 * e] = 0; wsn[2 * e + 1] = 0;
    for (int b = 0; b < NB; ++b) slp[(size_t)e * NB + b] = 0;
    episode[e] += 1;
  }
}

/* pre_physics_step after resets.  Fingers = EMA of the scaled actions.  TG (TG:1617-1636): arm = IK for (0.2 a[0:3] -- from step 60
 * on a pure 0.1 lift --, 5 x the orientation error to hand_target_quat); from step 91 on the arm goes to insert_prep0
 * (= arm_hand_insertion_prepare_dof_pos_list[0]) and the fingers hold their previous targets.  TO (TO:1471-1473): the arm holds its
 * previous target. */
void sdxo_tool_pre_physics(const sdx_scene_t* S, int n, int orient, const float* actions_in, float* actions, float* dof, const float* link,
                           const float* jac7, const int64_t* progress) {
  for (int e = 0; e < n; ++e) {
    const float* a = actions_in + 23 * (size_t)e;
    float* d = dof + (size_t)e * 72;
    float cur[23];
    for (int k = 0; k < 23; ++k) actions[23 * (size_t)e + k] = a[k];
    for (int i = 0; i < 16; ++i) {
      float t = scalef(a[7 + i], S->dof_lo[7 + i], S->dof_hi[7 + i]);
      cur[7 + i] = S->act_moving_average * t + (1.0f - S->act_moving_average) * d[48 + 7 + i];
    }
    if (orient) {
      for (int j = 0; j < 7; ++j) cur[j] = clampf(d[48 + j], S->dof_lo[j], S->dof_hi[j]);
    } else {
      const int64_t pg = progress[e];
      float dpose[6] = {a[0] * 0.2f, a[1] * 0.2f, a[2] * 0.2f, 0.0f, 0.0f, 0.0f};
      if (pg >= 60) { dpose[2] = 0.1f; dpose[0] = 0.0f; dpose[1] = 0.0f; }
      const float* hb = link + ((size_t)e * SDX_NL + 7) * 13;
      q4 want = {S->hand_target_quat[0], S->hand_target_quat[1], S->hand_target_quat[2], S->hand_target_quat[3]};
      q4 hq = {hb[3], hb[4], hb[5], hb[6]};
      v3 re = orientation_error(want, hq);
      dpose[3] = re.x * 5.0f; dpose[4] = re.y * 5.0f; dpose[5] = re.z * 5.0f;
      float u[7];
      control_ik(jac7 + 42 * (size_t)e, dpose, u);
      for (int j = 0; j < 7; ++j) cur[j] = d[j] + u[j];
      if (pg > 90) {
        for (int j = 0; j < 7; ++j) cur[j] = S->insert_prep0[j];
        for (int i = 7; i < 23; ++i) cur[i] = d[48 + i];
      }
    }
    for (int j = 0; j < 23; ++j) d[48 + j] = clampf(cur[j], S->dof_lo[j], S->dof_hi[j]);
  }
}

/* post_physics_step: progress += 1, observations, reward, reset flags.  obs [n][468], states [n][564] (UNCLAMPED task buffers; the
 * two older frames of each are the previous calls' newest, TG:1334-1336, 1366-1368).  TG writes successes (TG:1866-1868). */
void sdxo_tool_post_physics(const sdx_scene_t* S, int n, int orient, const float* brick, const float* dof, const float* link,
                            const float* actions, const float* target_init, const float* plate, int64_t* progress, int64_t* reset,
                            float* obs, float* states, float* rew, float* qcam, float* finger_dist_out, float* successes, float* consec) {
  int64_t num_resets = 0; float finished = 0.0f;
  for (int e = 0; e < n; ++e) {
    progress[e] += 1;
    const int64_t pg = progress[e];
    float* o = obs + (size_t)e * 3 * TOOL_OBS;
    float* s = states + (size_t)e * 3 * STATE_FRAME;
    for (int k = 2 * TOOL_OBS - 1; k >= 0; --k) o[TOOL_OBS + k] = o[k];
    for (int k = 2 * STATE_FRAME - 1; k >= 0; --k) s[STATE_FRAME + k] = s[k];
    const float* L = link + (size_t)e * SDX_NL * 13;
    const float* d = dof + (size_t)e * 72;
    const float* hb = L + 7 * 13;
    const float* ff = L + 11 * 13; const float* mf = L + 19 * 13; const float* rf = L + 23 * 13; const float* th = L + 15 * 13; /* TG:218-221 */
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, TOOL_BODY, tg);
    v3 tp = V3(tg[0], tg[1], tg[2]); q4 tq = {tg[3], tg[4], tg[5], tg[6]};
    v3 tip[4]; const float* fsr[4] = {ff, mf, rf, th};
    float nrm[4];
    for (int i = 0; i < 4; ++i) { /* TG:1200-1203 */
      q4 fq = {fsr[i][3], fsr[i][4], fsr[i][5], fsr[i][6]};
      tip[i] = vadd(V3(fsr[i][0], fsr[i][1], fsr[i][2]), qrot(fq, V3(0.0f, 0.0f, 1.0f * 0.04f)));
      v3 dd = vsub(tp, tip[i]); nrm[i] = sqrtf(vdot(dd, dd));
    }
    float fdist = nrm[0] + nrm[1] + nrm[2] + nrm[3]; /* TG:1221-1222 */
    finger_dist_out[e] = fdist;
    q4 hq = {hb[3], hb[4], hb[5], hb[6]}; v3 hp = V3(hb[0], hb[1], hb[2]);
    q4 cq0 = {S->cam_off_quat[0], S->cam_off_quat[1], S->cam_off_quat[2], S->cam_off_quat[3]};
    q4 cq = qmul(hq, cq0); v3 cp = vadd(qrot(hq, V3(S->cam_off_pos[0], S->cam_off_pos[1], S->cam_off_pos[2])), hp);
    q4 cqi = qconj(cq); v3 cpi = vneg(qrot(cqi, cp));
    q4 cvq = qmul(cqi, tq); v3 cvp = vadd(qrot(cqi, tp), cpi);
    qcam[4 * e] = cvq.x; qcam[4 * e + 1] = cvq.y; qcam[4 * e + 2] = cvq.z; qcam[4 * e + 3] = cvq.w;
    const float* ti = target_init + 7 * e;
    const float* pl = plate + 7 * e;
    v3 ep = V3(pl[0], pl[1], pl[2]); q4 eq = {pl[3], pl[4], pl[5], pl[6]};
    /* ---- observation frame (TG:1338-1364 = TO:1201-1232) */
    for (int j = 0; j < 23; ++j) { o[j] = unscalef(d[j], S->dof_lo[j], S->dof_hi[j]); o[23 + j] = actions[23 * (size_t)e + j]; }
    for (int k = 0; k < 7; ++k) { o[46 + k] = hb[k]; o[53 + k] = tg[k]; o[61 + k] = pl[k]; }
    o[60] = (float)pg / (float)S->max_episode_length;
    o[68] = tp.x - ep.x; o[69] = tp.y - ep.y; o[70] = tp.z - ep.z;
    { q4 r = qmul(tq, qconj(eq)); o[71] = r.x; o[72] = r.y; o[73] = r.z; o[74] = r.w; }
    for (int k = 0; k < 13; ++k) { o[75 + k] = ff[k]; o[88 + k] = rf[k]; o[101 + k] = mf[k]; o[114 + k] = th[k]; }
    for (int j = 0; j < 23; ++j) o[127 + j] = S->vel_obs_scale * d[24 + j];
    for (int k = 0; k < 6; ++k) o[150 + k] = tg[7 + k];
    /* ---- privileged frame (TG:1274-1332; TO:1137-1195 differs in 181:188); slot 141 is never written */
    for (int j = 0; j < 23; ++j) { s[j] = unscalef(d[j], S->dof_lo[j], S->dof_hi[j]); s[23 + j] = S->vel_obs_scale * d[24 + j]; }
    s[46] = tip[0].x; s[47] = tip[0].y; s[48] = tip[0].z;
    s[49] = tip[2].x; s[50] = tip[2].y; s[51] = tip[2].z;
    s[52] = tip[1].x; s[53] = tip[1].y; s[54] = tip[1].z;
    s[55] = tip[3].x; s[56] = tip[3].y; s[57] = tip[3].z;
    for (int k = 0; k < 23; ++k) s[58 + k] = actions[23 * (size_t)e + k];
    for (int k = 0; k < 7; ++k) { s[81 + k] = hb[k]; s[88 + k] = tg[k]; }
    for (int k = 0; k < 6; ++k) s[95 + k] = hb[7 + k];
    for (int k = 0; k < 4; ++k) { s[101 + k] = ff[3 + k]; s[111 + k] = mf[3 + k]; s[121 + k] = rf[3 + k]; s[131 + k] = th[3 + k]; }
    for (int k = 0; k < 6; ++k) { s[105 + k] = ff[7 + k]; s[115 + k] = mf[7 + k]; s[125 + k] = rf[7 + k]; s[135 + k] = th[7 + k]; }
    for (int k = 0; k < 6; ++k) s[142 + k] = tg[7 + k];
    s[148] = ti[0]; s[149] = ti[1]; s[150] = ti[2];
    s[151] = tp.x - ti[0]; s[152] = tp.y - ti[1]; s[153] = tp.z - ti[2];
    s[154] = hp.x - tp.x; s[155] = hp.y - tp.y; s[156] = hp.z - tp.z;
    { q4 rel = qmul(hq, qconj(tq)); s[157] = rel.x; s[158] = rel.y; s[159] = rel.z; s[160] = rel.w; }
    { v3 a = vsub(tp, tip[0]), b = vsub(tp, tip[2]), c = vsub(tp, tip[1]), dd = vsub(tp, tip[3]);
      s[161] = a.x; s[162] = a.y; s[163] = a.z; s[164] = b.x; s[165] = b.y; s[166] = b.z;
      s[167] = c.x; s[168] = c.y; s[169] = c.z; s[170] = dd.x; s[171] = dd.y; s[172] = dd.z; }
    s[173] = fdist;
    s[174] = cvp.x; s[175] = cvp.y; s[176] = cvp.z; s[177] = cvq.x; s[178] = cvq.y; s[179] = cvq.z; s[180] = cvq.w;
    if (orient) { for (int k = 0; k < 7; ++k) s[181 + k] = pl[k]; }
    else { s[181] = cvp.x; s[182] = cvp.y; s[183] = cvp.z; s[184] = cvq.x; s[185] = cvq.y; s[186] = cvq.z; s[187] = cvq.w; }
    /* ---- reward / resets */
    float dist = nrm[0] + nrm[1] + nrm[2] + 3.0f * nrm[3];
    float rd = tool_rot_dist(tq, eq);
    int64_t rs = reset[e];
    float sc = successes[e];
    if (orient) { /* TO:1574-1626 */
      if (dist >= 20.0f) rs = 1;
      if ((float)pg >= (float)S->max_episode_length - 1.0f) rs = 1;
      rew[e] = (rd < 0.2f ? 1.0f : 0.0f) + sdx_exp(-1.0f * rd);
    } else { /* TG:1741-1893 */
      float zal = tool_signed_sq(qrot(tq, V3(0.0f, 0.0f, 1.0f)).z);
      if (dist <= -1.0f) rs = 1;
      if (pg >= 150 && zal <= 0.75f) rs = 1;
      if (pg >= 150 && dist >= 0.4f) rs = 1;
      float ay = tp.y - ti[1], ax = tp.x - ti[0];
      if (pg <= 90 && (ay < 0.0f ? -ay : ay) >= 0.08f) rs = 1;
      if (pg <= 90 && (ax < 0.0f ? -ax : ax) >= 0.08f) rs = 1;
      if ((float)pg >= (float)S->max_episode_length - 1.0f) rs = 1;
      float Mi[9], Mc[9];
      q4 tiq = {ti[3], ti[4], ti[5], ti[6]};
      tool_p3d_matrix(tiq, Mi);
      tool_p3d_matrix(tq, Mc);
      float i0 = 0.0f * Mi[0] + 1.0f * Mi[3] + 0.0f * Mi[6], i1 = 0.0f * Mi[1] + 1.0f * Mi[4] + 0.0f * Mi[7], i2 = 0.0f * Mi[2] + 1.0f * Mi[5] + 0.0f * Mi[8];
      float c0 = Mc[0] * i0 + Mc[1] * i1 + Mc[2] * i2, c1 = Mc[3] * i0 + Mc[4] * i1 + Mc[5] * i2, c2 = Mc[6] * i0 + Mc[7] * i1 + Mc[8] * i2;
      float angle_difference = (0.0f * c0 + 1.0f * c1 + 0.0f * c2) - 1.0f;
      sc = -angle_difference >= 1.95f ? 1.0f : 0.0f;
      float cl = dist - 0.4f; if (cl < 0.0f) cl = 0.0f;
      float cr = rd - 0.5f; if (cr < 0.0f) cr = 0.0f;
      float up = clampf(tp.z - 0.6f, 0.0f, 0.2f);
      rew[e] = sdx_exp(-1.0f * (5.0f * cl + cr)) * (1.0f + 10.0f * up) + (rd < 0.5f ? 1.0f : 0.0f);
      successes[e] = sc;
    }
    reset[e] = rs;
    num_resets += rs; finished += sc * (float)rs;
  }
  if (num_resets > 0) consec[0] = S->av_factor * finished / (float)num_resets + (1.0f - S->av_factor) * consec[0];
}

/* ToolPositioningOrient, online t-value update (TO:1305-1316): success_buf for ALL envs from the current state; label = the column that is 1 */
void sdxo_tool_tvalue_labels(const sdx_scene_t* S, int n, const float* brick, const float* plate, float* success_buf, int* label) {
  for (int e = 0; e < n; ++e) {
    float tg[13];
    brick_root_row(S, brick + (size_t)e * 13 * NB, TOOL_BODY, tg);
    const float* pl = plate + 7 * e;
    q4 tq = {tg[3], tg[4], tg[5], tg[6]}, eq = {pl[3], pl[4], pl[5], pl[6]};
    float rd = rot_dist_sym(tq, eq);
    v3 dp = vsub(V3(pl[0], pl[1], pl[2]), V3(tg[0], tg[1], tg[2]));
    float ok = (sqrtf(vdot(dp, dp)) < 0.01f && rd < 0.1f) ? 1.0f : 0.0f;
    success_buf[2 * e] = ok; success_buf[2 * e + 1] = ok <= 0.5f ? 1.0f : 0.0f;
    label[e] = ok > 0.5f ? 0 : 1;
  }
}

/* ToolPositioningChain, compute_insertion_observations (TC:1404-1440): the frame compute_contact_observations has just written, with the
 * inner policy's last actions in 23:46 and the inner clock in slot 60, over the buffer's own history frames */
void sdxo_tool_insertion_obs(int n, const float* obs, const float* ins_actions, const int64_t* ins_progress, int ins_max_len, float* ins_obs) {
  for (int e = 0; e < n; ++e) {
    const float* o = obs + (size_t)e * 3 * TOOL_OBS;
    float* io = ins_obs + (size_t)e * 3 * TOOL_OBS;
    for (int k = 2 * TOOL_OBS - 1; k >= 0; --k) io[TOOL_OBS + k] = io[k];
    for (int k = 0; k < TOOL_OBS; ++k) io[k] = o[k];
    for (int k = 0; k < 23; ++k) io[23 + k] = ins_actions[23 * (size_t)e + k];
    io[60] = (float)ins_progress[e] / (float)ins_max_len;
  }
}
