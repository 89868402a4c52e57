/* seqdex_b200 -- C-ABI of the B200-native SeqDex hot path.
 *
 * Every entry point replaces a call the reference makes into Isaac Gym /
 * rl_games (both closed / external).  Citations are into
 * /root/reference/dexteroushandenvs:
 *   BT = tasks/hand_base/base_task.py
 *   GS = tasks/block_assembly/allegro_hand_block_assembly_grasp_sim.py
 *   VR = tasks/hand_base/vec_task_rlgames.py
 *   RGC = utils/rl_games_custom.py   TVF = policy_sequencing/terminal_value_function.py
 *
 * Plain pointers and sizes only: no torch types.  All `*_dev` pointers are
 * device pointers on the env's device; `*_host` pointers are host memory
 * (pinned for full speed).  Every function returns 0 on success, non-zero on
 * failure (message via sdx_last_error()); nothing falls back to the CPU.
 */
#ifndef SEQDEX_B200_H
#define SEQDEX_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDX_MAX_BRICKS 72
#define SDX_MAX_FIXED 60
#define SDX_NL 24          /* robot bodies after collapse_fixed_joints (GS:540-557) */
#define SDX_ND 23          /* robot DoFs */
#define SDX_MAX_RSHAPES 32
#define SDX_MAX_STATIC 80
#define SDX_ACTORS_PER_ENV 142
#define SDX_RB_PER_ENV 165 /* 24 robot + object + goal + table + 5 bin + 132 bricks + base-plate */
#define SDX_OBS_FRAME 132  /* GS:193 */
#define SDX_STATE_FRAME 188/* GS:204 */
#define SDX_STACK 3        /* GS:189 */
#define SDX_NUM_ACTIONS 23 /* GS:211 */
#define SDX_MAX_CONTACTS 1024
#define SDX_TVALUE_PARAMS 42562 /* 4->256->128->64->2 (TVF:30-46) */
#define SDX_GRASP_BANK 11024    /* 10000+1024 rows per brick type (GS:391-394) */

/* Constant scene tables (seqdex_b200/scene.py builds them; layout is ABI). */
typedef struct sdx_scene_t {
  int n_bricks, n_fixed, n_rshapes, n_static;
  int substeps, iters, max_episode_length, sleep_substeps;   /* sleep_substeps: quiet sub-steps before a brick sleeps; 0 = never */
  float dt, gravity_z, contact_offset, friction, baumgarte, slop, max_depen_vel, brick_ang_damp, max_ang_vel,
      max_lin_vel, brick_lin_damp, sleep_energy;   /* sleep_energy: mass-normalised kinetic energy below which a brick is quiet */
  float base_pos[3], base_quat[4];
  float face_margin;   /* a sample point counts as over the reference face up to this far beyond its edge (NOT the speculative contact_offset) */
  int body_parent[SDX_NL];
  unsigned link_anc_mask[SDX_NL];            /* bit j: DoF j moves link */
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
  float act_moving_average, av_factor, vel_obs_scale, warm_start, wake_energy;   /* wake_energy: energy above which a brick wakes what it touches */
  int task;                   /* SDX_TASK_*: which task's pre-physics / observation / reward / reset ops run around the contact step */
  float hand_target_quat[4];  /* Orient: quat_from_euler_xyz(target_euler = (0, 3.1415, 1.571)) the arm IK tracks (OR:484, 1738) */
  int bank_sample_range;      /* Orient: reset samples heap rows [0, range) of the bank (OR:1564 env_rand_range = range(0, 500)) */
  int pad3[2];
  float default_dof[SDX_ND];  /* Search: arm_hand_default_dof_pos, the pose that parks the hand beside the bin (SE:207-211) */
  float prepare_dof[SDX_ND];  /* Search: arm_hand_prepare_dof_pos_list[0], where an episode starts (SE:220-223, 316) */
  float insert_plate_zw[2];   /* InsertSim: (z, w) of gymapi.Quat.from_euler_zyx(0, 0, 1.57), the base-plate's second yaw (IS:1436-1437) */
  int st_mod[SDX_MAX_STATIC], st_rem[SDX_MAX_STATIC];   /* static s exists only in envs with env % st_mod == st_rem (st_mod 0: in every env);
                                                            InsertSim's base-plate is 4x4x{1,2,4} by env % 3 (IS:971-977) */
  /* COMPOUND free bodies: body b (n_bricks of them: state, mass, inertia) is the union of the collision boxes a with bs_body[a] == b
   * (n_bshapes boxes, those of one body consecutive; half extents br_half[a], centre bs_c[a] in the body's COM frame, axes = the body's).
   * n_bshapes == 0 (every BlockAssembly scene): each body is ONE box -- n_bricks boxes, box a = body a, centred on its COM. */
  int n_bshapes;
  int bs_body[SDX_MAX_BRICKS];
  float bs_c[SDX_MAX_BRICKS * 3];
  /* ToolPositioningGrasp / Orient (TG = tasks/tool_positioning/allegro_hand_tool_positioning_grasp.py, TO = ..._orient.py) */
  float tool_reset_pos[3];    /* where reset_idx puts the tool: (0.29, 0.19, 0.675) (TG:1496-1498) */
  float tool_pitch_sc[8];     /* (sin, cos) of k * 1.571 / 2, k = 0..3: the tool's pitch at reset is target_rot_rand * 1.571 (TG:1493-1495) */
  float tool_plate_pose[7];   /* the "extra lego" the tool's orientation is measured against: (0.25, -0.2, 0.618), Quat.from_euler_zyx(0, 3.1415, 0) (TG:1505-1512) */
  /* EDGE-EDGE contacts (k_simulate's narrow phase): edge_contacts > 0.5 completes the corner-vs-face test with the separating-axis test
   * over the nine edge-pair axes; a pair whose axis of least overlap is an edge pair -- by more than edge_pref [m] over every face axis --
   * gets one contact at the closest points of the two edges, normal = that axis */
  float edge_contacts, edge_pref;
  /* warm start of a contact that involves a robot link or a HOT brick (touched by the robot or faster than the wake threshold in the
   * previous sub-step): such contacts change too fast for their last impulse to be trusted at warm_start (resting contacts: 0.98) */
  float warm_start_hot;
} sdx_scene_t;

/* tasks sharing the scene, the contact step and the PPO engine (SURVEY.md section 8a "per-task dimensions"):
 *   OR = tasks/block_assembly/allegro_hand_block_assembly_orient.py */
#define SDX_TASK_GRASP_SIM 0   /* BlockAssemblyGraspSim: obs 132 x 3, states 188 x 3, episode 150 (GS:191-211) */
#define SDX_TASK_ORIENT 1      /* BlockAssemblyOrient:   obs  62 x 3, states 188 x 3, episode  75 (OR:189-214)  */
#define SDX_TASK_SEARCH 2      /* BlockAssemblySearch:   obs  62 x 3, states 188 x 3, episode  75 (SE:149-175); BASELINE configs[0] */
#define SDX_TASK_INSERT_SIM 3  /* BlockAssemblyInsertSim: obs 75 x 1, states 188 x 1, episode 125 (IS:172-193); last link of configs[3] */
#define SDX_TASK_TOOL_GRASP 4  /* ToolPositioningGrasp:   obs 156 x 3, states 188 x 3, episode 150 (TG:224-243); BASELINE configs[4] */
#define SDX_TASK_TOOL_ORIENT 5 /* ToolPositioningOrient:  obs 156 x 3, states 188 x 3, episode 125 (TO:170-189); BASELINE configs[4] */
#define SDX_INSERT_OBS_FRAME 75
#define SDX_TOOL_OBS_FRAME 156
#define SDX_TOOL_BANK_WRAP 10000     /* TG:1453-1454: the grasp ring's index returns to 0 after slot 10000 */
#define SDX_ORIENT_OBS_FRAME 62
#define SDX_ORIENT_BANK_WRAP 10000   /* OR:1478-1479: ring index returns to 0 after slot 10000 */

/* Tensor kinds for sdx_tensor(): device buffers owned by the env. dtype 0=f32 1=i64 2=i32 3=u8 */
enum {
  SDX_T_BRICK = 0,      /* f32 [N][13][72]  free-brick COM state, SoA inside an env block            */
  SDX_T_DOF = 1,        /* f32 [N][3][24]   q | qd | position target (GS:313-316, 328-329)           */
  SDX_T_LINK = 2,       /* f32 [N][24][13]  robot rigid-body rows (rigid_body_states GS:318)         */
  SDX_T_JAC7 = 3,       /* f32 [N][6][7]    jacobian_tensor[:, link7-1, :, :7] (GS:1601)             */
  SDX_T_NETF = 4,       /* f32 [N][24][3]   net contact force on robot links (GS:1159)               */
  SDX_T_ACTIONS = 5,    /* f32 [N][23]                                                              */
  SDX_T_OBS = 6,        /* f32 [N][396]  obs_buf (BT:57); [N][186] for SDX_TASK_ORIENT                */
  SDX_T_STATES = 7,     /* f32 [N][564]  states_buf (BT:59)                                          */
  SDX_T_REW = 8,        /* f32 [N]       rew_buf                                                     */
  SDX_T_RESET = 9,      /* i64 [N]       reset_buf (BT:63)                                           */
  SDX_T_PROGRESS = 10,  /* i64 [N]       progress_buf                                                */
  SDX_T_TVALUE = 11,    /* f32 [N]       sigmoid(t_value(q_cam))[:,1] (GS:1200-1201)                  */
  SDX_T_TARGET_INIT = 12,/* f32 [N][7]   segmentation_target_init_{pos,rot} (GS:1547-1548)           */
  SDX_T_SUCCESSES = 13, /* f32 [N]                                                                   */
  SDX_T_CONSEC = 14,    /* f32 [1]       consecutive_successes                                       */
  SDX_T_NCONTACT = 15,  /* i32 [N][4]    contacts in the last sub-step | contacts beyond SDX_MAX_CONTACTS after shedding | most shedding of
                           speculative contacts any sub-step of the step needed (0 none .. 3 touching contacts only) | candidate pairs
                           beyond the per-owner cap (low 16 bits; of those, pairs against statics: high 16 bits, always 0) */
  SDX_T_ROOT = 16,      /* f32 [N*142][13] actor_root_state_tensor, filled by sdx_refresh            */
  SDX_T_RB = 17,        /* f32 [N*165][13] rigid_body_state_tensor, filled by sdx_refresh            */
  SDX_T_DOF_STATE = 18, /* f32 [N*23][2]   dof_state_tensor, filled by sdx_refresh                    */
  SDX_T_JACOBIAN = 19,  /* f32 [N][23][6][23] jacobian tensor, filled by sdx_refresh                  */
  SDX_T_EPISODE = 20,   /* i32 [N]       per-env episode counter (keys the reset RNG)                */
  SDX_T_CONTACTS = 21,  /* f32 [N][SDX_MAX_CONTACTS][8] debug dump of the last sub-step's contacts   */
  SDX_T_WS = 22,        /* f32 [N][2][SDX_MAX_CONTACTS][4] contact-impulse cache (key bits, f.xyz), double buffered  */
  SDX_T_WSN = 23,       /* i32 [N][2]    entries in each cache buffer                                           */
  SDX_T_SLEEP = 24,     /* u8  [N][72]   sub-steps since each free brick was last hot (0 = hot; >= sleep_substeps = asleep)  */
  SDX_T_SEG = 25,       /* i32 [N][3]    Search: pixels showing the target | centre row | centre column of the last render (SE:1231-1241) */
  SDX_T_EMERGENCE = 26, /* f32 [N]       Search: emergence reward = 5 x (pixels now - pixels at the last render) (SE:1640-1646)   */
  SDX_T_TVOBS = 27,     /* f32 [N][650]  Search: the transition-feasibility gate's input, 10 frames x 65 (SE:400, 1154-1166)      */
  SDX_T_PLATE = 28,     /* f32 [N][7]    InsertSim / tool tasks: root pose of the base-plate ("extra lego", IS:1438-1446, TG:1505-1512)      */
  SDX_T_ROT_ERR = 29,   /* f32 [N][3]    InsertSim: wrist orientation error of the last pre_physics_step (IS:1531), read by the reward      */
  SDX_T_SUCCESS = 30,   /* f32 [N][2]    InsertSim / tool tasks: success_buf written at reset (IS:1348-1350; TG:1428, TO:1282 column 0 only) */
  SDX_T_COUNT = 31
};

/* pinhole camera of the segmentation features (see sdx_segmentation_features) */
typedef struct sdx_camera_t { float pos[3], fwd[3], right[3], up[3]; float inv_focal; int width, height; } sdx_camera_t;
typedef struct sdx_env sdx_env_t;

const char* sdx_last_error(void);

/* gym.create_sim + _create_envs + prepare_sim (BT:122-126, GS:505-523, BT:84). */
int sdx_create(const sdx_scene_t* scene, int num_envs, int device, uint64_t seed, sdx_env_t** out);
void sdx_destroy(sdx_env_t* env);
/* CUDA stream all subsequent launches go to (cudaStream_t as void*). */
int sdx_set_stream(sdx_env_t* env, void* stream);
int sdx_num_envs(const sdx_env_t* env);

/* acquire_*_tensor + gymtorch.wrap_tensor (GS:237-246): zero-copy device view. */
int sdx_tensor(sdx_env_t* env, int kind, void** dev_ptr, int64_t shape[4], int* ndim, int* dtype);
/* refresh_{actor_root_state,rigid_body_state,dof_state,jacobian}_tensor (GS:1091-1095). */
int sdx_refresh(sdx_env_t* env, int kind);

/* set_actor_root_state_tensor_indexed (GS:1514): rows of the facade root tensor named by
 * sim-domain actor indices (int32) are written back into the simulation state. */
int sdx_set_actor_root_state_indexed(sdx_env_t* env, const float* root_dev, const int32_t* actor_idx_dev, int n);
/* set_dof_state_tensor_indexed / set_dof_position_target_tensor_indexed (GS:1539-1545):
 * indices are sim-domain actor indices of the hand actors (= env*142). */
int sdx_set_dof_state_indexed(sdx_env_t* env, const float* dof_state_dev, const int32_t* actor_idx_dev, int n);
int sdx_set_dof_target_indexed(sdx_env_t* env, const float* targets_dev, const int32_t* actor_idx_dev, int n);
/* set_dof_position_target_tensor (GS:1638). targets [N][23]. */
int sdx_set_dof_targets(sdx_env_t* env, const float* targets_dev);

/* Terminal-state heap bank the task samples on reset: the stand-in for
 * saved_searching_ternimal_states_*.pkl = list[8] of [B,132,13] (GS:412-413, 1507-1511).
 * bank_host: f32 [8][per_type][72][13] root-frame rows of the 72 free bricks. */
int sdx_set_heap_bank(sdx_env_t* env, const float* bank_host, int per_type);
/* GraspInsertTValue parameters, torch state_dict order: W1[256][4] b1 W2[128][256] b2 W3[64][128] b3 W4[2][64] b4. */
int sdx_set_tvalue_weights(sdx_env_t* env, const float* weights_host);
/* Put every env into the scene's initial state (lattice of bricks GS:737-742, robot at the
 * prepare pose GS:267-272), progress 0, reset 1 (BT:63). */
int sdx_reset_all(sdx_env_t* env);

/* The three phases of BaseTask.step (BT:130-150). */
int sdx_pre_physics(sdx_env_t* env, const float* actions_dev); /* GS:1555-1638 incl. reset_idx GS:1361-1553 */
int sdx_simulate(sdx_env_t* env);                              /* gym.simulate BT:140 */
int sdx_post_physics(sdx_env_t* env);                          /* GS:1640-1645: obs, states, reward, reset  */
int sdx_step(sdx_env_t* env, const float* actions_dev);        /* all three */
/* RLgamesVecTaskPython.step with host buffers (VR:165-177): H2D actions, step, clamp +-5, D2H results. */
int sdx_step_host(sdx_env_t* env, const float* actions_host, float* obs_host, float* states_host, float* rew_host,
                  int64_t* reset_host);
/* settle the heap: n steps of sdx_simulate with targets frozen (bank generation). */
int sdx_simulate_n(sdx_env_t* env, int n);
/* number of kernels this library has launched on behalf of env since creation */
int64_t sdx_launch_count(const sdx_env_t* env);

/* dst = clamp(tensor(kind), -lim, lim): VecTask's clip_obs into caller memory (VR:174-175) */
int sdx_clamped_copy(sdx_env_t* env, int kind, float* dst_dev, float lim);
/* Domain randomisation, non-physical half.  BaseTask.step adds noise to the actions before pre_physics_step
 * (BT:131-132) and to obs_buf after post_physics_step (BT:149-150); apply_randomizations prepares its parameters (BT:263-340).
 *   sdx_dr_randn : dst[i] ~ N(0,1) -- the correlated-noise tensor, redrawn when the parameters are refreshed (BT:293-296)
 *   sdx_dr_noise : dst = op(src, (corr * a_corr + b_corr) + white * a + b)
 *                  distribution 0 gaussian (a_corr, b_corr, a, b = var_corr, mu_corr, var, mu; white ~ N(0,1)),
 *                               1 uniform  (hi_corr - lo_corr, lo_corr, hi - lo, lo;           white ~ U[0,1));
 *                  operation 0 additive, 1 scaling.  White noise: Philox(seed, element / 4, counter) -- pass a fresh counter per call.
 *   sdx_set_gravity : sim_params.gravity (BT:342-355), z component, effective from the next sdx_simulate. */
int sdx_dr_randn(sdx_env_t* env, float* dst_dev, int64_t n, uint64_t seed, uint32_t counter);
int sdx_dr_noise(sdx_env_t* env, float* dst_dev, const float* src_dev, const float* corr_dev, int64_t n, float a_corr, float b_corr,
                 float a, float b, int distribution, int operation, uint64_t seed, uint32_t counter);
int sdx_set_gravity(sdx_env_t* env, float gravity_z);
/* rows [142][13] of the actors that carry no simulation state (hand base, table, bin, fixed bricks, ...) for the facade */
int sdx_set_static_rows(sdx_env_t* env, const float* rows_host);
/* sdx_set_heap_bank from a device buffer (bank synthesised on the GPU) */
int sdx_set_heap_bank_dev(sdx_env_t* env, const float* bank_dev, int per_type);
/* device pointers of the grasp terminal-state rings reset_idx fills (GS:1399-1445): hand [8][11024][23][2], obj [8][11024][13], index [8] */
int sdx_grasp_bank(sdx_env_t* env, void** hand_dev, void** obj_dev, void** index_dev);
/* t-value training data (GS:1402-1438 under save_hdf5; read back by TVT:132-168): capacity > 0 allocates two device rings
 * succ/fail [capacity][4] f32 + counts i64[2] (rows ever written) and records from the next reset on; 0 stops recording */
int sdx_tvalue_dataset(sdx_env_t* env, int capacity, void** succ_dev, void** fail_dev, void** counts_dev);
int sdx_aux(sdx_env_t* env, void** qcam_dev, void** finger_dist_dev);
/* Orient (OR:1390-1695).  reset_idx there is not a per-env scatter: it drives the WHOLE sim through a scripted sequence --
 * 50 steps lifting the hand above the target (OR:1430-1458), one extra compute_observations + banking of the re-oriented heaps
 * (OR:1460-1511), the state reset (OR:1523-1605) and post_reset's 2 + 1 + 50 settle / approach steps (OR:1612-1695).
 * sdx_pre_physics runs that sequence itself when any reset flag is set (one 4-byte device->host read per step, the
 * reference's reset_buf.nonzero()).  The entry points below expose its two data products:
 * terminal heaps of envs whose brick ended face up (saved_digging_ternimal_states_list -> ..._good_mo_tvalue.pkl, OR:1465-1513):
 * capacity > 0 allocates rings [8][capacity + 1][72][13] f32 (free-brick root rows) + index i32[8] and records from then on */
int sdx_orient_heap_bank(sdx_env_t* env, int capacity, void** rows_dev, void** index_dev);
/* Search (SE = tasks/block_assembly/allegro_hand_block_assembly_search.py).  Its reset_idx settles the freshly dropped heap for 60
 * contact steps and renders the overview camera (SE:1435-1455); at the end of an episode compute_observations parks the hand,
 * steps once and renders again (SE:989-1019).  sdx_pre_physics / sdx_post_physics run both; the camera is set here
 * (gym.create_camera_sensor + set_camera_location, SE:873-875).  sdx_search_bank: the heaps whose target brick became visible
 * enough, with the hand's DoF state (saved_searching_{,hand_}ternimal_states_list -> ..._medium_mo_tvalue.pkl, SE:1305-1352):
 * capacity > 0 allocates rings rows [8][capacity + 1][72][13], hand [8][capacity + 1][23][2] f32, index i32[8] */
int sdx_set_camera(sdx_env_t* env, const sdx_camera_t* cam);
int sdx_search_bank(sdx_env_t* env, int capacity, void** rows_dev, void** hand_dev, void** index_dev);
/* InsertSim (IS = tasks/block_assembly/allegro_hand_block_assembly_insert_sim.py): the banked grasps its reset_idx restores
 * (saved_grasping_{object,hand}_ternimal_states_*.pkl, IS:372-375, 1449-1453): obj [8][per_type][13] root rows of the grasped brick,
 * hand [8][per_type][23][2] DoF states.  Device or host pointers (is_device). */
int sdx_set_grasp_bank(sdx_env_t* env, const float* hand, const float* obj, int per_type, int is_device);
/* test hook of the InsertSim parity tests: the bank slot every env restores on its next resets (NULL: drawn from Philox) and the
 * base-plate yaw index of the next reset_idx calls (-1: drawn) */
int sdx_insert_test_hooks(sdx_env_t* env, const int* slot_by_env_host, int plate_yaw);
/* ToolPositioningGrasp / ToolPositioningOrient (BASELINE configs[4]; TG, TO as above).  One free body, the tool (a compound of boxes,
 * n_bshapes > 0), the same arm and hand, the same 188-slot privileged frame; observation frame 156 x 3.
 *   TG: sdx_pre_physics = banking of good grasps into the sdx_grasp_bank rings (tool above 0.8 m, fingers within 0.4, within 1 rad of
 *       the plate's orientation; ring wraps after slot SDX_TOOL_BANK_WRAP; TG:1436-1457) -> reset_idx (tool to tool_reset_pos with pitch
 *       k x 1.571, k one draw per call, and a per-env yaw; hand to prepare_arm + finger_reset_unscaled; observation history zeroed;
 *       TG:1412-1578) -> targets (arm IK on 0.2 a[0:3], lift from step 60, parked at insert_prep0 from step 91; TG:1580-1675).
 *   TO: sdx_pre_physics = reset_idx restoring a banked grasp, tool root row INCLUDING its velocities and the hand's DoF positions and
 *       velocities (sdx_set_grasp_bank; TO:1265-1436) -> targets (fingers only; the arm holds its previous target, TO:1438-1509).
 * SDX_T_PLATE holds the plate pose (extra_target_{pos,rot}), SDX_T_SUCCESS[:, 0] success_buf (TG:1428, TO:1282).
 * sdx_tool_test_hooks: parity-test hook like sdx_insert_test_hooks -- bank slot per env (NULL: drawn), pitch index k of the next
 * reset_idx calls (-1: drawn), yaw draw u in [-1, 1) per env (NULL: drawn). */
int sdx_tool_test_hooks(sdx_env_t* env, const int* slot_by_env_host, int pitch_k, const float* yaw_u_host);
/* ToolPositioningOrient's online t-value update (TO:1305-1350; `if_t_value`, hard-wired False at TO:377, so opt-in here): the labels
 * reset_idx computes for ALL envs from their current state -- success = the tool within 1 cm of the plate's position and within 0.1 rad
 * of its orientation or that orientation turned by pi about z (TO:1306-1316).  Writes SDX_T_SUCCESS = [success, not success] and
 * label_dev i32[N] = 0 (success) / 1 (failure), the class index sdx_tvalue_bce takes.  The rows are SDX_T_TARGET_INIT
 * (t_value_obs_buf = the pose each episode started from, TO:1400); the five Adam steps run on sdx_mlp_* (tasks/tool_positioning.py). */
int sdx_tool_tvalue_labels(sdx_env_t* env, int* label_dev);
/* ToolPositioningChain's inner-policy call (TC = tasks/tool_positioning/allegro_hand_tool_positioning_chain.py:1733-1768): at step 118 of
 * env 0's clock pre_physics_step runs 125 steps of a FROZEN policy inside the outer step -- predict on insertion_obs_buf, fingers = the
 * scaled actions, arm = its previous target, gym.simulate.  sdx_tool_inner_step is one such step (actions -> targets -> contact step; the
 * task's own ACTIONS tensor, observations and reward are untouched); sdx_tool_insertion_obs is compute_insertion_observations
 * (TC:1404-1440): the observation frame just written to SDX_T_OBS with the inner policy's last actions in 23:46 and the inner episode
 * clock (ins_progress / ins_max_len) in slot 60, shifted into the caller's own [N][468] history buffer. */
int sdx_tool_inner_step(sdx_env_t* env, const float* actions_dev);
int sdx_tool_insertion_obs(sdx_env_t* env, const float* ins_actions_dev, const int64_t* ins_progress_dev, int ins_max_len, float* ins_obs_dev);
/* number of contact steps the last sdx_pre_physics spent inside reset_idx (0 when nobody reset; 103 for a full Orient reset) */
int sdx_last_reset_sim_steps(const sdx_env_t* env);
/* BlockAssemblySearch's camera features (SE = tasks/block_assembly/allegro_hand_block_assembly_search.py): the reference renders
 * a 128x128 segmentation image per env (gym.create_camera_sensor / set_camera_location / get_camera_image_gpu_tensor
 * IMAGE_SEGMENTATION, SE:755-758, 873-878) and keeps three integers of it -- the number of pixels that show the target brick
 * and the centroid (row, column) of those pixels (SE:1231-1241, 1640-1646).  sdx_segmentation_features computes exactly those
 * by ray casting against the scene's boxes.  The camera is a pinhole: unit vectors fwd / right / up (right = fwd x world-up,
 * up = right x fwd), inv_focal = tan(horizontal_fov / 2) / (width / 2) (Isaac Gym's default horizontal_fov is 90 degrees). */
int sdx_segmentation_features(sdx_env_t* env, const sdx_camera_t* cam, int32_t* out_dev /* [N][3]: pixels, centre row, centre column */);
int sdx_scene_size(void);
int sdx_sim_smem_bytes(void);

/* ---- PPO (rl_games 1.5.2 semantics; RGC:1394-1483, 1767-1911, 2115-2132) ---- */
/* discount_values: GAE sweep. rewards/values/dones [H][N], last_values/last_dones [N] -> advantages [H][N]. */
int sdx_gae(const float* rewards_dev, const float* values_dev, const float* dones_dev, const float* last_values_dev,
            const float* last_dones_dev, float* adv_dev, float* returns_dev, int horizon, int n, float gamma,
            float tau, void* stream);

/* bf16 tensor-core GEMM D[M,N] = A[M,K] . B[N,K]^T (both K-major); fused epilogues by `mode` (csrc/sdx_gemm.cuh):
 * 0 ELU(acc+bias)->bf16 (+transposed copy)  1 acc*ELU'(h)->bf16 (+transposed)  2 fp32 += acc (split-K)  3 fp32 = acc
 * 4 fp32 = acc + bias.  Replaces the cuBLAS calls under rl_games' network forward/backward (RGC:1697-1723, 1767-1911). */
int sdx_gemm_bf16_tn(int mode, const void* A, int M, int K, int lda, const void* B, int N, int ldb, const float* bias,
                     const void* h, int ldh, void* out, int ldo, void* out_t, int ldt, float* outf, int ldf, int splits,
                     void* stream);

/* MLP in -> 1024 -> 512 -> 256 -> out (ELU), fp32 master parameters in torch state_dict order, bf16 compute
 * (cfg/lego/ppo_continuous_grasp.yaml:21-23, 74-95).  has_sigma appends the logstd vector (fixed_sigma: True). */
typedef struct sdx_mlp sdx_mlp_t;
int sdx_mlp_create(int in_dim, int out_dim, int max_rows, int has_sigma, sdx_mlp_t** out);
int sdx_mlp_create_ex(int in_dim, int out_dim, int h1, int h2, int h3, int max_rows, int has_sigma, sdx_mlp_t** out);
/* t-value trainer loss (TVT:199-201,226): BCEWithLogits(ELU(z), onehot(label)), mean over M x 2; dz through the ELU */
int sdx_tvalue_bce(const float* z, const int* label, int M, float* dz, float* stats, void* stream);
void sdx_mlp_destroy(sdx_mlp_t* m);
int sdx_mlp_info(sdx_mlp_t* m, int64_t* nparams, void** params, void** grads, void** out, void** adam_m, void** adam_v);
int sdx_mlp_sync(sdx_mlp_t* m, void* stream);
int sdx_mlp_forward(sdx_mlp_t* m, const float* x_dev, int M, const float* mean, const float* var, int train, void* stream);
/* whole-batch input conversion (once per PPO iteration) and forward on a row range of it */
int sdx_mlp_convert_batch(sdx_mlp_t* m, const float* x_dev, int B, const float* mean, const float* var, void* xb_bf16, void* xt_bf16, void* stream);
/* x_dev is a TIME-major rollout buffer [horizon][B / horizon][in]; rows of the converted batch are ENV-major
 * (rl_games' swap_and_flatten01, RGC:1480-1481): row n * horizon + t <- x[t][n] */
int sdx_mlp_convert_batch_env_major(sdx_mlp_t* m, const float* x_dev, int B, int horizon, const float* mean, const float* var, void* xb_bf16,
                                    void* xt_bf16, void* stream);
int sdx_mlp_forward_pre(sdx_mlp_t* m, const void* xb_bf16, const void* xt_bf16, int B, int row0, int M, int train, void* stream);
int sdx_mlp_backward(sdx_mlp_t* m, const float* dout_dev, int M, void* stream);
/* the same gradients published LAYER BY LAYER (output layer first): layer l's slice of the flat gradient vector is final right after its dW
 * GEMM, an event marks it, and sdx_mlp_wait_layer lets another stream wait for it -- the data-parallel caller all-reduces layer l over NCCL
 * while the layers below are still being differentiated (rl_games' multi_gpu path all-reduces after the whole backward, RGC:1860-1870) */
int sdx_mlp_backward_pipelined(sdx_mlp_t* m, const float* dout_dev, int M, void* stream);
int sdx_mlp_wait_layer(sdx_mlp_t* m, int layer, void* waiting_stream);
int sdx_mlp_layer_range(sdx_mlp_t* m, int layer, int64_t* begin, int64_t* end);
int sdx_mlp_adam(sdx_mlp_t* m, float lr, float b1, float b2, float eps, float max_norm, void* stream);   /* RGC:1102, 1866-1872 */
/* same step with the learning rate read from device memory when the kernel runs (no host round trip for the adaptive schedule) */
int sdx_mlp_adam_dev(sdx_mlp_t* m, const float* lr_dev, float b1, float b2, float eps, float max_norm, void* stream);
/* rl_games' AdaptiveScheduler (schedule_type 'legacy' = after EVERY minibatch, RGC:1360-1365) on the device: kl = stats[2] * inv_count;
 * kl > 2 thr: lr = max(lr / 1.5, lr_min); kl < 0.5 thr: lr = min(lr * 1.5, lr_max).  stats[0..4) are then added to accum[0..4)
 * (accum[4] counts minibatches, accum[5] = the last kl) and cleared.  adaptive = 0 only moves the statistics. */
int sdx_ppo_adaptive_lr(float* stats_dev, float inv_count, float kl_threshold, float lr_min, float lr_max, float* lr_dev,
                        float* accum_dev, int adaptive, void* stream);
/* the gradient buffer sdx_mlp_info returns has SDX_GRAD_TAIL extra floats behind the nparams gradients: loss statistics written
 * there are summed by the same all-reduce as the gradients (RGC:1360-1363 averages the KL over ranks before the scheduler) */
#define SDX_GRAD_TAIL 16
/* Adam step counter of the MLP's optimiser (torch.optim.Adam state['step'], saved in rl_games checkpoints under 'optimizer',
 * RGC:1913-1933): set < 0 reads it, set >= 0 overwrites it (restore) */
long long sdx_mlp_adam_step(sdx_mlp_t* m, long long set);
long long sdx_ppo_launch_count(void);
/* a CUDA graph captured from the sdx_mlp_* / sdx_ppo_* entry points replays their kernels without passing through them: the caller adds
 * the number of kernels one replay launches (counted during the capture) so that sdx_ppo_launch_count stays the number actually launched */
void sdx_ppo_add_launches(long long n);

/* a = mu + exp(logstd) N(0,1) (Philox), neglogp (RGC:2115-2127) */
int sdx_ppo_sample(const float* mu, const float* logstd, int M, int A, uint64_t seed, uint32_t counter, float* actions,
                   float* neglogp, void* stream);
/* clipped surrogate + bounds loss + KL (RGC:1767-1911, 2130-2132): dmu [M,A], dlogstd [A] (+=), stats[4] (+=) */
int sdx_ppo_actor_loss(const float* mu, const float* logstd, const float* actions, const float* old_mu,
                       const float* old_logstd, const float* old_neglogp, const float* adv, int M, int A, float e_clip,
                       float bounds_coef, float inv_batch, float* dmu, float* dlogstd, float* stats, void* stream);
int sdx_ppo_value_loss(const float* v, const float* v_old, const float* ret, int M, float e_clip, int clip_value,
                       float scale, float* dv, float* stats, void* stream);
/* advantage normalisation (RGC:1651): moments -> (x - mean) / (std + 1e-8) */
int sdx_moments(const float* x, int64_t n, double* mom, void* stream);
int sdx_normalize(float* x, int64_t n, const double* mom, double count, void* stream);
/* RunningMeanStd of the central-value input (yaml:80 normalize_input) */
int sdx_col_moments(const float* x, int B, int D, double* colmom, void* stream);
int sdx_rms_merge(float* mean, float* var, double* count, const double* colmom, int D, double bcount, void* stream);

#ifdef __cplusplus
}
#endif
#endif
