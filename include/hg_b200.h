/* hg_b200.h -- C ABI of libhg_b200.so: the B200 (sm_100a) kernels behind the
 * humanoid-gym `humanoid_ppo` hot path.
 *
 * Conventions
 *   - plain C: POD structs, raw device pointers, sizes; no C++/torch types.
 *   - every buffer is allocated and owned by the caller (torch tensors in the
 *     Python host layer); the library never allocates or frees user data.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - return value: 0 ok, <0 argument error (see HG_E_*), >0 a cudaError_t.
 *     hg_last_error() returns a thread-local message for the last failure.
 *   - all floating point data is fp32, row-major, env-major (N, ...).
 *
 * Each entry cites the reference function it replaces; paths are relative to
 * roboterax/humanoid-gym `humanoid/`.
 */
#ifndef HG_B200_H
#define HG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_VERSION 200
#define HG_NUM_DOF 12
#define HG_NUM_REWARDS 22
#define HG_OBS1 47          /* single-frame observation width   */
#define HG_PRIV1 73         /* single-frame privileged obs width */
#define HG_OBS_FRAMES 15
#define HG_PRIV_FRAMES 3
#define HG_MAX_CONTACT_BODIES 4
#define HG_MAX_LAYERS 8

#define HG_E_NULL (-1)      /* required pointer is NULL           */
#define HG_E_ALIGN (-2)     /* pointer not 16-byte aligned        */
#define HG_E_SIZE (-3)      /* N <= 0 or inconsistent dimensions  */
#define HG_E_ARG (-4)       /* other invalid argument             */
#define HG_E_STATE (-5)     /* call sequence error                */

/* ------------------------------------------------------------------------ */
/* Environment side                                                         */
/* ------------------------------------------------------------------------ */

/* Constants of XBotLCfg as LeggedRobot._parse_cfg / _init_buffers derive them
 * (envs/base/legged_robot.py:434-541,710-720; envs/custom/humanoid_config.py). */
typedef struct HgEnvParams {
    float dt;                          /* float32(decimation * sim_params.dt)          */
    float cycle_time;
    float clip_actions, clip_obs, action_scale;
    float action_delay, action_noise;
    /* uniform ranges as (lo, span): torch_rand_float(lo,hi) = (hi-lo)*u + lo with the
     * span (hi-lo) formed in DOUBLE by the host, then rounded to fp32, as Python does  */
    float cmd_x_lo, cmd_x_span, cmd_y_lo, cmd_y_span, cmd_heading_lo, cmd_heading_span;
    float push_vel_lo, push_vel_span, push_ang_lo, push_ang_span;
    float dof_reset_lo, dof_reset_span;   /* q = q0 + U(-0.1, 0.1) on reset             */
    float target_joint_pos_scale, target_feet_height, base_height_target;
    float min_dist, max_dist, max_contact_force, tracking_sigma;
    float obs_scale_lin_vel, obs_scale_ang_vel, obs_scale_dof_pos, obs_scale_dof_vel, obs_scale_quat;
    float noise_level;
    float max_episode_length_s;
    int32_t add_noise, only_positive_rewards, heading_command, push_robots;
    int32_t resample_period;           /* int(resampling_time / dt) = 799              */
    int32_t push_interval;             /* ceil(push_interval_s / dt) = 400             */
    int64_t max_episode_length;        /* ceil(episode_length_s / dt) = 2400           */
    int32_t num_bodies;                /* 13 after fixed-joint collapse                */
    int32_t feet[2], knees[2];
    int32_t n_term, term_bodies[HG_MAX_CONTACT_BODIES];
    int32_t n_pen, pen_bodies[HG_MAX_CONTACT_BODIES];
    float reward_scales[HG_NUM_REWARDS];   /* scale*dt, alphabetical term order        */
    float p_gains[HG_NUM_DOF], d_gains[HG_NUM_DOF], torque_limits[HG_NUM_DOF];
    float default_dof_pos[HG_NUM_DOF];
    float noise_scale_vec[HG_OBS1];
    float base_init_state[13];
} HgEnvParams;

/* Live state tensors of LeggedRobot / XBotLFreeEnv. */
typedef struct HgEnvBuffers {
    /* physics tensors (Isaac Gym tensor-API layout) */
    float* root_states;        /* (N,13) pos3 quat_xyzw4 linvel3 angvel3           */
    float* dof_state;          /* (N,12,2) interleaved (pos, vel)                  */
    const float* contact_forces;   /* (N,num_bodies,3)                             */
    const float* rigid_state;      /* (N,num_bodies,13)                            */
    /* env state */
    float* actions;            /* (N,12) */
    float* last_actions;       /* (N,12) */
    float* last_last_actions;  /* (N,12) */
    float* torques;            /* (N,12) */
    float* last_dof_vel;       /* (N,12) */
    float* last_root_vel;      /* (N,6)  */
    float* commands;           /* (N,4)  */
    int64_t* episode_length_buf;   /* (N) */
    uint8_t* reset_buf;        /* (N) bool */
    uint8_t* time_out_buf;     /* (N) bool */
    uint8_t* extras_time_outs; /* (N) bool, refreshed only on steps with >=1 reset */
    float* base_lin_vel;       /* (N,3) */
    float* base_ang_vel;       /* (N,3) */
    float* projected_gravity;  /* (N,3) */
    float* base_euler_xyz;     /* (N,3) */
    float* feet_air_time;      /* (N,2) */
    uint8_t* last_contacts;    /* (N,2) bool */
    float* feet_height;        /* (N,2) */
    float* last_feet_z;        /* (N,2) */
    float* ref_dof_pos;        /* (N,12) */
    float* rand_push_force;    /* (N,3) */
    float* rand_push_torque;   /* (N,3) */
    const float* env_frictions;    /* (N,1) */
    const float* body_mass;        /* (N,1) */
    const float* env_origins;      /* (N,3) */
    float* episode_sums;       /* (22,N) one row per reward term, alphabetical     */
    float* episode_means;      /* (22)   extras["episode"]["rew_*"]                */
    float* rew_terms;          /* (22,N) optional (may be NULL): per-term scaled reward of this step */
    float* obs_buf;            /* (N,705) 15 frames oldest->newest, clipped: history INPUT  */
    float* privileged_obs_buf; /* (N,219) 3 frames: history INPUT                           */
    float* obs_out;            /* (N,705) shifted history + new frame is written here; must
                                  not alias obs_buf (ping-pong pair, or a rollout-storage slab) */
    float* priv_out;           /* (N,219) likewise                                          */
    float* rew_buf;            /* (N)                                              */
    int32_t* reset_ids;        /* (N) compacted ids of envs reset this step (any order) */
    int32_t* scratch;          /* (32) int32, zero-initialised once by the caller:
                                  [0] running reset count, [1] CTA ticket, [3] reset
                                  count of the last step, [4..5] int64 common_step_counter,
                                  [6..7] uint64 noise step, [8..29] float accumulators  */
    int64_t obs_pitch;         /* row pitch (elements) of obs_buf AND obs_out; 0 = dense 705.  A multiple of 4
                                  (708) makes the rows TMA-addressable for the tensor-core actor forward      */
    int64_t priv_pitch;        /* likewise for privileged_obs_buf / priv_out; 0 = dense 219 (220 for TMA)     */
} HgEnvBuffers;

/* Injected random draws (parity mode).  Any pointer may be NULL: the kernel
 * then draws the numbers itself from Philox4x32-10 keyed by (seed, step, env). */
typedef struct HgEnvNoise {
    const float* u_cmd_cb;     /* (N,3) U[0,1): command resample in the step callback */
    const float* u_cmd_rs;     /* (N,3) U[0,1): command resample on reset             */
    const float* u_dof;        /* (N,12) U[0,1): joint positions on reset             */
    const float* u_push;       /* (N,5) U[0,1): push lin-vel xy, ang-vel xyz          */
    const float* z_obs;        /* (N,47) N(0,1): observation noise                    */
    uint64_t seed;
    uint64_t step;             /* distinct per call                                   */
    int32_t use_device_counters;   /* 1: `step` and `common_step_counter` are taken from
                                      scratch[6..7] / scratch[4..5] (int64 each) and bumped
                                      by the kernel -> the launch is CUDA-graph replayable */
    int32_t _pad;
} HgEnvNoise;

/* phases of hg_env_post_physics (bit mask) */
#define HG_PHASE_COUNTERS  0x01u  /* episode_length++ ; base-frame quantities (legged_robot.py:128-136) */
#define HG_PHASE_CALLBACK  0x02u  /* command resample, heading command, push  (:304-320, humanoid_env.py:83-98) */
#define HG_PHASE_TERMINATE 0x04u  /* check_termination (:156-161)                                         */
#define HG_PHASE_REWARD    0x08u  /* compute_reward + 22 terms (:217-235, humanoid_env.py:272-540)        */
#define HG_PHASE_RESET     0x10u  /* reset_idx for envs with reset_buf set (:163-215, humanoid_env.py:264-269) */
#define HG_PHASE_OBS       0x20u  /* compute_observations (humanoid_env.py:200-262)                       */
#define HG_PHASE_LAST      0x40u  /* last_* copies (:147-151) and the +-18 clip (:104-108)                */
#define HG_PHASE_STEP_ALL  0x7Fu

/* XBotLFreeEnv.step prologue (humanoid_env.py:189-197) + LeggedRobot.step clip
 * (legged_robot.py:90-91): clip, delay-mix with the previous actions,
 * multiplicative noise, clip.  Writes B->actions.
 * u_delay (N,1) U[0,1) and z_act (N,12) N(0,1) may be NULL (Philox).
 * step == UINT64_MAX: take the noise step from scratch[6..7] (device counter). */
int32_t hg_env_pre_physics(const HgEnvBuffers* B, const HgEnvParams* P, const float* actions_in,
                           const float* u_delay, const float* z_act, uint64_t seed, uint64_t step,
                           int64_t N, void* stream);

/* LeggedRobot._compute_torques (legged_robot.py:340-356); called once per
 * decimation sub-step.  Reads B->actions, B->dof_state; writes B->torques. */
int32_t hg_env_compute_torques(const HgEnvBuffers* B, const HgEnvParams* P, int64_t N, void* stream);

/* debug aid: the following hg_env_post_physics launches stamp %globaltimer at their phase boundaries, buf = [grid][12] int64 */
void hg_env_set_trace(long long* buf);

/* Synthetic-physics fast path of the decimation loop (legged_robot.py:94-101) + the three state refreshes of
 * post_physics_step (:124-126): with an open-loop frame source the `decimation` x {PD torque, set forces, simulate,
 * refresh dof} sub-steps collapse into ONE launch.  dof_frames: (decimation, N*12, 2) the dof state after each
 * sub-step; root / contact / rigid frame: the state after the last one.  Sub-step d's PD law is evaluated against
 * the state sub-step d-1 left (d = 0: the live B->dof_state); B->torques receives the last torque, B->dof_state the
 * last frame -- exactly what the loop leaves behind.  A real simulator keeps the loop (hg_env_compute_torques). */
int32_t hg_env_synth_decimation(const HgEnvBuffers* B, const HgEnvParams* P, const float* dof_frames, int32_t decimation,
                                const float* root_frame, const float* contact_frame, const float* rigid_frame,
                                int64_t N, void* stream);

/* LeggedRobot.post_physics_step (legged_robot.py:119-154) with everything it
 * calls, fused into one launch; `phases` selects sub-sequences so that
 * reset_idx() / compute_observations() stay callable on their own.
 * `common_step_counter` is the value AFTER this step's increment. */
int32_t hg_env_post_physics(const HgEnvBuffers* B, const HgEnvParams* P, const HgEnvNoise* Z,
                            uint32_t phases, int64_t common_step_counter, int64_t N, void* stream);

/* ---- rough terrain: mesh_type 'heightfield' / 'trimesh' -------------------------------
 * The height field is the int16 grid HumanoidTerrain builds on the host (utils/terrain.py:38-63); everything
 * the env does with it per step runs here.  In this mode LeggedRobot splits hg_env_post_physics into
 * {COUNTERS|CALLBACK|TERMINATE|REWARD} and {RESET|OBS|LAST} with hg_terrain_reset_prepare in between, because
 * the curriculum moves env_origins of the envs that terminated BEFORE their root state is re-initialised
 * (legged_robot.py:175-186). */
typedef struct HgTerrain {
    const int16_t* height_samples;  /* (rows, cols) row-major = Terrain.heightsamples (legged_robot.py:570,586)   */
    int32_t rows, cols;             /* tot_rows, tot_cols                                                         */
    float border_size;              /* [m]                                                                        */
    float horizontal_scale;         /* [m per cell]                                                               */
    float vertical_scale;           /* [m per int16 unit]                                                         */
    const float* terrain_origins;   /* (num_levels, num_types, 3) = Terrain.env_origins as fp32 (:696)            */
    int32_t num_levels, num_types;  /* cfg.terrain.num_rows (= max_terrain_level), num_cols                       */
    float half_env_length;          /* terrain.env_length / 2: walking further moves an env up one level (:412)   */
    float max_episode_length_s;
    int32_t curriculum;             /* cfg.terrain.curriculum                                                      */
    int32_t _pad;
} HgTerrain;

/* LeggedRobot._get_heights (legged_robot.py:759-795): for every env the P base-frame grid points
 * `points_xy` (P,2) are rotated by the base yaw (quat_apply_yaw, utils/math.py:38-43), moved to the root
 * position, shifted by the border, divided by the horizontal scale and truncated (.long()); the height is the
 * minimum of the three samples (px,py), (px+1,py), (px,py+1) with px / py clipped to [0, rows-2] / [0, cols-2],
 * times the vertical scale.  heights: (N,P). */
int32_t hg_terrain_get_heights(const HgTerrain* T, const float* root_states, const float* points_xy, int32_t P,
                               float* heights, int64_t N, void* stream);

/* For the envs with reset_buf set: LeggedRobot._update_terrain_curriculum (legged_robot.py:400-420, skipped when
 * T->curriculum == 0) on terrain_levels / env_origins, then the spawn position _reset_root_states adds to
 * base_init_state under custom origins (:381-384): spawn = env_origins + (U(-1,1), U(-1,1), 0).  `spawn` (N,3)
 * is what the following hg_env_post_physics(RESET|...) launch must see as B->env_origins.
 * r_level (N) int64: the torch.randint_like draw of an env that solved the last level; u_root (N,2) U[0,1).
 * Either may be NULL: Philox4x32-10 keyed by (seed, step, env). */
int32_t hg_terrain_reset_prepare(const HgTerrain* T, const uint8_t* reset_buf, const float* root_states,
                                 const float* commands, int64_t* terrain_levels, const int64_t* terrain_types,
                                 float* env_origins, float* spawn, const int64_t* r_level, const float* u_root,
                                 uint64_t seed, uint64_t step, const uint64_t* step_dev, int64_t N, void* stream);

/* XBotLFreeEnv.compute_observations with measure_heights (humanoid_env.py:246-248,253-258,264-269): the critic
 * frame becomes [obs_buf of the PREVIOUS step (num_obs wide, already clipped) | clip(root_z - 0.5 - heights, -1, 1)
 * * height_scale], W = num_obs + P wide; priv_out = the `frames`-deep history (oldest first) with that frame
 * appended, rows of reset envs zeroed before the append (reset_buf may be NULL: a stand-alone compute_observations),
 * everything clipped to +-clip_obs.  Pitches in elements. */
int32_t hg_terrain_priv_frames(const float* obs_prev, int64_t obs_pitch, int32_t num_obs, const float* root_states,
                               const float* heights, int32_t P, float height_scale, float clip_obs,
                               const uint8_t* reset_buf, const float* priv_in, float* priv_out, int64_t priv_pitch,
                               int32_t frames, int64_t N, void* stream);

/* ------------------------------------------------------------------------ */
/* Learning side                                                            */
/* ------------------------------------------------------------------------ */

/* One ELU-MLP (nn.Sequential(Linear, ELU, ..., Linear), algo/ppo/actor_critic.py:54-77).
 * Parameters live in ONE flat fp32 buffer shared by actor, critic and std so
 * that gradient all-reduce / Adam touch a single range. */
typedef struct HgMlpDesc {
    int32_t n_layers;                   /* number of Linear layers                   */
    int32_t dims[HG_MAX_LAYERS + 1];    /* in, hidden..., out                        */
    int64_t w_off[HG_MAX_LAYERS];       /* element offset of weight l (out,in) row-major in the flat buffer */
    int64_t b_off[HG_MAX_LAYERS];       /* element offset of bias l                  */
    int64_t ldw[HG_MAX_LAYERS];         /* row pitch of weight l in elements (>= dims[l]); a multiple of 4 with a
                                           16-byte aligned w_off lets TMA address it -> tensor-core path         */
} HgMlpDesc;

/* Y = MLP(X).  `hidden`: caller scratch receiving every hidden layer's post-ELU
 * output, (M, sum(dims[1..n-1])) laid out layer after layer [layer l at element
 * offset M*sum(dims[1..l-1])] (kept for the backward pass); `out` (M, dims[n])
 * receives the network output and may point straight into a rollout-storage slab.
 * Replaces nn.Sequential.forward as used by ActorCritic.act / evaluate
 * (actor_critic.py:111-128). */
int32_t hg_mlp_forward(const HgMlpDesc* net, const float* params, const float* X, int64_t ldx,
                       float* hidden, float* out, int64_t M, void* stream);

/* Backward of the same MLP.  dY (M, dims[n]) is the loss gradient w.r.t. the
 * network output; grads (same flat layout as params) receives dW, db
 * (overwritten, not accumulated).  dhidden: scratch of the same size as hidden.
 * Replaces autograd through nn.Sequential in PPO.update (ppo.py:170-173). */
int32_t hg_mlp_backward(const HgMlpDesc* net, const float* params, const float* X, int64_t ldx,
                        const float* hidden, const float* dY, float* dhidden, float* grads,
                        int64_t M, void* stream);

/* Tensor-core GEMM building block of hg_mlp_forward / hg_mlp_backward (tcgen05.mma kind::tf32, accumulator
 * in TMEM, operands by TMA):  C (M x N, pitch ldc) = A x B reduced over K, fp32 in / fp32 out.
 *   a_mn_major = 0: A is (M, K) row-major  (pitch lda);  1: A is (K, M) row-major, i.e. the M index is contiguous
 *   b_mn_major = 0: B is (N, K) row-major  (pitch ldb);  1: B is (K, N) row-major
 *   passes     = 3: 3xTF32 split compensation (fp32-class accuracy);  1: plain TF32
 *   epilogue   : 0 store, 1 +bias[N], 2 +bias then ELU, 3 multiply by ELU'(z) recovered from H = ELU(z) (pitch ldh),
 *                4 atomicAdd into C (required when split_k > 1; C must be zeroed by the caller),
 *                5 +bias then the fused PPO.act sampling epilogue (sample_* fields)
 *   trust_hw_truncation: 1 = feed the raw fp32 tile as the "hi" operand (the tensor core drops the low 13
 *                mantissa bits itself); 0 = rewrite it with an explicit truncation first
 * Requirements: A, B 16-byte aligned with lda, ldb multiples of 4 (TMA); violations return HG_E_ALIGN. */
typedef struct HgGemm {
    const float* A; const float* B; float* C; const float* bias; const float* H;
    int32_t M, N, K;
    int64_t lda, ldb, ldc, ldh;
    int32_t a_mn_major, b_mn_major, epilogue, passes, split_k, trust_hw_truncation;
    const float* B_lo;              /* optional, K-major B with passes = 3: rna_tf32(B - trunc_tf32(B)) in the same layout
                                       (hg_tf32_residual): the tile arrives by TMA and the kernel splits A only         */
    /* epilogue 5 (N <= 32): + bias, then PPO.act on the row -- C receives the mean, see hg_policy_sample */
    const float* sample_std; const float* sample_eps; float* sample_actions; float* sample_log_prob; float* sample_sigma;
    uint64_t sample_seed, sample_step; const uint64_t* sample_step_dev;
} HgGemm;
int32_t hg_gemm_tf32(const HgGemm* d, void* stream);
/* dst[i] = rna_tf32(src[i] - trunc_tf32(src[i])): the "lo" operand of the 3xTF32 scheme for a whole buffer (weights). */
int32_t hg_tf32_residual(const float* src, float* dst, int64_t n, void* stream);

/* hg_mlp_forward with options for the rollout (PPO.act, ppo.py:91-101): params_lo = hg_tf32_residual(params) lets
 * every layer load its weight residuals by TMA (the in-kernel splitter then handles the activations only); when
 * `actions` is set the output layer's epilogue also samples a = mu + sigma z, log-prob and sigma (out receives mu). */
typedef struct HgMlpFwdOpts {
    const float* params_lo;
    const float* std; const float* eps; float* actions; float* log_prob; float* sigma;
    uint64_t seed, step; const uint64_t* step_dev;
} HgMlpFwdOpts;
int32_t hg_mlp_forward_ex(const HgMlpDesc* net, const float* params, const float* X, int64_t ldx, float* hidden, float* out,
                          int64_t M, const HgMlpFwdOpts* opts, void* stream);

/* PPO.act as ONE launch (ppo.py:91-101): the actor and critic MLPs of hg_mlp_forward_ex, all layers of both, in a
 * single persistent tcgen05 3xTF32 kernel whose layer-to-layer dependencies are resolved on the device (tile counters
 * in `counters`, hg_actor_critic_counters_size(M) int32, zero-initialised once by the caller; the kernel leaves them
 * zero).  Same arithmetic, tile shapes and results as two hg_mlp_forward_ex calls; ~3x less wall time at rollout
 * batch sizes, where each layer is a few microseconds of tensor work behind ~8 us of per-launch fixed cost and the
 * shared-memory footprint forbids two GEMM kernels per SM.  hidden_a / hidden_c as `hidden` of hg_mlp_forward;
 * hidden_lo_a / hidden_lo_c: scratch of the same size receiving the tf32 residuals of the hidden activations (the
 * next layer's A_lo tiles then arrive by TMA: no in-kernel splitting except for the network inputs).
 * sample (may be NULL / actions == NULL): the actor's output epilogue samples as in hg_policy_sample.
 * Either net may be NULL (its arguments are then ignored): the runner launches the actor alone on the critical path and the
 * critic alone on a side stream, where it overlaps the env step.
 * Returns HG_E_ALIGN when an operand is not TMA-addressable or the nets have more than 8 layers in total. */
int64_t hg_actor_critic_counters_size(int64_t M);
/* debug aid: per-item %globaltimer stamps of the following launches go to buf ([148][16][16] int64, device); NULL = off */
void hg_actor_critic_set_trace(long long* buf);
/* profiling only: [grid][8] int64 cycle sums per CTA of the next hg_gemm_bf16x3 launches (NULL = off), tools/bf3_trace.py:
 * 0 MMA-thread loop cycles, 1 of which waiting for operands (full), 2 waiting for a drained accumulator, 3 TMA thread waiting
 * for a free stage, 4 epilogue warp 0 waiting for an accumulator, 5 epilogue warp 0 busy, 6 items, 7 k-blocks */
void hg_gemm_bf16x3_set_trace(long long* buf);
int32_t hg_actor_critic_forward(const HgMlpDesc* actor, const HgMlpDesc* critic, const float* params, const float* params_lo,
                                const float* obs, int64_t ld_obs, const float* cobs, int64_t ld_cobs, float* hidden_a,
                                float* hidden_c, float* hidden_lo_a, float* hidden_lo_c, float* mu, float* value,
                                const HgMlpFwdOpts* sample, int32_t* counters, int64_t M, void* stream);

/* GEMM engine of hg_mlp_forward / hg_mlp_backward: 0 = exact-fp32 CUDA-core path, 1 = tcgen05 3xTF32,
 * 2 = tcgen05 plain TF32, 4 (default) = 3xTF32 for this fp32 API (rollout forward, 1e-5 bar) AND a hint to the host
 * layer to run PPO.update on the split-precision API below (hg_mlp_*_split, 1e-4 gradient bar).
 * Layers whose operands TMA cannot address fall back to 0.  Returns the previous mode. */
int32_t hg_set_gemm_mode(int32_t mode);

/* ---- split-precision ("bf16x3") learning path -------------------------------------------------------------
 * A split tensor stores x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 significant bits) as two bf16
 * planes: element (r, c) of plane k (0 = hi, 1 = lo) is p[k * plane + r * ld + c].  ld and plane are multiples
 * of 8 elements and p is 16-byte aligned (TMA).  Every producer on the update path (minibatch gather, GEMM
 * epilogues, output-head backward, the post-Adam weight split) writes this format directly, so the tensor-core
 * GEMM needs no conversion pass: D += A_lo B_hi + A_hi B_lo + A_hi B_hi as three tcgen05 kind::f16 MMAs with
 * fp32 accumulation (relative error ~5e-6 per product; gradients ~1e-5 against the 1e-4 bar). */
typedef struct HgSplit { uint16_t* p; int64_t ld; int64_t plane; } HgSplit;
int32_t hg_split_bf16(const float* src, int64_t ld_src, const HgSplit* dst, int64_t rows, int64_t cols, void* stream);
int32_t hg_unsplit_bf16(const HgSplit* src, float* dst, int64_t ld_dst, int64_t rows, int64_t cols, void* stream);

/* C (M x N) = op(A) op(B) reduced over K on split operands (tcgen05.mma kind::f16, TMEM accumulators, TMA).
 *   a_mn_major = 0: A is (M, K) row-major;  1: A is (K, M) row-major.   b_mn_major likewise with N.
 *   epilogue 0: C fp32 store            1: C fp32 = acc + bias[N]
 *            2: Cs split = ELU(acc + bias)                         (forward of a hidden Linear layer)
 *            3: Cs split = acc * ELU'(h), h = Hs ~= ELU(z); colsum[N] += column sums of the result (bias gradient; may be NULL)
 *            4: atomicAdd into C fp32 (required for split_k > 1; caller zeroes C)      (weight gradient)
 *            5: Cs split = acc */
typedef struct HgGemmSplit {
    HgSplit A, B;
    float* C; int64_t ldc;
    HgSplit Cs;
    const float* bias;
    HgSplit Hs;
    float* colsum;
    int32_t M, N, K;
    int32_t a_mn_major, b_mn_major, epilogue, split_k;
} HgGemmSplit;
int32_t hg_gemm_bf16x3(const HgGemmSplit* d, void* stream);

/* The MLP of hg_mlp_forward / hg_mlp_backward on split tensors (PPO.update path).  `wsplit` is the split image of
 * the flat parameter buffer (same element offsets: w_off / ldw of HgMlpDesc, ldw % 8 == 0 required), `w_plane` its
 * plane distance.  hidden / dhidden: split scratch, hidden layer l (1 <= l < n_layers) at element offset
 * 2 * M * sum(dims[1..l-1]) as planes (M, dims[l]), ld = dims[l], plane = M * dims[l]; every hidden width must be a
 * multiple of 8.  The output layer (<= 16 wide) runs on CUDA cores straight from the last hidden split tensor; in
 * the backward pass one kernel produces its weight / bias gradients, the previous layer's dZ (split) and that
 * layer's bias gradient in a single pass over the activations.  Bias gradients of the hidden layers come out of
 * the dgrad epilogues (no separate column-sum pass).  Returns HG_E_ALIGN when the net is not eligible. */
int32_t hg_mlp_forward_split(const HgMlpDesc* net, const float* params, const uint16_t* wsplit, int64_t w_plane,
                             const HgSplit* X, uint16_t* hidden, float* out, int64_t M, void* stream);
int32_t hg_mlp_backward_split(const HgMlpDesc* net, const float* params, const uint16_t* wsplit, int64_t w_plane,
                              const HgSplit* X, const uint16_t* hidden, const float* dY, uint16_t* dhidden,
                              float* grads, int64_t M, void* stream);

/* PPO.act epilogue (ppo.py:91-101, actor_critic.py:111-120): actions =
 * mean + std*eps, log-prob summed over actions, sigma broadcast.
 * eps (M,A) may be NULL (Philox with seed/step).  step_dev, when not NULL, is a device counter read instead
 * of `step` (the env's noise-step counter: the launch then carries no per-step host value -> graph replayable). */
int32_t hg_policy_sample(const float* mean, const float* std, const float* eps, uint64_t seed, uint64_t step,
                         const uint64_t* step_dev, float* actions, float* log_prob, float* sigma_out, int64_t M,
                         int32_t A, void* stream);

/* RolloutStorage.add_transitions (rollout_storage.py:87-100) fused with the
 * time-out bootstrap of PPO.process_env_step (ppo.py:107-108): copies one
 * step's tensors into slab `t` of the (T,N,.) storage in a single launch.
 * Any HgTransition pointer may be NULL or already equal to its slab (the
 * producer wrote in place): that tensor is skipped. */
typedef struct HgTransition {
    const float* obs; const float* priv_obs; const float* actions; const float* rewards;
    const uint8_t* dones; const uint8_t* time_outs; const float* values; const float* log_prob;
    const float* mu; const float* sigma;
    int64_t obs_pitch, priv_pitch;      /* row pitch of obs / priv_obs in elements (0 = dense)            */
} HgTransition;
typedef struct HgStorage {
    float* observations; float* privileged_observations; float* actions; float* rewards;
    uint8_t* dones; float* values; float* actions_log_prob; float* mu; float* sigma;
    float* returns; float* advantages;
    int32_t T, num_obs, num_priv, num_actions;
} HgStorage;
int32_t hg_storage_add(const HgStorage* S, const HgTransition* tr, int32_t t, float gamma, int64_t N, void* stream);

/* RolloutStorage.compute_returns (rollout_storage.py:122-136): reverse GAE
 * scan over T for each env + global advantage normalisation (unbiased std).
 * stats: (4) fp32+fp64-free scratch [sum, sumsq, count, pad] -- if
 * `normalise`==0 only the raw sums are produced (multi-GPU: all-reduce them,
 * then call hg_adv_normalise). */
int32_t hg_gae(const HgStorage* S, const float* last_values, float gamma, float lam, double* stats,
               int32_t normalise, int64_t N, void* stream);
/* The reverse scan runs as a WARP SCAN OVER TIME (default): the recurrence adv_t = d_t + c_t adv_{t+1} is a chain of
 * affine maps, composed associatively with shuffles (one warp per env, 32 envs per CTA staged through shared memory,
 * each lane replays its own steps with the reference's serial formula from the scanned carry-in).
 * hg_set_gae_mode(0) selects the one-thread-per-env serial walk, 1 the warp scan, 2 picks by N (scan up to 32768 envs: a
 * thread-per-env walk leaves SMs idle behind T dependent steps there), -1 re-reads HG_GAE=scan|serial|auto (default
 * auto) from the environment; returns the previous setting. */
int32_t hg_set_gae_mode(int32_t scan);
int32_t hg_adv_normalise(const HgStorage* S, const double* stats, int64_t N, void* stream);

/* mini_batch_generator gather (rollout_storage.py:146-182): rows idx[0..B) of
 * the flattened (T*N,.) storage into contiguous minibatch tensors, one launch. */
typedef struct HgMiniBatch {
    float* obs; float* priv_obs; float* actions; float* values; float* advantages; float* returns;
    float* old_log_prob; float* old_mu; float* old_sigma;
    int64_t ld_obs, ld_priv;            /* row pitch of obs / priv_obs in elements (0 = dense); a multiple of 4
                                           makes the rows TMA-addressable for the tensor-core MLP path        */
    HgSplit obs_split, priv_split;      /* optional (p may be NULL): the observations are ALSO / INSTEAD (obs == NULL)
                                           written as split bf16 planes for the bf16x3 update path              */
} HgMiniBatch;
int32_t hg_minibatch_gather(const HgStorage* S, const int64_t* idx, const HgMiniBatch* mb, int64_t B, void* stream);

/* PPO.update loss (ppo.py:133-168) forward + analytic backward in one kernel:
 * log-prob, ratio, clipped surrogate, clipped value loss, entropy, KL.
 * Outputs d loss/d mean (B,A), d loss/d value (B,1), d loss/d std (A) and
 * `scalars` (8 fp32, zeroed by the call): [0] surrogate loss, [1] value loss,
 * [2] entropy, [3] KL mean.  Means use `inv_B` = 1/B_global so that partial
 * results of env-sharded ranks SUM to the single-process value. */
typedef struct HgPpoLossArgs {
    const float* mean; const float* value; const float* std;
    const float* actions; const float* target_values; const float* advantages; const float* returns;
    const float* old_log_prob; const float* old_mu; const float* old_sigma;
    float* d_mean; float* d_value; float* grad_std; float* scalars;
    float clip_param, value_loss_coef, entropy_coef;
    int32_t use_clipped_value_loss;
    int32_t num_actions;
    float inv_B;
} HgPpoLossArgs;
int32_t hg_ppo_loss_fwd_bwd(const HgPpoLossArgs* a, int64_t B, void* stream);

/* clip_grad_norm_ + Adam (ppo.py:172-173; torch.optim.Adam defaults) over the
 * flat parameter buffer: a squared-norm reduction, then the fused clip+Adam
 * update.  The learning rate (double) and the Adam step count live in DEVICE
 * memory so that the adaptive-KL schedule needs no host sync; the update
 * kernel increments step_dev[0] and re-zeroes sqnorm[0] for the next step.
 * Effective gradient = grads * grad_scale (1.0 unless the caller all-reduced
 * un-normalised partial gradients). */
int32_t hg_grad_sqnorm(const float* grads, int64_t n, double* sqnorm_out, void* stream);
int32_t hg_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                          double* sqnorm, float max_grad_norm, const double* lr_dev, int32_t* step_dev,
                          float beta1, float beta2, float eps, float grad_scale, int64_t n, void* stream);
/* Same, and the one-thread tail kernel also adds this minibatch's n_stats loss statistics `stats` to the running sums
 * `stats_sum` that PPO.update averages at its end (ppo.py:175-184): no separate accumulation op per optimizer step. */
int32_t hg_clip_adam_step_stats(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                double* sqnorm, float max_grad_norm, const double* lr_dev, int32_t* step_dev,
                                float beta1, float beta2, float eps, float grad_scale, int64_t n,
                                const float* stats, float* stats_sum, int32_t n_stats, void* stream);

/* Minibatch permutation of PPO.update (rollout_storage.py:155 `torch.randperm`): out[0..n) = a pseudo-random permutation
 * of 0..n-1 determined by (seed, counter), in ONE launch and without a sort (keyed Feistel bijection of the enclosing
 * power-of-two range + cycle walking).  It is a different generator than torch's, so the minibatch COMPOSITION differs
 * from a torch run with the same seed -- as it does between any two torch seeds; every sample is still used exactly once
 * per epoch. */
int32_t hg_randperm(int64_t n, uint64_t seed, uint64_t counter, int64_t* out, void* stream);

/* OnPolicyRunner.learn's per-step episode bookkeeping (on_policy_runner.py:140-154), one launch, no host sync:
 * cur_reward_sum += rewards; cur_episode_length += 1; for finished envs (dones != 0) the totals are written to
 * done_rew_t / done_len_t (row t of (T, N) slabs; NaN where the env did not finish) and the running values restart at 0;
 * the n_infos extras["episode"] scalars of this step (infos_in, may be NULL) are copied to infos_out_t (row t of (T, n_infos)).
 * The runner reads the slabs back once per iteration instead of `.cpu()`-ing every step. */
int32_t hg_episode_book_step(const float* rewards, const uint8_t* dones, float* cur_reward_sum, float* cur_episode_length,
                             float* done_rew_t, float* done_len_t, const float* infos_in, float* infos_out_t,
                             int32_t n_infos, int64_t N, void* stream);

/* Adaptive-KL learning-rate rule (ppo.py:142-148) evaluated on the device:
 * reads kl_mean_dev[0] (fp32 KL mean of the minibatch), updates lr_dev[0]. */
int32_t hg_adapt_lr(const float* kl_mean_dev, double desired_kl, double* lr_dev, void* stream);

/* ------------------------------------------------------------------------ */
/* misc                                                                     */
/* ------------------------------------------------------------------------ */
int32_t hg_version(void);
/* fp16x3 form of hg_actor_critic_forward (the rollout default): same persistent multi-layer kernel and dependency scheme,
 * operands as two fp16 planes x ~= hi + lo (22 significant bits, like the hi / lo pair of 3xTF32) multiplied as three
 * tcgen05 kind::f16 MMAs -- half the operand bytes per k and twice the MMA rate of 3xTF32 on fp32 tiles, same ~1e-7
 * network-level error.  fp16's narrow exponent is handled by storing the WEIGHTS scaled by hg_f16_weight_scale() (2^10;
 * the epilogue scales back exactly); activations must stay below 65504 in magnitude (they saturate there).
 *   w16 / w16_plane : hg_split_f16(params, scale = hg_f16_weight_scale()) image of the whole flat parameter buffer
 *                     (plane k at w16 + k * w16_plane), refreshed by the caller after every optimizer step;
 *   scratch16       : hg_actor_critic_f16_scratch_elems(actor, critic, M) uint16 of staging (input + hidden planes);
 *   obs / cobs      : fp32 network inputs (split into planes by a first small launch of this call).
 * Everything else as hg_actor_critic_forward (counters, sample, either net may be NULL). */
float hg_f16_weight_scale(void);
int32_t hg_split_f16(const float* src, int64_t ld_src, const HgSplit* dst, int64_t rows, int64_t cols, float scale, void* stream);
int64_t hg_actor_critic_f16_scratch_elems(const HgMlpDesc* actor, const HgMlpDesc* critic, int64_t M);
int32_t hg_actor_critic_forward_f16(const HgMlpDesc* actor, const HgMlpDesc* critic, const float* params, const uint16_t* w16,
                                    int64_t w16_plane, const float* obs, int64_t ld_obs, const float* cobs, int64_t ld_cobs,
                                    uint16_t* scratch16, float* mu, float* value, const HgMlpFwdOpts* sample, int32_t* counters,
                                    int64_t M, void* stream);

/* sizeof() of the ABI structs, for binding self-checks: 0 HgEnvParams, 1 HgEnvBuffers,
 * 2 HgEnvNoise, 3 HgMlpDesc, 4 HgTransition, 5 HgStorage, 6 HgMiniBatch, 7 HgPpoLossArgs, 8 HgGemm, 9 HgSplit,
 * 10 HgGemmSplit, 11 HgMlpFwdOpts */
int64_t hg_struct_size(int32_t which);
const char* hg_last_error(void);
/* number of kernel launches issued by this library in the calling process */
int64_t hg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HG_B200_H */
