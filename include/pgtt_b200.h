/* pgtt_b200.h - C ABI of the B200-native GO2 PGTT environment (libpgtt_b200.so).
 *
 * Drop-in boundary for the reference's batched env hot path.  The reference has no FFI of its own
 * (it is Python over JAX); the entry points below are what a binding for that path needs, one per
 * reference call:
 *
 *   pgtt_create            Go2Env.__init__ / Joystick.__init__      go2/base.py:45-113, go2/joystick_pgtt.py:38-48
 *   pgtt_set_terrain_table jnp.load(terrain_file) + partial(domain_randomize, terrain_matrix=...)
 *                                                                    training/train.py:165-170
 *   pgtt_randomize         domain_randomize(model, rng, terrain)     go2/randomize.py:23-171, randomize_simple.py:24-138
 *   pgtt_reset             Joystick.reset (+ wrapper resets)         go2/joystick_pgtt.py:50-131
 *   pgtt_step              wrapped env.step                          go2/joystick_pgtt.py:141-231, training/train.py:255
 *   pgtt_heightscan        create_sensor_matrix                      go2/heightmap.py:25-67
 *   pgtt_forward           mjx.forward on the current state          go2/joystick_pgtt.py:78
 *   pgtt_policy_act /      brax ppo acting step + generate_unroll    training/train.py:135-161,242-263
 *   pgtt_rollout
 *   pgtt_get_buffers       State / info / data field access          go2/joystick_pgtt.py:101-131
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns 0 on success or
 * a negative pgtt_status; pgtt_last_error() gives the message of the last failure on the calling
 * thread.  A handle is bound to one device and is NOT thread-safe.  All launches are asynchronous on
 * the caller's stream (a cudaStream_t passed as void*; NULL = default stream); no entry point
 * synchronises the host except pgtt_create / pgtt_destroy / pgtt_set_terrain_table / pgtt_sync.
 * Per-env arrays are row-major [num_envs][dim] in DEVICE memory owned by the handle.
 */
#ifndef PGTT_B200_H
#define PGTT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGTT_NQ 19
#define PGTT_NV 18
#define PGTT_NU 12
#define PGTT_NBOX 100
#define PGTT_NRAY 117          /* 13 x 9 */
#define PGTT_NOBS 171          /* phase-guided task; the baseline variant has 162 / 206 (pgtt_obs_dims) */
#define PGTT_NPRIV 215
#define PGTT_NMETRIC 22        /* 21 reward terms (order of go2/configs.py:31-59) + swing_peak */
#define PGTT_NSENSOR 49
#define PGTT_NCON 8            /* 4 foot-plane + <=4 foot-box contacts (max_contact_points) */
#define PGTT_NEPMETRIC 24      /* sum_reward, length, 22 metrics */

typedef enum {
  PGTT_OK = 0,
  PGTT_ERR_ARG = -1,           /* bad argument / model outside the supported family */
  PGTT_ERR_CUDA = -2,          /* CUDA runtime error (message has the cudaError string) */
  PGTT_ERR_STATE = -3,         /* call order (e.g. stairs task stepped before a terrain table was set) */
  PGTT_ERR_NOMEM = -4
} pgtt_status;

/* Compiled model constants (float64 on the host side; see model.py:Go2Model). Bodies 1..13 =
 * base, FL_{hip,thigh,calf}, FR_*, RL_*, RR_*; hinges 0..11 in that (qpos) order; actuators in MJCF
 * order FR, FL, RR, RL. */
typedef struct {
  double timestep, gravity[3], impratio, tolerance, ls_tolerance, meaninertia;
  int iterations, ls_iterations, max_geom_pairs, max_contact_points, n_boxes;
  double body_pos[14][3], body_ipos[14][3], body_iquat[14][4], body_mass[14], body_inertia[14][3];
  double body_invweight0[14][2];
  double jnt_range[12][2], jnt_solref[2], jnt_solimp[5];
  double qpos0[19], dof_armature[18], dof_damping[18], dof_invweight0[18];
  int act_dof[12];
  double act_gain[12], act_bias[12][3], act_ctrlrange[12][2], act_forcerange[12][2];
  int foot_geom_id[4];         /* FL FR RL RR */
  double foot_pos[3], foot_radius, foot_friction[3], foot_solref[2], foot_solimp[5], foot_margin;
  int floor_geom_id, box_geom_id0;
  double floor_friction[3], floor_solref[2], floor_solimp[5];
  double box_rbound, box_friction[3], box_solref[2], box_solimp[5];
  double imu_pos[3];
  int n_model_bodies;          /* nbody of the MJCF scene (14 or 114): length of the body-mass DR draw */
} pgtt_model_desc;

/* Task configuration = go2/configs.py:default_config() flattened. */
typedef struct {
  double ctrl_dt, action_scale, noise_level;
  double noise_joint_pos, noise_joint_vel, noise_gyro, noise_gravity, noise_linvel, noise_heightscan;
  double reward_scale[21];     /* order of go2/configs.py:31-59 */
  double tracking_sigma, swing_height, base_feet_distance, phase_sigma;
  double cmd_u_max[3], cmd_u_min[3], cmd_b[3], gait_freq[2];
  double soft_limit_factor;
  double default_pose[12], home_qpos[19];
  int history_update_steps, episode_length, n_substeps;
  int rng_partitionable;       /* jax_threefry_partitionable: 1 = JAX >= 0.5 default */
  int variant;                 /* 0 = go2/joystick_pgtt.py (obs 171 / 215), 1 = go2/joystick.py baseline task (obs 162 / 206) */
} pgtt_task_desc;

/* Device pointers into the handle's state, [num_envs][dim] row-major. float unless noted. */
typedef struct {
  int num_envs;
  /* mjx.Data fields the task reads or exposes */
  float *qpos, *qvel, *qacc, *qacc_warmstart, *ctrl, *time;
  float *sensordata, *actuator_force, *site_xpos /*[5][3] imu FL FR RL RR*/, *site_xmat /*[9] imu*/;
  float *contact_dist /*[8]*/;
  int32_t *contact_geom /*[8][2]*/;
  int32_t *solver_niter /*[n_substeps]*/;
  /* State */
  float *obs_state, *obs_privileged, *reward, *done, *metrics;
  /* info (go2/joystick_pgtt.py:101-120) */
  uint32_t *rng /*[2]*/;
  float *command;
  int32_t *step, *steps_until_next_cmd;
  float *phase, *phase_dt, *gait_freq, *last_act, *last_last_act, *feet_air_time;
  int32_t *last_contact /*[4]*/;
  float *swing_peak, *H_max, *H_min, *heightscan /*[117][3]*/, *motor_targets;
  float *qpos_error_history /*[24]*/, *qvel_history /*[24]*/;
  int32_t *contact /*[4] FR FL RR RL, this step*/, *first_contact /*[4]*/;
  /* wrapper keys (brax EpisodeWrapper / playground BraxAutoResetWrapper) */
  float *steps, *truncation, *episode_done, *episode_metrics /*[24]*/;
  float *first_qpos, *first_qvel, *first_qacc_warmstart, *first_obs_state, *first_obs_privileged;
  /* per-env model written by pgtt_randomize (go2/randomize.py:140-169) */
  float *body_mass /*[13] bodies 1..13*/, *body_ipos_base /*[3]*/, *dof_armature /*[12]*/, *dof_damping /*[12]*/;
  float *actuator_gain /*[12]*/, *actuator_bias1 /*[12]*/, *qpos0 /*[12]*/, *box_friction /*[100]*/, *floor_friction /*[1]*/;
  int32_t *terrain_index;
} pgtt_buffers;

typedef struct pgtt_env pgtt_env;

const char* pgtt_last_error(void);
int pgtt_version(void);

int pgtt_create(const pgtt_model_desc* model, const pgtt_task_desc* task, int device, int num_envs, pgtt_env** out);
int pgtt_destroy(pgtt_env* env);
int pgtt_sync(pgtt_env* env, void* stream);

/* boxes: HOST float32 [n_terrains][100][10] = pos xyz | quat wxyz | half-size xyz (terrains/level*.npy).
 * Only yaw-rotated boxes (quat x = y = 0) are supported - true for every file the generator writes. */
int pgtt_set_terrain_table(pgtt_env* env, const float* boxes, int n_terrains);

/* keys: DEVICE uint32 [num_envs][2] (jax PRNG keys). dynamics = 0 assigns terrain only. */
int pgtt_randomize(pgtt_env* env, const uint32_t* keys, int dynamics, void* stream);
int pgtt_reset(pgtt_env* env, const uint32_t* keys, void* stream);

/* action: DEVICE float [num_envs][12]. wrapped != 0 applies EpisodeWrapper + auto-reset semantics.
 * Two launches: the physics kernel (n_substeps x mjx.step, go2/joystick_pgtt.py:145-148) and the task kernel (contact flags,
 * ray grid, observation, rewards, info, wrappers; go2/joystick_pgtt.py:149-231). */
int pgtt_step(pgtt_env* env, const float* action, int wrapped, void* stream);
/* pgtt_step that also files the step's transition into ONE time-major rollout slot, written by the task kernel itself (no
 * extra launch): obs_state_dst / obs_priv_dst <- next observation, reward_dst <- reward, discount_dst <- 1 - done,
 * truncation_dst <- truncation. DEVICE pointers to [num_envs][dim] rows; any may be NULL (brax generate_unroll,
 * training/train.py:135-161). */
int pgtt_step_record(pgtt_env* env, const float* action, int wrapped, float* obs_state_dst, float* obs_priv_dst, float* reward_dst,
                     float* discount_dst, float* truncation_dst, void* stream);

/* mjx.forward on the current qpos/qvel/ctrl (refreshes sensordata, contacts, qacc, warmstart). */
int pgtt_forward(pgtt_env* env, void* stream);

/* create_sensor_matrix: center DEVICE [num_envs][3], yaw DEVICE [num_envs] -> out DEVICE [num_envs][117][3]. */
int pgtt_heightscan(pgtt_env* env, const float* center, const float* yaw, float* out, void* stream);

int pgtt_get_buffers(pgtt_env* env, pgtt_buffers* out);
/* Row lengths of obs_state / obs_privileged for this handle's task variant (171 / 215 or 162 / 206). */
int pgtt_obs_dims(pgtt_env* env, int* nobs, int* npriv);

/* Debug / parity probe: one mjx.forward with every intermediate written to `out`
 * (DEVICE float [num_envs][PGTT_DEBUG_FLOATS]); layout in csrc/pgtt_debug.h. */
#define PGTT_DEBUG_FLOATS 2048
int pgtt_debug_forward(pgtt_env* env, float* out, void* stream);

/* Counters the bench reports: kernels launched by this handle since creation. */
int64_t pgtt_launch_count(pgtt_env* env);
/* Which physics kernel this handle launches: 0 = warp-per-env (pgtt_env_kernel), 1 = quad-per-env (pgtt_quad_kernel).
 * Chosen at creation from num_envs (measured crossover) or PGTT_KERNEL=warp|quad. */
int pgtt_step_kernel_generation(pgtt_env* env);
/* Kernel launches per pgtt_step: 1 = generation 1 with the task layer fused behind the physics (default; PGTT_FUSE_TASK=0 splits it), 2 = physics kernel + task kernel. */
int pgtt_step_launches(pgtt_env* env);

/* ---- rollout collector: acting step of brax ppo (training/train.py:135-161,242-263; network spec as re-hosted by
 * deploy/policy_net.py:35-64). One fused tcgen05 kernel: normalise obs -> MLP (swish) -> NormalTanh sample. ---- */
typedef struct pgtt_policy pgtt_policy;
const char* pgtt_policy_last_error(void);
/* sizes[n_layers + 1] = {obs_dim, hidden..., 2 * act_dim}, e.g. {171, 512, 256, 128, 24} */
int pgtt_policy_create(int device, const int* sizes, int n_layers, pgtt_policy** out);
int pgtt_policy_destroy(pgtt_policy* p);
/* kernels[l]: HOST float [in][out] row-major (flax `kernel`), biases[l]: HOST float [out];
 * obs_mean / obs_std: HOST float [obs_dim] (brax running statistics) or NULL for identity. */
int pgtt_policy_set_params(pgtt_policy* p, const float* const* kernels, const float* const* biases, const float* obs_mean, const float* obs_std);
/* the same from DEVICE fp32 arrays, packed by a kernel on `stream` (no host copy, no synchronisation): the learner -> collector hand-over */
int pgtt_policy_set_params_device(pgtt_policy* p, const float* const* kernels, const float* const* biases, const float* obs_mean, const float* obs_std, void* stream);
/* obs DEVICE [n][obs_dim]. eps DEVICE [n][act_dim] or NULL (internal counter-based N(0,1) keyed by seed, step).
 * Outputs DEVICE: action [n][act_dim] = tanh(raw); raw_action [n][act_dim], log_prob [n], logits [n][2 act_dim] may be NULL. */
int pgtt_policy_act(pgtt_policy* p, const float* obs, int n, uint64_t seed, uint64_t step, int deterministic, const float* eps,
                    float* action, float* raw_action, float* log_prob, float* logits, void* stream);
int64_t pgtt_policy_launch_count(pgtt_policy* p);

/* Copies the CURRENT observation / reward / done of the handle into a time-major rollout slot (device-to-device copies;
 * used for slot 0 of an unroll - the per-step slots are written by pgtt_step_record): obs_state_dst / obs_priv_dst <- obs,
 * reward_dst <- reward, discount_dst <- 1 - done, truncation_dst <- truncation.
 * All DEVICE pointers to ONE slot ([num_envs][dim]); any may be NULL. */
int pgtt_record(pgtt_env* env, float* obs_state_dst, float* obs_priv_dst, float* reward_dst, float* discount_dst, float* truncation_dst,
                void* stream);

/* brax generate_unroll (training/train.py:135-161,242-263): T x (policy act -> wrapped env step that records its own
 * transition), all launches issued natively on `stream` (3 kernels per control step: policy, physics, task; nothing
 * returns to the host in between).
 * Buffers are DEVICE, time-major: obs_* [T + 1][N][dim] (slot 0 = the observation the unroll starts from, so
 * next_observation[t] = observation[t + 1]); action / raw_action [T][N][12]; log_prob / reward / discount /
 * truncation [T][N]. Internal exploration noise is keyed by (seed, step0 + t). The 3 T launches (+ the slot-0 copies) are captured into a
 * CUDA graph on first use (per env / buffers / T) and replayed on an internal stream ordered after and before the
 * caller's stream (PGTT_ROLLOUT_GRAPH=0 issues them directly). */
typedef struct {
  float *obs_state, *obs_privileged, *action, *raw_action, *log_prob, *reward, *discount, *truncation;
} pgtt_rollout_buffers;
int pgtt_rollout(pgtt_env* env, pgtt_policy* policy, int T, uint64_t seed, uint64_t step0, int deterministic,
                 const pgtt_rollout_buffers* out, void* stream);

/* ---- learner helpers (brax ppo losses as driven by training/train.py:135-161) ---- */
/* Generalised advantage estimation with brax's truncation handling (compute_gae), one thread per trajectory segment.
 * All DEVICE, time-major: truncation / discount (= 1 - done) / reward [T][B], values [T + 1][B] (last row = bootstrap);
 * outputs vs (value targets) and adv [T][B]. termination = (1 - discount) * (1 - truncation). */
int pgtt_gae(const float* truncation, const float* discount, const float* reward, const float* values, int T, int B,
             float lambda, float gamma, float reward_scaling, float* vs, float* adv, void* stream);

/* out [2][cols] (DEVICE double) = column sums and column sums of squares of x [rows][cols] (DEVICE fp32, row stride ld), accumulated in float64 in a fixed order:
 * the batch statistics brax's running_statistics.update folds into the observation normaliser (training/train.py:140). scratch: DEVICE double
 * [pgtt_col_moments_scratch_doubles(cols)]. Two launches on `stream`. */
long long pgtt_col_moments_scratch_doubles(int cols);
int pgtt_col_moments(const float* x, long long rows, int cols, int ld, double* out, double* scratch, void* stream);

/* The multi-rank form of pgtt_gae_moments (the north-star's advantage-normalisation all-reduce): pgtt_gae_sums leaves this rank's (count, sum, sum of squares)
 * of the advantages in sums3 (DEVICE double [3]); the caller all-reduces them (NCCL, SUM) and pgtt_moments_finalize writes moments [2] = (mean, std) of ALL ranks. */
int pgtt_gae_sums(const float* truncation, const float* discount, const float* reward, const float* values, int T, int B, float lambda, float gamma,
                  float reward_scaling, float* vs, float* adv, double* sums3, void* stream);
int pgtt_moments_finalize(const double* sums3, float* moments, void* stream);

/* Everything of minibatch number *mbi that is not an observation, in one launch (all DEVICE): idx [mb] = perm[*mbi][:] (int64 segment ids, also what
 * pgtt_mlp_forward_gather takes), raw [T][mb][A] from raw_all [T][S][A], scal [n_scal][T][mb] from scal_all [n_scal][T][S], eps [T][mb][A] = eps_all[*mbi]
 * (brax `sgd_step`: shuffle, reshape into minibatches, slice; training/train.py:135-161). *mbi is read on the device, so a captured graph serves every minibatch. */
int pgtt_minibatch_gather(const long long* perm, const long long* mbi, int mb, int S, int T, int A, int n_scal, const float* raw_all, const float* scal_all,
                          const float* eps_all, long long* idx, float* raw, float* scal, float* eps, void* stream);

/* pgtt_gae plus what ppo_loss normalises the advantages with: moments [2] (DEVICE) = population mean and standard deviation of
 * adv over the T x B minibatch, from the same launch (one block; B <= 1024 segments). */
int pgtt_gae_moments(const float* truncation, const float* discount, const float* reward, const float* values, int T, int B, float lambda, float gamma,
                     float reward_scaling, float* vs, float* adv, float* moments, void* stream);

/* Fused PPO loss head (brax ppo losses: NormalTanh log-prob of the stored raw action, importance ratio, clipped surrogate,
 * 0.25 MSE value loss, one-sample entropy estimate) for M = T * B transitions with A action dims, forward AND the
 * gradients with respect to the logits [M][2A] and the value predictions [M] in one launch. adv_moments: DEVICE float[2] =
 * (mean, std) used to normalise the advantages (std + 1e-8). sums: DEVICE float[4], zeroed by the call, receives
 * (total loss, policy loss, value loss, entropy), each already divided by M. */
int pgtt_ppo_head(const float* logits, const float* baseline, const float* raw_action, const float* old_log_prob, const float* adv,
                  const float* vs, const float* eps, const float* adv_moments, int M, int A, float clip_eps, float entropy_cost,
                  float min_std, float* grad_logits, float* grad_baseline, float* sums, void* stream);

/* One optimiser step over a flat parameter vector of n floats (all DEVICE): g' = grad * grad_scale, clipped to a global
 * L2 norm of max_norm (coefficient min(1, max_norm / (|g'| + 1e-6)); max_norm <= 0 = no clip), then Adam with bias correction
 * (optax.adam / torch.optim.Adam: p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)) as configured by brax ppo.train
 * (training/train.py:135-161: learning_rate 3e-4, max_grad_norm 1.0). step: DEVICE float counter t, incremented by the call;
 * scratch: DEVICE float[pgtt_adam_scratch_floats()]. Two launches, no host synchronisation, capturable in a CUDA graph. */
int pgtt_adam_clip(float* param, const float* grad, float* m, float* v, float* step, float* scratch, long long n, float lr, float beta1,
                   float beta2, float eps, float max_norm, float grad_scale, void* stream);
int pgtt_adam_scratch_floats(void);

/* Dense layers of the learner's MLPs (policy 171-512-256-128-24, value 215-512-256-128-1; brax ppo networks as configured by
 * training/train.py:135-161), hand-written tcgen05 GEMMs with fp32-level accuracy: every fp32 operand is split into two
 * bf16 parts (hi + lo) and a k-slice issues three bf16 MMAs (hi hi + hi lo + lo hi) with fp32 accumulation in TMEM - ~16
 * mantissa bits per product (the reference runs these at jax_default_matmul_precision=highest, training/train.py:93-94).
 * All pointers DEVICE fp32, row-major. w [K][N] is the flax `kernel` ([in, out]); x rows may be padded (ldx >= K).
 *   forward:         y [M][N] = x [M][:K] w + b;  silu != 0: z receives the pre-activation and y = z * sigmoid(z)
 *   backward_input:  dx [M][:K] (row stride lddx) = dy [M][N] w^T; z_in != NULL ([M][lddx], the pre-activation whose SiLU is this
 *                    layer's input): dx is multiplied by SiLU'(z_in) in the epilogue, i.e. it is the gradient wrt z_in
 *   backward_params: dw [K][N] = x^T dy and db [N] = column sums of dy (may be NULL), from one GEMM launch (the loader threads
 *                    of the dy operand also sum what they load): the M rows are split over CTAs, the split partials in
 *                    `scratch` (pgtt_linear_backward_params_scratch(M, K, N) floats) are summed in a fixed order: deterministic
 *   silu_backward:   dz = dy * silu'(z), elementwise over n values */
const char* pgtt_learner_last_error(void);
int pgtt_linear_forward(const float* x, int ldx, const float* w, const float* b, int M, int K, int N, int silu, float* y, float* z, void* stream);
int pgtt_linear_backward_input(const float* dy, const float* w, int M, int K, int N, float* dx, int lddx, const float* z_in, void* stream);
int pgtt_linear_backward_params_splits(int M);
long long pgtt_linear_backward_params_scratch(int M, int K, int N);
int pgtt_linear_backward_params(const float* x, int ldx, const float* dy, int M, int K, int N, float* dw, float* db, float* scratch, void* stream);
int pgtt_silu_backward(const float* dy, const float* z, float* dz, long long n, void* stream);

/* Whole-MLP forward / backward of the learner on blocked split-bf16 operands (csrc/pgtt_mlp.cu): what brax's ppo.train
 * evaluates per SGD step for the policy and the value network (training/train.py:135-161: hidden sizes (512, 256, 128), SiLU,
 * fp32-grade products as `jax_default_matmul_precision=highest`, train.py:93-94). A handle owns the device workspace for ONE
 * row count (the minibatch size): blocked copies of the input, of every weight matrix and of every activation / gradient,
 * the pre-activations, and the split partials of the weight gradients. dims [n_layers + 1] = in, hidden..., out.
 *   forward:  y [rows][dims[L]] = MLP(x [rows][:dims[0]], row stride ldx);  w[l] [dims[l]][dims[l+1]] (the flax `kernel`), b[l]
 *   backward: dw[l], db[l] (overwritten) from dy [rows][dims[L]] = dLoss/dy of the LAST forward; parameters unchanged in between.
 * All pointers DEVICE fp32; launches on `stream`, no host synchronisation (create / destroy excepted). */
typedef struct pgtt_mlp pgtt_mlp;
const char* pgtt_mlp_last_error(void);
int pgtt_mlp_create(int n_layers, const int* dims, int rows, int device, pgtt_mlp** out);
void pgtt_mlp_destroy(pgtt_mlp* m);
int pgtt_mlp_rows(const pgtt_mlp* m);
int pgtt_mlp_forward(pgtt_mlp* m, const float* x, int ldx, const float* const* w, const float* const* b, float* y, void* stream);
/* forward with the learner's input fused in: row (t, j) of the minibatch = row t * S + idx[j] of the time-major store `data` [T][S][ld]
 * (idx: DEVICE int64 [mb]; rows = T_used * mb), normalised as (x - mean[c]) * inv_std[c] when both are given (brax running_statistics) */
int pgtt_mlp_forward_gather(pgtt_mlp* m, const float* data, int ld, int S, const long long* idx, int mb, const float* mean, const float* inv_std,
                            const float* const* w, const float* const* b, float* y, void* stream);
int pgtt_mlp_backward(pgtt_mlp* m, const float* dy, float* const* dw, float* const* db, void* stream);
int pgtt_mlp_debug_trace(pgtt_mlp* m, unsigned long long* out, int max_launches);   /* development aid, see csrc/pgtt_mlp.cu */

#ifdef __cplusplus
}
#endif
#endif
