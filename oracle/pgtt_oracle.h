/* pgtt_oracle.h - data layout of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this directory.  The product path (CUDA, csrc/) never links it.
 *
 * PARITY UNPINNED: the reference ships no tests / golden vectors and its arithmetic lives in
 * un-vendored third-party packages (mujoco-mjx, mujoco_playground, brax, jax) that are not
 * installable here.  This file restates their published algorithms for the one
 * model family the reference uses; each function cites the reference call site it stands in for.
 */
#ifndef PGTT_ORACLE_H
#define PGTT_ORACLE_H

#include <stdint.h>

#ifdef ORC_F32
typedef float real;
#else
typedef double real;
#endif

#define NBODY 14
#define NV 18
#define NQ 19
#define NU 12
#define NHINGE 12
#define NBOX 100
#define NFOOT 4
#define NCON 8
#define NEFC (NHINGE + 4 * NCON) /* 12 limit rows + 8 contacts x 4 pyramid edges */
#define NRAY_H 13
#define NRAY_W 9
#define NRAY (NRAY_H * NRAY_W)
#define NOBS 171
#define NPRIV 215
#define NREW 21
#define NMETRIC 22
#define NSENSOR 49
#define NHIST 24

typedef struct {
  /* options (go2_mjx_feetonly.xml:4-19, base.py:57) */
  real timestep, gravity[3], impratio, tolerance, ls_tolerance, meaninertia;
  int iterations, ls_iterations, max_geom_pairs, max_contact_points, n_boxes;
  /* kinematic tree */
  int body_parent[NBODY];
  real body_pos[NBODY][3], body_quat[NBODY][4], body_ipos[NBODY][3], body_iquat[NBODY][4];
  real body_mass[NBODY], body_inertia[NBODY][3], body_invweight0[NBODY][2];
  int jnt_body[NHINGE];
  real jnt_axis[NHINGE][3], jnt_range[NHINGE][2], jnt_solref[2], jnt_solimp[5];
  real qpos0[NQ], dof_armature[NV], dof_damping[NV], dof_invweight0[NV];
  /* actuators, MJCF order FR FL RR RL */
  int act_dof[NU];
  real act_gain[NU], act_bias[NU][3], act_ctrlrange[NU][2], act_forcerange[NU][2];
  /* feet (geom-id order FL FR RL RR), floor, boxes */
  int foot_body[NFOOT], foot_geom_id[NFOOT];
  real foot_pos[3], foot_radius, foot_friction[3], foot_solref[2], foot_solimp[5], foot_margin;
  int floor_geom_id, box_geom_id0;
  real floor_friction[3], floor_solref[2], floor_solimp[5];
  real box_rbound, box_solref[2], box_solimp[5];
  real box_friction[NBOX][3];
  real box_pos[NBOX][3], box_quat[NBOX][4], box_size[NBOX][3];
  real imu_pos[3];
} Model;

typedef struct {
  real dist, pos[3], frame[9], mu, solref[2], solimp[5], includemargin;
  int geom1, geom2, body, foot, box; /* body = calf body id; box = -1 for the floor plane */
} Contact;

typedef struct {
  /* state */
  real qpos[NQ], qvel[NV], ctrl[NU], qacc[NV], qacc_warmstart[NV], time;
  /* position stage */
  real xpos[NBODY][3], xquat[NBODY][4], xmat[NBODY][9], xipos[NBODY][3], ximat[NBODY][9];
  real xanchor[NHINGE][3], xaxis[NHINGE][3];
  real subtree_com[3];
  real cinert[NBODY][10], crb[NBODY][10], cdof[NV][6];
  real qM[NV][NV], qL[NV][NV]; /* dense inertia and its Cholesky factor */
  real foot_xpos[NFOOT][3], site_xpos[5][3], site_xmat[9]; /* sites: imu, FL, FR, RL, RR; all-site xmat = body xmat */
  Contact contact[NCON];
  int ncon;
  real efc_J[NEFC][NV], efc_D[NEFC], efc_aref[NEFC], efc_pos[NEFC], efc_force[NEFC];
  /* velocity / acceleration stage */
  real cvel[NBODY][6], cdof_dot[NV][6], cacc[NBODY][6];
  real qfrc_bias[NV], qfrc_passive[NV], qfrc_actuator[NV], qfrc_smooth[NV], qacc_smooth[NV];
  real qfrc_constraint[NV];
  real actuator_force[NU];
  real sensordata[NSENSOR];
  int solver_niter;
} Data;

typedef struct {
  real ctrl_dt, action_scale, noise_level;
  real noise_joint_pos, noise_joint_vel, noise_gyro, noise_gravity, noise_linvel, noise_heightscan;
  real reward_scale[NREW]; /* order of go2/configs.py:31-59 */
  real tracking_sigma, swing_height, base_feet_distance, phase_sigma;
  real cmd_u_max[3], cmd_u_min[3], cmd_b[3], gait_freq[2];
  real soft_limit_factor;
  real default_pose[NHINGE], home_qpos[NQ];
  int history_update_steps, episode_length, n_substeps, rng_partitionable;
  int variant; /* 0 = go2/joystick_pgtt.py, 1 = go2/joystick.py (baseline task: obs 162 / 206, first entries of the arrays) */
} TaskCfg;

typedef struct {
  uint32_t rng[2];
  real command[3];
  int step, steps_until_next_cmd;
  real phase[4], phase_dt, gait_freq;
  real last_act[NU], last_last_act[NU], feet_air_time[4];
  int last_contact[4];
  real swing_peak[4], H_max[4], H_min[4], heightscan[NRAY][3], motor_targets[NU];
  real qpos_error_history[NHIST], qvel_history[NHIST];
  /* wrapper keys (brax EpisodeWrapper / playground BraxAutoResetWrapper) */
  real steps, truncation, episode_done, episode_metrics[2 + NMETRIC]; /* sum_reward, length, metrics */
} Info;

typedef struct {
  Model m;
  Data d;
  Info info;
  real obs_state[NOBS], obs_priv[NPRIV], reward, done, metrics[NMETRIC];
  int contact_flags[4], first_contact[4]; /* FR FL RR RL */
  int terrain_index;
  /* auto-reset cache */
  Data first_data;
  real first_obs_state[NOBS], first_obs_priv[NPRIV];
  /* optional externally supplied uniform noise in [0,1) replacing the 5 obs-noise draws (parity mode) */
} Env;

#endif
