// pgtt_types.h - constant table, per-env buffer table and per-warp workspace of the env kernels.
#pragma once
#include <stdint.h>

#define NLEG 4
#define NB 13     // moving bodies: base, then (hip, thigh, calf) x FL FR RL RR
#define NV 18
#define NQ 19
#define NU 12
#define NBOX 100
#define NCON 8
#define NRAY_H 13
#define NRAY_W 9
#define NRAY 117
#define NOBS 171
#define NPRIV 215
#define NREW 21
#define NMETRIC 22
#define NSENSOR 49
#define MAXCAND 32
#define BOXF 8    // floats per preprocessed box: px py pz hx hy hz cos sin

// Everything that is identical for all envs (float copies of model.py:Go2Model + task config).
struct ModelConst {
  float dt, gravity_z, impratio, tolerance, ls_tolerance, meaninertia, solver_scale;
  int iterations, ls_iterations, max_geom_pairs, max_contact_points, n_boxes, n_substeps;
  int sync_mask;   // which stages end in a CTA barrier (tuning knob, pgtt_api.cu)
  int quad_ls_vote;    // quad kernel: line-search loop exits on a CTA vote per iteration (PGTT_QUAD_LS_VOTE=1) instead of running ls_iterations rounds
  int quad_fullscan;   // quad kernel: always scan all boxes instead of the per-step near lists (test knob, PGTT_QUAD_FULLSCAN=1)
  float body_pos[NB][3], body_ipos[NB][3], body_I[NB][6];  // body-frame inertia tensor xx yy zz xy xz yz
  float jnt_lo[12], jnt_hi[12], dof_invw[12], calf_invw[4];
  float lim_solref[2], lim_solimp[5];
  float act_bias0[12], act_bias2[12], ctrl_lo[12], ctrl_hi[12], frc_lo[12], frc_hi[12];  // actuator order
  int act_of_hinge[12], hinge_of_act[12];
  float foot_pos[3], foot_r, foot_mu, includemargin, box_rbound;
  float floor_solref[2], floor_solimp[5], box_solref[2], box_solimp[5];  // already mixed with the foot's
  int foot_geom[4], floor_geom, box_geom0;
  float imu_pos[3];
  // nominal per-env parameters (what pgtt_randomize starts from)
  float nom_mass[NB], nom_ipos_base[3], nom_armature[12], nom_damping[12], nom_gain[12], nom_bias1[12], nom_qpos0[12];
  float nom_box_mu, nom_floor_mu;
  int n_model_bodies;
  // task (go2/configs.py)
  float ctrl_dt, action_scale, noise_level;
  float noise_joint_pos, noise_joint_vel, noise_gyro, noise_gravity, noise_linvel, noise_heightscan;
  float reward_scale[NREW], tracking_sigma, swing_height, base_feet_distance, phase_sigma;
  float cmd_u_max[3], cmd_u_min[3], cmd_b[3], gait_freq[2];
  float soft_lo[12], soft_hi[12], default_pose[12], home_qpos[NQ];
  int history_update_steps, episode_length, rng_partitionable;
  // task variant: 0 = phase-guided (go2/joystick_pgtt.py, obs 171 / 215), 1 = baseline (go2/joystick.py, obs 162 / 206: no phase,
  // no gait_freq; H_max = quadrant max; world-frame clearance; air-time threshold 0.5)
  int variant, nobs, npriv;
};

// Device pointers, all [N][dim] row-major (see include/pgtt_b200.h:pgtt_buffers).
struct EnvBuffers {
  int N;
  float *qpos, *qvel, *qacc, *warm, *ctrl, *time;
  float *sensordata, *actuator_force, *site_xpos, *site_xmat, *contact_dist;
  int *contact_geom, *solver_niter;
  float *obs_state, *obs_priv, *reward, *done, *metrics;
  uint32_t* rng;
  float* command;
  int *step, *steps_until;
  float *phase, *phase_dt, *gait_freq, *last_act, *last_last_act, *feet_air_time;
  int* last_contact;
  float *swing_peak, *H_max, *H_min, *heightscan, *motor_targets, *qpos_err_hist, *qvel_hist;
  int *contact, *first_contact;
  float *steps, *truncation, *episode_done, *episode_metrics;
  float *first_qpos, *first_qvel, *first_warm, *first_obs_state, *first_obs_priv;
  float *first_sensordata, *first_actuator_force, *first_site_xpos, *first_site_xmat, *first_contact_dist, *first_qacc;
  int* first_contact_geom;
  float *m_mass, *m_ipos, *m_armature, *m_damping, *m_gain, *m_bias1, *m_qpos0, *m_boxfric, *m_floorfric;
  int* terrain_index;
  const float* terrain;   // [T][100][BOXF] preprocessed boxes
  int n_terrains;
};

// near-box lists of the warp-per-env collision stage (per foot; overflow falls back to the full scan)
#define W_QPEN 8           // boxes whose surface can reach the foot within one control step
#define W_QCEN 16          // boxes whose centre can come within W_RCEN of the foot
#define W_MARGIN 0.12f     // foot travel covered by the lists (checked every substep)
#define W_RCEN 0.9f        // broad-phase thresholds below W_RCEN^2 are ranked from the centre lists

// Per-warp shared-memory workspace of the warp-per-env physics kernel (one env per warp). Kept under 8 KB so that 28
// warps (two CTAs of 14) fit one SM: 4096 envs are ONE resident wave on 148 SMs. Arrays that are only alive in one
// phase of mjx.forward share storage: position-stage scratch (link frames, composite inertias, RNE forces, collision
// candidates) with the factorisation / Hessian scratch of the solver.
struct WS {
  // state
  float qpos[20], qvel[NV], qacc[NV], warm[NV], ctrl[NU];
  // per-env model
  float mass[NB], ipos0[3], armature[12], damping[12], gain[12], bias1[12], qpos0[12];
  float mtot_inv, floor_mu;
  // kinematics that outlive the position stage
  float xpos0[3], xmat0[9], com[3];                 // base frame (= imu site orientation), subtree COM
  float cinert[NB][10], cdof[NV][6];
  float cdofd_base[6][6], cvel_base[6], cvel_calf[NLEG][6];   // what the sensors read of cdof_dot / cvel
  float foot[NLEG][3];
  // arrow inertia matrix: base block [6][6], coupling [leg][6][3], leg blocks [leg][3][3]
  float MB[36], MC[72], MA[36];
  // vectors
  float qs[NV], qas[NV], Ma[NV], grad[NV], search[NV], mv[NV], qfc[NV];
  float actf[NU];
  // contacts: slots 0..3 = foot g vs floor, 4..7 = selected foot-box contacts
  int c_leg[NCON], c_box[NCON];
  float c_dist[NCON], c_pos[NCON][3], c_frame[NCON][9], c_mu[NCON];
  float Jc[NCON][3][9], Ac[NCON][5], fc[NCON][3], limD[12];
  int nact, actlist[NCON];
  // near-box lists, built at the first substep of a control step
  int pen_list[NLEG][W_QPEN], cen_list[NLEG][W_QCEN];
  float near_f0[NLEG][3];
  int near_npen[NLEG], near_ncen[NLEG], near_ok;
  float sens[NSENSOR];
  int niter[4];
  int bar_threads;  // threads of this CTA that take part in stage barriers (32 x live warps)
  int tix;          // next slot of the stage-timestamp trace (development aid, tools/stage_trace.py)
  long long* trace; // null in normal launches
  union {
    struct {        // position / velocity stage (index 0 = base, 1+3g+t = leg g link t)
      float xpos[NB][3], xmat[NB][9], xipos[NB][3];
      union {
        float crb[NB][10];
        struct { int pair[MAXCAND]; float dist[MAXCAND], cd2[MAXCAND]; } cand;   // collision runs before the inertia stages
      };
      float F[NV][6], bias[NV];
    };
    struct {        // factorisation / solver stage
      float HB[36], HC[72], HA[36];
      float fLA[NLEG][6];       // Cholesky of A_g: l00 l10 l11 l20 l21 l22 (diagonals stored as reciprocals)
      float fY[NLEG][6][3];     // C_g A_g^-1
      float fS[36];             // Schur complement of the base block
      float fL[24];             // its Cholesky factor, row-packed lower, diagonals as reciprocals (21 used)
      float tb[6];
    };
  };
};
