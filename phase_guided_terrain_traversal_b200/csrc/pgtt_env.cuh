// pgtt_env.cuh - warp-per-env physics step (generation 1), reset and domain randomisation (device code).
//
// Reference semantics: go2/joystick_pgtt.py:50-148, go2/randomize.py:23-171, jax.random (threefry2x32),
// playground BraxAutoResetWrapper's first-state cache. Checked against oracle/pgtt_oracle.c.
// The task layer that follows mjx.step inside Joystick.step lives in pgtt_task.cuh (its own kernel).
#pragma once
#include "pgtt_task.cuh"

// ----------------------------------------------------------------------------------------------
DEV void store_data(WS& w, const EnvBuffers& B, int env, int lane) {
  if (lane < NQ) B.qpos[env * NQ + lane] = w.qpos[lane];
  if (lane < NV) { B.qvel[env * NV + lane] = w.qvel[lane]; B.warm[env * NV + lane] = w.warm[lane]; B.qacc[env * NV + lane] = w.qacc[lane]; }
  if (lane < NU) { B.ctrl[env * NU + lane] = w.ctrl[lane]; B.actuator_force[env * NU + lane] = w.actf[lane]; }
  for (int i = lane; i < NSENSOR; i += 32) B.sensordata[env * NSENSOR + i] = w.sens[i];
  if (lane < 3) B.site_xpos[env * 15 + lane] = w.sens[10 + lane];
  if (lane < 12) {  // sites FL FR RL RR = leg order
    B.site_xpos[env * 15 + 3 + lane] = w.foot[lane / 3][lane % 3];
  }
  if (lane < 9) B.site_xmat[env * 9 + lane] = w.xmat0[lane];
  if (lane < NCON) {
    const int c = lane;
    const int box = w.c_box[c];
    B.contact_dist[env * NCON + c] = (box == -2) ? 0.f : w.c_dist[c];
    int g1 = -1, g2 = -1;
    if (box == -1) { g1 = GC.floor_geom; g2 = GC.foot_geom[w.c_leg[c]]; }
    else if (box >= 0) { g1 = GC.foot_geom[w.c_leg[c]]; g2 = GC.box_geom0 + box; }
    B.contact_geom[env * NCON * 2 + 2 * c] = g1; B.contact_geom[env * NCON * 2 + 2 * c + 1] = g2;
  }
}
// ----------------------------------------------------------------------------------------------
// mjx_env.step of Joystick.step (joystick_pgtt.py:145-148): motor targets, n_substeps x mjx.step, one warp per env.
// Leaves mjx.Data (qpos, qvel, warm start, qacc, ctrl, sensordata, actuator_force, site frames, contact list) in the
// handle's buffers for the task kernel.
// ----------------------------------------------------------------------------------------------
DEV void env_physics(WS& w, const EnvBuffers& B, const float* action_all, int env, int lane, long long* trace = nullptr) {
  if (lane == 0) { w.trace = trace ? trace + (size_t)env * 40 : nullptr; w.tix = 0; }
  syncwarp();
  load_model(w, B, env, lane);
  load_state(w, B, env, lane);
  if (lane < NU) {
    const float mt = GC.default_pose[lane] + action_all[(size_t)env * NU + lane] * GC.action_scale;
    w.ctrl[lane] = mt; B.motor_targets[env * NU + lane] = mt;
  }
  syncwarp();
  for (int s = 0; s < GC.n_substeps; s++) {
    const int ni = forward(w, B, env, lane, s == GC.n_substeps - 1, s == 0);
    if (lane == 0) w.niter[s & 3] = ni;
    euler(w, lane);
  }
  if (lane < 4 && lane < GC.n_substeps) B.solver_niter[env * 4 + lane] = w.niter[lane];
  store_data(w, B, env, lane);
  STAGE_TRACE(w, lane);   // 28: stored
}

// mjx.Data of the workspace -> the task layer's staging area (reset runs both layers in one kernel)
DEV void stage_task_from_ws(TaskWS& t, const WS& w, int lane) {
  if (lane < NQ) t.qpos[lane] = w.qpos[lane];
  if (lane < NV) t.qvel[lane] = w.qvel[lane];
  for (int i = lane; i < NSENSOR; i += 32) t.sens[i] = w.sens[i];
  if (lane < NU) { t.actf[lane] = w.actf[lane]; (&t.foot[0][0])[lane] = (&w.foot[0][0])[lane]; }
  if (lane < 9) t.xmat0[lane] = w.xmat0[lane];
  syncwarp();
}

// ----------------------------------------------------------------------------------------------
// Joystick.reset (joystick_pgtt.py:50-131) + wrapper resets
// ----------------------------------------------------------------------------------------------
DEV void env_reset(WS& w, TaskWS& t, const EnvBuffers& B, const uint32_t* keys, int env, int lane) {
  const float* boxes = t_env_boxes(B, env);
  load_model(w, B, env, lane);
  Key rng; rng.a = keys[2 * env]; rng.b = keys[2 * env + 1];
  Key key;
  key = rng_split(rng, 2, 1); rng = rng_split(rng, 2, 0);
  if (lane < NQ) w.qpos[lane] = GC.home_qpos[lane];
  if (lane < NV) { w.qvel[lane] = 0.f; w.warm[lane] = 0.f; }
  syncwarp();
  if (lane < 2) w.qpos[lane] += rng_uniform(key, 2, lane, -0.5f, 0.5f);
  key = rng_split(rng, 2, 1); rng = rng_split(rng, 2, 0);
  if (lane == 3) {
    const float yaw = rng_uniform(key, 1, 0, -3.14f, 3.14f);
    float s, c;
    sincos_(yaw * 0.5f, &s, &c);
    const float aw = w.qpos[3], ax = w.qpos[4], ay = w.qpos[5], az = w.qpos[6];
    // quat_mul(home_quat, (c, 0, 0, s))
    w.qpos[3] = aw * c - az * s; w.qpos[4] = ax * c + ay * s; w.qpos[5] = ay * c - ax * s; w.qpos[6] = aw * s + az * c;
  }
  key = rng_split(rng, 2, 1); rng = rng_split(rng, 2, 0);
  if (lane < 6) w.qvel[lane] = rng_uniform(key, 6, lane, -0.1f, 0.1f);
  syncwarp();
  if (lane < NU) w.ctrl[lane] = w.qpos[7 + lane];
  syncwarp();
  forward(w, B, env, lane, false, true);                 // mjx_env.init
  float* hs = B.heightscan + (size_t)env * NRAY * 3;
  t_heightscan(t, boxes, w.qpos[0], w.qpos[1], w.qpos[2], 0.f, hs, lane);     // yaw = 0 (Q10)
  float zmax = -__int_as_float(0x7f800000);
  for (int r = lane; r < NRAY; r += 32) zmax = fmaxf(zmax, t.scan[r]);
  zmax = warp_max(zmax);
  if (lane == 0) w.qpos[2] += zmax;
  syncwarp();
  forward(w, B, env, lane, true, true);                  // mjx.forward at the lifted pose
  const Key key1 = rng_split(rng, 3, 1), key2 = rng_split(rng, 3, 2);
  rng = rng_split(rng, 3, 0);
  const float tcmd = -log1pf(-rng_unit(key1, 1, 0)) * 5.0f;
  const int steps_until = (int)rintf(tcmd / GC.ctrl_dt);
  float command = 0.f;
  if (lane < 3) command = rng_uniform(key2, 3, lane, GC.cmd_u_min[lane], GC.cmd_u_max[lane]);
  key = rng_split(rng, 2, 1); rng = rng_split(rng, 2, 0);
  const float gait_freq = rng_uniform(key, 1, 0, GC.gait_freq[0], GC.gait_freq[1]);
  t_heightscan(t, boxes, w.qpos[0], w.qpos[1], w.qpos[2], 0.f, hs, lane);
  stage_task_from_ws(t, w, lane);
  // info
  if (lane < 4) {
    const float ph = (lane == 1 || lane == 2) ? PGTT_PI : 0.f;
    B.phase[env * 4 + lane] = ph;
    B.feet_air_time[env * 4 + lane] = 0.f; B.last_contact[env * 4 + lane] = 0; B.swing_peak[env * 4 + lane] = 0.f;
    B.H_max[env * 4 + lane] = 0.1f; B.H_min[env * 4 + lane] = 0.f;
    B.contact[env * 4 + lane] = 0; B.first_contact[env * 4 + lane] = 0;
    t.phase[lane] = ph; t.air[lane] = 0.f; t.last_contact[lane] = 0;
  }
  if (lane < NU) { B.last_act[env * NU + lane] = 0.f; B.last_last_act[env * NU + lane] = 0.f; B.motor_targets[env * NU + lane] = 0.f; t.last_act[lane] = 0.f; }
  if (lane < 24) { B.qpos_err_hist[env * 24 + lane] = 0.f; B.qvel_hist[env * 24 + lane] = 0.f; }
  if (lane < 3) { B.command[env * 3 + lane] = command; t.command[lane] = command; }
  if (lane < NMETRIC) B.metrics[env * NMETRIC + lane] = 0.f;
  if (lane < 24) B.episode_metrics[env * 24 + lane] = 0.f;
  syncwarp();
  t_write_obs(t, B, env, rng, gait_freq, lane);
  t_update_history(t, B, env, 0, B.motor_targets + (size_t)env * NU, lane);
  if (lane == 0) {
    B.rng[env * 2] = rng.a; B.rng[env * 2 + 1] = rng.b;
    B.step[env] = 0; B.steps_until[env] = steps_until;
    B.phase_dt[env] = 2.f * PGTT_PI * GC.ctrl_dt * gait_freq;
    B.gait_freq[env] = gait_freq;
    B.reward[env] = 0.f; B.done[env] = 0.f; B.time[env] = 0.f;
    B.steps[env] = 0.f; B.truncation[env] = 0.f; B.episode_done[env] = 0.f;
    for (int s = 0; s < 4; s++) B.solver_niter[env * 4 + s] = 0;
  }
  store_data(w, B, env, lane);
  syncwarp();
  // auto-reset cache
  if (lane < NQ) B.first_qpos[env * NQ + lane] = w.qpos[lane];
  if (lane < NV) { B.first_qvel[env * NV + lane] = w.qvel[lane]; B.first_warm[env * NV + lane] = w.warm[lane]; B.first_qacc[env * NV + lane] = w.qacc[lane]; }
  if (lane < NU) B.first_actuator_force[env * NU + lane] = w.actf[lane];
  for (int i = lane; i < NSENSOR; i += 32) B.first_sensordata[env * NSENSOR + i] = w.sens[i];
  if (lane < 15) B.first_site_xpos[env * 15 + lane] = B.site_xpos[env * 15 + lane];
  if (lane < 9) B.first_site_xmat[env * 9 + lane] = w.xmat0[lane];
  if (lane < NCON) B.first_contact_dist[env * NCON + lane] = B.contact_dist[env * NCON + lane];
  if (lane < 2 * NCON) B.first_contact_geom[env * NCON * 2 + lane] = B.contact_geom[env * NCON * 2 + lane];
  for (int i = lane; i < GC.nobs; i += 32) B.first_obs_state[(size_t)env * GC.nobs + i] = B.obs_state[(size_t)env * GC.nobs + i];
  for (int i = lane; i < GC.npriv; i += 32) B.first_obs_priv[(size_t)env * GC.npriv + i] = B.obs_priv[(size_t)env * GC.npriv + i];
}

// ----------------------------------------------------------------------------------------------
// domain_randomize (go2/randomize.py:23-171 / randomize_simple.py:24-138): one THREAD per env
// ----------------------------------------------------------------------------------------------
DEV void env_randomize(const EnvBuffers& B, const uint32_t* keys, int env, int dyn) {
  Key rng; rng.a = keys[2 * env]; rng.b = keys[2 * env + 1];
  Key key;
  const bool stairs = GC.n_boxes > 0;
#define NEXT_KEY() { key = rng_split(rng, 2, 1); rng = rng_split(rng, 2, 0); }
  NEXT_KEY();  // floor friction: discarded by the stairs variant (Q5)
  B.m_floorfric[env] = (!stairs && dyn) ? rng_uniform(key, 1, 0, 0.4f, 1.0f) : GC.nom_floor_mu;
  if (stairs) {
    NEXT_KEY();
    for (int k = 0; k < NBOX; k++) B.m_boxfric[env * NBOX + k] = dyn ? rng_uniform(key, NBOX, k, 0.4f, 1.0f) : GC.nom_box_mu;
  }
  NEXT_KEY();  // frictionloss scaling: frictionloss is 0
  NEXT_KEY();
  for (int i = 0; i < 12; i++) B.m_armature[env * 12 + i] = dyn ? GC.nom_armature[i] * rng_uniform(key, 12, i, 1.0f, 1.05f) : GC.nom_armature[i];
  NEXT_KEY();
  for (int i = 0; i < 3; i++) B.m_ipos[env * 3 + i] = dyn ? GC.nom_ipos_base[i] + rng_uniform(key, 3, i, -0.05f, 0.05f) : GC.nom_ipos_base[i];
  NEXT_KEY();
  for (int b = 0; b < NB; b++) B.m_mass[env * NB + b] = dyn ? GC.nom_mass[b] * rng_uniform(key, GC.n_model_bodies, b + 1, 0.9f, 1.1f) : GC.nom_mass[b];
  NEXT_KEY();
  if (dyn) B.m_mass[env * NB] += rng_uniform(key, 1, 0, -1.0f, 1.0f);
  NEXT_KEY();
  for (int i = 0; i < 12; i++) B.m_qpos0[env * 12 + i] = dyn ? GC.nom_qpos0[i] + rng_uniform(key, 12, i, -0.05f, 0.05f) : GC.nom_qpos0[i];
  NEXT_KEY();
  for (int i = 0; i < 12; i++) B.m_damping[env * 12 + i] = dyn ? GC.nom_damping[i] * rng_uniform(key, 12, i, 0.9f, 1.1f) : GC.nom_damping[i];
  NEXT_KEY();
  for (int a = 0; a < 12; a++) {
    const float g = dyn ? rng_uniform(key, 12, a, 0.9f, 1.1f) : 1.0f;
    B.m_gain[env * 12 + a] = GC.nom_gain[a] * g;
    B.m_bias1[env * 12 + a] = GC.nom_bias1[a] * g;
  }
  int tidx = 0;
  if (stairs) {
    NEXT_KEY();
    tidx = rng_randint(key, 0, B.n_terrains);
  }
  B.terrain_index[env] = tidx;
#undef NEXT_KEY
}
// mjx.forward on the stored state (refreshes derived data fields)
DEV void env_forward(WS& w, const EnvBuffers& B, int env, int lane) {
  load_model(w, B, env, lane);
  load_state(w, B, env, lane);
  if (lane < NU) w.ctrl[lane] = B.ctrl[env * NU + lane];
  syncwarp();
  forward(w, B, env, lane, true, true);
  store_data(w, B, env, lane);
}
