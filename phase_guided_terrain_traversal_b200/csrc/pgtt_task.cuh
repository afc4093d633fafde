// pgtt_task.cuh - the TASK kernel: everything of Joystick.step that is not mjx.step, one warp per env.
//
// A control step is two launches: the physics kernel of the handle's generation (pgtt_env_kernel / pgtt_quad_kernel:
// 4 x mjx.step, go2/joystick_pgtt.py:145-148) leaves mjx.Data in the handle's buffers, and this kernel does the rest of
// go2/joystick_pgtt.py:149-231 from those buffers: compute_contact (go2/base.py:153-171), the 117-ray height scan
// (go2/heightmap.py:25-67) with its quadrant statistics, the observation with threefry noise, the 21 rewards, the info
// bookkeeping (gait phase, command resampling), brax's EpisodeWrapper / playground's BraxAutoResetWrapper and - for the
// rollout collector (training/train.py:135-161) - the transition slot of generate_unroll, so no separate record launch
// exists. The work here is wide (117 rays, 171 noise draws, 386 observation words per env) where the physics is a long
// dependent chain: one warp per env with every lane busy, instead of riding at the end of the physics warps' chain.
#pragma once
#include "pgtt_env_common.cuh"

// per-warp shared memory of the task kernel
struct TaskWS {
  float box[NBOX][BOXF];   // the boxes whose footprint can reach the ray grid, compacted
  float scan[NRAY];
  float qpos[20], qvel[NV], sens[NSENSOR], actf[NU], foot[NLEG][3], xmat0[9];
  float phase[4], air[4], last_act[NU], command[4];
  int last_contact[4];
};

// where generate_unroll files this step (one time-major slot; any pointer may be null)
struct RecordSlot { float *obs_state, *obs_priv, *reward, *discount, *truncation; };

// ----------------------------------------------------------------------------------------------
// create_sensor_matrix (go2/heightmap.py:25-67): 13x9 vertical rays against floor + boxes of the env's terrain
// (`boxes`: global, [100][BOXF]). Writes hit points to `out` ([117][3], may be null) and their z to t.scan.
// Broad phase: the boxes whose footprint can reach the grid are compacted into t.box (one ballot per 32 boxes);
// the box loop is the OUTER loop and a lane's rays stay in registers (each listed box is read once per lane).
// ----------------------------------------------------------------------------------------------
DEV void t_heightscan(TaskWS& t, const float* boxes, float cx, float cyy, float cz, float yaw, float* out, int lane) {
  float sy, cy;
  sincos_(yaw, &sy, &cy);
  const float oz = cz + 0.6f;
  const int nb = GC.n_boxes;
  int nl = 0;
  if (nb > 0) {
    const float4* bp = reinterpret_cast<const float4*>(boxes);
    const unsigned lt = (1u << lane) - 1u;
    float4 c0[4], c1[4];
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int k = it * 32 + lane, kc = k < nb ? k : nb - 1;
      c0[it] = ldg4(bp + 2 * kc); c1[it] = ldg4(bp + 2 * kc + 1);
    }
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int k = it * 32 + lane;
      bool near = false;
      if (k < nb) {
        const float dx = c0[it].x - cx, dy = c0[it].y - cyy;
        const float rad = sqrtf(c0[it].w * c0[it].w + c1[it].x * c1[it].x) + 0.75f;  // grid half-diagonal 0.7211 + slack
        near = dx * dx + dy * dy <= rad * rad;
      }
      const unsigned m = wballot(near);
      if (near) {
        float4* dst = reinterpret_cast<float4*>(t.box[nl + popc(m & lt)]);
        dst[0] = c0[it]; dst[1] = c1[it];
      }
      nl += popc(m);
    }
    syncwarp();
  }
  constexpr int RPL = (NRAY + 31) / 32;
  float ox[RPL], oy[RPL], best[RPL];
#pragma unroll
  for (int q = 0; q < RPL; q++) {
    const int r = (lane + 32 * q < NRAY) ? lane + 32 * q : NRAY - 1;
    const int i = r / NRAY_W, j = r % NRAY_W;
    const float p = (6.0f - (float)i) * 0.1f, k = (4.0f - (float)j) * 0.1f;
    ox[q] = cx + (p * cy - k * sy); oy[q] = cyy + (p * sy + k * cy);
    if (i == 6 && j == 4) { ox[q] = cx; oy[q] = cyy; }
    best[q] = __int_as_float(0x7f800000);
    if (oz >= 0.f) best[q] = oz;  // floor plane z = 0
  }
#pragma unroll 1
  for (int b = 0; b < nl; b++) {
    const float4 b0 = *reinterpret_cast<const float4*>(t.box[b]), b1 = *reinterpret_cast<const float4*>(t.box[b] + 4);
    const float lz = oz - b0.z;
    const float ttop = lz - b1.y, tbot = lz + b1.y;  // (+-hz - lz) / -1
#pragma unroll
    for (int q = 0; q < RPL; q++) {
      const float rx = ox[q] - b0.x, ry = oy[q] - b0.y;
      const float lx = b1.z * rx + b1.w * ry, ly = -b1.w * rx + b1.z * ry;
      if (fabsf(lx) <= b0.w && fabsf(ly) <= b1.x) {
        if (ttop >= 0.f) best[q] = fminf(best[q], ttop);
        else if (tbot >= 0.f) best[q] = fminf(best[q], tbot);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < RPL; q++) {
    const int r = lane + 32 * q;
    if (r < NRAY) {
      const float z = oz - best[q];
      if (out) { out[3 * r] = ox[q]; out[3 * r + 1] = oy[q]; out[3 * r + 2] = z; }
      t.scan[r] = z;
    }
  }
  syncwarp();
}

DEV const float* t_env_boxes(const EnvBuffers& B, int env) {
  return GC.n_boxes > 0 ? B.terrain + (size_t)B.terrain_index[env] * NBOX * BOXF : nullptr;
}

// ----------------------------------------------------------------------------------------------
// observation (joystick_pgtt.py:238-370). `rng` is advanced by the five splits of _get_obs. Reads the staged
// mjx.Data / info values of t (sens, qpos, qvel, actf, xmat0, scan, phase, air, last_act, command, last_contact).
// ----------------------------------------------------------------------------------------------
DEV void t_write_obs(TaskWS& t, const EnvBuffers& B, int env, Key& rng, float gait_freq, int lane) {
  // five sequential (rng, key) = split(rng) of _get_obs: the rng chain is computed by every lane, but a
  // lane derives only the noise key of its own slot group (and the height-scan key), so the number of
  // threefry evaluations per lane is 5 + 1 + 1 + 1 + 4 instead of 10 + 4 + 4
  Key chain[5];
#pragma unroll
  for (int s = 0; s < 5; s++) { chain[s] = rng; rng = rng_split(rng, 2, 0); }
  float* o = B.obs_state + (size_t)env * GC.nobs;
  float* pr = B.obs_priv + (size_t)env * GC.npriv;
  const float lvl = GC.noise_level;
  const float* R = t.xmat0;
  // layout: the baseline variant (go2/joystick.py:333-341) has no phase block and no gait_freq
  const bool base_v = GC.variant != 0;
  const int o_scan = base_v ? 30 : 38, o_last = base_v ? 147 : 156, o_cmd = base_v ? 159 : 168, px = GC.nobs;
  {
    // slot group of this lane: gyro 0..2 | gravity 3..5 | joint pos 6..17 | joint vel 18..29
    const int grp = lane < 3 ? 0 : (lane < 6 ? 1 : (lane < 18 ? 2 : 3));
    const int idx = lane < 3 ? lane : (lane < 6 ? lane - 3 : (lane < 18 ? lane - 6 : lane - 18));
    const int cnt = grp < 2 ? 3 : 12;
    Key base = chain[0];
    if (grp == 1) base = chain[1];
    if (grp == 2) base = chain[2];
    if (grp == 3) base = chain[3];
    const float u = 2.f * rng_unit(rng_split(base, 2, 1), cnt, idx < cnt ? idx : 0) - 1.f;
    float v = 0.f;
    if (lane < 3) v = t.sens[lane] + u * lvl * GC.noise_gyro;
    else if (lane < 6) v = -R[6 + idx] + u * lvl * GC.noise_gravity;                       // xmat^T (0,0,-1)
    else if (lane < 18) v = (t.qpos[7 + idx] + u * lvl * GC.noise_joint_pos) - GC.default_pose[idx];
    else if (lane < 30) v = t.qvel[6 + idx] + u * lvl * GC.noise_joint_vel;
    if (lane < 30) { o[lane] = v; pr[lane] = v; }
  }
  const Key scan_key = rng_split(chain[4], 2, 1);   // the linvel key, re-used for the height scan (Q9)
  if (lane < 4 && !base_v) {
    float s, c;
    sincos_(t.phase[lane], &s, &c);
    o[30 + lane] = c; o[34 + lane] = s; pr[30 + lane] = c; pr[34 + lane] = s;
  }
  float zmin = __int_as_float(0x7f800000);
  for (int r = lane; r < NRAY; r += 32) zmin = fminf(zmin, t.scan[r]);
  zmin = warp_min(zmin);
  for (int r = lane; r < NRAY; r += 32) {
    const float z = (t.scan[r] - zmin) + (2.f * rng_unit(scan_key, NRAY, r) - 1.f) * lvl * GC.noise_heightscan;
    o[o_scan + r] = z; pr[o_scan + r] = z;
  }
  if (lane == 0 && !base_v) { o[155] = gait_freq; pr[155] = gait_freq; }
  if (lane < 12) { o[o_last + lane] = t.last_act[lane]; pr[o_last + lane] = t.last_act[lane]; }
  if (lane < 3) {
    o[o_cmd + lane] = t.command[lane]; pr[o_cmd + lane] = t.command[lane];
    pr[px + lane] = t.sens[19 + lane]; pr[px + 3 + lane] = t.sens[3 + lane]; pr[px + 6 + lane] = t.sens[16 + lane];
    pr[px + 41 + lane] = 0.f;
  }
  if (lane < 12) { pr[px + 9 + lane] = t.actf[lane]; pr[px + 25 + lane] = t.sens[37 + lane]; }
  if (lane < 4) { pr[px + 21 + lane] = (float)t.last_contact[lane]; pr[px + 37 + lane] = t.air[lane]; }
}

// history rolls of _get_obs (joystick_pgtt.py:319-334), `step` is info["step"] BEFORE the increment
DEV void t_update_history(const TaskWS& t, const EnvBuffers& B, int env, int step, const float* motor_targets, int lane) {
  const bool upd = (step % GC.history_update_steps == 0) && lane < 12;
  float* qv = B.qvel_hist + (size_t)env * 24;
  float* qe = B.qpos_err_hist + (size_t)env * 24;
  float a = 0.f, b = 0.f;
  if (upd) { a = qv[lane]; b = qe[lane]; }
  syncwarp();
  if (upd) {
    qv[12 + lane] = a; qe[12 + lane] = b;
    qv[lane] = t.qvel[6 + lane];
    qe[lane] = t.qpos[7 + lane] - motor_targets[lane];
  }
}

// mjx.Data of this env as the physics kernel left it -> shared memory
DEV void t_stage_data(TaskWS& t, const EnvBuffers& B, int env, int lane) {
  if (lane < NQ) t.qpos[lane] = B.qpos[env * NQ + lane];
  if (lane < NV) t.qvel[lane] = B.qvel[env * NV + lane];
  for (int i = lane; i < NSENSOR; i += 32) t.sens[i] = B.sensordata[env * NSENSOR + i];
  if (lane < NU) { t.actf[lane] = B.actuator_force[env * NU + lane]; (&t.foot[0][0])[lane] = B.site_xpos[env * 15 + 3 + lane]; }
  if (lane < 9) t.xmat0[lane] = B.site_xmat[env * 9 + lane];
  syncwarp();
}

// ----------------------------------------------------------------------------------------------
// Joystick.step after mjx_env.step (joystick_pgtt.py:149-231) + EpisodeWrapper + BraxAutoResetWrapper + record
// ----------------------------------------------------------------------------------------------
DEV void task_step(TaskWS& t, const EnvBuffers& B, const float* action_all, int env, int lane, int wrapped, const RecordSlot& rec) {
  const float dt = GC.ctrl_dt;
  // every independent global load of the step is issued up front (one DRAM round trip instead of a dozen dependent ones)
  const float* boxes = t_env_boxes(B, env);
  const float steps_in = wrapped ? B.steps[env] : 0.f, done_in = wrapped ? B.done[env] : 0.f;
  const float action = lane < NU ? action_all[(size_t)env * NU + lane] : 0.f;
  const int c_geom = lane < NCON ? B.contact_geom[env * NCON * 2 + 2 * lane + (lane < 4 ? 1 : 0)] : -1;
  const float c_dist = lane < NCON ? B.contact_dist[env * NCON + lane] : 0.f;
  const int last_contact = lane < 4 ? B.last_contact[env * 4 + lane] : 0;
  float air = lane < 4 ? B.feet_air_time[env * 4 + lane] : 0.f;
  const float swing_peak_in = lane < 4 ? B.swing_peak[env * 4 + lane] : 0.f;
  Key rng; rng.a = B.rng[env * 2]; rng.b = B.rng[env * 2 + 1];
  const float phase = lane < 4 ? B.phase[env * 4 + lane] : 0.f;
  const float last_act = lane < NU ? B.last_act[env * NU + lane] : 0.f;
  const float command = lane < 3 ? B.command[env * 3 + lane] : 0.f;
  const int step = B.step[env];
  const float gait_freq = B.gait_freq[env], phase_dt = B.phase_dt[env];
  const int steps_until_in = B.steps_until[env];
  const float prev_ep_done = wrapped ? B.episode_done[env] : 0.f;
  const float em_in = (wrapped && lane < 24) ? B.episode_metrics[(size_t)env * 24 + lane] : 0.f;
  const float time_in = B.time[env];
  t_stage_data(t, B, env, lane);
  // BraxAutoResetWrapper.step: steps <- 0 where the previous step ended an episode; done cleared
  float steps = (done_in != 0.f) ? 0.f : steps_in;
  // compute_contact (base.py:153-171): any listed contact of the foot with dist < 0. Flag order FR FL RR RL, contact
  // list = 4 foot/plane slots (floor geom, foot geom) + 4 foot/box slots (foot geom, box geom); empty slots hold -1.
  // Lane c < 8 holds contact c: one ballot per foot.
  int contact = 0, first_contact = 0;
  float swing_peak = 0.f;
  {
    const bool hit = lane < NCON && c_dist < 0.f;
    unsigned m[4];
#pragma unroll
    for (int f = 0; f < 4; f++) m[f] = wballot(hit && c_geom == GC.foot_geom[f]);
    if (lane < 4) {
      const int g = lane ^ 1;
#pragma unroll
      for (int f = 0; f < 4; f++) if (f == g) contact = m[f] != 0u;
      first_contact = (air > 0.f) && (contact | last_contact);
      air += dt;
      swing_peak = fmaxf(swing_peak_in, t.sens[25 + 3 * lane + 2]);
    }
  }
  // height scan at the post-step pose, quadrant statistics (joystick_pgtt.py:167-190, Q8)
  t_heightscan(t, boxes, t.qpos[0], t.qpos[1], t.qpos[2], quat_to_yaw(t.qpos + 3), B.heightscan + (size_t)env * NRAY * 3, lane);
  float hmax = 0.f;
  {
    // quadrant q of this lane group: rows/cols per joystick_pgtt.py:171-174 (n = 6)
    const int q = lane >> 3, sub = lane & 7;
    const int r0 = (q < 2) ? 0 : 7, r1 = (q < 2) ? 6 : 13, c0 = (q & 1) ? 0 : 7, c1 = (q & 1) ? 6 : 9;
    float mx = -__int_as_float(0x7f800000), mn = __int_as_float(0x7f800000);
    if (q & 1) {            // left quadrants: 6 rows x 6 columns (compile-time divisors)
      for (int k = sub; k < 36; k += 8) { const float z = t.scan[(r0 + k / 6) * NRAY_W + c0 + k % 6]; mx = fmaxf(mx, z); mn = fminf(mn, z); }
    } else {                // right quadrants: 6 rows x 2 columns
      for (int k = sub; k < 12; k += 8) { const float z = t.scan[(r0 + (k >> 1)) * NRAY_W + c0 + (k & 1)]; mx = fmaxf(mx, z); mn = fminf(mn, z); }
    }
    (void)r1; (void)c1;
    for (int o = 4; o > 0; o >>= 1) { mx = fmaxf(mx, shfl_xor(mx, o)); mn = fminf(mn, shfl_xor(mn, o)); }
    const float hm = GC.variant ? mx : mx - mn;   // joystick.py:186 vs joystick_pgtt.py:189
    // lane k (< 4) needs quadrant k
    hmax = shfl(hm, (lane & 3) * 8);
    const float hmin = shfl(mn, (lane & 3) * 8);
    if (lane < 4) { B.H_max[env * 4 + lane] = hmax; B.H_min[env * 4 + lane] = hmin; }
  }
  // observation (uses info BEFORE the bookkeeping below, except feet_air_time which is already += dt)
  if (lane < 4) { t.phase[lane] = phase; t.air[lane] = air; t.last_contact[lane] = last_contact; }
  if (lane < NU) t.last_act[lane] = last_act;
  if (lane < 3) t.command[lane] = command;
  syncwarp();
  t_write_obs(t, B, env, rng, gait_freq, lane);
  t_update_history(t, B, env, step, B.motor_targets + (size_t)env * NU, lane);
  // termination (joystick_pgtt.py:233-236) + a failure guard the reference does not have: a non-finite or
  // absurd generalised state ends the episode, so the auto-reset wrapper restores the env instead of carrying NaNs forever
  bool bad = false;
  if (lane < NQ) bad = !(fabsf(t.qpos[lane]) < 1e6f);
  if (lane < NV) bad |= !(fabsf(t.qvel[lane]) < 1e6f);
  const bool poisoned = any_lane(bad);
  const int done = (t.sens[24] < 0.f) || poisoned;
  // rewards (joystick_pgtt.py:372-599); lanes 0..3 hold per-foot partials, lanes 0..11 per-joint partials
  const float* sdat = t.sens;
  const float cmd0 = t.command[0], cmd1 = t.command[1], cmd2 = t.command[2];
  const float cmd_norm = sqrtf(cmd0 * cmd0 + cmd1 * cmd1 + cmd2 * cmd2);
  float ss = 0.f, pose = 0.f, lim = 0.f, t2 = 0.f, t1 = 0.f, ar = 0.f, en = 0.f;
  if (lane < 12) {
    const float q = t.qpos[7 + lane], dq = q - GC.default_pose[lane], af = t.actf[lane];
    ss = fabsf(dq);
    pose = dq * dq * ((lane % 3 == 0) ? 1.0f : 0.1f);
    const float a = q - GC.soft_lo[lane], b = q - GC.soft_hi[lane];
    lim = -(a < 0.f ? a : 0.f) + (b > 0.f ? b : 0.f);
    t2 = af * af; t1 = fabsf(af);
    ar = (action - last_act) * (action - last_act);
    en = fabsf(t.qvel[6 + lane]) * fabsf(af);
  }
  float slip = 0.f, clr = 0.f, perr = 0.f, swing = 0.f, airr = 0.f, con = 0.f, center = 0.f, fh = 0.f, footz = __int_as_float(0x7f800000);
  if (lane < 4) {
    const float* v = sdat + 37 + 3 * lane; const float* pf = sdat + 25 + 3 * lane;
    const float vxy2 = v[0] * v[0] + v[1] * v[1];
    slip = vxy2 * (float)contact;
    clr = (GC.variant ? fabsf(t.foot[lane ^ 1][2] - (hmax - GC.base_feet_distance + GC.swing_height))   // world-frame foot height, joystick.py:569-572
                      : fabsf(pf[2] - (hmax + GC.swing_height))) * sqrtf(sqrtf(vxy2));
    const float rz = gait_get_z(phase, hmax + GC.swing_height, GC.base_feet_distance);
    perr = (pf[2] - rz) * (pf[2] - rz);
    const int swing_mask = (phase / (2.f * PGTT_PI)) >= 0.5f;
    swing = (pf[2] - GC.swing_height) * (pf[2] - GC.swing_height) * (float)swing_mask;
    airr = (air - (GC.variant ? 0.5f : 0.1f)) * (float)first_contact;   // joystick.py:591 vs joystick_pgtt.py:597
    con = (float)(swing_mask && contact);
    center = pf[0] * pf[0] + pf[1] * pf[1];
    const float er = swing_peak / GC.swing_height - 1.f;
    fh = er * er * (float)first_contact;
    footz = t.foot[lane ^ 1][2];
  }
  float sums[15];
  {
    float vals[15] = {ss, pose, lim, t2, t1, ar, en, slip, clr, perr, swing, airr, con, center, fh};
#pragma unroll
    for (int k = 0; k < 15; k++) sums[k] = warp_sum(vals[k]);
  }
  footz = warp_min(footz);
  float reward = 0.f;
  float rw[NREW];
  {
    const float le = (cmd0 - sdat[19]) * (cmd0 - sdat[19]) + (cmd1 - sdat[20]) * (cmd1 - sdat[20]);
    rw[0] = expf(-le / GC.tracking_sigma);
    rw[1] = expf(-((cmd2 - sdat[2]) * (cmd2 - sdat[2])) / GC.tracking_sigma);
    rw[2] = sdat[15] * sdat[15];
    rw[3] = sdat[16] * sdat[16] + sdat[17] * sdat[17];
    rw[4] = sdat[22] * sdat[22] + sdat[23] * sdat[23];
    rw[5] = sums[2];
    rw[6] = sums[1];
    rw[7] = (float)done;
    rw[8] = sums[0] * (float)(cmd_norm < 0.01f);
    rw[9] = sqrtf(sums[3]) + sums[4];
    rw[10] = sums[5];
    rw[11] = sums[6];
    rw[12] = sums[8];
    rw[13] = sums[14] * (float)(cmd_norm > 0.01f);
    rw[14] = sums[7] * (float)(cmd_norm > 0.01f);
    rw[15] = sums[11] * (float)(cmd_norm > 0.01f);
    rw[16] = expf(-sums[9] / GC.phase_sigma);
    rw[17] = sums[10];
    const float bh = t.qpos[2] - footz - 0.27f;
    rw[18] = bh * bh;
    rw[19] = -sums[12];
    rw[20] = sums[13];
#pragma unroll
    for (int k = 0; k < NREW; k++) rw[k] *= GC.reward_scale[k];
    // sum in the dict order of _get_reward
    const float total = ((((((((((((((((((((rw[0] + rw[1]) + rw[2]) + rw[3]) + rw[4]) + rw[8]) + rw[7]) + rw[6]) + rw[9]) + rw[10]) + rw[11]) +
                        rw[14]) + rw[12]) + rw[16]) + rw[18]) + rw[17]) + rw[15]) + rw[5]) + rw[19]) + rw[20]) + rw[13]);
    reward = fminf(fmaxf(total * dt, 0.f), 10000.f);
  }
  // info bookkeeping (joystick_pgtt.py:205-224)
  if (lane < NU) { B.last_last_act[env * NU + lane] = last_act; B.last_act[env * NU + lane] = action; }
  if (lane < 4) B.phase[env * 4 + lane] = fmodf(phase + phase_dt, 2.f * PGTT_PI);
  int steps_until = steps_until_in - 1;
  const Key rng3 = rng;   // rng, key1, key2 = split(rng, 3): key1 / key2 are derived only when consumed
  rng = rng_split(rng, 3, 0);
  if (steps_until <= 0) {  // sample_command (joystick_pgtt.py:603-611)
    const Key key1 = rng_split(rng3, 3, 1);
    const Key y_rng = rng_split(key1, 4, 1), w_rng = rng_split(key1, 4, 2), z_rng = rng_split(key1, 4, 3);
    if (lane < 3) {
      const float y = rng_uniform(y_rng, 3, lane, GC.cmd_u_min[lane], GC.cmd_u_max[lane]);
      const float z = (float)(rng_unit(z_rng, 3, lane) < GC.cmd_b[lane]);
      const float ww = (float)(rng_unit(w_rng, 3, lane) < 0.5f);
      B.command[env * 3 + lane] = command - ww * (command - y * z);
    }
  }
  if (done || steps_until <= 0) steps_until = (int)rintf(-log1pf(-rng_unit(rng_split(rng3, 3, 2), 1, 0)) * 5.0f / dt);
  float sp_mean = 0.f;
  if (lane < 4) {
    air *= (float)(!contact);
    swing_peak *= (float)(!contact);
    B.feet_air_time[env * 4 + lane] = air;
    B.last_contact[env * 4 + lane] = contact;
    B.swing_peak[env * 4 + lane] = swing_peak;
    B.contact[env * 4 + lane] = contact;
    B.first_contact[env * 4 + lane] = first_contact;
    sp_mean = swing_peak;
  }
  sp_mean = warp_sum(sp_mean) * 0.25f;
  float metric = 0.f;
  if (lane < NREW) {
#pragma unroll
    for (int k = 0; k < NREW; k++) if (lane == k) metric = rw[k];
  } else if (lane == NREW) metric = sp_mean;
  if (lane < NMETRIC) B.metrics[env * NMETRIC + lane] = metric;
  if (lane == 0) {
    B.rng[env * 2] = rng.a; B.rng[env * 2 + 1] = rng.b;
    B.step[env] = step + 1;
    B.steps_until[env] = steps_until;
    B.time[env] = time_in + GC.dt * (float)GC.n_substeps;
  }
  float done_out = (float)done, trunc_out = 0.f;
  bool restored = false;
  if (wrapped) {
    // EpisodeWrapper.step
    steps += 1.f;
    const float done_inner = done_out;
    const bool over = steps >= (float)GC.episode_length;
    done_out = over ? 1.f : done_inner;
    trunc_out = over ? 1.f - done_inner : 0.f;
    // episode_metrics rows: [sum_reward, length, 22 metrics]; lane k holds row k, the step's increments come from lanes k - 2
    {
      const float inc_m = shfl(metric, lane >= 2 ? lane - 2 : 0);
      const float inc = lane == 0 ? reward : (lane == 1 ? 1.f : inc_m);
      if (lane < 24) B.episode_metrics[(size_t)env * 24 + lane] = (em_in + inc) * (1.f - prev_ep_done);
    }
    if (lane == 0) {
      B.truncation[env] = trunc_out;
      B.steps[env] = steps;
      B.episode_done[env] = done_out;
    }
    // auto-reset: restore the cached first data / obs only (info is NOT reset)
    if (done_out != 0.f) {
      restored = true;
      if (lane < NQ) B.qpos[env * NQ + lane] = B.first_qpos[env * NQ + lane];
      if (lane < NV) {
        B.qvel[env * NV + lane] = B.first_qvel[env * NV + lane];
        B.warm[env * NV + lane] = B.first_warm[env * NV + lane];
        B.qacc[env * NV + lane] = B.first_qacc[env * NV + lane];
      }
      if (lane < NU) { B.actuator_force[env * NU + lane] = B.first_actuator_force[env * NU + lane]; B.ctrl[env * NU + lane] = GC.home_qpos[7 + lane]; }
      for (int i = lane; i < NSENSOR; i += 32) B.sensordata[env * NSENSOR + i] = B.first_sensordata[env * NSENSOR + i];
      if (lane < 15) B.site_xpos[env * 15 + lane] = B.first_site_xpos[env * 15 + lane];
      if (lane < 9) B.site_xmat[env * 9 + lane] = B.first_site_xmat[env * 9 + lane];
      if (lane < NCON) B.contact_dist[env * NCON + lane] = B.first_contact_dist[env * NCON + lane];
      if (lane < 2 * NCON) B.contact_geom[env * NCON * 2 + lane] = B.first_contact_geom[env * NCON * 2 + lane];
      if (lane == 0) B.time[env] = 0.f;
    }
  }
  if (lane == 0) { B.reward[env] = reward; B.done[env] = done_out; }
  // observation rows: the restored first observation where the episode ended; filed into the rollout slot in the same pass
  // (generate_unroll: next_observation, reward, discount = 1 - done, truncation)
  const int nobs = GC.nobs, npriv = GC.npriv;
  float* o = B.obs_state + (size_t)env * nobs;
  float* pr = B.obs_priv + (size_t)env * npriv;
  if (restored) {
    const float* fo = B.first_obs_state + (size_t)env * nobs;
    const float* fp = B.first_obs_priv + (size_t)env * npriv;
    float* ro = rec.obs_state ? rec.obs_state + (size_t)env * nobs : nullptr;
    float* rp = rec.obs_priv ? rec.obs_priv + (size_t)env * npriv : nullptr;
    for (int i = lane; i < nobs; i += 32) { const float v = fo[i]; o[i] = v; if (ro) ro[i] = v; }
    for (int i = lane; i < npriv; i += 32) { const float v = fp[i]; pr[i] = v; if (rp) rp[i] = v; }
  } else if (rec.obs_state || rec.obs_priv) {
    syncwarp();   // the rows were written by other lanes of this warp
    if (rec.obs_state) { float* ro = rec.obs_state + (size_t)env * nobs; for (int i = lane; i < nobs; i += 32) ro[i] = o[i]; }
    if (rec.obs_priv) { float* rp = rec.obs_priv + (size_t)env * npriv; for (int i = lane; i < npriv; i += 32) rp[i] = pr[i]; }
  }
  if (lane == 0) {
    if (rec.reward) rec.reward[env] = reward;
    if (rec.discount) rec.discount[env] = 1.0f - done_out;
    if (rec.truncation) rec.truncation[env] = trunc_out;
  }
}

// create_sensor_matrix for caller-supplied centres / yaws (pgtt_heightscan)
DEV void task_scan(TaskWS& t, const EnvBuffers& B, const float* center, const float* yaw, float* out, int env, int lane) {
  t_heightscan(t, t_env_boxes(B, env), center[env * 3], center[env * 3 + 1], center[env * 3 + 2], yaw[env], out + (size_t)env * NRAY * 3, lane);
}
