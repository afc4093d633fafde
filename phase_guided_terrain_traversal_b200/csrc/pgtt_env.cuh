// pgtt_env.cuh - task layer fused around the physics: ray grid, gait phase, observation, rewards,
// info bookkeeping, episode / auto-reset wrappers, reset and domain randomisation (device code).
//
// Reference semantics: go2/joystick_pgtt.py:50-611, go2/heightmap.py:25-67, go2/gait.py:8-49,
// go2/base.py:153-171, go2/randomize.py:23-171, jax.random (threefry2x32), brax EpisodeWrapper and
// playground BraxAutoResetWrapper. Checked against oracle/pgtt_oracle.c.
#pragma once
#include "pgtt_physics.cuh"

// ----------------------------------------------------------------------------------------------
// jax.random
// ----------------------------------------------------------------------------------------------
struct Key { uint32_t a, b; };

DEV uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
// noinline: ~170 instructions, called from ~25 sites of the obs / command / reset code
DEV_NOINLINE Key threefry(Key k, uint32_t x0, uint32_t x1) {
  const uint32_t ks0 = k.a, ks1 = k.b, ks2 = k.a ^ k.b ^ 0x1BD11BDAu;
  x0 += ks0; x1 += ks1;
#define TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)
  x0 += ks1; x1 += ks2 + 1u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)
  x0 += ks2; x1 += ks0 + 2u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)
  x0 += ks0; x1 += ks1 + 3u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)
  x0 += ks1; x1 += ks2 + 4u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef TF_R
  Key o; o.a = x0; o.b = x1;
  return o;
}
// split(key, num)[i]
DEV Key rng_split(Key k, int num, int i) {
  if (GC.rng_partitionable) return threefry(k, 0u, (uint32_t)i);
  Key o;
  {
    const int flat = 2 * i, pair = flat % num, which = flat / num;
    const Key t = threefry(k, (uint32_t)pair, (uint32_t)(pair + num));
    o.a = which ? t.b : t.a;
  }
  {
    const int flat = 2 * i + 1, pair = flat % num, which = flat / num;
    const Key t = threefry(k, (uint32_t)pair, (uint32_t)(pair + num));
    o.b = which ? t.b : t.a;
  }
  return o;
}
// random_bits(key, 32, (n,))[i]
DEV uint32_t rng_bits(Key k, int n, int i) {
  if (GC.rng_partitionable) { const Key t = threefry(k, 0u, (uint32_t)i); return t.a ^ t.b; }
  const int half = (n + 1) / 2, pair = i % half, which = i / half;
  uint32_t x1 = (uint32_t)(pair + half);
  if ((n & 1) && pair + half >= n) x1 = 0u;
  const Key t = threefry(k, (uint32_t)pair, x1);
  return which ? t.b : t.a;
}
DEV float rng_unit(Key k, int n, int i) { return __uint_as_float((rng_bits(k, n, i) >> 9) | 0x3F800000u) - 1.0f; }
DEV float rng_uniform(Key k, int n, int i, float lo, float hi) { return fmaxf(lo, mul_add_nofma(rng_unit(k, n, i), hi - lo, lo)); }
DEV int rng_randint(Key k, int lo, int hi) {  // shape (1,)
  const Key k1 = rng_split(k, 2, 0), k2 = rng_split(k, 2, 1);
  const uint32_t hb = rng_bits(k1, 1, 0), lb = rng_bits(k2, 1, 0);
  uint32_t span = (uint32_t)(hi - lo);
  if (hi <= lo) span = 1u;
  const uint32_t mult = ((65536u % span) * (65536u % span)) % span;
  return lo + (int)(((hb % span) * mult + (lb % span)) % span);
}

// ----------------------------------------------------------------------------------------------
// go2/gait.py:27-49 and go2/utility.py:4-8
// ----------------------------------------------------------------------------------------------
DEV float gait_get_z(float phi, float h_max, float stance) {
  const float T_swing = 2.f * PGTT_PI * (1.f - 0.5f) / 2.f, T_peak = 2.f * PGTT_PI * (1.f + 0.5f) / 2.f, T_stance = 2.f * PGTT_PI * 0.5f;
  if (phi <= T_stance) return stance;
  float p0, p1, t;
  if (phi <= T_peak) { p0 = stance; p1 = h_max; t = (phi - T_stance) / T_swing; }
  else { p0 = h_max; p1 = stance; t = (phi - T_peak) / T_swing; }
  const float t2 = t * t, t3 = t2 * t;
  return (2.f * t3 - 3.f * t2 + 1.f) * p0 + (-2.f * t3 + 3.f * t2) * p1;
}
DEV float quat_to_yaw(const float* q) {
  return atan2f(2.f * (q[0] * q[3] + q[1] * q[2]), 1.f - 2.f * (q[2] * q[2] + q[3] * q[3]));
}

// ----------------------------------------------------------------------------------------------
// create_sensor_matrix (go2/heightmap.py:25-67): 13x9 vertical rays against floor + boxes.
// Writes hit points to `out` ([117][3], global) and their z to w.scan. `boxlist` receives the
// boxes whose footprint can reach the grid (broad phase over the 100 boxes, one ballot per 32).
// ----------------------------------------------------------------------------------------------
DEV void heightscan(WS& w, const float* center, float yaw, float* out, int* boxlist, int lane) {
  float sy, cy;
  sincos_(yaw, &sy, &cy);
  const float cx = center[0], cyy = center[1], oz = center[2] + 0.6f;
  const int nb = GC.n_boxes;
  int nl = 0;
  if (nb > 0) {
    const unsigned lt = (1u << lane) - 1u;
    for (int it = 0; it < 4; it++) {
      const int k = it * 32 + lane;
      bool near = false;
      if (k < nb) {
        const float* bx = w.box[k];
        const float dx = bx[0] - cx, dy = bx[1] - cyy;
        const float rad = sqrtf(bx[3] * bx[3] + bx[4] * bx[4]) + 0.75f;  // grid half-diagonal 0.7211 + slack
        near = dx * dx + dy * dy <= rad * rad;
      }
      const unsigned m = wballot(near);
      if (near) boxlist[nl + popc(m & lt)] = k;
      nl += popc(m);
    }
    syncwarp();
  }
  for (int r = lane; r < NRAY; r += 32) {
    const int i = r / NRAY_W, j = r % NRAY_W;
    const float p = (6.0f - (float)i) * 0.1f, k = (4.0f - (float)j) * 0.1f;
    float ox = cx + (p * cy - k * sy), oy = cyy + (p * sy + k * cy);
    if (i == 6 && j == 4) { ox = cx; oy = cyy; }
    float best = __int_as_float(0x7f800000);
    if (oz >= 0.f) best = oz;  // floor plane z = 0
    for (int t = 0; t < nl; t++) {
      const float* bx = w.box[boxlist[t]];
      const float rx = ox - bx[0], ry = oy - bx[1], lz = oz - bx[2];
      const float lx = bx[6] * rx + bx[7] * ry, ly = -bx[7] * rx + bx[6] * ry;
      if (fabsf(lx) <= bx[3] && fabsf(ly) <= bx[4]) {
        const float ttop = lz - bx[5], tbot = lz + bx[5];  // (+-hz - lz) / -1
        if (ttop >= 0.f) best = fminf(best, ttop);
        else if (tbot >= 0.f) best = fminf(best, tbot);
      }
    }
    const float z = oz - best;
    out[3 * r] = ox; out[3 * r + 1] = oy; out[3 * r + 2] = z;
    w.scan[r] = z;
  }
  syncwarp();
}

// ----------------------------------------------------------------------------------------------
// observation (joystick_pgtt.py:238-370). `rng` is advanced by the five splits of _get_obs.
// ----------------------------------------------------------------------------------------------
DEV void write_obs(WS& w, const EnvBuffers& B, int env, Key& rng, const float* phase, float gait_freq, const float* last_act,
                   const float* command, const int* last_contact, const float* feet_air_time, int lane) {
  // five sequential (rng, key) = split(rng) of _get_obs: the rng chain is computed by every lane, but a
  // lane derives only the noise key of its own slot group (and the height-scan key), so the number of
  // threefry evaluations per lane is 5 + 1 + 1 + 1 + 4 instead of 10 + 4 + 4
  Key chain[5];
#pragma unroll
  for (int s = 0; s < 5; s++) { chain[s] = rng; rng = rng_split(rng, 2, 0); }
  float* o = B.obs_state + (size_t)env * GC.nobs;
  float* pr = B.obs_priv + (size_t)env * GC.npriv;
  const float lvl = GC.noise_level;
  const float* R = w.xmat[0];
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
    if (lane < 3) v = w.sens[lane] + u * lvl * GC.noise_gyro;
    else if (lane < 6) v = -R[6 + idx] + u * lvl * GC.noise_gravity;                       // xmat^T (0,0,-1)
    else if (lane < 18) v = (w.qpos[7 + idx] + u * lvl * GC.noise_joint_pos) - GC.default_pose[idx];
    else if (lane < 30) v = w.qvel[6 + idx] + u * lvl * GC.noise_joint_vel;
    if (lane < 30) { o[lane] = v; pr[lane] = v; }
  }
  const Key scan_key = rng_split(chain[4], 2, 1);   // the linvel key, re-used for the height scan (Q9)
  if (lane < 4 && !base_v) {
    float s, c;
    sincos_(phase[lane], &s, &c);
    o[30 + lane] = c; o[34 + lane] = s; pr[30 + lane] = c; pr[34 + lane] = s;
  }
  float zmin = __int_as_float(0x7f800000);
  for (int r = lane; r < NRAY; r += 32) zmin = fminf(zmin, w.scan[r]);
  zmin = warp_min(zmin);
  for (int r = lane; r < NRAY; r += 32) {
    const float z = (w.scan[r] - zmin) + (2.f * rng_unit(scan_key, NRAY, r) - 1.f) * lvl * GC.noise_heightscan;
    o[o_scan + r] = z; pr[o_scan + r] = z;
  }
  if (lane == 0 && !base_v) { o[155] = gait_freq; pr[155] = gait_freq; }
  if (lane < 12) { o[o_last + lane] = last_act[lane]; pr[o_last + lane] = last_act[lane]; }
  if (lane < 3) {
    o[o_cmd + lane] = command[lane]; pr[o_cmd + lane] = command[lane];
    pr[px + lane] = w.sens[19 + lane]; pr[px + 3 + lane] = w.sens[3 + lane]; pr[px + 6 + lane] = w.sens[16 + lane];
    pr[px + 41 + lane] = 0.f;
  }
  if (lane < 12) { pr[px + 9 + lane] = w.actf[lane]; pr[px + 25 + lane] = w.sens[37 + lane]; }
  if (lane < 4) { pr[px + 21 + lane] = (float)last_contact[lane]; pr[px + 37 + lane] = feet_air_time[lane]; }
}

// history rolls of _get_obs (joystick_pgtt.py:319-334), `step` is info["step"] BEFORE the increment
DEV void update_history(WS& w, const EnvBuffers& B, int env, int step, const float* motor_targets, int lane) {
  const bool upd = (step % GC.history_update_steps == 0) && lane < 12;
  float* qv = B.qvel_hist + (size_t)env * 24;
  float* qe = B.qpos_err_hist + (size_t)env * 24;
  float a = 0.f, b = 0.f;
  if (upd) { a = qv[lane]; b = qe[lane]; }
  syncwarp();
  if (upd) {
    qv[12 + lane] = a; qe[12 + lane] = b;
    qv[lane] = w.qvel[6 + lane];
    qe[lane] = w.qpos[7 + lane] - motor_targets[lane];
  }
}

DEV void store_data(WS& w, const EnvBuffers& B, int env, int lane) {
  if (lane < NQ) B.qpos[env * NQ + lane] = w.qpos[lane];
  if (lane < NV) { B.qvel[env * NV + lane] = w.qvel[lane]; B.warm[env * NV + lane] = w.warm[lane]; B.qacc[env * NV + lane] = w.qacc[lane]; }
  if (lane < NU) { B.ctrl[env * NU + lane] = w.ctrl[lane]; B.actuator_force[env * NU + lane] = w.actf[lane]; }
  for (int i = lane; i < NSENSOR; i += 32) B.sensordata[env * NSENSOR + i] = w.sens[i];
  if (lane < 3) B.site_xpos[env * 15 + lane] = w.sens[10 + lane];
  if (lane < 12) {  // sites FL FR RL RR = leg order
    B.site_xpos[env * 15 + 3 + lane] = w.foot[lane / 3][lane % 3];
  }
  if (lane < 9) B.site_xmat[env * 9 + lane] = w.xmat[0][lane];
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
// Joystick.step + training wrappers, one warp per env
// ----------------------------------------------------------------------------------------------
DEV void env_step(WS& w, const EnvBuffers& B, const float* action_all, int env, int lane, int wrapped) {
  const float dt = GC.ctrl_dt;
  load_model(w, B, env, lane);
  load_state(w, B, env, lane);
  // BraxAutoResetWrapper.step: steps <- 0 where the previous step ended an episode; done cleared
  float steps = 0.f, prev_done_flag = 0.f;
  if (wrapped) { steps = B.steps[env]; prev_done_flag = B.done[env]; if (prev_done_flag != 0.f) steps = 0.f; }
  const float* act = action_all + (size_t)env * NU;
  float action = 0.f, mt = 0.f;
  if (lane < NU) { action = act[lane]; mt = GC.default_pose[lane] + action * GC.action_scale; w.ctrl[lane] = mt; B.motor_targets[env * NU + lane] = mt; }
  syncwarp();
  for (int s = 0; s < GC.n_substeps; s++) {
    const int ni = forward(w, B, env, lane, s == GC.n_substeps - 1);
    if (lane == 0) w.niter[s & 3] = ni;
    euler(w, lane);
  }
  if (lane < 4 && lane < GC.n_substeps) B.solver_niter[env * 4 + lane] = w.niter[lane];
  // compute_contact (base.py:153-171): flag order FR FL RR RL, leg order FL FR RL RR
  int contact = 0, first_contact = 0, last_contact = 0;
  float air = 0.f, swing_peak = 0.f;
  if (lane < 4) {
    const int g = lane ^ 1;
    for (int c = 0; c < NCON; c++) if (w.c_box[c] != -2 && w.c_leg[c] == g && w.c_dist[c] < 0.f) contact = 1;
    last_contact = B.last_contact[env * 4 + lane];
    air = B.feet_air_time[env * 4 + lane];
    first_contact = (air > 0.f) && (contact | last_contact);
    air += dt;
    swing_peak = fmaxf(B.swing_peak[env * 4 + lane], w.sens[25 + 3 * lane + 2]);
  }
  // height scan at the post-step pose, quadrant statistics (joystick_pgtt.py:167-190, Q8)
  heightscan(w, w.qpos, quat_to_yaw(w.qpos + 3), B.heightscan + (size_t)env * NRAY * 3, w.boxlist, lane);
  float hmax = 0.f;
  {
    // quadrant q of this lane group: rows/cols per joystick_pgtt.py:171-174 (n = 6)
    const int q = lane >> 3, sub = lane & 7;
    const int r0 = (q < 2) ? 0 : 7, r1 = (q < 2) ? 6 : 13, c0 = (q & 1) ? 0 : 7, c1 = (q & 1) ? 6 : 9;
    float mx = -__int_as_float(0x7f800000), mn = __int_as_float(0x7f800000);
    const int ncol = c1 - c0, ncell = (r1 - r0) * ncol;
    for (int t = sub; t < ncell; t += 8) {
      const float z = w.scan[(r0 + t / ncol) * NRAY_W + c0 + t % ncol];
      mx = fmaxf(mx, z); mn = fminf(mn, z);
    }
    for (int o = 4; o > 0; o >>= 1) { mx = fmaxf(mx, shfl_xor(mx, o)); mn = fminf(mn, shfl_xor(mn, o)); }
    const float hm = GC.variant ? mx : mx - mn;   // joystick.py:186 vs joystick_pgtt.py:189
    // lane k (< 4) needs quadrant k
    hmax = shfl(hm, (lane & 3) * 8);
    const float hmin = shfl(mn, (lane & 3) * 8);
    if (lane < 4) { B.H_max[env * 4 + lane] = hmax; B.H_min[env * 4 + lane] = hmin; }
  }
  // observation (uses info BEFORE the bookkeeping below, except feet_air_time which is already += dt)
  Key rng; rng.a = B.rng[env * 2]; rng.b = B.rng[env * 2 + 1];
  float phase = 0.f, last_act = 0.f, command = 0.f;
  if (lane < 4) phase = B.phase[env * 4 + lane];
  if (lane < NU) last_act = B.last_act[env * NU + lane];
  if (lane < 3) command = B.command[env * 3 + lane];
  const int step = B.step[env];
  {
    // stage per-lane info values in shared scratch so write_obs can index them
    float* sc = w.mv;  // free between solver calls
    int* sci = w.cand_cnt;
    if (lane < 4) { sc[lane] = phase; sc[4 + lane] = air; sci[lane] = last_contact; }
    syncwarp();
    float* la = w.grad;
    if (lane < NU) la[lane] = last_act;
    if (lane < 3) w.tb[lane] = command;
    syncwarp();
    write_obs(w, B, env, rng, sc, B.gait_freq[env], la, w.tb, sci, sc + 4, lane);
    update_history(w, B, env, step, B.motor_targets + (size_t)env * NU, lane);
  }
  // termination (joystick_pgtt.py:233-236) + a failure guard the reference does not have: a non-finite or
  // absurd generalised state ends the episode, so the auto-reset wrapper restores the env instead of carrying NaNs forever
  bool bad = false;
  if (lane < NQ) bad = !(fabsf(w.qpos[lane]) < 1e6f);
  if (lane < NV) bad |= !(fabsf(w.qvel[lane]) < 1e6f);
  const bool poisoned = any_lane(bad);
  const int done = (w.sens[24] < 0.f) || poisoned;
  // rewards (joystick_pgtt.py:372-599); lanes 0..3 hold per-foot partials, lanes 0..11 per-joint partials
  const float* sdat = w.sens;
  const float cmd0 = w.tb[0], cmd1 = w.tb[1], cmd2 = w.tb[2];
  const float cmd_norm = sqrtf(cmd0 * cmd0 + cmd1 * cmd1 + cmd2 * cmd2);
  float ss = 0.f, pose = 0.f, lim = 0.f, t2 = 0.f, t1 = 0.f, ar = 0.f, en = 0.f;
  if (lane < 12) {
    const float q = w.qpos[7 + lane], dq = q - GC.default_pose[lane], af = w.actf[lane];
    ss = fabsf(dq);
    pose = dq * dq * ((lane % 3 == 0) ? 1.0f : 0.1f);
    const float a = q - GC.soft_lo[lane], b = q - GC.soft_hi[lane];
    lim = -(a < 0.f ? a : 0.f) + (b > 0.f ? b : 0.f);
    t2 = af * af; t1 = fabsf(af);
    ar = (action - last_act) * (action - last_act);
    en = fabsf(w.qvel[6 + lane]) * fabsf(af);
  }
  float slip = 0.f, clr = 0.f, perr = 0.f, swing = 0.f, airr = 0.f, con = 0.f, center = 0.f, fh = 0.f, footz = __int_as_float(0x7f800000);
  if (lane < 4) {
    const float* v = sdat + 37 + 3 * lane; const float* pf = sdat + 25 + 3 * lane;
    const float vxy2 = v[0] * v[0] + v[1] * v[1];
    slip = vxy2 * (float)contact;
    clr = (GC.variant ? fabsf(w.foot[lane ^ 1][2] - (hmax - GC.base_feet_distance + GC.swing_height))   // world-frame foot height, joystick.py:569-572
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
    footz = w.foot[lane ^ 1][2];
  }
  // ordered (lane 0, 1, 2, ...) sums so the result does not depend on a butterfly order
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
    const float bh = w.qpos[2] - footz - 0.27f;
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
  if (lane < 4) B.phase[env * 4 + lane] = fmodf(phase + B.phase_dt[env], 2.f * PGTT_PI);
  int steps_until = B.steps_until[env] - 1;
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
    B.time[env] += GC.dt * (float)GC.n_substeps;
  }
  store_data(w, B, env, lane);
  float done_out = (float)done;
  if (wrapped) {
    // EpisodeWrapper.step
    steps += 1.f;
    const float done_inner = done_out;
    const bool over = steps >= (float)GC.episode_length;
    done_out = over ? 1.f : done_inner;
    const float prev_done = B.episode_done[env];
    float* em = B.episode_metrics + (size_t)env * 24;
    if (lane == 0) {
      B.truncation[env] = over ? 1.f - done_inner : 0.f;
      B.steps[env] = steps;
      em[0] = (em[0] + reward) * (1.f - prev_done);
      em[1] = (em[1] + 1.f) * (1.f - prev_done);
    }
    if (lane < NMETRIC) em[2 + lane] = (em[2 + lane] + metric) * (1.f - prev_done);
    syncwarp();
    if (lane == 0) B.episode_done[env] = done_out;
    // auto-reset: restore the cached first data / obs only (info is NOT reset)
    if (done_out != 0.f) {
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
      for (int i = lane; i < GC.nobs; i += 32) B.obs_state[(size_t)env * GC.nobs + i] = B.first_obs_state[(size_t)env * GC.nobs + i];
      for (int i = lane; i < GC.npriv; i += 32) B.obs_priv[(size_t)env * GC.npriv + i] = B.first_obs_priv[(size_t)env * GC.npriv + i];
      if (lane == 0) B.time[env] = 0.f;
    }
  }
  if (lane == 0) { B.reward[env] = reward; B.done[env] = done_out; }
}

// ----------------------------------------------------------------------------------------------
// Joystick.reset (joystick_pgtt.py:50-131) + wrapper resets
// ----------------------------------------------------------------------------------------------
DEV void env_reset(WS& w, const EnvBuffers& B, const uint32_t* keys, int env, int lane) {
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
  forward(w, B, env, lane, false);                       // mjx_env.init
  float* hs = B.heightscan + (size_t)env * NRAY * 3;
  heightscan(w, w.qpos, 0.f, hs, w.boxlist, lane);     // yaw = 0 (Q10)
  float zmax = -__int_as_float(0x7f800000);
  for (int r = lane; r < NRAY; r += 32) zmax = fmaxf(zmax, w.scan[r]);
  zmax = warp_max(zmax);
  if (lane == 0) w.qpos[2] += zmax;
  syncwarp();
  forward(w, B, env, lane, true);                        // mjx.forward at the lifted pose
  const Key key1 = rng_split(rng, 3, 1), key2 = rng_split(rng, 3, 2);
  rng = rng_split(rng, 3, 0);
  const float tcmd = -log1pf(-rng_unit(key1, 1, 0)) * 5.0f;
  const int steps_until = (int)rintf(tcmd / GC.ctrl_dt);
  float command = 0.f;
  if (lane < 3) command = rng_uniform(key2, 3, lane, GC.cmd_u_min[lane], GC.cmd_u_max[lane]);
  key = rng_split(rng, 2, 1); rng = rng_split(rng, 2, 0);
  const float gait_freq = rng_uniform(key, 1, 0, GC.gait_freq[0], GC.gait_freq[1]);
  heightscan(w, w.qpos, 0.f, hs, w.boxlist, lane);
  // info
  if (lane < 4) {
    const float ph = (lane == 1 || lane == 2) ? PGTT_PI : 0.f;
    B.phase[env * 4 + lane] = ph;
    B.feet_air_time[env * 4 + lane] = 0.f; B.last_contact[env * 4 + lane] = 0; B.swing_peak[env * 4 + lane] = 0.f;
    B.H_max[env * 4 + lane] = 0.1f; B.H_min[env * 4 + lane] = 0.f;
    B.contact[env * 4 + lane] = 0; B.first_contact[env * 4 + lane] = 0;
    w.mv[lane] = ph; w.mv[4 + lane] = 0.f; w.cand_cnt[lane] = 0;
  }
  if (lane < NU) { B.last_act[env * NU + lane] = 0.f; B.last_last_act[env * NU + lane] = 0.f; B.motor_targets[env * NU + lane] = 0.f; w.grad[lane] = 0.f; }
  if (lane < 24) { B.qpos_err_hist[env * 24 + lane] = 0.f; B.qvel_hist[env * 24 + lane] = 0.f; }
  if (lane < 3) { B.command[env * 3 + lane] = command; w.tb[lane] = command; }
  if (lane < NMETRIC) B.metrics[env * NMETRIC + lane] = 0.f;
  if (lane < 24) B.episode_metrics[env * 24 + lane] = 0.f;
  syncwarp();
  write_obs(w, B, env, rng, w.mv, gait_freq, w.grad, w.tb, w.cand_cnt, w.mv + 4, lane);
  update_history(w, B, env, 0, B.motor_targets + (size_t)env * NU, lane);
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
  if (lane < 9) B.first_site_xmat[env * 9 + lane] = w.xmat[0][lane];
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
  forward(w, B, env, lane, true);
  store_data(w, B, env, lane);
}

DEV void env_scan(WS& w, const EnvBuffers& B, const float* center, const float* yaw, float* out, int env, int lane) {
  load_model(w, B, env, lane);
  if (lane < 3) w.tb[lane] = center[env * 3 + lane];
  syncwarp();
  heightscan(w, w.tb, yaw[env], out + (size_t)env * NRAY * 3, w.boxlist, lane);
}
