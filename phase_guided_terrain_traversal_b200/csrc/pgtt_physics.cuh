// pgtt_physics.cuh - one-warp-per-env rigid-body step for the GO2 (device code, sm_100a).
//
// Stands in for mjx.forward / mjx.step as the reference calls them (go2/joystick_pgtt.py:72,78,
// 146-148); algorithm per SURVEY.md Appendix A, checked stage by stage against oracle/pgtt_oracle.c.
// Design (B200-first, not a port of MJX's dense XLA program):
//   * one warp owns one env; all per-env working data lives in a ~13 KB shared-memory workspace,
//     HBM is touched once per control step (load state, store state/obs);
//   * the 18x18 inertia / Newton Hessian are kept as ARROW matrices (6x6 base block, four 6x3
//     couplings, four 3x3 leg blocks): contacts and joint limits preserve that sparsity, so the
//     factorisation is four independent 3x3 Choleskys + one 6x6 Schur complement;
//   * lanes are grouped 8 per leg for the kinematic / RNE chains, one lane per constraint row
//     (8 contacts x 4 pyramid edges = 32 rows + 12 limit rows) in the solver.
// Every collective below is issued with the full warp from warp-uniform control flow.
#pragma once
#include "pgtt_types.h"
#include "simt.h"

#ifdef PGTT_HOST_EMU
extern ModelConst g_mc;
#else
__constant__ ModelConst g_mc;
#endif
#define GC g_mc

// stage barrier ids (bit positions of ModelConst::sync_mask)
enum { ST_KIN = 0, ST_COLLIDE, ST_COM, ST_CRB, ST_RNE, ST_SMOOTH, ST_ROWS, ST_PRESOLVE, ST_POSTSOLVE, ST_SENSORS, ST_EULER };
// CTA barrier after stage `id` if enabled (see simt.h:cta_bar), else just the warp-level sync the stage needs
DEV void stage_sync(const WS& w, int id) {
  if ((GC.sync_mask >> id) & 1) cta_bar(w.bar_threads); else syncwarp();
}

// stage-timestamp trace (development aid): lane 0 of the warp appends the SM clock after a stage when tracing is on
#ifdef PGTT_HOST_EMU
#define STAGE_TRACE(w, lane) do { } while (0)
#else
#define STAGE_TRACE(w, lane) do { if ((w).trace && (lane) == 0) (w).trace[(w).tix++] = clock64(); } while (0)
#endif

#define PGTT_MINVAL 1e-15f
#define PGTT_PI 3.14159265358979323846f

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
DEV float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
DEV float dot6(const float* a, const float* b) { return dot3(a, b) + dot3(a + 3, b + 3); }
DEV void cross3(float* r, const float* a, const float* b) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
// res = I * v for the 10-number spatial inertia (Ixx Iyy Izz Ixy Ixz Iyz | m*c | m) about the robot COM
DEV void inert_mul(float* res, const float* I, const float* v) {
  float hxl[3], hxw[3];
  cross3(hxl, I + 6, v + 3);
  cross3(hxw, I + 6, v);
  res[0] = I[0] * v[0] + I[3] * v[1] + I[4] * v[2] + hxl[0];
  res[1] = I[3] * v[0] + I[1] * v[1] + I[5] * v[2] + hxl[1];
  res[2] = I[4] * v[0] + I[5] * v[1] + I[2] * v[2] + hxl[2];
  res[3] = I[9] * v[3] - hxw[0]; res[4] = I[9] * v[4] - hxw[1]; res[5] = I[9] * v[5] - hxw[2];
}
DEV void cross_motion(float* res, const float* vel, const float* v) {
  float a[3], b[3], c[3];
  cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
  res[0] = a[0]; res[1] = a[1]; res[2] = a[2];
  res[3] = b[0] + c[0]; res[4] = b[1] + c[1]; res[5] = b[2] + c[2];
}
DEV void cross_force(float* res, const float* vel, const float* f) {
  float a[3], b[3], c[3];
  cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
  res[0] = a[0] + b[0]; res[1] = a[1] + b[1]; res[2] = a[2] + b[2];
  res[3] = c[0]; res[4] = c[1]; res[5] = c[2];
}
// dof index of column `col` (0..8) of a contact on leg `leg`: 6 base dofs then the leg's 3 hinges
DEV int col_dof(int col, int leg) { return col < 6 ? col : 3 * leg + col; }

// ----------------------------------------------------------------------------------------------
// load / store of the per-env state
// ----------------------------------------------------------------------------------------------
DEV void load_model(WS& w, const EnvBuffers& B, int env, int lane) {
  if (lane < NB) w.mass[lane] = B.m_mass[env * NB + lane];
  if (lane < 3) w.ipos0[lane] = B.m_ipos[env * 3 + lane];
  if (lane < 12) {
    w.armature[lane] = B.m_armature[env * 12 + lane];
    w.damping[lane] = B.m_damping[env * 12 + lane];
    w.gain[lane] = B.m_gain[env * 12 + lane];
    w.bias1[lane] = B.m_bias1[env * 12 + lane];
    w.qpos0[lane] = B.m_qpos0[env * 12 + lane];
  }
  float m = lane < NB ? B.m_mass[env * NB + lane] : 0.f;
  m = warp_sum(m);
  if (lane == 0) { w.mtot_inv = 1.0f / m; w.floor_mu = B.m_floorfric[env]; w.near_ok = 0; }
  syncwarp();
}

DEV void load_state(WS& w, const EnvBuffers& B, int env, int lane) {
  if (lane < NQ) w.qpos[lane] = B.qpos[env * NQ + lane];
  if (lane < NV) { w.qvel[lane] = B.qvel[env * NV + lane]; w.warm[lane] = B.warm[env * NV + lane]; }
  syncwarp();
}

// ----------------------------------------------------------------------------------------------
// position stage: kinematics, COM-frame inertias, motion axes, CRB, inertia matrix (App. A1-A3)
// ----------------------------------------------------------------------------------------------
DEV void kinematics(WS& w, int lane) {
  const int g = lane >> 3, sub = lane & 7;
  float qw = w.qpos[3], qx = w.qpos[4], qy = w.qpos[5], qz = w.qpos[6];
  const float qn = 1.0f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  qw *= qn; qx *= qn; qy *= qn; qz *= qn;
  float R[9];
  R[0] = qw * qw + qx * qx - qy * qy - qz * qz; R[1] = 2 * (qx * qy - qw * qz); R[2] = 2 * (qx * qz + qw * qy);
  R[3] = 2 * (qx * qy + qw * qz); R[4] = qw * qw - qx * qx + qy * qy - qz * qz; R[5] = 2 * (qy * qz - qw * qx);
  R[6] = 2 * (qx * qz - qw * qy); R[7] = 2 * (qy * qz + qw * qx); R[8] = qw * qw - qx * qx - qy * qy + qz * qz;
  float p[3] = {w.qpos[0], w.qpos[1], w.qpos[2]};
  if (lane == 1) {
    for (int i = 0; i < 3; i++) { w.xpos[0][i] = p[i]; w.xpos0[i] = p[i]; }
    for (int i = 0; i < 9; i++) { w.xmat[0][i] = R[i]; w.xmat0[i] = R[i]; }
  }
  const int bh = 1 + 3 * g;
  float s, c;
  // hip: rotation about the parent's x axis
  {
    const float* o = GC.body_pos[bh];
    for (int i = 0; i < 3; i++) p[i] += R[3 * i] * o[0] + R[3 * i + 1] * o[1] + R[3 * i + 2] * o[2];
    sincos_(w.qpos[7 + 3 * g] - w.qpos0[3 * g], &s, &c);
    for (int i = 0; i < 3; i++) {
      const float c1 = R[3 * i + 1], c2 = R[3 * i + 2];
      R[3 * i + 1] = c1 * c + c2 * s;
      R[3 * i + 2] = -c1 * s + c2 * c;
    }
    if (sub == 0) {
      for (int i = 0; i < 3; i++) w.xpos[bh][i] = p[i];
      for (int i = 0; i < 9; i++) w.xmat[bh][i] = R[i];
    }
  }
  // thigh, calf: rotation about the local y axis
  for (int t = 1; t < 3; t++) {
    const float* o = GC.body_pos[bh + t];
    for (int i = 0; i < 3; i++) p[i] += R[3 * i] * o[0] + R[3 * i + 1] * o[1] + R[3 * i + 2] * o[2];
    sincos_(w.qpos[7 + 3 * g + t] - w.qpos0[3 * g + t], &s, &c);
    for (int i = 0; i < 3; i++) {
      const float c0 = R[3 * i], c2 = R[3 * i + 2];
      R[3 * i] = c0 * c - c2 * s;
      R[3 * i + 2] = c0 * s + c2 * c;
    }
    if (sub == t) {
      for (int i = 0; i < 3; i++) w.xpos[bh + t][i] = p[i];
      for (int i = 0; i < 9; i++) w.xmat[bh + t][i] = R[i];
    }
  }
  if (sub == 3) {
    const float* o = GC.foot_pos;
    for (int i = 0; i < 3; i++) w.foot[g][i] = p[i] + R[3 * i] * o[0] + R[3 * i + 1] * o[1] + R[3 * i + 2] * o[2];
  }
  stage_sync(w, ST_KIN);
}

DEV void com_inertia_cdof(WS& w, int lane) {
  float m = 0.f, xi[3] = {0.f, 0.f, 0.f};
  if (lane < NB) {
    const int b = lane;
    m = w.mass[b];
    const float* ip = (b == 0) ? w.ipos0 : GC.body_ipos[b];
    const float* R = w.xmat[b];
    for (int i = 0; i < 3; i++) xi[i] = w.xpos[b][i] + R[3 * i] * ip[0] + R[3 * i + 1] * ip[1] + R[3 * i + 2] * ip[2];
    for (int i = 0; i < 3; i++) w.xipos[b][i] = xi[i];
  }
  float com[3];
  for (int i = 0; i < 3; i++) com[i] = warp_sum(m * xi[i]) * w.mtot_inv;
  if (lane == 0) for (int i = 0; i < 3; i++) w.com[i] = com[i];
  if (lane < NB) {
    const int b = lane;
    const float* R = w.xmat[b];
    const float* I = GC.body_I[b];  // body-frame tensor xx yy zz xy xz yz
    float T[9];                     // T = R * I_b
    for (int i = 0; i < 3; i++) {
      const float r0 = R[3 * i], r1 = R[3 * i + 1], r2 = R[3 * i + 2];
      T[3 * i] = r0 * I[0] + r1 * I[3] + r2 * I[4];
      T[3 * i + 1] = r0 * I[3] + r1 * I[1] + r2 * I[5];
      T[3 * i + 2] = r0 * I[4] + r1 * I[5] + r2 * I[2];
    }
    float o[3] = {xi[0] - com[0], xi[1] - com[1], xi[2] - com[2]};
    const float oo = dot3(o, o);
    float* ci = w.cinert[b];
    ci[0] = dot3(T, R) + m * (oo - o[0] * o[0]);
    ci[1] = dot3(T + 3, R + 3) + m * (oo - o[1] * o[1]);
    ci[2] = dot3(T + 6, R + 6) + m * (oo - o[2] * o[2]);
    ci[3] = dot3(T, R + 3) - m * o[0] * o[1];
    ci[4] = dot3(T, R + 6) - m * o[0] * o[2];
    ci[5] = dot3(T + 3, R + 6) - m * o[1] * o[2];
    ci[6] = m * o[0]; ci[7] = m * o[1]; ci[8] = m * o[2]; ci[9] = m;
  }
  if (lane < NV) {
    const int d = lane;
    float* cd = w.cdof[d];
    if (d < 3) {
      for (int i = 0; i < 6; i++) cd[i] = 0.f;
      cd[3 + d] = 1.f;
    } else {
      int b, colidx;
      if (d < 6) { b = 0; colidx = d - 3; } else { b = d - 5; colidx = ((d - 6) % 3 == 0) ? 0 : 1; }
      const float* R = w.xmat[b];
      float ax[3] = {R[colidx], R[3 + colidx], R[6 + colidx]};
      float off[3] = {com[0] - w.xpos[b][0], com[1] - w.xpos[b][1], com[2] - w.xpos[b][2]};
      cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
      cross3(cd + 3, ax, off);
    }
  }
  stage_sync(w, ST_COM);
}

DEV void crb_and_inertia(WS& w, int lane) {
  const int g = lane >> 3, sub = lane & 7;
  for (int k = sub; k < 10; k += 8) {
    const float c3 = w.cinert[3 + 3 * g][k];
    const float c2 = w.cinert[2 + 3 * g][k] + c3;
    const float c1 = w.cinert[1 + 3 * g][k] + c2;
    w.crb[3 + 3 * g][k] = c3; w.crb[2 + 3 * g][k] = c2; w.crb[1 + 3 * g][k] = c1;
  }
  syncwarp();
  if (lane < 10) w.crb[0][lane] = w.cinert[0][lane] + w.crb[1][lane] + w.crb[4][lane] + w.crb[7][lane] + w.crb[10][lane];
  syncwarp();
  if (lane < NV) {
    const int b = lane < 6 ? 0 : lane - 5;
    float cd[6], f[6];
    for (int i = 0; i < 6; i++) cd[i] = w.cdof[lane][i];
    inert_mul(f, w.crb[b], cd);
    for (int i = 0; i < 6; i++) w.F[lane][i] = f[i];
  }
  syncwarp();
  for (int e = lane; e < 144; e += 32) {
    if (e < 36) {
      w.MB[e] = dot6(w.cdof[e / 6], w.F[e % 6]);
    } else if (e < 108) {
      const int e2 = e - 36, gg = e2 / 18, a = (e2 % 18) / 3, j = e2 % 3;
      w.MC[e2] = dot6(w.cdof[a], w.F[6 + 3 * gg + j]);
    } else {
      const int e3 = e - 108, gg = e3 / 9, j = (e3 % 9) / 3, k = e3 % 3;
      const int hi = j > k ? j : k, lo = j > k ? k : j;
      float v = dot6(w.cdof[6 + 3 * gg + lo], w.F[6 + 3 * gg + hi]);
      if (j == k) v += w.armature[3 * gg + j];
      w.MA[e3] = v;
    }
  }
  stage_sync(w, ST_CRB);
}

// y = M x with the arrow blocks (x, y in shared memory; caller syncs)
DEV void arrow_mul(const float* Bm, const float* Cm, const float* Am, const float* x, float* y, int lane) {
  if (lane < 6) {
    float s = 0.f;
    for (int b = 0; b < 6; b++) s += Bm[lane * 6 + b] * x[b];
    for (int gg = 0; gg < 4; gg++)
      for (int j = 0; j < 3; j++) s += Cm[gg * 18 + lane * 3 + j] * x[6 + 3 * gg + j];
    y[lane] = s;
  } else if (lane < NV) {
    const int gg = (lane - 6) / 3, j = (lane - 6) % 3;
    float s = 0.f;
    for (int a = 0; a < 6; a++) s += Cm[gg * 18 + a * 3 + j] * x[a];
    for (int k = 0; k < 3; k++) s += Am[gg * 9 + j * 3 + k] * x[6 + 3 * gg + k];
    y[lane] = s;
  }
}

// ----------------------------------------------------------------------------------------------
// arrow factorisation / solve
// ----------------------------------------------------------------------------------------------
struct ArrowFac {
  float L[21];  // Cholesky of the 6x6 Schur complement, row-packed lower, diagonals as reciprocals
  float la[6];  // Cholesky of this lane's leg block: i00 l10 i11 l20 l21 i22 (reciprocal diagonals)
};

// Factor lives in shared memory (w.fLA, w.fY, w.fL) so it can be re-used by later solves.
DEV void arrow_factor(WS& w, const float* Bm, const float* Cm, const float* Am, int lane) {
  ArrowFac F;
  const int g = lane >> 3, sub = lane & 7;
  const float* Ag = Am + 9 * g;
  const float i00 = rsqrt_(fmaxf(Ag[0], PGTT_MINVAL));
  const float l10 = Ag[3] * i00, l20 = Ag[6] * i00;
  const float i11 = rsqrt_(fmaxf(Ag[4] - l10 * l10, PGTT_MINVAL));
  const float l21 = (Ag[7] - l20 * l10) * i11;
  const float i22 = rsqrt_(fmaxf(Ag[8] - l20 * l20 - l21 * l21, PGTT_MINVAL));
  F.la[0] = i00; F.la[1] = l10; F.la[2] = i11; F.la[3] = l20; F.la[4] = l21; F.la[5] = i22;
  float pr[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (sub < 6) {
    const float* cr = Cm + g * 18 + sub * 3;
    const float z0 = cr[0] * i00, z1 = (cr[1] - l10 * z0) * i11, z2 = (cr[2] - l20 * z0 - l21 * z1) * i22;
    const float y2 = z2 * i22, y1 = (z1 - l21 * y2) * i11, y0 = (z0 - l10 * y1 - l20 * y2) * i00;
    w.fY[g][sub][0] = y0; w.fY[g][sub][1] = y1; w.fY[g][sub][2] = y2;
    for (int b = 0; b < 6; b++) {
      const float* cb = Cm + g * 18 + b * 3;
      pr[b] = y0 * cb[0] + y1 * cb[1] + y2 * cb[2];
    }
  }
  for (int b = 0; b < 6; b++) { pr[b] += shfl_xor(pr[b], 8); pr[b] += shfl_xor(pr[b], 16); }
  if (g == 0 && sub < 6) for (int b = 0; b < 6; b++) w.fS[sub * 6 + b] = Bm[sub * 6 + b] - pr[b];
  syncwarp();
  // every lane factors the 6x6 redundantly in registers (no communication in the solves)
  int idx = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int j = 0; j <= i; j++) {
      float s = w.fS[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= F.L[i * (i + 1) / 2 + k] * F.L[j * (j + 1) / 2 + k];
      F.L[idx++] = (i == j) ? rsqrt_(fmaxf(s, PGTT_MINVAL)) : s * F.L[j * (j + 1) / 2 + j];
    }
  }
  if (lane < 21) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 21; k++) v = (lane == k) ? F.L[k] : v;
    w.fL[lane] = v;
  }
  if (sub == 6) {
#pragma unroll
    for (int k = 0; k < 6; k++) w.fLA[g][k] = F.la[k];
  }
  syncwarp();
}

// x = H^-1 r  (r, x in shared memory, may alias). Ends with a syncwarp.
DEV void arrow_solve(WS& w, const float* r, float* x, int lane) {
  const int g = lane >> 3, sub = lane & 7;
  ArrowFac F;
#pragma unroll
  for (int k = 0; k < 21; k++) F.L[k] = w.fL[k];
#pragma unroll
  for (int k = 0; k < 6; k++) F.la[k] = w.fLA[g][k];
  const float r0 = r[6 + 3 * g], r1 = r[7 + 3 * g], r2 = r[8 + 3 * g];
  float t = 0.f;
  if (sub < 6) t = w.fY[g][sub][0] * r0 + w.fY[g][sub][1] * r1 + w.fY[g][sub][2] * r2;
  t += shfl_xor(t, 8);
  t += shfl_xor(t, 16);
  const float rb = (lane < 6) ? r[lane] : 0.f;
  syncwarp();  // all reads of r done before x (possibly aliasing r) is written
  if (lane < 6) w.tb[lane] = rb - t;
  syncwarp();
  float xb[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float s = w.tb[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= F.L[i * (i + 1) / 2 + k] * xb[k];
    xb[i] = s * F.L[i * (i + 1) / 2 + i];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    float s = xb[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s -= F.L[k * (k + 1) / 2 + i] * xb[k];
    xb[i] = s * F.L[i * (i + 1) / 2 + i];
  }
  // leg part: A_g^-1 r_g - Y_g^T x_b
  const float z0 = r0 * F.la[0], z1 = (r1 - F.la[1] * z0) * F.la[2], z2 = (r2 - F.la[3] * z0 - F.la[4] * z1) * F.la[5];
  const float y2 = z2 * F.la[5], y1 = (z1 - F.la[4] * y2) * F.la[2], y0 = (z0 - F.la[1] * y1 - F.la[3] * y2) * F.la[0];
  if (sub < 3) {
    float v = sub == 0 ? y0 : (sub == 1 ? y1 : y2);
    for (int a = 0; a < 6; a++) v -= w.fY[g][a][sub] * xb[a];
    x[6 + 3 * g + sub] = v;
  }
  if (lane < 6) x[lane] = xb[lane];
  syncwarp();
}

// ----------------------------------------------------------------------------------------------
// velocity stage + RNE bias forces (App. A6); per-leg chains run redundantly on the 8 lanes of a group
// ----------------------------------------------------------------------------------------------
DEV void velocity_rne(WS& w, int lane) {
  const int g = lane >> 3, sub = lane & 7;
  float vb[6] = {0.f, 0.f, 0.f, w.qvel[0], w.qvel[1], w.qvel[2]};
  float ab[6] = {0.f, 0.f, 0.f, 0.f, 0.f, -GC.gravity_z};
  float cdd[3][3];
  for (int k = 0; k < 3; k++) {  // cdof_dot of the rotational free dofs = [0; v_lin x axis]
    cross3(cdd[k], vb + 3, w.cdof[3 + k]);
    const float qd = w.qvel[3 + k];
    ab[3] += cdd[k][0] * qd; ab[4] += cdd[k][1] * qd; ab[5] += cdd[k][2] * qd;
  }
  for (int k = 0; k < 3; k++) {
    const float qd = w.qvel[3 + k];
    for (int i = 0; i < 6; i++) vb[i] += w.cdof[3 + k][i] * qd;
  }
  if (lane < 3) {
    for (int i = 0; i < 6; i++) w.cdofd_base[lane][i] = 0.f;
    w.cdofd_base[3 + lane][0] = 0.f; w.cdofd_base[3 + lane][1] = 0.f; w.cdofd_base[3 + lane][2] = 0.f;
    for (int i = 0; i < 3; i++) w.cdofd_base[3 + lane][3 + i] = cdd[lane][i];
  }
  if (lane == 3) for (int i = 0; i < 6; i++) w.cvel_base[i] = vb[i];
  float fb[6], tmp[6], tmp2[6];
  inert_mul(fb, w.cinert[0], ab);
  inert_mul(tmp, w.cinert[0], vb);
  cross_force(tmp2, vb, tmp);
  for (int i = 0; i < 6; i++) fb[i] += tmp2[i];
  float vp[6], ap[6], fl[3][6];
  for (int i = 0; i < 6; i++) { vp[i] = vb[i]; ap[i] = ab[i]; }
  for (int t = 0; t < 3; t++) {
    const int d = 6 + 3 * g + t, b = 1 + 3 * g + t;
    float cd[6], cdot[6];
    for (int i = 0; i < 6; i++) cd[i] = w.cdof[d][i];
    cross_motion(cdot, vp, cd);
    const float qd = w.qvel[d];
    for (int i = 0; i < 6; i++) { vp[i] += cd[i] * qd; ap[i] += cdot[i] * qd; }
    inert_mul(fl[t], w.cinert[b], ap);
    inert_mul(tmp, w.cinert[b], vp);
    cross_force(tmp2, vp, tmp);
    for (int i = 0; i < 6; i++) fl[t][i] += tmp2[i];
    if (t == 2 && sub == 2) {   // the sensors read the calf velocity (foot linear velocity)
      for (int i = 0; i < 6; i++) w.cvel_calf[g][i] = vp[i];
    }
  }
  for (int i = 0; i < 6; i++) { fl[1][i] += fl[2][i]; fl[0][i] += fl[1][i]; }
  if (sub < 3) w.bias[6 + 3 * g + sub] = dot6(w.cdof[6 + 3 * g + sub], fl[sub]);
  float tot[6];
  for (int i = 0; i < 6; i++) {
    float v = fl[0][i];
    v += shfl_xor(v, 8);
    v += shfl_xor(v, 16);
    tot[i] = v + fb[i];
  }
  if (lane < 6) w.bias[lane] = dot6(w.cdof[lane], tot);
  stage_sync(w, ST_RNE);
}

// passive + actuator + bias -> qfrc_smooth; leaves actuator_force in w.actf
DEV void smooth_forces(WS& w, int lane) {
  if (lane < NV) {
    const int d = lane;
    float f = -w.bias[d];
    if (d >= 6) {
      const int j = d - 6, a = GC.act_of_hinge[j];
      float c = fminf(fmaxf(w.ctrl[a], GC.ctrl_lo[a]), GC.ctrl_hi[a]);
      float af = w.gain[a] * c + GC.act_bias0[a] + w.bias1[a] * w.qpos[7 + j] + GC.act_bias2[a] * w.qvel[d];
      af = fminf(fmaxf(af, GC.frc_lo[a]), GC.frc_hi[a]);
      w.actf[a] = af;
      f = (-w.damping[j] * w.qvel[d] - w.bias[d]) + af;
    }
    w.qs[d] = f;
  }
  stage_sync(w, ST_SMOOTH);
}

// ----------------------------------------------------------------------------------------------
// collision (App. A4)
// ----------------------------------------------------------------------------------------------
DEV void make_frame(float* fr, const float* n) {
  float a[3] = {n[0], n[1], n[2]};
  float nn = sqrtf(dot3(a, a));
  if (nn < PGTT_MINVAL) { a[0] = a[1] = a[2] = 0.f; } else { a[0] /= nn; a[1] /= nn; a[2] /= nn; }
  float b[3] = {0.f, 0.f, 0.f};
  if (-0.5f < a[1] && a[1] < 0.5f) b[1] = 1.f; else b[2] = 1.f;
  const float ab = dot3(a, b);
  for (int i = 0; i < 3; i++) b[i] -= a[i] * ab;
  nn = sqrtf(dot3(b, b));
  if (nn < PGTT_MINVAL) { b[0] = b[1] = b[2] = 0.f; } else { b[0] /= nn; b[1] /= nn; b[2] /= nn; }
  float c[3];
  cross3(c, a, b);
  for (int i = 0; i < 3; i++) { fr[i] = a[i]; fr[3 + i] = b[i]; fr[6 + i] = c[i]; }
}

// sphere (centre p, radius r) against a yaw-rotated box: returns dist; local closest-point data
DEV float sphere_box_local(const float* bx, const float* p, float r, float* l, float* pt) {
  const float rx = p[0] - bx[0], ry = p[1] - bx[1], rz = p[2] - bx[2];
  l[0] = bx[6] * rx + bx[7] * ry;
  l[1] = -bx[7] * rx + bx[6] * ry;
  l[2] = rz;
  const float d0 = fabsf(l[0]) - bx[3], d1 = fabsf(l[1]) - bx[4], d2 = fabsf(l[2]) - bx[5];
  if (d0 <= 0.f && d1 <= 0.f && d2 <= 0.f) {
    // centre inside the box: least-penetrated face (first of +x,-x,+y,-y,+z,-z on ties)
    int ax = 0; float dm = d0;
    if (d1 > dm) { ax = 1; dm = d1; }
    if (d2 > dm) { ax = 2; dm = d2; }
    pt[0] = l[0]; pt[1] = l[1]; pt[2] = l[2];
    pt[ax] = (l[ax] >= 0.f ? 1.f : -1.f) * bx[3 + ax];
    return -dm - r;
  }
  pt[0] = fminf(fmaxf(l[0], -bx[3]), bx[3]);
  pt[1] = fminf(fmaxf(l[1], -bx[4]), bx[4]);
  pt[2] = fminf(fmaxf(l[2], -bx[5]), bx[5]);
  const float e0 = pt[0] - l[0], e1 = pt[1] - l[1], e2 = pt[2] - l[2];
  return sqrtf(e0 * e0 + e1 * e1 + e2 * e2) - r;
}

// squared centre distance, the broad-phase key (same order as mjx's |d| - (r + rbound): rbound is one constant).
// Un-contracted so that the list pass and the full scan (and every kernel generation) get the same bits.
DEV float sqdist3(float dx, float dy, float dz) { return mul_add_nofma(dz, dz, mul_add_nofma(dy, dy, dx * dx)); }

// Per control step and foot: the boxes whose SURFACE can come within the foot radius (penetration tests) and the boxes
// whose CENTRE can come within W_RCEN (broad-phase ranks) while the foot travels at most W_MARGIN from where the lists
// were built. Both are conservative supersets, so scanning them gives bit-identical contact bookkeeping to scanning all
// 100 boxes; travel beyond the margin, list overflow or a threshold above W_RCEN^2 falls back to the full scan.
// The env's box table (3.2 KB, global memory) is read once per control step here; the per-substep passes only touch
// the listed boxes (L1-resident).
DEV void build_near(WS& w, const float4* bp, int lane) {
  const int nb = GC.n_boxes;
  const float rp = (GC.foot_r + W_MARGIN) * 1.001f, rc = (W_RCEN + W_MARGIN) * 1.001f;
  const float rp2 = rp * rp, rc2 = rc * rc;
  const unsigned lt = (1u << lane) - 1u;
  float ft[4][3];
#pragma unroll
  for (int f = 0; f < 4; f++) { ft[f][0] = w.foot[f][0]; ft[f][1] = w.foot[f][1]; ft[f][2] = w.foot[f][2]; }
  int npen[4] = {0, 0, 0, 0}, ncen[4] = {0, 0, 0, 0};
  float4 c0[4], c1[4];
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int k = it * 32 + lane, kc = k < nb ? k : nb - 1;
    c0[it] = ldg4(bp + 2 * kc); c1[it] = ldg4(bp + 2 * kc + 1);
  }
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int k = it * 32 + lane;
    const bool valid = k < nb;
    const float4 b0 = c0[it], b1 = c1[it];
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const float dx = b0.x - ft[f][0], dy = b0.y - ft[f][1], dz = b0.z - ft[f][2];
      const float e0 = fmaxf(fabsf(b1.z * dx + b1.w * dy) - b0.w, 0.f), e1 = fmaxf(fabsf(b1.z * dy - b1.w * dx) - b1.x, 0.f);
      const float e2 = fmaxf(fabsf(dz) - b1.y, 0.f);
      const bool pen = valid && (e0 * e0 + e1 * e1 + e2 * e2 < rp2), cen = valid && (dx * dx + dy * dy + dz * dz < rc2);
      const unsigned mp = wballot(pen), mc = wballot(cen);
      if (pen) { const int i = npen[f] + popc(mp & lt); if (i < W_QPEN) w.pen_list[f][i] = k; }
      if (cen) { const int i = ncen[f] + popc(mc & lt); if (i < W_QCEN) w.cen_list[f][i] = k; }
      npen[f] += popc(mp); ncen[f] += popc(mc);
    }
  }
  bool over = false;
#pragma unroll
  for (int f = 0; f < 4; f++) over |= (npen[f] > W_QPEN) || (ncen[f] > W_QCEN);
  if (lane < 4) {
#pragma unroll
    for (int f = 0; f < 4; f++) if (lane == f) { w.near_npen[f] = npen[f]; w.near_ncen[f] = ncen[f]; }
    w.near_f0[lane][0] = w.foot[lane][0]; w.near_f0[lane][1] = w.foot[lane][1]; w.near_f0[lane][2] = w.foot[lane][2];
  }
  if (lane == 0) w.near_ok = over ? 0 : 1;
  syncwarp();
}

DEV void collide_boxes(WS& w, const EnvBuffers& B, int env, int lane, bool build) {
  const float r = GC.foot_r;
  const int nb = GC.n_boxes;
  const float4* bp = reinterpret_cast<const float4*>(B.terrain + (size_t)B.terrain_index[env] * NBOX * BOXF);
  if (build) build_near(w, bp, lane);
  // lists usable: built without overflow and no foot further than the margin from where they were built (warp-uniform)
  bool lists_ok = w.near_ok != 0;
#pragma unroll
  for (int f = 0; f < 4; f++) {
    const float tx = w.foot[f][0] - w.near_f0[f][0], ty = w.foot[f][1] - w.near_f0[f][1], tz = w.foot[f][2] - w.near_f0[f][2];
    lists_ok = lists_ok && (tx * tx + ty * ty + tz * tz <= W_MARGIN * W_MARGIN);
  }
  const bool full = GC.quad_fullscan || !all_lanes(lists_ok);
  int ncand = 0;
  const unsigned lt = (1u << lane) - 1u;
  const float r2 = r * r * 1.0001f;  // conservative pre-filter; the exact test runs only where it passes
  if (!full) {
    // narrow phase over the penetration lists: 8 lanes per foot, one listed box per lane
    const int f = lane >> 3, j = lane & 7;
    const bool have = j < w.near_npen[f];
    const int k = have ? w.pen_list[f][j] : 0;
    const float4 b0 = ldg4(bp + 2 * k), b1 = ldg4(bp + 2 * k + 1);
    const float fx = w.foot[f][0], fy = w.foot[f][1], fz = w.foot[f][2];
    const float dx = b0.x - fx, dy = b0.y - fy, dz = b0.z - fz;
    const float e0 = fmaxf(fabsf(b1.z * dx + b1.w * dy) - b0.w, 0.f), e1 = fmaxf(fabsf(b1.z * dy - b1.w * dx) - b1.x, 0.f);
    const float e2 = fmaxf(fabsf(dz) - b1.y, 0.f);
    const bool maybe = have && (e0 * e0 + e1 * e1 + e2 * e2 < r2);
    if (any_lane(maybe)) {
      const float bx[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float l[3], pt[3];
      const float dist = maybe ? sphere_box_local(bx, w.foot[f], r, l, pt) : 1.f;
      const bool hit = maybe && dist < 0.f;
      const unsigned m = wballot(hit);
      if (hit) {
        const int idx = popc(m & lt);
        if (idx < MAXCAND) { w.cand.pair[idx] = f * NBOX + k; w.cand.dist[idx] = dist; w.cand.cd2[idx] = sqdist3(dx, dy, dz); }
      }
      ncand = popc(m);
    }
  } else {
    float ft[4][3];
#pragma unroll
    for (int f = 0; f < 4; f++) { ft[f][0] = w.foot[f][0]; ft[f][1] = w.foot[f][1]; ft[f][2] = w.foot[f][2]; }
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
      const int k = it * 32 + lane;
      const bool valid = k < nb;
      const int kc = valid ? k : 0;
      const float4 b0 = ldg4(bp + 2 * kc), b1 = ldg4(bp + 2 * kc + 1);
      unsigned maybe = 0u;
#pragma unroll
      for (int f = 0; f < 4; f++) {
        const float dx = b0.x - ft[f][0], dy = b0.y - ft[f][1], dz = b0.z - ft[f][2];
        const float e0 = fmaxf(fabsf(b1.z * dx + b1.w * dy) - b0.w, 0.f), e1 = fmaxf(fabsf(b1.z * dy - b1.w * dx) - b1.x, 0.f);
        const float e2 = fmaxf(fabsf(dz) - b1.y, 0.f);
        if (valid && (e0 * e0 + e1 * e1 + e2 * e2 < r2)) maybe |= 1u << f;
      }
      if (any_lane(maybe != 0u)) {
        const float bx[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll 1
        for (int f = 0; f < 4; f++) {
          const bool mb = (maybe >> f) & 1u;
          float l[3], pt[3];
          const float dist = mb ? sphere_box_local(bx, w.foot[f], r, l, pt) : 1.f;
          const bool hit = mb && dist < 0.f;
          const unsigned m = wballot(hit);
          if (hit) {
            const int idx = ncand + popc(m & lt);
            if (idx < MAXCAND) {
              w.cand.pair[idx] = f * NBOX + k; w.cand.dist[idx] = dist;
              w.cand.cd2[idx] = sqdist3(b0.x - ft[f][0], b0.y - ft[f][1], b0.z - ft[f][2]);
            }
          }
          ncand += popc(m);
        }
      }
    }
  }
  if (ncand > MAXCAND) ncand = MAXCAND;
  syncwarp();
  if (ncand == 0) return;
  // broad-phase rank of every penetrating pair among all 4*nb pairs (keep the max_geom_pairs nearest
  // centres); lane c ends up owning candidate c
  const bool cull = (GC.max_geom_pairs > -1) && (4 * nb > GC.max_geom_pairs);
  const float inf = __int_as_float(0x7f800000);
  int mycnt = 0;
  if (cull) {
    const float myt = lane < ncand ? w.cand.cd2[lane] : 0.f;
    // every pair that can precede a candidate with threshold < W_RCEN^2 has its centre within W_RCEN of its foot: it is in
    // that foot's centre list. Each lane keys (at most) two listed pairs of its foot ONCE; a rank is then two ballots.
    if (!full && all_lanes(myt < W_RCEN * W_RCEN)) {
      const int f = lane >> 3, j = lane & 7, n = w.near_ncen[f];
      const float fx = w.foot[f][0], fy = w.foot[f][1], fz = w.foot[f][2];
      float v0 = inf, v1 = inf;
      int id0 = 0x7fffffff, id1 = 0x7fffffff;
      if (j < n) {
        const int k = w.cen_list[f][j];
        const float4 b0 = ldg4(bp + 2 * k);
        v0 = sqdist3(b0.x - fx, b0.y - fy, b0.z - fz); id0 = f * NBOX + k;
      }
      if (j + 8 < n) {
        const int k = w.cen_list[f][j + 8];
        const float4 b0 = ldg4(bp + 2 * k);
        v1 = sqdist3(b0.x - fx, b0.y - fy, b0.z - fz); id1 = f * NBOX + k;
      }
      static_assert(W_QCEN == 16, "two listed pairs per lane");
#pragma unroll 1
      for (int c = 0; c < ncand; c++) {
        const float t = w.cand.cd2[c];
        const int pi = w.cand.pair[c];
        const int cnt = popc(wballot((v0 < t) || (v0 == t && id0 < pi))) + popc(wballot((v1 < t) || (v1 == t && id1 < pi)));
        if (lane == c) mycnt = cnt;
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < ncand; c++) {
        int cnt = 0;
        const float t = w.cand.cd2[c];
        const int pi = w.cand.pair[c];
#pragma unroll 1
        for (int k = lane; k < nb; k += 32) {
          const float4 b0 = ldg4(bp + 2 * k);
#pragma unroll
          for (int f = 0; f < 4; f++) {
            const float v = sqdist3(b0.x - w.foot[f][0], b0.y - w.foot[f][1], b0.z - w.foot[f][2]);
            const int id = f * NBOX + k;
            cnt += (v < t) || (v == t && id < pi);
          }
        }
        cnt = warp_sum_i(cnt);
        if (lane == c) mycnt = cnt;
      }
    }
  }
  // keep the max_contact_points deepest of the surviving pairs (ties -> broad-phase order, then slot order): every
  // candidate counts the candidates that precede it in that order; the first max_contact_points fill the box slots, all
  // at once (each winner computes its own contact frame)
  const bool mine = lane < ncand && !(cull && mycnt >= GC.max_geom_pairs);
  const float myd = mine ? w.cand.dist[lane] : inf;
  const int mypair = w.cand.pair[lane < ncand ? lane : 0];
  const int maxc = GC.max_contact_points < 4 ? GC.max_contact_points : 4;
  syncwarp();
  if (lane < ncand) { w.cand.dist[lane] = myd; w.cand.cd2[lane] = __int_as_float(mycnt); }
  syncwarp();
  int order = 0;
#pragma unroll 1
  for (int c = 0; c < ncand; c++) {
    const float d = w.cand.dist[c];
    const int n = __float_as_int(w.cand.cd2[c]);
    order += (d < myd) || (d == myd && (n < mycnt || (n == mycnt && c < lane)));
  }
  if (mine && order < maxc) {
    const int c = 4 + order, f = mypair / NBOX, k = mypair % NBOX;
    const float4 b0 = ldg4(bp + 2 * k), b1 = ldg4(bp + 2 * k + 1);
    const float bx[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float l[3], pt[3];
    const float dist = sphere_box_local(bx, w.foot[f], r, l, pt);
    float nl[3] = {pt[0] - l[0], pt[1] - l[1], pt[2] - l[2]};
    const float dn = sqrtf(dot3(nl, nl));
    if (dn < PGTT_MINVAL) { nl[0] = nl[1] = nl[2] = 0.f; } else { nl[0] /= dn; nl[1] /= dn; nl[2] /= dn; }
    // contact point: midway between the box point and the sphere surface point
    const float pl0 = 0.5f * (pt[0] + l[0] + nl[0] * r), pl1 = 0.5f * (pt[1] + l[1] + nl[1] * r), pl2 = 0.5f * (pt[2] + l[2] + nl[2] * r);
    float nw[3] = {bx[6] * nl[0] - bx[7] * nl[1], bx[7] * nl[0] + bx[6] * nl[1], nl[2]};
    w.c_pos[c][0] = bx[0] + bx[6] * pl0 - bx[7] * pl1;
    w.c_pos[c][1] = bx[1] + bx[7] * pl0 + bx[6] * pl1;
    w.c_pos[c][2] = bx[2] + pl2;
    make_frame(w.c_frame[c], nw);
    w.c_dist[c] = dist; w.c_leg[c] = f; w.c_box[c] = k;
    w.c_mu[c] = fmaxf(GC.foot_mu, B.m_boxfric[env * NBOX + k]);
  }
}

DEV void collision(WS& w, const EnvBuffers& B, int env, int lane, bool build_lists) {
  const float r = GC.foot_r;
  if (lane < 4) {
    const int c = lane;
    const float dist = w.foot[c][2] - r;
    w.c_dist[c] = dist; w.c_leg[c] = c; w.c_box[c] = -1;
    w.c_pos[c][0] = w.foot[c][0]; w.c_pos[c][1] = w.foot[c][1]; w.c_pos[c][2] = w.foot[c][2] - (r + 0.5f * dist);
    float* fr = w.c_frame[c];
    fr[0] = 0.f; fr[1] = 0.f; fr[2] = 1.f; fr[3] = 0.f; fr[4] = 1.f; fr[5] = 0.f; fr[6] = -1.f; fr[7] = 0.f; fr[8] = 0.f;
    w.c_mu[c] = fmaxf(GC.foot_mu, w.floor_mu);
  } else if (lane < 8) {
    w.c_dist[lane] = 1.f; w.c_leg[lane] = 0; w.c_box[lane] = -2;  // empty slot
  }
  const int nb = GC.n_boxes;
  if (nb > 0) collide_boxes(w, B, env, lane, build_lists);
  STAGE_TRACE(w, lane);   // collision work done, before its barrier
  stage_sync(w, ST_COLLIDE);
}

// ----------------------------------------------------------------------------------------------
// constraint rows (App. A5): one lane per pyramid edge (32) + one lane per joint limit (12)
// ----------------------------------------------------------------------------------------------
struct Rows {
  float jr[9];  // this lane's contact row: 6 base columns + 3 columns of the contact's leg
  float D, aref, jaref, jv;
  int leg, active;
  float lD, laref, ljaref, ljv, lsign;  // joint-limit row of hinge `lane` (lane < 12)
  int lactive;
};

// impedance curve: only solimp power = 2 (the MuJoCo default; every geom / joint of the GO2 scenes) is supported -
// pgtt_create rejects other models, which keeps two powf expansions (~600 instructions) out of the kernel
DEV void kbi(const float* solref, const float* solimp, float pos, float* k, float* b, float* imp) {
  float timeconst = fmaxf(solref[0], 2.f * GC.dt);
  const float dampratio = solref[1];
  const float dmin = fminf(fmaxf(solimp[0], 1e-4f), 0.9999f), dmax = fminf(fmaxf(solimp[1], 1e-4f), 0.9999f);
  const float width = fmaxf(solimp[2], PGTT_MINVAL), mid = fminf(fmaxf(solimp[3], 1e-4f), 0.9999f);
  *k = 1.f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  *b = 2.f / (dmax * timeconst);
  const float x = fabsf(pos) / width;
  const float y = x < mid ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);
  float im = dmin + y * (dmax - dmin);
  im = fminf(fmaxf(im, dmin), dmax);
  if (x > 1.f) im = dmax;
  *imp = im;
}

DEV float row_dot(const Rows& R, const float* x) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; i++) s += R.jr[i] * x[i];
#pragma unroll
  for (int i = 0; i < 3; i++) s += R.jr[6 + i] * x[6 + 3 * R.leg + i];
  return s;
}

DEV void make_rows(WS& w, Rows& R, int lane) {
  // contact-frame Jacobians Jc[c][axis][col]
  for (int item = lane; item < NCON * 9; item += 32) {
    const int c = item / 9, col = item % 9;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (w.c_dist[c] < GC.includemargin) {
      const int d = col_dof(col, w.c_leg[c]);
      const float* cd = w.cdof[d];
      float off[3] = {w.c_pos[c][0] - w.com[0], w.c_pos[c][1] - w.com[1], w.c_pos[c][2] - w.com[2]};
      float jp[3];
      cross3(jp, cd, off);
      jp[0] += cd[3]; jp[1] += cd[4]; jp[2] += cd[5];
      const float sgn = (c < 4) ? 1.f : -1.f;  // plane: body2 = calf; box: body1 = calf
      const float* fr = w.c_frame[c];
      v0 = sgn * dot3(fr, jp); v1 = sgn * dot3(fr + 3, jp); v2 = sgn * dot3(fr + 6, jp);
    }
    w.Jc[c][0][col] = v0; w.Jc[c][1][col] = v1; w.Jc[c][2][col] = v2;
  }
  syncwarp();
  {
    const int c = lane >> 2, e = lane & 3;
    const float pos = w.c_dist[c] - GC.includemargin;
    R.active = pos < 0.f;
    R.leg = w.c_leg[c];
    R.D = 0.f; R.aref = 0.f;
    for (int i = 0; i < 9; i++) R.jr[i] = 0.f;
    if (R.active) {
      const float mu = w.c_mu[c], f = (e & 1) ? -mu : mu;
      const int tn = 1 + (e >> 1);
      for (int i = 0; i < 9; i++) R.jr[i] = w.Jc[c][0][i] + w.Jc[c][tn][i] * f;
      const float t = GC.calf_invw[R.leg];
      const float invweight = (t + f * f * t) * 2.f * f * f / GC.impratio;
      float k, b, imp;
      kbi(c < 4 ? GC.floor_solref : GC.box_solref, c < 4 ? GC.floor_solimp : GC.box_solimp, pos, &k, &b, &imp);
      const float Rr = fmaxf(invweight * (1.f - imp) / imp, PGTT_MINVAL);
      const float vel = row_dot(R, w.qvel);
      R.aref = -b * vel - k * imp * pos;
      R.D = 1.f / Rr;
    }
  }
  R.lactive = 0; R.lD = 0.f; R.laref = 0.f; R.lsign = 0.f;
  if (lane < 12) {
    const float q = w.qpos[7 + lane];
    const float dlo = q - GC.jnt_lo[lane], dhi = GC.jnt_hi[lane] - q;
    const float pos = fminf(dlo, dhi);
    if (pos < 0.f) {
      R.lactive = 1;
      R.lsign = dlo < dhi ? 1.f : -1.f;
      float k, b, imp;
      kbi(GC.lim_solref, GC.lim_solimp, pos, &k, &b, &imp);
      const float Rr = fmaxf(GC.dof_invw[lane] * (1.f - imp) / imp, PGTT_MINVAL);
      R.laref = -b * (R.lsign * w.qvel[6 + lane]) - k * imp * pos;
      R.lD = 1.f / Rr;
    }
  }
  // compact list of active contacts (warp-uniform)
  const unsigned m = wballot((lane & 3) == 0 && R.active);
  if (lane == 0) {
    int n = 0;
    for (int c = 0; c < NCON; c++) if ((m >> (4 * c)) & 1u) w.actlist[n++] = c;
    w.nact = n;
  }
  stage_sync(w, ST_ROWS);
}

// ----------------------------------------------------------------------------------------------
// Newton solver with mjx's bracketed line search (App. A7)
// ----------------------------------------------------------------------------------------------
struct SolveState { float cost, prev_cost, gauss; unsigned am, lam, fac_am, fac_lam; bool fac_valid; };

// jaref already holds J*qacc - aref in R; w.Ma = M*qacc. Computes forces, qfc, cost.
DEV void update_constraint(WS& w, Rows& R, SolveState& S, bool need_force, int lane) {
  const int act = R.active && (R.jaref < 0.f);
  const int lact = R.lactive && (R.ljaref < 0.f);
  float cpart = 0.f, gpart = 0.f;
  if (act) cpart += R.D * R.jaref * R.jaref;
  if (lact) cpart += R.lD * R.ljaref * R.ljaref;
  if (lane < NV) gpart = (w.Ma[lane] - w.qs[lane]) * (w.qacc[lane] - w.qas[lane]);
  cpart = warp_sum(cpart);
  gpart = warp_sum(gpart);
  S.gauss = 0.5f * gpart;
  S.prev_cost = S.cost;
  S.cost = 0.5f * cpart + S.gauss;
  if (!need_force) return;
  S.am = wballot(act); S.lam = wballot(lact);
  // contact-space force and Hessian weights per contact, gathered inside each 4-lane group
  const float f = act ? R.D * -R.jaref : 0.f;
  const float wgt = act ? R.D : 0.f;
  const int e = lane & 3, c = lane >> 2;
  const float fs = f + shfl_xor(f, 1), ws = wgt + shfl_xor(wgt, 1);
  const float fdv = (e & 1) ? -f : f, wdv = (e & 1) ? -wgt : wgt;
  const float fd = fdv + shfl_xor(fdv, 1), wd = wdv + shfl_xor(wdv, 1);
  const float fn = fs + shfl_xor(fs, 2), wn = ws + shfl_xor(ws, 2);
  const float fd2 = shfl_xor(fd, 2), wd2 = shfl_xor(wd, 2), ws2 = shfl_xor(ws, 2);
  if (e == 0) {
    const float mu = w.c_mu[c];
    w.fc[c][0] = fn; w.fc[c][1] = mu * fd; w.fc[c][2] = mu * fd2;
    w.Ac[c][0] = wn; w.Ac[c][1] = mu * wd; w.Ac[c][2] = mu * wd2; w.Ac[c][3] = mu * mu * ws; w.Ac[c][4] = mu * mu * ws2;
  }
  const float lf = lact ? R.lsign * R.lD * -R.ljaref : 0.f;
  if (lane < 12) w.limD[lane] = lact ? R.lD : 0.f;   // Hessian diagonal of the joint-limit rows
  syncwarp();
  if (lane < NV) {
    float s = 0.f;
    const int myleg = (lane - 6) / 3;
    for (int i = 0; i < w.nact; i++) {
      const int cc = w.actlist[i];
      int col = lane;
      if (lane >= 6) { if (w.c_leg[cc] != myleg) continue; col = 6 + (lane - 6) % 3; }
      s += w.Jc[cc][0][col] * w.fc[cc][0] + w.Jc[cc][1][col] * w.fc[cc][1] + w.Jc[cc][2][col] * w.fc[cc][2];
    }
    w.qfc[lane] = s;
  }
  syncwarp();
  if (lane < 12) w.qfc[6 + lane] += lf;
  syncwarp();
}

// element (row ci, column cj) of the 9x9 contact-local Hessian block that this lane accumulates:
// e in [0,36) -> base block, [36,54) -> base x leg coupling, [54,63) -> leg block
DEV void hess_elem(int e, int* ci, int* cj) {
  if (e < 36) { *ci = e / 6; *cj = e % 6; }
  else if (e < 54) { *ci = (e - 36) / 3; *cj = 6 + (e - 36) % 3; }
  else { *ci = 6 + (e - 54) / 3; *cj = 6 + (e - 54) % 3; }
}

// Newton direction: Hessian H = M + J^T diag(D active) J in arrow form, factor, solve.
// Each lane accumulates two fixed elements of the contact-local 9x9 blocks in registers over the
// active contacts (leg-dependent destinations get one accumulator per leg), then writes H once.
// H only depends on the active set (J, D are fixed within a solve), so the factor of the previous
// iteration is re-used bit for bit when the set did not change.
DEV void newton_direction(WS& w, SolveState& S, int lane) {
  syncwarp();   // w.grad was just written by lanes < NV
  if (!(S.fac_valid && S.am == S.fac_am && S.lam == S.fac_lam)) {
    const int e1 = lane + 32;
    int ci0, cj0, ci1, cj1;
    hess_elem(lane, &ci0, &cj0);
    hess_elem(e1 < 63 ? e1 : 62, &ci1, &cj1);
    float acc0 = 0.f, accb = 0.f, accl[4] = {0.f, 0.f, 0.f, 0.f};
    const int nact = w.nact;
#pragma unroll 1
    for (int i = 0; i < nact; i++) {
      const int c = w.actlist[i], leg = w.c_leg[c];
      const float A0 = w.Ac[c][0], A1 = w.Ac[c][1], A2 = w.Ac[c][2], A3 = w.Ac[c][3], A4 = w.Ac[c][4];
      {
        const float j0 = w.Jc[c][0][cj0], j1 = w.Jc[c][1][cj0], j2 = w.Jc[c][2][cj0];
        const float g0 = A0 * j0 + A1 * j1 + A2 * j2, g1 = A1 * j0 + A3 * j1, g2 = A2 * j0 + A4 * j2;
        acc0 += w.Jc[c][0][ci0] * g0 + w.Jc[c][1][ci0] * g1 + w.Jc[c][2][ci0] * g2;
      }
      {
        const float j0 = w.Jc[c][0][cj1], j1 = w.Jc[c][1][cj1], j2 = w.Jc[c][2][cj1];
        const float g0 = A0 * j0 + A1 * j1 + A2 * j2, g1 = A1 * j0 + A3 * j1, g2 = A2 * j0 + A4 * j2;
        const float v = w.Jc[c][0][ci1] * g0 + w.Jc[c][1][ci1] * g1 + w.Jc[c][2][ci1] * g2;
        accb += v;
#pragma unroll
        for (int g = 0; g < 4; g++) accl[g] += (leg == g) ? v : 0.f;
      }
    }
    w.HB[lane] = w.MB[lane] + acc0;
    if (e1 < 36) w.HB[e1] = w.MB[e1] + accb;
    else if (e1 < 54) {
#pragma unroll
      for (int g = 0; g < 4; g++) w.HC[g * 18 + (e1 - 36)] = w.MC[g * 18 + (e1 - 36)] + accl[g];
    } else if (e1 < 63) {
      const int e3 = e1 - 54, j = e3 / 3;
      const bool diag = (e3 % 3) == j;
#pragma unroll
      for (int g = 0; g < 4; g++) w.HA[g * 9 + e3] = w.MA[g * 9 + e3] + accl[g] + (diag ? w.limD[3 * g + j] : 0.f);
    }
    syncwarp();
    arrow_factor(w, w.HB, w.HC, w.HA, lane);
    S.fac_valid = true; S.fac_am = S.am; S.fac_lam = S.lam;
  }
  arrow_solve(w, w.grad, w.search, lane);
  if (lane < NV) w.search[lane] = -w.search[lane];
  syncwarp();
}

// 1-D restriction of the cost along the search direction: cost / derivatives at alpha. Per-row
// quadratics live in registers; their sum over the rows active at alpha only depends on the active
// SET, so the butterfly sums are cached per (contact-row mask, limit-row mask) - most evaluations of
// one line search see the set of alpha = 0 or of the Newton step. Cached sums are the identical
// butterfly results, so caching does not change a single bit.
struct LSPoint { float alpha, cost, d0, d1; };
struct LSCache { unsigned m[2], lm[2]; float s[2][3]; };

DEV LSPoint ls_eval(float alpha, const Rows& R, const float* q, const float* lq, const float* qg, LSCache& C, int lane) {
  const unsigned m = wballot(R.active && (R.jaref + alpha * R.jv < 0.f));
  const unsigned lm = wballot(R.lactive && (R.ljaref + alpha * R.ljv < 0.f));
  float s0, s1, s2;
  if (m == C.m[0] && lm == C.lm[0]) { s0 = C.s[0][0]; s1 = C.s[0][1]; s2 = C.s[0][2]; }
  else if (m == C.m[1] && lm == C.lm[1]) { s0 = C.s[1][0]; s1 = C.s[1][1]; s2 = C.s[1][2]; }
  else {
    const bool on = (m >> lane) & 1u, lon = (lm >> lane) & 1u;
    s0 = 0.f; s1 = 0.f; s2 = 0.f;
    if (on) { s0 += q[0]; s1 += q[1]; s2 += q[2]; }
    if (lon) { s0 += lq[0]; s1 += lq[1]; s2 += lq[2]; }
    s0 = warp_sum(s0) + qg[0]; s1 = warp_sum(s1) + qg[1]; s2 = warp_sum(s2) + qg[2];
    C.m[1] = m; C.lm[1] = lm; C.s[1][0] = s0; C.s[1][1] = s1; C.s[1][2] = s2;
  }
  LSPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * s2 + alpha * s1 + s0;
  p.d0 = 2.f * alpha * s2 + s1;
  p.d1 = 2.f * s2 + (s2 == 0.f ? PGTT_MINVAL : 0.f);
  return p;
}

DEV void linesearch(WS& w, Rows& R, const SolveState& S, int lane) {
  arrow_mul(w.MB, w.MC, w.MA, w.search, w.mv, lane);
  syncwarp();
  R.jv = R.active ? row_dot(R, w.search) : 0.f;
  R.ljv = R.lactive ? R.lsign * w.search[6 + (lane < 12 ? lane : 0)] : 0.f;
  float a = 0.f, b = 0.f, c2 = 0.f;
  if (lane < NV) {
    const float sv = w.search[lane];
    a = sv * sv; b = sv * (w.Ma[lane] - w.qs[lane]); c2 = sv * w.mv[lane];
  }
  a = warp_sum(a); b = warp_sum(b); c2 = warp_sum(c2);
  const float smag = sqrtf(a) * GC.solver_scale;
  const float gtol = GC.tolerance * GC.ls_tolerance * smag;
  const float qg[3] = {S.gauss, b, 0.5f * c2};
  const float q[3] = {0.5f * R.jaref * R.jaref * R.D, R.jv * R.jaref * R.D, 0.5f * R.jv * R.jv * R.D};
  const float lq[3] = {0.5f * R.ljaref * R.ljaref * R.lD, R.ljv * R.ljaref * R.lD, 0.5f * R.ljv * R.ljv * R.lD};
  LSCache C;
  {  // prime both cache entries with the alpha = 0 active set
    const unsigned m = wballot(R.active && (R.jaref < 0.f)), lm = wballot(R.lactive && (R.ljaref < 0.f));
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if ((m >> lane) & 1u) { s0 += q[0]; s1 += q[1]; s2 += q[2]; }
    if ((lm >> lane) & 1u) { s0 += lq[0]; s1 += lq[1]; s2 += lq[2]; }
    s0 = warp_sum(s0) + qg[0]; s1 = warp_sum(s1) + qg[1]; s2 = warp_sum(s2) + qg[2];
    for (int t = 0; t < 2; t++) { C.m[t] = m; C.lm[t] = lm; C.s[t][0] = s0; C.s[t][1] = s1; C.s[t][2] = s2; }
  }
  const LSPoint p0 = ls_eval(0.f, R, q, lq, qg, C, lane);
  const LSPoint l0 = ls_eval(p0.alpha - fdiv_(p0.d0, p0.d1), R, q, lq, qg, C, lane);
  const bool lesser = l0.d0 < p0.d0;
  LSPoint hi = lesser ? p0 : l0, lo = lesser ? l0 : p0;
  bool swap = true;
  int it = 0;
#pragma unroll 1
  for (;;) {
    bool done = it >= GC.ls_iterations;
    done |= (!swap) && (it > 0);
    done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
    done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
    if (all_lanes(done)) break;
    const LSPoint lo_next = ls_eval(lo.alpha - fdiv_(lo.d0, lo.d1), R, q, lq, qg, C, lane);
    const LSPoint hi_next = ls_eval(hi.alpha - fdiv_(hi.d0, hi.d1), R, q, lq, qg, C, lane);
    const LSPoint mid = ls_eval(0.5f * (lo.alpha + hi.alpha), R, q, lq, qg, C, lane);
    const bool s_lo_next = (lo.d0 > 0.f) || (lo.d0 < lo_next.d0);
    if (s_lo_next) lo = lo_next;
    const bool s_lo_mid = (mid.d0 < 0.f) && (lo.d0 < mid.d0);
    if (s_lo_mid) lo = mid;
    const bool s_hi_next = (hi.d0 < 0.f) || (hi.d0 > hi_next.d0);
    if (s_hi_next) hi = hi_next;
    const bool s_hi_mid = (mid.d0 > 0.f) && (hi.d0 > mid.d0);
    if (s_hi_mid) hi = mid;
    swap = s_lo_next || s_lo_mid || s_hi_next || s_hi_mid;
    it++;
  }
  const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  const float alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (improved) {
    if (lane < NV) { w.qacc[lane] += w.search[lane] * alpha; w.Ma[lane] += w.mv[lane] * alpha; }
    R.jaref += R.jv * alpha;
    R.ljaref += R.ljv * alpha;
  }
  syncwarp();
}

// set qacc = x (shared vector), recompute Ma and jaref
DEV void ctx_init(WS& w, Rows& R, const float* x, int lane) {
  if (lane < NV) w.qacc[lane] = x[lane];
  syncwarp();
  arrow_mul(w.MB, w.MC, w.MA, w.qacc, w.Ma, lane);
  R.jaref = R.active ? row_dot(R, w.qacc) - R.aref : 0.f;
  R.ljaref = R.lactive ? R.lsign * w.qacc[6 + (lane < 12 ? lane : 0)] - R.laref : 0.f;
  syncwarp();
}

DEV int solve(WS& w, Rows& R, int lane) {
  SolveState S;
  // warm start: whichever of qacc_warmstart / qacc_smooth has the lower cost
  S.cost = 0.f; S.prev_cost = 0.f;
  float c2[2];
#pragma unroll 1
  for (int t = 0; t < 2; t++) {
    ctx_init(w, R, t == 0 ? w.warm : w.qas, lane);
    update_constraint(w, R, S, false, lane);
    c2[t] = S.cost;
  }
  if (all_lanes(c2[0] < c2[1])) ctx_init(w, R, w.warm, lane);
  S.cost = __int_as_float(0x7f800000);  // +inf
  S.prev_cost = 0.f;
  S.fac_valid = false;
  int niter = 0;
#pragma unroll 1
  for (;;) {
    // mjx order is constraint update -> gradient + Newton direction -> convergence test; the direction
    // is only consumed by the next line search, so it is computed after the test (same results)
    update_constraint(w, R, S, true, lane);
    float gn = 0.f;
    if (lane < NV) { const float gr = w.Ma[lane] - w.qs[lane] - w.qfc[lane]; w.grad[lane] = gr; gn = gr * gr; }
    if (niter > 0 && GC.iterations == 1) break;
    const float improvement = fdiv_(S.prev_cost - S.cost, GC.solver_scale);
    gn = warp_sum(gn);
    const float gradient = fdiv_(sqrtf(gn), GC.solver_scale);
    bool done = niter >= GC.iterations;
    done |= improvement < GC.tolerance;
    done |= gradient < GC.tolerance;
    if (all_lanes(done) && GC.iterations != 1) break;
    newton_direction(w, S, lane);
    linesearch(w, R, S, lane);
    niter++;
  }
  if (lane < NV) w.warm[lane] = w.qacc[lane];
  STAGE_TRACE(w, lane);   // 5: solver done, before the barrier
  stage_sync(w, ST_POSTSOLVE);
  return niter;
}

// ----------------------------------------------------------------------------------------------
// sensors (App. A9) - evaluated from the forward pass of the substep, i.e. pre-integration (Q2)
// ----------------------------------------------------------------------------------------------
DEV void sensors(WS& w, int lane) {
  const float* R = w.xmat0;
  if (lane < 4) {
    // sensor slot k = FR FL RR RL  ->  leg (qpos order FL FR RL RR)
    const int k = lane, g = k ^ 1;
    float imu[3];
    for (int i = 0; i < 3; i++) imu[i] = w.xpos0[i] + R[3 * i] * GC.imu_pos[0] + R[3 * i + 1] * GC.imu_pos[1] + R[3 * i + 2] * GC.imu_pos[2];
    const float rel[3] = {w.foot[g][0] - imu[0], w.foot[g][1] - imu[1], w.foot[g][2] - imu[2]};
    for (int i = 0; i < 3; i++) w.sens[25 + 3 * k + i] = R[i] * rel[0] + R[3 + i] * rel[1] + R[6 + i] * rel[2];
    const float* cv = w.cvel_calf[g];
    const float off[3] = {w.foot[g][0] - w.com[0], w.foot[g][1] - w.com[1], w.foot[g][2] - w.com[2]};
    float c[3];
    cross3(c, cv, off);
    for (int i = 0; i < 3; i++) w.sens[37 + 3 * k + i] = cv[3 + i] + c[i];
  } else if (lane == 4) {
    float imu[3];
    for (int i = 0; i < 3; i++) imu[i] = w.xpos0[i] + R[3 * i] * GC.imu_pos[0] + R[3 * i + 1] * GC.imu_pos[1] + R[3 * i + 2] * GC.imu_pos[2];
    const float* cv = w.cvel_base;
    const float off[3] = {imu[0] - w.com[0], imu[1] - w.com[1], imu[2] - w.com[2]};
    float c[3], lin[3];
    cross3(c, cv, off);
    for (int i = 0; i < 3; i++) lin[i] = cv[3 + i] + c[i];
    float ang_l[3], lin_l[3];
    for (int i = 0; i < 3; i++) {
      ang_l[i] = R[i] * cv[0] + R[3 + i] * cv[1] + R[6 + i] * cv[2];
      lin_l[i] = R[i] * lin[0] + R[3 + i] * lin[1] + R[6 + i] * lin[2];
    }
    for (int i = 0; i < 3; i++) { w.sens[i] = ang_l[i]; w.sens[10 + i] = imu[i]; w.sens[13 + i] = lin[i]; w.sens[16 + i] = cv[i]; w.sens[19 + i] = lin_l[i]; }
    {  // framequat of the imu site = normalised base quaternion
      float qw = w.qpos[3], qx = w.qpos[4], qy = w.qpos[5], qz = w.qpos[6];
      const float qn = 1.0f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
      w.sens[6] = qw * qn; w.sens[7] = qx * qn; w.sens[8] = qy * qn; w.sens[9] = qz * qn;
    }
    w.sens[22] = R[2]; w.sens[23] = R[5]; w.sens[24] = R[8];
    // accelerometer: cacc of the base from the solved qacc, moved to the site, plus w x v
    float cacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, -GC.gravity_z};
    for (int k = 0; k < 6; k++) {
      const float qd = w.qvel[k], qa = w.qacc[k];
      for (int i = 0; i < 6; i++) cacc[i] += w.cdofd_base[k][i] * qd + w.cdof[k][i] * qa;
    }
    float c1[3], corr[3];
    cross3(c1, cacc, off);
    cross3(corr, ang_l, lin_l);
    for (int i = 0; i < 3; i++) {
      const float aw0 = cacc[3] + c1[0], aw1 = cacc[4] + c1[1], aw2 = cacc[5] + c1[2];
      w.sens[3 + i] = (R[i] * aw0 + R[3 + i] * aw1 + R[6 + i] * aw2) + corr[i];
    }
  }
  stage_sync(w, ST_SENSORS);
}

// ----------------------------------------------------------------------------------------------
// mjx.forward and the Euler update
// ----------------------------------------------------------------------------------------------
DEV int forward(WS& w, const EnvBuffers& B, int env, int lane, bool with_sensors, bool build_lists) {
  // fixed-duration stages first, then the two whose duration depends on the env (collision: number of penetrating
  // boxes; solver: Newton iterations) back to back, so that the lockstep CTA waits once per substep, not twice
  STAGE_TRACE(w, lane);   // 0
  kinematics(w, lane);
  com_inertia_cdof(w, lane);
  crb_and_inertia(w, lane);
  STAGE_TRACE(w, lane);   // 1: position stage done
  velocity_rne(w, lane);
  smooth_forces(w, lane);
  STAGE_TRACE(w, lane);   // 2: velocity stage done
  collision(w, B, env, lane, build_lists);   // after the inertia stages: its scratch aliases w.crb
  STAGE_TRACE(w, lane);   // 3: collision done (incl. its barrier, if enabled)
  Rows R;
  make_rows(w, R, lane);
  arrow_factor(w, w.MB, w.MC, w.MA, lane);
  arrow_solve(w, w.qs, w.qas, lane);
  stage_sync(w, ST_PRESOLVE);
  STAGE_TRACE(w, lane);   // 4: rows + smooth solve done
  const int niter = solve(w, R, lane);
  STAGE_TRACE(w, lane);   // 6: after the post-solver barrier (5 is taken inside solve(), before the barrier)
  if (with_sensors) sensors(w, lane);
  return niter;
}

DEV void euler(WS& w, int lane) {
  const float dt = GC.dt;
  if (lane < NV) w.qvel[lane] += dt * w.qacc[lane];
  syncwarp();
  if (lane < 3) w.qpos[lane] += dt * w.qvel[lane];
  else if (lane == 3) {
    float v[3] = {w.qvel[3], w.qvel[4], w.qvel[5]};
    float n = sqrtf(dot3(v, v));
    if (n < PGTT_MINVAL) { v[0] = v[1] = v[2] = 0.f; n = 0.f; } else { v[0] /= n; v[1] /= n; v[2] /= n; }
    float s, c;
    sincos_(0.5f * dt * n, &s, &c);
    const float bw = c, bx = v[0] * s, by = v[1] * s, bz = v[2] * s;
    const float aw = w.qpos[3], ax = w.qpos[4], ay = w.qpos[5], az = w.qpos[6];
    float rw = aw * bw - ax * bx - ay * by - az * bz;
    float rx = aw * bx + ax * bw + ay * bz - az * by;
    float ry = aw * by - ax * bz + ay * bw + az * bx;
    float rz = aw * bz + ax * by - ay * bx + az * bw;
    const float rn = 1.0f / sqrtf(rw * rw + rx * rx + ry * ry + rz * rz);
    w.qpos[3] = rw * rn; w.qpos[4] = rx * rn; w.qpos[5] = ry * rn; w.qpos[6] = rz * rn;
  } else if (lane >= 6 && lane < NV) w.qpos[lane + 1] += dt * w.qvel[lane];
  stage_sync(w, ST_EULER);
}
