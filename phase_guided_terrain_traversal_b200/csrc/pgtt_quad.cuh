// pgtt_quad.cuh - generation-2 env kernel: FOUR LANES PER ENV (one per leg), eight envs per warp.
//
// Same arithmetic as pgtt_physics.cuh / pgtt_env.cuh (reference semantics: go2/joystick_pgtt.py:50-611 over
// mjx.step) re-mapped onto the machine: lane g of a quad owns leg g (FL FR RL RR) - its three
// bodies, three hinges, its foot/plane contact and box-contact slot g - and runs the leg's kinematic, inertia and
// RNE chains serially IN REGISTERS with full instruction-level parallelism; base-block quantities are computed
// redundantly by the four lanes (bitwise identical: every cross-lane sum is an xor butterfly) so the only
// communication is 2-shuffle quad reductions. There is no shared-memory workspace per env, no CTA barrier and no
// idle lane in the chains; one warp advances eight envs, so 4096 envs occupy 512 of the 592 warp schedulers of a
// B200 once instead of two waves of fourteen-warp CTAs (profiles/r01b -> r01c).
//
// Layout inside a warp: lane = 4 * slot + g, env = 8 * warp + slot. Loops whose trip count depends on the env
// (Newton iterations, line search, candidate rounds) run until every env of the warp is done, with per-env
// predicates on the state updates.
#pragma once
#include "pgtt_env.cuh"

#define QENV 8
#define QCAND 8            // penetrating boxes remembered per foot
#define QPEN 10            // near-list capacities per foot (overflow falls back to the full scan)
#define QCEN 32
#define Q_MARGIN 0.12f     // foot travel within one control step covered by the near lists (checked every substep)
#define Q_RCEN 0.9f        // broad-phase thresholds below Q_RCEN^2 are ranked from the centre lists
#define Q_INF __int_as_float(0x7f800000)

// ----------------------------------------------------------------------------------------------
// quad collectives (full-warp shuffles issued from warp-uniform control flow)
// ----------------------------------------------------------------------------------------------
DEV float qsum(float v) { v += shfl_xor(v, 1); v += shfl_xor(v, 2); return v; }
DEV int qsum_i(int v) { v += shfl_xor(v, 1); v += shfl_xor(v, 2); return v; }
DEV float qmin(float v) { v = fminf(v, shfl_xor(v, 1)); v = fminf(v, shfl_xor(v, 2)); return v; }
DEV float qmax(float v) { v = fmaxf(v, shfl_xor(v, 1)); v = fmaxf(v, shfl_xor(v, 2)); return v; }
DEV int qmin_i(int v) { int u = shfl_xor(v, 1); v = u < v ? u : v; u = shfl_xor(v, 2); v = u < v ? u : v; return v; }
DEV bool qany(bool p, int qbase) { return ((wballot(p) >> qbase) & 0xFu) != 0u; }

// ----------------------------------------------------------------------------------------------
// per-leg constants, staged once per warp in shared memory (constant memory would serialise on the
// four different leg addresses of a warp; shared memory serves them in one wavefront)
// ----------------------------------------------------------------------------------------------
struct LegC {
  float body_pos[3][3], body_ipos[3][3], body_I[3][6];
  float jnt_lo[3], jnt_hi[3], dof_invw[3], calf_invw;
  float act_bias0[3], act_bias2[3], ctrl_lo[3], ctrl_hi[3], frc_lo[3], frc_hi[3];   // of the actuator driving hinge t
  float default_pose[3], soft_lo[3], soft_hi[3], mt_default[3];
  int act[3];
  int foot_geom;
  float pad[1];   // 81 words: odd stride -> the four legs land in different banks
};

struct QShared {
  LegC leg[4];
  // lane-private columns ([k][lane]: conflict-free, no synchronisation needed)
  float cand_dist[QCAND][32], cand_cd2[QCAND][32];
  int cand_box[QCAND][32], cand_rank[QCAND][32];
  int pen_list[QPEN][32], cen_list[QCEN][32];
};

DEV void q_stage_consts(QShared& S, int lane) {
  if (lane < 4) {
    const int g = lane;
    LegC& L = S.leg[g];
    for (int t = 0; t < 3; t++) {
      const int b = 1 + 3 * g + t, j = 3 * g + t, a = GC.act_of_hinge[j];
      for (int i = 0; i < 3; i++) { L.body_pos[t][i] = GC.body_pos[b][i]; L.body_ipos[t][i] = GC.body_ipos[b][i]; }
      for (int i = 0; i < 6; i++) L.body_I[t][i] = GC.body_I[b][i];
      L.jnt_lo[t] = GC.jnt_lo[j]; L.jnt_hi[t] = GC.jnt_hi[j]; L.dof_invw[t] = GC.dof_invw[j];
      L.act_bias0[t] = GC.act_bias0[a]; L.act_bias2[t] = GC.act_bias2[a];
      L.ctrl_lo[t] = GC.ctrl_lo[a]; L.ctrl_hi[t] = GC.ctrl_hi[a]; L.frc_lo[t] = GC.frc_lo[a]; L.frc_hi[t] = GC.frc_hi[a];
      L.default_pose[t] = GC.default_pose[j]; L.soft_lo[t] = GC.soft_lo[j]; L.soft_hi[t] = GC.soft_hi[j];
      L.act[t] = a; L.mt_default[t] = GC.default_pose[a];
    }
    L.calf_invw = GC.calf_invw[g];
    L.foot_geom = GC.foot_geom[g];
  }
  syncwarp();
}

// ----------------------------------------------------------------------------------------------
// register-resident per-lane data
// ----------------------------------------------------------------------------------------------
struct QModel {                 // per-env model (pgtt_randomize), the slice this lane needs
  float m0, m[3], ipos0[3];
  float arm[3], damp[3], gain[3], bias1[3], q0[3];
  float mtot_inv, floor_mu;
  const float* box;             // this env's terrain: [100][BOXF]
  const float* boxfric;         // this env's box frictions [100]
};
struct QState {                 // generalised coordinates: base part replicated on the 4 lanes, own leg part
  float qb[7], ql[3], vb[6], vl[3], wb[6], wl[3], ctrl[3];
};
struct QVec { float b[6], l[3]; };

struct QKin {
  float pb[3], Rb[9], com[3];
  float p[3][3], R[3][9], foot[3];
  float cinb[10], cin[3][10];
  float cdb[6][6];              // base dofs (0..2 translation = unit vectors, 3..5 rotation)
  float cd[3][6];               // leg dofs
};
struct QMass {                  // arrow inertia matrix slice: MB packed lower [21], MC [6][3], MA packed lower [6]
  float MB[21], MC[6][3], MA[6];
};
struct QFac {                   // arrow factor: leg Cholesky (reciprocal diagonals), Y = C A^-1, Schur Cholesky
  float la[6], Y[6][3], L[21];
};
struct QCon {                   // one contact slot
  float J[3][9];                // contact-frame Jacobian: 6 base columns, 3 columns of the contact's leg
  float D, mu, aref[4], jaref[4], jv[4];
  int leg, active;
};
struct QLim { float D[3], aref[3], jaref[3], jv[3], sign[3]; int active[3]; };
struct QSens {                  // what the task layer reads after the last substep
  float gyro[3], acc[3], quat[4], gpos[3], glin[3], gang[3], llin[3], up[3];
  float fpos[3], fvel[3];       // own foot: position in the imu frame, world linear velocity
  float fworld[3], Rb[9];       // own foot world position (site_xpos), imu site orientation (site_xmat)
  float accG[3][6], acc0[3];    // accelerometer = acc0 + accG * qacc_base (affine in the solver output)
};

#define PK(i, j) ((i) * ((i) + 1) / 2 + (j))   // packed lower index, i >= j

// ----------------------------------------------------------------------------------------------
// load / store
// ----------------------------------------------------------------------------------------------
DEV void q_load_model(QModel& M, const EnvBuffers& B, const LegC& L, int env, int g) {
  M.m0 = B.m_mass[env * NB];
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const int j = 3 * g + t, a = L.act[t];
    M.m[t] = B.m_mass[env * NB + 1 + j];
    M.arm[t] = B.m_armature[env * 12 + j]; M.damp[t] = B.m_damping[env * 12 + j]; M.q0[t] = B.m_qpos0[env * 12 + j];
    M.gain[t] = B.m_gain[env * 12 + a]; M.bias1[t] = B.m_bias1[env * 12 + a];
    M.ipos0[t] = B.m_ipos[env * 3 + t];
  }
  const float ml = (M.m[0] + M.m[1]) + M.m[2];
  M.mtot_inv = 1.0f / (M.m0 + qsum(ml));
  M.floor_mu = B.m_floorfric[env];
  M.box = nullptr; M.boxfric = nullptr;
  if (GC.n_boxes > 0) {
    M.box = B.terrain + (size_t)B.terrain_index[env] * NBOX * BOXF;
    M.boxfric = B.m_boxfric + (size_t)env * NBOX;
  }
}

DEV void q_load_state(QState& X, const EnvBuffers& B, int env, int g) {
#pragma unroll
  for (int i = 0; i < 7; i++) X.qb[i] = B.qpos[env * NQ + i];
#pragma unroll
  for (int i = 0; i < 6; i++) { X.vb[i] = B.qvel[env * NV + i]; X.wb[i] = B.warm[env * NV + i]; }
#pragma unroll
  for (int t = 0; t < 3; t++) {
    X.ql[t] = B.qpos[env * NQ + 7 + 3 * g + t]; X.vl[t] = B.qvel[env * NV + 6 + 3 * g + t]; X.wl[t] = B.warm[env * NV + 6 + 3 * g + t];
  }
}

// ----------------------------------------------------------------------------------------------
// position stage (App. A1-A3)
// ----------------------------------------------------------------------------------------------
DEV void q_kinematics(QKin& K, const QState& X, const QModel& M, const LegC& L) {
  float qw = X.qb[3], qx = X.qb[4], qy = X.qb[5], qz = X.qb[6];
  const float qn = 1.0f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  qw *= qn; qx *= qn; qy *= qn; qz *= qn;
  float* Rb = K.Rb;
  Rb[0] = qw * qw + qx * qx - qy * qy - qz * qz; Rb[1] = 2 * (qx * qy - qw * qz); Rb[2] = 2 * (qx * qz + qw * qy);
  Rb[3] = 2 * (qx * qy + qw * qz); Rb[4] = qw * qw - qx * qx + qy * qy - qz * qz; Rb[5] = 2 * (qy * qz - qw * qx);
  Rb[6] = 2 * (qx * qz - qw * qy); Rb[7] = 2 * (qy * qz + qw * qx); Rb[8] = qw * qw - qx * qx - qy * qy + qz * qz;
  float p[3] = {X.qb[0], X.qb[1], X.qb[2]}, R[9];
#pragma unroll
  for (int i = 0; i < 3; i++) K.pb[i] = p[i];
#pragma unroll
  for (int i = 0; i < 9; i++) R[i] = Rb[i];
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const float* o = L.body_pos[t];
#pragma unroll
    for (int i = 0; i < 3; i++) p[i] += R[3 * i] * o[0] + R[3 * i + 1] * o[1] + R[3 * i + 2] * o[2];
    float s, c;
    sincos_(X.ql[t] - M.q0[t], &s, &c);
    if (t == 0) {  // hip: rotation about the local x axis
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const float c1 = R[3 * i + 1], c2 = R[3 * i + 2];
        R[3 * i + 1] = c1 * c + c2 * s;
        R[3 * i + 2] = -c1 * s + c2 * c;
      }
    } else {       // thigh, calf: rotation about the local y axis
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const float c0 = R[3 * i], c2 = R[3 * i + 2];
        R[3 * i] = c0 * c - c2 * s;
        R[3 * i + 2] = c0 * s + c2 * c;
      }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) K.p[t][i] = p[i];
#pragma unroll
    for (int i = 0; i < 9; i++) K.R[t][i] = R[i];
  }
  const float* o = GC.foot_pos;
#pragma unroll
  for (int i = 0; i < 3; i++) K.foot[i] = p[i] + R[3 * i] * o[0] + R[3 * i + 1] * o[1] + R[3 * i + 2] * o[2];
}

// COM-frame inertia (10 numbers, see inert_mul) of a body with mass m, body-frame tensor I, orientation R, COM offset o
DEV void q_cinert(float* ci, float m, const float* I, const float* R, const float* o) {
  float T[9];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float r0 = R[3 * i], r1 = R[3 * i + 1], r2 = R[3 * i + 2];
    T[3 * i] = r0 * I[0] + r1 * I[3] + r2 * I[4];
    T[3 * i + 1] = r0 * I[3] + r1 * I[1] + r2 * I[5];
    T[3 * i + 2] = r0 * I[4] + r1 * I[5] + r2 * I[2];
  }
  const float oo = dot3(o, o);
  ci[0] = dot3(T, R) + m * (oo - o[0] * o[0]);
  ci[1] = dot3(T + 3, R + 3) + m * (oo - o[1] * o[1]);
  ci[2] = dot3(T + 6, R + 6) + m * (oo - o[2] * o[2]);
  ci[3] = dot3(T, R + 3) - m * o[0] * o[1];
  ci[4] = dot3(T, R + 6) - m * o[0] * o[2];
  ci[5] = dot3(T + 3, R + 6) - m * o[1] * o[2];
  ci[6] = m * o[0]; ci[7] = m * o[1]; ci[8] = m * o[2]; ci[9] = m;
}

DEV void q_com_inertia_cdof(QKin& K, const QModel& M, const LegC& L, float xi_out[4][3]) {
  float xib[3], xi[3][3], ms[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 3; i++) xib[i] = K.pb[i] + K.Rb[3 * i] * M.ipos0[0] + K.Rb[3 * i + 1] * M.ipos0[1] + K.Rb[3 * i + 2] * M.ipos0[2];
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const float* ip = L.body_ipos[t];
    const float* R = K.R[t];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      xi[t][i] = K.p[t][i] + R[3 * i] * ip[0] + R[3 * i + 1] * ip[1] + R[3 * i + 2] * ip[2];
      ms[i] += M.m[t] * xi[t][i];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++) K.com[i] = (M.m0 * xib[i] + qsum(ms[i])) * M.mtot_inv;
  {
    const float o[3] = {xib[0] - K.com[0], xib[1] - K.com[1], xib[2] - K.com[2]};
    q_cinert(K.cinb, M.m0, GC.body_I[0], K.Rb, o);
  }
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const float o[3] = {xi[t][0] - K.com[0], xi[t][1] - K.com[1], xi[t][2] - K.com[2]};
    q_cinert(K.cin[t], M.m[t], L.body_I[t], K.R[t], o);
  }
  // motion axes about the COM
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int i = 0; i < 6; i++) K.cdb[d][i] = (i == 3 + d) ? 1.f : 0.f;
    const float ax[3] = {K.Rb[d], K.Rb[3 + d], K.Rb[6 + d]};
    const float off[3] = {K.com[0] - K.pb[0], K.com[1] - K.pb[1], K.com[2] - K.pb[2]};
    K.cdb[3 + d][0] = ax[0]; K.cdb[3 + d][1] = ax[1]; K.cdb[3 + d][2] = ax[2];
    cross3(K.cdb[3 + d] + 3, ax, off);
  }
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const int col = (t == 0) ? 0 : 1;
    const float* R = K.R[t];
    const float ax[3] = {R[col], R[3 + col], R[6 + col]};
    const float off[3] = {K.com[0] - K.p[t][0], K.com[1] - K.p[t][1], K.com[2] - K.p[t][2]};
    K.cd[t][0] = ax[0]; K.cd[t][1] = ax[1]; K.cd[t][2] = ax[2];
    cross3(K.cd[t] + 3, ax, off);
  }
  if (xi_out) {
#pragma unroll
    for (int i = 0; i < 3; i++) { xi_out[0][i] = xib[i]; xi_out[1][i] = xi[0][i]; xi_out[2][i] = xi[1][i]; xi_out[3][i] = xi[2][i]; }
  }
}

// composite inertias and the arrow inertia matrix
DEV void q_mass_matrix(QMass& Mm, const QKin& K, const QModel& M) {
  float crb[3][10], crbb[10];
#pragma unroll
  for (int k = 0; k < 10; k++) {
    crb[2][k] = K.cin[2][k];
    crb[1][k] = K.cin[1][k] + crb[2][k];
    crb[0][k] = K.cin[0][k] + crb[1][k];
    crbb[k] = K.cinb[k] + qsum(crb[0][k]);
  }
  float F[3][6];
#pragma unroll
  for (int t = 0; t < 3; t++) inert_mul(F[t], crb[t], K.cd[t]);
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int k = 0; k <= j; k++) Mm.MA[PK(j, k)] = dot6(K.cd[k], F[j]) + (j == k ? M.arm[j] : 0.f);
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int j = 0; j < 3; j++) Mm.MC[a][j] = dot6(K.cdb[a], F[j]);
  float FB[6][6];
#pragma unroll
  for (int a = 0; a < 6; a++) inert_mul(FB[a], crbb, K.cdb[a]);
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = 0; b <= a; b++) Mm.MB[PK(a, b)] = dot6(K.cdb[a], FB[b]);
}

// y = M x
DEV void q_mul(QVec& y, const QMass& Mm, const QVec& x) {
#pragma unroll
  for (int a = 0; a < 6; a++) {
    float s = 0.f, c = 0.f;
#pragma unroll
    for (int b = 0; b < 6; b++) s += Mm.MB[a >= b ? PK(a, b) : PK(b, a)] * x.b[b];
#pragma unroll
    for (int j = 0; j < 3; j++) c += Mm.MC[a][j] * x.l[j];
    y.b[a] = s + qsum(c);
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 6; a++) s += Mm.MC[a][j] * x.b[a];
#pragma unroll
    for (int k = 0; k < 3; k++) s += Mm.MA[j >= k ? PK(j, k) : PK(k, j)] * x.l[k];
    y.l[j] = s;
  }
}

// ----------------------------------------------------------------------------------------------
// arrow factorisation / solve: HB [21], HC [6][3], HA [6] -> QFac
// ----------------------------------------------------------------------------------------------
DEV void q_factor(QFac& F, const float* HB, const float (*HC)[3], const float* HA) {
  const float i00 = rsqrt_(fmaxf(HA[PK(0, 0)], PGTT_MINVAL));
  const float l10 = HA[PK(1, 0)] * i00, l20 = HA[PK(2, 0)] * i00;
  const float i11 = rsqrt_(fmaxf(HA[PK(1, 1)] - l10 * l10, PGTT_MINVAL));
  const float l21 = (HA[PK(2, 1)] - l20 * l10) * i11;
  const float i22 = rsqrt_(fmaxf(HA[PK(2, 2)] - l20 * l20 - l21 * l21, PGTT_MINVAL));
  F.la[0] = i00; F.la[1] = l10; F.la[2] = i11; F.la[3] = l20; F.la[4] = l21; F.la[5] = i22;
  float pr[21];
#pragma unroll
  for (int k = 0; k < 21; k++) pr[k] = 0.f;
#pragma unroll
  for (int a = 0; a < 6; a++) {
    const float* cr = HC[a];
    const float z0 = cr[0] * i00, z1 = (cr[1] - l10 * z0) * i11, z2 = (cr[2] - l20 * z0 - l21 * z1) * i22;
    const float y2 = z2 * i22, y1 = (z1 - l21 * y2) * i11, y0 = (z0 - l10 * y1 - l20 * y2) * i00;
    F.Y[a][0] = y0; F.Y[a][1] = y1; F.Y[a][2] = y2;
#pragma unroll
    for (int b = 0; b <= a; b++) pr[PK(a, b)] = y0 * HC[b][0] + y1 * HC[b][1] + y2 * HC[b][2];
  }
  float S[21];
#pragma unroll
  for (int k = 0; k < 21; k++) S[k] = HB[k] - qsum(pr[k]);
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int j = 0; j <= i; j++) {
      float s = S[PK(i, j)];
#pragma unroll
      for (int k = 0; k < j; k++) s -= F.L[PK(i, k)] * F.L[PK(j, k)];
      F.L[PK(i, j)] = (i == j) ? rsqrt_(fmaxf(s, PGTT_MINVAL)) : s * F.L[PK(j, j)];
    }
  }
}

// x = H^-1 r
DEV void q_solve(QVec& x, const QFac& F, const QVec& r) {
  float tb[6];
#pragma unroll
  for (int a = 0; a < 6; a++) {
    const float t = F.Y[a][0] * r.l[0] + F.Y[a][1] * r.l[1] + F.Y[a][2] * r.l[2];
    tb[a] = r.b[a] - qsum(t);
  }
  float xb[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float s = tb[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= F.L[PK(i, k)] * xb[k];
    xb[i] = s * F.L[PK(i, i)];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    float s = xb[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s -= F.L[PK(k, i)] * xb[k];
    xb[i] = s * F.L[PK(i, i)];
  }
  const float r0 = r.l[0], r1 = r.l[1], r2 = r.l[2];
  const float z0 = r0 * F.la[0], z1 = (r1 - F.la[1] * z0) * F.la[2], z2 = (r2 - F.la[3] * z0 - F.la[4] * z1) * F.la[5];
  const float y2 = z2 * F.la[5], y1 = (z1 - F.la[4] * y2) * F.la[2], y0 = (z0 - F.la[1] * y1 - F.la[3] * y2) * F.la[0];
  float v[3] = {y0, y1, y2};
#pragma unroll
  for (int j = 0; j < 3; j++) {
#pragma unroll
    for (int a = 0; a < 6; a++) v[j] -= F.Y[a][j] * xb[a];
    x.l[j] = v[j];
  }
#pragma unroll
  for (int a = 0; a < 6; a++) x.b[a] = xb[a];
}

// ----------------------------------------------------------------------------------------------
// velocity stage + RNE bias forces (App. A6) + passive / actuator forces
// ----------------------------------------------------------------------------------------------
struct QVel { float cvb[6], cvcalf[6], cdd[3][3]; };   // base / calf spatial velocity, cdof_dot (linear part) of the base rotations

DEV void q_rne_smooth(QVec& qs, float* actf, QVel& V, const QKin& K, const QState& X, const QModel& M, const LegC& L) {
  float vb[6] = {0.f, 0.f, 0.f, X.vb[0], X.vb[1], X.vb[2]};
  float ab[6] = {0.f, 0.f, 0.f, 0.f, 0.f, -GC.gravity_z};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    cross3(V.cdd[k], vb + 3, K.cdb[3 + k]);
    const float qd = X.vb[3 + k];
    ab[3] += V.cdd[k][0] * qd; ab[4] += V.cdd[k][1] * qd; ab[5] += V.cdd[k][2] * qd;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float qd = X.vb[3 + k];
#pragma unroll
    for (int i = 0; i < 6; i++) vb[i] += K.cdb[3 + k][i] * qd;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) V.cvb[i] = vb[i];
  float fb[6], tmp[6], tmp2[6];
  inert_mul(fb, K.cinb, ab);
  inert_mul(tmp, K.cinb, vb);
  cross_force(tmp2, vb, tmp);
#pragma unroll
  for (int i = 0; i < 6; i++) fb[i] += tmp2[i];
  float vp[6], ap[6], fl[3][6];
#pragma unroll
  for (int i = 0; i < 6; i++) { vp[i] = vb[i]; ap[i] = ab[i]; }
#pragma unroll
  for (int t = 0; t < 3; t++) {
    float cdot[6];
    cross_motion(cdot, vp, K.cd[t]);
    const float qd = X.vl[t];
#pragma unroll
    for (int i = 0; i < 6; i++) { vp[i] += K.cd[t][i] * qd; ap[i] += cdot[i] * qd; }
    inert_mul(fl[t], K.cin[t], ap);
    inert_mul(tmp, K.cin[t], vp);
    cross_force(tmp2, vp, tmp);
#pragma unroll
    for (int i = 0; i < 6; i++) fl[t][i] += tmp2[i];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) { V.cvcalf[i] = vp[i]; fl[1][i] += fl[2][i]; fl[0][i] += fl[1][i]; }
  float tot[6];
#pragma unroll
  for (int i = 0; i < 6; i++) tot[i] = qsum(fl[0][i]) + fb[i];
#pragma unroll
  for (int a = 0; a < 6; a++) qs.b[a] = -dot6(K.cdb[a], tot);
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const float bias = dot6(K.cd[t], fl[t]);
    const float c = fminf(fmaxf(X.ctrl[t], L.ctrl_lo[t]), L.ctrl_hi[t]);
    float af = M.gain[t] * c + L.act_bias0[t] + M.bias1[t] * X.ql[t] + L.act_bias2[t] * X.vl[t];
    af = fminf(fmaxf(af, L.frc_lo[t]), L.frc_hi[t]);
    actf[t] = af;
    qs.l[t] = (-M.damp[t] * X.vl[t] - bias) + af;
  }
}

// ----------------------------------------------------------------------------------------------
// collision (App. A4): lane g owns the foot-g/plane contact and box-contact slot g
// ----------------------------------------------------------------------------------------------
struct QGeo { float dist, pos[3], fr[9], mu; int leg, box; };   // box: -1 plane, -2 empty slot, >= 0 box index

// squared centre distance, the broad-phase key. Un-contracted so that both passes (and both builds) get the same bits.
DEV float q_sqdist3(float dx, float dy, float dz) { return mul_add_nofma(dz, dz, mul_add_nofma(dy, dy, dx * dx)); }

DEV void q_collide_plane(QGeo& GP, const QKin& K, const QModel& M, int g) {
  const float r = GC.foot_r, dist = K.foot[2] - r;
  GP.dist = dist; GP.leg = g; GP.box = -1;
  GP.pos[0] = K.foot[0]; GP.pos[1] = K.foot[1]; GP.pos[2] = K.foot[2] - (r + 0.5f * dist);
  GP.fr[0] = 0.f; GP.fr[1] = 0.f; GP.fr[2] = 1.f; GP.fr[3] = 0.f; GP.fr[4] = 1.f; GP.fr[5] = 0.f; GP.fr[6] = -1.f; GP.fr[7] = 0.f; GP.fr[8] = 0.f;
  GP.mu = fmaxf(GC.foot_mu, M.floor_mu);
}

// Per control step and foot: the boxes whose SURFACE can come within the foot radius (penetration tests) and the boxes
// whose CENTRE can come within Q_RCEN (broad-phase ranks) while the foot travels at most Q_MARGIN from where the lists
// were built. Both are conservative supersets, so scanning them gives bit-identical contact bookkeeping to scanning all
// 100 boxes; travel beyond the margin, list overflow or a threshold above Q_RCEN^2 falls back to the full scan.
struct QNear { float f0[3]; int npen, ncen; bool over; };

DEV void q_build_near(QNear& Nr, QShared& S, const QModel& M, const float* foot, int lane) {
  const int nb = GC.n_boxes;
  const float4* bp = reinterpret_cast<const float4*>(M.box);
  const float rp = (GC.foot_r + Q_MARGIN) * 1.001f, rc = (Q_RCEN + Q_MARGIN) * 1.001f;
  const float rp2 = rp * rp, rc2 = rc * rc;
  int npen = 0, ncen = 0;
  // the box table of an env (3.2 KB, one of 100 per level) is L2-resident, not L1-resident: fetch ten boxes (twenty
  // 16-byte loads in flight) per round trip instead of leaving the trip count to the unroller
  constexpr int CH = 10;
#pragma unroll 1
  for (int k0 = 0; k0 < nb; k0 += CH) {
    float4 c0[CH], c1[CH];
#pragma unroll
    for (int u = 0; u < CH; u++) {
      const int k = (k0 + u < nb) ? k0 + u : nb - 1;
      c0[u] = ldg4(bp + 2 * k); c1[u] = ldg4(bp + 2 * k + 1);
    }
#pragma unroll
    for (int u = 0; u < CH; u++) {
      const int k = k0 + u;
      if (k < nb) {
        const float4 b0 = c0[u], b1 = c1[u];
        const float dx = b0.x - foot[0], dy = b0.y - foot[1], dz = b0.z - foot[2];
        const float e0 = fmaxf(fabsf(b1.z * dx + b1.w * dy) - b0.w, 0.f), e1 = fmaxf(fabsf(b1.z * dy - b1.w * dx) - b1.x, 0.f);
        const float e2 = fmaxf(fabsf(dz) - b1.y, 0.f);
        if (e0 * e0 + e1 * e1 + e2 * e2 < rp2) { if (npen < QPEN) S.pen_list[npen][lane] = k; npen++; }
        if (dx * dx + dy * dy + dz * dz < rc2) { if (ncen < QCEN) S.cen_list[ncen][lane] = k; ncen++; }
      }
    }
  }
  Nr.f0[0] = foot[0]; Nr.f0[1] = foot[1]; Nr.f0[2] = foot[2];
  Nr.npen = npen; Nr.ncen = ncen; Nr.over = (npen > QPEN) || (ncen > QCEN);
}

DEV void q_collide_boxes(QGeo& GX, QNear& Nr, bool build, QShared& S, const QKin& K, const QModel& M, int lane, int g, int qbase) {
  GX.dist = 1.f; GX.leg = 0; GX.box = -2; GX.mu = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) GX.pos[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 9; i++) GX.fr[i] = 0.f;
  const int nb = GC.n_boxes;
  if (nb <= 0) return;
  if (build) q_build_near(Nr, S, M, K.foot, lane);
  const float r = GC.foot_r, r2 = r * r * 1.0001f;   // conservative pre-filter; the exact test runs only where it passes
  const float fx = K.foot[0], fy = K.foot[1], fz = K.foot[2];
  const float4* bp = reinterpret_cast<const float4*>(M.box);
  bool lists_ok;
  {
    const float tx = fx - Nr.f0[0], ty = fy - Nr.f0[1], tz = fz - Nr.f0[2];
    lists_ok = !Nr.over && (tx * tx + ty * ty + tz * tz <= Q_MARGIN * Q_MARGIN);
  }
  const bool full = GC.quad_fullscan || any_lane(!lists_ok);
  int ncand = 0;
  {
    const int cnt = full ? nb : Nr.npen;
#pragma unroll 2
    for (int i = 0; i < cnt; i++) {
      const int k = full ? i : S.pen_list[i][lane];
      const float4 b0 = ldg4(bp + 2 * k), b1 = ldg4(bp + 2 * k + 1);
      const float dx = b0.x - fx, dy = b0.y - fy, dz = b0.z - fz;
      const float e0 = fmaxf(fabsf(b1.z * dx + b1.w * dy) - b0.w, 0.f), e1 = fmaxf(fabsf(b1.z * dy - b1.w * dx) - b1.x, 0.f);
      const float e2 = fmaxf(fabsf(dz) - b1.y, 0.f);
      if (e0 * e0 + e1 * e1 + e2 * e2 < r2) {
        const float bx[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float l[3], pt[3];
        const float dist = sphere_box_local(bx, K.foot, r, l, pt);
        if (dist < 0.f && ncand < QCAND) {
          S.cand_box[ncand][lane] = k; S.cand_dist[ncand][lane] = dist; S.cand_cd2[ncand][lane] = q_sqdist3(dx, dy, dz);
          S.cand_rank[ncand][lane] = g * NBOX + k;
          ncand++;
        }
      }
    }
  }
  if (!any_lane(ncand > 0)) return;
  // broad-phase rank of every penetrating pair among the env's 4 * nb pairs (mjx keeps the max_geom_pairs nearest
  // centres): each lane counts over its own foot's pairs for the candidates of all four feet, one round per slot
  const bool cull = (GC.max_geom_pairs > -1) && (4 * nb > GC.max_geom_pairs);
  if (cull) {
#pragma unroll 1
    for (int c = 0; c < QCAND; c++) {
      if (!any_lane(c < ncand)) break;
      const bool have = c < ncand;
      const float myt = have ? S.cand_cd2[c][lane] : -1.f;
      const int myid = have ? g * NBOX + S.cand_box[c][lane] : 0;
      const bool full2 = full || any_lane(have && !(myt < Q_RCEN * Q_RCEN));
      float t[4]; int id[4], cnt[4];
#pragma unroll
      for (int s = 0; s < 4; s++) { t[s] = shfl(myt, qbase | s); id[s] = shfl(myid, qbase | s); cnt[s] = 0; }
      const int n2 = full2 ? nb : Nr.ncen;
#pragma unroll 4
      for (int i = 0; i < n2; i++) {
        const int k = full2 ? i : S.cen_list[i][lane];
        const float4 b0 = ldg4(bp + 2 * k);
        const float v = q_sqdist3(b0.x - fx, b0.y - fy, b0.z - fz);
        const int idk = g * NBOX + k;
#pragma unroll
        for (int s = 0; s < 4; s++) cnt[s] += (v < t[s]) || (v == t[s] && idk < id[s]);
      }
      int mine = 0;
#pragma unroll
      for (int s = 0; s < 4; s++) { const int rs = qsum_i(cnt[s]); if (s == g) mine = rs; }
      if (have) {
        S.cand_rank[c][lane] = mine;
        if (mine >= GC.max_geom_pairs) S.cand_dist[c][lane] = Q_INF;
      }
    }
  }
  // the max_contact_points deepest survivors, deepest first (ties -> broad-phase order); selection s fills slot s
  const int maxc = GC.max_contact_points < 4 ? GC.max_contact_points : 4;
  int selk = 0, self = g, selv = 0;
#pragma unroll 1
  for (int s = 0; s < maxc; s++) {
    float bd = Q_INF; int br = 0x7fffffff, bc = 0;
    for (int c = 0; c < ncand; c++) {
      const float d = S.cand_dist[c][lane]; const int rk = S.cand_rank[c][lane];
      if (d < bd || (d == bd && d < Q_INF && rk < br)) { bd = d; br = rk; bc = c; }
    }
    const float dmin = qmin(bd);
    if (all_lanes(!(dmin < Q_INF))) break;
    const int rmin = qmin_i((bd == dmin) ? br : 0x7fffffff);
    const bool win = (dmin < Q_INF) && (bd == dmin) && (br == rmin);
    const int kw = qsum_i(win ? S.cand_box[bc][lane] : 0), fw = qsum_i(win ? g : 0), vw = qsum_i(win ? 1 : 0);
    if (win) S.cand_dist[bc][lane] = Q_INF;
    if (g == s && vw) { selk = kw; self = fw; selv = 1; }
  }
  const float ox = shfl(fx, qbase | self), oy = shfl(fy, qbase | self), oz = shfl(fz, qbase | self);
  if (selv) {
    const float4 b0 = ldg4(bp + 2 * selk), b1 = ldg4(bp + 2 * selk + 1);
    const float bx[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float fo[3] = {ox, oy, oz};
    float l[3], pt[3];
    const float dist = sphere_box_local(bx, fo, r, l, pt);
    float nl[3] = {pt[0] - l[0], pt[1] - l[1], pt[2] - l[2]};
    const float dn = sqrtf(dot3(nl, nl));
    if (dn < PGTT_MINVAL) { nl[0] = nl[1] = nl[2] = 0.f; } else { nl[0] /= dn; nl[1] /= dn; nl[2] /= dn; }
    // contact point: midway between the box point and the sphere surface point
    const float pl0 = 0.5f * (pt[0] + l[0] + nl[0] * r), pl1 = 0.5f * (pt[1] + l[1] + nl[1] * r), pl2 = 0.5f * (pt[2] + l[2] + nl[2] * r);
    const float nw[3] = {bx[6] * nl[0] - bx[7] * nl[1], bx[7] * nl[0] + bx[6] * nl[1], nl[2]};
    GX.pos[0] = bx[0] + bx[6] * pl0 - bx[7] * pl1;
    GX.pos[1] = bx[1] + bx[7] * pl0 + bx[6] * pl1;
    GX.pos[2] = bx[2] + pl2;
    make_frame(GX.fr, nw);
    GX.dist = dist; GX.leg = self; GX.box = selk;
    GX.mu = fmaxf(GC.foot_mu, M.boxfric[selk]);
  }
}

// ----------------------------------------------------------------------------------------------
// constraint rows (App. A5)
// ----------------------------------------------------------------------------------------------
// u = Jc x for a contact whose leg part of x is xl
DEV void q_jdot(float* u, const QCon& C, const float* xb, const float* xl) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 6; a++) s += C.J[i][a] * xb[a];
#pragma unroll
    for (int j = 0; j < 3; j++) s += C.J[i][6 + j] * xl[j];
    u[i] = s;
  }
}
// pyramid edge e of a contact from the frame-space product u: (u0 + mu u1, u0 - mu u1, u0 + mu u2, u0 - mu u2)
DEV float q_edge(const float* u, float mu, int e) { return u[0] + ((e & 1) ? -mu : mu) * u[1 + (e >> 1)]; }

// cdl: motion axes of the contact's leg, vl: joint velocities of that leg, sgn: +1 plane (body2 = calf), -1 box (body1 = calf)
DEV void q_contact_rows(QCon& C, const QGeo& G, const QKin& K, const float (*cdl)[6], const float* vb, const float* vl, float calf_invw, float sgn, bool plane) {
  C.leg = G.leg; C.mu = G.mu; C.D = 0.f;
  const float pos = G.dist - GC.includemargin;
  C.active = (G.box != -2) && (pos < 0.f);
  const float off[3] = {G.pos[0] - K.com[0], G.pos[1] - K.com[1], G.pos[2] - K.com[2]};
#pragma unroll
  for (int col = 0; col < 9; col++) {
    const float* cd = col < 6 ? K.cdb[col] : cdl[col - 6];
    float jp[3];
    cross3(jp, cd, off);
    jp[0] += cd[3]; jp[1] += cd[4]; jp[2] += cd[5];
#pragma unroll
    for (int i = 0; i < 3; i++) C.J[i][col] = C.active ? sgn * dot3(G.fr + 3 * i, jp) : 0.f;
  }
#pragma unroll
  for (int e = 0; e < 4; e++) { C.aref[e] = 0.f; C.jaref[e] = 0.f; C.jv[e] = 0.f; }
  if (C.active) {
    const float mu = G.mu, t = calf_invw;
    const float invweight = (t + mu * mu * t) * 2.f * mu * mu / GC.impratio;
    float k, b, imp;
    kbi(plane ? GC.floor_solref : GC.box_solref, plane ? GC.floor_solimp : GC.box_solimp, pos, &k, &b, &imp);
    const float Rr = fmaxf(invweight * (1.f - imp) / imp, PGTT_MINVAL);
    C.D = 1.f / Rr;
    float u[3];
    q_jdot(u, C, vb, vl);
#pragma unroll
    for (int e = 0; e < 4; e++) C.aref[e] = -b * q_edge(u, mu, e) - k * imp * pos;
  }
}

DEV void q_limit_rows(QLim& Lm, const QState& X, const LegC& L) {
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const float q = X.ql[t];
    const float dlo = q - L.jnt_lo[t], dhi = L.jnt_hi[t] - q;
    const float pos = fminf(dlo, dhi);
    Lm.active[t] = 0; Lm.D[t] = 0.f; Lm.aref[t] = 0.f; Lm.sign[t] = 0.f; Lm.jaref[t] = 0.f; Lm.jv[t] = 0.f;
    if (pos < 0.f) {
      Lm.active[t] = 1;
      Lm.sign[t] = dlo < dhi ? 1.f : -1.f;
      float k, b, imp;
      kbi(GC.lim_solref, GC.lim_solimp, pos, &k, &b, &imp);
      const float Rr = fmaxf(L.dof_invw[t] * (1.f - imp) / imp, PGTT_MINVAL);
      Lm.aref[t] = -b * (Lm.sign[t] * X.vl[t]) - k * imp * pos;
      Lm.D[t] = 1.f / Rr;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Newton solver with mjx's bracketed line search (App. A7), eight envs in lockstep
// ----------------------------------------------------------------------------------------------
struct QSol {
  QVec qacc, Ma, qfc;
  float cost, prev, gauss;
  unsigned bits, fac_bits;
  bool fac_valid;
};

// leg part of x for the box contact of this lane's slot (owned by another leg)
DEV void q_gather_leg(float* xl, const QVec& x, int leg, int qbase) {
#pragma unroll
  for (int j = 0; j < 3; j++) xl[j] = shfl(x.l[j], qbase | leg);
}

// qacc = x; Ma = M x; jaref = J x - aref
DEV void q_ctx_init(QSol& S, QCon& CP, QCon& CX, QLim& Lm, const QMass& Mm, const QVec& x, int qbase, bool anyX) {
  S.qacc = x;
  q_mul(S.Ma, Mm, x);
  float u[3];
  q_jdot(u, CP, x.b, x.l);
#pragma unroll
  for (int e = 0; e < 4; e++) CP.jaref[e] = CP.active ? q_edge(u, CP.mu, e) - CP.aref[e] : 0.f;
  if (anyX) {
    float xl[3];
    q_gather_leg(xl, x, CX.leg, qbase);
    q_jdot(u, CX, x.b, xl);
#pragma unroll
    for (int e = 0; e < 4; e++) CX.jaref[e] = CX.active ? q_edge(u, CX.mu, e) - CX.aref[e] : 0.f;
  }
#pragma unroll
  for (int t = 0; t < 3; t++) Lm.jaref[t] = Lm.active[t] ? Lm.sign[t] * x.l[t] - Lm.aref[t] : 0.f;
}

struct QForce { float fcP[3], AP[5], fcX[3], AX[5], limD[3]; };

DEV void q_contact_force(float* fc, float* A, const QCon& C, unsigned& bits, int shift) {
  float f[4], w[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const bool act = C.active && (C.jaref[e] < 0.f);
    f[e] = act ? C.D * -C.jaref[e] : 0.f;
    w[e] = act ? C.D : 0.f;
    bits |= (act ? 1u : 0u) << (shift + e);
  }
  const float mu = C.mu;
  fc[0] = (f[0] + f[1]) + (f[2] + f[3]); fc[1] = mu * (f[0] - f[1]); fc[2] = mu * (f[2] - f[3]);
  A[0] = (w[0] + w[1]) + (w[2] + w[3]); A[1] = mu * (w[0] - w[1]); A[2] = mu * (w[2] - w[3]);
  A[3] = mu * mu * (w[0] + w[1]); A[4] = mu * mu * (w[2] + w[3]);
}

// cost at the current qacc; with need_force also the constraint forces, qfrc_constraint and the active-set bits
DEV void q_update_constraint(QSol& S, QForce& Fo, const QCon& CP, const QCon& CX, const QLim& Lm, const QVec& qs, const QVec& qas,
                             bool need_force, int g, int qbase, bool anyX) {
  float cpart = 0.f;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    if (CP.active && CP.jaref[e] < 0.f) cpart += CP.D * CP.jaref[e] * CP.jaref[e];
    if (CX.active && CX.jaref[e] < 0.f) cpart += CX.D * CX.jaref[e] * CX.jaref[e];
  }
#pragma unroll
  for (int t = 0; t < 3; t++) if (Lm.active[t] && Lm.jaref[t] < 0.f) cpart += Lm.D[t] * Lm.jaref[t] * Lm.jaref[t];
  float gl = 0.f, gb = 0.f;
#pragma unroll
  for (int t = 0; t < 3; t++) gl += (S.Ma.l[t] - qs.l[t]) * (S.qacc.l[t] - qas.l[t]);
#pragma unroll
  for (int a = 0; a < 6; a++) gb += (S.Ma.b[a] - qs.b[a]) * (S.qacc.b[a] - qas.b[a]);
  cpart = qsum(cpart);
  const float gpart = gb + qsum(gl);
  S.gauss = 0.5f * gpart;
  S.prev = S.cost;
  S.cost = 0.5f * cpart + S.gauss;
  if (!need_force) return;
  unsigned bits = 0u;
  q_contact_force(Fo.fcP, Fo.AP, CP, bits, 0);
  q_contact_force(Fo.fcX, Fo.AX, CX, bits, 4);
  float lf[3];
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const bool act = Lm.active[t] && (Lm.jaref[t] < 0.f);
    lf[t] = act ? Lm.sign[t] * Lm.D[t] * -Lm.jaref[t] : 0.f;
    Fo.limD[t] = act ? Lm.D[t] : 0.f;
    bits |= (act ? 1u : 0u) << (8 + t);
  }
  S.bits = bits;
#pragma unroll
  for (int a = 0; a < 6; a++) {
    float s = CP.J[0][a] * Fo.fcP[0] + CP.J[1][a] * Fo.fcP[1] + CP.J[2][a] * Fo.fcP[2];
    s += CX.J[0][a] * Fo.fcX[0] + CX.J[1][a] * Fo.fcX[1] + CX.J[2][a] * Fo.fcX[2];
    S.qfc.b[a] = qsum(s);
  }
  float wx[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    S.qfc.l[j] = (CP.J[0][6 + j] * Fo.fcP[0] + CP.J[1][6 + j] * Fo.fcP[1] + CP.J[2][6 + j] * Fo.fcP[2]) + lf[j];
    wx[j] = CX.J[0][6 + j] * Fo.fcX[0] + CX.J[1][6 + j] * Fo.fcX[1] + CX.J[2][6 + j] * Fo.fcX[2];
  }
  if (anyX) {
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
      if (!any_lane(g == s && CX.active)) continue;
      const int lf_ = shfl(CX.leg, qbase | s);
#pragma unroll
      for (int j = 0; j < 3; j++) { const float v = shfl(wx[j], qbase | s); if (lf_ == g) S.qfc.l[j] += v; }
    }
  }
}

// one element of the contact-local 9x9 Hessian block J^T A J
DEV float q_hloc(const QCon& C, const float* A, int ci, int cj) {
  const float j0 = C.J[0][cj], j1 = C.J[1][cj], j2 = C.J[2][cj];
  const float g0 = A[0] * j0 + A[1] * j1 + A[2] * j2, g1 = A[1] * j0 + A[3] * j1, g2 = A[2] * j0 + A[4] * j2;
  return C.J[0][ci] * g0 + C.J[1][ci] * g1 + C.J[2][ci] * g2;
}

// H = M + J^T diag(D active) J in arrow form, factorised. H only depends on the active set.
DEV void q_hessian_factor(QFac& F, const QMass& Mm, const QForce& Fo, const QCon& CP, const QCon& CX, int g, int qbase, bool anyX) {
  float HB[21], HC[6][3], HA[6];
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = 0; b <= a; b++) {
      float v = q_hloc(CP, Fo.AP, a, b);
      if (anyX) v += q_hloc(CX, Fo.AX, a, b);
      HB[PK(a, b)] = Mm.MB[PK(a, b)] + qsum(v);
    }
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int j = 0; j < 3; j++) HC[a][j] = Mm.MC[a][j] + q_hloc(CP, Fo.AP, a, 6 + j);
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int k = 0; k <= j; k++) HA[PK(j, k)] = Mm.MA[PK(j, k)] + q_hloc(CP, Fo.AP, 6 + j, 6 + k) + (j == k ? Fo.limD[j] : 0.f);
  if (anyX) {
    float hx[24];
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int j = 0; j < 3; j++) hx[3 * a + j] = q_hloc(CX, Fo.AX, a, 6 + j);
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int k = 0; k <= j; k++) hx[18 + PK(j, k)] = q_hloc(CX, Fo.AX, 6 + j, 6 + k);
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
      if (!any_lane(g == s && CX.active)) continue;
      const bool mine = shfl(CX.leg, qbase | s) == g;
#pragma unroll
      for (int i = 0; i < 24; i++) {
        const float v = shfl(hx[i], qbase | s);
        if (mine) { if (i < 18) HC[i / 3][i % 3] += v; else HA[i - 18] += v; }
      }
    }
  }
  q_factor(F, HB, HC, HA);
}

struct QLSPoint { float alpha, cost, d0, d1; };

// NA line-search points at once (the three candidates of one bracketing iteration are independent: evaluating them
// together triples the instruction-level parallelism of the row loop and of the quad reductions). Inactive rows have
// jaref = jv = 0, so their test is false without looking at the active flag.
template <int NA>
DEV void q_ls_eval(QLSPoint* out, const float* alpha, const QCon& CP, const QCon& CX, const QLim& Lm, const float (*qP)[3], const float (*qX)[3],
                   const float (*qL)[3], const float* qg, bool anyX, bool anyL) {
  float s[NA][3];
#pragma unroll
  for (int a = 0; a < NA; a++) { s[a][0] = 0.f; s[a][1] = 0.f; s[a][2] = 0.f; }
#pragma unroll
  for (int e = 0; e < 4; e++)
#pragma unroll
    for (int a = 0; a < NA; a++)
      if (CP.jaref[e] + alpha[a] * CP.jv[e] < 0.f) { s[a][0] += qP[e][0]; s[a][1] += qP[e][1]; s[a][2] += qP[e][2]; }
  if (anyX) {
#pragma unroll
    for (int e = 0; e < 4; e++)
#pragma unroll
      for (int a = 0; a < NA; a++)
        if (CX.jaref[e] + alpha[a] * CX.jv[e] < 0.f) { s[a][0] += qX[e][0]; s[a][1] += qX[e][1]; s[a][2] += qX[e][2]; }
  }
  if (anyL) {
#pragma unroll
    for (int t = 0; t < 3; t++)
#pragma unroll
      for (int a = 0; a < NA; a++)
        if (Lm.jaref[t] + alpha[a] * Lm.jv[t] < 0.f) { s[a][0] += qL[t][0]; s[a][1] += qL[t][1]; s[a][2] += qL[t][2]; }
  }
#pragma unroll
  for (int a = 0; a < NA; a++) {
    const float s0 = qsum(s[a][0]) + qg[0], s1 = qsum(s[a][1]) + qg[1], s2 = qsum(s[a][2]) + qg[2];
    out[a].alpha = alpha[a];
    out[a].cost = alpha[a] * alpha[a] * s2 + alpha[a] * s1 + s0;
    out[a].d0 = 2.f * alpha[a] * s2 + s1;
    out[a].d1 = 2.f * s2 + (s2 == 0.f ? PGTT_MINVAL : 0.f);
  }
}

// `live`: this env still iterates (envs that converged ride along without changing state)
DEV void q_linesearch(QSol& S, QCon& CP, QCon& CX, QLim& Lm, const QMass& Mm, const QVec& search, const QVec& qs, bool live, int qbase, bool anyX, bool anyL) {
  QVec mv;
  q_mul(mv, Mm, search);
  {
    float u[3];
    q_jdot(u, CP, search.b, search.l);
#pragma unroll
    for (int e = 0; e < 4; e++) CP.jv[e] = CP.active ? q_edge(u, CP.mu, e) : 0.f;
    if (anyX) {
      float xl[3];
      q_gather_leg(xl, search, CX.leg, qbase);
      q_jdot(u, CX, search.b, xl);
#pragma unroll
      for (int e = 0; e < 4; e++) CX.jv[e] = CX.active ? q_edge(u, CX.mu, e) : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 3; t++) Lm.jv[t] = Lm.active[t] ? Lm.sign[t] * search.l[t] : 0.f;
  }
  float al = 0.f, bl = 0.f, cl = 0.f, ab = 0.f, bb = 0.f, cb = 0.f;
#pragma unroll
  for (int t = 0; t < 3; t++) { const float sv = search.l[t]; al += sv * sv; bl += sv * (S.Ma.l[t] - qs.l[t]); cl += sv * mv.l[t]; }
#pragma unroll
  for (int a = 0; a < 6; a++) { const float sv = search.b[a]; ab += sv * sv; bb += sv * (S.Ma.b[a] - qs.b[a]); cb += sv * mv.b[a]; }
  const float a = ab + qsum(al), b = bb + qsum(bl), c2 = cb + qsum(cl);
  const float smag = sqrtf(a) * GC.solver_scale;
  const float gtol = GC.tolerance * GC.ls_tolerance * smag;
  const float qg[3] = {S.gauss, b, 0.5f * c2};
  float qP[4][3], qX[4][3], qL[3][3];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    qP[e][0] = 0.5f * CP.jaref[e] * CP.jaref[e] * CP.D; qP[e][1] = CP.jv[e] * CP.jaref[e] * CP.D; qP[e][2] = 0.5f * CP.jv[e] * CP.jv[e] * CP.D;
    qX[e][0] = 0.5f * CX.jaref[e] * CX.jaref[e] * CX.D; qX[e][1] = CX.jv[e] * CX.jaref[e] * CX.D; qX[e][2] = 0.5f * CX.jv[e] * CX.jv[e] * CX.D;
  }
#pragma unroll
  for (int t = 0; t < 3; t++) {
    qL[t][0] = 0.5f * Lm.jaref[t] * Lm.jaref[t] * Lm.D[t]; qL[t][1] = Lm.jv[t] * Lm.jaref[t] * Lm.D[t]; qL[t][2] = 0.5f * Lm.jv[t] * Lm.jv[t] * Lm.D[t];
  }
  QLSPoint p0, l0;
  {
    const float a0 = 0.f;
    q_ls_eval<1>(&p0, &a0, CP, CX, Lm, qP, qX, qL, qg, anyX, anyL);
    const float a1 = p0.alpha - fdiv_(p0.d0, p0.d1);
    q_ls_eval<1>(&l0, &a1, CP, CX, Lm, qP, qX, qL, qg, anyX, anyL);
  }
  const bool lesser = l0.d0 < p0.d0;
  QLSPoint hi = lesser ? p0 : l0, lo = lesser ? l0 : p0;
  bool swap = true;
  int it = 0;
  // An env takes at most ls_iterations rounds and the 64 envs of a lockstep CTA practically always contain one that takes
  // them all (8 envs: 4.88 of 5 on average), so the loop runs the full count without a CTA vote per round (one barrier
  // less per round; rounds of finished envs are predicated off, as they were under the vote). The Newton-level vote
  // re-aligns the warps once per Newton iteration.
#pragma unroll 1
  for (int round = 0;; round++) {
    bool done = !live || it >= GC.ls_iterations;
    done |= (!swap) && (it > 0);
    done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
    done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
    if (GC.quad_ls_vote ? cta_all(done) : (round >= GC.ls_iterations)) break;
    QLSPoint pts[3];
    const float al3[3] = {lo.alpha - fdiv_(lo.d0, lo.d1), hi.alpha - fdiv_(hi.d0, hi.d1), 0.5f * (lo.alpha + hi.alpha)};
    q_ls_eval<3>(pts, al3, CP, CX, Lm, qP, qX, qL, qg, anyX, anyL);
    const QLSPoint lo_next = pts[0], hi_next = pts[1], mid = pts[2];
    if (!done) {
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
  }
  const bool improved = live && ((lo.cost < p0.cost) || (hi.cost < p0.cost));
  const float alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (improved) {
#pragma unroll
    for (int i = 0; i < 6; i++) { S.qacc.b[i] += search.b[i] * alpha; S.Ma.b[i] += mv.b[i] * alpha; }
#pragma unroll
    for (int t = 0; t < 3; t++) { S.qacc.l[t] += search.l[t] * alpha; S.Ma.l[t] += mv.l[t] * alpha; Lm.jaref[t] += Lm.jv[t] * alpha; }
#pragma unroll
    for (int e = 0; e < 4; e++) { CP.jaref[e] += CP.jv[e] * alpha; CX.jaref[e] += CX.jv[e] * alpha; }
  }
}

// returns the number of Newton iterations of this env; S.qacc is the solution.
// One loop, one inlined copy of every stage: phase 0 / 1 evaluate the cost at qacc_warmstart / qacc_smooth (the
// warm start is whichever is lower), phases >= 2 are the Newton iterations.
DEV int q_solve_constraints(QSol& S, QCon& CP, QCon& CX, QLim& Lm, const QMass& Mm, const QVec& qs, const QVec& qas, const QVec& warm,
                            int g, int qbase) {
  const bool anyX = any_lane(CX.active);
  const bool anyL = any_lane(Lm.active[0] | Lm.active[1] | Lm.active[2]);
  QForce Fo;
  QFac F = {};
  QVec keep_q = warm, keep_Ma = warm;
  float keep_j[11], cw = 0.f;
#pragma unroll
  for (int i = 0; i < 11; i++) keep_j[i] = 0.f;
  S.cost = 0.f; S.prev = 0.f;
  S.fac_valid = false; S.fac_bits = 0u; S.bits = 0u;
  int niter = 0;
  bool live = true;
#pragma unroll 1
  for (int phase = 0;; phase++) {
    if (phase < 2) {
      QVec x0;
#pragma unroll
      for (int i = 0; i < 6; i++) x0.b[i] = phase == 0 ? warm.b[i] : qas.b[i];
#pragma unroll
      for (int t = 0; t < 3; t++) x0.l[t] = phase == 0 ? warm.l[t] : qas.l[t];
      q_ctx_init(S, CP, CX, Lm, Mm, x0, qbase, anyX);
    }
    // mjx order is constraint update -> gradient + Newton direction -> convergence test; the direction is only
    // consumed by the next line search, so it is computed after the test (same results)
    q_update_constraint(S, Fo, CP, CX, Lm, qs, qas, phase >= 2, g, qbase, anyX);
    if (phase == 0) {
      cw = S.cost; keep_q = S.qacc; keep_Ma = S.Ma;
#pragma unroll
      for (int e = 0; e < 4; e++) { keep_j[e] = CP.jaref[e]; keep_j[4 + e] = CX.jaref[e]; }
#pragma unroll
      for (int t = 0; t < 3; t++) keep_j[8 + t] = Lm.jaref[t];
      continue;
    }
    if (phase == 1) {
      if (cw < S.cost) {   // per-env choice (identical on the four lanes)
        S.qacc = keep_q; S.Ma = keep_Ma;
#pragma unroll
        for (int e = 0; e < 4; e++) { CP.jaref[e] = keep_j[e]; CX.jaref[e] = keep_j[4 + e]; }
#pragma unroll
        for (int t = 0; t < 3; t++) Lm.jaref[t] = keep_j[8 + t];
      }
      S.cost = Q_INF; S.prev = 0.f;
      continue;
    }
    QVec grad;
    float gnl = 0.f, gnb = 0.f;
#pragma unroll
    for (int t = 0; t < 3; t++) { grad.l[t] = S.Ma.l[t] - qs.l[t] - S.qfc.l[t]; gnl += grad.l[t] * grad.l[t]; }
#pragma unroll
    for (int a = 0; a < 6; a++) { grad.b[a] = S.Ma.b[a] - qs.b[a] - S.qfc.b[a]; gnb += grad.b[a] * grad.b[a]; }
    const float gn = gnb + qsum(gnl);
    if (live) {
      const float improvement = fdiv_(S.prev - S.cost, GC.solver_scale);
      const float gradient = fdiv_(sqrtf(gn), GC.solver_scale);
      bool done = niter >= GC.iterations;
      if (GC.iterations == 1) done = niter > 0;
      else { done |= improvement < GC.tolerance; done |= gradient < GC.tolerance; }
      live = !done;
    }
    if (!cta_any(live)) break;
    // Newton direction; the factor is re-used while the active set of the env is unchanged. A lane whose own rows
    // did not change still takes the new factor when a sibling's rows did (the decision is per quad).
    const bool stale = live && !(S.fac_valid && S.bits == S.fac_bits);
    const unsigned stale_m = wballot(stale);
    if (stale_m != 0u) {
      const bool refac = ((stale_m >> qbase) & 0xFu) != 0u;
      QFac Fn;
      q_hessian_factor(Fn, Mm, Fo, CP, CX, g, qbase, anyX);
      if (refac) { F = Fn; S.fac_valid = true; S.fac_bits = S.bits; }
    }
    QVec search;
    q_solve(search, F, grad);
#pragma unroll
    for (int a = 0; a < 6; a++) search.b[a] = -search.b[a];
#pragma unroll
    for (int t = 0; t < 3; t++) search.l[t] = -search.l[t];
    q_linesearch(S, CP, CX, Lm, Mm, search, qs, live, qbase, anyX, anyL);
    if (live) niter++;
  }
  return niter;
}

// ----------------------------------------------------------------------------------------------
// sensors (App. A9): everything but the accelerometer is known before the solver; the accelerometer is affine
// in the base part of qacc, so its map is prepared here and evaluated after the solve
// ----------------------------------------------------------------------------------------------
DEV void q_sensors_pre(QSens& Z, const QKin& K, const QVel& V, const QState& X) {
  const float* R = K.Rb;
  float imu[3];
#pragma unroll
  for (int i = 0; i < 9; i++) Z.Rb[i] = R[i];
#pragma unroll
  for (int i = 0; i < 3; i++) Z.fworld[i] = K.foot[i];
#pragma unroll
  for (int i = 0; i < 3; i++) imu[i] = K.pb[i] + R[3 * i] * GC.imu_pos[0] + R[3 * i + 1] * GC.imu_pos[1] + R[3 * i + 2] * GC.imu_pos[2];
  {
    const float rel[3] = {K.foot[0] - imu[0], K.foot[1] - imu[1], K.foot[2] - imu[2]};
#pragma unroll
    for (int i = 0; i < 3; i++) Z.fpos[i] = R[i] * rel[0] + R[3 + i] * rel[1] + R[6 + i] * rel[2];
    const float off[3] = {K.foot[0] - K.com[0], K.foot[1] - K.com[1], K.foot[2] - K.com[2]};
    float c[3];
    cross3(c, V.cvcalf, off);
#pragma unroll
    for (int i = 0; i < 3; i++) Z.fvel[i] = V.cvcalf[3 + i] + c[i];
  }
  const float* cv = V.cvb;
  const float off[3] = {imu[0] - K.com[0], imu[1] - K.com[1], imu[2] - K.com[2]};
  float c[3], lin[3];
  cross3(c, cv, off);
#pragma unroll
  for (int i = 0; i < 3; i++) lin[i] = cv[3 + i] + c[i];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Z.gyro[i] = R[i] * cv[0] + R[3 + i] * cv[1] + R[6 + i] * cv[2];
    Z.llin[i] = R[i] * lin[0] + R[3 + i] * lin[1] + R[6 + i] * lin[2];
    Z.gpos[i] = imu[i]; Z.glin[i] = lin[i]; Z.gang[i] = cv[i];
  }
  {
    float qw = X.qb[3], qx = X.qb[4], qy = X.qb[5], qz = X.qb[6];
    const float qn = 1.0f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
    Z.quat[0] = qw * qn; Z.quat[1] = qx * qn; Z.quat[2] = qy * qn; Z.quat[3] = qz * qn;
  }
  Z.up[0] = R[2]; Z.up[1] = R[5]; Z.up[2] = R[8];
  // accelerometer: cacc of the base from qacc, moved to the site, rotated, plus w x v
  float a0[3] = {0.f, 0.f, -GC.gravity_z};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float qd = X.vb[3 + k];
    a0[0] += V.cdd[k][0] * qd; a0[1] += V.cdd[k][1] * qd; a0[2] += V.cdd[k][2] * qd;
  }
  float corr[3];
  cross3(corr, Z.gyro, Z.llin);
#pragma unroll
  for (int i = 0; i < 3; i++) Z.acc0[i] = (R[i] * a0[0] + R[3 + i] * a0[1] + R[6 + i] * a0[2]) + corr[i];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float ck[3];
    cross3(ck, K.cdb[k], off);
    ck[0] += K.cdb[k][3]; ck[1] += K.cdb[k][4]; ck[2] += K.cdb[k][5];
#pragma unroll
    for (int i = 0; i < 3; i++) Z.accG[i][k] = R[i] * ck[0] + R[3 + i] * ck[1] + R[6 + i] * ck[2];
  }
}
DEV void q_sensors_post(QSens& Z, const QVec& qacc) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float s = Z.acc0[i];
#pragma unroll
    for (int k = 0; k < 6; k++) s += Z.accG[i][k] * qacc.b[k];
    Z.acc[i] = s;
  }
}

// semi-implicit Euler with quaternion integration; the base part runs redundantly on the four lanes
DEV void q_euler(QState& X, const QVec& qacc) {
  const float dt = GC.dt;
#pragma unroll
  for (int i = 0; i < 6; i++) X.vb[i] += dt * qacc.b[i];
#pragma unroll
  for (int t = 0; t < 3; t++) { X.vl[t] += dt * qacc.l[t]; X.ql[t] += dt * X.vl[t]; }
#pragma unroll
  for (int i = 0; i < 3; i++) X.qb[i] += dt * X.vb[i];
  float v[3] = {X.vb[3], X.vb[4], X.vb[5]};
  float n = sqrtf(dot3(v, v));
  if (n < PGTT_MINVAL) { v[0] = v[1] = v[2] = 0.f; n = 0.f; } else { v[0] /= n; v[1] /= n; v[2] /= n; }
  float s, c;
  sincos_(0.5f * dt * n, &s, &c);
  const float bw = c, bx = v[0] * s, by = v[1] * s, bz = v[2] * s;
  const float aw = X.qb[3], ax = X.qb[4], ay = X.qb[5], az = X.qb[6];
  const float rw = aw * bw - ax * bx - ay * by - az * bz;
  const float rx = aw * bx + ax * bw + ay * bz - az * by;
  const float ry = aw * by - ax * bz + ay * bw + az * bx;
  const float rz = aw * bz + ax * by - ay * bx + az * bw;
  const float rn = 1.0f / sqrtf(rw * rw + rx * rx + ry * ry + rz * rz);
  X.qb[3] = rw * rn; X.qb[4] = rx * rn; X.qb[5] = ry * rn; X.qb[6] = rz * rn;
}

// ----------------------------------------------------------------------------------------------
// mjx.forward for eight envs. Outputs: qacc (S.qacc), actuator forces, contacts, optional sensors / debug dump
// ----------------------------------------------------------------------------------------------
struct QFwd { QGeo GP, GX; float actf[3]; int niter; };

template <bool DBG>
DEV void q_forward(QFwd& O, QSol& S, QSens& Z, QNear& Nr, bool build_near, QShared& Sh, const QState& X, const QModel& M, const LegC& L, bool sens, float* dbg, int lane, int g, int qbase) {
  QMass Mm;
  QVec qs, qas;
  QCon CP, CX;
  QLim Lm;
  {
    QKin K;
    q_kinematics(K, X, M, L);
    q_collide_plane(O.GP, K, M, g);
    q_collide_boxes(O.GX, Nr, build_near, Sh, K, M, lane, g, qbase);
    float xi[4][3];
    q_com_inertia_cdof(K, M, L, DBG ? xi : nullptr);
    q_mass_matrix(Mm, K, M);
    QVel V;
    q_rne_smooth(qs, O.actf, V, K, X, M, L);
    if (sens) q_sensors_pre(Z, K, V, X);
    // constraint rows: own foot/plane contact, then the box contact of slot g (owned by leg GX.leg)
    q_contact_rows(CP, O.GP, K, K.cd, X.vb, X.vl, L.calf_invw, 1.f, true);
    const bool anyX = any_lane(O.GX.box >= 0);
    if (anyX) {
      float cdx[3][6], vlx[3];
      const int src = qbase | O.GX.leg;
#pragma unroll
      for (int t = 0; t < 3; t++) {
        vlx[t] = shfl(X.vl[t], src);
#pragma unroll
        for (int i = 0; i < 6; i++) cdx[t][i] = shfl(K.cd[t][i], src);
      }
      q_contact_rows(CX, O.GX, K, cdx, X.vb, vlx, Sh.leg[O.GX.leg].calf_invw, -1.f, false);
    } else {
      CX.leg = 0; CX.mu = 0.f; CX.D = 0.f; CX.active = 0;
#pragma unroll
      for (int e = 0; e < 4; e++) { CX.aref[e] = 0.f; CX.jaref[e] = 0.f; CX.jv[e] = 0.f; }
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int c = 0; c < 9; c++) CX.J[i][c] = 0.f;
    }
    q_limit_rows(Lm, X, L);
    if (DBG) {
      float* o = dbg;
      if (g == 0) {
        for (int i = 0; i < 3; i++) { o[DBG_XPOS + i] = K.pb[i]; o[DBG_XIPOS + i] = xi[0][i]; o[DBG_COM + i] = K.com[i]; }
        for (int i = 0; i < 9; i++) o[DBG_XMAT + i] = K.Rb[i];
        for (int i = 0; i < 10; i++) o[DBG_CINERT + i] = K.cinb[i];
        for (int d = 0; d < 6; d++) for (int i = 0; i < 6; i++) o[DBG_CDOF + 6 * d + i] = K.cdb[d][i];
        for (int a = 0; a < 6; a++) { o[DBG_BIAS + a] = -qs.b[a]; o[DBG_QS + a] = qs.b[a]; }
        for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) o[DBG_QM + 18 * a + b] = Mm.MB[a >= b ? PK(a, b) : PK(b, a)];
      }
      for (int t = 0; t < 3; t++) {
        const int b = 1 + 3 * g + t, d = 6 + 3 * g + t;
        for (int i = 0; i < 3; i++) { o[DBG_XPOS + 3 * b + i] = K.p[t][i]; o[DBG_XIPOS + 3 * b + i] = xi[1 + t][i]; }
        for (int i = 0; i < 9; i++) o[DBG_XMAT + 9 * b + i] = K.R[t][i];
        for (int i = 0; i < 10; i++) o[DBG_CINERT + 10 * b + i] = K.cin[t][i];
        for (int i = 0; i < 6; i++) o[DBG_CDOF + 6 * d + i] = K.cd[t][i];
        o[DBG_QS + d] = qs.l[t];
        o[DBG_BIAS + d] = O.actf[t] - M.damp[t] * X.vl[t] - qs.l[t];
        for (int a = 0; a < 6; a++) { o[DBG_QM + 18 * a + d] = Mm.MC[a][t]; o[DBG_QM + 18 * d + a] = Mm.MC[a][t]; }
        for (int k = 0; k < 3; k++) o[DBG_QM + 18 * d + 6 + 3 * g + k] = Mm.MA[t >= k ? PK(t, k) : PK(k, t)];
        o[DBG_ACTF + L.act[t]] = O.actf[t];
        o[DBG_FOOT + 3 * g + t] = K.foot[t];
      }
    }
  }
  {
    QFac F;
    q_factor(F, Mm.MB, Mm.MC, Mm.MA);
    q_solve(qas, F, qs);
  }
  if (DBG) {
    float* o = dbg;
    if (g == 0) for (int a = 0; a < 6; a++) o[DBG_QAS + a] = qas.b[a];
    for (int t = 0; t < 3; t++) o[DBG_QAS + 6 + 3 * g + t] = qas.l[t];
    for (int which = 0; which < 2; which++) {
      const QGeo& G = which ? O.GX : O.GP;
      const QCon& C = which ? CX : CP;
      const int c = which ? 4 + g : g;
      float* cc = o + DBG_CONTACT + 16 * c;
      cc[0] = G.dist;
      for (int i = 0; i < 3; i++) cc[1 + i] = G.pos[i];
      for (int i = 0; i < 9; i++) cc[4 + i] = G.fr[i];
      cc[13] = G.mu; cc[14] = (float)G.leg; cc[15] = (float)G.box;
      for (int e = 0; e < 4; e++) {
        const int row = 12 + 4 * c + e;
        o[DBG_EFC_D + row] = C.active ? C.D : 0.f; o[DBG_EFC_AREF + row] = C.aref[e];
        const float f = (e & 1) ? -C.mu : C.mu;
        for (int col = 0; col < 9; col++) o[DBG_EFC_J + row * 18 + col_dof(col, C.leg)] = C.J[0][col] + C.J[1 + (e >> 1)][col] * f;
      }
    }
    for (int t = 0; t < 3; t++) {
      const int row = 3 * g + t;
      o[DBG_EFC_D + row] = Lm.D[t]; o[DBG_EFC_AREF + row] = Lm.aref[t]; o[DBG_EFC_J + row * 18 + 6 + row] = Lm.sign[t];
    }
  }
  QVec warm;
#pragma unroll
  for (int i = 0; i < 6; i++) warm.b[i] = X.wb[i];
#pragma unroll
  for (int t = 0; t < 3; t++) warm.l[t] = X.wl[t];
  O.niter = q_solve_constraints(S, CP, CX, Lm, Mm, qs, qas, warm, g, qbase);
  if (sens) q_sensors_post(Z, S.qacc);
}

// store the mjx.Data fields the task / API exposes
DEV void q_store_data(const EnvBuffers& B, int env, int g, const QState& X, const QSol& S, const QSens& Z, const QFwd& O, const LegC& L, const float* ctrl3) {
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < 7; i++) B.qpos[env * NQ + i] = X.qb[i];
#pragma unroll
    for (int i = 0; i < 6; i++) { B.qvel[env * NV + i] = X.vb[i]; B.warm[env * NV + i] = S.qacc.b[i]; B.qacc[env * NV + i] = S.qacc.b[i]; }
    float* sd = B.sensordata + (size_t)env * NSENSOR;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      sd[i] = Z.gyro[i]; sd[3 + i] = Z.acc[i]; sd[10 + i] = Z.gpos[i]; sd[13 + i] = Z.glin[i]; sd[16 + i] = Z.gang[i]; sd[19 + i] = Z.llin[i]; sd[22 + i] = Z.up[i];
      B.site_xpos[env * 15 + i] = Z.gpos[i];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) B.site_xmat[env * 9 + i] = Z.Rb[i];
#pragma unroll
    for (int i = 0; i < 4; i++) sd[6 + i] = Z.quat[i];
  }
  const int kf = g ^ 1;
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const int j = 3 * g + t;
    B.qpos[env * NQ + 7 + j] = X.ql[t]; B.qvel[env * NV + 6 + j] = X.vl[t];
    B.warm[env * NV + 6 + j] = S.qacc.l[t]; B.qacc[env * NV + 6 + j] = S.qacc.l[t];
    B.ctrl[env * NU + L.act[t]] = ctrl3[t]; B.actuator_force[env * NU + L.act[t]] = O.actf[t];
    B.sensordata[(size_t)env * NSENSOR + 25 + 3 * kf + t] = Z.fpos[t];
    B.sensordata[(size_t)env * NSENSOR + 37 + 3 * kf + t] = Z.fvel[t];
    B.site_xpos[env * 15 + 3 + j] = Z.fworld[t];   // sites FL FR RL RR = leg order
  }
  // contact list: slots 0..3 = foot g vs floor, 4..7 = box slots
  {
    B.contact_dist[env * NCON + g] = O.GP.dist;
    B.contact_geom[env * NCON * 2 + 2 * g] = GC.floor_geom; B.contact_geom[env * NCON * 2 + 2 * g + 1] = L.foot_geom;
    const int c = 4 + g, box = O.GX.box;
    B.contact_dist[env * NCON + c] = (box == -2) ? 0.f : O.GX.dist;
    B.contact_geom[env * NCON * 2 + 2 * c] = box >= 0 ? GC.foot_geom[O.GX.leg] : -1;
    B.contact_geom[env * NCON * 2 + 2 * c + 1] = box >= 0 ? GC.box_geom0 + box : -1;
  }
}

// ----------------------------------------------------------------------------------------------
// mjx_env.step of Joystick.step (joystick_pgtt.py:145-148) for the eight envs of a warp: motor targets, n_substeps x
// mjx.step; mjx.Data goes to the handle's buffers for the task kernel (pgtt_task.cuh)
// ----------------------------------------------------------------------------------------------
DEV void q_env_physics(QShared& Sh, const EnvBuffers& B, const float* action_all, int env_raw, int lane) {
  const int g = lane & 3, qbase = lane & ~3;
  const bool ok = env_raw < B.N;
  const int env = ok ? env_raw : B.N - 1;          // surplus quads shadow the last env and store nothing
  const LegC& L = Sh.leg[g];
  QModel M;
  QState X;
  q_load_model(M, B, L, env, g);
  q_load_state(X, B, env, g);
  const float* act = action_all + (size_t)env * NU;
#pragma unroll
  for (int t = 0; t < 3; t++) {   // motor target with array (actuator-order) index 3g+t; X.ctrl: of the actuator driving hinge t
    const float mtA = GC.default_pose[3 * g + t] + act[3 * g + t] * GC.action_scale;
    X.ctrl[t] = L.mt_default[t] + act[L.act[t]] * GC.action_scale;
    if (ok) B.motor_targets[env * NU + 3 * g + t] = mtA;
  }
  QFwd O;
  QSol S;
  QSens Z;
  QNear Nr;
  int niter[4] = {0, 0, 0, 0};
  const int nsub = GC.n_substeps;
#pragma unroll 1
  for (int s = 0; s < nsub; s++) {
    q_forward<false>(O, S, Z, Nr, s == 0, Sh, X, M, L, s == nsub - 1, nullptr, lane, g, qbase);
#pragma unroll
    for (int i = 0; i < 4; i++) if (i == s) niter[i] = O.niter;
    q_euler(X, S.qacc);
#pragma unroll
    for (int i = 0; i < 6; i++) X.wb[i] = S.qacc.b[i];
#pragma unroll
    for (int t = 0; t < 3; t++) X.wl[t] = S.qacc.l[t];
  }
  if (ok) {
    if (g == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++) if (i < nsub) B.solver_niter[env * 4 + i] = niter[i];
    }
    q_store_data(B, env, g, X, S, Z, O, L, X.ctrl);
  }
}

// mjx.forward with every intermediate dumped (parity probe; layout in pgtt_debug.h)
DEV void q_env_debug_forward(QShared& Sh, const EnvBuffers& B, float* out_all, int env_raw, int lane) {
  const int g = lane & 3, qbase = lane & ~3;
  const bool ok = env_raw < B.N;
  const int env = ok ? env_raw : B.N - 1;
  const LegC& L = Sh.leg[g];
  float* o = out_all + (size_t)env * PGTT_DEBUG_FLOATS;
  QModel M;
  QState X;
  q_load_model(M, B, L, env, g);
  q_load_state(X, B, env, g);
#pragma unroll
  for (int t = 0; t < 3; t++) X.ctrl[t] = B.ctrl[env * NU + L.act[t]];
  if (ok) for (int i = g; i < 44 * 18; i += 4) o[DBG_EFC_J + i] = 0.f;
  syncwarp();
  QFwd O; QSol S; QSens Z; QNear Nr;
  // surplus quads shadow the last env: they write the same values to the same record
  q_forward<true>(O, S, Z, Nr, true, Sh, X, M, L, true, o, lane, g, qbase);
  if (g == 0) {
    for (int a = 0; a < 6; a++) o[DBG_QACC + a] = S.qacc.b[a];
    for (int i = 0; i < 3; i++) {
      o[DBG_SENS + i] = Z.gyro[i]; o[DBG_SENS + 3 + i] = Z.acc[i]; o[DBG_SENS + 10 + i] = Z.gpos[i]; o[DBG_SENS + 13 + i] = Z.glin[i];
      o[DBG_SENS + 16 + i] = Z.gang[i]; o[DBG_SENS + 19 + i] = Z.llin[i]; o[DBG_SENS + 22 + i] = Z.up[i];
    }
    for (int i = 0; i < 4; i++) o[DBG_SENS + 6 + i] = Z.quat[i];
    o[DBG_NITER] = (float)O.niter;
  }
  for (int t = 0; t < 3; t++) {
    o[DBG_QACC + 6 + 3 * g + t] = S.qacc.l[t];
    o[DBG_SENS + 25 + 3 * (g ^ 1) + t] = Z.fpos[t]; o[DBG_SENS + 37 + 3 * (g ^ 1) + t] = Z.fvel[t];
  }
}
