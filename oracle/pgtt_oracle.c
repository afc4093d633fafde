/* pgtt_oracle.c - CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).
 *
 * What it restates (file:line are relative to the reference checkout):
 *   physics      mjx.step / mjx.forward as called by go2/joystick_pgtt.py:72,78,146-148
 *                (mujoco-mjx is NOT vendored in the reference; algorithm per SURVEY.md App. A)
 *   contacts     Go2Env.compute_contact, go2/base.py:153-171
 *   ray grid     go2/heightmap.py:10-67
 *   task         Joystick.reset/step/_get_obs/_get_reward/sample_command, go2/joystick_pgtt.py:50-611
 *   gait         go2/gait.py:8-49
 *   randomiser   go2/randomize.py:23-171, go2/randomize_simple.py:24-138
 *   wrappers     brax EpisodeWrapper + playground BraxAutoResetWrapper (call site training/train.py:255)
 *   rng          jax.random threefry2x32 (split / uniform / exponential / bernoulli / randint)
 *
 * PARITY UNPINNED (see pgtt_oracle.h).  Written in the generic MuJoCo form - body tree,
 * 6-D spatial vectors about the subtree COM, dense nv x nv matrices, dense nefc x nv Jacobian -
 * on purpose: the CUDA kernel uses leg-specialised arrow matrices, so agreement between the two
 * is evidence, not tautology.  Build twice: -DORC_F32 (fp32, like the reference) and fp64.
 */
#include "pgtt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_F32
#define R_(x) x##f
#else
#define R_(x) x
#endif
#define SQRT R_(sqrt)
#define SIN R_(sin)
#define COS R_(cos)
#define EXP R_(exp)
#define FABS R_(fabs)
#define FMOD R_(fmod)
#define ATAN2 R_(atan2)
#define LOG1P R_(log1p)
#define RINT R_(rint)
#define POW R_(pow)
#define MINVAL ((real)1e-15)
#define MINIMP ((real)0.0001)
#define MAXIMP ((real)0.9999)
#define PI_R ((real)3.14159265358979323846)

/* ------------------------------------------------------------------------------------------
 * small vector / quaternion helpers
 * ---------------------------------------------------------------------------------------- */
static inline real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(real* r, const real* a, const real* b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline real dot6(const real* a, const real* b) { return dot3(a, b) + dot3(a + 3, b + 3); }
static inline void mat_vec(real* r, const real* m, const real* v) { /* row-major 3x3 */
  real x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  real y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  real z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void matT_vec(real* r, const real* m, const real* v) {
  real x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  real y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  real z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mat_mul(real* r, const real* a, const real* b) {
  real t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(r, t, sizeof(t));
}
static inline void quat_mul(real* r, const real* a, const real* b) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void quat_normalize(real* q) {
  real n = SQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int i = 0; i < 4; i++) q[i] /= n;
}
static inline void quat_to_mat(real* m, const real* q) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void axis_angle_to_quat(real* q, const real* axis, real angle) {
  real s = SIN(angle * (real)0.5), c = COS(angle * (real)0.5);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
static inline real normalize3(real* v) {
  real n = SQRT(dot3(v, v));
  if (n < MINVAL) { v[0] = v[1] = v[2] = 0; return 0; } /* mjx math.normalize_with_norm: safe zero */
  v[0] /= n; v[1] /= n; v[2] /= n;
  return n;
}

/* spatial helpers: 10-number inertia (Ixx,Iyy,Izz,Ixy,Ixz,Iyz, hx,hy,hz, m) about the subtree COM */
static void inert_mul(real* res, const real* I, const real* v) {
  const real* w = v; const real* l = v + 3; const real* h = I + 6; real m = I[9];
  real hxl[3], hxw[3];
  cross3(hxl, h, l); cross3(hxw, h, w);
  res[0] = I[0] * w[0] + I[3] * w[1] + I[4] * w[2] + hxl[0];
  res[1] = I[3] * w[0] + I[1] * w[1] + I[5] * w[2] + hxl[1];
  res[2] = I[4] * w[0] + I[5] * w[1] + I[2] * w[2] + hxl[2];
  res[3] = m * l[0] - hxw[0]; res[4] = m * l[1] - hxw[1]; res[5] = m * l[2] - hxw[2];
}
static void cross_motion(real* res, const real* vel, const real* v) {
  real a[3], b[3], c[3];
  cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
  res[0] = a[0]; res[1] = a[1]; res[2] = a[2];
  res[3] = b[0] + c[0]; res[4] = b[1] + c[1]; res[5] = b[2] + c[2];
}
static void cross_force(real* res, const real* vel, const real* f) {
  real a[3], b[3], c[3];
  cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
  res[0] = a[0] + b[0]; res[1] = a[1] + b[1]; res[2] = a[2] + b[2];
  res[3] = c[0]; res[4] = c[1]; res[5] = c[2];
}

/* dof bookkeeping for this tree */
static inline int dof_body(const Model* m, int d) { return d < 6 ? 1 : m->jnt_body[d - 6]; }
static int dof_parent(const Model* m, int d) {
  if (d == 0) return -1;
  if (d < 6) return d - 1;
  int pb = m->body_parent[m->jnt_body[d - 6]];
  if (pb == 1) return 5;
  for (int j = 0; j < NHINGE; j++) if (m->jnt_body[j] == pb) return 6 + j;
  return -1;
}
static int body_hinge(const Model* m, int b) {
  for (int j = 0; j < NHINGE; j++) if (m->jnt_body[j] == b) return j;
  return -1;
}

/* ------------------------------------------------------------------------------------------
 * mjx.kinematics + com_pos + crb + factor_m      
 * ---------------------------------------------------------------------------------------- */
static void kinematics(const Model* m, Data* d) {
  memset(d->xpos[0], 0, sizeof(real) * 3);
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat_to_mat(d->xmat[0], d->xquat[0]);
  for (int b = 1; b < NBODY; b++) {
    if (b == 1) {
      for (int i = 0; i < 3; i++) d->xpos[b][i] = d->qpos[i];
      for (int i = 0; i < 4; i++) d->xquat[b][i] = d->qpos[3 + i];
      quat_normalize(d->xquat[b]);
    } else {
      int p = m->body_parent[b], j = body_hinge(m, b);
      real off[3], q[4], ql[4];
      mat_vec(off, d->xmat[p], m->body_pos[b]);
      for (int i = 0; i < 3; i++) d->xpos[b][i] = d->xpos[p][i] + off[i];
      quat_mul(q, d->xquat[p], m->body_quat[b]);
      real rm[9];
      quat_to_mat(rm, q);
      for (int i = 0; i < 3; i++) d->xanchor[j][i] = d->xpos[b][i]; /* jnt_pos == 0 */
      mat_vec(d->xaxis[j], rm, m->jnt_axis[j]);
      axis_angle_to_quat(ql, m->jnt_axis[j], d->qpos[7 + j] - m->qpos0[7 + j]);
      quat_mul(d->xquat[b], q, ql);
      quat_normalize(d->xquat[b]);
    }
    quat_to_mat(d->xmat[b], d->xquat[b]);
    real off[3], im[9];
    mat_vec(off, d->xmat[b], m->body_ipos[b]);
    for (int i = 0; i < 3; i++) d->xipos[b][i] = d->xpos[b][i] + off[i];
    quat_to_mat(im, m->body_iquat[b]);
    mat_mul(d->ximat[b], d->xmat[b], im);
  }
  /* geoms and sites that matter: 4 foot spheres (= foot sites) and the imu site */
  for (int f = 0; f < NFOOT; f++) {
    int b = m->foot_body[f];
    real off[3];
    mat_vec(off, d->xmat[b], m->foot_pos);
    for (int i = 0; i < 3; i++) d->foot_xpos[f][i] = d->site_xpos[1 + f][i] = d->xpos[b][i] + off[i];
  }
  real off[3];
  mat_vec(off, d->xmat[1], m->imu_pos);
  for (int i = 0; i < 3; i++) d->site_xpos[0][i] = d->xpos[1][i] + off[i];
  memcpy(d->site_xmat, d->xmat[1], sizeof(real) * 9);
}

static void com_pos(const Model* m, Data* d) {
  real mt = 0, c[3] = {0, 0, 0};
  for (int b = 1; b < NBODY; b++) {
    mt += m->body_mass[b];
    for (int i = 0; i < 3; i++) c[i] += m->body_mass[b] * d->xipos[b][i];
  }
  for (int i = 0; i < 3; i++) d->subtree_com[i] = c[i] / mt;
  memset(d->cinert[0], 0, sizeof(real) * 10);
  for (int b = 1; b < NBODY; b++) {
    const real* R = d->ximat[b]; const real* I = m->body_inertia[b]; real mass = m->body_mass[b];
    real o[3];
    for (int i = 0; i < 3; i++) o[i] = d->xipos[b][i] - d->subtree_com[i];
    real Iw[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        Iw[3 * i + j] = R[3 * i] * I[0] * R[3 * j] + R[3 * i + 1] * I[1] * R[3 * j + 1] + R[3 * i + 2] * I[2] * R[3 * j + 2];
    real oo = dot3(o, o);
    real* ci = d->cinert[b];
    ci[0] = Iw[0] + mass * (oo - o[0] * o[0]);
    ci[1] = Iw[4] + mass * (oo - o[1] * o[1]);
    ci[2] = Iw[8] + mass * (oo - o[2] * o[2]);
    ci[3] = Iw[1] - mass * o[0] * o[1];
    ci[4] = Iw[2] - mass * o[0] * o[2];
    ci[5] = Iw[5] - mass * o[1] * o[2];
    ci[6] = mass * o[0]; ci[7] = mass * o[1]; ci[8] = mass * o[2]; ci[9] = mass;
  }
  /* cdof: free joint = 3 world translations + 3 rotations about the body axes; hinges */
  memset(d->cdof, 0, sizeof(d->cdof));
  real off[3];
  for (int i = 0; i < 3; i++) off[i] = d->subtree_com[i] - d->xpos[1][i];
  for (int k = 0; k < 3; k++) {
    d->cdof[k][3 + k] = 1;
    real ax[3] = {d->xmat[1][k], d->xmat[1][3 + k], d->xmat[1][6 + k]};
    for (int i = 0; i < 3; i++) d->cdof[3 + k][i] = ax[i];
    cross3(d->cdof[3 + k] + 3, ax, off);
  }
  for (int j = 0; j < NHINGE; j++) {
    for (int i = 0; i < 3; i++) { d->cdof[6 + j][i] = d->xaxis[j][i]; off[i] = d->subtree_com[i] - d->xanchor[j][i]; }
    cross3(d->cdof[6 + j] + 3, d->xaxis[j], off);
  }
}

static int cholesky(real* L, const real* A, int n, int ld) { /* lower, row-major; returns rank deficiency */
  int bad = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      real s = A[i * ld + j];
      for (int k = 0; k < j; k++) s -= L[i * ld + k] * L[j * ld + k];
      if (i == j) { if (s < MINVAL) { s = MINVAL; bad++; } L[i * ld + i] = SQRT(s); }
      else L[i * ld + j] = s / L[j * ld + j];
    }
  return bad;
}
static void chol_solve(real* x, const real* L, const real* b, int n, int ld) {
  real y[NV];
  for (int i = 0; i < n; i++) { real s = b[i]; for (int k = 0; k < i; k++) s -= L[i * ld + k] * y[k]; y[i] = s / L[i * ld + i]; }
  for (int i = n - 1; i >= 0; i--) { real s = y[i]; for (int k = i + 1; k < n; k++) s -= L[k * ld + i] * x[k]; x[i] = s / L[i * ld + i]; }
}

static void crb_and_factor(const Model* m, Data* d) {
  memcpy(d->crb, d->cinert, sizeof(d->crb));
  for (int b = NBODY - 1; b > 1; b--)
    for (int k = 0; k < 10; k++) d->crb[m->body_parent[b]][k] += d->crb[b][k];
  memset(d->qM, 0, sizeof(d->qM));
  for (int i = 0; i < NV; i++) {
    real f[6];
    inert_mul(f, d->crb[dof_body(m, i)], d->cdof[i]);
    for (int j = i; j >= 0; j = dof_parent(m, j)) d->qM[i][j] = d->qM[j][i] = dot6(d->cdof[j], f);
    d->qM[i][i] += m->dof_armature[i];
  }
  cholesky(&d->qL[0][0], &d->qM[0][0], NV, NV);
}

static void mul_M(const Data* d, real* res, const real* v) {
  for (int i = 0; i < NV; i++) { real s = 0; for (int j = 0; j < NV; j++) s += d->qM[i][j] * v[j]; res[i] = s; }
}

/* jacp (3 x nv) of a world point attached to body b */
static void jac_point(const Model* m, const Data* d, real jacp[3][NV], const real* p, int b) {
  memset(jacp, 0, sizeof(real) * 3 * NV);
  real off[3];
  for (int i = 0; i < 3; i++) off[i] = p[i] - d->subtree_com[i];
  int dof = (b == 1) ? 5 : 6 + body_hinge(m, b);
  for (; dof >= 0; dof = dof_parent(m, dof)) {
    real c[3];
    cross3(c, d->cdof[dof], off);
    for (int i = 0; i < 3; i++) jacp[i][dof] = d->cdof[dof][3 + i] + c[i];
  }
}

/* ------------------------------------------------------------------------------------------
 * collision: plane-sphere x4, sphere-box 400 -> 25 -> 4    
 * ---------------------------------------------------------------------------------------- */
static void make_frame(real* frame, const real* n) {
  real a[3] = {n[0], n[1], n[2]};
  normalize3(a);
  real b[3] = {0, 0, 0};
  if (-(real)0.5 < a[1] && a[1] < (real)0.5) b[1] = 1; else b[2] = 1;
  real ab = dot3(a, b);
  for (int i = 0; i < 3; i++) b[i] -= a[i] * ab;
  normalize3(b);
  real c[3];
  cross3(c, a, b);
  for (int i = 0; i < 3; i++) { frame[i] = a[i]; frame[3 + i] = b[i]; frame[6 + i] = c[i]; }
}

/* mjx collision_convex._sphere_convex with the box as a 6-face hull. Returns dist; pos, n in world. */
static real sphere_box(const real* spos, real radius, const real* bpos, const real* bmat, const real* size,
                       real* pos, real* n) {
  real rel[3] = {spos[0] - bpos[0], spos[1] - bpos[1], spos[2] - bpos[2]}, c[3];
  matT_vec(c, bmat, rel);
  /* faces: +x,-x,+y,-y,+z,-z ; vertex order irrelevant here because faces are axis-aligned rectangles */
  int best = -1; real best_support = 0;
  for (int f = 0; f < 6; f++) {
    int ax = f / 2; real sgn = (f % 2 == 0) ? 1 : -1;
    real support = sgn * c[ax] - size[ax] - radius; /* ((c - r n) - v0) . n */
    if (support >= 0) support = (real)-1e12;         /* "minimal penetration as long as it has support" */
    if (best < 0 || support > best_support) { best = f; best_support = support; }
  }
  int ax = best / 2; real sgn = (best % 2 == 0) ? 1 : -1;
  real pt[3] = {c[0], c[1], c[2]};
  pt[ax] = sgn * size[ax]; /* project centre onto the face plane */
  /* if outside the face rectangle: closest point on the nearest violated edge == clamp the
   * coordinate with the smallest positive violation first, then clamp along the edge segment.  */
  int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
  real v1 = FABS(pt[a1]) - size[a1], v2 = FABS(pt[a2]) - size[a2];
  int inside = (v1 <= 0) && (v2 <= 0);
  if (!inside) {
    /* edge with the smallest positive distance (edges the point is behind are skipped) */
    int e;
    if (v1 > 0 && v2 > 0) e = (v1 <= v2) ? a1 : a2; else e = (v1 > 0) ? a1 : a2;
    int o = (e == a1) ? a2 : a1;
    pt[e] = (pt[e] > 0 ? 1 : -1) * size[e];
    if (pt[o] > size[o]) pt[o] = size[o];   /* closest_segment_point clamps to the edge ends */
    if (pt[o] < -size[o]) pt[o] = -size[o];
  }
  real nl[3] = {pt[0] - c[0], pt[1] - c[1], pt[2] - c[2]};
  real dn = normalize3(nl);
  real spt[3] = {c[0] + nl[0] * radius, c[1] + nl[1] * radius, c[2] + nl[2] * radius};
  real pl[3] = {(pt[0] + spt[0]) * (real)0.5, (pt[1] + spt[1]) * (real)0.5, (pt[2] + spt[2]) * (real)0.5};
  mat_vec(n, bmat, nl);
  mat_vec(pos, bmat, pl);
  for (int i = 0; i < 3; i++) pos[i] += bpos[i];
  return dn - radius;
}

static void fill_contact_params(const Model* m, Contact* c, const real* fr2, const real* solref2, const real* solimp2) {
  /* condim = max(1,3) = 3; friction = element-wise max; solref/solimp mixed 50/50 (equal solmix/priority) */
  c->mu = m->foot_friction[0] > fr2[0] ? m->foot_friction[0] : fr2[0];
  for (int i = 0; i < 2; i++) c->solref[i] = (real)0.5 * m->foot_solref[i] + (real)0.5 * solref2[i];
  for (int i = 0; i < 5; i++) c->solimp[i] = (real)0.5 * m->foot_solimp[i] + (real)0.5 * solimp2[i];
  real mg = m->foot_margin > 0 ? m->foot_margin : 0; /* max(margin1, margin2), other geom margin = 0 */
  c->includemargin = mg; /* gap = 0 */
}

static void collision(const Model* m, Data* d) {
  int nc = 0;
  /* group (PLANE, SPHERE): 4 pairs, no culling (4 <= max_geom_pairs, 4 <= max_contact_points) */
  for (int f = 0; f < NFOOT; f++) {
    Contact* c = &d->contact[nc++];
    real nrm[3] = {0, 0, 1};
    c->dist = d->foot_xpos[f][2] - m->foot_radius; /* plane at z=0, normal +z */
    for (int i = 0; i < 3; i++) c->pos[i] = d->foot_xpos[f][i] - nrm[i] * (m->foot_radius + (real)0.5 * c->dist);
    make_frame(c->frame, nrm);
    c->geom1 = m->floor_geom_id; c->geom2 = m->foot_geom_id[f];
    c->body = m->foot_body[f]; c->foot = f; c->box = -1;
    fill_contact_params(m, c, m->floor_friction, m->floor_solref, m->floor_solimp);
  }
  /* group (SPHERE, BOX): candidates ordered [foot FL,FR,RL,RR] x [box 0..n) */
  int nb = m->n_boxes, ncand = NFOOT * nb;
  if (nb > 0) {
    static _Thread_local real bs[NFOOT * NBOX];
    static _Thread_local int keep[NFOOT * NBOX];
    int nkeep;
    for (int f = 0; f < NFOOT; f++)
      for (int k = 0; k < nb; k++) {
        real dx[3] = {m->box_pos[k][0] - d->foot_xpos[f][0], m->box_pos[k][1] - d->foot_xpos[f][1], m->box_pos[k][2] - d->foot_xpos[f][2]};
        bs[f * nb + k] = SQRT(dot3(dx, dx)) - (m->foot_radius + m->box_rbound);
      }
    if (m->max_geom_pairs > -1 && ncand > m->max_geom_pairs) {
      /* top_k(-dist, k): k smallest, ties -> lowest index, result sorted ascending */
      nkeep = m->max_geom_pairs;
      static _Thread_local char used[NFOOT * NBOX];
      memset(used, 0, sizeof(used));
      for (int s = 0; s < nkeep; s++) {
        int bi = -1;
        for (int i = 0; i < ncand; i++) if (!used[i] && (bi < 0 || bs[i] < bs[bi])) bi = i;
        used[bi] = 1; keep[s] = bi;
      }
    } else { nkeep = ncand; for (int i = 0; i < ncand; i++) keep[i] = i; }
    real cd[NFOOT * NBOX], cpos[NFOOT * NBOX][3], cn[NFOOT * NBOX][3];
    for (int s = 0; s < nkeep; s++) {
      int f = keep[s] / nb, k = keep[s] % nb;
      real bq[4] = {m->box_quat[k][0], m->box_quat[k][1], m->box_quat[k][2], m->box_quat[k][3]}, bm[9];
      quat_normalize(bq);
      quat_to_mat(bm, bq);
      cd[s] = sphere_box(d->foot_xpos[f], m->foot_radius, m->box_pos[k], bm, m->box_size[k], cpos[s], cn[s]);
    }
    int nsel = nkeep;
    int sel[NFOOT * NBOX];
    if (m->max_contact_points > -1 && nkeep > m->max_contact_points) {
      nsel = m->max_contact_points;
      char used2[NFOOT * NBOX];
      memset(used2, 0, sizeof(used2));
      for (int s = 0; s < nsel; s++) {
        int bi = -1;
        for (int i = 0; i < nkeep; i++) if (!used2[i] && (bi < 0 || cd[i] < cd[bi])) bi = i;
        used2[bi] = 1; sel[s] = bi;
      }
    } else for (int i = 0; i < nkeep; i++) sel[i] = i;
    for (int s = 0; s < nsel && nc < NCON; s++) {
      int i = sel[s], f = keep[i] / nb, k = keep[i] % nb;
      Contact* c = &d->contact[nc++];
      c->dist = cd[i];
      for (int a = 0; a < 3; a++) c->pos[a] = cpos[i][a];
      make_frame(c->frame, cn[i]);
      c->geom1 = m->foot_geom_id[f]; c->geom2 = m->box_geom_id0 + k;
      c->body = m->foot_body[f]; c->foot = f; c->box = k;
      fill_contact_params(m, c, m->box_friction[k], m->box_solref, m->box_solimp);
    }
  }
  d->ncon = nc;
}

/* ------------------------------------------------------------------------------------------
 * make_constraint: 12 joint-limit rows + pyramidal contact rows    
 * ---------------------------------------------------------------------------------------- */
static void kbi(const Model* m, const real* solref, const real* solimp, real pos, real* k, real* b, real* imp) {
  real timeconst = solref[0], dampratio = solref[1];
  if (timeconst < 2 * m->timestep) timeconst = 2 * m->timestep; /* refsafe */
  real dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  dmin = dmin < MINIMP ? MINIMP : (dmin > MAXIMP ? MAXIMP : dmin);
  dmax = dmax < MINIMP ? MINIMP : (dmax > MAXIMP ? MAXIMP : dmax);
  if (width < MINVAL) width = MINVAL;
  mid = mid < MINIMP ? MINIMP : (mid > MAXIMP ? MAXIMP : mid);
  if (power < 1) power = 1;
  *k = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  *b = 2 / (dmax * timeconst);
  if (solref[0] <= 0) *k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0) *b = -solref[1] / dmax;
  real x = FABS(pos) / width;
  real ia = (1 / POW(mid, power - 1)) * POW(x, power);
  real ib = 1 - (1 / POW(1 - mid, power - 1)) * POW(1 - x, power);
  real y = x < mid ? ia : ib;
  real im = dmin + y * (dmax - dmin);
  im = im < dmin ? dmin : (im > dmax ? dmax : im);
  if (x > 1) im = dmax;
  *imp = im;
}

static void finish_row(const Model* m, Data* d, int r, real pos, real invweight, const real* solref, const real* solimp, int active) {
  if (!active) { /* every factor is multiplied by `active` in mjx: zero row, zero aref, contributes nothing */
    memset(d->efc_J[r], 0, sizeof(real) * NV);
    d->efc_pos[r] = 0; d->efc_aref[r] = 0; d->efc_D[r] = 0;
    return;
  }
  real k, b, imp;
  kbi(m, solref, solimp, pos, &k, &b, &imp);
  real R = invweight * (1 - imp) / imp;
  if (R < MINVAL) R = MINVAL;
  real vel = 0;
  for (int i = 0; i < NV; i++) vel += d->efc_J[r][i] * d->qvel[i];
  d->efc_pos[r] = pos;
  d->efc_aref[r] = -b * vel - k * imp * pos;
  d->efc_D[r] = 1 / R;
}

static void make_constraint(const Model* m, Data* d) {
  int r = 0;
  for (int j = 0; j < NHINGE; j++, r++) {
    real q = d->qpos[7 + j];
    real dmin = q - m->jnt_range[j][0], dmax = m->jnt_range[j][1] - q;
    real pos = dmin < dmax ? dmin : dmax; /* jnt_margin = 0 */
    int active = pos < 0;
    memset(d->efc_J[r], 0, sizeof(real) * NV);
    d->efc_J[r][6 + j] = (dmin < dmax) ? 1 : -1;
    finish_row(m, d, r, pos, m->dof_invweight0[6 + j], m->jnt_solref, m->jnt_solimp, active);
  }
  for (int ci = 0; ci < NCON; ci++) {
    if (ci >= d->ncon) { for (int e = 0; e < 4; e++, r++) finish_row(m, d, r, 0, 0, m->jnt_solref, m->jnt_solimp, 0); continue; }
    const Contact* c = &d->contact[ci];
    real pos = c->dist - c->includemargin;
    int active = pos < 0;
    real jp[3][NV], jc[3][NV];
    jac_point(m, d, jp, c->pos, c->body);
    /* diff = jac(body2) - jac(body1): plane contact has body2 = calf (+), box contact body1 = calf (-) */
    real sgn = (c->box < 0) ? 1 : -1;
    for (int a = 0; a < 3; a++)
      for (int i = 0; i < NV; i++)
        jc[a][i] = sgn * (c->frame[3 * a] * jp[0][i] + c->frame[3 * a + 1] * jp[1][i] + c->frame[3 * a + 2] * jp[2][i]);
    real t = m->body_invweight0[c->body][0]; /* the other body is static: invweight 0 */
    for (int e = 0; e < 4; e++, r++) {
      int tan = 1 + e / 2; real f = (e % 2 == 0) ? c->mu : -c->mu;
      for (int i = 0; i < NV; i++) d->efc_J[r][i] = jc[0][i] + jc[tan][i] * f;
      real invweight = (t + f * f * t) * 2 * f * f / m->impratio;
      finish_row(m, d, r, pos, invweight, c->solref, c->solimp, active);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * velocity, actuation, acceleration     
 * ---------------------------------------------------------------------------------------- */
static void com_vel(const Model* m, Data* d) {
  memset(d->cvel[0], 0, sizeof(real) * 6);
  memset(d->cdof_dot, 0, sizeof(d->cdof_dot));
  for (int b = 1; b < NBODY; b++) {
    real v[6];
    memcpy(v, d->cvel[m->body_parent[b]], sizeof(v));
    if (b == 1) {
      for (int k = 0; k < 3; k++) for (int i = 0; i < 6; i++) v[i] += d->cdof[k][i] * d->qvel[k]; /* cdof_dot = 0 */
      for (int k = 3; k < 6; k++) cross_motion(d->cdof_dot[k], v, d->cdof[k]);
      for (int k = 3; k < 6; k++) for (int i = 0; i < 6; i++) v[i] += d->cdof[k][i] * d->qvel[k];
    } else {
      int dd = 6 + body_hinge(m, b);
      cross_motion(d->cdof_dot[dd], v, d->cdof[dd]);
      for (int i = 0; i < 6; i++) v[i] += d->cdof[dd][i] * d->qvel[dd];
    }
    memcpy(d->cvel[b], v, sizeof(v));
  }
}

static void rne_bias(const Model* m, Data* d) {
  real cacc[NBODY][6], cfrc[NBODY][6];
  memset(cacc, 0, sizeof(cacc)); memset(cfrc, 0, sizeof(cfrc));
  for (int i = 0; i < 3; i++) cacc[0][3 + i] = -m->gravity[i];
  for (int b = 1; b < NBODY; b++) {
    memcpy(cacc[b], cacc[m->body_parent[b]], sizeof(real) * 6);
    int d0 = (b == 1) ? 0 : 6 + body_hinge(m, b), d1 = (b == 1) ? 6 : d0 + 1;
    for (int k = d0; k < d1; k++) for (int i = 0; i < 6; i++) cacc[b][i] += d->cdof_dot[k][i] * d->qvel[k];
    real Ia[6], Iv[6], cf[6];
    inert_mul(Ia, d->cinert[b], cacc[b]);
    inert_mul(Iv, d->cinert[b], d->cvel[b]);
    cross_force(cf, d->cvel[b], Iv);
    for (int i = 0; i < 6; i++) cfrc[b][i] = Ia[i] + cf[i];
  }
  for (int b = NBODY - 1; b > 1; b--) for (int i = 0; i < 6; i++) cfrc[m->body_parent[b]][i] += cfrc[b][i];
  for (int k = 0; k < NV; k++) d->qfrc_bias[k] = dot6(d->cdof[k], cfrc[dof_body(m, k)]);
}

static void fwd_actuation(const Model* m, Data* d) {
  memset(d->qfrc_actuator, 0, sizeof(d->qfrc_actuator));
  for (int a = 0; a < NU; a++) {
    real c = d->ctrl[a];
    if (c < m->act_ctrlrange[a][0]) c = m->act_ctrlrange[a][0];
    if (c > m->act_ctrlrange[a][1]) c = m->act_ctrlrange[a][1];
    int dof = m->act_dof[a];
    real f = m->act_gain[a] * c + m->act_bias[a][0] + m->act_bias[a][1] * d->qpos[dof + 1] + m->act_bias[a][2] * d->qvel[dof];
    if (f < m->act_forcerange[a][0]) f = m->act_forcerange[a][0];
    if (f > m->act_forcerange[a][1]) f = m->act_forcerange[a][1];
    d->actuator_force[a] = f;
    d->qfrc_actuator[dof] += f;
  }
}

/* ------------------------------------------------------------------------------------------
 * Newton solver with mjx's bracketed line search    
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  real qacc[NV], Ma[NV], Jaref[NEFC], grad[NV], Mgrad[NV], search[NV], qfrc_constraint[NV], efc_force[NEFC];
  real gauss, cost, prev_cost;
  int active[NEFC], niter;
} SolCtx;

static void ctx_update_constraint(const Data* d, SolCtx* c) {
  real cost = 0;
  for (int r = 0; r < NEFC; r++) {
    c->active[r] = c->Jaref[r] < 0;
    c->efc_force[r] = c->active[r] ? d->efc_D[r] * -c->Jaref[r] : 0;
    if (c->active[r]) cost += d->efc_D[r] * c->Jaref[r] * c->Jaref[r];
  }
  for (int i = 0; i < NV; i++) { real s = 0; for (int r = 0; r < NEFC; r++) s += d->efc_J[r][i] * c->efc_force[r]; c->qfrc_constraint[i] = s; }
  real g = 0;
  for (int i = 0; i < NV; i++) g += (c->Ma[i] - d->qfrc_smooth[i]) * (c->qacc[i] - d->qacc_smooth[i]);
  c->gauss = (real)0.5 * g;
  c->prev_cost = c->cost;
  c->cost = (real)0.5 * cost + c->gauss;
}

static void ctx_update_gradient(const Data* d, SolCtx* c) {
  real H[NV][NV], L[NV][NV];
  for (int i = 0; i < NV; i++) c->grad[i] = c->Ma[i] - d->qfrc_smooth[i] - c->qfrc_constraint[i];
  memcpy(H, d->qM, sizeof(H));
  for (int r = 0; r < NEFC; r++) if (c->active[r]) {
    real w = d->efc_D[r];
    for (int i = 0; i < NV; i++) { real ji = d->efc_J[r][i]; if (ji == 0) continue; for (int j = 0; j < NV; j++) H[i][j] += w * ji * d->efc_J[r][j]; }
  }
  cholesky(&L[0][0], &H[0][0], NV, NV);
  chol_solve(c->Mgrad, &L[0][0], c->grad, NV, NV);
}

static void ctx_create(const Data* d, SolCtx* c, const real* qacc, int grad) {
  memcpy(c->qacc, qacc, sizeof(real) * NV);
  for (int r = 0; r < NEFC; r++) { real s = 0; for (int i = 0; i < NV; i++) s += d->efc_J[r][i] * qacc[i]; c->Jaref[r] = s - d->efc_aref[r]; }
  mul_M(d, c->Ma, qacc);
  memset(c->grad, 0, sizeof(c->grad)); memset(c->Mgrad, 0, sizeof(c->Mgrad)); memset(c->search, 0, sizeof(c->search));
  c->gauss = 0; c->cost = INFINITY; c->prev_cost = 0; c->niter = 0;
  ctx_update_constraint(d, c);
  if (grad) { ctx_update_gradient(d, c); for (int i = 0; i < NV; i++) c->search[i] = -c->Mgrad[i]; }
}

typedef struct { real alpha, cost, deriv_0, deriv_1; } LSPoint;

static LSPoint ls_point(const SolCtx* c, real alpha, const real* jv, const real quad[][3], const real* quad_gauss) {
  real q0 = quad_gauss[0], q1 = quad_gauss[1], q2 = quad_gauss[2];
  for (int r = 0; r < NEFC; r++) {
    real x = c->Jaref[r] + alpha * jv[r];
    if (x < 0) { q0 += quad[r][0]; q1 += quad[r][1]; q2 += quad[r][2]; }
  }
  LSPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.deriv_0 = 2 * alpha * q2 + q1;
  p.deriv_1 = 2 * q2 + (q2 == 0 ? MINVAL : 0);
  return p;
}

static void linesearch(const Model* m, const Data* d, SolCtx* c) {
  real snorm = 0;
  for (int i = 0; i < NV; i++) snorm += c->search[i] * c->search[i];
  real smag = SQRT(snorm) * m->meaninertia * NV;
  real gtol = m->tolerance * m->ls_tolerance * smag;
  real mv[NV], jv[NEFC], quad[NEFC][3], qg[3];
  mul_M(d, mv, c->search);
  for (int r = 0; r < NEFC; r++) { real s = 0; for (int i = 0; i < NV; i++) s += d->efc_J[r][i] * c->search[i]; jv[r] = s; }
  real sMa = 0, sq = 0, sMv = 0;
  for (int i = 0; i < NV; i++) { sMa += c->search[i] * c->Ma[i]; sq += c->search[i] * d->qfrc_smooth[i]; sMv += c->search[i] * mv[i]; }
  qg[0] = c->gauss; qg[1] = sMa - sq; qg[2] = (real)0.5 * sMv;
  for (int r = 0; r < NEFC; r++) {
    quad[r][0] = (real)0.5 * c->Jaref[r] * c->Jaref[r] * d->efc_D[r];
    quad[r][1] = jv[r] * c->Jaref[r] * d->efc_D[r];
    quad[r][2] = (real)0.5 * jv[r] * jv[r] * d->efc_D[r];
  }
  LSPoint p0 = ls_point(c, 0, jv, quad, qg);
  LSPoint lo0 = ls_point(c, p0.alpha - p0.deriv_0 / p0.deriv_1, jv, quad, qg);
  int lesser = lo0.deriv_0 < p0.deriv_0;
  LSPoint hi = lesser ? p0 : lo0, lo = lesser ? lo0 : p0;
  int swap = 1, it = 0;
  for (;;) {
    int done = it >= m->ls_iterations;
    done |= (!swap) && (it > 0);
    done |= (lo.deriv_0 < 0) && (lo.deriv_0 > -gtol);
    done |= (hi.deriv_0 > 0) && (hi.deriv_0 < gtol);
    if (done) break;
    LSPoint lo_next = ls_point(c, lo.alpha - lo.deriv_0 / lo.deriv_1, jv, quad, qg);
    LSPoint hi_next = ls_point(c, hi.alpha - hi.deriv_0 / hi.deriv_1, jv, quad, qg);
    LSPoint mid = ls_point(c, (real)0.5 * (lo.alpha + hi.alpha), jv, quad, qg);
    int swap_lo_next = (lo.deriv_0 > 0) | (lo.deriv_0 < lo_next.deriv_0);
    if (swap_lo_next) lo = lo_next;
    int swap_lo_mid = (mid.deriv_0 < 0) & (lo.deriv_0 < mid.deriv_0);
    if (swap_lo_mid) lo = mid;
    int swap_hi_next = (hi.deriv_0 < 0) | (hi.deriv_0 > hi_next.deriv_0);
    if (swap_hi_next) hi = hi_next;
    int swap_hi_mid = (mid.deriv_0 > 0) & (hi.deriv_0 > mid.deriv_0);
    if (swap_hi_mid) hi = mid;
    swap = swap_lo_next | swap_lo_mid | swap_hi_next | swap_hi_mid;
    it++;
  }
  int improved = (lo.cost < p0.cost) | (hi.cost < p0.cost);
  real alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (improved) {
    for (int i = 0; i < NV; i++) { c->qacc[i] += c->search[i] * alpha; c->Ma[i] += mv[i] * alpha; }
    for (int r = 0; r < NEFC; r++) c->Jaref[r] += jv[r] * alpha;
  }
}

static void solve(const Model* m, Data* d) {
  SolCtx warm, smth, ctx;
  ctx_create(d, &warm, d->qacc_warmstart, 0);
  ctx_create(d, &smth, d->qacc_smooth, 0);
  const real* q0 = warm.cost < smth.cost ? d->qacc_warmstart : d->qacc_smooth;
  ctx_create(d, &ctx, q0, 1);
  real scale = m->meaninertia * NV;
  for (;;) {
    real improvement = (ctx.prev_cost - ctx.cost) / scale;
    real gn = 0;
    for (int i = 0; i < NV; i++) gn += ctx.grad[i] * ctx.grad[i];
    real gradient = SQRT(gn) / scale;
    int done = ctx.niter >= m->iterations;
    done |= improvement < m->tolerance;
    done |= gradient < m->tolerance;
    if (done && m->iterations != 1) break;
    linesearch(m, d, &ctx);
    ctx_update_constraint(d, &ctx);
    ctx_update_gradient(d, &ctx);
    for (int i = 0; i < NV; i++) ctx.search[i] = -ctx.Mgrad[i];
    ctx.niter++;
    if (m->iterations == 1) break;
  }
  memcpy(d->qacc, ctx.qacc, sizeof(real) * NV);
  memcpy(d->qacc_warmstart, ctx.qacc, sizeof(real) * NV);
  memcpy(d->qfrc_constraint, ctx.qfrc_constraint, sizeof(real) * NV);
  memcpy(d->efc_force, ctx.efc_force, sizeof(real) * NEFC);
  d->solver_niter = ctx.niter;
}

/* ------------------------------------------------------------------------------------------
 * sensors (go2_mjx_feetonly.xml:258-274)     
 * ---------------------------------------------------------------------------------------- */
static void site_linvel(const Data* d, real* out, const real* cvel, const real* spos) {
  real off[3] = {spos[0] - d->subtree_com[0], spos[1] - d->subtree_com[1], spos[2] - d->subtree_com[2]}, c[3];
  cross3(c, cvel, off); /* w x r */
  for (int i = 0; i < 3; i++) out[i] = cvel[3 + i] + c[i];
}

static void sensors(const Model* m, Data* d) {
  real* s = d->sensordata;
  const real* R = d->site_xmat; const real* imu = d->site_xpos[0]; const real* cv = d->cvel[1];
  real lin[3];
  matT_vec(s + 0, R, cv);                                   /* gyro */
  memcpy(s + 6, d->xquat[1], sizeof(real) * 4);             /* framequat (site quat = identity) */
  memcpy(s + 10, imu, sizeof(real) * 3);                    /* framepos */
  site_linvel(d, lin, cv, imu);
  memcpy(s + 13, lin, sizeof(real) * 3);                    /* framelinvel (world) */
  memcpy(s + 16, cv, sizeof(real) * 3);                     /* frameangvel (world) */
  matT_vec(s + 19, R, lin);                                 /* velocimeter */
  s[22] = R[2]; s[23] = R[5]; s[24] = R[8];                 /* framezaxis */
  static const int foot_of_sensor[4] = {1, 0, 3, 2};        /* sensors FR FL RR RL -> feet FL FR RL RR */
  for (int k = 0; k < 4; k++) {
    int f = foot_of_sensor[k];
    real rel[3] = {d->foot_xpos[f][0] - imu[0], d->foot_xpos[f][1] - imu[1], d->foot_xpos[f][2] - imu[2]};
    matT_vec(s + 25 + 3 * k, R, rel);                       /* framepos relative to imu site */
    site_linvel(d, s + 37 + 3 * k, d->cvel[m->foot_body[f]], d->foot_xpos[f]);
  }
  /* accelerometer: rne_postconstraint cacc of the base + classical-acceleration correction */
  real cacc[6] = {0, 0, 0, -m->gravity[0], -m->gravity[1], -m->gravity[2]};
  for (int k = 0; k < 6; k++) for (int i = 0; i < 6; i++) cacc[i] += d->cdof_dot[k][i] * d->qvel[k] + d->cdof[k][i] * d->qacc[k];
  memcpy(d->cacc[1], cacc, sizeof(cacc));
  real off[3] = {imu[0] - d->subtree_com[0], imu[1] - d->subtree_com[1], imu[2] - d->subtree_com[2]};
  real c1[3], acc_w[3], ang_l[3], lin_l[3], acc_l[3], corr[3];
  cross3(c1, cacc, off);
  for (int i = 0; i < 3; i++) acc_w[i] = cacc[3 + i] + c1[i];
  matT_vec(acc_l, R, acc_w); matT_vec(ang_l, R, cv); matT_vec(lin_l, R, lin);
  cross3(corr, ang_l, lin_l);
  for (int i = 0; i < 3; i++) s[3 + i] = acc_l[i] + corr[i];
}

/* ------------------------------------------------------------------------------------------
 * forward / step
 * ---------------------------------------------------------------------------------------- */
void orc_forward(const Model* m, Data* d) {
  kinematics(m, d); com_pos(m, d); crb_and_factor(m, d); collision(m, d); make_constraint(m, d);
  com_vel(m, d);
  for (int i = 0; i < NV; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  rne_bias(m, d);
  fwd_actuation(m, d);
  for (int i = 0; i < NV; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  chol_solve(d->qacc_smooth, &d->qL[0][0], d->qfrc_smooth, NV, NV);
  solve(m, d);
  sensors(m, d);
}

static void euler(const Model* m, Data* d) {
  real dt = m->timestep;
  for (int i = 0; i < NV; i++) d->qvel[i] += dt * d->qacc[i];
  for (int i = 0; i < 3; i++) d->qpos[i] += dt * d->qvel[i];
  real w[3] = {d->qvel[3], d->qvel[4], d->qvel[5]};
  real n = normalize3(w);
  real dq[4], q[4];
  axis_angle_to_quat(dq, w, dt * n);
  quat_mul(q, d->qpos + 3, dq);
  quat_normalize(q);
  memcpy(d->qpos + 3, q, sizeof(q));
  for (int j = 0; j < NHINGE; j++) d->qpos[7 + j] += dt * d->qvel[6 + j];
  d->time += dt;
}

void orc_step(const Model* m, Data* d) { orc_forward(m, d); euler(m, d); }

/* ------------------------------------------------------------------------------------------
 * mjx.ray restricted to group-0 geoms = floor plane + boxes (go2/heightmap.py:10-23)   App. A10
 * ---------------------------------------------------------------------------------------- */
static real ray_scene(const Model* m, const real* pnt, const real* vec) {
  real best = INFINITY;
  /* plane z = 0, normal +z: hit iff moving towards it from above */
  if (vec[2] < 0 && pnt[2] >= 0) { real x = -pnt[2] / vec[2]; if (x >= 0 && x < best) best = x; }
  for (int k = 0; k < m->n_boxes; k++) {
    real bq[4] = {m->box_quat[k][0], m->box_quat[k][1], m->box_quat[k][2], m->box_quat[k][3]}, bm[9];
    quat_normalize(bq); quat_to_mat(bm, bq);
    real rel[3] = {pnt[0] - m->box_pos[k][0], pnt[1] - m->box_pos[k][1], pnt[2] - m->box_pos[k][2]}, lp[3], lv[3];
    matT_vec(lp, bm, rel); matT_vec(lv, bm, vec);
    const real* sz = m->box_size[k];
    for (int f = 0; f < 6; f++) {
      int ax = f % 3; real sg = f < 3 ? 1 : -1;
      if (lv[ax] == 0) continue;
      real x = (sg * sz[ax] - lp[ax]) / lv[ax];
      if (!(x >= 0)) continue;
      int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
      real p1 = lp[a1] + x * lv[a1], p2 = lp[a2] + x * lv[a2];
      if (FABS(p1) <= sz[a1] && FABS(p2) <= sz[a2] && x < best) best = x;
    }
  }
  return best;
}

/* create_sensor_matrix(mx, dx, center, yaw): hit points [13*9][3], row-major (heightmap.py:25-67) */
void orc_heightscan(const Model* m, const real* center, real yaw, real out[NRAY][3]) {
  real c = COS(yaw), s = SIN(yaw);
  real ref[3] = {center[0], center[1], center[2] + (real)0.6};
  real dir[3] = {0, 0, -1};
  real ch = (NRAY_H - 1) / (real)2, cw = (NRAY_W - 1) / (real)2;
  for (int i = 0; i < NRAY_H; i++)
    for (int j = 0; j < NRAY_W; j++) {
      real p = (ch - i) * (real)0.1, k = (cw - j) * (real)0.1;
      /* offsets @ [[c, s], [-s, c]] */
      real o[3] = {ref[0] + (p * c - k * s), ref[1] + (p * s + k * c), ref[2]};
      if (i == (NRAY_H - 1) / 2 && j == (NRAY_W - 1) / 2) { o[0] = ref[0]; o[1] = ref[1]; }
      real dist = ray_scene(m, o, dir);
      for (int a = 0; a < 3; a++) out[i * NRAY_W + j][a] = o[a] + dir[a] * dist;
    }
}

/* ------------------------------------------------------------------------------------------
 * jax.random (threefry2x32)     
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
void orc_threefry2x32(const uint32_t key[2], uint32_t x0, uint32_t x1, uint32_t out[2]) {
  static const int R[8] = {13, 15, 26, 6, 17, 29, 16, 24};
  uint32_t ks[3] = {key[0], key[1], key[0] ^ key[1] ^ 0x1BD11BDAu};
  x0 += ks[0]; x1 += ks[1];
  for (int g = 0; g < 5; g++) {
    const int* rot = (g % 2 == 0) ? R : R + 4;
    for (int i = 0; i < 4; i++) { x0 += x1; x1 = rotl32(x1, rot[i]); x1 ^= x0; }
    x0 += ks[(g + 1) % 3]; x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  out[0] = x0; out[1] = x1;
}
/* split(key, num)[i] */
static void rng_split(int part, const uint32_t key[2], int num, int i, uint32_t out[2]) {
  if (part) { orc_threefry2x32(key, 0, (uint32_t)i, out); return; }
  /* original: counts = iota(2*num) split in halves x0 = counts[:num], x1 = counts[num:]; out = concat(y0, y1).reshape(num, 2) */
  uint32_t o[2];
  for (int w = 0; w < 2; w++) {
    int flat = 2 * i + w; /* index into concat(y0, y1) */
    int pair = flat % num, which = flat / num;
    orc_threefry2x32(key, (uint32_t)pair, (uint32_t)(pair + num), o);
    out[w] = o[which];
  }
}
/* random_bits(key, 32, (n,))[i] */
static uint32_t rng_bits(int part, const uint32_t key[2], int n, int i) {
  uint32_t o[2];
  if (part) { orc_threefry2x32(key, 0, (uint32_t)i, o); return o[0] ^ o[1]; }
  int half = (n + 1) / 2;
  int pair = i % half, which = i / half;
  uint32_t x1 = (uint32_t)(pair + half);
  if ((n & 1) && pair + half >= n) x1 = 0; /* zero padding of odd-sized counts */
  orc_threefry2x32(key, (uint32_t)pair, x1, o);
  return o[which];
}
static inline real bits_to_unit(uint32_t b) { /* [0,1): mantissa trick, always in fp32 like jax */
  union { uint32_t u; float f; } v;
  v.u = (b >> 9) | 0x3F800000u;
  return (real)(v.f - 1.0f);
}
static real rng_uniform(int part, const uint32_t key[2], int n, int i, real lo, real hi) {
  real u = bits_to_unit(rng_bits(part, key, n, i));
#ifdef ORC_F32
  real v = u * (hi - lo) + lo;
#else
  real v = u * (hi - lo) + lo;
#endif
  return v < lo ? lo : v;
}
static int rng_randint(int part, const uint32_t key[2], int lo, int hi) { /* shape (1,) */
  uint32_t k1[2], k2[2];
  rng_split(part, key, 2, 0, k1); rng_split(part, key, 2, 1, k2);
  uint32_t hb = rng_bits(part, k1, 1, 0), lb = rng_bits(part, k2, 1, 0);
  uint32_t span = (uint32_t)(hi - lo);
  if (hi <= lo) span = 1;
  uint32_t mult = ((65536u % span) * (65536u % span)) % span;
  uint32_t off = ((hb % span) * mult + (lb % span)) % span;
  return lo + (int)off;
}

/* ------------------------------------------------------------------------------------------
 * task: go2/joystick_pgtt.py
 * ---------------------------------------------------------------------------------------- */
static real quat_to_yaw(const real* q) { /* go2/utility.py:4-8: scipy as_euler('xyz')[2] */
  real w = q[0], x = q[1], y = q[2], z = q[3];
  return ATAN2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z));
}

static real gait_get_z(real phi, real h_max, real stance) { /* go2/gait.py:27-49, p_stance = 0.5 */
  real T_swing = 2 * PI_R * (1 - (real)0.5) / 2, T_peak = 2 * PI_R * (1 + (real)0.5) / 2, T_stance = 2 * PI_R * (real)0.5;
  if (phi <= T_stance) return stance;
  real p0, p1, t;
  if (phi <= T_peak) { p0 = stance; p1 = h_max; t = (phi - T_stance) / T_swing; }
  else { p0 = h_max; p1 = stance; t = (phi - T_peak) / T_swing; }
  real t2 = t * t, t3 = t2 * t;
  real h00 = 2 * t3 - 3 * t2 + 1, h01 = -2 * t3 + 3 * t2;
  return h00 * p0 + h01 * p1; /* tangents are zero */
}

static void split2(int part, uint32_t rng[2], uint32_t key[2]) { /* rng, key = split(rng) */
  uint32_t a[2], b[2];
  rng_split(part, rng, 2, 0, a); rng_split(part, rng, 2, 1, b);
  rng[0] = a[0]; rng[1] = a[1]; key[0] = b[0]; key[1] = b[1];
}

static void compute_contact_flags(const Env* e, int flags[4]) { /* base.py:153-171, order FR FL RR RL */
  static const int foot_of_flag[4] = {1, 0, 3, 2};
  for (int k = 0; k < 4; k++) {
    flags[k] = 0;
    for (int c = 0; c < e->d.ncon; c++)
      if (e->d.contact[c].foot == foot_of_flag[k] && e->d.contact[c].dist < 0) flags[k] = 1;
  }
}

static void get_obs(const TaskCfg* t, Env* e) { /* joystick_pgtt.py:238-370 */
  Data* d = &e->d; Info* in = &e->info; int part = t->rng_partitionable;
  uint32_t key[2];
  real* o = e->obs_state;
  real lvl = t->noise_level;
  split2(part, in->rng, key);
  for (int i = 0; i < 3; i++) o[i] = d->sensordata[i] + (2 * rng_uniform(part, key, 3, i, 0, 1) - 1) * lvl * t->noise_gyro;
  real g[3], down[3] = {0, 0, -1};
  matT_vec(g, d->site_xmat, down);
  split2(part, in->rng, key);
  for (int i = 0; i < 3; i++) o[3 + i] = g[i] + (2 * rng_uniform(part, key, 3, i, 0, 1) - 1) * lvl * t->noise_gravity;
  split2(part, in->rng, key);
  for (int i = 0; i < 12; i++)
    o[6 + i] = (d->qpos[7 + i] + (2 * rng_uniform(part, key, 12, i, 0, 1) - 1) * lvl * t->noise_joint_pos) - t->default_pose[i];
  split2(part, in->rng, key);
  for (int i = 0; i < 12; i++) o[18 + i] = d->qvel[6 + i] + (2 * rng_uniform(part, key, 12, i, 0, 1) - 1) * lvl * t->noise_joint_vel;
  split2(part, in->rng, key); /* linvel noise key, re-used for the heightscan */
  /* baseline task (go2/joystick.py:333-341): no phase block, no gait_freq -> 162 / 206 entries */
  const int bv = t->variant != 0, o_scan = bv ? 30 : 38, o_last = bv ? 147 : 156, o_cmd = bv ? 159 : 168, nobs = bv ? NOBS - 9 : NOBS;
  if (!bv) for (int i = 0; i < 4; i++) { o[30 + i] = COS(in->phase[i]); o[34 + i] = SIN(in->phase[i]); }
  real zmin = INFINITY;
  for (int i = 0; i < NRAY; i++) if (in->heightscan[i][2] < zmin) zmin = in->heightscan[i][2];
  for (int i = 0; i < NRAY; i++)
    o[o_scan + i] = (in->heightscan[i][2] - zmin) + (2 * rng_uniform(part, key, NRAY, i, 0, 1) - 1) * lvl * t->noise_heightscan;
  if (in->step % t->history_update_steps == 0) {
    memmove(in->qvel_history + 12, in->qvel_history, sizeof(real) * 12);
    memmove(in->qpos_error_history + 12, in->qpos_error_history, sizeof(real) * 12);
    for (int i = 0; i < 12; i++) { in->qvel_history[i] = d->qvel[6 + i]; in->qpos_error_history[i] = d->qpos[7 + i] - in->motor_targets[i]; }
  }
  if (!bv) o[155] = in->gait_freq;
  for (int i = 0; i < 12; i++) o[o_last + i] = in->last_act[i];
  for (int i = 0; i < 3; i++) o[o_cmd + i] = in->command[i];
  for (int i = nobs; i < NOBS; i++) o[i] = 0;
  real* p = e->obs_priv;
  memcpy(p, o, sizeof(real) * nobs);
  real* x = p + nobs;
  for (int i = 0; i < 3; i++) { x[i] = d->sensordata[19 + i]; x[3 + i] = d->sensordata[3 + i]; x[6 + i] = d->sensordata[16 + i]; }
  for (int i = 0; i < 12; i++) x[9 + i] = d->actuator_force[i];
  for (int i = 0; i < 4; i++) x[21 + i] = (real)in->last_contact[i];
  for (int i = 0; i < 12; i++) x[25 + i] = d->sensordata[37 + i];
  for (int i = 0; i < 4; i++) x[37 + i] = in->feet_air_time[i];
  x[41] = x[42] = x[43] = 0; /* xfrc_applied[torso,:3] is never written in this fork */
  for (int i = nobs + 44; i < NPRIV; i++) p[i] = 0;
}

static void quadrant_stats(const TaskCfg* t, Info* in) { /* joystick_pgtt.py:169-190, n = 6 for both axes */
  const int n = (NRAY_H - 1) / 2;
  const int r0[4] = {0, 0, n + 1, n + 1}, r1[4] = {n, n, NRAY_H, NRAY_H};
  const int c0[4] = {n + 1, 0, n + 1, 0}, c1[4] = {NRAY_W, n, NRAY_W, n};
  for (int q = 0; q < 4; q++) {
    real mx = -INFINITY, mn = INFINITY;
    for (int i = r0[q]; i < r1[q]; i++)
      for (int j = c0[q]; j < c1[q]; j++) { real z = in->heightscan[i * NRAY_W + j][2]; if (z > mx) mx = z; if (z < mn) mn = z; }
    in->H_max[q] = t->variant ? mx : mx - mn; in->H_min[q] = mn; /* joystick.py:186 vs joystick_pgtt.py:189 */
  }
}

void orc_env_reset(const TaskCfg* t, Env* e, const uint32_t rng_in[2]) { /* joystick_pgtt.py:50-131 */
  const Model* m = &e->m; Data* d = &e->d; Info* in = &e->info; int part = t->rng_partitionable;
  uint32_t rng[2] = {rng_in[0], rng_in[1]}, key[2], key1[2], key2[2];
  memset(d, 0, sizeof(Data)); memset(in, 0, sizeof(Info));
  memcpy(d->qpos, t->home_qpos, sizeof(real) * NQ);
  split2(part, rng, key);
  for (int i = 0; i < 2; i++) d->qpos[i] += rng_uniform(part, key, 2, i, (real)-0.5, (real)0.5);
  split2(part, rng, key);
  real yaw = rng_uniform(part, key, 1, 0, (real)-3.14, (real)3.14);
  real zax[3] = {0, 0, 1}, qy[4], q[4];
  axis_angle_to_quat(qy, zax, yaw);
  quat_mul(q, d->qpos + 3, qy);
  memcpy(d->qpos + 3, q, sizeof(q));
  split2(part, rng, key);
  for (int i = 0; i < 6; i++) d->qvel[i] = rng_uniform(part, key, 6, i, (real)-0.1, (real)0.1);
  memcpy(d->ctrl, d->qpos + 7, sizeof(real) * NU);
  orc_forward(m, d);                                    /* mjx_env.init */
  orc_heightscan(m, d->qpos, 0, in->heightscan);        /* yaw = 0 */
  real zmax = -INFINITY;
  for (int i = 0; i < NRAY; i++) if (in->heightscan[i][2] > zmax) zmax = in->heightscan[i][2];
  d->qpos[2] += zmax;
  orc_forward(m, d);
  /* rng, key1, key2 = split(rng, 3) */
  { uint32_t a[2]; rng_split(part, rng, 3, 1, key1); rng_split(part, rng, 3, 2, key2); rng_split(part, rng, 3, 0, a); rng[0] = a[0]; rng[1] = a[1]; }
  real tcmd = -LOG1P(-rng_uniform(part, key1, 1, 0, 0, 1)) * (real)5.0;
  in->steps_until_next_cmd = (int)RINT(tcmd / t->ctrl_dt);
  for (int i = 0; i < 3; i++) in->command[i] = rng_uniform(part, key2, 3, i, t->cmd_u_min[i], t->cmd_u_max[i]);
  split2(part, rng, key);
  in->gait_freq = rng_uniform(part, key, 1, 0, t->gait_freq[0], t->gait_freq[1]);
  orc_heightscan(m, d->qpos, 0, in->heightscan);
  in->rng[0] = rng[0]; in->rng[1] = rng[1];
  in->step = 0;
  in->phase[0] = 0; in->phase[1] = PI_R; in->phase[2] = PI_R; in->phase[3] = 0;
  in->phase_dt = 2 * PI_R * t->ctrl_dt * in->gait_freq;
  for (int i = 0; i < 4; i++) { in->H_max[i] = (real)0.1; in->H_min[i] = 0; }
  memset(e->metrics, 0, sizeof(e->metrics));
  get_obs(t, e);
  e->reward = 0; e->done = 0;
  /* wrappers */
  in->steps = 0; in->truncation = 0; in->episode_done = 0;
  memset(in->episode_metrics, 0, sizeof(in->episode_metrics));
  memcpy(&e->first_data, d, sizeof(Data));
  memcpy(e->first_obs_state, e->obs_state, sizeof(e->obs_state));
  memcpy(e->first_obs_priv, e->obs_priv, sizeof(e->obs_priv));
  memset(e->contact_flags, 0, sizeof(e->contact_flags)); memset(e->first_contact, 0, sizeof(e->first_contact));
}

/* Joystick.step (joystick_pgtt.py:141-231) without the wrappers */
void orc_task_step(const TaskCfg* t, Env* e, const real* action) {
  const Model* m = &e->m; Data* d = &e->d; Info* in = &e->info; int part = t->rng_partitionable;
  real dt = t->ctrl_dt;
  for (int i = 0; i < NU; i++) { in->motor_targets[i] = t->default_pose[i] + action[i] * t->action_scale; }
  for (int s = 0; s < t->n_substeps; s++) { memcpy(d->ctrl, in->motor_targets, sizeof(real) * NU); orc_step(m, d); }
  int contact[4], first_contact[4];
  compute_contact_flags(e, contact);
  for (int i = 0; i < 4; i++) {
    int filt = contact[i] | in->last_contact[i];
    first_contact[i] = (in->feet_air_time[i] > 0) * filt;
    in->feet_air_time[i] += dt;
  }
  const real* s = d->sensordata;
  real pfz[4];
  for (int i = 0; i < 4; i++) { pfz[i] = s[25 + 3 * i + 2]; if (pfz[i] > in->swing_peak[i]) in->swing_peak[i] = pfz[i]; }
  orc_heightscan(m, d->qpos, quat_to_yaw(d->qpos + 3), in->heightscan);
  quadrant_stats(t, in);
  get_obs(t, e);
  int done = s[24] < 0; /* upvector z */
  /* failure guard shared with the kernels (not in the reference): a non-finite or absurd state ends the episode */
  for (int i = 0; i < NQ; i++) if (!(FABS(d->qpos[i]) < (real)1e6)) done = 1;
  for (int i = 0; i < NV; i++) if (!(FABS(d->qvel[i]) < (real)1e6)) done = 1;
  /* rewards, in the dict order of _get_reward (joystick_pgtt.py:382-420) */
  const real* cmd = in->command; const real* q = d->qpos + 7; const real* qd = d->qvel + 6; const real* af = d->actuator_force;
  real cmd_norm = SQRT(dot3(cmd, cmd));
  real rw[NREW]; /* indexed in config order */
  enum { TRACK_LIN, TRACK_ANG, LIN_VEL_Z, ANG_VEL_XY, ORIENT, DOF_LIM, POSE, TERMINATION, STAND_STILL, TORQUES, ACTION_RATE,
         ENERGY, FEET_CLEAR, FEET_HEIGHT, FEET_SLIP, FEET_AIR, FEET_PHASE, FEET_SWING, BODY_HEIGHT, CONTACT, CENTER };
  real le = (cmd[0] - s[19]) * (cmd[0] - s[19]) + (cmd[1] - s[20]) * (cmd[1] - s[20]);
  rw[TRACK_LIN] = EXP(-le / t->tracking_sigma);
  rw[TRACK_ANG] = EXP(-((cmd[2] - s[2]) * (cmd[2] - s[2])) / t->tracking_sigma);
  rw[LIN_VEL_Z] = s[15] * s[15];
  rw[ANG_VEL_XY] = s[16] * s[16] + s[17] * s[17];
  rw[ORIENT] = s[22] * s[22] + s[23] * s[23];
  real ss = 0, pose = 0, lim = 0, t2 = 0, t1 = 0, ar = 0, en = 0;
  for (int i = 0; i < 12; i++) {
    real dq = q[i] - t->default_pose[i];
    ss += FABS(dq);
    pose += dq * dq * ((i % 3 == 0) ? (real)1.0 : (real)0.1);
    real lo = m->jnt_range[i][0] * t->soft_limit_factor, hi = m->jnt_range[i][1] * t->soft_limit_factor;
    real a = q[i] - lo, b = q[i] - hi;
    lim += -(a < 0 ? a : 0) + (b > 0 ? b : 0);
    t2 += af[i] * af[i]; t1 += FABS(af[i]);
    ar += (action[i] - in->last_act[i]) * (action[i] - in->last_act[i]);
    en += FABS(qd[i]) * FABS(af[i]); /* joint order x actuator order, as in the reference */
  }
  rw[STAND_STILL] = ss * (cmd_norm < (real)0.01);
  rw[TERMINATION] = (real)done;
  rw[POSE] = pose;
  rw[TORQUES] = SQRT(t2) + t1;
  rw[ACTION_RATE] = ar;
  rw[ENERGY] = en;
  rw[DOF_LIM] = lim;
  real slip = 0, clear = 0, phase_err = 0, swing = 0, air = 0, con = 0, center = 0, fh = 0;
  real minfoot = INFINITY;
  static const int foot_of_sensor[4] = {1, 0, 3, 2};
  for (int i = 0; i < 4; i++) {
    const real* v = s + 37 + 3 * i; const real* pf = s + 25 + 3 * i;
    real vxy2 = v[0] * v[0] + v[1] * v[1];
    slip += vxy2 * contact[i];
    if (t->variant) clear += FABS(d->foot_xpos[foot_of_sensor[i]][2] - (in->H_max[i] - t->base_feet_distance + t->swing_height)) * SQRT(SQRT(vxy2)); /* joystick.py:569-572 */
    else clear += FABS(pf[2] - (in->H_max[i] + t->swing_height)) * SQRT(SQRT(vxy2));
    real rz = gait_get_z(in->phase[i], in->H_max[i] + t->swing_height, t->base_feet_distance);
    phase_err += (pf[2] - rz) * (pf[2] - rz);
    int swing_mask = (in->phase[i] / (2 * PI_R)) >= (real)0.5;
    swing += (pf[2] - t->swing_height) * (pf[2] - t->swing_height) * swing_mask;
    air += (in->feet_air_time[i] - (t->variant ? (real)0.5 : (real)0.1)) * first_contact[i]; /* joystick.py:591 vs joystick_pgtt.py:597 */
    con += (real)(swing_mask && contact[i]);
    center += pf[0] * pf[0] + pf[1] * pf[1]; /* init_feet_pos stays zero (joystick_pgtt.py:130 is a no-op) */
    real er = in->swing_peak[i] / t->swing_height - 1;
    fh += er * er * first_contact[i];
    real fz = d->foot_xpos[foot_of_sensor[i]][2];
    if (fz < minfoot) minfoot = fz;
  }
  rw[FEET_SLIP] = slip * (cmd_norm > (real)0.01);
  rw[FEET_CLEAR] = clear;
  rw[FEET_PHASE] = EXP(-phase_err / t->phase_sigma);
  real bh = d->qpos[2] - minfoot - (real)0.27;
  rw[BODY_HEIGHT] = bh * bh;
  rw[FEET_SWING] = swing;
  rw[FEET_AIR] = air * (cmd_norm > (real)0.01);
  rw[CONTACT] = -con;
  rw[CENTER] = center;
  rw[FEET_HEIGHT] = fh * (cmd_norm > (real)0.01);
  static const int sum_order[NREW] = {TRACK_LIN, TRACK_ANG, LIN_VEL_Z, ANG_VEL_XY, ORIENT, STAND_STILL, TERMINATION, POSE, TORQUES,
                                      ACTION_RATE, ENERGY, FEET_SLIP, FEET_CLEAR, FEET_PHASE, BODY_HEIGHT, FEET_SWING, FEET_AIR,
                                      DOF_LIM, CONTACT, CENTER, FEET_HEIGHT};
  real total = 0;
  for (int k = 0; k < NREW; k++) { rw[k] *= t->reward_scale[k]; }
  for (int k = 0; k < NREW; k++) total += rw[sum_order[k]];
  real reward = total * dt;
  reward = reward < 0 ? 0 : (reward > 10000 ? 10000 : reward);
  /* info bookkeeping, joystick_pgtt.py:205-224 */
  memcpy(in->last_last_act, in->last_act, sizeof(in->last_act));
  memcpy(in->last_act, action, sizeof(real) * NU);
  in->step += 1;
  for (int i = 0; i < 4; i++) in->phase[i] = FMOD(in->phase[i] + in->phase_dt, 2 * PI_R);
  in->steps_until_next_cmd -= 1;
  uint32_t a[2], key1[2], key2[2];
  rng_split(part, in->rng, 3, 1, key1); rng_split(part, in->rng, 3, 2, key2); rng_split(part, in->rng, 3, 0, a);
  in->rng[0] = a[0]; in->rng[1] = a[1];
  if (in->steps_until_next_cmd <= 0) { /* sample_command, joystick_pgtt.py:603-611 */
    uint32_t y_rng[2], w_rng[2], z_rng[2];
    rng_split(part, key1, 4, 1, y_rng); rng_split(part, key1, 4, 2, w_rng); rng_split(part, key1, 4, 3, z_rng);
    for (int i = 0; i < 3; i++) {
      real y = rng_uniform(part, y_rng, 3, i, t->cmd_u_min[i], t->cmd_u_max[i]);
      real z = (real)(rng_uniform(part, z_rng, 3, i, 0, 1) < t->cmd_b[i]);
      real w = (real)(rng_uniform(part, w_rng, 3, i, 0, 1) < (real)0.5);
      in->command[i] = in->command[i] - w * (in->command[i] - y * z);
    }
  }
  if (done || in->steps_until_next_cmd <= 0)
    in->steps_until_next_cmd = (int)RINT(-LOG1P(-rng_uniform(part, key2, 1, 0, 0, 1)) * (real)5.0 / dt);
  real sp = 0;
  for (int i = 0; i < 4; i++) {
    in->feet_air_time[i] *= (real)(!contact[i]);
    in->last_contact[i] = contact[i];
    in->swing_peak[i] *= (real)(!contact[i]);
    sp += in->swing_peak[i];
    e->contact_flags[i] = contact[i]; e->first_contact[i] = first_contact[i];
  }
  for (int k = 0; k < NREW; k++) e->metrics[k] = rw[k];
  e->metrics[NREW] = sp / 4;
  e->reward = reward; e->done = (real)done;
}

/* wrap_for_brax_training(env, episode_length, action_repeat=1): auto-reset + episode wrapper */
void orc_env_step(const TaskCfg* t, Env* e, const real* action) {
  Info* in = &e->info;
  if (e->done != 0) in->steps = 0;          /* BraxAutoResetWrapper.step: zero steps where done, clear done */
  e->done = 0;
  orc_task_step(t, e, action);
  /* EpisodeWrapper.step */
  real steps = in->steps + 1;
  real done_inner = e->done;
  real done = steps >= (real)t->episode_length ? 1 : done_inner;
  in->truncation = steps >= (real)t->episode_length ? 1 - done_inner : 0;
  in->steps = steps;
  real prev_done = in->episode_done;
  in->episode_metrics[0] = (in->episode_metrics[0] + e->reward) * (1 - prev_done);
  in->episode_metrics[1] = (in->episode_metrics[1] + 1) * (1 - prev_done);
  for (int k = 0; k < NMETRIC; k++) in->episode_metrics[2 + k] = (in->episode_metrics[2 + k] + e->metrics[k]) * (1 - prev_done);
  in->episode_done = done;
  e->done = done;
  if (done != 0) {                          /* restore cached first data / obs only */
    memcpy(&e->d, &e->first_data, sizeof(Data));
    memcpy(e->obs_state, e->first_obs_state, sizeof(e->obs_state));
    memcpy(e->obs_priv, e->first_obs_priv, sizeof(e->obs_priv));
  }
}

/* ------------------------------------------------------------------------------------------
 * domain_randomize (go2/randomize.py:23-171 with terrain; randomize_simple.py:24-138 without)
 * `nominal` holds the un-randomised model; dyn = 0 keeps every dynamics multiplier at 1 and only
 * assigns the terrain ("no DR" of BASELINE config 2).
 * ---------------------------------------------------------------------------------------- */
void orc_domain_randomize(const Model* nominal, Model* out, const uint32_t rng_in[2], const float* terrain, int n_terrains,
                          int n_model_bodies, int part, int dyn, int* terrain_index) {
  uint32_t rng[2] = {rng_in[0], rng_in[1]}, key[2];
  *out = *nominal;
  int stairs = terrain != NULL;
  split2(part, rng, key); /* floor friction: the stairs variant discards it */
  if (!stairs && dyn) out->floor_friction[0] = rng_uniform(part, key, 1, 0, (real)0.4, (real)1.0);
  if (stairs) {
    split2(part, rng, key);
    if (dyn) for (int k = 0; k < NBOX; k++) out->box_friction[k][0] = rng_uniform(part, key, NBOX, k, (real)0.4, (real)1.0);
  }
  split2(part, rng, key); /* frictionloss * U(.9,1.1): frictionloss is 0 in the model */
  split2(part, rng, key);
  if (dyn) for (int i = 0; i < 12; i++) out->dof_armature[6 + i] = nominal->dof_armature[6 + i] * rng_uniform(part, key, 12, i, (real)1.0, (real)1.05);
  split2(part, rng, key);
  if (dyn) for (int i = 0; i < 3; i++) out->body_ipos[1][i] = nominal->body_ipos[1][i] + rng_uniform(part, key, 3, i, (real)-0.05, (real)0.05);
  split2(part, rng, key);
  if (dyn) for (int b = 0; b < NBODY; b++) out->body_mass[b] = nominal->body_mass[b] * rng_uniform(part, key, n_model_bodies, b, (real)0.9, (real)1.1);
  split2(part, rng, key);
  if (dyn) out->body_mass[1] += rng_uniform(part, key, 1, 0, (real)-1.0, (real)1.0);
  split2(part, rng, key);
  if (dyn) for (int i = 0; i < 12; i++) out->qpos0[7 + i] = nominal->qpos0[7 + i] + rng_uniform(part, key, 12, i, (real)-0.05, (real)0.05);
  split2(part, rng, key);
  if (dyn) for (int i = 0; i < 12; i++) out->dof_damping[6 + i] = nominal->dof_damping[6 + i] * rng_uniform(part, key, 12, i, (real)0.9, (real)1.1);
  split2(part, rng, key);
  if (dyn) for (int a = 0; a < 12; a++) {
    real g = rng_uniform(part, key, 12, a, (real)0.9, (real)1.1);
    out->act_gain[a] = nominal->act_gain[a] * g;
    out->act_bias[a][1] = nominal->act_bias[a][1] * g;
  }
  *terrain_index = -1;
  if (stairs) {
    split2(part, rng, key);
    int idx = rng_randint(part, key, 0, n_terrains);
    *terrain_index = idx;
    const float* T = terrain + (size_t)idx * NBOX * 10;
    for (int k = 0; k < NBOX; k++) {
      for (int i = 0; i < 3; i++) out->box_pos[k][i] = T[k * 10 + i];
      for (int i = 0; i < 4; i++) out->box_quat[k][i] = T[k * 10 + 3 + i];
      for (int i = 0; i < 3; i++) out->box_size[k][i] = T[k * 10 + 7 + i];
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * flat, ctypes-friendly API (all numeric I/O as double / int32 arrays)
 * ---------------------------------------------------------------------------------------- */
typedef struct { int n; Env* env; Model nominal; TaskCfg cfg; int n_model_bodies; } Handle;

size_t orc_sizeof_real(void) { return sizeof(real); }

/* model constants arrive as one flat double vector in the order of oracle/oracle.py:pack_model */
#define TAKE(dst, cnt) do { for (int _i = 0; _i < (cnt); _i++) ((real*)(dst))[_i] = (real)*p++; } while (0)
#define TAKEI(dst, cnt) do { for (int _i = 0; _i < (cnt); _i++) ((int*)(dst))[_i] = (int)*p++; } while (0)

void* orc_create(int n_envs, const double* mc, int n_mc, const double* tc, int n_tc) {
  Handle* h = (Handle*)calloc(1, sizeof(Handle));
  h->n = n_envs;
  h->env = (Env*)calloc((size_t)n_envs, sizeof(Env));
  Model* m = &h->nominal;
  const double* p = mc;
  TAKE(&m->timestep, 1); TAKE(m->gravity, 3); TAKE(&m->impratio, 1); TAKE(&m->tolerance, 1); TAKE(&m->ls_tolerance, 1); TAKE(&m->meaninertia, 1);
  TAKEI(&m->iterations, 1); TAKEI(&m->ls_iterations, 1); TAKEI(&m->max_geom_pairs, 1); TAKEI(&m->max_contact_points, 1); TAKEI(&m->n_boxes, 1);
  TAKEI(m->body_parent, NBODY); TAKE(m->body_pos, NBODY * 3); TAKE(m->body_quat, NBODY * 4); TAKE(m->body_ipos, NBODY * 3); TAKE(m->body_iquat, NBODY * 4);
  TAKE(m->body_mass, NBODY); TAKE(m->body_inertia, NBODY * 3); TAKE(m->body_invweight0, NBODY * 2);
  TAKEI(m->jnt_body, NHINGE); TAKE(m->jnt_axis, NHINGE * 3); TAKE(m->jnt_range, NHINGE * 2); TAKE(m->jnt_solref, 2); TAKE(m->jnt_solimp, 5);
  TAKE(m->qpos0, NQ); TAKE(m->dof_armature, NV); TAKE(m->dof_damping, NV); TAKE(m->dof_invweight0, NV);
  TAKEI(m->act_dof, NU); TAKE(m->act_gain, NU); TAKE(m->act_bias, NU * 3); TAKE(m->act_ctrlrange, NU * 2); TAKE(m->act_forcerange, NU * 2);
  TAKEI(m->foot_body, NFOOT); TAKEI(m->foot_geom_id, NFOOT); TAKE(m->foot_pos, 3); TAKE(&m->foot_radius, 1); TAKE(m->foot_friction, 3);
  TAKE(m->foot_solref, 2); TAKE(m->foot_solimp, 5); TAKE(&m->foot_margin, 1);
  TAKEI(&m->floor_geom_id, 1); TAKEI(&m->box_geom_id0, 1);
  TAKE(m->floor_friction, 3); TAKE(m->floor_solref, 2); TAKE(m->floor_solimp, 5);
  TAKE(&m->box_rbound, 1); TAKE(m->box_solref, 2); TAKE(m->box_solimp, 5);
  real bf[3]; TAKE(bf, 3);
  for (int k = 0; k < NBOX; k++) for (int i = 0; i < 3; i++) m->box_friction[k][i] = bf[i];
  TAKE(m->box_pos, NBOX * 3); TAKE(m->box_quat, NBOX * 4); TAKE(m->box_size, NBOX * 3);
  TAKE(m->imu_pos, 3);
  TAKEI(&h->n_model_bodies, 1);
  if ((int)(p - mc) != n_mc) { fprintf(stderr, "orc_create: model constant count mismatch %d vs %d\n", (int)(p - mc), n_mc); free(h->env); free(h); return NULL; }
  TaskCfg* t = &h->cfg;
  p = tc;
  TAKE(&t->ctrl_dt, 1); TAKE(&t->action_scale, 1); TAKE(&t->noise_level, 1);
  TAKE(&t->noise_joint_pos, 1); TAKE(&t->noise_joint_vel, 1); TAKE(&t->noise_gyro, 1); TAKE(&t->noise_gravity, 1); TAKE(&t->noise_linvel, 1); TAKE(&t->noise_heightscan, 1);
  TAKE(t->reward_scale, NREW); TAKE(&t->tracking_sigma, 1); TAKE(&t->swing_height, 1); TAKE(&t->base_feet_distance, 1); TAKE(&t->phase_sigma, 1);
  TAKE(t->cmd_u_max, 3); TAKE(t->cmd_u_min, 3); TAKE(t->cmd_b, 3); TAKE(t->gait_freq, 2); TAKE(&t->soft_limit_factor, 1);
  TAKE(t->default_pose, NHINGE); TAKE(t->home_qpos, NQ);
  TAKEI(&t->history_update_steps, 1); TAKEI(&t->episode_length, 1); TAKEI(&t->n_substeps, 1); TAKEI(&t->rng_partitionable, 1); TAKEI(&t->variant, 1);
  if ((int)(p - tc) != n_tc) { fprintf(stderr, "orc_create: task constant count mismatch %d vs %d\n", (int)(p - tc), n_tc); free(h->env); free(h); return NULL; }
  for (int i = 0; i < n_envs; i++) { h->env[i].m = *m; h->env[i].terrain_index = -1; }
  return h;
}

void orc_destroy(void* hv) { Handle* h = (Handle*)hv; if (h) { free(h->env); free(h); } }

void orc_randomize(void* hv, const uint32_t* keys, const float* terrain, int n_terrains, int dyn) {
  Handle* h = (Handle*)hv;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < h->n; i++)
    orc_domain_randomize(&h->nominal, &h->env[i].m, keys + 2 * i, terrain, n_terrains, h->n_model_bodies, h->cfg.rng_partitionable, dyn,
                         &h->env[i].terrain_index);
}

void orc_reset(void* hv, const uint32_t* keys) {
  Handle* h = (Handle*)hv;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < h->n; i++) orc_env_reset(&h->cfg, &h->env[i], keys + 2 * i);
}

/* action [n,12] double; wrapped = 1 applies the training wrappers, 0 = bare Joystick.step */
void orc_step_envs(void* hv, const double* action, int wrapped) {
  Handle* h = (Handle*)hv;
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < h->n; i++) {
    real a[NU];
    for (int k = 0; k < NU; k++) a[k] = (real)action[i * NU + k];
    if (wrapped) orc_env_step(&h->cfg, &h->env[i], a); else orc_task_step(&h->cfg, &h->env[i], a);
  }
}

/* bare physics: forward (do_step = 0) or mjx.step (do_step = 1) on every env with its current qpos/qvel/ctrl */
void orc_physics(void* hv, int do_step) {
  Handle* h = (Handle*)hv;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < h->n; i++) { if (do_step) orc_step(&h->env[i].m, &h->env[i].d); else orc_forward(&h->env[i].m, &h->env[i].d); }
}

void orc_scan(void* hv, const double* center, const double* yaw, double* out) { /* [n,3],[n] -> [n,117,3] */
  Handle* h = (Handle*)hv;
  for (int i = 0; i < h->n; i++) {
    real c[3] = {(real)center[3 * i], (real)center[3 * i + 1], (real)center[3 * i + 2]}, o[NRAY][3];
    orc_heightscan(&h->env[i].m, c, (real)yaw[i], o);
    for (int k = 0; k < NRAY * 3; k++) out[(size_t)i * NRAY * 3 + k] = (double)(&o[0][0])[k];
  }
}

/* mjx.ray(m, d, pnt, vec, geomgroup=(1,0,0,0,1,1)) for env i: distance to the nearest floor / box hit, -1 if none */
double orc_ray(void* hv, int i, const double* pnt, const double* vec) {
  Handle* h = (Handle*)hv;
  real p[3] = {(real)pnt[0], (real)pnt[1], (real)pnt[2]}, v[3] = {(real)vec[0], (real)vec[1], (real)vec[2]};
  real dist = ray_scene(&h->env[i].m, p, v);
  return isinf((double)dist) ? -1.0 : (double)dist;
}

/* ---- field access by name: kind 0 = real, 1 = int32, 2 = uint32 ----------------------------- */
typedef struct { const char* name; size_t off; int count; int kind; } Field;
#define F_(name, member, cnt, kind) {name, offsetof(Env, member), cnt, kind}
#include <stddef.h>
static const Field FIELDS[] = {
    F_("qpos", d.qpos, NQ, 0), F_("qvel", d.qvel, NV, 0), F_("ctrl", d.ctrl, NU, 0), F_("qacc", d.qacc, NV, 0),
    F_("qacc_warmstart", d.qacc_warmstart, NV, 0), F_("time", d.time, 1, 0),
    F_("xpos", d.xpos, NBODY * 3, 0), F_("xquat", d.xquat, NBODY * 4, 0), F_("xmat", d.xmat, NBODY * 9, 0), F_("xipos", d.xipos, NBODY * 3, 0),
    F_("ximat", d.ximat, NBODY * 9, 0), F_("xanchor", d.xanchor, NHINGE * 3, 0), F_("xaxis", d.xaxis, NHINGE * 3, 0),
    F_("subtree_com", d.subtree_com, 3, 0), F_("cinert", d.cinert, NBODY * 10, 0), F_("cdof", d.cdof, NV * 6, 0),
    F_("qM", d.qM, NV * NV, 0), F_("foot_xpos", d.foot_xpos, NFOOT * 3, 0), F_("site_xpos", d.site_xpos, 15, 0), F_("site_xmat", d.site_xmat, 9, 0),
    F_("ncon", d.ncon, 1, 1), F_("efc_J", d.efc_J, NEFC * NV, 0), F_("efc_D", d.efc_D, NEFC, 0), F_("efc_aref", d.efc_aref, NEFC, 0),
    F_("efc_pos", d.efc_pos, NEFC, 0), F_("efc_force", d.efc_force, NEFC, 0), F_("cvel", d.cvel, NBODY * 6, 0), F_("cdof_dot", d.cdof_dot, NV * 6, 0),
    F_("qfrc_bias", d.qfrc_bias, NV, 0), F_("qfrc_passive", d.qfrc_passive, NV, 0), F_("qfrc_actuator", d.qfrc_actuator, NV, 0),
    F_("qfrc_smooth", d.qfrc_smooth, NV, 0), F_("qacc_smooth", d.qacc_smooth, NV, 0), F_("qfrc_constraint", d.qfrc_constraint, NV, 0),
    F_("actuator_force", d.actuator_force, NU, 0), F_("sensordata", d.sensordata, NSENSOR, 0), F_("solver_niter", d.solver_niter, 1, 1),
    F_("rng", info.rng, 2, 2), F_("command", info.command, 3, 0), F_("step", info.step, 1, 1), F_("steps_until_next_cmd", info.steps_until_next_cmd, 1, 1),
    F_("phase", info.phase, 4, 0), F_("phase_dt", info.phase_dt, 1, 0), F_("gait_freq", info.gait_freq, 1, 0), F_("last_act", info.last_act, NU, 0),
    F_("last_last_act", info.last_last_act, NU, 0), F_("feet_air_time", info.feet_air_time, 4, 0), F_("last_contact", info.last_contact, 4, 1),
    F_("swing_peak", info.swing_peak, 4, 0), F_("H_max", info.H_max, 4, 0), F_("H_min", info.H_min, 4, 0), F_("heightscan", info.heightscan, NRAY * 3, 0),
    F_("motor_targets", info.motor_targets, NU, 0), F_("qpos_error_history", info.qpos_error_history, NHIST, 0), F_("qvel_history", info.qvel_history, NHIST, 0),
    F_("steps", info.steps, 1, 0), F_("truncation", info.truncation, 1, 0), F_("episode_done", info.episode_done, 1, 0),
    F_("episode_metrics", info.episode_metrics, 2 + NMETRIC, 0),
    F_("obs_state", obs_state, NOBS, 0), F_("obs_priv", obs_priv, NPRIV, 0), F_("reward", reward, 1, 0), F_("done", done, 1, 0),
    F_("metrics", metrics, NMETRIC, 0), F_("contact_flags", contact_flags, 4, 1), F_("first_contact", first_contact, 4, 1),
    F_("terrain_index", terrain_index, 1, 1),
    F_("first_qpos", first_data.qpos, NQ, 0), F_("first_qvel", first_data.qvel, NV, 0), F_("first_obs_state", first_obs_state, NOBS, 0),
    /* per-env model */
    F_("m_body_mass", m.body_mass, NBODY, 0), F_("m_body_ipos", m.body_ipos, NBODY * 3, 0), F_("m_qpos0", m.qpos0, NQ, 0),
    F_("m_dof_armature", m.dof_armature, NV, 0), F_("m_dof_damping", m.dof_damping, NV, 0), F_("m_act_gain", m.act_gain, NU, 0),
    F_("m_act_bias", m.act_bias, NU * 3, 0), F_("m_box_friction", m.box_friction, NBOX * 3, 0), F_("m_box_pos", m.box_pos, NBOX * 3, 0),
    F_("m_box_quat", m.box_quat, NBOX * 4, 0), F_("m_box_size", m.box_size, NBOX * 3, 0), F_("m_floor_friction", m.floor_friction, 3, 0),
    F_("m_n_boxes", m.n_boxes, 1, 1),
};
#define NFIELDS ((int)(sizeof(FIELDS) / sizeof(FIELDS[0])))

static const Field* find_field(const char* name) {
  for (int i = 0; i < NFIELDS; i++) if (strcmp(FIELDS[i].name, name) == 0) return &FIELDS[i];
  return NULL;
}
int orc_field_count(const char* name) { const Field* f = find_field(name); return f ? f->count : -1; }
int orc_get(void* hv, const char* name, double* out) {
  Handle* h = (Handle*)hv; const Field* f = find_field(name);
  if (!f) return -1;
  for (int i = 0; i < h->n; i++) {
    const char* base = (const char*)&h->env[i] + f->off;
    for (int k = 0; k < f->count; k++)
      out[(size_t)i * f->count + k] = f->kind == 0 ? (double)((const real*)base)[k] : f->kind == 1 ? (double)((const int*)base)[k] : (double)((const uint32_t*)base)[k];
  }
  return 0;
}
int orc_set(void* hv, const char* name, const double* in) {
  Handle* h = (Handle*)hv; const Field* f = find_field(name);
  if (!f) return -1;
  for (int i = 0; i < h->n; i++) {
    char* base = (char*)&h->env[i] + f->off;
    for (int k = 0; k < f->count; k++) {
      double v = in[(size_t)i * f->count + k];
      if (f->kind == 0) ((real*)base)[k] = (real)v; else if (f->kind == 1) ((int*)base)[k] = (int)v; else ((uint32_t*)base)[k] = (uint32_t)v;
    }
  }
  return 0;
}

/* contact list of env i: out_f [8][17] = dist, pos3, frame9, mu, includemargin, solimp0, solref0 ; out_i [8][5] = geom1, geom2, body, foot, box */
void orc_get_contacts(void* hv, int i, double* out_f, int* out_i) {
  Handle* h = (Handle*)hv; const Data* d = &h->env[i].d;
  for (int c = 0; c < NCON; c++) {
    const Contact* k = &d->contact[c];
    double* o = out_f + c * 17;
    if (c >= d->ncon) { for (int a = 0; a < 17; a++) o[a] = 0; for (int a = 0; a < 5; a++) out_i[c * 5 + a] = -1; continue; }
    o[0] = k->dist; for (int a = 0; a < 3; a++) o[1 + a] = k->pos[a]; for (int a = 0; a < 9; a++) o[4 + a] = k->frame[a];
    o[13] = k->mu; o[14] = k->includemargin; o[15] = k->solimp[0]; o[16] = k->solref[0];
    out_i[c * 5] = k->geom1; out_i[c * 5 + 1] = k->geom2; out_i[c * 5 + 2] = k->body; out_i[c * 5 + 3] = k->foot; out_i[c * 5 + 4] = k->box;
  }
}

/* jax.random probes for tests */
void orc_rng_split(int part, const uint32_t* key, int num, uint32_t* out) { for (int i = 0; i < num; i++) rng_split(part, key, num, i, out + 2 * i); }
void orc_rng_uniform(int part, const uint32_t* key, int n, double lo, double hi, double* out) { for (int i = 0; i < n; i++) out[i] = rng_uniform(part, key, n, i, (real)lo, (real)hi); }
void orc_rng_bits(int part, const uint32_t* key, int n, uint32_t* out) { for (int i = 0; i < n; i++) out[i] = rng_bits(part, key, n, i); }
int orc_rng_randint(int part, const uint32_t* key, int lo, int hi) { return rng_randint(part, key, lo, hi); }
double orc_gait_get_z(double phi, double h_max, double stance) { return (double)gait_get_z((real)phi, (real)h_max, (real)stance); }
double orc_quat_to_yaw(const double* q) { real qq[4] = {(real)q[0], (real)q[1], (real)q[2], (real)q[3]}; return (double)quat_to_yaw(qq); }
