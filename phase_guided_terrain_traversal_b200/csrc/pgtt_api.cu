// pgtt_api.cu - C ABI (include/pgtt_b200.h) over the env kernels.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a -> libpgtt_b200.so (CUDA only; there
// is no CPU path behind these entry points). With -DPGTT_HOST_EMU the same file is compiled by g++
// into tests/simt_emu/libpgtt_emu.so, where "device" memory is host memory and a launch runs every
// warp through the fiber emulator - test infrastructure for GPU-less CI, never shipped or loaded by
// the package.
#include "../../include/pgtt_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "pgtt_debug.h"
#include "pgtt_env.cuh"
#include "pgtt_quad.cuh"

#define MAX_WARPS_PER_BLOCK 28   // x 72 registers x 7.3 KB workspace = one SM: 28 resident envs per SM
#define ENV_CTAS_PER_SM 1
#define WS_BYTES ((sizeof(WS) + 15) / 16 * 16)

static thread_local std::string g_err;

// Warps (= envs) per CTA of the warp-per-env kernels. An SM holds 28 such warps (registers: 28 x 32 x 72; shared memory:
// 28 x 7.3 KB), i.e. 4096 envs are ONE resident wave on 148 SMs; the warps of a CTA run the stages in lockstep so the
// instruction stream is fetched once per CTA. Picks the count in [7, 14] that needs the fewest waves, then the fewest
// idle warp slots. PGTT_WARPS_PER_BLOCK overrides (tuning / tests).
static int pick_warps_per_block(int n_envs) {
  if (const char* s = getenv("PGTT_WARPS_PER_BLOCK")) {
    const int v = atoi(s);
    if (v >= 1 && v <= MAX_WARPS_PER_BLOCK) return v;
  }
  int n_sm = 148;
#ifndef PGTT_HOST_EMU
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
#endif
  int best = MAX_WARPS_PER_BLOCK;
  double best_cost = 1e30;
  for (int w = MAX_WARPS_PER_BLOCK; w >= 7; w--) {
    if ((MAX_WARPS_PER_BLOCK * ENV_CTAS_PER_SM) % w) continue;   // CTAs must tile the 28 warp slots of an SM
    const long ctas = (n_envs + w - 1) / w, per_sm = (MAX_WARPS_PER_BLOCK * ENV_CTAS_PER_SM) / w, waves = (ctas + n_sm * per_sm - 1) / (n_sm * per_sm);
    // fewest waves; then every SM busy (small batches want small CTAs); then the larger lockstep CTA
    const double cost = (double)waves + (ctas * 10 < n_sm * 9 ? 0.5 : 0.0) + 0.001 * (MAX_WARPS_PER_BLOCK - w);
    if (cost < best_cost - 1e-12) { best_cost = cost; best = w; }
  }
  return best;
}
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

// ----------------------------------------------------------------------------------------------
// debug forward (device): mjx.forward with every intermediate dumped
// ----------------------------------------------------------------------------------------------
DEV void env_debug_forward(WS& w, const EnvBuffers& B, float* out_all, int env, int lane) {
  float* o = out_all + (size_t)env * PGTT_DEBUG_FLOATS;
  load_model(w, B, env, lane);
  load_state(w, B, env, lane);
  if (lane < NU) w.ctrl[lane] = B.ctrl[env * NU + lane];
  syncwarp();
  kinematics(w, lane);
  com_inertia_cdof(w, lane);
  crb_and_inertia(w, lane);
  velocity_rne(w, lane);
  smooth_forces(w, lane);
  collision(w, B, env, lane, true);
  Rows R;
  make_rows(w, R, lane);
  // the position-stage arrays share storage with the factorisation scratch: dump them before the factor is formed
  for (int i = lane; i < 39; i += 32) { o[DBG_XPOS + i] = (&w.xpos[0][0])[i]; o[DBG_XIPOS + i] = (&w.xipos[0][0])[i]; }
  for (int i = lane; i < 117; i += 32) o[DBG_XMAT + i] = (&w.xmat[0][0])[i];
  if (lane < 3) o[DBG_COM + lane] = w.com[lane];
  for (int i = lane; i < 130; i += 32) o[DBG_CINERT + i] = (&w.cinert[0][0])[i];
  for (int i = lane; i < 108; i += 32) o[DBG_CDOF + i] = (&w.cdof[0][0])[i];
  if (lane < NV) o[DBG_BIAS + lane] = w.bias[lane];
  syncwarp();
  arrow_factor(w, w.MB, w.MC, w.MA, lane);
  arrow_solve(w, w.qs, w.qas, lane);
  for (int e = lane; e < 324; e += 32) {
    const int i = e / 18, j = e % 18;
    float v = 0.f;
    if (i < 6 && j < 6) v = w.MB[i * 6 + j];
    else if (i < 6) v = w.MC[((j - 6) / 3) * 18 + i * 3 + (j - 6) % 3];
    else if (j < 6) v = w.MC[((i - 6) / 3) * 18 + j * 3 + (i - 6) % 3];
    else if ((i - 6) / 3 == (j - 6) / 3) v = w.MA[((i - 6) / 3) * 9 + ((i - 6) % 3) * 3 + (j - 6) % 3];
    o[DBG_QM + e] = v;
  }
  if (lane < NV) { o[DBG_QS + lane] = w.qs[lane]; o[DBG_QAS + lane] = w.qas[lane]; }
  if (lane < NCON) {
    float* c = o + DBG_CONTACT + 16 * lane;
    c[0] = w.c_dist[lane];
    for (int i = 0; i < 3; i++) c[1 + i] = w.c_pos[lane][i];
    for (int i = 0; i < 9; i++) c[4 + i] = w.c_frame[lane][i];
    c[13] = w.c_mu[lane]; c[14] = (float)w.c_leg[lane]; c[15] = (float)w.c_box[lane];
  }
  for (int i = lane; i < 44 * 18; i += 32) o[DBG_EFC_J + i] = 0.f;
  syncwarp();
  o[DBG_EFC_D + 12 + lane] = R.D; o[DBG_EFC_AREF + 12 + lane] = R.aref;
  for (int i = 0; i < 9; i++) o[DBG_EFC_J + (12 + lane) * 18 + col_dof(i, R.leg)] = R.jr[i];
  if (lane < 12) { o[DBG_EFC_D + lane] = R.lD; o[DBG_EFC_AREF + lane] = R.laref; o[DBG_EFC_J + lane * 18 + 6 + lane] = R.lsign; }
  const int niter = solve(w, R, lane);
  sensors(w, lane);
  if (lane < NV) o[DBG_QACC + lane] = w.qacc[lane];
  for (int i = lane; i < NSENSOR; i += 32) o[DBG_SENS + i] = w.sens[i];
  if (lane == 0) o[DBG_NITER] = (float)niter;
  if (lane < NU) { o[DBG_ACTF + lane] = w.actf[lane]; o[DBG_FOOT + lane] = (&w.foot[0][0])[lane]; }
}

// ----------------------------------------------------------------------------------------------
// kernels / launch shims
// ----------------------------------------------------------------------------------------------
enum { OP_STEP = 0, OP_RESET, OP_FORWARD, OP_DEBUG, OP_TASK, OP_SCAN, OP_STEP_TASK };   // OP_TASK / OP_SCAN run in the task kernel; OP_STEP_TASK = physics + task in one launch (generation 1)
struct LaunchArgs {
  EnvBuffers B;
  int op, wrapped;
  const float* action;
  const uint32_t* keys;
  const float* center;
  const float* yaw;
  float* out;
  RecordSlot rec;
};
#define TASK_BYTES ((sizeof(TaskWS) + 15) / 16 * 16)
#define TASK_WARPS 7   // warps (= envs) per CTA of the task kernel; 4 CTAs x 7 warps x 72 registers per SM: 4096 envs are one wave of 586 CTAs
#define SMEM_MAX (227 * 1024)
// warps per CTA of the reset kernel: each carries a physics workspace and a task workspace
#define RESET_WARPS ((int)(SMEM_MAX / (WS_BYTES + TASK_BYTES)) < MAX_WARPS_PER_BLOCK ? (int)(SMEM_MAX / (WS_BYTES + TASK_BYTES)) : MAX_WARPS_PER_BLOCK)

DEV void dispatch(WS& w, TaskWS& t, const LaunchArgs& a, int env, int lane) {
  switch (a.op) {
    case OP_STEP: env_physics(w, a.B, a.action, env, lane); break;
    case OP_RESET: env_reset(w, t, a.B, a.keys, env, lane); break;
    case OP_FORWARD: env_forward(w, a.B, env, lane); break;
    case OP_DEBUG: env_debug_forward(w, a.B, a.out, env, lane); break;
    case OP_TASK: task_step(t, a.B, a.action, env, lane, a.wrapped, a.rec); break;
    case OP_SCAN: task_scan(t, a.B, a.center, a.yaw, a.out, env, lane); break;
    case OP_STEP_TASK: env_physics(w, a.B, a.action, env, lane); task_step(t, a.B, a.action, env, lane, a.wrapped, a.rec); break;
  }
}

#ifndef PGTT_HOST_EMU
#define CUDA_OK(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return fail(PGTT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// generation-1 physics / reset kernels: one warp per env, workspace in shared memory (reset also carries a TaskWS per warp)
template <int OP>
__global__ void __launch_bounds__(MAX_WARPS_PER_BLOCK * 32, ENV_CTAS_PER_SM) pgtt_env_kernel(LaunchArgs a) {
  extern __shared__ float4 smem4[];
  const int wpb = blockDim.x >> 5;
  const int warp = warp_index();
  int lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));   // volatile: one register for the whole kernel instead of S2R + LOP3 re-derived at every use
  const int env = blockIdx.x * wpb + warp;
  if (env >= a.B.N) return;   // warp-uniform (warp_index is a broadcast): no collective sees a partial warp
  WS& w = *reinterpret_cast<WS*>(reinterpret_cast<char*>(smem4) + (size_t)warp * WS_BYTES);
  {
    const int live = a.B.N - blockIdx.x * wpb;
    if (lane == 0) { w.bar_threads = 32 * (live < wpb ? live : wpb); w.trace = nullptr; w.tix = 0; }
    syncwarp();
  }
  if (OP == OP_STEP) env_physics(w, a.B, a.action, env, lane, reinterpret_cast<long long*>(a.out));
  else if (OP == OP_STEP_TASK) {
    // the whole control step in one launch: the task layer of THIS env right behind its physics, in the same warp. The physics
    // workspace is dead by then, so the task workspace aliases it; mjx.Data travels through the handle's buffers exactly as between the
    // two kernels (same warp: __syncwarp orders its global writes before its reads). A CTA barrier keeps the warps of the SM in step
    // on the task layer's instruction stream as well.
    static_assert(sizeof(TaskWS) <= sizeof(WS), "the task workspace must fit the physics workspace it aliases");
    const int bar_threads = w.bar_threads;
    env_physics(w, a.B, a.action, env, lane, nullptr);
    syncwarp();
    cta_bar(bar_threads);
    task_step(*reinterpret_cast<TaskWS*>(&w), a.B, a.action, env, lane, a.wrapped, a.rec);
  }
  else if (OP == OP_RESET) {
    TaskWS& t = *reinterpret_cast<TaskWS*>(reinterpret_cast<char*>(smem4) + (size_t)wpb * WS_BYTES + (size_t)warp * TASK_BYTES);
    env_reset(w, t, a.B, a.keys, env, lane);
  }
  else if (OP == OP_FORWARD) env_forward(w, a.B, env, lane);
  else env_debug_forward(w, a.B, a.out, env, lane);
}

// generation-2 physics kernels: eight envs per warp (pgtt_quad.cuh), grid = ceil(N / 8) warps in CTAs of qw warps
#define QWARPS_MAX 8
template <int OP>
__global__ void __launch_bounds__(32 * QWARPS_MAX) pgtt_quad_kernel(LaunchArgs a) {
  extern __shared__ float4 smem4[];
  const int lane = threadIdx.x & 31, warp = warp_index();
  QShared& sh = *reinterpret_cast<QShared*>(reinterpret_cast<char*>(smem4) + (size_t)warp * ((sizeof(QShared) + 15) / 16 * 16));
  q_stage_consts(sh, lane);
  const int env = (blockIdx.x * (blockDim.x >> 5) + warp) * QENV + (lane >> 2);
  if (OP == OP_STEP) q_env_physics(sh, a.B, a.action, env, lane);
  else q_env_debug_forward(sh, a.B, a.out, env, lane);
}

// task kernel (pgtt_task.cuh): one warp per env, TASK_WARPS envs per CTA
template <int OP>
__global__ void __launch_bounds__(32 * TASK_WARPS, 4) pgtt_task_kernel(LaunchArgs a) {
  extern __shared__ float4 smem4[];
  const int warp = warp_index(), lane = threadIdx.x & 31;
  const int env = blockIdx.x * TASK_WARPS + warp;
  if (env >= a.B.N) return;
  TaskWS& t = *reinterpret_cast<TaskWS*>(reinterpret_cast<char*>(smem4) + (size_t)warp * TASK_BYTES);
  if (OP == OP_TASK) task_step(t, a.B, a.action, env, lane, a.wrapped, a.rec);
  else task_scan(t, a.B, a.center, a.yaw, a.out, env, lane);
}

__global__ void pgtt_discount_kernel(const float* __restrict__ done, float* __restrict__ dc, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dc[i] = 1.0f - done[i];
}

__global__ void pgtt_randomize_kernel(EnvBuffers B, const uint32_t* keys, int dyn) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env < B.N) env_randomize(B, keys, env, dyn);
}
#else
ModelConst g_mc;
#define CUDA_OK(call) do { (void)0; } while (0)
struct WarpJob { const LaunchArgs* a; int env; WS* w; TaskWS* t; };
static void warp_entry(void* p, int lane) {
  WarpJob* j = (WarpJob*)p;
  dispatch(*j->w, *j->t, *j->a, j->env, lane);
}
struct QuadJob { const LaunchArgs* a; int env0; QShared* sh; };
static void quad_entry(void* p, int lane) {
  QuadJob* j = (QuadJob*)p;
  q_stage_consts(*j->sh, lane);
  const int env = j->env0 + (lane >> 2);
  if (j->a->op == OP_STEP) q_env_physics(*j->sh, j->a->B, j->a->action, env, lane);
  else q_env_debug_forward(*j->sh, j->a->B, j->a->out, env, lane);
}
#endif

struct pgtt_env {
  int device, N, wpb;
  int quad;   // 1: generation-2 quad-per-env physics kernel for step / debug-forward, 0: warp-per-env (PGTT_KERNEL=warp|quad overrides)
  int fuse_task;   // generation 1: the task layer runs behind the physics in the same launch (PGTT_FUSE_TASK=0|1 overrides)
  ModelConst mc;
  EnvBuffers B;
  std::vector<void*> allocs;
  float* terrain_dev;
  int n_terrains;
  int64_t launches;
  uint64_t serial;   // unique per pgtt_create (graph caches key on it: a freed handle's address can be re-used)
  bool randomized;
};

// The model / task constants live in __constant__ memory (one table per device). The handle whose constants are resident is
// tracked per device; when another handle launches, the table is replaced by a STREAM-ORDERED copy on the launching stream
// that first waits (device side, cudaStreamWaitEvent) for the last launch of the previous owner - so two handles on one
// device (a training and an evaluation env, training/train.py:242-263) alternate without any host synchronisation.
// Events are only recorded once a second handle exists on the device (its creation drains the device once).
#define PGTT_MAX_DEVICES 64
struct DeviceConsts {
  pgtt_env* owner = nullptr;
  int handles = 0;
#ifndef PGTT_HOST_EMU
  cudaEvent_t last = nullptr;   // after the most recent launch of `owner`
  bool have_last = false;
#endif
};
static DeviceConsts g_dev[PGTT_MAX_DEVICES];
static DeviceConsts& dev_consts(const pgtt_env* e) { return g_dev[e->device >= 0 && e->device < PGTT_MAX_DEVICES ? e->device : 0]; }

static void* dev_alloc(pgtt_env* e, size_t bytes) {
  void* p = nullptr;
#ifndef PGTT_HOST_EMU
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  cudaMemset(p, 0, bytes);
#else
  p = calloc(1, bytes);
#endif
  if (p) e->allocs.push_back(p);
  return p;
}

// make e's constant table the resident one for launches on `stream` (no host synchronisation)
static int make_resident(pgtt_env* e, void* stream) {
  DeviceConsts& d = dev_consts(e);
  if (d.owner == e) return 0;
#ifndef PGTT_HOST_EMU
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_OK(cudaSetDevice(e->device));
  if (d.have_last) CUDA_OK(cudaStreamWaitEvent(st, d.last, 0));
  CUDA_OK(cudaMemcpyToSymbolAsync(g_mc, &e->mc, sizeof(ModelConst), 0, cudaMemcpyHostToDevice, st));
#else
  (void)stream;
  g_mc = e->mc;
#endif
  d.owner = e;
  return 0;
}
// after the launches of one entry point: lets the next owner order itself behind them
static int mark_launched(pgtt_env* e, void* stream) {
#ifndef PGTT_HOST_EMU
  DeviceConsts& d = dev_consts(e);
  if (d.handles > 1) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs == cudaStreamCaptureStatusNone) {   // a capturing stream gets its record from pgtt_internal_mark_launched after the replay
      if (!d.last) CUDA_OK(cudaEventCreateWithFlags(&d.last, cudaEventDisableTiming));
      CUDA_OK(cudaEventRecord(d.last, st));
      d.have_last = true;
    }
  }
#else
  (void)e; (void)stream;
#endif
  return 0;
}

static int launch(pgtt_env* e, LaunchArgs& a, void* stream) {
  if (int rc = make_resident(e, stream)) return rc;
  a.B = e->B;
  const bool task = a.op == OP_TASK || a.op == OP_SCAN;
#ifndef PGTT_HOST_EMU
  cudaStream_t st = (cudaStream_t)stream;
  if (task) {
    const int blocks = (e->N + TASK_WARPS - 1) / TASK_WARPS;
    const size_t smem = TASK_WARPS * TASK_BYTES;
    if (a.op == OP_TASK) pgtt_task_kernel<OP_TASK><<<blocks, 32 * TASK_WARPS, smem, st>>>(a);
    else pgtt_task_kernel<OP_SCAN><<<blocks, 32 * TASK_WARPS, smem, st>>>(a);
  } else if (e->quad && (a.op == OP_STEP || a.op == OP_DEBUG)) {
    int qw = e->N >= 5000 ? QWARPS_MAX : 4;   // lockstep CTAs (shared instruction fetch)
    if (const char* s = getenv("PGTT_QUAD_WARPS")) { const int v = atoi(s); if (v >= 1 && v <= QWARPS_MAX) qw = v; }
    const int qwarps = (e->N + QENV - 1) / QENV, qblocks = (qwarps + qw - 1) / qw;
    const size_t qsmem = qw * ((sizeof(QShared) + 15) / 16 * 16);
    if (a.op == OP_STEP) pgtt_quad_kernel<OP_STEP><<<qblocks, 32 * qw, qsmem, st>>>(a);
    else pgtt_quad_kernel<OP_DEBUG><<<qblocks, 32 * qw, qsmem, st>>>(a);
  } else {
    const int wpb = a.op == OP_RESET ? (e->wpb < RESET_WARPS ? e->wpb : RESET_WARPS) : e->wpb;
    const int blocks = (e->N + wpb - 1) / wpb;
    const size_t smem = wpb * WS_BYTES;
    switch (a.op) {
      case OP_STEP: pgtt_env_kernel<OP_STEP><<<blocks, wpb * 32, smem, st>>>(a); break;
      case OP_STEP_TASK: pgtt_env_kernel<OP_STEP_TASK><<<blocks, wpb * 32, smem, st>>>(a); break;
      case OP_RESET: pgtt_env_kernel<OP_RESET><<<blocks, wpb * 32, smem + wpb * TASK_BYTES, st>>>(a); break;
      case OP_FORWARD: pgtt_env_kernel<OP_FORWARD><<<blocks, wpb * 32, smem, st>>>(a); break;
      default: pgtt_env_kernel<OP_DEBUG><<<blocks, wpb * 32, smem, st>>>(a); break;
    }
  }
  CUDA_OK(cudaGetLastError());
#else
  (void)stream;
  if (!task && e->quad && (a.op == OP_STEP || a.op == OP_DEBUG)) {
    const int nwarps = (e->N + QENV - 1) / QENV;
#pragma omp parallel
    {
      QShared* sh = (QShared*)aligned_alloc(16, (sizeof(QShared) + 15) / 16 * 16);
#pragma omp for schedule(dynamic, 1)
      for (int wi = 0; wi < nwarps; wi++) {
        memset(sh, 0xCD, sizeof(QShared));
        QuadJob j = {&a, wi * QENV, sh};
        emu_run_warp(quad_entry, &j);
      }
      free(sh);
    }
  } else {
#pragma omp parallel
    {
      WS* w = (WS*)aligned_alloc(16, WS_BYTES);
      TaskWS* t = (TaskWS*)aligned_alloc(16, TASK_BYTES);
#pragma omp for schedule(dynamic, 1)
      for (int env = 0; env < e->N; env++) {
        memset(w, 0xCD, WS_BYTES);  // poison: uninitialised reads show up as garbage, like on the GPU
        memset(t, 0xCD, TASK_BYTES);
        WarpJob j = {&a, env, w, t};
        emu_run_warp(warp_entry, &j);
      }
      free(w);
      free(t);
    }
  }
#endif
  e->launches++;
  return 0;
}

// ----------------------------------------------------------------------------------------------
// ABI
// ----------------------------------------------------------------------------------------------
extern "C" {

const char* pgtt_last_error(void) { return g_err.c_str(); }
int pgtt_version(void) { return 100; }

static void quat_to_mat_d(const double* q, double* m) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}

int pgtt_create(const pgtt_model_desc* m, const pgtt_task_desc* t, int device, int num_envs, pgtt_env** out) {
  if (!m || !t || !out || num_envs <= 0) return fail(PGTT_ERR_ARG, "pgtt_create: null argument or num_envs <= 0");
  if (m->n_boxes != 0 && m->n_boxes != NBOX) return fail(PGTT_ERR_ARG, "pgtt_create: n_boxes must be 0 (flat) or 100 (stairs)");
  if (t->n_substeps < 1 || t->n_substeps > 4) return fail(PGTT_ERR_ARG, "pgtt_create: n_substeps must be in 1..4");
  if (m->gravity[0] != 0 || m->gravity[1] != 0) return fail(PGTT_ERR_ARG, "pgtt_create: gravity must be along z");
  if (m->jnt_solimp[4] != 2.0 || m->foot_solimp[4] != 2.0 || m->floor_solimp[4] != 2.0 || (m->n_boxes && m->box_solimp[4] != 2.0))
    return fail(PGTT_ERR_ARG, "pgtt_create: only solimp power = 2 (the MuJoCo default) is supported");
#ifndef PGTT_HOST_EMU
  {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return fail(PGTT_ERR_CUDA, std::string("pgtt_create: no CUDA device: ") + cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(PGTT_ERR_ARG, "pgtt_create: bad device index");
    CUDA_OK(cudaSetDevice(device));
    const size_t smem = MAX_WARPS_PER_BLOCK * WS_BYTES;
    CUDA_OK(cudaFuncSetAttribute(pgtt_env_kernel<OP_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(pgtt_env_kernel<OP_STEP_TASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(pgtt_env_kernel<OP_RESET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RESET_WARPS * (WS_BYTES + TASK_BYTES))));
    CUDA_OK(cudaFuncSetAttribute(pgtt_env_kernel<OP_FORWARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(pgtt_env_kernel<OP_DEBUG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t qsmem = QWARPS_MAX * ((sizeof(QShared) + 15) / 16 * 16);
    CUDA_OK(cudaFuncSetAttribute(pgtt_quad_kernel<OP_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem));
    CUDA_OK(cudaFuncSetAttribute(pgtt_quad_kernel<OP_DEBUG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem));
  }
#endif
  pgtt_env* e = new pgtt_env();
  {
    DeviceConsts& d = g_dev[device >= 0 && device < PGTT_MAX_DEVICES ? device : 0];
    d.handles++;
#ifndef PGTT_HOST_EMU
    if (d.handles == 2) cudaDeviceSynchronize();   // launches of the first handle were not followed by event records so far
#endif
  }
  { static uint64_t next_serial = 1; e->serial = next_serial++; }
  e->device = device; e->N = num_envs; e->wpb = pick_warps_per_block(num_envs); e->terrain_dev = nullptr; e->n_terrains = 0; e->launches = 0; e->randomized = false;
  // Physics-kernel generation for step / debug-forward. Measured on B200 (profiles/r02): the warp-per-env kernel holds 28
  // envs per SM, so up to 148 x 28 = 4144 envs are ONE resident wave (0.37 ms); above that it needs a second wave (0.70 ms at
  // 8192) while the quad kernel (8 envs per warp, latency-bound at ~0.46 - 0.63 ms up to 8192 envs) does not.
  // PGTT_KERNEL=warp|quad overrides.
  {
    int n_sm = 148;
#ifndef PGTT_HOST_EMU
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
#endif
    e->quad = num_envs > n_sm * MAX_WARPS_PER_BLOCK * ENV_CTAS_PER_SM;
  }
  if (const char* k = getenv("PGTT_KERNEL")) e->quad = strcmp(k, "quad") == 0;
  // generation 1 runs the task layer behind the physics in the same launch: -2 % per step at 4096 envs (no second launch, no reload of
  // mjx.Data, SMs that finish their physics early start their task layer early; profiles/r02_summary.md)
  e->fuse_task = 1;
  if (const char* f = getenv("PGTT_FUSE_TASK")) e->fuse_task = atoi(f) != 0;
  ModelConst& c = e->mc;
  memset(&c, 0, sizeof(c));
  c.dt = (float)m->timestep; c.gravity_z = (float)m->gravity[2]; c.impratio = (float)m->impratio;
  c.tolerance = (float)m->tolerance; c.ls_tolerance = (float)m->ls_tolerance; c.meaninertia = (float)m->meaninertia;
  c.solver_scale = (float)(m->meaninertia * NV);
  // CTA barriers after collision and after the Newton solver: the stages whose duration varies between envs; in
  // between the warps of a CTA stay aligned by themselves (measured: profiles/r01b; r01e: the pre-solver barrier of the
  // earlier default costs 0.7 % - 0.5221 vs 0.5185 ms at 4096 envs, three interleaved runs each)
  c.sync_mask = (1 << ST_COLLIDE) | (1 << ST_POSTSOLVE);
  if (const char* sm = getenv("PGTT_SYNC_MASK")) c.sync_mask = (int)strtol(sm, nullptr, 0);
  if (const char* fs = getenv("PGTT_QUAD_FULLSCAN")) c.quad_fullscan = atoi(fs) != 0;
  if (const char* lv = getenv("PGTT_QUAD_LS_VOTE")) c.quad_ls_vote = atoi(lv) != 0;
  c.iterations = m->iterations; c.ls_iterations = m->ls_iterations; c.max_geom_pairs = m->max_geom_pairs;
  c.max_contact_points = m->max_contact_points; c.n_boxes = m->n_boxes; c.n_substeps = t->n_substeps;
  for (int b = 0; b < NB; b++) {
    double R[9];
    quat_to_mat_d(m->body_iquat[b + 1], R);
    const double* I = m->body_inertia[b + 1];
    double T[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[3 * i + j] = R[3 * i] * I[0] * R[3 * j] + R[3 * i + 1] * I[1] * R[3 * j + 1] + R[3 * i + 2] * I[2] * R[3 * j + 2];
    c.body_I[b][0] = (float)T[0]; c.body_I[b][1] = (float)T[4]; c.body_I[b][2] = (float)T[8];
    c.body_I[b][3] = (float)T[1]; c.body_I[b][4] = (float)T[2]; c.body_I[b][5] = (float)T[5];
    for (int i = 0; i < 3; i++) { c.body_pos[b][i] = (float)m->body_pos[b + 1][i]; c.body_ipos[b][i] = (float)m->body_ipos[b + 1][i]; }
    c.nom_mass[b] = (float)m->body_mass[b + 1];
  }
  for (int i = 0; i < 3; i++) c.nom_ipos_base[i] = (float)m->body_ipos[1][i];
  for (int j = 0; j < 12; j++) {
    c.jnt_lo[j] = (float)m->jnt_range[j][0]; c.jnt_hi[j] = (float)m->jnt_range[j][1];
    c.dof_invw[j] = (float)m->dof_invweight0[6 + j];
    c.nom_armature[j] = (float)m->dof_armature[6 + j]; c.nom_damping[j] = (float)m->dof_damping[6 + j];
    c.nom_qpos0[j] = (float)m->qpos0[7 + j];
    c.soft_lo[j] = (float)(m->jnt_range[j][0] * t->soft_limit_factor); c.soft_hi[j] = (float)(m->jnt_range[j][1] * t->soft_limit_factor);
    c.default_pose[j] = (float)t->default_pose[j];
  }
  for (int g = 0; g < 4; g++) c.calf_invw[g] = (float)m->body_invweight0[4 + 3 * g][0];
  for (int i = 0; i < 2; i++) c.lim_solref[i] = (float)m->jnt_solref[i];
  for (int i = 0; i < 5; i++) c.lim_solimp[i] = (float)m->jnt_solimp[i];
  for (int a = 0; a < 12; a++) {
    const int hinge = m->act_dof[a] - 6;
    if (hinge < 0 || hinge >= 12) { pgtt_destroy(e); return fail(PGTT_ERR_ARG, "pgtt_create: actuator must drive a hinge"); }
    c.hinge_of_act[a] = hinge; c.act_of_hinge[hinge] = a;
    c.nom_gain[a] = (float)m->act_gain[a]; c.act_bias0[a] = (float)m->act_bias[a][0]; c.nom_bias1[a] = (float)m->act_bias[a][1];
    c.act_bias2[a] = (float)m->act_bias[a][2];
    c.ctrl_lo[a] = (float)m->act_ctrlrange[a][0]; c.ctrl_hi[a] = (float)m->act_ctrlrange[a][1];
    c.frc_lo[a] = (float)m->act_forcerange[a][0]; c.frc_hi[a] = (float)m->act_forcerange[a][1];
  }
  for (int i = 0; i < 3; i++) { c.foot_pos[i] = (float)m->foot_pos[i]; c.imu_pos[i] = (float)m->imu_pos[i]; }
  c.foot_r = (float)m->foot_radius; c.foot_mu = (float)m->foot_friction[0];
  c.includemargin = (float)(m->foot_margin > 0 ? m->foot_margin : 0.0);
  for (int i = 0; i < 2; i++) { c.floor_solref[i] = (float)(0.5 * m->foot_solref[i] + 0.5 * m->floor_solref[i]); c.box_solref[i] = (float)(0.5 * m->foot_solref[i] + 0.5 * m->box_solref[i]); }
  for (int i = 0; i < 5; i++) { c.floor_solimp[i] = (float)(0.5 * m->foot_solimp[i] + 0.5 * m->floor_solimp[i]); c.box_solimp[i] = (float)(0.5 * m->foot_solimp[i] + 0.5 * m->box_solimp[i]); }
  for (int g = 0; g < 4; g++) c.foot_geom[g] = m->foot_geom_id[g];
  c.floor_geom = m->floor_geom_id; c.box_geom0 = m->box_geom_id0;
  c.box_rbound = (float)m->box_rbound;
  c.nom_box_mu = (float)m->box_friction[0]; c.nom_floor_mu = (float)m->floor_friction[0];
  c.n_model_bodies = m->n_model_bodies;
  c.ctrl_dt = (float)t->ctrl_dt; c.action_scale = (float)t->action_scale; c.noise_level = (float)t->noise_level;
  c.noise_joint_pos = (float)t->noise_joint_pos; c.noise_joint_vel = (float)t->noise_joint_vel; c.noise_gyro = (float)t->noise_gyro;
  c.noise_gravity = (float)t->noise_gravity; c.noise_linvel = (float)t->noise_linvel; c.noise_heightscan = (float)t->noise_heightscan;
  for (int k = 0; k < NREW; k++) c.reward_scale[k] = (float)t->reward_scale[k];
  c.tracking_sigma = (float)t->tracking_sigma; c.swing_height = (float)t->swing_height;
  c.base_feet_distance = (float)t->base_feet_distance; c.phase_sigma = (float)t->phase_sigma;
  for (int i = 0; i < 3; i++) { c.cmd_u_max[i] = (float)t->cmd_u_max[i]; c.cmd_u_min[i] = (float)t->cmd_u_min[i]; c.cmd_b[i] = (float)t->cmd_b[i]; }
  c.gait_freq[0] = (float)t->gait_freq[0]; c.gait_freq[1] = (float)t->gait_freq[1];
  for (int i = 0; i < NQ; i++) c.home_qpos[i] = (float)t->home_qpos[i];
  c.history_update_steps = t->history_update_steps; c.episode_length = t->episode_length; c.rng_partitionable = t->rng_partitionable;
  c.variant = t->variant ? 1 : 0; c.nobs = c.variant ? NOBS - 9 : NOBS; c.npriv = c.nobs + 44;
  // the kernels assume the GO2 tree: x-axis abduction, y-axis hip/knee, identity body quats (checked by the Python model compiler)

  EnvBuffers& B = e->B;
  memset(&B, 0, sizeof(B));
  B.N = num_envs;
  const size_t N = (size_t)num_envs;
  bool ok = true;
#define ALLOC(field, type, dim) ok = ok && ((B.field = (type*)dev_alloc(e, N * (dim) * sizeof(type))) != nullptr)
  ALLOC(qpos, float, NQ); ALLOC(qvel, float, NV); ALLOC(qacc, float, NV); ALLOC(warm, float, NV); ALLOC(ctrl, float, NU); ALLOC(time, float, 1);
  ALLOC(sensordata, float, NSENSOR); ALLOC(actuator_force, float, NU); ALLOC(site_xpos, float, 15); ALLOC(site_xmat, float, 9);
  ALLOC(contact_dist, float, NCON); ALLOC(contact_geom, int, 2 * NCON); ALLOC(solver_niter, int, 4);
  ALLOC(obs_state, float, NOBS); ALLOC(obs_priv, float, NPRIV); ALLOC(reward, float, 1); ALLOC(done, float, 1); ALLOC(metrics, float, NMETRIC);
  ALLOC(rng, uint32_t, 2); ALLOC(command, float, 3); ALLOC(step, int, 1); ALLOC(steps_until, int, 1);
  ALLOC(phase, float, 4); ALLOC(phase_dt, float, 1); ALLOC(gait_freq, float, 1); ALLOC(last_act, float, NU); ALLOC(last_last_act, float, NU);
  ALLOC(feet_air_time, float, 4); ALLOC(last_contact, int, 4); ALLOC(swing_peak, float, 4); ALLOC(H_max, float, 4); ALLOC(H_min, float, 4);
  ALLOC(heightscan, float, NRAY * 3); ALLOC(motor_targets, float, NU); ALLOC(qpos_err_hist, float, 24); ALLOC(qvel_hist, float, 24);
  ALLOC(contact, int, 4); ALLOC(first_contact, int, 4);
  ALLOC(steps, float, 1); ALLOC(truncation, float, 1); ALLOC(episode_done, float, 1); ALLOC(episode_metrics, float, 24);
  ALLOC(first_qpos, float, NQ); ALLOC(first_qvel, float, NV); ALLOC(first_warm, float, NV); ALLOC(first_qacc, float, NV);
  ALLOC(first_obs_state, float, NOBS); ALLOC(first_obs_priv, float, NPRIV);
  ALLOC(first_sensordata, float, NSENSOR); ALLOC(first_actuator_force, float, NU); ALLOC(first_site_xpos, float, 15); ALLOC(first_site_xmat, float, 9);
  ALLOC(first_contact_dist, float, NCON); ALLOC(first_contact_geom, int, 2 * NCON);
  ALLOC(m_mass, float, NB); ALLOC(m_ipos, float, 3); ALLOC(m_armature, float, 12); ALLOC(m_damping, float, 12); ALLOC(m_gain, float, 12);
  ALLOC(m_bias1, float, 12); ALLOC(m_qpos0, float, 12); ALLOC(m_boxfric, float, NBOX); ALLOC(m_floorfric, float, 1); ALLOC(terrain_index, int, 1);
#undef ALLOC
  if (!ok) { pgtt_destroy(e); return fail(PGTT_ERR_NOMEM, "pgtt_create: device allocation failed"); }
  *out = e;
  // nominal per-env model so that an un-randomised env (flat task, no DR) is immediately usable
  {
    uint32_t* dk = (uint32_t*)dev_alloc(e, 2 * N * sizeof(uint32_t));
    if (!dk) { pgtt_destroy(e); return fail(PGTT_ERR_NOMEM, "pgtt_create: device allocation failed"); }
    if (m->n_boxes == 0) {
      int rc = pgtt_randomize(e, dk, 0, nullptr);
      if (rc) { pgtt_destroy(e); return rc; }
      rc = pgtt_sync(e, nullptr);
      if (rc) { pgtt_destroy(e); return rc; }
    }
  }
  return PGTT_OK;
}

int pgtt_destroy(pgtt_env* e) {
  if (!e) return PGTT_OK;
  {
    DeviceConsts& d = dev_consts(e);
    if (d.owner == e) d.owner = nullptr;
    if (d.handles > 0) d.handles--;
  }
#ifndef PGTT_HOST_EMU
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (void* p : e->allocs) cudaFree(p);
  if (e->terrain_dev) cudaFree(e->terrain_dev);
#else
  for (void* p : e->allocs) free(p);
  if (e->terrain_dev) free(e->terrain_dev);
#endif
  delete e;
  return PGTT_OK;
}

int pgtt_sync(pgtt_env* e, void* stream) {
  if (!e) return fail(PGTT_ERR_ARG, "pgtt_sync: null handle");
#ifndef PGTT_HOST_EMU
  CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
#else
  (void)stream;
#endif
  return PGTT_OK;
}

int pgtt_set_terrain_table(pgtt_env* e, const float* boxes, int T) {
  if (!e || !boxes || T <= 0) return fail(PGTT_ERR_ARG, "pgtt_set_terrain_table: null argument or T <= 0");
  if (e->mc.n_boxes == 0) return fail(PGTT_ERR_STATE, "pgtt_set_terrain_table: the flat_terrain scene has no boxes");
  std::vector<float> pre((size_t)T * NBOX * BOXF);
  for (size_t i = 0; i < (size_t)T * NBOX; i++) {
    const float* b = boxes + i * 10;
    if (b[4] != 0.f || b[5] != 0.f) return fail(PGTT_ERR_ARG, "pgtt_set_terrain_table: only yaw-rotated boxes (quat x = y = 0) are supported");
    const double n = sqrt((double)b[3] * b[3] + (double)b[6] * b[6]);
    if (n < 1e-9) return fail(PGTT_ERR_ARG, "pgtt_set_terrain_table: zero quaternion");
    const double w = b[3] / n, z = b[6] / n;
    float* o = &pre[i * BOXF];
    o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[7]; o[4] = b[8]; o[5] = b[9];
    o[6] = (float)(w * w - z * z); o[7] = (float)(2.0 * w * z);
  }
#ifndef PGTT_HOST_EMU
  CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaDeviceSynchronize());
  if (e->terrain_dev) CUDA_OK(cudaFree(e->terrain_dev));
  CUDA_OK(cudaMalloc((void**)&e->terrain_dev, pre.size() * sizeof(float)));
  CUDA_OK(cudaMemcpy(e->terrain_dev, pre.data(), pre.size() * sizeof(float), cudaMemcpyHostToDevice));
#else
  if (e->terrain_dev) free(e->terrain_dev);
  e->terrain_dev = (float*)malloc(pre.size() * sizeof(float));
  memcpy(e->terrain_dev, pre.data(), pre.size() * sizeof(float));
#endif
  e->n_terrains = T;
  e->B.terrain = e->terrain_dev;
  e->B.n_terrains = T;
  e->randomized = false;
  return PGTT_OK;
}

int pgtt_randomize(pgtt_env* e, const uint32_t* keys, int dynamics, void* stream) {
  if (!e || !keys) return fail(PGTT_ERR_ARG, "pgtt_randomize: null argument");
  if (e->mc.n_boxes > 0 && !e->terrain_dev) return fail(PGTT_ERR_STATE, "pgtt_randomize: stairs task needs pgtt_set_terrain_table first");
  if (int rc = make_resident(e, stream)) return rc;
#ifndef PGTT_HOST_EMU
  pgtt_randomize_kernel<<<(e->N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->B, keys, dynamics);
  CUDA_OK(cudaGetLastError());
  if (int rc = mark_launched(e, stream)) return rc;
#else
  (void)stream;
#pragma omp parallel for
  for (int env = 0; env < e->N; env++) env_randomize(e->B, keys, env, dynamics);
#endif
  e->launches++;
  e->randomized = true;
  return PGTT_OK;
}

static int check_ready(pgtt_env* e, const char* who) {
  if (!e) return fail(PGTT_ERR_ARG, std::string(who) + ": null handle");
  if (e->mc.n_boxes > 0 && !e->randomized)
    return fail(PGTT_ERR_STATE, std::string(who) + ": stairs task needs pgtt_set_terrain_table + pgtt_randomize first (terrain only enters through the randomiser)");
  return 0;
}

int pgtt_reset(pgtt_env* e, const uint32_t* keys, void* stream) {
  if (int rc = check_ready(e, "pgtt_reset")) return rc;
  if (!keys) return fail(PGTT_ERR_ARG, "pgtt_reset: null keys");
  LaunchArgs a; memset(&a, 0, sizeof(a));
  a.op = OP_RESET; a.keys = keys;
  if (int rc = launch(e, a, stream)) return rc;
  return mark_launched(e, stream);
}

int pgtt_step_record(pgtt_env* e, const float* action, int wrapped, float* os, float* op, float* rw, float* dc, float* tr, void* stream) {
  if (int rc = check_ready(e, "pgtt_step")) return rc;
  if (!action) return fail(PGTT_ERR_ARG, "pgtt_step: null action");
  LaunchArgs a; memset(&a, 0, sizeof(a));
  a.op = OP_STEP; a.action = action; a.wrapped = wrapped;
  a.rec.obs_state = os; a.rec.obs_priv = op; a.rec.reward = rw; a.rec.discount = dc; a.rec.truncation = tr;
  if (!e->quad && e->fuse_task) {                 // generation 1: physics + task layer in one launch
    a.op = OP_STEP_TASK;
    if (int rc = launch(e, a, stream)) return rc;
    return mark_launched(e, stream);
  }
  if (int rc = launch(e, a, stream)) return rc;   // physics: n_substeps x mjx.step
  a.op = OP_TASK;
  if (int rc = launch(e, a, stream)) return rc;   // task layer, wrappers, transition slot
  return mark_launched(e, stream);
}

int pgtt_step(pgtt_env* e, const float* action, int wrapped, void* stream) {
  return pgtt_step_record(e, action, wrapped, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

int pgtt_forward(pgtt_env* e, void* stream) {
  if (int rc = check_ready(e, "pgtt_forward")) return rc;
  LaunchArgs a; memset(&a, 0, sizeof(a));
  a.op = OP_FORWARD;
  if (int rc = launch(e, a, stream)) return rc;
  return mark_launched(e, stream);
}

int pgtt_heightscan(pgtt_env* e, const float* center, const float* yaw, float* out, void* stream) {
  if (int rc = check_ready(e, "pgtt_heightscan")) return rc;
  if (!center || !yaw || !out) return fail(PGTT_ERR_ARG, "pgtt_heightscan: null argument");
  LaunchArgs a; memset(&a, 0, sizeof(a));
  a.op = OP_SCAN; a.center = center; a.yaw = yaw; a.out = out;
  if (int rc = launch(e, a, stream)) return rc;
  return mark_launched(e, stream);
}

int pgtt_debug_forward(pgtt_env* e, float* out, void* stream) {
  if (int rc = check_ready(e, "pgtt_debug_forward")) return rc;
  if (!out) return fail(PGTT_ERR_ARG, "pgtt_debug_forward: null out");
  LaunchArgs a; memset(&a, 0, sizeof(a));
  a.op = OP_DEBUG; a.out = out;
  if (int rc = launch(e, a, stream)) return rc;
  return mark_launched(e, stream);
}

int pgtt_get_buffers(pgtt_env* e, pgtt_buffers* o) {
  if (!e || !o) return fail(PGTT_ERR_ARG, "pgtt_get_buffers: null argument");
  const EnvBuffers& B = e->B;
  memset(o, 0, sizeof(*o));
  o->num_envs = e->N;
  o->qpos = B.qpos; o->qvel = B.qvel; o->qacc = B.qacc; o->qacc_warmstart = B.warm; o->ctrl = B.ctrl; o->time = B.time;
  o->sensordata = B.sensordata; o->actuator_force = B.actuator_force; o->site_xpos = B.site_xpos; o->site_xmat = B.site_xmat;
  o->contact_dist = B.contact_dist; o->contact_geom = B.contact_geom; o->solver_niter = B.solver_niter;
  o->obs_state = B.obs_state; o->obs_privileged = B.obs_priv; o->reward = B.reward; o->done = B.done; o->metrics = B.metrics;
  o->rng = B.rng; o->command = B.command; o->step = B.step; o->steps_until_next_cmd = B.steps_until;
  o->phase = B.phase; o->phase_dt = B.phase_dt; o->gait_freq = B.gait_freq; o->last_act = B.last_act; o->last_last_act = B.last_last_act;
  o->feet_air_time = B.feet_air_time; o->last_contact = B.last_contact; o->swing_peak = B.swing_peak; o->H_max = B.H_max; o->H_min = B.H_min;
  o->heightscan = B.heightscan; o->motor_targets = B.motor_targets; o->qpos_error_history = B.qpos_err_hist; o->qvel_history = B.qvel_hist;
  o->contact = B.contact; o->first_contact = B.first_contact;
  o->steps = B.steps; o->truncation = B.truncation; o->episode_done = B.episode_done; o->episode_metrics = B.episode_metrics;
  o->first_qpos = B.first_qpos; o->first_qvel = B.first_qvel; o->first_qacc_warmstart = B.first_warm;
  o->first_obs_state = B.first_obs_state; o->first_obs_privileged = B.first_obs_priv;
  o->body_mass = B.m_mass; o->body_ipos_base = B.m_ipos; o->dof_armature = B.m_armature; o->dof_damping = B.m_damping;
  o->actuator_gain = B.m_gain; o->actuator_bias1 = B.m_bias1; o->qpos0 = B.m_qpos0; o->box_friction = B.m_boxfric; o->floor_friction = B.m_floorfric;
  o->terrain_index = B.terrain_index;
  return PGTT_OK;
}

int64_t pgtt_launch_count(pgtt_env* e) { return e ? e->launches : 0; }
int pgtt_obs_dims(pgtt_env* e, int* nobs, int* npriv) {
  if (!e) return fail(PGTT_ERR_ARG, "pgtt_obs_dims: null handle");
  if (nobs) *nobs = e->mc.nobs;
  if (npriv) *npriv = e->mc.npriv;
  return PGTT_OK;
}
int pgtt_step_kernel_generation(pgtt_env* e) { return e ? e->quad : -1; }
int pgtt_step_launches(pgtt_env* e) { return e ? ((!e->quad && e->fuse_task) ? 1 : 2) : -1; }
// graph replays of the rollout launch this handle's kernels without going through launch(): keep the counter honest
void pgtt_internal_count_launches(pgtt_env* e, int64_t n) { if (e) e->launches += n; }

// development aid (tools/stage_trace.py): one warp-per-env physics launch that records the SM clock of every warp after each
// stage of every substep into trace[num_envs][40] (DEVICE, int64)
int pgtt_internal_physics_trace(pgtt_env* e, const float* action, long long* trace, void* stream) {
  if (int rc = check_ready(e, "pgtt_internal_physics_trace")) return rc;
  if (e->quad) return fail(PGTT_ERR_STATE, "pgtt_internal_physics_trace: warp-per-env kernel only");
  LaunchArgs a; memset(&a, 0, sizeof(a));
  a.op = OP_STEP; a.action = action; a.out = reinterpret_cast<float*>(trace);
  if (int rc = launch(e, a, stream)) return rc;
  return mark_launched(e, stream);
}
// development / profiling aid (tools/kernel_times.py): one half of pgtt_step. part 0 = physics kernel, 1 = task kernel
int pgtt_internal_step_part(pgtt_env* e, const float* action, int wrapped, int part, void* stream) {
  if (int rc = check_ready(e, "pgtt_internal_step_part")) return rc;
  LaunchArgs a; memset(&a, 0, sizeof(a));
  a.op = part ? OP_TASK : OP_STEP; a.action = action; a.wrapped = wrapped;
  if (int rc = launch(e, a, stream)) return rc;
  return mark_launched(e, stream);
}
uint64_t pgtt_internal_serial(pgtt_env* e) { return e ? e->serial : 0; }
int pgtt_internal_make_resident(pgtt_env* e, void* stream) { return e ? make_resident(e, stream) : PGTT_OK; }
int pgtt_internal_mark_launched(pgtt_env* e, void* stream) { return e ? mark_launched(e, stream) : PGTT_OK; }

// Files the CURRENT observation / reward / done into a rollout slot with plain device copies (the per-step slots of an
// unroll are written by the task kernel itself, pgtt_step_record; this entry point fills slot 0 of an unroll).
int pgtt_record(pgtt_env* e, float* os, float* op, float* rw, float* dc, float* tr, void* stream) {
  if (!e) return fail(PGTT_ERR_ARG, "pgtt_record: null handle");
  const EnvBuffers& B = e->B;
  const size_t N = (size_t)B.N;
#ifndef PGTT_HOST_EMU
  cudaStream_t st = (cudaStream_t)stream;
  if (os) CUDA_OK(cudaMemcpyAsync(os, B.obs_state, N * e->mc.nobs * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (op) CUDA_OK(cudaMemcpyAsync(op, B.obs_priv, N * e->mc.npriv * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (rw) CUDA_OK(cudaMemcpyAsync(rw, B.reward, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (tr) CUDA_OK(cudaMemcpyAsync(tr, B.truncation, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (dc) {
    pgtt_discount_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(B.done, dc, (int)N);
    CUDA_OK(cudaGetLastError());
    e->launches++;
  }
#else
  (void)stream;
  if (os) memcpy(os, B.obs_state, N * e->mc.nobs * sizeof(float));
  if (op) memcpy(op, B.obs_priv, N * e->mc.npriv * sizeof(float));
  for (size_t i = 0; i < N; i++) {
    if (rw) rw[i] = B.reward[i];
    if (dc) dc[i] = 1.0f - B.done[i];
    if (tr) tr[i] = B.truncation[i];
  }
#endif
  return PGTT_OK;
}

}  // extern "C"
