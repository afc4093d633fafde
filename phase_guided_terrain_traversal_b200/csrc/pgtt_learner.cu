// pgtt_learner.cu - dense layers of the PPO learner on the 5th-gen tensor cores (sm_100a), hand-written.
//
// brax's ppo.train (as driven by training/train.py:135-161,242-263) evaluates two small MLPs, policy 171-512-256-128-24 and
// value 215-512-256-128-1, forward and backward on minibatches of 5120 transitions, 4 x 32 times per training step, at
// `jax_default_matmul_precision=highest` (train.py:93-94): fp32-accurate products. The three GEMM shapes of a dense layer
//     forward      Y [M, N]  = X [M, K] W [K, N] + b
//     input grad   dX [M, K] = dY [M, N] W^T
//     weight grad  dW [K, N] = X^T dY            (reduction over the M = 5120 rows, split over CTAs), db = column sums of dY
// all run through ONE kernel: C [Mc, Nc] = A B^T with both operands addressed by (row, k) strides, so a transposed view costs
// nothing - the loader threads read fp32 from global memory with lanes along whichever index is contiguous and write the
// K-major canonical (no-swizzle) UMMA layout to shared memory themselves (no TMA: the operands change every SGD step
// and need a precision split on the way in).
//
// Precision: every fp32 operand is split x = hi + lo with hi = bf16(x), lo = bf16(x - hi), and a k-slice issues three
// `tcgen05.mma.kind::f16` (hi hi + hi lo + lo hi, fp32 accumulation in TMEM): products carry ~16 mantissa bits, i.e.
// relative error ~2^-17 per product - two orders below the 1e-4 gradient-parity bar of tests/test_ppo.py and far inside
// what Adam's 1e-8 epsilon and PPO's noise can see - at 3 bf16 MMAs per slice (= 1.5 tf32 MMAs) instead of an fp32 SIMT GEMM.
//
// One CTA = one 128 x 128 output tile (x one K split): 256 threads stage 32-wide K chunks (A, B as hi / lo: 32 KB per stage,
// 3 stages, two CTAs per SM) while thread 0 issues the MMAs of the previous chunk; accumulators live in 128 TMEM columns;
// the epilogue reads them with tcgen05.ld, transposes the tile through shared memory (the pipeline stages are free by
// then) and stores fp32 rows fully coalesced with bias / SiLU applied (or the split's partial tile; a second small kernel
// sums the splits in a fixed order - deterministic - and forms db).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/pgtt_b200.h"

#define LG_TM 128
#define LG_TN 128
#define LG_KC 32                         // k per stage
#define LG_STAGES 3                      // 3 x 32 KB: two CTAs per SM (a 5120 x 512 layer is 160 tiles: one wave)
#define LG_THREADS 256
#define LG_OP_BYTES (LG_TM * LG_KC * 2)   // one operand, one precision part: 16 KB
#define LG_STAGE_BYTES (4 * LG_OP_BYTES)  // A hi, A lo, B hi, B lo
#define LG_SMEM (LG_STAGES * LG_STAGE_BYTES + 256)

static thread_local std::string g_lerr;
static int lfail(int code, const std::string& m) { g_lerr = m; return code; }
#define LCUDA(call)                                                                                        \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) return lfail(PGTT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  for (long it = 0; it < (1L << 28); it++) {   // bounded spin: a protocol bug traps instead of hanging the GPU
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle: start >> 4 | LBO >> 4 @16 (next core matrix along K) |
// SBO >> 4 @32 (next 8-row group) | version 1 @46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: D fp32 @4, A / B bf16 @7 / @10, both K-major, N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

struct GemmArgs {
  const float* A;      // element (row, k) at A[row * a_row + k * a_k]
  const float* B;      // element (row, k) at B[row * b_row + k * b_k]   (C = A B^T over k)
  float* C;            // [Mc][ldc] (k_splits == 1) or partial tiles [split][Mc][ldc]
  const float* bias;   // [Nc] added in the epilogue (k_splits == 1 only), may be null
  int Mc, Nc, K, ldc;
  long long a_row, a_k, b_row, b_k;
  int k_splits, k_per_split;   // K range of split z: [z * k_per_split, min(K, (z + 1) * k_per_split))
  int act;             // 1: epilogue writes the pre-activation to Z and SiLU of it to C (forward of a hidden layer);
                       // 2: epilogue multiplies by SiLU'(Z) (input gradient flowing into the SiLU that produced this layer's input)
  float* Z;            // [Mc][ldc] pre-activations (written for act == 1, read for act == 2)
  float* colsum;       // [k_splits][Nc]: per-split sums over k of every B row (db of the weight-gradient GEMM), may be null
};

// eight consecutive k of one operand row: global loads only (issued one chunk ahead of their use). `p` points at the
// octet; `mode`: 0 = row out of range (zeros), 1 = k contiguous and 16-byte aligned (two vector loads), 2 = strided
__device__ __forceinline__ void load_octet(float* v, const float* __restrict__ p, long long k_stride, int mode, int k_left) {
  if (mode == 0 || k_left <= 0) {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 0.f;
  } else if (k_left >= 8) {
    if (mode == 1) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = __ldg(p + (long long)i * k_stride);
    }
  } else {     // K tail
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (i < k_left) ? __ldg(p + (long long)i * k_stride) : 0.f;
  }
}
// ... -> hi / lo bf16 octets in the canonical K-major layout of a [128 rows][LG_KC] tile: 8-row groups LG_KC * 16 B apart,
// core matrices (8 k) 128 B apart
__device__ __forceinline__ void store_octet(const float* v, uint8_t* dst_hi, uint8_t* dst_lo, int r_local, int kg) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
  }
  const uint32_t off = (uint32_t)(r_local >> 3) * (LG_KC * 16u) + (uint32_t)kg * 128u + (uint32_t)(r_local & 7) * 16u;
  *reinterpret_cast<uint4*>(dst_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(dst_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(LG_THREADS, 2) pgtt_gemm_kernel(GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + LG_STAGES * LG_STAGE_BYTES);   // [LG_STAGES]: the MMAs that read stage s are complete
  uint64_t* all_done = mma_done + LG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(all_done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * LG_TM, n0 = blockIdx.y * LG_TN, split = blockIdx.z;
  const int k_begin = split * g.k_per_split, k_end = min(g.K, k_begin + g.k_per_split);
  const int nchunks = (k_end - k_begin + LG_KC - 1) / LG_KC;

  if (tid == 0) {
    for (int i = 0; i < LG_STAGES; i++) mbar_init(&mma_done[i], 1);
    mbar_init(all_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(LG_TN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = make_idesc(LG_TM, LG_TN);

  // thread tid stages the octets (row r = tid & 127, k group kg = tid >> 7 and kg + 2) of both operands; their raw values
  // are fetched one chunk ahead (registers), so the global-load latency overlaps the barrier and the MMAs of the chunk before
  constexpr int NOCT = (LG_TM * (LG_KC / 8)) / LG_THREADS;   // 2
  const int r_st = tid & (LG_TM - 1), kg_st = tid >> 7;
  float pa[2][NOCT][8], pb[2][NOCT][8];   // two chunks in flight
  float colsum = 0.f;    // weight-gradient GEMM: running sum of this thread's B row (= a column of dY) -> db
  // per-thread operand cursors (octet 0 of chunk 0); they advance by LG_KC k per chunk
  const float* cA = g.A + (long long)(m0 + r_st) * g.a_row + (long long)(k_begin + kg_st * 8) * g.a_k;
  const float* cB = g.B + (long long)(n0 + r_st) * g.b_row + (long long)(k_begin + kg_st * 8) * g.b_k;
  const int modeA = (m0 + r_st >= g.Mc) ? 0 : ((g.a_k == 1 && ((reinterpret_cast<uintptr_t>(cA) | (uintptr_t)(g.a_row * 4)) & 15) == 0) ? 1 : 2);
  const int modeB = (n0 + r_st >= g.Nc) ? 0 : ((g.b_k == 1 && ((reinterpret_cast<uintptr_t>(cB) | (uintptr_t)(g.b_row * 4)) & 15) == 0) ? 1 : 2);
  const long long stepA = 16 * g.a_k, stepB = 16 * g.b_k;     // second octet of the thread: k group + 2
#define LG_PREFETCH(c, buf)                                                                                                   \
  {                                                                                                                           \
    const int kl_ = k_end - (k_begin + (c) * LG_KC + kg_st * 8);                                                             \
    const float* a_ = cA + (long long)(c) * LG_KC * g.a_k;                                                                   \
    const float* b_ = cB + (long long)(c) * LG_KC * g.b_k;                                                                   \
    _Pragma("unroll") for (int i = 0; i < NOCT; i++) {                                                                       \
      load_octet(pa[buf][i], a_ + i * stepA, g.a_k, modeA, kl_ - 16 * i);                                                     \
      load_octet(pb[buf][i], b_ + i * stepB, g.b_k, modeB, kl_ - 16 * i);                                                     \
    }                                                                                                                         \
  }
#define LG_CONSUME(buf)                                                                                                       \
  _Pragma("unroll") for (int i = 0; i < NOCT; i++) {                                                                         \
    store_octet(pa[buf][i], st, st + LG_OP_BYTES, r_st, kg_st + 2 * i);                                                       \
    store_octet(pb[buf][i], st + 2 * LG_OP_BYTES, st + 3 * LG_OP_BYTES, r_st, kg_st + 2 * i);                                 \
    if (g.colsum) { _Pragma("unroll") for (int j = 0; j < 8; j++) colsum += pb[buf][i][j]; }                                 \
  }
  if (nchunks > 0) LG_PREFETCH(0, 0);
  if (nchunks > 1) LG_PREFETCH(1, 1);
  for (int c = 0; c < nchunks; c++) {
    const int s = c % LG_STAGES;
    uint8_t* st = smem + (size_t)s * LG_STAGE_BYTES;
    if (c >= LG_STAGES) mbar_wait(&mma_done[s], (uint32_t)((c / LG_STAGES - 1) & 1));   // the MMAs of chunk c - LG_STAGES released this stage
    if (c & 1) {          // (static register indices: the two buffers are distinct register sets)
      LG_CONSUME(1);
      if (c + 2 < nchunks) LG_PREFETCH(c + 2, 1);
    } else {
      LG_CONSUME(0);
      if (c + 2 < nchunks) LG_PREFETCH(c + 2, 0);
    }
    fence_proxy_async();      // generic-proxy stores -> visible to the tensor-core (async) proxy
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + LG_OP_BYTES, b_hi = a_hi + 2 * LG_OP_BYTES, b_lo = a_hi + 3 * LG_OP_BYTES;
#pragma unroll
      for (int ks = 0; ks < LG_KC / 16; ks++) {
        const uint32_t o = (uint32_t)ks * 256u;   // two core matrices (16 k) per slice
        const uint64_t dah = make_desc(a_hi + o, 128u, LG_KC * 16u), dal = make_desc(a_lo + o, 128u, LG_KC * 16u);
        const uint64_t dbh = make_desc(b_hi + o, 128u, LG_KC * 16u), dbl = make_desc(b_lo + o, 128u, LG_KC * 16u);
        umma_bf16(tmem, dal, dbh, idesc, (uint32_t)(c > 0 || ks > 0));   // small terms first
        umma_bf16(tmem, dah, dbl, idesc, 1u);
        umma_bf16(tmem, dah, dbh, idesc, 1u);
      }
      umma_commit(&mma_done[s]);
      if (c == nchunks - 1) umma_commit(all_done);
    }
  }
  if (g.colsum && blockIdx.x == 0) {   // db partial of this split: the two threads of a row meet in shared memory after the last MMA
    mbar_wait(all_done, 0);
    float* cs = reinterpret_cast<float*>(smem + LG_STAGES * LG_STAGE_BYTES - 1024);   // last KB of the stage area (free: every MMA completed)
    if (tid >= LG_TM) cs[tid - LG_TM] = colsum;
    __syncthreads();
    if (tid < LG_TM && n0 + tid < g.Nc) g.colsum[(size_t)split * g.Nc + n0 + tid] = colsum + cs[tid];
    __syncthreads();
  }
  mbar_wait(all_done, 0);
  tc_fence_after();
  // epilogue: warp w reads TMEM lanes 32 (w & 3) .. (= tile rows) and the column half w >> 2, 32 columns at a time; a thread
  // owns 32 consecutive columns of its row and writes them as 16-byte vectors (bias / SiLU / SiLU' applied in registers)
  {
    const int q = warp & 3, half = warp >> 2;
    const int row = m0 + q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    float* Cb = g.C + (g.k_splits > 1 ? (size_t)split * g.Mc * g.ldc : 0);
    const bool add_bias = g.bias && g.k_splits == 1;
    const bool vec = (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(Cb) | (g.Z ? reinterpret_cast<uintptr_t>(g.Z) : 0)) & 15) == 0;
#pragma unroll 1
    for (int cb = half * 64; cb < half * 64 + 64; cb += 32) {
      if (n0 + cb >= g.Nc) break;       // (warp-uniform)
      float v[32];
      tmem_ld32(trow + (uint32_t)cb, v);
      if (row < g.Mc) {
        float* crow = Cb + (size_t)row * g.ldc + n0 + cb;
        float* zrow = g.Z ? g.Z + (size_t)row * g.ldc + n0 + cb : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int col = n0 + cb + j;
          if (col >= g.Nc) break;
          const bool full = vec && col + 3 < g.Nc;
          float x[4] = {v[j], v[j + 1], v[j + 2], v[j + 3]};
          if (add_bias) {
#pragma unroll
            for (int t = 0; t < 4; t++) if (col + t < g.Nc) x[t] += __ldg(g.bias + col + t);
          }
          if (g.act == 1) {            // forward of a hidden layer: keep the pre-activation, emit SiLU
            if (full) *reinterpret_cast<float4*>(zrow + j) = make_float4(x[0], x[1], x[2], x[3]);
            else { for (int t = 0; t < 4; t++) if (col + t < g.Nc) zrow[j + t] = x[t]; }
#pragma unroll
            for (int t = 0; t < 4; t++) x[t] = x[t] / (1.f + __expf(-x[t]));
          } else if (g.act == 2) {     // input gradient that feeds a SiLU: multiply by SiLU'(pre-activation of that input)
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            if (full) { const float4 zz = *reinterpret_cast<const float4*>(zrow + j); z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w; }
            else { for (int t = 0; t < 4; t++) if (col + t < g.Nc) z[t] = zrow[j + t]; }
#pragma unroll
            for (int t = 0; t < 4; t++) { const float sg = 1.f / (1.f + __expf(-z[t])); x[t] *= sg * (1.f + z[t] * (1.f - sg)); }
          }
          if (full) *reinterpret_cast<float4*>(crow + j) = make_float4(x[0], x[1], x[2], x[3]);
          else { for (int t = 0; t < 4; t++) if (col + t < g.Nc) crow[j + t] = x[t]; }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(LG_TN));
}

// dW = sum of the split partials in split order (deterministic), db[n] = sum over rows of dY[:, n]
__global__ void pgtt_splitsum_kernel(const float* __restrict__ part, float* __restrict__ out, int n_elem, int splits,
                                     const float* __restrict__ bpart, float* __restrict__ bout, int n_bias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_elem) {
    float s = 0.f;
    for (int z = 0; z < splits; z++) s += part[(size_t)z * n_elem + i];
    out[i] = s;
  } else if (bout && i - n_elem < n_bias) {
    const int j = i - n_elem;
    float s = 0.f;
    for (int z = 0; z < splits; z++) s += bpart[(size_t)z * n_bias + j];
    bout[j] = s;
  }
}
// dZ = dY * silu'(Z) in place of dY's consumer (backward of the fused forward activation)
__global__ void pgtt_silu_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Z, float* __restrict__ dZ, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float z = Z[i], sg = 1.f / (1.f + __expf(-z));
    dZ[i] = dY[i] * (sg * (1.f + z * (1.f - sg)));
  }
}

int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    LCUDA(cudaFuncSetAttribute(pgtt_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_SMEM));
    attr = true;
  }
  dim3 grid((g.Mc + LG_TM - 1) / LG_TM, (g.Nc + LG_TN - 1) / LG_TN, g.k_splits);
  pgtt_gemm_kernel<<<grid, LG_THREADS, LG_SMEM, st>>>(g);
  LCUDA(cudaGetLastError());
  return PGTT_OK;
}

}  // namespace

extern "C" {

const char* pgtt_learner_last_error(void) { return g_lerr.c_str(); }

int pgtt_linear_forward(const float* x, int ldx, const float* w, const float* b, int M, int K, int N, int silu, float* y, float* z, void* stream) {
  if (!x || !w || !y || M <= 0 || K <= 0 || N <= 0 || ldx < K || (silu && !z)) return lfail(PGTT_ERR_ARG, "pgtt_linear_forward: bad argument");
  GemmArgs g = {};
  g.A = x; g.a_row = ldx; g.a_k = 1;
  g.B = w; g.b_row = 1; g.b_k = N;          // B(row n, k) = W[k][n]
  g.C = y; g.bias = b; g.Mc = M; g.Nc = N; g.K = K; g.ldc = N; g.k_splits = 1; g.k_per_split = K; g.act = silu ? 1 : 0; g.Z = z;
  return launch_gemm(g, (cudaStream_t)stream);
}

int pgtt_linear_backward_input(const float* dy, const float* w, int M, int K, int N, float* dx, int lddx, const float* z_in, void* stream) {
  if (!dy || !w || !dx || M <= 0 || K <= 0 || N <= 0 || lddx < K) return lfail(PGTT_ERR_ARG, "pgtt_linear_backward_input: bad argument");
  GemmArgs g = {};
  g.A = dy; g.a_row = N; g.a_k = 1;          // reduce over the N outputs
  g.B = w; g.b_row = N; g.b_k = 1;           // B(row k_in, n) = W[k_in][n]
  g.C = dx; g.bias = nullptr; g.Mc = M; g.Nc = K; g.K = N; g.ldc = lddx; g.k_splits = 1; g.k_per_split = N;
  if (z_in) { g.act = 2; g.Z = const_cast<float*>(z_in); }   // z_in [M][lddx]: pre-activation whose SiLU is this layer's input
  return launch_gemm(g, (cudaStream_t)stream);
}

int pgtt_linear_backward_params_splits(int M) { const int s = (M + 511) / 512; return s < 1 ? 1 : (s > 16 ? 16 : s); }
// floats of scratch pgtt_linear_backward_params needs: split partials of dW and of db
long long pgtt_linear_backward_params_scratch(int M, int K, int N) { return (long long)pgtt_linear_backward_params_splits(M) * ((long long)K * N + N); }

int pgtt_linear_backward_params(const float* x, int ldx, const float* dy, int M, int K, int N, float* dw, float* db, float* scratch, void* stream) {
  if (!x || !dy || !dw || !scratch || M <= 0 || K <= 0 || N <= 0 || ldx < K) return lfail(PGTT_ERR_ARG, "pgtt_linear_backward_params: bad argument");
  const int splits = pgtt_linear_backward_params_splits(M);
  GemmArgs g = {};
  g.A = x; g.a_row = 1; g.a_k = ldx;         // A(row k_in, m) = X[m][k_in]
  g.B = dy; g.b_row = 1; g.b_k = N;          // B(row n, m) = dY[m][n]
  g.C = scratch; g.bias = nullptr; g.Mc = K; g.Nc = N; g.K = M; g.ldc = N; g.k_splits = splits;
  g.k_per_split = ((M + splits - 1) / splits + LG_KC - 1) / LG_KC * LG_KC;
  const int n_elem = K * N;
  g.colsum = db ? scratch + (size_t)splits * n_elem : nullptr;      // db partials behind the dW partials
  if (int rc = launch_gemm(g, (cudaStream_t)stream)) return rc;
  pgtt_splitsum_kernel<<<(n_elem + N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scratch, dw, n_elem, splits, g.colsum, db, N);
  LCUDA(cudaGetLastError());
  return PGTT_OK;
}

int pgtt_silu_backward(const float* dy, const float* z, float* dz, long long n, void* stream) {
  if (!dy || !z || !dz || n <= 0) return lfail(PGTT_ERR_ARG, "pgtt_silu_backward: bad argument");
  pgtt_silu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dy, z, dz, (size_t)n);
  LCUDA(cudaGetLastError());
  return PGTT_OK;
}

}  // extern "C"
