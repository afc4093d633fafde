// pgtt_mlp.cu - forward and backward of the PPO learner's MLPs on the 5th-gen tensor cores (sm_100a), hand-written.
//
// brax's ppo.train (as driven by training/train.py:135-161,242-263) evaluates two small MLPs, policy 171-512-256-128-24 and
// value 215-512-256-128-1 (SiLU), forward and backward on minibatches of 5120 / 5376 rows, 4 x 32 times per training step,
// at `jax_default_matmul_precision=highest` (train.py:93-94): fp32-accurate products.
//
// Precision: every fp32 value is split x = hi + lo with hi = bf16(x), lo = bf16(x - hi); a k-slice issues three
// `tcgen05.mma.kind::f16` (lo hi + hi lo + hi hi, fp32 accumulation in TMEM): products carry ~16 mantissa bits.
//
// Data flow (what makes this fast: nobody converts inside a GEMM, and every operand tile is ONE bulk copy):
//   every matrix that is a GEMM operand lives in a BLOCKED SPLIT format: 128 x 128 blocks of 64 KB, each block four
//   k-chunks of 32, each chunk [hi 8 KB | lo 8 KB], each part a grid of 8 x 8 bf16 core matrices (128 B: eight 16-byte
//   rows) - 4 cores along the reduction index (stride 2 KB) x 16 cores along the other index (stride 128 B). The SAME
//   8 x 8 core is a K-major core (its rows are M/N indices) and an MN-major core (its rows are k indices), so a matrix is
//   stored in two variants that differ only in which of its indices picks the chunk:
//     variant C ("k along columns"): operand of a GEMM that reduces over the matrix's columns (K-major descriptor),
//     variant R ("k along rows"):    operand of a GEMM that reduces over the matrix's rows (MN-major descriptor),
//   and the 16 KB [hi | lo] chunk of a 128-wide tile is contiguous in global memory in both: a pipeline stage is two
//   `cp.async.bulk` (A chunk, B chunk) landing in shared memory exactly as `tcgen05.mma` reads them (no swizzle,
//   LBO = 2 KB along k, SBO = 128 B along M/N).
//   forward      z_l = in_l W_l + b_l:     A = in_l.C,  B = W_l.R;   epilogue writes z_l (fp32) and SiLU(z_l) as in_{l+1}.C / .R
//   input grad   dz_{l-1} = (dz_l W_l^T) * SiLU'(z_{l-1}):  A = dz_l.C, B = W_l.C;  epilogue writes dz_{l-1}.C / .R
//   weight grad  dW_l = in_l^T dz_l (rows split over CTAs): A = in_l.R, B = dz_l.R;  in_l.R carries a column of ones behind
//                its last feature, so the row behind dW_l's last row is db_l - the bias gradient comes out of the same GEMM;
//                split partials are summed in a fixed order (deterministic)
//   The fp32 inputs of a step (observations, the head's gradient, the weights) enter through `pgtt_bsplit_kernel`.
//
// One CTA = one 128 x TN output tile: warp 0 lane 0 = bulk-copy producer, warp 1 lane 0 = MMA issuer (+ TMEM owner),
// warps 2-17 = epilogue (TMEM lane quarter = warp % 4, column group of 32 = (warp - 2) / 4): the epilogue (bias, SiLU, the
// hi / lo split and the stores in both variants, ~1400 instructions per 32 columns of a row) is spread over 16 warps.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/pgtt_b200.h"

#define BS_BLOCK_BYTES 65536
#define BS_CHUNK_BYTES 16384
#define BS_PART_BYTES 8192
#define BG_STAGES 4
#define BG_STAGE_BYTES (2 * BS_CHUNK_BYTES)
#define BG_THREADS 576                  // producer warp + MMA warp + 16 epilogue warps (4 TMEM lane quarters x 4 column groups of 32)
#define BG_SMEM (BG_STAGES * BG_STAGE_BYTES + 256)
#define MLP_MAX_LAYERS 8

static thread_local std::string g_merr;
static int mfail(int code, const std::string& m) { g_merr = m; return code; }
#define MCUDA(call)                                                                                        \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) return mfail(PGTT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 @16 (next core matrix
// along k) | SBO >> 4 @32 (next core matrix along M/N) | version 1 @46. Same fields for K-major and MN-major operands
// (cute/atom/mma_traits_sm100.hpp: INTERLEAVE layouts ((8,m),(T,2)):((1T,SBO),(1,LBO)) and ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)))
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((2048u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
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

// eight fp32 -> [hi | lo] bf16 octets
__device__ __forceinline__ void split8(const float* v, uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// byte offset of the 16-byte octet (row r, columns 8 c8 .. 8 c8 + 7) of a blocked matrix with `ncb` block columns;
// the hi part - the lo part sits BS_PART_BYTES behind
__device__ __forceinline__ size_t bs_off_C(int r, int c8, int ncb) {   // k along columns
  const int rb = r >> 7, rr = r & 127, cb = c8 >> 4, cc = c8 & 15;
  return ((size_t)rb * ncb + cb) * BS_BLOCK_BYTES + (size_t)(cc >> 2) * BS_CHUNK_BYTES + (size_t)(cc & 3) * 2048 + (size_t)(rr >> 3) * 128 + (size_t)(rr & 7) * 16;
}
__device__ __forceinline__ size_t bs_off_R(int r, int c8, int ncb) {   // k along rows
  const int rb = r >> 7, rr = r & 127, cb = c8 >> 4, cc = c8 & 15;
  return ((size_t)rb * ncb + cb) * BS_BLOCK_BYTES + (size_t)(rr >> 5) * BS_CHUNK_BYTES + (size_t)((rr >> 3) & 3) * 2048 + (size_t)cc * 128 + (size_t)(rr & 7) * 16;
}

// pre-activations live in the handle only, so their layout is the epilogue's: groups of 32 rows x 4 columns (512 B), a warp's
// float4 access (lane = row) is one contiguous 512-byte piece instead of 32 pieces of 16 bytes 2 KB apart
__device__ __forceinline__ size_t z_off(int row, int col, int n4) { return ((size_t)(row >> 5) * n4 + (col >> 2)) * 128 + (size_t)(row & 31) * 4 + (col & 3); }

// ---- fp32 row-major -> blocked split ----------------------------------------------------------------------------------
struct SplitJob {
  const float* src; int rows, cols, ld;
  uint8_t* dstC; uint8_t* dstR; int ncb;
  int ones_col;      // variant R gets 1.0 in this column for every row (the bias-gradient trick), -1: none
  // optional fused minibatch gather + observation normalisation (the learner's input): row r = (t, j) = (r / gather_mb, r % gather_mb)
  // reads source row t * gather_S + gather[j]; value = (x - mean[c]) * inv_std[c]
  const long long* gather; int gather_mb, gather_S;
  const float* mean; const float* inv_std;
};
struct SplitJobs { SplitJob j[MLP_MAX_LAYERS]; };

__global__ void pgtt_bsplit_kernel(SplitJobs J) {
  const SplitJob& s = J.j[blockIdx.y];
  const int width = s.ones_col >= s.cols ? s.ones_col + 1 : s.cols;
  const int noct = (width + 7) >> 3;
  const long long total = (long long)s.rows * noct;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / noct), c8 = (int)(i % noct);
    float v[8];
    const size_t sr = s.gather ? (size_t)(r / s.gather_mb) * s.gather_S + (size_t)__ldg(s.gather + r % s.gather_mb) : (size_t)r;
    const float* p = s.src + sr * s.ld + c8 * 8;
    if (c8 * 8 + 8 <= s.cols && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = (c8 * 8 + k < s.cols) ? __ldg(p + k) : 0.f;
    }
    if (s.mean) {
#pragma unroll
      for (int k = 0; k < 8; k++) if (c8 * 8 + k < s.cols) v[k] = (v[k] - __ldg(s.mean + c8 * 8 + k)) * __ldg(s.inv_std + c8 * 8 + k);
    }
    uint4 hi, lo;
    if (s.dstC) {
      split8(v, &hi, &lo);
      uint8_t* d = s.dstC + bs_off_C(r, c8, s.ncb);
      *reinterpret_cast<uint4*>(d) = hi; *reinterpret_cast<uint4*>(d + BS_PART_BYTES) = lo;
    }
    if (s.dstR) {
      if (s.ones_col >= c8 * 8 && s.ones_col < c8 * 8 + 8) v[s.ones_col - c8 * 8] = 1.f;
      split8(v, &hi, &lo);
      uint8_t* d = s.dstR + bs_off_R(r, c8, s.ncb);
      *reinterpret_cast<uint4*>(d) = hi; *reinterpret_cast<uint4*>(d + BS_PART_BYTES) = lo;
    }
  }
}
// the column of ones of an activation matrix (variant R): written once, the epilogues never touch it... unless it shares an
// octet with valid columns, which the epilogue then preserves
__global__ void pgtt_bones_kernel(uint8_t* dstR, int rows, int col, int ncb) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  uint8_t* d = dstR + bs_off_R(r, col >> 3, ncb) + (col & 7) * 2;
  *reinterpret_cast<uint16_t*>(d) = 0x3F80;                       // bf16(1.0); the lo part stays 0
}

// ---- the GEMM ---------------------------------------------------------------------------------------------------------
struct BOperand { const uint8_t* base; long long mn_stride, kb_stride; };   // tile (block b, chunk j) at base + b mn_stride + (j / 4) kb_stride + (j % 4) 16 KB
enum { EP_OUT = 0, EP_HIDDEN = 1, EP_DX = 2, EP_PART = 3 };
struct BGemm {
  BOperand A, B;
  uint32_t idesc;          // without N: D fp32, A / B bf16, the two major bits, M = 128
  int TN, nsub;            // output tile width (32 / 64 / 128), sub-tiles per 128-block of B
  int chunks_total, chunks_per_split;
  int mode;
  int Mc, Nc;              // valid output rows / columns
  float* out; int ldo;     // EP_OUT: [Mc][ldo] (+ bias); EP_PART: [split][Mc][ldo]
  const float* bias;       // EP_OUT / EP_HIDDEN
  float* z; int ldz;       // EP_HIDDEN: pre-activations written; EP_DX: read (z_off layout, ldz = float4 groups per row)
  uint8_t* outC; uint8_t* outR; int out_ncb, ones_col;   // EP_HIDDEN / EP_DX: blocked split outputs
  unsigned long long* trace;   // development aid: CTA (0, 0, 0) stores 6 %globaltimer stamps here (may be null)
};
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define BG_TRACE(i) do { if (g.trace && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) g.trace[i] = gtime(); } while (0)

__global__ void __launch_bounds__(BG_THREADS, 1) pgtt_bgemm_kernel(BGemm g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BG_STAGES * BG_STAGE_BYTES);
  uint64_t* empty = full + BG_STAGES;
  uint64_t* done = empty + BG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mb = blockIdx.x, nb = blockIdx.y / g.nsub, sub = blockIdx.y % g.nsub, split = blockIdx.z;
  const int c_begin = split * g.chunks_per_split;
  const int nchunks = min(g.chunks_total, c_begin + g.chunks_per_split) - c_begin;
  const uint32_t tmem_cols = g.TN < 32 ? 32u : (uint32_t)g.TN;

  if (tid == 0) {
    BG_TRACE(0);                                            // kernel entered
    for (int i = 0; i < BG_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {                                        // producer
      const uint8_t* a0 = g.A.base + (long long)mb * g.A.mn_stride;
      const uint8_t* b0 = g.B.base + (long long)nb * g.B.mn_stride;
      for (int c = 0; c < nchunks; c++) {
        const int s = c % BG_STAGES, j = c_begin + c;
        if (c >= BG_STAGES) mbar_wait(&empty[s], (uint32_t)((c / BG_STAGES - 1) & 1));
        const uint32_t st = smem_u32(smem + (size_t)s * BG_STAGE_BYTES);
        mbar_expect_tx(&full[s], BG_STAGE_BYTES);
        bulk_g2s(st, a0 + (long long)(j >> 2) * g.A.kb_stride + (long long)(j & 3) * BS_CHUNK_BYTES, BS_CHUNK_BYTES, &full[s]);
        bulk_g2s(st + BS_CHUNK_BYTES, b0 + (long long)(j >> 2) * g.B.kb_stride + (long long)(j & 3) * BS_CHUNK_BYTES, BS_CHUNK_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                        // MMA issuer
      const uint32_t idesc = g.idesc | ((uint32_t)(g.TN >> 3) << 17);
      for (int c = 0; c < nchunks; c++) {
        const int s = c % BG_STAGES;
        mbar_wait(&full[s], (uint32_t)((c / BG_STAGES) & 1));
        if (c == 0) BG_TRACE(1);                            // first operand chunk landed
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + (size_t)s * BG_STAGE_BYTES), a_lo = a_hi + BS_PART_BYTES;
        const uint32_t b_hi = a_hi + BS_CHUNK_BYTES + (uint32_t)(sub * g.TN) * 16u, b_lo = b_hi + BS_PART_BYTES;
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {                    // 16 k = two core matrices along k
          const uint32_t o = (uint32_t)ks * 4096u;
          const uint64_t dah = make_desc(a_hi + o), dal = make_desc(a_lo + o), dbh = make_desc(b_hi + o), dbl = make_desc(b_lo + o);
          umma_bf16(tmem, dal, dbh, idesc, (uint32_t)(c > 0 || ks > 0));   // small terms first
          umma_bf16(tmem, dah, dbl, idesc, 1u);
          umma_bf16(tmem, dah, dbh, idesc, 1u);
        }
        umma_commit(&empty[s]);
        if (c == nchunks - 1) { umma_commit(done); BG_TRACE(2); }   // every MMA issued
      }
    }
  } else {                                                  // epilogue: TMEM lane quarter = warp % 4, columns 32 ((warp - 2) / 4) ..
    const int q = warp & 3;
    const int row = mb * 128 + q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int cb = ((warp - 2) >> 2) * 32;
    const int col0 = nb * 128 + sub * g.TN + cb;
    const bool row_ok = row < g.Mc;
    if (cb < g.TN && col0 < g.Nc) {                         // (warp-uniform)
      mbar_wait(done, 0);
      if (tid == 64) BG_TRACE(3);                           // accumulators complete
      tc_fence_after();
      float v[32];
      tmem_ld32(trow + (uint32_t)cb, v);
      if (row_ok && (g.mode == EP_OUT || g.mode == EP_PART)) {
        float* o = g.out + ((size_t)(g.mode == EP_PART ? split : 0) * g.Mc + row) * g.ldo + col0;
        const bool vec = ((reinterpret_cast<uintptr_t>(o) & 15) == 0) && col0 + 32 <= g.Nc;
        if (g.bias) {
          if (col0 + 32 <= g.Nc && (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) { const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + i)); v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w; }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++) if (col0 + i < g.Nc) v[i] += __ldg(g.bias + col0 + i);
          }
        }
        if (vec) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) if (col0 + i < g.Nc) o[i] = v[i];
        }
      } else if (row_ok) {
        float* zr = g.z + z_off(row, col0, g.ldz);          // (ldz = float4 groups per row)
        const bool full = col0 + 32 <= g.Nc;                // (warp-uniform) every column of the group is valid: no per-element predicates
        if (g.mode == EP_HIDDEN) {                            // z = acc + b (kept), y = SiLU(z)
          if (full && (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + i));     // (col0 is a multiple of 32)
              v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
              *reinterpret_cast<float4*>(zr + i * 32) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = (col0 + i < g.Nc) ? v[i] + __ldg(g.bias + col0 + i) : 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) if (col0 + i < g.Nc) *reinterpret_cast<float4*>(zr + i * 32) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = __fdividef(v[i], 1.f + __expf(-v[i]));
        } else {                                              // EP_DX: dz = acc * SiLU'(z)
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (full || col0 + i < g.Nc) t = *reinterpret_cast<const float4*>(zr + i * 32);
            const float zz[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const float sg = __fdividef(1.f, 1.f + __expf(-zz[k]));
              const float d = v[i + k] * (sg * (1.f + zz[k] * (1.f - sg)));
              v[i + k] = (full || col0 + i + k < g.Nc) ? d : 0.f;
            }
          }
        }
        // blocked split outputs, four octets in both variants. col0 is a multiple of 32, so the four octets are consecutive k-cores (variant C: 2 KB apart)
        // respectively consecutive M/N-cores (variant R: 128 B apart) of one chunk: two base addresses, constant offsets
        const int c80 = col0 >> 3;
        uint8_t* pc = g.outC + bs_off_C(row, c80, g.out_ncb);
        uint8_t* pr = g.outR + bs_off_R(row, c80, g.out_ncb);
        const int ones_o8 = (g.ones_col >= col0 && g.ones_col < col0 + 32) ? ((g.ones_col - col0) >> 3) : -1;   // the octet that holds the column of ones, if any
        const int n_oct = min(4, g.out_ncb * 16 - c80);
#pragma unroll
        for (int o8 = 0; o8 < 4; o8++) {
          if (o8 < n_oct) {
            uint4 hi, lo;
            split8(v + 8 * o8, &hi, &lo);
            *reinterpret_cast<uint4*>(pc + o8 * 2048) = hi; *reinterpret_cast<uint4*>(pc + o8 * 2048 + BS_PART_BYTES) = lo;
            if (o8 == ones_o8) {
              float t[8];
#pragma unroll
              for (int i = 0; i < 8; i++) t[i] = (col0 + 8 * o8 + i == g.ones_col) ? 1.f : v[8 * o8 + i];
              split8(t, &hi, &lo);
            }
            *reinterpret_cast<uint4*>(pr + o8 * 128) = hi; *reinterpret_cast<uint4*>(pr + o8 * 128 + BS_PART_BYTES) = lo;
          }
        }
      }
    }
    tc_fence_before();
  }
  if (tid == 64) BG_TRACE(4);                               // first epilogue warp done
  __syncthreads();
  if (tid == 0) BG_TRACE(5);                                // CTA done
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

// dW = sum of the split partials in split order (deterministic); the row behind dW's last row is db
__global__ void pgtt_bsum_kernel(const float* __restrict__ part, int splits, int rows_p, int K, int N, float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (K + 1) * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; z++) s += part[(size_t)z * rows_p * N + i];
  if (i < K * N) dw[i] = s; else if (db) db[i - K * N] = s;
}

}  // namespace

// ---- host side ----------------------------------------------------------------------------------------------------------
struct pgtt_mlp {
  int L, dims[MLP_MAX_LAYERS + 1], rows, device;
  int nrb;                                   // row blocks
  // blocked operands (device): in[l] = input of layer l (x for l = 0, SiLU(z_{l-1}) otherwise); dz[l] = gradient wrt z_l
  uint8_t *inC[MLP_MAX_LAYERS], *inR[MLP_MAX_LAYERS]; int in_ncb[MLP_MAX_LAYERS];
  uint8_t *wC[MLP_MAX_LAYERS], *wR[MLP_MAX_LAYERS]; int w_ncb[MLP_MAX_LAYERS];
  uint8_t *dzC[MLP_MAX_LAYERS], *dzR[MLP_MAX_LAYERS]; int dz_ncb[MLP_MAX_LAYERS];
  float* z[MLP_MAX_LAYERS];                  // pre-activations of the hidden layers (z_off layout)
  float* part; int splits, chunks_per_split;  // split partials of the weight-gradient GEMMs
  std::vector<void*> allocs;
  bool forward_done;
  unsigned long long* trace; int trace_on, n_launch;   // development aid (PGTT_MLP_TRACE=1): stamps of the GEMM launches of the last forward + backward
  cudaStream_t aux; cudaEvent_t ev[MLP_MAX_LAYERS + 2];  // weight-gradient GEMMs run beside the input-gradient chain
};

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

static int launch_bgemm(const BGemm& g, int m_blocks, int n_tiles, int splits, cudaStream_t st) {
  pgtt_bgemm_kernel<<<dim3(m_blocks, n_tiles, splits), BG_THREADS, BG_SMEM, st>>>(g);
  MCUDA(cudaGetLastError());
  return PGTT_OK;
}
// the widest tile the matrix fills: the main loops are bound by the L2 -> shared-memory traffic of the operand chunks (a 128-wide
// A chunk is re-read once per tile column), so fewer, wider tiles beat more CTAs (measured, profiles/r02f: forward + backward of
// the policy network 126 us with 128-wide tiles against 161 us with tiles narrowed to put a CTA on every SM)
static inline int pick_tn(int n) { return n > 64 ? 128 : (n > 32 ? 64 : 32); }
static inline uint32_t idesc_base(int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(128 >> 4) << 24);
}
static inline BOperand op_C(const uint8_t* base, int ncb) { return BOperand{base, (long long)ncb * BS_BLOCK_BYTES, (long long)BS_BLOCK_BYTES}; }   // blocks along rows, k along columns
static inline BOperand op_R(const uint8_t* base, int ncb) { return BOperand{base, (long long)BS_BLOCK_BYTES, (long long)ncb * BS_BLOCK_BYTES}; }   // blocks along columns, k along rows

extern "C" {

const char* pgtt_mlp_last_error(void) { return g_merr.c_str(); }
void pgtt_mlp_destroy(pgtt_mlp* m);

int pgtt_mlp_create(int n_layers, const int* dims, int rows, int device, pgtt_mlp** out) {
  if (!dims || !out || n_layers < 1 || n_layers >= MLP_MAX_LAYERS || rows < 1) return mfail(PGTT_ERR_ARG, "pgtt_mlp_create: bad argument");
  for (int i = 0; i <= n_layers; i++) if (dims[i] < 1 || dims[i] > 4096) return mfail(PGTT_ERR_ARG, "pgtt_mlp_create: layer widths must be in 1..4096");
  MCUDA(cudaSetDevice(device));
  MCUDA(cudaFuncSetAttribute(pgtt_bgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BG_SMEM));   // (per device: every handle asks for it)
  pgtt_mlp* m = new pgtt_mlp();
  m->L = n_layers; m->rows = rows; m->device = device; m->nrb = cdiv(rows, 128); m->forward_done = false;
  m->trace = nullptr; m->trace_on = 0; m->n_launch = 0; m->aux = nullptr;
  for (int i = 0; i < MLP_MAX_LAYERS + 2; i++) m->ev[i] = nullptr;
  if (const char* t = getenv("PGTT_MLP_TRACE")) m->trace_on = atoi(t) != 0;
  for (int i = 0; i <= n_layers; i++) m->dims[i] = dims[i];
  auto alloc = [&](size_t bytes, void** p) -> int {
    MCUDA(cudaMalloc(p, bytes));
    m->allocs.push_back(*p);
    MCUDA(cudaMemset(*p, 0, bytes));
    return PGTT_OK;
  };
  int rc = PGTT_OK;
  size_t part_floats = 0;
  for (int l = 0; l < n_layers && rc == PGTT_OK; l++) {
    const int K = dims[l], N = dims[l + 1];
    m->in_ncb[l] = cdiv(K + 1, 128);                         // + the column of ones
    m->w_ncb[l] = cdiv(N, 128);
    m->dz_ncb[l] = cdiv(N, 128);
    const size_t in_b = (size_t)m->nrb * m->in_ncb[l] * BS_BLOCK_BYTES, w_b = (size_t)cdiv(K, 128) * m->w_ncb[l] * BS_BLOCK_BYTES;
    const size_t dz_b = (size_t)m->nrb * m->dz_ncb[l] * BS_BLOCK_BYTES;
    if ((rc = alloc(in_b, (void**)&m->inC[l]))) break;
    if ((rc = alloc(in_b, (void**)&m->inR[l]))) break;
    if ((rc = alloc(w_b, (void**)&m->wC[l]))) break;
    if ((rc = alloc(w_b, (void**)&m->wR[l]))) break;
    if ((rc = alloc(dz_b, (void**)&m->dzC[l]))) break;
    if ((rc = alloc(dz_b, (void**)&m->dzR[l]))) break;
    m->z[l] = nullptr;
    if (l + 1 < n_layers && (rc = alloc((size_t)cdiv(rows, 32) * cdiv(N, 4) * 128 * sizeof(float), (void**)&m->z[l]))) break;
    part_floats += (size_t)(K + 1) * N;                      // every layer has its own partials: the weight-gradient GEMMs overlap
  }
  if (rc == PGTT_OK) {
    const int chunks = cdiv(rows, 32);
    m->splits = chunks >= 160 ? 10 : (chunks >= 16 ? cdiv(chunks, 16) : 1);
    m->chunks_per_split = cdiv(chunks, m->splits);
    m->splits = cdiv(chunks, m->chunks_per_split);
    rc = alloc((size_t)m->splits * part_floats * sizeof(float), (void**)&m->part);
  }
  if (rc == PGTT_OK) rc = alloc(sizeof(unsigned long long) * 8 * 4 * MLP_MAX_LAYERS, (void**)&m->trace);
  if (rc == PGTT_OK) {
    cudaError_t e = cudaStreamCreateWithFlags(&m->aux, cudaStreamNonBlocking);
    for (int i = 0; i < MLP_MAX_LAYERS + 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&m->ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) rc = mfail(PGTT_ERR_CUDA, std::string("pgtt_mlp_create: ") + cudaGetErrorString(e));
  }
  if (rc == PGTT_OK) {
    // columns of ones of the hidden activations (layer 0's comes with every split of the input)
    for (int l = 1; l < n_layers; l++) pgtt_bones_kernel<<<cdiv(rows, 256), 256>>>(m->inR[l], rows, dims[l], m->in_ncb[l]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = mfail(PGTT_ERR_CUDA, std::string("pgtt_mlp_create: ") + cudaGetErrorString(e));
  }
  if (rc != PGTT_OK) {
    pgtt_mlp_destroy(m);
    return rc;
  }
  *out = m;
  return PGTT_OK;
}

void pgtt_mlp_destroy(pgtt_mlp* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  for (void* p : m->allocs) cudaFree(p);
  for (int i = 0; i < MLP_MAX_LAYERS + 2; i++) if (m->ev[i]) cudaEventDestroy(m->ev[i]);
  if (m->aux) cudaStreamDestroy(m->aux);
  delete m;
}

int pgtt_mlp_rows(const pgtt_mlp* m) { return m ? m->rows : 0; }

/* development aid (PGTT_MLP_TRACE=1 at create): 6 %globaltimer stamps (ns) of CTA (0, 0, 0) of each GEMM launch of the last
 * forward + backward, in launch order: entered, first chunk landed, MMAs issued, accumulators complete, epilogue warp done, CTA done.
 * Returns the number of launches recorded; synchronises the device. */
int pgtt_mlp_debug_trace(pgtt_mlp* m, unsigned long long* out, int max_launches) {
  if (!m || !out || !m->trace_on) return 0;
  cudaDeviceSynchronize();
  const int n = m->n_launch < max_launches ? m->n_launch : max_launches;
  cudaMemcpy(out, m->trace, sizeof(unsigned long long) * 8 * n, cudaMemcpyDeviceToHost);
  return n;
}

static unsigned long long* next_trace(pgtt_mlp* m) {
  if (!m->trace_on || m->n_launch >= 4 * MLP_MAX_LAYERS) return nullptr;
  return m->trace + 8 * (m->n_launch++);
}

/* y [rows][dims[L]] = MLP(x [rows][:dims[0]] (row stride ldx)); w[l] [dims[l]][dims[l + 1]] (the flax `kernel`), b[l] [dims[l + 1]];
 * SiLU between layers, none after the last. Keeps what pgtt_mlp_backward needs inside the handle. */
static int mlp_forward(pgtt_mlp* m, const SplitJob& input, const float* const* w, const float* const* b, float* y, cudaStream_t st) {
  m->n_launch = 0;
  SplitJobs J = {};                                         // the input and every weight matrix: one launch
  J.j[0] = input;
  for (int l = 0; l < m->L; l++) {
    if (!w[l] || !b[l]) return mfail(PGTT_ERR_ARG, "pgtt_mlp_forward: null parameter");
    SplitJob& jw = J.j[l + 1];
    jw.src = w[l]; jw.rows = m->dims[l]; jw.cols = m->dims[l + 1]; jw.ld = m->dims[l + 1]; jw.dstC = m->wC[l]; jw.dstR = m->wR[l]; jw.ncb = m->w_ncb[l]; jw.ones_col = -1;
  }
  pgtt_bsplit_kernel<<<dim3(592, m->L + 1), 256, 0, st>>>(J);   // (the input job is ~150 k octets: one grid-stride pass)
  MCUDA(cudaGetLastError());
  for (int l = 0; l < m->L; l++) {
    const int K = m->dims[l], N = m->dims[l + 1];
    const bool last = l + 1 == m->L;
    BGemm g = {};
    g.A = op_C(m->inC[l], m->in_ncb[l]);
    g.B = op_R(m->wR[l], m->w_ncb[l]);
    g.idesc = idesc_base(0, 1);
    g.TN = pick_tn(N); g.nsub = 128 / g.TN;
    g.chunks_total = cdiv(K, 32); g.chunks_per_split = g.chunks_total;
    g.Mc = m->rows; g.Nc = N; g.bias = b[l];
    g.trace = next_trace(m);
    if (last) { g.mode = EP_OUT; g.out = y; g.ldo = N; }
    else { g.mode = EP_HIDDEN; g.z = m->z[l]; g.ldz = cdiv(N, 4); g.outC = m->inC[l + 1]; g.outR = m->inR[l + 1]; g.out_ncb = m->in_ncb[l + 1]; g.ones_col = N; }
    if (int rc = launch_bgemm(g, m->nrb, cdiv(N, g.TN), 1, st)) return rc;
  }
  m->forward_done = true;
  return PGTT_OK;
}

int pgtt_mlp_forward(pgtt_mlp* m, const float* x, int ldx, const float* const* w, const float* const* b, float* y, void* stream) {
  if (!m || !x || !w || !b || !y || ldx < m->dims[0]) return mfail(PGTT_ERR_ARG, "pgtt_mlp_forward: bad argument");
  SplitJob in = {};
  in.src = x; in.rows = m->rows; in.cols = m->dims[0]; in.ld = ldx; in.dstC = m->inC[0]; in.dstR = m->inR[0]; in.ncb = m->in_ncb[0]; in.ones_col = m->dims[0];
  return mlp_forward(m, in, w, b, y, (cudaStream_t)stream);
}

/* The same with the learner's input fused in: row (t, j) of the minibatch (t < rows / mb, j < mb) is row t * S + idx[j] of the
 * time-major transition store `data` [T][S][ld] (idx: DEVICE int64 [mb] segment ids of this minibatch), normalised on the way in
 * as (x - mean[c]) * inv_std[c] when mean / inv_std (DEVICE fp32 [dims[0]]) are given - brax `running_statistics.normalize`. */
int pgtt_mlp_forward_gather(pgtt_mlp* m, const float* data, int ld, int S, const long long* idx, int mb, const float* mean, const float* inv_std,
                            const float* const* w, const float* const* b, float* y, void* stream) {
  if (!m || !data || !idx || !w || !b || !y || ld < m->dims[0] || mb < 1 || S < 1 || m->rows % mb != 0 || (!mean) != (!inv_std))
    return mfail(PGTT_ERR_ARG, "pgtt_mlp_forward_gather: bad argument (rows must be a multiple of mb)");
  SplitJob in = {};
  in.src = data; in.rows = m->rows; in.cols = m->dims[0]; in.ld = ld; in.dstC = m->inC[0]; in.dstR = m->inR[0]; in.ncb = m->in_ncb[0]; in.ones_col = m->dims[0];
  in.gather = idx; in.gather_mb = mb; in.gather_S = S; in.mean = mean; in.inv_std = inv_std;
  return mlp_forward(m, in, w, b, y, (cudaStream_t)stream);
}

/* Gradients of the last pgtt_mlp_forward: dy [rows][dims[L]] = dLoss/dy; dw[l] [dims[l]][dims[l + 1]], db[l] [dims[l + 1]] are
 * overwritten. The parameters must not have changed since the forward (their blocked copies are re-used). The input-gradient
 * chain runs on `stream`; the weight-gradient GEMMs run beside it on the handle's own stream (forked and joined with events,
 * so the call is stream-ordered on `stream` as a whole and can be captured into a CUDA graph). */
int pgtt_mlp_backward(pgtt_mlp* m, const float* dy, float* const* dw, float* const* db, void* stream) {
  if (!m || !dy || !dw || !db) return mfail(PGTT_ERR_ARG, "pgtt_mlp_backward: bad argument");
  if (!m->forward_done) return mfail(PGTT_ERR_ARG, "pgtt_mlp_backward: no forward pass to differentiate");
  cudaStream_t st = (cudaStream_t)stream, aux = m->aux;
  const int L = m->L;
  for (int l = 0; l < L; l++) if (!dw[l] || !db[l]) return mfail(PGTT_ERR_ARG, "pgtt_mlp_backward: null gradient");
  SplitJobs J = {};
  { SplitJob& jd = J.j[0]; jd.src = dy; jd.rows = m->rows; jd.cols = m->dims[L]; jd.ld = m->dims[L]; jd.dstC = m->dzC[L - 1]; jd.dstR = m->dzR[L - 1]; jd.ncb = m->dz_ncb[L - 1]; jd.ones_col = -1; }
  pgtt_bsplit_kernel<<<dim3(296, 1), 256, 0, st>>>(J);
  MCUDA(cudaGetLastError());
  size_t part_off = 0;
  for (int l = L - 1; l >= 0; l--) {
    const int K = m->dims[l], N = m->dims[l + 1];
    MCUDA(cudaEventRecord(m->ev[l], st));                   // dz_l is complete
    MCUDA(cudaStreamWaitEvent(aux, m->ev[l], 0));
    {                                                       // [dW_l; db_l] = [in_l | 1]^T dz_l, beside the chain
      float* part = m->part + part_off;
      part_off += (size_t)m->splits * (K + 1) * N;
      BGemm g = {};
      g.A = op_R(m->inR[l], m->in_ncb[l]);
      g.B = op_R(m->dzR[l], m->dz_ncb[l]);
      g.idesc = idesc_base(1, 1);
      g.TN = pick_tn(N); g.nsub = 128 / g.TN;
      g.chunks_total = cdiv(m->rows, 32); g.chunks_per_split = m->chunks_per_split;
      g.mode = EP_PART; g.Mc = K + 1; g.Nc = N; g.out = part; g.ldo = N;
      g.trace = next_trace(m);
      if (int rc = launch_bgemm(g, cdiv(K + 1, 128), cdiv(N, g.TN), m->splits, aux)) return rc;
      pgtt_bsum_kernel<<<cdiv((K + 1) * N, 256), 256, 0, aux>>>(part, m->splits, K + 1, K, N, dw[l], db[l]);
      MCUDA(cudaGetLastError());
    }
    if (l > 0) {                                            // dz_{l-1} = (dz_l W_l^T) * SiLU'(z_{l-1})
      BGemm g = {};
      g.A = op_C(m->dzC[l], m->dz_ncb[l]);
      g.B = op_C(m->wC[l], m->w_ncb[l]);
      g.idesc = idesc_base(0, 0);
      g.TN = pick_tn(K); g.nsub = 128 / g.TN;
      g.chunks_total = cdiv(N, 32); g.chunks_per_split = g.chunks_total;
      g.mode = EP_DX; g.Mc = m->rows; g.Nc = K;
      g.z = m->z[l - 1]; g.ldz = cdiv(K, 4); g.outC = m->dzC[l - 1]; g.outR = m->dzR[l - 1]; g.out_ncb = m->dz_ncb[l - 1]; g.ones_col = -1;
      g.trace = next_trace(m);
      if (int rc = launch_bgemm(g, m->nrb, cdiv(K, g.TN), 1, st)) return rc;
    }
  }
  MCUDA(cudaEventRecord(m->ev[L], aux));                    // join
  MCUDA(cudaStreamWaitEvent(st, m->ev[L], 0));
  return PGTT_OK;
}

}  // extern "C"
