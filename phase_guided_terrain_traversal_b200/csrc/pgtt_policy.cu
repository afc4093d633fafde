// pgtt_policy.cu - acting step of the PPO rollout collector on the 5th-gen tensor cores (sm_100a).
//
// Stands in for brax's `actor_step` as configured by training/train.py:135-161 (policy MLP
// (512,256,128) on obs["state"], swish, NormalTanh distribution) and for deploy/policy_net.py:35-64
// (same network re-hosted in torch): normalise obs -> 171->512->256->128->24 MLP -> loc | scale ->
// raw = loc + (softplus(scale) + 0.001) * eps -> action = tanh(raw), log_prob.
//
// One CTA (256 threads) owns a tile of 128 envs and runs the WHOLE network for it in one launch:
//   * activations live in shared memory as the K-major A operand (bf16, canonical no-swizzle UMMA layout),
//   * weights are pre-packed on the host into the same canonical layout (B operand, K-chunks of <= 32 KB) and
//     stream through a double buffer with TMA bulk copies (cp.async.bulk + mbarrier complete_tx) that run one
//     chunk ahead of the MMAs,
//   * `tcgen05.mma.cta_group::1.kind::f16` (M = 128, N <= 256, K = 16 per instruction) issued by one
//     thread accumulates a whole layer in TMEM (fp32, up to 512 columns),
//   * after `tcgen05.commit` -> mbarrier, the eight warps read their TMEM lane quarter / column half with `tcgen05.ld` x32,
//     apply bias + SiLU and write the next layer's A operand in place; the last layer's epilogue does
//     the distribution math and writes action / raw_action / log_prob.
// bf16 operands, fp32 accumulation (north-star: "policy MLP on tensor cores, bf16 in, fp32 accumulate").
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/pgtt_b200.h"

// pgtt_api.cu (not part of the ABI header)
extern "C" void pgtt_internal_count_launches(pgtt_env* env, int64_t n);
extern "C" int pgtt_step_launches(pgtt_env* env);
extern "C" int pgtt_internal_make_resident(pgtt_env* env, void* stream);
extern "C" int pgtt_internal_mark_launched(pgtt_env* env, void* stream);
extern "C" uint64_t pgtt_internal_serial(pgtt_env* env);   // unique per pgtt_create: a handle re-created at the same address is a different env

#define POL_TM 128            // envs per CTA = MMA M
#define POL_MAXK 512          // widest activation
#define POL_MAXLAYERS 6
#define POL_BCHUNK 16384      // bytes per weight chunk buffer
#define POL_NBUF 5            // chunk buffers: up to 4 TMA bulk copies in flight ahead of the MMAs
#define POL_MAXBIAS 1024      // padded biases of all layers, staged in shared memory
#define POL_A_BYTES (POL_TM * POL_MAXK * 2)
#define POL_SMEM (POL_A_BYTES + POL_NBUF * POL_BCHUNK + POL_MAXBIAS * 4 + 128)   // + 2 NBUF + 1 mbarriers and the TMEM slot

struct PolicyLayer {
  int K, N, Kp, Np, Kc, nchunks;   // true sizes, padded sizes, K per chunk
  size_t w_off, b_off;             // offsets (elements) into the packed weight / bias arrays
  // cluster layout (4 CTAs per 128-row tile, each owning Ns = Np / 4 output columns; the head layer stays whole on rank 0)
  int Ns, cKc, cnchunks;
  size_t cw_off, cw_rstride;       // element offset of rank 0's chunks in `cw`, elements per rank
};

struct PolicyParams {
  int n_layers, obs_dim, act_dim;
  PolicyLayer L[POL_MAXLAYERS];
  const __nv_bfloat16* w;   // packed chunks, canonical UMMA K-major layout
  const __nv_bfloat16* cw;  // the same weights packed per cluster rank (NULL: cluster kernel unavailable for these sizes)
  const float* bias;        // padded biases
  const float* mean;        // [obs_dim]
  const float* inv_std;     // [obs_dim]
};

// ----------------------------------------------------------------------------------------------
// PTX helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  // bounded spin: a protocol bug traps instead of hanging the GPU
  for (long it = 0; it < (1L << 28); it++) {
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

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address [0,14) >> 4 | leading byte offset [16,30) >> 4 (next core matrix along K) |
// stride byte offset [32,46) >> 4 (next 8-row group along M/N) | version [46,48) = 1 | layout type [61,64) = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @4, a/b_format BF16 = 1 @7/@10,
// a/b major K = 0 @15/@16, N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}

// counter-based standard normal: threefry2x32 (same block function as the env's jax.random port) + Box-Muller
__device__ __forceinline__ uint32_t rotl32p(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__device__ void threefry_p(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int R[8] = {13, 15, 26, 6, 17, 29, 16, 24};
  x0 += ks[0]; x1 += ks[1];
#pragma unroll 1
  for (int g = 0; g < 5; g++) {
    for (int i = 0; i < 4; i++) { x0 += x1; x1 = rotl32p(x1, R[(g & 1) * 4 + i]); x1 ^= x0; }
    x0 += ks[(g + 1) % 3]; x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
}
__device__ float normal_draw(uint64_t seed, uint64_t step, uint32_t row, uint32_t j) {
  uint32_t x0 = row, x1 = j ^ ((uint32_t)step << 8);
  threefry_p((uint32_t)seed ^ (uint32_t)(step >> 24), (uint32_t)(seed >> 32) + 0x9E3779B9u, x0, x1);
  const float u1 = ((x0 >> 8) + 1u) * (1.0f / 16777216.0f);   // (0, 1]
  const float u2 = (x1 >> 8) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ----------------------------------------------------------------------------------------------
// kernel
// ----------------------------------------------------------------------------------------------
#define POL_EPI_THREADS 512   // 16 epilogue warps: warp w reads TMEM lane quarter w & 3 (rows 32 (w & 3) ..), column quarter w >> 2
#define POL_THREADS 544       // + 1 control warp whose lane 0 is the TMA producer and the MMA issuer. It must not share a warp
                              // with threads that spin on an mbarrier: a diverged warp runs one path at a time and
                              // mbarrier.try_wait suspends, which stalled the issuer ~3 us per chunk (profiles/r01c)
#define POL_CTRL_TID 512

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit; completion is signalled on `bar` (complete_tx)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// issue only; the destination registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
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

// Pipeline: thread 0 is both the TMA producer (weight chunk g + 1 is in flight while the MMAs of chunk g run; buffers
// are recycled on the tcgen05.commit of the MMAs that read them) and the MMA issuer. full[b] / empty[b]: chunk g uses
// buffer b = g & 1 and completion number g >> 1 of both barriers.
__global__ void __launch_bounds__(POL_THREADS, 1)
pgtt_policy_kernel(PolicyParams P, const float* __restrict__ obs, int n_rows, unsigned long long seed, unsigned long long step_in,
                   const unsigned long long* __restrict__ step_base, int deterministic, const float* __restrict__ eps_in, float* __restrict__ action, float* __restrict__ raw_action,
                   float* __restrict__ log_prob, float* __restrict__ logits_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const unsigned long long step = step_in + (step_base ? *step_base : 0ull);   // noise counter; the base lives on the device for graph replays
  uint8_t* sA = smem;
  uint8_t* sB = smem + POL_A_BYTES;                                                  // [POL_NBUF][POL_BCHUNK]
  float* sBias = reinterpret_cast<float*>(smem + POL_A_BYTES + POL_NBUF * POL_BCHUNK);
  uint64_t* full = reinterpret_cast<uint64_t*>(sBias + POL_MAXBIAS);                 // [POL_NBUF]
  uint64_t* empty = full + POL_NBUF;                                                 // [POL_NBUF]
  uint64_t* layer_done = empty + POL_NBUF;                                           // completion number l = layer l accumulated
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(layer_done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * POL_TM;

  if (tid == 0) {
    for (int i = 0; i < POL_NBUF; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(layer_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();   // barriers initialised before the first bulk copy is issued
  // producer state (thread 0): the next chunk to load, as (layer, chunk in layer, global index)
  int ld_l = 0, ld_c = 0, ld_g = 0;
  auto top_up = [&](int upto) {   // keep the weight stream `POL_NBUF - 1` chunks ahead of the chunk being multiplied
    while (ld_l < P.n_layers && ld_g < upto) {
      const PolicyLayer& Ln = P.L[ld_l];
      const int b = ld_g % POL_NBUF;
      if (ld_g >= POL_NBUF) mbar_wait(&empty[b], (uint32_t)((ld_g / POL_NBUF - 1) & 1));   // the MMAs that read this buffer are done
      const uint32_t bytes = (uint32_t)(Ln.Np * Ln.Kc * 2);
      mbar_expect_tx(&full[b], bytes);
      bulk_g2s(sB + (size_t)b * POL_BCHUNK, reinterpret_cast<const uint8_t*>(P.w + Ln.w_off) + (size_t)ld_c * bytes, bytes, &full[b]);
      ld_g++;
      if (++ld_c == Ln.nchunks) { ld_c = 0; ld_l++; }
    }
  };
  if (tid == POL_CTRL_TID) top_up(POL_NBUF - 1);   // the first weight chunks travel while the observations are staged
  {
    int nb = 0;
    for (int l = 0; l < P.n_layers; l++) nb += P.L[l].Np;
    for (int i = tid; i < nb; i += POL_THREADS) sBias[i] = __ldg(P.bias + i);
  }
  // layer-0 A operand: normalised obs, bf16, canonical K-major layout (LBO = 128 B, SBO = Kp * 16 B). A warp store
  // covers 8 rows x 4 k-groups = 512 contiguous bytes (conflict-free); each lane converts 8 consecutive features.
  {
    const PolicyLayer& L0 = P.L[0];
    const uint32_t sbo = (uint32_t)L0.Kp * 16u;
    const int kgroups = L0.Kp >> 3;
    for (int rg = warp; rg < POL_TM / 8; rg += POL_THREADS / 32) {   // 16 row groups, one per warp
      const int r = rg * 8 + (lane & 7), row = row0 + r;
      const float* orow = obs + (size_t)row * P.obs_dim;
      for (int kg = lane >> 3; kg < kgroups; kg += 4) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int k = kg * 8 + i;
          v[i] = (row < n_rows && k < L0.K) ? (__ldg(orow + k) - __ldg(P.mean + k)) * __ldg(P.inv_std + k) : 0.f;
        }
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); pk[i] = *reinterpret_cast<const uint32_t*>(&h); }
        *reinterpret_cast<uint4*>(sA + (uint32_t)rg * sbo + (uint32_t)kg * 128u + (uint32_t)(lane & 7) * 16u) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  int g = 0;   // global chunk counter
  for (int l = 0; l < P.n_layers; l++) {
    const PolicyLayer& L = P.L[l];
    if (tid == POL_CTRL_TID) {
      const uint32_t a_sbo = (uint32_t)L.Kp * 16u, b_sbo = (uint32_t)L.Kc * 16u;
      for (int c = 0; c < L.nchunks; c++) {
        const int gc = g + c, bi = gc % POL_NBUF;
        top_up(gc + POL_NBUF - 1);
        mbar_wait(&full[bi], (uint32_t)((gc / POL_NBUF) & 1));
        tc_fence_after();
        const uint8_t* buf = sB + (size_t)bi * POL_BCHUNK;
        const int nslices = L.Kc / 16;
        for (int s = 0; s < nslices; s++) {
          const int kslice = c * nslices + s;
          const uint32_t a_addr = smem_u32(sA) + (uint32_t)kslice * 256u;   // 2 core matrices (16 k) per slice
          const uint32_t b_addr = smem_u32(buf) + (uint32_t)s * 256u;
          for (int n0 = 0; n0 < L.Np; n0 += 256) {
            const int nn = (L.Np - n0) < 256 ? (L.Np - n0) : 256;
            umma_bf16(tmem + (uint32_t)n0, make_desc(a_addr, 128u, a_sbo), make_desc(b_addr + (uint32_t)(n0 >> 3) * b_sbo, 128u, b_sbo),
                      make_idesc(POL_TM, nn), (uint32_t)(kslice > 0));
          }
        }
        umma_commit(&empty[bi]);
      }
      umma_commit(layer_done);
    }
    g += L.nchunks;
    // whole layer accumulated. A barrier of its own: the warps that do not issue run a whole layer ahead of the
    // empty[] phases, and a parity wait is only meaningful for the current or the previous phase.
    if (warp < POL_EPI_THREADS / 32) {
    mbar_wait(layer_done, (uint32_t)(l & 1));
    tc_fence_after();
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane, row = row0 + r;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const float* bias = sBias + L.b_off;
    if (l + 1 < P.n_layers) {
      const uint32_t sbo_next = (uint32_t)P.L[l + 1].Kp * 16u;   // == Np of this layer
      const int ncol = L.Np >> 2;                                  // this warp's column quarter (Np is a multiple of 128)
      for (int n0 = cq * ncol; n0 < (cq + 1) * ncol; n0 += 32) {
        float vv[32];
        tmem_ld32(trow + (uint32_t)n0, vv);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int n = n0 + 8 * j + 2 * i;
            float a = vv[8 * j + 2 * i] + bias[n], b = vv[8 * j + 2 * i + 1] + bias[n + 1];
            a = (n < L.N) ? __fdividef(a, 1.f + __expf(-a)) : 0.f;          // SiLU (swish)
            b = (n + 1 < L.N) ? __fdividef(b, 1.f + __expf(-b)) : 0.f;
            const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            pk[i] = *reinterpret_cast<const uint32_t*>(&h);
          }
          const uint32_t off = (uint32_t)(r >> 3) * sbo_next + (uint32_t)((n0 >> 3) + j) * 128u + (uint32_t)(r & 7) * 16u;
          *reinterpret_cast<uint4*>(sA + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
    } else {
      // distribution head: logits = [loc | scale_raw], brax NormalTanhDistribution(min_std = 0.001). The four warps of a
      // lane quarter split the actions; log-prob partial sums meet in shared memory (sA is free: the last MMAs are done)
      float out[32];
      tmem_ld32(trow, out);
      const int A = P.act_dim, per = (A + 3) >> 2;
      float* lp_part = reinterpret_cast<float*>(sA);            // [4][POL_TM]
      float lp = 0.f;
      if (row < n_rows) {
        for (int j = cq * per; j < (cq + 1) * per && j < A; j++) {
          float ol = 0.f, os = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i++) { ol = (i == j) ? out[i] : ol; os = (i == A + j) ? out[i] : os; }
          const float loc = ol + bias[j], sr = os + bias[A + j];
          if (logits_out) { logits_out[(size_t)row * 2 * A + j] = loc; logits_out[(size_t)row * 2 * A + A + j] = sr; }
          const float scale = (sr > 20.f ? sr : log1pf(expf(sr))) + 0.001f;
          float e = 0.f;
          if (!deterministic) e = eps_in ? eps_in[(size_t)row * A + j] : normal_draw(seed, step, (uint32_t)row, (uint32_t)j);
          const float raw = loc + scale * e;
          // log N(raw; loc, scale) - log|d tanh / d raw|, with log det = 2 (log 2 - raw - softplus(-2 raw))
          const float m2 = -2.f * raw;
          const float sp = m2 > 20.f ? m2 : log1pf(expf(m2));
          lp += -0.5f * e * e - logf(scale) - 0.9189385332046727f - 2.f * (0.6931471805599453f - raw - sp);
          action[(size_t)row * A + j] = tanhf(raw);
          if (raw_action) raw_action[(size_t)row * A + j] = raw;
        }
      }
      lp_part[cq * POL_TM + r] = lp;
      asm volatile("bar.sync 1, %0;" ::"r"(POL_EPI_THREADS) : "memory");
      if (cq == 0 && row < n_rows && log_prob) log_prob[row] = ((lp_part[r] + lp_part[POL_TM + r]) + lp_part[2 * POL_TM + r]) + lp_part[3 * POL_TM + r];
      tc_fence_before();
    }
    }   // epilogue warps
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ----------------------------------------------------------------------------------------------
// cluster variant: 4 CTAs (one thread-block cluster) share a 128-row tile. Each CTA computes a quarter of every hidden
// layer's columns (a quarter of the MMAs, of the weight stream and of the SiLU epilogue) and writes its bf16 activation
// slice into the shared memory of all four CTAs through distributed shared memory (mapa + st.shared::cluster); two
// cluster barriers per layer order "everybody's MMAs have read the old A operand" -> slice exchange -> next layer. The
// 24-wide head runs on rank 0. 4096 rows occupy 128 SMs instead of 32.
// ----------------------------------------------------------------------------------------------
#define POL_CL 4
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  // not .aligned: the control warp reaches it diverged (lane 0 issues the MMAs, lanes 1..31 arrive early)
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t cta_addr, uint32_t rank, uint4 v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(cta_addr), "r"(rank));
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __cluster_dims__(POL_CL, 1, 1) __launch_bounds__(POL_THREADS, 1)
pgtt_policy_cluster_kernel(PolicyParams P, const float* __restrict__ obs, int n_rows, unsigned long long seed, unsigned long long step_in,
                           const unsigned long long* __restrict__ step_base, int deterministic, const float* __restrict__ eps_in, float* __restrict__ action, float* __restrict__ raw_action,
                           float* __restrict__ log_prob, float* __restrict__ logits_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const unsigned long long step = step_in + (step_base ? *step_base : 0ull);   // noise counter; the base lives on the device for graph replays
  uint8_t* sA = smem;
  uint8_t* sB = smem + POL_A_BYTES;
  float* sBias = reinterpret_cast<float*>(smem + POL_A_BYTES + POL_NBUF * POL_BCHUNK);
  uint64_t* full = reinterpret_cast<uint64_t*>(sBias + POL_MAXBIAS);
  uint64_t* empty = full + POL_NBUF;
  uint64_t* layer_done = empty + POL_NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(layer_done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_rank();
  const int row0 = (blockIdx.x / POL_CL) * POL_TM;
  const int my_layers = rank == 0 ? P.n_layers : P.n_layers - 1;      // the head layer runs on rank 0 only

  if (tid == 0) {
    for (int i = 0; i < POL_NBUF; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(layer_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  int ld_l = 0, ld_c = 0, ld_g = 0;
  auto top_up = [&](int upto) {
    while (ld_l < my_layers && ld_g < upto) {
      const PolicyLayer& Ln = P.L[ld_l];
      const int b = ld_g % POL_NBUF;
      if (ld_g >= POL_NBUF) mbar_wait(&empty[b], (uint32_t)((ld_g / POL_NBUF - 1) & 1));
      const uint32_t bytes = (uint32_t)(Ln.Ns * Ln.cKc * 2);
      const __nv_bfloat16* src = P.cw + Ln.cw_off + (size_t)rank * Ln.cw_rstride + (size_t)ld_c * (bytes / 2);
      mbar_expect_tx(&full[b], bytes);
      bulk_g2s(sB + (size_t)b * POL_BCHUNK, src, bytes, &full[b]);
      ld_g++;
      if (++ld_c == Ln.cnchunks) { ld_c = 0; ld_l++; }
    }
  };
  if (tid == POL_CTRL_TID) top_up(POL_NBUF - 1);
  {
    int nb = 0;
    for (int l = 0; l < P.n_layers; l++) nb += P.L[l].Np;
    for (int i = tid; i < nb; i += POL_THREADS) sBias[i] = __ldg(P.bias + i);
  }
  {  // every CTA of the cluster stages the whole normalised observation tile (layer 0 needs all of K)
    const PolicyLayer& L0 = P.L[0];
    const uint32_t sbo = (uint32_t)L0.Kp * 16u;
    const int kgroups = L0.Kp >> 3;
    for (int rg = warp; rg < POL_TM / 8; rg += POL_THREADS / 32) {
      const int r = rg * 8 + (lane & 7), row = row0 + r;
      const float* orow = obs + (size_t)row * P.obs_dim;
      for (int kg = lane >> 3; kg < kgroups; kg += 4) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int k = kg * 8 + i;
          v[i] = (row < n_rows && k < L0.K) ? (__ldg(orow + k) - __ldg(P.mean + k)) * __ldg(P.inv_std + k) : 0.f;
        }
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); pk[i] = *reinterpret_cast<const uint32_t*>(&h); }
        *reinterpret_cast<uint4*>(sA + (uint32_t)rg * sbo + (uint32_t)kg * 128u + (uint32_t)(lane & 7) * 16u) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  int g = 0;
  for (int l = 0; l < P.n_layers; l++) {
    const PolicyLayer& L = P.L[l];
    const bool last = l + 1 == P.n_layers;
    if (last && rank != 0) break;                       // ranks 1..3 are done: nobody touches their shared memory any more
    if (tid == POL_CTRL_TID) {
      const uint32_t a_sbo = (uint32_t)L.Kp * 16u, b_sbo = (uint32_t)L.cKc * 16u;
      for (int c = 0; c < L.cnchunks; c++) {
        const int gc = g + c, bi = gc % POL_NBUF;
        top_up(gc + POL_NBUF - 1);
        mbar_wait(&full[bi], (uint32_t)((gc / POL_NBUF) & 1));
        tc_fence_after();
        const uint8_t* buf = sB + (size_t)bi * POL_BCHUNK;
        const int nslices = L.cKc / 16;
        for (int s = 0; s < nslices; s++) {
          const int kslice = c * nslices + s;
          umma_bf16(tmem, make_desc(smem_u32(sA) + (uint32_t)kslice * 256u, 128u, a_sbo), make_desc(smem_u32(buf) + (uint32_t)s * 256u, 128u, b_sbo),
                    make_idesc(POL_TM, L.Ns), (uint32_t)(kslice > 0));
        }
        umma_commit(&empty[bi]);
      }
      umma_commit(layer_done);
    }
    g += L.cnchunks;
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane, row = row0 + r;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const bool epi = warp < POL_EPI_THREADS / 32;
    if (epi) { mbar_wait(layer_done, (uint32_t)(l & 1)); tc_fence_after(); }
    if (!last) {
      // (A) every CTA of the cluster has finished the MMAs that read the current A operand
      cluster_sync_all();
      if (epi) {
        const float* bias = sBias + L.b_off + rank * L.Ns;
        const uint32_t sbo_next = (uint32_t)P.L[l + 1].Kp * 16u;     // == Np of this layer
        const bool to_all = l + 2 < P.n_layers;                        // the head's input is only needed by rank 0
        for (int n0 = cq * 32; n0 < L.Ns; n0 += 128) {
          float vv[32];
          tmem_ld32(trow + (uint32_t)n0, vv);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int n = n0 + 8 * j + 2 * i, ng = rank * L.Ns + n;
              float a = vv[8 * j + 2 * i] + bias[n], b = vv[8 * j + 2 * i + 1] + bias[n + 1];
              a = (ng < L.N) ? __fdividef(a, 1.f + __expf(-a)) : 0.f;          // SiLU (swish)
              b = (ng + 1 < L.N) ? __fdividef(b, 1.f + __expf(-b)) : 0.f;
              const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
              pk[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            const uint32_t kg = (uint32_t)((rank * L.Ns + n0) >> 3) + (uint32_t)j;
            const uint32_t addr = smem_u32(sA) + (uint32_t)(r >> 3) * sbo_next + kg * 128u + (uint32_t)(r & 7) * 16u;
            const uint4 val = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (to_all) {
#pragma unroll
              for (int d = 0; d < POL_CL; d++) st_cluster_v4(addr, (uint32_t)d, val);
            } else {
              st_cluster_v4(addr, 0u, val);
            }
          }
        }
        asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy stores (local and remote) -> the async proxy that feeds the MMAs
        tc_fence_before();
      }
      // (B) all slices have arrived everywhere
      cluster_sync_all();
      tc_fence_after();
    } else {
      if (epi) {
        const float* bias = sBias + L.b_off;
        float out[32];
        tmem_ld32(trow, out);
        const int A = P.act_dim, per = (A + 3) >> 2;
        float* lp_part = reinterpret_cast<float*>(sA);
        float lp = 0.f;
        if (row < n_rows) {
          for (int j = cq * per; j < (cq + 1) * per && j < A; j++) {
            float ol = 0.f, os = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i++) { ol = (i == j) ? out[i] : ol; os = (i == A + j) ? out[i] : os; }
            const float loc = ol + bias[j], sr = os + bias[A + j];
            if (logits_out) { logits_out[(size_t)row * 2 * A + j] = loc; logits_out[(size_t)row * 2 * A + A + j] = sr; }
            const float scale = (sr > 20.f ? sr : log1pf(expf(sr))) + 0.001f;
            float e = 0.f;
            if (!deterministic) e = eps_in ? eps_in[(size_t)row * A + j] : normal_draw(seed, step, (uint32_t)row, (uint32_t)j);
            const float raw = loc + scale * e;
            const float m2 = -2.f * raw;
            const float sp = m2 > 20.f ? m2 : log1pf(expf(m2));
            lp += -0.5f * e * e - logf(scale) - 0.9189385332046727f - 2.f * (0.6931471805599453f - raw - sp);
            action[(size_t)row * A + j] = tanhf(raw);
            if (raw_action) raw_action[(size_t)row * A + j] = raw;
          }
        }
        lp_part[cq * POL_TM + r] = lp;
        asm volatile("bar.sync 1, %0;" ::"r"(POL_EPI_THREADS) : "memory");
        if (cq == 0 && row < n_rows && log_prob) log_prob[row] = ((lp_part[r] + lp_part[POL_TM + r]) + lp_part[2 * POL_TM + r]) + lp_part[3 * POL_TM + r];
        tc_fence_before();
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// ----------------------------------------------------------------------------------------------
// host side + C ABI
// ----------------------------------------------------------------------------------------------
struct pgtt_policy {
  int device;
  PolicyParams P;
  __nv_bfloat16* cw_dev;
  size_t cw_elems;
  int use_cluster;   // 1 / 0 forced by PGTT_POLICY_CLUSTER, -1 = by row count
  int n_sm;
  // rollout graph: the T-step unroll (1 + 3 T launches) captured once per (env, buffers, T, deterministic) and replayed
  const unsigned long long* step_base_arg;   // what pgtt_policy_act passes to the kernels (NULL outside a rollout graph)
  unsigned long long* step_dev;              // device-side noise counter base of the graph
  cudaStream_t gstream;
  cudaEvent_t ev_in, ev_out;
  cudaGraphExec_t gexec;
  struct { pgtt_env* env; uint64_t env_serial; int T, deterministic; uint64_t seed; pgtt_rollout_buffers o; } gkey;
  int use_graph;
  __nv_bfloat16* w_dev;
  float *bias_dev, *mean_dev, *istd_dev;
  size_t w_elems, b_elems;
  bool has_params;
  int64_t launches;
};

static thread_local std::string g_perr;
static int pfail(int code, const std::string& m) { g_perr = m; return code; }
#define PCUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return pfail(PGTT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

static uint16_t f2bf(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

extern "C" {

const char* pgtt_policy_last_error(void) { return g_perr.c_str(); }

int pgtt_policy_create(int device, const int* sizes, int n_layers, pgtt_policy** out) {
  if (!sizes || !out || n_layers < 2 || n_layers > POL_MAXLAYERS) return pfail(PGTT_ERR_ARG, "pgtt_policy_create: need 2..6 layers");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return pfail(PGTT_ERR_CUDA, "pgtt_policy_create: no CUDA device");
  if (device < 0 || device >= ndev) return pfail(PGTT_ERR_ARG, "pgtt_policy_create: bad device");
  pgtt_policy* p = new pgtt_policy();
  memset(&p->P, 0, sizeof(p->P));
  p->device = device; p->has_params = false; p->launches = 0;
  p->w_dev = nullptr; p->cw_dev = nullptr; p->bias_dev = p->mean_dev = p->istd_dev = nullptr;
  PolicyParams& P = p->P;
  P.n_layers = n_layers; P.obs_dim = sizes[0]; P.act_dim = sizes[n_layers] / 2;
  size_t w_off = 0, b_off = 0;
  for (int l = 0; l < n_layers; l++) {
    PolicyLayer& L = P.L[l];
    L.K = sizes[l]; L.N = sizes[l + 1];
    L.Np = (L.N + 15) / 16 * 16;
    int kc = (POL_BCHUNK / 2 / L.Np) / 16 * 16;
    const int kp16 = (L.K + 15) / 16 * 16;
    if (kc > kp16) kc = kp16;
    if (kc < 16 || L.Np > 512 || (l > 0 && L.K != P.L[l - 1].N)) { delete p; return pfail(PGTT_ERR_ARG, "pgtt_policy_create: unsupported layer sizes"); }
    L.Kc = kc;
    L.Kp = (l == 0) ? (L.K + kc - 1) / kc * kc : P.L[l - 1].Np;   // layer l > 0 reads the padded output of layer l - 1
    if (L.Kp % kc != 0) { L.Kc = 16; }
    L.nchunks = L.Kp / L.Kc;
    if (L.Kp > POL_MAXK) { delete p; return pfail(PGTT_ERR_ARG, "pgtt_policy_create: layer wider than 512"); }
    L.w_off = w_off; L.b_off = b_off;
    w_off += (size_t)L.Np * L.Kp; b_off += L.Np;
  }
  {  // cluster layout: hidden layers split 4 ways over the columns, the head whole on rank 0
    size_t cw_off = 0;
    for (int l = 0; l < n_layers; l++) {
      PolicyLayer& L = P.L[l];
      const bool head = l + 1 == n_layers;
      L.Ns = head ? L.Np : L.Np / POL_CL;
      int kc = (POL_BCHUNK / 2 / L.Ns) / 16 * 16;
      if (kc > L.Kp) kc = L.Kp;
      while (kc > 16 && L.Kp % kc != 0) kc -= 16;
      L.cKc = kc; L.cnchunks = L.Kp / kc;
      L.cw_off = cw_off; L.cw_rstride = head ? 0 : (size_t)L.Ns * L.Kp;
      cw_off += (size_t)L.Ns * L.Kp * (head ? 1 : POL_CL);
    }
    p->cw_elems = cw_off;
  }
  for (int l = 0; l + 1 < n_layers; l++)
    if (P.L[l].Np % 128 != 0) { delete p; return pfail(PGTT_ERR_ARG, "pgtt_policy_create: hidden widths must be multiples of 128"); }
  if (b_off > POL_MAXBIAS) { delete p; return pfail(PGTT_ERR_ARG, "pgtt_policy_create: total layer width above 1024"); }
  if (P.L[n_layers - 1].Np > 32 || sizes[n_layers] % 2) { delete p; return pfail(PGTT_ERR_ARG, "pgtt_policy_create: head must be 2 * act_dim <= 32"); }
  p->w_elems = w_off; p->b_elems = b_off;
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p->w_dev, w_off * 2) != cudaSuccess || cudaMalloc(&p->bias_dev, b_off * 4) != cudaSuccess ||
      cudaMalloc(&p->mean_dev, P.obs_dim * 4) != cudaSuccess || cudaMalloc(&p->istd_dev, P.obs_dim * 4) != cudaSuccess ||
      cudaMalloc(&p->cw_dev, p->cw_elems * 2) != cudaSuccess ||
      cudaFuncSetAttribute(pgtt_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POL_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(pgtt_policy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POL_SMEM) != cudaSuccess) {
    delete p; return pfail(PGTT_ERR_CUDA, std::string("pgtt_policy_create: ") + cudaGetErrorString(cudaGetLastError()));
  }
  P.w = p->w_dev; P.cw = p->cw_dev; P.bias = p->bias_dev; P.mean = p->mean_dev; P.inv_std = p->istd_dev;
  // cluster kernel while its 4 CTAs per tile fit the SMs in one wave (<= 4736 rows on 148 SMs: 42 us vs 57 us); beyond that
  // the single-CTA-per-tile kernel has fewer waves (8192 rows: 59 us vs ~75 us). PGTT_POLICY_CLUSTER=0|1 forces one.
  p->use_cluster = -1;
  if (const char* e = getenv("PGTT_POLICY_CLUSTER")) p->use_cluster = atoi(e) != 0;
  p->n_sm = 148;
  cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, device);
  p->step_base_arg = nullptr; p->step_dev = nullptr; p->gstream = nullptr; p->ev_in = p->ev_out = nullptr; p->gexec = nullptr;
  memset(&p->gkey, 0, sizeof(p->gkey));
  p->use_graph = 1;
  if (const char* e = getenv("PGTT_ROLLOUT_GRAPH")) p->use_graph = atoi(e) != 0;
  *out = p;
  return PGTT_OK;
}

int pgtt_policy_destroy(pgtt_policy* p) {
  if (!p) return PGTT_OK;
  cudaSetDevice(p->device); cudaDeviceSynchronize();
  if (p->gexec) cudaGraphExecDestroy(p->gexec);
  if (p->gstream) cudaStreamDestroy(p->gstream);
  if (p->ev_in) cudaEventDestroy(p->ev_in);
  if (p->ev_out) cudaEventDestroy(p->ev_out);
  cudaFree(p->step_dev);
  cudaFree(p->cw_dev); cudaFree(p->w_dev); cudaFree(p->bias_dev); cudaFree(p->mean_dev); cudaFree(p->istd_dev);
  delete p;
  return PGTT_OK;
}

// kernels[l]: HOST fp32 [in][out] row-major (brax / flax `kernel` layout), biases[l]: HOST fp32 [out];
// obs_mean / obs_std: HOST fp32 [obs_dim] (running-statistics normaliser; NULL = identity)
int pgtt_policy_set_params(pgtt_policy* p, const float* const* kernels, const float* const* biases, const float* obs_mean, const float* obs_std) {
  if (!p || !kernels || !biases) return pfail(PGTT_ERR_ARG, "pgtt_policy_set_params: null argument");
  const PolicyParams& P = p->P;
  std::vector<uint16_t> w(p->w_elems, 0);
  std::vector<float> b(p->b_elems, 0.f), mean(P.obs_dim, 0.f), istd(P.obs_dim, 1.f);
  for (int l = 0; l < P.n_layers; l++) {
    const PolicyLayer& L = P.L[l];
    const size_t chunk_elems = (size_t)L.Np * L.Kc;
    for (int k = 0; k < L.K; k++)
      for (int n = 0; n < L.N; n++) {
        const int c = k / L.Kc, kk = k % L.Kc;
        const size_t off_bytes = (size_t)(n >> 3) * (L.Kc * 16) + (size_t)(kk >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(kk & 7) * 2;
        w[L.w_off + c * chunk_elems + off_bytes / 2] = f2bf(kernels[l][(size_t)k * L.N + n]);
      }
    for (int n = 0; n < L.N; n++) b[L.b_off + n] = biases[l][n];
  }
  std::vector<uint16_t> cw(p->cw_elems, 0);
  for (int l = 0; l < P.n_layers; l++) {
    const PolicyLayer& L = P.L[l];
    const size_t chunk_elems = (size_t)L.Ns * L.cKc;
    for (int k = 0; k < L.K; k++)
      for (int n = 0; n < L.N; n++) {
        const int rank = n / L.Ns, nn = n % L.Ns, c = k / L.cKc, kk = k % L.cKc;
        const size_t off_bytes = (size_t)(nn >> 3) * (L.cKc * 16) + (size_t)(kk >> 3) * 128 + (size_t)(nn & 7) * 16 + (size_t)(kk & 7) * 2;
        cw[L.cw_off + (size_t)rank * L.cw_rstride + c * chunk_elems + off_bytes / 2] = f2bf(kernels[l][(size_t)k * L.N + n]);
      }
  }
  if (obs_mean) for (int i = 0; i < P.obs_dim; i++) mean[i] = obs_mean[i];
  if (obs_std) for (int i = 0; i < P.obs_dim; i++) istd[i] = 1.0f / obs_std[i];
  PCUDA(cudaSetDevice(p->device));
  PCUDA(cudaDeviceSynchronize());
  PCUDA(cudaMemcpy(p->w_dev, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  PCUDA(cudaMemcpy(p->cw_dev, cw.data(), cw.size() * 2, cudaMemcpyHostToDevice));
  PCUDA(cudaMemcpy(p->bias_dev, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  PCUDA(cudaMemcpy(p->mean_dev, mean.data(), mean.size() * 4, cudaMemcpyHostToDevice));
  PCUDA(cudaMemcpy(p->istd_dev, istd.data(), istd.size() * 4, cudaMemcpyHostToDevice));
  p->has_params = true;
  return PGTT_OK;
}

// The same packing from DEVICE fp32 parameters, stream-ordered: what the learner calls after every training step (the host variant
// costs eight device-to-host copies, a host repack and a device synchronisation: 1.5 ms per training step).
struct PolicySrc { const float* k[POL_MAXLAYERS]; const float* b[POL_MAXLAYERS]; const float* mean; const float* std; };
__global__ void pgtt_policy_pack_kernel(PolicyParams P, PolicySrc S, __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ cw, float* __restrict__ bias,
                                        float* __restrict__ mean, float* __restrict__ istd) {
  const int l = blockIdx.y;
  const PolicyLayer& L = P.L[l];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const size_t chunk_elems = (size_t)L.Np * L.Kc, cchunk_elems = (size_t)L.Ns * L.cKc;
  for (int i = tid; i < L.K * L.N; i += nth) {
    const int k = i / L.N, n = i % L.N;
    const __nv_bfloat16 v = __float2bfloat16_rn(S.k[l][i]);
    {
      const int c = k / L.Kc, kk = k % L.Kc;
      const size_t off_bytes = (size_t)(n >> 3) * (L.Kc * 16) + (size_t)(kk >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(kk & 7) * 2;
      w[L.w_off + c * chunk_elems + off_bytes / 2] = v;
    }
    if (cw) {
      const int rank = n / L.Ns, nn = n % L.Ns, c = k / L.cKc, kk = k % L.cKc;
      const size_t off_bytes = (size_t)(nn >> 3) * (L.cKc * 16) + (size_t)(kk >> 3) * 128 + (size_t)(nn & 7) * 16 + (size_t)(kk & 7) * 2;
      cw[L.cw_off + (size_t)rank * L.cw_rstride + c * cchunk_elems + off_bytes / 2] = v;
    }
  }
  for (int n = tid; n < L.N; n += nth) bias[L.b_off + n] = S.b[l][n];
  if (l == 0)
    for (int i = tid; i < P.obs_dim; i += nth) { mean[i] = S.mean ? S.mean[i] : 0.f; istd[i] = S.std ? 1.0f / S.std[i] : 1.f; }
}

int pgtt_policy_set_params_device(pgtt_policy* p, const float* const* kernels, const float* const* biases, const float* obs_mean, const float* obs_std, void* stream) {
  if (!p || !kernels || !biases) return pfail(PGTT_ERR_ARG, "pgtt_policy_set_params_device: null argument");
  const PolicyParams& P = p->P;
  PolicySrc S = {};
  for (int l = 0; l < P.n_layers; l++) {
    if (!kernels[l] || !biases[l]) return pfail(PGTT_ERR_ARG, "pgtt_policy_set_params_device: null parameter");
    S.k[l] = kernels[l]; S.b[l] = biases[l];
  }
  S.mean = obs_mean; S.std = obs_std;
  cudaStream_t st = (cudaStream_t)stream;
  if (!p->has_params) {                                     // the padding of the packed arrays is zero and stays zero
    PCUDA(cudaMemsetAsync(p->w_dev, 0, p->w_elems * 2, st));
    if (p->cw_dev) PCUDA(cudaMemsetAsync(p->cw_dev, 0, p->cw_elems * 2, st));
    PCUDA(cudaMemsetAsync(p->bias_dev, 0, p->b_elems * 4, st));
  }
  pgtt_policy_pack_kernel<<<dim3(64, P.n_layers), 256, 0, st>>>(P, S, p->w_dev, p->cw_dev, p->bias_dev, p->mean_dev, p->istd_dev);
  PCUDA(cudaGetLastError());
  p->has_params = true;
  return PGTT_OK;
}

// obs DEVICE [n][obs_dim]; eps DEVICE [n][act_dim] or NULL (internal counter-based normal draws keyed by
// seed/step); outputs DEVICE: action [n][act_dim], raw_action [n][act_dim] or NULL, log_prob [n] or NULL,
// logits [n][2 act_dim] or NULL.
int pgtt_policy_act(pgtt_policy* p, const float* obs, int n, uint64_t seed, uint64_t step, int deterministic, const float* eps,
                    float* action, float* raw_action, float* log_prob, float* logits, void* stream) {
  if (!p || !obs || !action || n <= 0) return pfail(PGTT_ERR_ARG, "pgtt_policy_act: null argument or n <= 0");
  if (!p->has_params) return pfail(PGTT_ERR_STATE, "pgtt_policy_act: pgtt_policy_set_params first");
  const int blocks = (n + POL_TM - 1) / POL_TM;
  const bool cluster = p->use_cluster >= 0 ? p->use_cluster != 0 : blocks * POL_CL <= p->n_sm;
  if (cluster)
    pgtt_policy_cluster_kernel<<<blocks * POL_CL, POL_THREADS, POL_SMEM, (cudaStream_t)stream>>>(p->P, obs, n, (unsigned long long)seed, (unsigned long long)step,
                                                                                          p->step_base_arg, deterministic, eps, action, raw_action, log_prob, logits);
  else
    pgtt_policy_kernel<<<blocks, POL_THREADS, POL_SMEM, (cudaStream_t)stream>>>(p->P, obs, n, (unsigned long long)seed, (unsigned long long)step,
                                                                          p->step_base_arg, deterministic, eps, action, raw_action, log_prob, logits);
  PCUDA(cudaGetLastError());
  p->launches++;
  return PGTT_OK;
}

int64_t pgtt_policy_launch_count(pgtt_policy* p) { return p ? p->launches : 0; }

// brax compute_gae: backward recursion over one trajectory segment per thread (coalesced across segments). `mom` (may be null;
// needs the whole batch in ONE block): population mean and standard deviation of the advantages, two-pass (mean, then squared
// deviations), every partial sum in a fixed order - what ppo_loss normalises the advantages with.
__global__ void pgtt_gae_kernel(const float* __restrict__ trunc, const float* __restrict__ disc, const float* __restrict__ rew,
                                const float* __restrict__ val, int T, int B, float lam, float gamma, float rscale,
                                float* __restrict__ vs, float* __restrict__ adv, float* __restrict__ mom, double* __restrict__ sums3) {
  __shared__ float red[33];
  __shared__ double dred[2][32];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float asum = 0.f;
  if (b < B) {
    float acc = 0.f, vs_next = val[(size_t)T * B + b];
    for (int t = T - 1; t >= 0; t--) {
      const size_t i = (size_t)t * B + b;
      const float tr = trunc[i], mask = 1.f - tr, term = (1.f - disc[i]) * mask, r = rew[i] * rscale, v = val[i];
      const float cont = gamma * (1.f - term);
      const float delta = (r + cont * val[i + B] - v) * mask;
      acc = delta + cont * mask * lam * acc;
      const float vst = acc + v;
      const float a = (r + cont * vs_next - v) * mask;
      adv[i] = a;
      asum += a;
      vs[i] = vst;
      vs_next = vst;
    }
  }
  if (sums3) {        // multi-rank: (count, sum, sum of squares) of this rank's advantages in float64; the ranks add them up (NCCL) and pgtt_moments_finalize forms mean / std
    double s1 = 0.0, s2 = 0.0;
    if (b < B) for (int t = 0; t < T; t++) { const double a = (double)adv[(size_t)t * B + b]; s1 += a; s2 += a * a; }
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { dred[0][threadIdx.x >> 5] = s1; dred[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a1 = 0.0, a2 = 0.0;
      for (int w = 0; w < (int)((blockDim.x + 31) / 32); w++) { a1 += dred[0][w]; a2 += dred[1][w]; }
      sums3[0] = (double)T * (double)B; sums3[1] = a1; sums3[2] = a2;
    }
    return;
  }
  if (!mom) return;                                          // (uniform)
  auto block_sum = [&](float v) -> float {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                         // red[] of the previous call has been read
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      float w = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.f;
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
      if (threadIdx.x == 0) red[32] = w;
    }
    __syncthreads();
    return red[32];
  };
  const float n = (float)T * (float)B;
  const float mean = block_sum(asum) / n;
  float dev = 0.f;
  if (b < B) for (int t = 0; t < T; t++) { const float d = adv[(size_t)t * B + b] - mean; dev += d * d; }   // (own writes: visible to this thread)
  const float var = block_sum(dev) / n;
  if (threadIdx.x == 0) { mom[0] = mean; mom[1] = sqrtf(var); }
}

int pgtt_gae(const float* truncation, const float* discount, const float* reward, const float* values, int T, int B, float lambda, float gamma,
             float reward_scaling, float* vs, float* adv, void* stream) {
  if (!truncation || !discount || !reward || !values || !vs || !adv || T <= 0 || B <= 0) return pfail(PGTT_ERR_ARG, "pgtt_gae: null argument or empty shape");
  pgtt_gae_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(truncation, discount, reward, values, T, B, lambda, gamma, reward_scaling, vs, adv, nullptr, nullptr);
  PCUDA(cudaGetLastError());
  return PGTT_OK;
}

// multi-rank variant of pgtt_gae_moments: this rank's (count, sum, sum of squares) of the advantages, float64, from the same launch; after the all-reduce
// pgtt_moments_finalize turns the global sums into (mean, std)
int pgtt_gae_sums(const float* truncation, const float* discount, const float* reward, const float* values, int T, int B, float lambda, float gamma,
                  float reward_scaling, float* vs, float* adv, double* sums3, void* stream) {
  if (!truncation || !discount || !reward || !values || !vs || !adv || !sums3 || T <= 0 || B <= 0) return pfail(PGTT_ERR_ARG, "pgtt_gae_sums: null argument or empty shape");
  if (B > 1024) return pfail(PGTT_ERR_ARG, "pgtt_gae_sums: more than 1024 segments per minibatch");
  pgtt_gae_kernel<<<1, (B + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(truncation, discount, reward, values, T, B, lambda, gamma, reward_scaling, vs, adv, nullptr, sums3);
  PCUDA(cudaGetLastError());
  return PGTT_OK;
}
__global__ void pgtt_moments_finalize_kernel(const double* __restrict__ s, float* __restrict__ mom) {
  const double mean = s[1] / s[0], var = s[2] / s[0] - mean * mean;
  mom[0] = (float)mean; mom[1] = (float)sqrt(var > 0.0 ? var : 0.0);
}
int pgtt_moments_finalize(const double* sums3, float* moments, void* stream) {
  if (!sums3 || !moments) return pfail(PGTT_ERR_ARG, "pgtt_moments_finalize: null argument");
  pgtt_moments_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums3, moments);
  PCUDA(cudaGetLastError());
  return PGTT_OK;
}

int pgtt_gae_moments(const float* truncation, const float* discount, const float* reward, const float* values, int T, int B, float lambda, float gamma,
                     float reward_scaling, float* vs, float* adv, float* moments, void* stream) {
  if (!truncation || !discount || !reward || !values || !vs || !adv || !moments || T <= 0 || B <= 0) return pfail(PGTT_ERR_ARG, "pgtt_gae_moments: null argument or empty shape");
  if (B > 1024) return pfail(PGTT_ERR_ARG, "pgtt_gae_moments: more than 1024 segments per minibatch (use pgtt_gae and reduce separately)");
  pgtt_gae_kernel<<<1, (B + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(truncation, discount, reward, values, T, B, lambda, gamma, reward_scaling, vs, adv, moments, nullptr);
  PCUDA(cudaGetLastError());
  return PGTT_OK;
}

// Column sums and sums of squares of an fp32 matrix [rows][cols] (row stride ld) in float64 - what brax's running_statistics.update needs of a
// batch of observations (training/train.py:140 normalize_observations=True). Two launches, fixed summation order (deterministic): per-block
// partials over a strided share of the rows (one column per thread, coalesced along the row), then one thread per column adds the partials.
#define CM_BLOCKS 592
#define CM_THREADS 256
__global__ void __launch_bounds__(CM_THREADS) pgtt_col_moments_kernel(const float* __restrict__ x, long long rows, int cols, int ld, double* __restrict__ part) {
  for (int c = threadIdx.x; c < cols; c += CM_THREADS) {
    double s = 0.0, q = 0.0;
    for (long long r = blockIdx.x; r < rows; r += CM_BLOCKS) { const double v = (double)__ldg(x + r * ld + c); s += v; q += v * v; }
    part[((size_t)blockIdx.x * 2) * cols + c] = s;
    part[((size_t)blockIdx.x * 2 + 1) * cols + c] = q;
  }
}
__global__ void pgtt_col_moments_sum_kernel(const double* __restrict__ part, int cols, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // i < 2 * cols: sums then sums of squares
  if (i >= 2 * cols) return;
  const int which = i / cols, c = i % cols;
  double s = 0.0;
  for (int b = 0; b < CM_BLOCKS; b++) s += part[((size_t)b * 2 + which) * cols + c];
  out[i] = s;
}
long long pgtt_col_moments_scratch_doubles(int cols) { return (long long)CM_BLOCKS * 2 * cols; }
int pgtt_col_moments(const float* x, long long rows, int cols, int ld, double* out, double* scratch, void* stream) {
  if (!x || !out || !scratch || rows < 1 || cols < 1 || ld < cols) return pfail(PGTT_ERR_ARG, "pgtt_col_moments: bad argument");
  pgtt_col_moments_kernel<<<CM_BLOCKS, CM_THREADS, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, scratch);
  PCUDA(cudaGetLastError());
  pgtt_col_moments_sum_kernel<<<(2 * cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scratch, cols, out);
  PCUDA(cudaGetLastError());
  return PGTT_OK;
}

// Everything of a minibatch that is not an observation, in one launch: segment ids idx[j] = perm[mbi][j], then for t < T, j < mb
// the raw actions [T][S][A] -> [T][mb][A], the n_scal per-transition scalars [n_scal][T][S] -> [n_scal][T][mb] and the entropy noise
// of minibatch mbi, eps_all[mbi] ([T][mb][A], copied). (brax sgd_step's `convert_data` + minibatch slicing, training/train.py:135-161.)
__global__ void pgtt_minibatch_gather_kernel(const long long* __restrict__ perm, const long long* __restrict__ mbi, int mb, int S, int T, int A, int n_scal,
                                             const float* __restrict__ raw_all, const float* __restrict__ scal_all, const float* __restrict__ eps_all,
                                             long long* __restrict__ idx, float* __restrict__ raw, float* __restrict__ scal, float* __restrict__ eps) {
  const long long m = *mbi;
  const long long* pm = perm + m * mb;
  const long long n_raw = (long long)T * mb * A, n_sc = (long long)n_scal * T * mb;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n_raw + n_sc + mb; i += (long long)gridDim.x * blockDim.x) {
    if (i < n_raw) {
      const int a = (int)(i % A); const long long tj = i / A; const int j = (int)(tj % mb); const long long t = tj / mb;
      raw[i] = raw_all[(t * S + pm[j]) * A + a];
    } else if (i < 2 * n_raw) {
      eps[i - n_raw] = eps_all[m * n_raw + (i - n_raw)];
    } else if (i < 2 * n_raw + n_sc) {
      const long long k = i - 2 * n_raw; const int j = (int)(k % mb); const long long ct = k / mb;     // ct = c * T + t
      scal[k] = scal_all[ct * S + pm[j]];
    } else {
      const long long j = i - 2 * n_raw - n_sc;
      idx[j] = pm[j];
    }
  }
}

int pgtt_minibatch_gather(const long long* perm, const long long* mbi, int mb, int S, int T, int A, int n_scal, const float* raw_all, const float* scal_all,
                          const float* eps_all, long long* idx, float* raw, float* scal, float* eps, void* stream) {
  if (!perm || !mbi || !raw_all || !scal_all || !eps_all || !idx || !raw || !scal || !eps || mb < 1 || S < 1 || T < 1 || A < 1 || n_scal < 1)
    return pfail(PGTT_ERR_ARG, "pgtt_minibatch_gather: null argument or empty shape");
  const long long n = 2LL * T * mb * A + (long long)n_scal * T * mb + mb;
  pgtt_minibatch_gather_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, (cudaStream_t)stream>>>(perm, mbi, mb, S, T, A, n_scal, raw_all, scal_all, eps_all,
                                                                                                                      idx, raw, scal, eps);
  PCUDA(cudaGetLastError());
  return PGTT_OK;
}

// Fused PPO head: one thread per transition; forward terms and gradients (see include/pgtt_b200.h)
#define HEAD_MAXA 16
__global__ void pgtt_ppo_head_kernel(const float* __restrict__ logits, const float* __restrict__ baseline, const float* __restrict__ raw,
                                     const float* __restrict__ old_lp, const float* __restrict__ adv, const float* __restrict__ vs,
                                     const float* __restrict__ eps, const float* __restrict__ mom, int M, int A, float clip_eps, float c_ent,
                                     float min_std, float* __restrict__ g_logits, float* __restrict__ g_base, float* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float pol = 0.f, vl = 0.f, ent = 0.f;
  if (i < M) {
    const float invM = 1.0f / (float)M;
    const float* lg = logits + (size_t)i * 2 * A;
    float scale[HEAD_MAXA], z[HEAD_MAXA], sig[HEAD_MAXA], th[HEAD_MAXA];
    float lp = 0.f, e = 0.f;
    for (int j = 0; j < A; j++) {
      const float loc = lg[j], sr = lg[A + j], x = raw[(size_t)i * A + j], ep = eps[(size_t)i * A + j];
      const float sp = sr > 20.f ? sr : log1pf(expf(sr));
      scale[j] = sp + min_std;
      sig[j] = 1.f / (1.f + expf(-sr));
      z[j] = (x - loc) / scale[j];
      const float m2 = -2.f * x;
      lp += -0.5f * z[j] * z[j] - logf(scale[j]) - 0.9189385332046727f - 2.f * (0.6931471805599453f - x - (m2 > 20.f ? m2 : log1pf(expf(m2))));
      const float s = loc + scale[j] * ep, n2 = -2.f * s;
      e += 1.4189385332046727f + logf(scale[j]) + 2.f * (0.6931471805599453f - s - (n2 > 20.f ? n2 : log1pf(expf(n2))));
      th[j] = tanhf(s);
    }
    const float a = (adv[i] - mom[0]) / (mom[1] + 1e-8f);
    const float rho = expf(lp - old_lp[i]);
    const float rc = fminf(fmaxf(rho, 1.f - clip_eps), 1.f + clip_eps);
    const float s1 = rho * a, s2 = rc * a;
    pol = -fminf(s1, s2) * invM;
    // d min(s1, s2) / d rho: s1 branch -> a; s2 branch -> a inside the clip range, 0 outside; a tie splits evenly (torch.minimum)
    const float in_range = (rho >= 1.f - clip_eps && rho <= 1.f + clip_eps) ? 1.f : 0.f;
    const float dmin = s1 < s2 ? a : (s1 > s2 ? a * in_range : 0.5f * (a + a * in_range));
    const float dlp = -dmin * rho * invM;          // d total / d log-prob
    const float verr = vs[i] - baseline[i];
    vl = 0.25f * verr * verr * invM;
    g_base[i] = -0.5f * verr * invM;
    ent = e * invM;
    const float ce = c_ent * invM;
    float* gl = g_logits + (size_t)i * 2 * A;
    for (int j = 0; j < A; j++) {
      const float ep = eps[(size_t)i * A + j];
      gl[j] = dlp * (z[j] / scale[j]) + ce * 2.f * th[j];
      gl[A + j] = (dlp * ((z[j] * z[j] - 1.f) / scale[j]) - ce * (1.f / scale[j] - 2.f * th[j] * ep)) * sig[j];
    }
  }
  // block sums -> 4 atomics per warp
  float tot = pol + vl - c_ent * ent;
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, o); pol += __shfl_xor_sync(0xffffffffu, pol, o);
    vl += __shfl_xor_sync(0xffffffffu, vl, o); ent += __shfl_xor_sync(0xffffffffu, ent, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(sums, tot); atomicAdd(sums + 1, pol); atomicAdd(sums + 2, vl); atomicAdd(sums + 3, ent); }
}

int pgtt_ppo_head(const float* logits, const float* baseline, const float* raw_action, const float* old_log_prob, const float* adv, const float* vs,
                  const float* eps, const float* adv_moments, int M, int A, float clip_eps, float entropy_cost, float min_std, float* grad_logits,
                  float* grad_baseline, float* sums, void* stream) {
  if (!logits || !baseline || !raw_action || !old_log_prob || !adv || !vs || !eps || !adv_moments || !grad_logits || !grad_baseline || !sums || M <= 0)
    return pfail(PGTT_ERR_ARG, "pgtt_ppo_head: null argument or M <= 0");
  if (A < 1 || A > HEAD_MAXA) return pfail(PGTT_ERR_ARG, "pgtt_ppo_head: action dim must be in 1..16");
  PCUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(float), (cudaStream_t)stream));
  pgtt_ppo_head_kernel<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logits, baseline, raw_action, old_log_prob, adv, vs, eps, adv_moments, M, A,
                                                                        clip_eps, entropy_cost, min_std, grad_logits, grad_baseline, sums);
  PCUDA(cudaGetLastError());
  return PGTT_OK;
}

// Optimiser step over one flat parameter vector: global-norm clip + Adam in two launches (see include/pgtt_b200.h).
// Launch 1: per-block partial sums of g^2 (fixed grid, fixed order: the norm is deterministic) and the step counter;
// launch 2: every block re-reduces the partials in the same order, then updates its grid-stride share.
#define ADAM_BLOCKS 296
#define ADAM_THREADS 256
__global__ void __launch_bounds__(ADAM_THREADS) pgtt_adam_sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial,
                                                                      float* __restrict__ step) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * ADAM_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * ADAM_THREADS) { const float x = g[i]; s += x * x; }
  __shared__ float sh[ADAM_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < ADAM_THREADS / 32; w++) t += sh[w];
    partial[blockIdx.x] = t;
    if (blockIdx.x == 0) *step += 1.f;
  }
}

__global__ void __launch_bounds__(ADAM_THREADS) pgtt_adam_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                                       float* __restrict__ v, const float* __restrict__ partial, const float* __restrict__ step,
                                                                       long long n, float lr, float b1, float b2, float eps, float max_norm, float grad_scale) {
  __shared__ float coef_s;
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int i = threadIdx.x; i < ADAM_BLOCKS; i += 32) s += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      const float norm = sqrtf(s) * grad_scale;   // norm of the scaled gradient
      coef_s = grad_scale * (max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f);
    }
  }
  __syncthreads();
  const float coef = coef_s, t = *step;
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, rs2 = 1.f / sqrtf(bc2);
  for (long long i = (long long)blockIdx.x * ADAM_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * ADAM_THREADS) {
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * rs2 + eps);
  }
}

int pgtt_adam_clip(float* param, const float* grad, float* m, float* v, float* step, float* scratch, long long n, float lr, float beta1, float beta2,
                   float eps, float max_norm, float grad_scale, void* stream) {
  if (!param || !grad || !m || !v || !step || !scratch || n <= 0) return pfail(PGTT_ERR_ARG, "pgtt_adam_clip: null argument or n <= 0");
  pgtt_adam_sumsq_kernel<<<ADAM_BLOCKS, ADAM_THREADS, 0, (cudaStream_t)stream>>>(grad, n, scratch, step);
  pgtt_adam_update_kernel<<<ADAM_BLOCKS, ADAM_THREADS, 0, (cudaStream_t)stream>>>(param, grad, m, v, scratch, step, n, lr, beta1, beta2, eps, max_norm, grad_scale);
  PCUDA(cudaGetLastError());
  return PGTT_OK;
}

int pgtt_adam_scratch_floats(void) { return ADAM_BLOCKS; }

// generate_unroll: T x (act -> wrapped step -> record); see include/pgtt_b200.h
static int rollout_issue(pgtt_env* env, pgtt_policy* pol, int T, uint64_t seed, uint64_t step0, int deterministic, const pgtt_rollout_buffers* o,
                         size_t N, int nobs, int npriv, void* stream) {
  const size_t A = (size_t)pol->P.act_dim;
  // slot 0: the observation the unroll starts from; reward / discount / truncation slots are only written after a step
  if (int rc = pgtt_record(env, o->obs_state, o->obs_privileged, nullptr, nullptr, nullptr, stream)) return pfail(rc, pgtt_last_error());
  for (int t = 0; t < T; t++) {
    float* act = o->action + (size_t)t * N * A;
    if (int rc = pgtt_policy_act(pol, o->obs_state + (size_t)t * N * nobs, (int)N, seed, step0 + (uint64_t)t, deterministic, nullptr, act,
                                 o->raw_action ? o->raw_action + (size_t)t * N * A : nullptr, o->log_prob ? o->log_prob + (size_t)t * N : nullptr,
                                 nullptr, stream)) return rc;
    // the task kernel of the step writes the transition slot itself
    if (int rc = pgtt_step_record(env, act, 1, o->obs_state + (size_t)(t + 1) * N * nobs,
                                  o->obs_privileged ? o->obs_privileged + (size_t)(t + 1) * N * npriv : nullptr,
                                  o->reward ? o->reward + (size_t)t * N : nullptr, o->discount ? o->discount + (size_t)t * N : nullptr,
                                  o->truncation ? o->truncation + (size_t)t * N : nullptr, stream)) return pfail(rc, pgtt_last_error());
  }
  return PGTT_OK;
}

__global__ void pgtt_set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }

int pgtt_rollout(pgtt_env* env, pgtt_policy* pol, int T, uint64_t seed, uint64_t step0, int deterministic, const pgtt_rollout_buffers* o, void* stream) {
  if (!env || !pol || !o || T <= 0) return pfail(PGTT_ERR_ARG, "pgtt_rollout: null argument or T <= 0");
  if (!o->obs_state || !o->action) return pfail(PGTT_ERR_ARG, "pgtt_rollout: obs_state and action buffers are required");
  pgtt_buffers b;
  if (int rc = pgtt_get_buffers(env, &b)) return pfail(rc, pgtt_last_error());
  const size_t N = (size_t)b.num_envs;
  int nobs = 0, npriv = 0;
  pgtt_obs_dims(env, &nobs, &npriv);
  if (pol->P.obs_dim != nobs || pol->P.act_dim != PGTT_NU) return pfail(PGTT_ERR_ARG, "pgtt_rollout: policy must map the env's obs[\"state\"] (171 or 162) -> 2 x 12 logits");
  if (!pol->use_graph) return rollout_issue(env, pol, T, seed, step0, deterministic, o, N, nobs, npriv, stream);

  // CUDA-graph path: the unroll is captured once on an internal stream (the caller's stream may be the legacy default
  // stream, which cannot be captured) and replayed; the exploration-noise counter base lives on the device.
  cudaStream_t user = (cudaStream_t)stream;
  if (!pol->gstream) {
    PCUDA(cudaStreamCreateWithFlags(&pol->gstream, cudaStreamNonBlocking));
    PCUDA(cudaEventCreateWithFlags(&pol->ev_in, cudaEventDisableTiming));
    PCUDA(cudaEventCreateWithFlags(&pol->ev_out, cudaEventDisableTiming));
    PCUDA(cudaMalloc(&pol->step_dev, sizeof(unsigned long long)));
  }
  const bool same = pol->gexec && pol->gkey.env == env && pol->gkey.env_serial == pgtt_internal_serial(env) && pol->gkey.T == T && pol->gkey.deterministic == deterministic && pol->gkey.seed == seed &&
                    memcmp(&pol->gkey.o, o, sizeof(*o)) == 0;
  if (!same) {
    if (pol->gexec) { cudaGraphExecDestroy(pol->gexec); pol->gexec = nullptr; }
    // this env's constant table must be resident before the capture starts (the stream-ordered upload is not captured)
    if (int rc = pgtt_internal_make_resident(env, user)) return pfail(rc, pgtt_last_error());
    PCUDA(cudaStreamSynchronize(user));
    cudaGraph_t graph = nullptr;
    pol->step_base_arg = pol->step_dev;
    cudaError_t ce = cudaStreamBeginCapture(pol->gstream, cudaStreamCaptureModeThreadLocal);
    int rc = PGTT_OK;
    if (ce == cudaSuccess) {
      rc = rollout_issue(env, pol, T, seed, 0, deterministic, o, N, nobs, npriv, pol->gstream);
      ce = cudaStreamEndCapture(pol->gstream, &graph);
    }
    pol->step_base_arg = nullptr;
    if (ce == cudaSuccess && rc == PGTT_OK && graph) ce = cudaGraphInstantiate(&pol->gexec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (ce != cudaSuccess || rc != PGTT_OK || !pol->gexec) {   // no graph: issue directly (and stop trying)
      cudaGetLastError();
      pol->gexec = nullptr; pol->use_graph = 0;
      return rollout_issue(env, pol, T, seed, step0, deterministic, o, N, nobs, npriv, stream);
    }
    pol->gkey.env = env; pol->gkey.env_serial = pgtt_internal_serial(env); pol->gkey.T = T; pol->gkey.deterministic = deterministic; pol->gkey.seed = seed; pol->gkey.o = *o;
  } else {
    pgtt_internal_count_launches(env, (int64_t)pgtt_step_launches(env) * T);   // capture counted the first unroll's launches
    pol->launches += T;
  }
  PCUDA(cudaEventRecord(pol->ev_in, user));
  PCUDA(cudaStreamWaitEvent(pol->gstream, pol->ev_in, 0));
  // another handle of this device may have run in between: its constant table is replaced (stream-ordered) before the replay
  if (int rc = pgtt_internal_make_resident(env, pol->gstream)) return pfail(rc, pgtt_last_error());
  pgtt_set_u64_kernel<<<1, 1, 0, pol->gstream>>>(pol->step_dev, (unsigned long long)step0);
  PCUDA(cudaGraphLaunch(pol->gexec, pol->gstream));
  if (int rc = pgtt_internal_mark_launched(env, pol->gstream)) return pfail(rc, pgtt_last_error());
  PCUDA(cudaEventRecord(pol->ev_out, pol->gstream));
  PCUDA(cudaStreamWaitEvent(user, pol->ev_out, 0));
  return PGTT_OK;
}

}  // extern "C"
