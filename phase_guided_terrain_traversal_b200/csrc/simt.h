// simt.h - the few warp primitives the env kernels use.
//
// Product build (nvcc, sm_100a): thin wrappers over the CUDA intrinsics.
// Test build (-DPGTT_HOST_EMU, g++): the same kernel source is compiled for the host and each warp
// is run as 32 cooperatively scheduled fibers (tests/simt_emu/simt_emu.h). That build exists so the
// warp-cooperative logic can be checked against the CPU oracle without a GPU; it is test
// infrastructure, never loaded by the package.
#pragma once

#ifdef PGTT_HOST_EMU
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define DEV __device__ __forceinline__
#define DEV_NOINLINE __device__ __noinline__
#define FULL_MASK 0xffffffffu
DEV float shfl(float v, int src) { return __shfl_sync(FULL_MASK, v, src); }
DEV int shfl(int v, int src) { return __shfl_sync(FULL_MASK, v, src); }
DEV float shfl_xor(float v, int m) { return __shfl_xor_sync(FULL_MASK, v, m); }
DEV int shfl_xor(int v, int m) { return __shfl_xor_sync(FULL_MASK, v, m); }
DEV unsigned wballot(bool p) { return __ballot_sync(FULL_MASK, p); }
DEV bool any_lane(bool p) { return __any_sync(FULL_MASK, p); }
// Control decisions go through a vote so that the compiler knows the branch is warp-uniform; a
// branch on a butterfly-reduced float is uniform in fact but not provably, and every shuffle
// behind it is then emitted as an out-of-line WARPSYNC.COLLECTIVE trampoline.
DEV bool all_lanes(bool p) { return __all_sync(FULL_MASK, p); }
DEV void syncwarp() { __syncwarp(); }
// Stage barrier across the warps (= envs) of a CTA. Envs are independent, so this is not needed for
// correctness beyond the warp-level ordering it implies; it keeps the warps of a CTA streaming
// through the same stretch of code so instruction-cache lines are fetched once per CTA, not once per warp.
DEV void cta_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
// CTA-wide votes that double as lockstep barriers (quad kernels: every warp of the CTA iterates the same number of times)
DEV bool cta_any(bool p) { return __syncthreads_or(p) != 0; }
DEV bool cta_all(bool p) { return __syncthreads_and(p) != 0; }
DEV int warp_index() { return __shfl_sync(FULL_MASK, (int)(threadIdx.x >> 5), 0); }
DEV float ldg(const float* p) { return __ldg(p); }
DEV float4 ldg4(const float4* p) { return __ldg(p); }
DEV int popc(unsigned x) { return __popc(x); }
DEV int ffs_(unsigned x) { return __ffs((int)x); }
DEV float rsqrt_(float x) { return rsqrtf(x); }
// sin and cos together: Cody-Waite reduction by pi/2 (three-constant split, exact products for |a| < 1e5 - the arguments
// here are joint angles, yaws and half rotation angles) and the Cephes single-precision kernels on [-pi/4, pi/4]
// (~1 ulp). ~25 instructions inline; the library sincosf costs ~110 per call plus a 100-instruction slow path per site.
DEV void sincos_(float a, float* s, float* c) {
  const float k = rintf(a * 0.636619772367581f);
  float r = fmaf(k, -1.5703125f, a);
  r = fmaf(k, -4.837512969970703125e-4f, r);
  r = fmaf(k, -7.54978995489188e-8f, r);
  const float r2 = r * r;
  float sn = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
  float cs = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2 * r2, fmaf(-0.5f, r2, 1.0f));
  const int q = (int)k;
  if (q & 1) { const float t = sn; sn = cs; cs = -t; }
  if (q & 2) { sn = -sn; cs = -cs; }
  *s = sn; *c = cs;
}
// approximate division (MUFU.RCP + multiply, 2 ulp) for the solver's scalar control quantities - Newton-step lengths of the line search,
// convergence measures: they steer an iteration whose result is compared at 1e-3, and the IEEE sequence is ~8 instructions on a
// latency-bound dependent chain (the host-emulated build divides exactly)
DEV float fdiv_(float a, float b) { return __fdividef(a, b); }
// un-contracted multiply-add: random draws must not depend on whether the compiler forms an FMA
DEV float mul_add_nofma(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
#endif

DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor(v, o);
  return v;
}
DEV int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor(v, o);
  return v;
}
DEV int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const int u = shfl_xor(v, o); v = u < v ? u : v; }
  return v;
}
DEV float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, shfl_xor(v, o));
  return v;
}
DEV float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, shfl_xor(v, o));
  return v;
}
