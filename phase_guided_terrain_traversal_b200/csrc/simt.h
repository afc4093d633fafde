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
DEV void syncwarp() { __syncwarp(); }
DEV float ldg(const float* p) { return __ldg(p); }
DEV int popc(unsigned x) { return __popc(x); }
DEV float rsqrt_(float x) { return rsqrtf(x); }
DEV void sincos_(float a, float* s, float* c) { sincosf(a, s, c); }
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
