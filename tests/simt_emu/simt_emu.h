// simt_emu.h - host emulation of one CUDA warp as 32 cooperatively scheduled fibers.
//
// TEST INFRASTRUCTURE ONLY. It lets tests/ compile the env kernel source (csrc/pgtt_device.cuh)
// with g++ and run it lane-for-lane against the CPU oracle on a machine without a GPU. The
// product (libpgtt_b200.so) is built by nvcc from the same source and never includes this file.
//
// Model: every warp collective (shuffle, ballot, syncwarp) is a barrier across the 32 fibers; the
// kernel code only issues full-mask collectives from warp-uniform control flow, which is asserted.
#pragma once
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#define DEV static inline
#define DEV_NOINLINE static
#define FULL_MASK 0xffffffffu
#define __restrict__
struct alignas(16) float4 { float x, y, z, w; };

struct EmuWarp {
  ucontext_t main_ctx;
  ucontext_t ctx[32];
  char* stacks[32];
  int finished[32];
  int nfinished;
  uint32_t xbuf[2][32];
  int parity[32];
  int arrived;
  unsigned gen;
  void (*fn)(void*, int);
  void* arg;
  long ncollectives;
};

extern thread_local EmuWarp* emu_warp;
extern thread_local int emu_lane;

static inline void emu_barrier() {
  EmuWarp* w = emu_warp;
  int lane = emu_lane;
  assert(w->nfinished == 0 && "a lane returned while others still issue collectives");
  unsigned mygen = w->gen;
  if (++w->arrived == 32) { w->arrived = 0; w->gen++; w->ncollectives++; return; }
  while (w->gen == mygen) { swapcontext(&w->ctx[lane], &w->main_ctx); emu_lane = lane; }
}

static inline uint32_t emu_exchange(uint32_t bits, int src) {
  EmuWarp* w = emu_warp;
  int lane = emu_lane;
  int p = w->parity[lane];
  w->parity[lane] ^= 1;
  w->xbuf[p][lane] = bits;
  emu_barrier();
  return emu_warp->xbuf[p][src & 31];
}

DEV float shfl(float v, int src) { uint32_t b; memcpy(&b, &v, 4); b = emu_exchange(b, src); float r; memcpy(&r, &b, 4); return r; }
DEV int shfl(int v, int src) { return (int)emu_exchange((uint32_t)v, src); }
DEV float shfl_xor(float v, int m) { return shfl(v, emu_lane ^ m); }
DEV int shfl_xor(int v, int m) { return shfl(v, emu_lane ^ m); }
DEV unsigned wballot(bool pr) {
  EmuWarp* w = emu_warp;
  int lane = emu_lane;
  int p = w->parity[lane];
  w->parity[lane] ^= 1;
  w->xbuf[p][lane] = pr ? 1u : 0u;
  emu_barrier();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (emu_warp->xbuf[p][i] & 1u) << i;
  return r;
}
DEV bool any_lane(bool p) { return wballot(p) != 0; }
DEV bool all_lanes(bool p) { return wballot(p) == FULL_MASK; }
DEV void syncwarp() { emu_barrier(); }
DEV void cta_bar(int) { emu_barrier(); }
DEV bool cta_any(bool p) { return any_lane(p); }   // the emulated CTA is one warp
DEV bool cta_all(bool p) { return all_lanes(p); }
DEV float ldg(const float* p) { return *p; }
DEV float4 ldg4(const float4* p) { return *p; }
DEV int popc(unsigned x) { return __builtin_popcount(x); }
DEV int ffs_(unsigned x) { return __builtin_ffs((int)x); }
DEV float rsqrt_(float x) { return 1.0f / sqrtf(x); }
DEV float fdiv_(float a, float b) { return a / b; }
DEV void sincos_(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
DEV float mul_add_nofma(float a, float b, float c) { return a * b + c; }  // built with -ffp-contract=off
DEV float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
DEV int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
DEV float __uint_as_float(uint32_t i) { float f; memcpy(&f, &i, 4); return f; }

#ifdef EMU_IMPL
thread_local EmuWarp* emu_warp = nullptr;
thread_local int emu_lane = 0;

static void emu_trampoline(int lane) {
  EmuWarp* w = emu_warp;
  emu_lane = lane;
  w->fn(w->arg, lane);
  w->finished[lane] = 1;
  w->nfinished++;
  swapcontext(&w->ctx[lane], &w->main_ctx);
}

// Run fn(arg, lane) for lanes 0..31 as one warp. Returns the number of collectives executed.
long emu_run_warp(void (*fn)(void*, int), void* arg) {
  static thread_local EmuWarp* w = nullptr;
  const size_t STK = 1 << 18;
  if (!w) {
    w = (EmuWarp*)calloc(1, sizeof(EmuWarp));
    for (int l = 0; l < 32; l++) w->stacks[l] = (char*)malloc(STK);
  }
  char* stacks[32];
  memcpy(stacks, w->stacks, sizeof(stacks));
  memset(w, 0, sizeof(EmuWarp));
  memcpy(w->stacks, stacks, sizeof(stacks));
  w->fn = fn; w->arg = arg;
  emu_warp = w;
  for (int l = 0; l < 32; l++) {
    getcontext(&w->ctx[l]);
    w->ctx[l].uc_stack.ss_sp = w->stacks[l];
    w->ctx[l].uc_stack.ss_size = STK;
    w->ctx[l].uc_link = &w->main_ctx;
    makecontext(&w->ctx[l], (void (*)())emu_trampoline, 1, l);
  }
  while (w->nfinished < 32) {
    for (int l = 0; l < 32; l++)
      if (!w->finished[l]) { emu_lane = l; swapcontext(&w->main_ctx, &w->ctx[l]); }
  }
  return w->ncollectives;
}
#else
long emu_run_warp(void (*fn)(void*, int), void* arg);
#endif
