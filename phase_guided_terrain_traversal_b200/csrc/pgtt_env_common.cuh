// pgtt_env_common.cuh - jax.random (threefry2x32), go2/gait.py:27-49 and go2/utility.py:4-8 as device functions,
// shared by the task kernel (pgtt_task.cuh), the reset / randomise kernels (pgtt_env.cuh) and the quad kernels.
#pragma once
#include "pgtt_physics.cuh"

// ----------------------------------------------------------------------------------------------
// jax.random
// ----------------------------------------------------------------------------------------------
struct Key { uint32_t a, b; };

#ifdef PGTT_HOST_EMU
DEV uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#else
DEV uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }   // one SHF
#endif
// inlined: ~105 instructions per site, and independent draws of one lane interleave (the physics kernels do not draw)
DEV Key threefry(Key k, uint32_t x0, uint32_t x1) {
  const uint32_t ks0 = k.a, ks1 = k.b, ks2 = k.a ^ k.b ^ 0x1BD11BDAu;
  x0 += ks0; x1 += ks1;
#define TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)
  x0 += ks1; x1 += ks2 + 1u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)
  x0 += ks2; x1 += ks0 + 2u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)
  x0 += ks0; x1 += ks1 + 3u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)
  x0 += ks1; x1 += ks2 + 4u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef TF_R
  Key o; o.a = x0; o.b = x1;
  return o;
}
// split(key, num)[i]
DEV Key rng_split(Key k, int num, int i) {
  if (GC.rng_partitionable) return threefry(k, 0u, (uint32_t)i);
  Key o;
  {
    const int flat = 2 * i, pair = flat % num, which = flat / num;
    const Key t = threefry(k, (uint32_t)pair, (uint32_t)(pair + num));
    o.a = which ? t.b : t.a;
  }
  {
    const int flat = 2 * i + 1, pair = flat % num, which = flat / num;
    const Key t = threefry(k, (uint32_t)pair, (uint32_t)(pair + num));
    o.b = which ? t.b : t.a;
  }
  return o;
}
// random_bits(key, 32, (n,))[i]
DEV uint32_t rng_bits(Key k, int n, int i) {
  if (GC.rng_partitionable) { const Key t = threefry(k, 0u, (uint32_t)i); return t.a ^ t.b; }
  const int half = (n + 1) / 2, pair = i % half, which = i / half;
  uint32_t x1 = (uint32_t)(pair + half);
  if ((n & 1) && pair + half >= n) x1 = 0u;
  const Key t = threefry(k, (uint32_t)pair, x1);
  return which ? t.b : t.a;
}
DEV float rng_unit(Key k, int n, int i) { return __uint_as_float((rng_bits(k, n, i) >> 9) | 0x3F800000u) - 1.0f; }
DEV float rng_uniform(Key k, int n, int i, float lo, float hi) { return fmaxf(lo, mul_add_nofma(rng_unit(k, n, i), hi - lo, lo)); }
DEV int rng_randint(Key k, int lo, int hi) {  // shape (1,)
  const Key k1 = rng_split(k, 2, 0), k2 = rng_split(k, 2, 1);
  const uint32_t hb = rng_bits(k1, 1, 0), lb = rng_bits(k2, 1, 0);
  uint32_t span = (uint32_t)(hi - lo);
  if (hi <= lo) span = 1u;
  const uint32_t mult = ((65536u % span) * (65536u % span)) % span;
  return lo + (int)(((hb % span) * mult + (lb % span)) % span);
}

// ----------------------------------------------------------------------------------------------
// go2/gait.py:27-49 and go2/utility.py:4-8
// ----------------------------------------------------------------------------------------------
DEV float gait_get_z(float phi, float h_max, float stance) {
  const float T_swing = 2.f * PGTT_PI * (1.f - 0.5f) / 2.f, T_peak = 2.f * PGTT_PI * (1.f + 0.5f) / 2.f, T_stance = 2.f * PGTT_PI * 0.5f;
  if (phi <= T_stance) return stance;
  float p0, p1, t;
  if (phi <= T_peak) { p0 = stance; p1 = h_max; t = (phi - T_stance) / T_swing; }
  else { p0 = h_max; p1 = stance; t = (phi - T_peak) / T_swing; }
  const float t2 = t * t, t3 = t2 * t;
  return (2.f * t3 - 3.f * t2 + 1.f) * p0 + (-2.f * t3 + 3.f * t2) * p1;
}
DEV float quat_to_yaw(const float* q) {
  return atan2f(2.f * (q[0] * q[3] + q[1] * q[2]), 1.f - 2.f * (q[2] * q[2] + q[3] * q[3]));
}

