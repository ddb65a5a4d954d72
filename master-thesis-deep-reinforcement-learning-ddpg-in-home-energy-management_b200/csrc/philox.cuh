// philox.cuh — Philox4x32-10 counter-based streams (this repo's RNG spec; the reference's
// Julia MersenneTwister streams cannot be reproduced).  Restated bit-exactly in
// oracle/shems_oracle.c (oracle_philox).  key = seed, counter = (id_lo, id_hi, ctr, stream).
#pragma once
#include <stdint.h>

enum : uint32_t { STREAM_RESET = 0x5245u, STREAM_ACTION = 0x4143u, STREAM_NOISE = 0x4e4fu, STREAM_SAMPLE = 0x534du, STREAM_INIT = 0x494eu };

__host__ __device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t id, uint32_t ctr, uint32_t stream, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32), c2 = ctr, c3 = stream;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 53-bit uniform in [0,1) from two words (stands in for Julia's Float64 rand())
__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return (double)((((uint64_t)(a >> 5)) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}
