// act_epilogue.cuh — clamp(actor(s) + noise, -1, 1) and scale_action (RL-SHEMS/algorithms/DDPG.jl:57-61, :172-184) for one state:
// shared by ddpg_act_epilogue_kernel (csrc/ddpg.cu) and the cluster-fused act kernel (csrc/ddpg_fused.cu), so both draw the same
// Philox/Box-Muller stream and round alike.
#pragma once
#include "philox.cuh"

struct ActOut { float a0, a1, s0, s1; float noise_mean; };  // noise_mean = mean(noise) of the two components (act returns it, DDPG.jl:172)
// noise: (nz0, nz1) given when have_noise, else σ·N(0,1) by Box-Muller on Philox(seed, env id, step) when sigma > 0, else none
__device__ __forceinline__ ActOut act_gauss_epilogue(float y0, float y1, bool have_noise, float nz0, float nz1, float sigma, unsigned long long seed,
                                                     long long step, long long env_id, float lo0, float lo1, float hi0, float hi1) {
  if (!have_noise) {
    nz0 = 0.0f; nz1 = 0.0f;
    if (sigma > 0.0f) {
      uint32_t w[4];
      philox4x32_10(seed, (uint64_t)env_id, (uint32_t)step, STREAM_NOISE, w);
      const double u1 = 1.0 - u53(w[0], w[1]), u2 = u53(w[2], w[3]);  // u1 in (0,1]
      const double rad = sqrt(-2.0 * log(u1));
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      nz0 = (float)((double)sigma * (rad * cs));  // Float32.(rand(Normal(μ=0, σ), 2))  (DDPG.jl:57-61)
      nz1 = (float)((double)sigma * (rad * sn));
    }
  }
  ActOut o;
  o.noise_mean = __fmul_rn(__fadd_rn(nz0, nz1), 0.5f);   // mean(Float32[nz0, nz1])
  float a0 = __fadd_rn(y0, nz0), a1 = __fadd_rn(y1, nz1);
  a0 = a0 > 1.0f ? 1.0f : (a0 < -1.0f ? -1.0f : a0);
  a1 = a1 > 1.0f ? 1.0f : (a1 < -1.0f ? -1.0f : a1);
  o.a0 = a0; o.a1 = a1;
  // Float32.(LO .+ (a .+ 1.0) .* 0.5 .* (HI .- LO)) in Float64
  const double sp0 = (double)__fsub_rn(hi0, lo0), sp1 = (double)__fsub_rn(hi1, lo1);
  o.s0 = (float)__dadd_rn((double)lo0, __dmul_rn(__dmul_rn(__dadd_rn((double)a0, 1.0), 0.5), sp0));
  o.s1 = (float)__dadd_rn((double)lo1, __dmul_rn(__dmul_rn(__dadd_rn((double)a1, 1.0), 0.5), sp1));
  return o;
}
