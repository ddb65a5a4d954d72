// shems_device.cuh — the shems_LU1 transition as straight-line device code.
//
// Reference: RL-SHEMS/RL_environments/envs/shems_LU1.jl — action(env, a) :283-316,
// action(env, track) :318-340, step! :343-485, next_state! :264-281.
//
// The Julia source is dynamically typed: which intermediates are Float32 and which are Float64
// depends on the branch (zeros(11) are Float64, b.rate_max is Float64, `pv_ = 0` is an Int ...).
// Here every flow branch ("leaf") only ASSIGNS values into a canonical set of registers — doubles
// holding exactly-representable Float32 or genuine Float64 values — and the common tail applies
// the rounding each leaf implies; the two leaves whose sums stay Float32 in Julia (A2a, B1a) are
// flagged.  Built with -fmad=false: Julia never contracts a*b+c, so neither may this file.
#pragma once
#include "common.h"

struct StepIn {
  float Soc_b, Soc_ev, c_ev, d_e, g_e, p_buy;
};
struct StepOut {
  float Soc_b, Soc_ev;  // endogenous part of s' (before next_state! overrides Soc_ev on arrival)
  double reward;        // env.reward (Float64)
};
struct StepTrace {  // the remaining `results` columns (:476-478); all values exact in double
  double EV_target, EV, profit, discomfort, penalty, PV_DE, B_DE, GR_DE, PV_B, PV_GR, PV_EV, B_EV, GR_EV, EX_EV, B, B_target;
};

__device__ __forceinline__ float jl_minf(float x, float y) { return (y < x) ? y : x; }     // Base.min(x, y), no NaN
__device__ __forceinline__ double jl_mind(double x, double y) { return (y < x) ? y : x; }
// Base.clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))
__device__ __forceinline__ double jl_clampd(double x, double lo, double hi) { return (x > hi) ? hi : ((x < lo) ? lo : x); }

// action(env, a::ShemsAction) :283-316 -> Float32.([B, EV])
__device__ __forceinline__ void shems_action_drl(const DevParams& P, float Soc_b, float Soc_ev, float c_ev, float d_e,
                                                 float g_e, float Bt, float EVt, float& B, float& EV) {
  const float perc = (Soc_b - P.smin) / P.span;                                   // :288
  EV = (c_ev > -1.0f && Soc_ev < EVt) ? jl_minf(P.evR, (EVt - Soc_ev) * P.C) : 0.0f;  // :292-297
  const float pv = (g_e - d_e) - EV;                                              // :301
  if (pv > 0.0f && perc < Bt) {                                                   // :304
    const float Btv = Bt * P.span + P.smin;                                       // :306 (two roundings, fmad off)
    const double hi = jl_mind(P.R, (double)((Btv - Soc_b) + P.loss));             // :307
    B = (float)jl_clampd((double)pv, 0.0, hi);
  } else if (Soc_b > 1e-3f) {                                                     // :309
    B = (float)(-jl_mind(P.R, (double)(P.one_m_l * Soc_b)));                      // :310
  } else {
    B = 0.0f;
  }
}

// action(env, track) rule-based controller :318-340 -> Float32.([B, EV])
__device__ __forceinline__ void shems_action_rule(const DevParams& P, float Soc_b, float Soc_ev, float d_e, float g_e,
                                                  float& B, float& EV) {
  EV = jl_minf(P.evR, (1.0f - Soc_ev) * P.C);                                     // :323
  const float pv = (g_e - d_e) - EV;                                              // :327
  if (pv > 0.0f && (double)Soc_b < P.smax95) {                                    // :330 (0.95 is Float64)
    const double hi = jl_mind(P.R, (double)((P.smax - Soc_b) + P.loss));          // :331
    B = (float)jl_clampd((double)pv, 0.0, hi);
  } else if (Soc_b > 1e-3f) {                                                     // :333
    B = (float)(-jl_mind(P.R, (double)(P.one_m_l * Soc_b)));                      // :334
  } else {
    B = 0.0f;
  }
}

// step! :343-485 given the feasible (B, EV) and the recorded targets (Bt, EVt).
// TRACK_NEG <=> track < 0 (no penalty in the reward, :466-471).
template <bool WANT_TRACE>
__device__ __forceinline__ StepOut shems_flows(const DevParams& P, const StepIn& s, float B, float EV, float EVt,
                                               bool track_neg, StepTrace* tr) {
  const float eta = P.b_eta;
  const double eta_d = P.eta_d;
  // :362-364  battery discharge amount (Float64)
  double BD = 0.0;
  if (B < -0.01f) {  // F32 < -0.01 (Float64 literal) <=> F32 < -0.01f0, see tests/test_oracle_env.py
    BD = jl_clampd((double)(-B), 0.001, jl_mind(P.R, (double)(P.one_m_l_e * s.Soc_b)));
  }
  // canonical registers (values exact in double)
  float PV_DE, PV_EV = 0.0f, pv = 0.0f;  // always Float32 (or a zero)
  double B_DE = 0.0, B_EV = 0.0, GR_DE = 0.0, GR_EV = 0.0;
  bool leafA2a = false, leafB1a = false;
  float f32sum = 0.0f;  // the Float32 partial sum of the flagged leaf
  const float ge = s.g_e * P.pv_eta;
  if (ge > s.d_e) {  // :368 PV covers the demand
    PV_DE = s.d_e;
    pv = ge - PV_DE;
    if (pv > EV) {   // :371
      PV_EV = EV;
      pv = pv - PV_EV;
    } else {         // :374
      PV_EV = pv;
      pv = 0.0f;
      const float rem = EV - PV_EV;
      const float need = rem / eta;
      if (BD > (double)need) {  // :377
        B_EV = (double)rem;
        BD = BD - (double)need;
        leafA2a = true;
        f32sum = PV_EV + rem;   // PV_EV + B_EV is a Float32 add in Julia (:435)
      } else {                  // :380
        B_EV = BD * eta_d;
        BD = 0.0;
        GR_EV = (double)rem - B_EV;
      }
    }
  } else {           // :388 PV short of the demand
    PV_DE = ge;
    const float d = s.d_e - PV_DE;
    const float dn = d / eta;
    if (BD > (double)dn) {      // :392
      B_DE = (double)d;
      BD = BD - (double)dn;
      const float en = EV / eta;
      if (BD > (double)en) {    // :395
        B_EV = (double)EV;
        BD = BD - (double)en;
        leafB1a = true;
        f32sum = d + EV;        // B_DE + B_EV is a Float32 add in Julia (:432)
      } else {                  // :398
        B_EV = BD * eta_d;
        BD = 0.0;
        GR_EV = (double)EV - B_EV;
      }
    } else {                    // :403
      B_DE = BD * eta_d;
      BD = 0.0;
      GR_DE = (double)d - B_DE;
      GR_EV = (double)EV;
    }
  }
  // :412-422 battery charging
  double PV_B = 0.0;        // value of PV_B
  double X = (double)s.Soc_b;  // Soc_b + PV_B + GR_B with Julia's rounding
  double pv_d = (double)pv;  // pv_ / PV_GR
  if (B > 0.01f) {  // F32 > 0.01 (Float64) <=> F32 > 0.01f0
    const double BC = jl_clampd((double)B, 0.001, jl_mind(P.R, (double)(P.smax - s.Soc_b)));
    const double thr = BC / eta_d;
    if (pv_d > thr) {  // :414
      PV_B = BC;
      pv_d = pv_d - thr;
      X = (double)s.Soc_b + BC;
    } else {           // :417 PV_B = pv_ * b.eta stays Float32 (pv_ is Float32, or the Int 0)
      const float pvb = pv * eta;
      PV_B = (double)pvb;
      pv_d = 0.0;
      X = (double)(s.Soc_b + pvb);
    }
  }
  // :432  (1 - loss) * (Soc_b + PV_B + GR_B - (B_DE + B_EV + B_GR) / eta)
  const double Y = leafB1a ? (double)(f32sum / eta) : (B_DE + B_EV) / eta_d;
  StepOut o;
  o.Soc_b = (float)(P.one_m_l_d * (X - Y));
  // :435  Soc_ev + (PV_EV + B_EV + GR_EV) / (ev.soc_max - ev.soc_min)
  const double T = leafA2a ? (double)f32sum : (((double)PV_EV + B_EV) + GR_EV);
  float Soc_ev_new = (float)((double)s.Soc_ev + T / P.C_d);
  // :438-449
  float disc = 0.0f, pen = 0.0f, EX_EV = 0.0f;
  if (s.c_ev == 0.0f && Soc_ev_new < 1.0f) {
    const float short_ = 1.0f - Soc_ev_new;
    disc = short_ * 100.0f;
    EX_EV = short_ * P.C;
    Soc_ev_new = 1.0f;
  } else if (s.c_ev < 0.0f && EVt < 0.99f) {  // F32 < 0.99 (Float64) <=> F32 < 0.99f0
    pen = (1.0f - EVt) * P.pw;
  }
  o.Soc_ev = Soc_ev_new;
  // :464  profit = (sell * p_buy * (PV_GR + B_GR)) - (p_buy * (GR_DE + GR_B + GR_EV + EX_EV))   [Float64]
  const double pb = (double)s.p_buy;
  const double profit = ((P.sell * pb) * pv_d) - (pb * ((GR_DE + GR_EV) + (double)EX_EV));
  const double dd = (double)disc;
  const double dpow = (P.pot == 2.0) ? dd * dd : pow(dd, P.pot);  // Float32^Float64 promotes; x*x is exact for F32 x
  const double base = profit - P.dw * dpow;
  if (track_neg) pen = 0.0f;  // :466-468
  o.reward = track_neg ? base : (base - (double)pen);
  if (WANT_TRACE) {
    tr->EV = (double)EV; tr->profit = profit; tr->discomfort = dd; tr->penalty = (double)pen;
    tr->PV_DE = (double)PV_DE; tr->B_DE = B_DE; tr->GR_DE = GR_DE; tr->PV_B = PV_B; tr->PV_GR = pv_d;
    tr->PV_EV = (double)PV_EV; tr->B_EV = B_EV; tr->GR_EV = GR_EV; tr->EX_EV = (double)EX_EV; tr->B = (double)B;
  }
  return o;
}
