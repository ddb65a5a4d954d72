/*
 * shems_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A literal restatement of the reference environment
 *   /root/reference/RL-SHEMS/RL_environments/envs/shems_LU1.jl
 * (constants :40-59,:67-99; reset :206-262; next_state! :264-281; action :283-316 and
 * :318-340; step! :343-485) written so that it can be read side by side with the Julia.
 *
 * PARITY UNPINNED: the reference ships no golden vectors for this path (every file under
 * RL-SHEMS/out/ is a git-LFS pointer) and Julia is not installed, so this oracle cannot be
 * checked against reference output.  It is pinned only by (i) hand-derived known-answer
 * vectors (tests/golden/kat_appendix_b.json), (ii) invariants that
 * follow from the source (energy balances) and (iii) line-by-line review.
 *
 * Julia is dynamically typed and the reference mixes Int, Float32 and Float64 (e.g.
 * `zeros(11)` makes Float64 zeros, `b.rate_max` is Float64, `pv_ = 0` is an Int), so the
 * type — and therefore the rounding — of an intermediate depends on the branch taken.
 * Instead of hand-deriving a static type per path, every value here is a tagged `jl`
 * value and every operation applies Julia's promotion rule (Int ⊕ Float32 → Float32,
 * anything ⊕ Float64 → Float64) and rounds in the promoted type, exactly as Julia does.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or call this file.  Build: see oracle/Makefile (-ffp-contract=off is required).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/shems_b200.h"
#include "oracle.h"
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1: the CPU baseline sets its thread count explicitly */
int oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* --------------------------------------------------------------- jl values */
typedef enum { K_INT = 0, K_F32 = 1, K_F64 = 2 } jkind;
typedef struct { jkind k; double v; } jl; /* v holds the value exactly (ints are small) */

static inline jl J_I(long long x) { jl r = {K_INT, (double)x}; return r; }
static inline jl J_F(float x) { jl r = {K_F32, (double)x}; return r; }
static inline jl J_D(double x) { jl r = {K_F64, x}; return r; }
static inline jkind prom(jl a, jl b) { return a.k > b.k ? a.k : b.k; }
static inline jl conv(jl a, jkind k) { /* convert(T, a) */
  jl r; r.k = k;
  if (k == K_F32) r.v = (double)(float)a.v; else r.v = a.v;
  return r;
}
static inline float to_f32(jl a) { return (float)a.v; } /* Float32(a): round-to-nearest-even */
static inline double to_f64(jl a) { return a.v; }

#define BINOP(name, OP)                                                            \
  static inline jl name(jl a, jl b) {                                              \
    jkind k = prom(a, b); jl r; r.k = k;                                           \
    if (k == K_F32) { const float x = (float)a.v, y = (float)b.v; const float z = x OP y; r.v = (double)z; }     \
    else if (k == K_F64) { r.v = a.v OP b.v; }                                                                    \
    else { r.v = (double)((long long)a.v OP (long long)b.v); }                      \
    return r;                                                                      \
  }
BINOP(jadd, +)
BINOP(jsub, -)
BINOP(jmul, *)
static inline jl jdiv(jl a, jl b) { /* Int/Int -> Float64 in Julia */
  jkind k = prom(a, b); jl r;
  if (k == K_INT) k = K_F64;
  r.k = k;
  if (k == K_F32) { const float x = (float)a.v, y = (float)b.v; const float z = x / y; r.v = (double)z; }
  else { r.v = a.v / b.v; }
  return r;
}
static inline jl jneg(jl a) { jl r = a; r.v = -a.v; return r; }
/* comparisons promote and are exact (every Int/Float32 here is exactly representable in Float64) */
static inline int jlt(jl a, jl b) { return a.v < b.v; }
static inline int jgt(jl a, jl b) { return a.v > b.v; }
static inline int jle(jl a, jl b) { return a.v <= b.v; }
static inline int jge(jl a, jl b) { return a.v >= b.v; }
static inline int jeq(jl a, jl b) { return a.v == b.v; }
/* Base.min for floats (after promotion): y if (y < x) | (signbit(y) > signbit(x)) else x; no NaNs here */
static inline jl jmin(jl a, jl b) {
  jkind k = prom(a, b);
  jl x = conv(a, k), y = conv(b, k);
  if (k == K_INT) return (y.v < x.v) ? y : x;
  return ((y.v < x.v) || (signbit(y.v) > signbit(x.v))) ? y : x;
}
/* Base.clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x)) in promote_type(X, L, H) */
static inline jl jclamp(jl x, jl lo, jl hi) {
  jkind k = prom(x, lo); if (hi.k > k) k = hi.k;
  if (jgt(x, hi)) return conv(hi, k);
  if (jlt(x, lo)) return conv(lo, k);
  return conv(x, k);
}
/* x ^ y with y::Float64 -> Float64 pow (Int ^ Float64 and Float32 ^ Float64 both promote) */
static inline jl jpow(jl x, jl y) { return J_D(pow(x.v, y.v)); }

/* --------------------------------------------------------------- constants */
/* capacities dict, shems_LU1.jl:47-59: (EV capacity F32, battery capacity F32 product, rate F64) */
int oracle_params_for_charger(int charger_id, ShemsParams* p) {
  float ev_cap, b_cap; double rate;
  switch (charger_id) {
    case 1: ev_cap = 48.250f; b_cap = 7.5f * 0.9f; rate = 3.3; break;
    case 2: ev_cap = 36.271f; b_cap = 10.f * 0.9f; rate = 3.3; break;
    case 3: ev_cap = 45.508f; b_cap = 10.f * 0.9f; rate = 3.3; break;
    case 4: ev_cap = 78.993f; b_cap = 11.f * 0.9f; rate = 4.6; break;
    case 5: ev_cap = 37.207f; b_cap = 10.f * 0.9f; rate = 4.6; break;
    case 6: ev_cap = 35.816f; b_cap = 15.f * 0.9f; rate = 4.6; break;
    case 7: ev_cap = 36.521f; b_cap = 12.f * 0.9f; rate = 3.3; break;
    case 8: ev_cap = 45.728f; b_cap = 10.f * 0.9f; rate = 3.3; break;
    case 9: ev_cap = 21.935f; b_cap = 7.5f * 0.9f; rate = 3.3; break;
    case 98: ev_cap = 35.816f; b_cap = 7.5f * 0.9f; rate = 3.3; break;
    case 97: ev_cap = 78.993f; b_cap = 11.f * 0.9f; rate = 4.6; break;
    default: return SHEMS_ERR_KEY; /* KeyError at :95 */
  }
  { volatile float a = b_cap; b_cap = a; } /* Float32 product, rounded once */
  p->pv_eta = 1.0f;                       /* :92 */
  p->b_eta = 0.95f; p->b_soc_min = 0.f; p->b_soc_max = b_cap; p->b_rate_max = rate; p->b_loss = 0.00003f; /* :95 */
  p->ev_soc_min = 0.f; p->ev_soc_max = ev_cap; p->ev_rate_max = 11.f;                                     /* :97 */
  p->penalty_weight = 0.1f;               /* :43 */
  p->sell_discount = (double)0.2f;        /* Market(0.2f0, …) with Float64 fields :85-89, :99 */
  p->discomfort_weight_ev = (double)0.01f; /* :40 */
  p->disc_pot = (double)2.0f;             /* :41 */
  p->penalty_weight_f64 = 0.0; p->penalty_in_f64 = 0; p->reward_form = 0;
  return 0;
}

/* module-level constants of the sibling environment files:
 *   shems_LU7.jl: ev_capacities :42-55; b = Battery(0.95f0, 0f0, 10f0, 4.6f0, 0.00003f0) :91 (rate_max::Float64);
 *                 m = Market(0.3f0, DISCOMFORT_WEIGHT_EV = 1) :94, :25; penalty_weight = 0.1 (Float64) :35; linear discomfort :465-468
 *   shems_LU1_input0607.jl: capacities :57-68 (no 97); DISC_POT = 1f0 :49; DISCOMFORT_WEIGHT_EV by JOB digit :38-47 (0.1f0 here);
 *                 penalty_weight = 0.1 (Float64) :52; (discomfort * w)^pot :481-484 */
int oracle_params_for_env(int variant, int charger_id, ShemsParams* p) {
  if (variant == SHEMS_ENV_LU1) return oracle_params_for_charger(charger_id, p);
  if (variant == SHEMS_ENV_LU1_INPUT0607) {
    if (charger_id == 97) return SHEMS_ERR_KEY;
    int st = oracle_params_for_charger(charger_id, p);
    if (st) return st;
    p->discomfort_weight_ev = (double)0.1f; p->disc_pot = (double)1.0f;
    p->penalty_weight_f64 = 0.1; p->penalty_in_f64 = 1; p->reward_form = 1;
    return 0;
  }
  if (variant != SHEMS_ENV_LU7) return SHEMS_ERR_INVALID;
  float ev_cap;
  switch (charger_id) {
    case 1: ev_cap = 48.250f; break; case 2: ev_cap = 36.271f; break; case 3: ev_cap = 45.508f; break; case 4: ev_cap = 78.993f; break;
    case 5: ev_cap = 37.207f; break; case 6: ev_cap = 35.816f; break; case 7: ev_cap = 36.521f; break; case 8: ev_cap = 45.728f; break;
    case 9: ev_cap = 21.935f; break; case 99: ev_cap = 35.816f; break; case 98: ev_cap = 35.816f; break;
    default: return SHEMS_ERR_KEY;
  }
  p->pv_eta = 1.0f;
  p->b_eta = 0.95f; p->b_soc_min = 0.f; p->b_soc_max = 10.f; p->b_rate_max = (double)4.6f; p->b_loss = 0.00003f;
  p->ev_soc_min = 0.f; p->ev_soc_max = ev_cap; p->ev_rate_max = 11.f;
  p->penalty_weight = 0.1f;
  p->sell_discount = (double)0.3f; p->discomfort_weight_ev = 1.0; p->disc_pot = 1.0;
  p->penalty_weight_f64 = 0.1; p->penalty_in_f64 = 1; p->reward_form = 0;
  return 0;
}

/* column c of row `row` (1-based), series is [8][nrows] Float32 */
#define COL(c, row) (series[(size_t)(c) * (size_t)nrows + (size_t)((row) - 1)])

/* ------------------------------------------------- action(env, a::ShemsAction) :283-316 */
void oracle_action_drl(const ShemsParams* P, const float* s, float B_target_f, float EV_target_f, float* out) {
  jl Soc_b = J_F(s[0]), Soc_ev = J_F(s[1]), c_ev = J_F(s[2]), d_e = J_F(s[3]), g_e = J_F(s[4]);
  jl B_target = J_F(B_target_f), EV_target = J_F(EV_target_f);
  jl b_soc_min = J_F(P->b_soc_min), b_soc_max = J_F(P->b_soc_max), b_rate_max = J_D(P->b_rate_max), b_loss = J_F(P->b_loss);
  jl ev_soc_min = J_F(P->ev_soc_min), ev_soc_max = J_F(P->ev_soc_max), ev_rate_max = J_F(P->ev_rate_max);
  jl B = J_D(0.0), EV = J_D(0.0); /* B, EV = zeros(2) :286 */

  jl Soc_b_perc = jdiv(jsub(Soc_b, b_soc_min), jsub(b_soc_max, b_soc_min)); /* :288 */
  if (jgt(c_ev, J_I(-1)) && jlt(Soc_ev, EV_target)) {                        /* :292 */
    EV = jmin(ev_rate_max, jmul(jsub(EV_target, Soc_ev), jsub(ev_soc_max, ev_soc_min))); /* :294 */
  } else {
    EV = J_I(0); /* :296 */
  }
  jl pv_ = jsub(jsub(g_e, d_e), EV); /* :301 */
  if (jgt(pv_, J_I(0)) && jlt(Soc_b_perc, B_target)) { /* :304 */
    jl B_target_value = jadd(jmul(B_target, jsub(b_soc_max, b_soc_min)), b_soc_min); /* :306 */
    B = jclamp(pv_, J_I(0), jmin(b_rate_max, jadd(jsub(B_target_value, Soc_b), b_loss))); /* :307 */
  } else if (jgt(Soc_b, J_F(1e-3f))) { /* :309 */
    B = jneg(jmin(b_rate_max, jmul(jsub(J_I(1), b_loss), Soc_b))); /* :310 */
  } else {
    B = J_I(0); /* :312 */
  }
  out[0] = to_f32(B); out[1] = to_f32(EV); /* Float32.([B, EV]) :315 */
}

/* ------------------------------------------------- action(env, track) rule-based :318-340 */
void oracle_action_rule(const ShemsParams* P, const float* s, float* out) {
  jl Soc_b = J_F(s[0]), Soc_ev = J_F(s[1]), d_e = J_F(s[3]), g_e = J_F(s[4]);
  jl b_soc_max = J_F(P->b_soc_max), b_rate_max = J_D(P->b_rate_max), b_loss = J_F(P->b_loss);
  jl ev_soc_min = J_F(P->ev_soc_min), ev_soc_max = J_F(P->ev_soc_max), ev_rate_max = J_F(P->ev_rate_max);
  jl B;
  jl EV = jmin(ev_rate_max, jmul(jsub(J_I(1), Soc_ev), jsub(ev_soc_max, ev_soc_min))); /* :323 */
  jl pv_ = jsub(jsub(g_e, d_e), EV);                                                  /* :327 */
  if (jgt(pv_, J_I(0)) && jlt(Soc_b, jmul(J_D(0.95), b_soc_max))) {                    /* :330 */
    B = jclamp(pv_, J_I(0), jmin(b_rate_max, jadd(jsub(b_soc_max, Soc_b), b_loss)));   /* :331 */
  } else if (jgt(Soc_b, J_F(1e-3f))) {                                                 /* :333 */
    B = jneg(jmin(b_rate_max, jmul(jsub(J_I(1), b_loss), Soc_b)));                     /* :334 */
  } else {
    B = J_I(0);
  }
  out[0] = to_f32(B); out[1] = to_f32(EV); /* :339 */
}

/* ------------------------------------------------- step!(env, s, a; track) :343-485
 * state: in/out 9 floats; idx: in/out 1-based row; a: 2 floats; track: the sign matters only.
 * Returns 0, or SHEMS_ERR_BOUNDS when idx+1 > nrows (BoundsError at :268) — state untouched. */
int oracle_step(const ShemsParams* P, const float* series, int nrows, float* state, int* idx_io,
                const float* a, double track, double* reward_out, double* trace /* 23 or NULL */) {
  const int idx_old = *idx_io;
  if (idx_old + 1 > nrows || idx_old < 1) return SHEMS_ERR_BOUNDS;

  jl Soc_b = J_F(state[0]), Soc_ev = J_F(state[1]), c_ev = J_F(state[2]), d_e = J_F(state[3]),
     g_e = J_F(state[4]), p_buy = J_F(state[5]); /* :344 */
  jl pv_eta = J_F(P->pv_eta), b_eta = J_F(P->b_eta), b_soc_max = J_F(P->b_soc_max),
     b_rate_max = J_D(P->b_rate_max), b_loss = J_F(P->b_loss);
  jl ev_soc_min = J_F(P->ev_soc_min), ev_soc_max = J_F(P->ev_soc_max);
  jl B_target, EV_target, B, EV;

  if (track >= 0) { /* :346-349 */
    float be[2];
    B_target = J_F(a[0]); EV_target = J_F(a[1]);
    oracle_action_drl(P, state, a[0], a[1], be);
    B = J_F(be[0]); EV = J_F(be[1]);
  } else { /* :350-353 */
    B_target = J_F(0.f); EV_target = J_F(0.f);
    B = J_F(a[0]); EV = J_F(a[1]);
  }

  /* :356-357 — zeros(8), zeros(11): Float64 zeros */
  jl pv_ = J_D(0.0), BD = J_D(0.0), BC = J_D(0.0), discomfort = J_D(0.0), profit = J_D(0.0), penalty = J_D(0.0);
  jl PV_DE = J_D(0.0), PV_B = J_D(0.0), PV_EV = J_D(0.0), PV_GR = J_D(0.0), B_DE = J_D(0.0), B_EV = J_D(0.0),
     B_GR = J_D(0.0), GR_DE = J_D(0.0), GR_EV = J_D(0.0), GR_B = J_D(0.0), EX_EV = J_D(0.0);
  (void)profit;

  if (jlt(B, J_D(-0.01))) { /* :362 */
    BD = jclamp(jneg(B), J_D(0.001),
                jmin(b_rate_max, jmul(jsub(jsub(J_I(1), b_loss), J_F(1e-7f)), Soc_b))); /* :363 */
  }

  if (jgt(jmul(g_e, pv_eta), d_e)) { /* :368 */
    PV_DE = d_e;                              /* :369 */
    pv_ = jsub(jmul(g_e, pv_eta), PV_DE);     /* :370 */
    if (jgt(pv_, EV)) {                       /* :371 */
      PV_EV = EV;                             /* :372 */
      pv_ = jsub(pv_, PV_EV);                 /* :373 */
    } else if (jle(pv_, EV)) {                /* :374 */
      PV_EV = pv_;                            /* :375 */
      pv_ = J_I(0);                           /* :376 */
      if (jgt(BD, jdiv(jsub(EV, PV_EV), b_eta))) {        /* :377 */
        B_EV = jsub(EV, PV_EV);                           /* :378 */
        BD = jsub(BD, jdiv(B_EV, b_eta));                 /* :379 */
      } else if (jle(BD, jdiv(jsub(EV, PV_EV), b_eta))) { /* :380 */
        B_EV = jmul(BD, b_eta);                           /* :381 */
        BD = J_I(0);                                      /* :382 */
        GR_EV = jsub(jsub(EV, PV_EV), B_EV);              /* :383 */
      }
    }
  } else if (jle(jmul(g_e, pv_eta), d_e)) { /* :388 */
    PV_DE = jmul(g_e, pv_eta);              /* :389 */
    pv_ = J_I(0);                           /* :390 */
    d_e = jsub(d_e, PV_DE);                 /* :391 */
    if (jgt(BD, jdiv(d_e, b_eta))) {        /* :392 */
      B_DE = d_e;                           /* :393 */
      BD = jsub(BD, jdiv(B_DE, b_eta));     /* :394 */
      if (jgt(BD, jdiv(EV, b_eta))) {       /* :395 */
        B_EV = EV;                          /* :396 */
        BD = jsub(BD, jdiv(B_EV, b_eta));   /* :397 */
      } else if (jle(BD, jdiv(EV, b_eta))) { /* :398 */
        B_EV = jmul(BD, b_eta);             /* :399 */
        BD = J_I(0);                        /* :400 */
        GR_EV = jsub(EV, B_EV);             /* :401 */
      }
    } else if (jle(BD, jdiv(d_e, b_eta))) { /* :403 */
      B_DE = jmul(BD, b_eta);               /* :404 */
      BD = J_I(0);                          /* :405 */
      GR_DE = jsub(d_e, B_DE);              /* :406 */
      GR_EV = EV;                           /* :407 */
    }
  }

  if (jgt(B, J_D(0.01))) { /* :412 */
    BC = jclamp(B, J_D(0.001), jmin(b_rate_max, jsub(b_soc_max, Soc_b))); /* :413 */
    if (jgt(pv_, jdiv(BC, b_eta))) {        /* :414 */
      PV_B = BC;                            /* :415 */
      pv_ = jsub(pv_, jdiv(BC, b_eta));     /* :416 */
    } else if (jle(pv_, jdiv(BC, b_eta))) { /* :417 */
      PV_B = jmul(pv_, b_eta);              /* :418 */
      pv_ = J_I(0);                         /* :419 */
      GR_B = J_I(0);                        /* :420 */
    }
  }
  PV_GR = pv_;   /* :424 */
  B_GR = J_I(0); /* :425 */

  /* :432  (1 - b.loss) * (Soc_b + PV_B + GR_B - ((B_DE + B_EV + B_GR) / b.eta)) -> Float32 field */
  float Soc_b_new = to_f32(jmul(jsub(J_I(1), b_loss),
                                jsub(jadd(jadd(Soc_b, PV_B), GR_B), jdiv(jadd(jadd(B_DE, B_EV), B_GR), b_eta))));
  /* :435  Soc_ev + (PV_EV + B_EV + GR_EV) / (ev.soc_max - ev.soc_min) -> Float32 field */
  float Soc_ev_new = to_f32(jadd(Soc_ev, jdiv(jadd(jadd(PV_EV, B_EV), GR_EV), jsub(ev_soc_max, ev_soc_min))));

  discomfort = J_I(0); penalty = J_I(0); EX_EV = J_I(0); /* :438-440 */
  if (jeq(c_ev, J_I(0)) && jlt(J_F(Soc_ev_new), J_I(1))) { /* :442 */
    discomfort = jmul(jsub(J_I(1), J_F(Soc_ev_new)), J_I(100));                       /* :444 */
    EX_EV = jmul(jsub(J_I(1), J_F(Soc_ev_new)), jsub(ev_soc_max, ev_soc_min));        /* :445 */
    Soc_ev_new = 1.0f;                                                                /* :446 */
  } else if (jlt(c_ev, J_I(0)) && jlt(EV_target, J_D(0.99))) {                        /* :447 */
    /* :448; penalty_weight is a Float32 in shems_LU1.jl:43 and a Float64 in shems_LU7.jl:35 / shems_LU1_input0607.jl:52 */
    penalty = jmul(jsub(J_I(1), EV_target), P->penalty_in_f64 ? J_D(P->penalty_weight_f64) : J_F(P->penalty_weight));
  }

  /* next_state!(env) :264-281 with idx = env.idx + 1 */
  {
    const int idx = idx_old + 1;
    float c_new = COL(SHEMS_COL_H_COUNTDOWN, idx);                          /* :268 */
    if (c_new >= 0 && COL(SHEMS_COL_H_COUNTDOWN, idx_old) == -1.0f) {       /* :270 */
      Soc_ev_new = COL(SHEMS_COL_SOC_EV, idx);                              /* :271 */
    }
    state[0] = Soc_b_new;
    state[1] = Soc_ev_new;
    state[2] = c_new;
    state[3] = COL(SHEMS_COL_ELECTKWH, idx);      /* :274 */
    state[4] = COL(SHEMS_COL_PV_GENERATION, idx); /* :275 */
    state[5] = COL(SHEMS_COL_P_BUY, idx);         /* :276 */
    state[8] = COL(SHEMS_COL_SEASON, idx);        /* :277 */
    state[6] = COL(SHEMS_COL_HOUR_COS, idx);      /* :278 */
    state[7] = COL(SHEMS_COL_HOUR_SIN, idx);      /* :279 */
    *idx_io = idx;                                /* :456 (env.step += 1 is the caller's counter) */
  }

  /* :464  profit = (m.sell_discount * p_buy * (PV_GR + B_GR)) - (p_buy * (GR_DE + GR_B + GR_EV + EX_EV)) */
  jl sell = J_D(P->sell_discount), dw = J_D(P->discomfort_weight_ev), pot = J_D(P->disc_pot);
  profit = jsub(jmul(jmul(sell, p_buy), jadd(PV_GR, B_GR)),
                jmul(p_buy, jadd(jadd(jadd(GR_DE, GR_B), GR_EV), EX_EV)));
  jl reward;
  /* shems_LU1.jl:467-470 m.discomfort_weight_ev * (discomfort ^ m.disc_pot); shems_LU7.jl:465-468 discomfort * m.discomfort_weight_ev
   * (pot = 1: identical value); shems_LU1_input0607.jl:481-484 (discomfort * m.discomfort_weight_ev) ^ (m.disc_pot) */
  jl dterm = P->reward_form ? jpow(jmul(discomfort, dw), pot) : jmul(dw, jpow(discomfort, pot));
  if (track < 0) { /* :466-468 */
    reward = jsub(profit, dterm);
    penalty = J_I(0);
  } else {         /* :470 */
    reward = jsub(jsub(profit, dterm), penalty);
  }
  *reward_out = to_f64(reward); /* env.reward::Float64 */

  if (trace) { /* :476-478 hcat(...) -> Matrix{Float64} */
    trace[SHEMS_T_INDEX] = (double)*idx_io; trace[SHEMS_T_C_EV] = c_ev.v; trace[SHEMS_T_EV_TARGET] = EV_target.v;
    trace[SHEMS_T_EV] = EV.v; trace[SHEMS_T_SOC_EV] = Soc_ev.v; trace[SHEMS_T_REWARD] = reward.v;
    trace[SHEMS_T_PROFIT] = profit.v; trace[SHEMS_T_DISCOMFORT] = discomfort.v; trace[SHEMS_T_PENALTY] = penalty.v;
    trace[SHEMS_T_PV_DE] = PV_DE.v; trace[SHEMS_T_B_DE] = B_DE.v; trace[SHEMS_T_GR_DE] = GR_DE.v;
    trace[SHEMS_T_PV_B] = PV_B.v; trace[SHEMS_T_PV_GR] = PV_GR.v; trace[SHEMS_T_PV_EV] = PV_EV.v;
    trace[SHEMS_T_B_EV] = B_EV.v; trace[SHEMS_T_GR_EV] = GR_EV.v; trace[SHEMS_T_EX_EV] = EX_EV.v;
    trace[SHEMS_T_GR_B] = GR_B.v; trace[SHEMS_T_B_GR] = B_GR.v; trace[SHEMS_T_B] = B.v;
    trace[SHEMS_T_B_TARGET] = B_target.v; trace[SHEMS_T_SOC_B] = Soc_b.v;
  }
  return 0;
}

/* ------------------------------------------------- reset!/reset_state! :206-262
 * deterministic != 0  <=> rng == -1.  Otherwise (idx0, u_or_socb) are the two draws of :224-225:
 * idx0 in 1..nrows-maxsteps, socb0 = Float32 value stored into Soc_b.  Returns the final idx. */
int oracle_reset(const ShemsParams* P, const float* series, int nrows, int maxsteps, int deterministic,
                 int idx0, float socb0, float* state, int* idx_out) {
  int idx;
  const int hi = nrows - maxsteps;
  if (hi < 1) return SHEMS_ERR_INVALID; /* rand(1:0) throws */
  if (deterministic) {
    state[0] = to_f32(jmul(J_D(0.5), jadd(J_F(P->b_soc_min), J_F(P->b_soc_max)))); /* :221 */
    idx = 1;                                                                       /* :222 */
  } else {
    if (idx0 < 1 || idx0 > hi) return SHEMS_ERR_INVALID;
    state[0] = socb0; /* :224 */
    idx = idx0;       /* :225 */
    float c_ev_end = COL(SHEMS_COL_H_COUNTDOWN, idx + maxsteps); /* :227 */
    int counter = 0; const int max_iterations = 100;
    while (c_ev_end > -1 && idx < hi) { /* :231 */
      idx += (int)(c_ev_end + 1);       /* :232 */
      if (idx > hi) idx = idx0;         /* :235-237: a fresh MersenneTwister(rng) re-draws the SAME index */
      c_ev_end = COL(SHEMS_COL_H_COUNTDOWN, idx + maxsteps); /* :239 */
      counter += 1;
      if (counter > max_iterations) break; /* :242-245 */
    }
  }
  state[1] = COL(SHEMS_COL_SOC_EV, idx);        /* :251 */
  state[2] = COL(SHEMS_COL_H_COUNTDOWN, idx);   /* :254 */
  state[3] = COL(SHEMS_COL_ELECTKWH, idx);      /* :255 */
  state[4] = COL(SHEMS_COL_PV_GENERATION, idx); /* :256 */
  state[5] = COL(SHEMS_COL_P_BUY, idx);         /* :257 */
  state[8] = COL(SHEMS_COL_SEASON, idx);        /* :258 */
  state[6] = COL(SHEMS_COL_HOUR_COS, idx);      /* :259 */
  state[7] = COL(SHEMS_COL_HOUR_SIN, idx);      /* :260 */
  *idx_out = idx;
  return 0;
}

/* ------------------------------------------------- Philox4x32-10 (this repo's device RNG spec)
 * The reference draws from Julia's MersenneTwister, which cannot be reproduced here; the
 * product defines its own counter-based streams and the oracle restates them bit-exactly. */
static inline void philox_round(uint32_t* c, uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void oracle_philox(uint64_t seed, uint64_t id, uint32_t ctr, uint32_t stream, uint32_t* out4) {
  uint32_t c[4] = {(uint32_t)id, (uint32_t)(id >> 32), ctr, stream};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  memcpy(out4, c, sizeof(c));
}
/* 53-bit uniform in [0,1) from two words — stands in for Julia's Float64 rand() */
double oracle_u53(uint32_t a, uint32_t b) {
  return (double)((((uint64_t)(a >> 5)) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

/* stream ids (shared with the CUDA side: csrc/philox.cuh) */
enum { STREAM_RESET = 0x5245u, STREAM_ACTION = 0x4143u, STREAM_NOISE = 0x4e4fu, STREAM_SAMPLE = 0x534du, STREAM_INIT = 0x494eu };

/* device-Philox reset draws: Soc_b = Float32(soc_min + (soc_max-soc_min)*u) (Uniform rand :224),
 * idx0 = 1 + floor(u2 * (nrows-maxsteps)) (:225) */
void oracle_reset_draws(const ShemsParams* P, int nrows, int maxsteps, uint64_t seed, uint64_t env_id,
                        int* idx0, float* socb0) {
  uint32_t r[4];
  oracle_philox(seed, env_id, 0u, STREAM_RESET, r);
  const double u1 = oracle_u53(r[0], r[1]), u2 = oracle_u53(r[2], r[3]);
  volatile float span = P->b_soc_max - P->b_soc_min;
  volatile double v = (double)P->b_soc_min + (double)span * u1;
  *socb0 = (float)v;
  int hi = nrows - maxsteps;
  int k = (int)(u2 * (double)hi);
  if (k >= hi) k = hi - 1;
  *idx0 = 1 + k;
}

/* random warm-up action a = Float32.(rand(2) .* 2 .- 1) (memory_plotting_saving.jl:17) at (env, step).
 * Spec: one Philox block serves two consecutive steps (ctr = step >> 1; words 0,1 for even, 2,3 for odd
 * steps); u = word * 2^-32 stands in for Julia's Float64 rand(). */
void oracle_random_action(uint64_t seed, uint64_t env_id, uint32_t step, float* a) {
  uint32_t r[4];
  oracle_philox(seed, env_id, step >> 1, STREAM_ACTION, r);
  const uint32_t w0 = r[(step & 1u) * 2u], w1 = r[(step & 1u) * 2u + 1u];
  volatile double a0 = ((double)w0 * (1.0 / 4294967296.0)) * 2.0; a0 = a0 - 1.0;
  volatile double a1 = ((double)w1 * (1.0 / 4294967296.0)) * 2.0; a1 = a1 - 1.0;
  a[0] = (float)a0; a[1] = (float)a1;
}

/* scale_action (DDPG.jl:178-184): Float32.(LO .+ (a .+ ones(2)) .* 0.5 .* (HI .- LO)) */
void oracle_scale_action(const float* a, const float* lo, const float* hi, float* out) {
  for (int i = 0; i < 2; ++i) {
    volatile float span = hi[i] - lo[i];
    volatile double t = (double)a[i] + 1.0;
    t = t * 0.5;
    t = t * (double)span;
    t = (double)lo[i] + t;
    out[i] = (float)t;
  }
}

/* ------------------------------------------------- batched loops (CPU baseline + parity) */
/* T-step rollout of n envs; policy as ShemsRolloutArgs.policy.  Any output pointer may be NULL.
 * Layouts match the ABI: obs [9][n] in/out, idx [n] in/out, tape [T][2][n], ep_return [n],
 * trans_* [T][k][n] (s, a_unscaled, r, s2), trace [T][23][n]. */
int oracle_rollout(const ShemsParams* P, const float* series, int nrows, long long n, float* obs, int* idx,
                   int policy, int T, uint64_t seed, long long env_id_base, const float* tape,
                   double* ep_return, float* tr_s, float* tr_a, float* tr_r, float* tr_s2, double* trace,
                   int step0) {
  int status = 0;
  const float lo[2] = {0.f, 0.f}, hi[2] = {1.f, 1.f}; /* ACTION_BOUND_LO/HI, input.jl:182-183 */
#pragma omp parallel for schedule(static)
  for (long long e = 0; e < n; ++e) {
    float s[9]; int id = idx[e]; double ret = 0.0;
    for (int k = 0; k < 9; ++k) s[k] = obs[(size_t)k * n + e];
    for (int t = 0; t < T; ++t) {
      float a_raw[2] = {0.f, 0.f}, a_env[2]; double track, r; double tr[23];
      float s_prev[9]; memcpy(s_prev, s, sizeof(s));
      if (policy == SHEMS_POLICY_RULE) { oracle_action_rule(P, s, a_env); track = -0.5; a_raw[0] = a_env[0]; a_raw[1] = a_env[1]; }
      else if (policy == SHEMS_POLICY_RANDOM) { oracle_random_action(seed, (uint64_t)(env_id_base + e), (uint32_t)(step0 + t), a_raw); oracle_scale_action(a_raw, lo, hi, a_env); track = 0; }
      else { a_env[0] = tape[((size_t)t * 2 + 0) * n + e]; a_env[1] = tape[((size_t)t * 2 + 1) * n + e]; a_raw[0] = a_env[0]; a_raw[1] = a_env[1]; track = 0; }
      int st = oracle_step(P, series, nrows, s, &id, a_env, track, &r, trace ? tr : NULL);
      if (st) {
#pragma omp critical
        status = st;
        break;
      }
      ret += r; /* reward_eps += r (Float64 accumulation, DDPG.jl:223) */
      if (tr_s) for (int k = 0; k < 9; ++k) tr_s[((size_t)t * 9 + k) * n + e] = s_prev[k];
      if (tr_a) for (int k = 0; k < 2; ++k) tr_a[((size_t)t * 2 + k) * n + e] = a_raw[k];
      if (tr_r) tr_r[(size_t)t * n + e] = (float)r;
      if (tr_s2) for (int k = 0; k < 9; ++k) tr_s2[((size_t)t * 9 + k) * n + e] = s[k];
      if (trace) for (int k = 0; k < 23; ++k) trace[((size_t)t * 23 + k) * n + e] = tr[k];
    }
    for (int k = 0; k < 9; ++k) obs[(size_t)k * n + e] = s[k];
    idx[e] = id;
    if (ep_return) ep_return[e] = ret;
  }
  return status;
}

/* one step of n envs with given actions [2][n]; reward [n] double, trace [23][n] */
int oracle_step_batch(const ShemsParams* P, const float* series, int nrows, long long n, float* obs, int* idx,
                      const float* act, double track, double* reward, double* trace) {
  int status = 0;
#pragma omp parallel for schedule(static)
  for (long long e = 0; e < n; ++e) {
    float s[9], a[2] = {act[e], act[n + e]}; int id = idx[e]; double r = 0, tr[23];
    for (int k = 0; k < 9; ++k) s[k] = obs[(size_t)k * n + e];
    int st = oracle_step(P, series, nrows, s, &id, a, track, &r, trace ? tr : NULL);
    if (st) {
#pragma omp critical
      status = st;
      continue;
    }
    for (int k = 0; k < 9; ++k) obs[(size_t)k * n + e] = s[k];
    idx[e] = id;
    if (reward) reward[e] = r;
    if (trace) for (int k = 0; k < 23; ++k) trace[(size_t)k * n + e] = tr[k];
  }
  return status;
}

int oracle_reset_batch(const ShemsParams* P, const float* series, int nrows, int maxsteps, long long n, int mode,
                       const int* idx0, const float* socb0, uint64_t seed, long long env_id_base, float* obs, int* idx) {
  int status = 0;
#pragma omp parallel for schedule(static)
  for (long long e = 0; e < n; ++e) {
    float s[9]; int id = 0, i0 = 1; float sb = 0.f;
    if (mode == SHEMS_RESET_HOST_DRAWS) { i0 = idx0[e]; sb = socb0[e]; }
    else if (mode == SHEMS_RESET_DEVICE_PHILOX) oracle_reset_draws(P, nrows, maxsteps, seed, (uint64_t)(env_id_base + e), &i0, &sb);
    int st = oracle_reset(P, series, nrows, maxsteps, mode == SHEMS_RESET_DETERMINISTIC, i0, sb, s, &id);
    if (st) {
#pragma omp critical
      status = st;
      continue;
    }
    for (int k = 0; k < 9; ++k) obs[(size_t)k * n + e] = s[k];
    idx[e] = id;
  }
  return status;
}

void oracle_action_batch(const ShemsParams* P, long long n, const float* obs, const float* target /* NULL: rule */, float* bev) {
#pragma omp parallel for schedule(static)
  for (long long e = 0; e < n; ++e) {
    float s[9], o[2];
    for (int k = 0; k < 9; ++k) s[k] = obs[(size_t)k * n + e];
    if (target) oracle_action_drl(P, s, target[e], target[n + e], o); else oracle_action_rule(P, s, o);
    bev[e] = o[0]; bev[n + e] = o[1];
  }
}
