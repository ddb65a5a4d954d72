/* oracle.h — prototypes of the CPU oracle (test infrastructure; see shems_oracle.c header). */
#ifndef SHEMS_ORACLE_H
#define SHEMS_ORACLE_H
#include <stdint.h>
#include "../include/shems_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int oracle_set_threads(int n);
int oracle_params_for_charger(int charger_id, ShemsParams* p);
int oracle_params_for_env(int variant, int charger_id, ShemsParams* p);
void oracle_action_drl(const ShemsParams* P, const float* s, float B_target, float EV_target, float* out);
void oracle_action_rule(const ShemsParams* P, const float* s, float* out);
int oracle_step(const ShemsParams* P, const float* series, int nrows, float* state, int* idx_io,
                const float* a, double track, double* reward_out, double* trace);
int oracle_reset(const ShemsParams* P, const float* series, int nrows, int maxsteps, int deterministic,
                 int idx0, float socb0, float* state, int* idx_out);
void oracle_philox(uint64_t seed, uint64_t id, uint32_t ctr, uint32_t stream, uint32_t* out4);
double oracle_u53(uint32_t a, uint32_t b);
void oracle_reset_draws(const ShemsParams* P, int nrows, int maxsteps, uint64_t seed, uint64_t env_id,
                        int* idx0, float* socb0);
void oracle_random_action(uint64_t seed, uint64_t env_id, uint32_t step, float* a);
void oracle_scale_action(const float* a, const float* lo, const float* hi, float* out);
int oracle_rollout(const ShemsParams* P, const float* series, int nrows, long long n, float* obs, int* idx,
                   int policy, int T, uint64_t seed, long long env_id_base, const float* tape,
                   double* ep_return, float* tr_s, float* tr_a, float* tr_r, float* tr_s2, double* trace,
                   int step0);
int oracle_step_batch(const ShemsParams* P, const float* series, int nrows, long long n, float* obs, int* idx,
                      const float* act, double track, double* reward, double* trace);
int oracle_reset_batch(const ShemsParams* P, const float* series, int nrows, int maxsteps, long long n, int mode,
                       const int* idx0, const float* socb0, uint64_t seed, long long env_id_base, float* obs, int* idx);
void oracle_action_batch(const ShemsParams* P, long long n, const float* obs, const float* target, float* bev);
#ifdef __cplusplus
}
#endif
#endif
