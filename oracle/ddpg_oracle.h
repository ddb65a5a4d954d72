/* ddpg_oracle.h — prototypes of the DDPG CPU oracle (test infrastructure). */
#ifndef DDPG_ORACLE_H
#define DDPG_ORACLE_H
#include <stdint.h>
#include "../include/shems_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct OracleDdpg OracleDdpg;
OracleDdpg* oracle_ddpg_create(const DdpgParams* p);
void oracle_ddpg_destroy(OracleDdpg* h);
void oracle_ddpg_set_layer(OracleDdpg* h, int net, int layer, const float* w, const float* b);
void oracle_ddpg_get_layer(const OracleDdpg* h, int net, int layer, float* w, float* b);
void oracle_ddpg_get_grad(const OracleDdpg* h, int net, int layer, float* w, float* b);
void oracle_ddpg_set_norm(OracleDdpg* h, const float* s_min, const float* s_max);
void oracle_ddpg_get_losses(const OracleDdpg* h, float* lc, float* la);
void oracle_ddpg_init(OracleDdpg* h, uint64_t seed);
void oracle_ddpg_act(OracleDdpg* h, const float* obs, int n, const float* noise, float* a_out, float* scaled_out);
void oracle_ou_noise(float theta, float mu, float sigma, float dt, float* ou_x, const double* z, int n, int A, float* noise_out);
void oracle_ddpg_update_batch(OracleDdpg* h, const float* s, const float* a, const float* r, const float* s2, const float* done);
void oracle_sample_indices(uint64_t seed, uint32_t update, long long len, int batch, int* idx_out);
#ifdef __cplusplus
}
#endif
#endif
