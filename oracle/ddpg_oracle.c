/*
 * ddpg_oracle.c — CPU ORACLE (test infrastructure, NOT product code) for the DDPG minibatch
 * update of the reference:
 *   /root/reference/RL-SHEMS/algorithms/DDPG.jl  nets :30-46, soft_update! :99-103,
 *   update_model! :105-108, loss_crit :114, loss_act :116-119, replay :121-145, act :148-176,
 *   scale_action :178-184;  src/memory_plotting_saving.jl  normalize :55-57.
 *
 * PARITY UNPINNED: the arithmetic lives in un-vendored Julia packages (Flux 0.12.1 Dense /
 * mse / ADAM, Zygote 0.6.12, NNlib 0.7.21 relu/tanh, CUBLAS SGEMM) and the reference holds no
 * test or golden vector for it.  Their published definitions are restated here:
 *   Dense: σ.(W*x .+ b), W is out×in;  mse = mean((ŷ .- y).^2);  relu'(z) = z > 0;
 *   ADAM.apply!: mt = β1*mt + (1-β1)*Δ; vt = β2*vt + (1-β2)*Δ^2;
 *                Δ = mt/(1-β1^t) / (√(vt/(1-β2^t)) + ϵ) * η  with β, ϵ Float64 (so the element
 *                math runs in Float64 and is stored back as Float32); x .-= Δ.
 * tests/test_ddpg_oracle.py cross-checks one whole update against torch autograd +
 * torch.optim.Adam on CPU (a second, independent restatement).
 *
 * Dot products are accumulated in double and rounded once to float — the oracle is the
 * "exact fp32-storage" answer; the CUDA kernels accumulate in fp32 in a different order and
 * are compared with a stated tolerance.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/shems_b200.h"
#include "oracle.h"
#include "ddpg_oracle.h"

enum { ACT_RELU = 0, ACT_TANH = 1, ACT_ID = 2 };

typedef struct { int in, out, act; float *W, *b; } Layer; /* W[o + out*i] (Flux column-major out×in) */
typedef struct { Layer l[3]; } Net;
typedef struct { float *m, *v; } Moments;

struct OracleDdpg {
  DdpgParams p;
  Net net[4];
  float* gW[2][3]; float* gb[2][3];     /* grads of ACTOR/CRITIC */
  Moments mW[2][3], mb[2][3];           /* ADAM state per param array */
  double beta_pow[2][2];                /* βp per optimiser (same for every array of one net) */
  float s_min[9], s_max[9];
  float loss_crit, loss_act;
};

static void layer_alloc(Layer* L, int in, int out, int act) {
  L->in = in; L->out = out; L->act = act;
  L->W = (float*)calloc((size_t)in * out, sizeof(float));
  L->b = (float*)calloc((size_t)out, sizeof(float));
}
static void net_alloc(Net* n, int in, int l1, int l2, int out, int last_act) {
  layer_alloc(&n->l[0], in, l1, ACT_RELU);
  layer_alloc(&n->l[1], l1, l2, ACT_RELU);
  layer_alloc(&n->l[2], l2, out, last_act);
}
static void net_copy(Net* dst, const Net* src) {
  for (int k = 0; k < 3; ++k) {
    memcpy(dst->l[k].W, src->l[k].W, sizeof(float) * (size_t)src->l[k].in * src->l[k].out);
    memcpy(dst->l[k].b, src->l[k].b, sizeof(float) * (size_t)src->l[k].out);
  }
}

OracleDdpg* oracle_ddpg_create(const DdpgParams* p) {
  OracleDdpg* h = (OracleDdpg*)calloc(1, sizeof(OracleDdpg));
  h->p = *p;
  const int S = p->state_size, A = p->action_size;
  net_alloc(&h->net[DDPG_NET_ACTOR], S, p->l1, p->l2, A, ACT_TANH);          /* DDPG.jl:30-36 */
  net_alloc(&h->net[DDPG_NET_CRITIC], S + A, p->l1, p->l2, 1, ACT_ID);       /* :41-44 */
  net_alloc(&h->net[DDPG_NET_ACTOR_TARGET], S, p->l1, p->l2, A, ACT_TANH);   /* :38 */
  net_alloc(&h->net[DDPG_NET_CRITIC_TARGET], S + A, p->l1, p->l2, 1, ACT_ID); /* :46 */
  for (int n = 0; n < 2; ++n)
    for (int k = 0; k < 3; ++k) {
      const Layer* L = &h->net[n].l[k];
      size_t nw = (size_t)L->in * L->out;
      h->gW[n][k] = (float*)calloc(nw, sizeof(float)); h->gb[n][k] = (float*)calloc(L->out, sizeof(float));
      h->mW[n][k].m = (float*)calloc(nw, sizeof(float)); h->mW[n][k].v = (float*)calloc(nw, sizeof(float));
      h->mb[n][k].m = (float*)calloc(L->out, sizeof(float)); h->mb[n][k].v = (float*)calloc(L->out, sizeof(float));
    }
  for (int n = 0; n < 2; ++n) { h->beta_pow[n][0] = p->adam_beta1; h->beta_pow[n][1] = p->adam_beta2; }
  for (int k = 0; k < 9; ++k) { h->s_min[k] = 0.f; h->s_max[k] = 1.f; }
  return h;
}
void oracle_ddpg_destroy(OracleDdpg* h) {
  if (!h) return;
  for (int n = 0; n < 4; ++n) for (int k = 0; k < 3; ++k) { free(h->net[n].l[k].W); free(h->net[n].l[k].b); }
  for (int n = 0; n < 2; ++n) for (int k = 0; k < 3; ++k) {
    free(h->gW[n][k]); free(h->gb[n][k]); free(h->mW[n][k].m); free(h->mW[n][k].v); free(h->mb[n][k].m); free(h->mb[n][k].v);
  }
  free(h);
}
void oracle_ddpg_set_layer(OracleDdpg* h, int net, int layer, const float* w, const float* b) {
  Layer* L = &h->net[net].l[layer];
  if (w) memcpy(L->W, w, sizeof(float) * (size_t)L->in * L->out);
  if (b) memcpy(L->b, b, sizeof(float) * (size_t)L->out);
}
void oracle_ddpg_get_layer(const OracleDdpg* h, int net, int layer, float* w, float* b) {
  const Layer* L = &h->net[net].l[layer];
  if (w) memcpy(w, L->W, sizeof(float) * (size_t)L->in * L->out);
  if (b) memcpy(b, L->b, sizeof(float) * (size_t)L->out);
}
void oracle_ddpg_get_grad(const OracleDdpg* h, int net, int layer, float* w, float* b) {
  const Layer* L = &h->net[net].l[layer];
  if (w) memcpy(w, h->gW[net][layer], sizeof(float) * (size_t)L->in * L->out);
  if (b) memcpy(b, h->gb[net][layer], sizeof(float) * (size_t)L->out);
}
void oracle_ddpg_set_norm(OracleDdpg* h, const float* s_min, const float* s_max) {
  memcpy(h->s_min, s_min, sizeof(float) * h->p.state_size);
  memcpy(h->s_max, s_max, sizeof(float) * h->p.state_size);
}
void oracle_ddpg_get_losses(const OracleDdpg* h, float* lc, float* la) { *lc = h->loss_crit; *la = h->loss_act; }

/* init (this repo's Philox spec; the distributions are DDPG.jl:21-22):
 * hidden: glorot_uniform = (rand(Float32) - 0.5f0) * sqrt(24f0/(fan_in+fan_out)); last: 6f-3*rand(Float32) - 3f-3 */
enum { STREAM_INIT = 0x494eu };
void oracle_ddpg_init(OracleDdpg* h, uint64_t seed) {
  for (int n = 0; n < 2; ++n)
    for (int k = 0; k < 3; ++k) {
      Layer* L = &h->net[n].l[k];
      const size_t nw = (size_t)L->in * L->out;
      const float scale = sqrtf(24.0f / (float)(L->in + L->out));
      for (size_t e = 0; e < nw; ++e) {
        uint32_t r[4];
        oracle_philox(seed, (uint64_t)e, (uint32_t)(n * 3 + k), STREAM_INIT, r);
        const float u = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
        if (k < 2) { const float t = u - 0.5f; L->W[e] = t * scale; }
        else { const float t = 6e-3f * u; L->W[e] = t - 3e-3f; }
      }
      memset(L->b, 0, sizeof(float) * (size_t)L->out);
    }
  net_copy(&h->net[DDPG_NET_ACTOR_TARGET], &h->net[DDPG_NET_ACTOR]);
  net_copy(&h->net[DDPG_NET_CRITIC_TARGET], &h->net[DDPG_NET_CRITIC]);
}

/* normalize (memory_plotting_saving.jl:55-57): (s - s_min) / (s_max - s_min + 1f-8), Float32 ops.
 * in: [S][B] SoA -> out x[b*ld + k] (sample-major rows) */
static void normalize_into(const OracleDdpg* h, const float* s, int B, float* x, int ld) {
  const int S = h->p.state_size;
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < S; ++k) {
      const float num = s[(size_t)k * B + b] - h->s_min[k];
      const float span = h->s_max[k] - h->s_min[k];
      const float den = span + 1e-8f;
      x[(size_t)b * ld + k] = num / den;
    }
}

/* Dense forward for B samples: x [B][in] -> y [B][out]; z kept only through y (relu/tanh invertible enough) */
static void dense_fwd(const Layer* L, const float* x, int B, float* y) {
  /* every output (b, o) is the Float64 sum over i = 0, 1, ... of W[o, i] * x[b, i], in that order; the loops run o innermost so that
   * the column-major Flux weight is walked contiguously (same sums, same order per element) */
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    double acc[2048];
    const int out = L->out;
    for (int o0 = 0; o0 < out; o0 += 2048) {
      const int on = out - o0 < 2048 ? out - o0 : 2048;
      for (int o = 0; o < on; ++o) acc[o] = 0.0;
      for (int i = 0; i < L->in; ++i) {
        const double xv = (double)x[(size_t)b * L->in + i];
        const float* w = L->W + (size_t)out * i + o0;
        for (int o = 0; o < on; ++o) acc[o] += (double)w[o] * xv;
      }
      for (int o = 0; o < on; ++o) {
        float z = (float)acc[o];   /* W*x */
        z = z + L->b[o0 + o];      /* .+ b */
        if (L->act == ACT_RELU) z = z > 0.f ? z : 0.f;
        else if (L->act == ACT_TANH) z = tanhf(z);
        y[(size_t)b * out + o0 + o] = z;
      }
    }
  }
}
/* dy (grad wrt layer output) -> dz in place, then gW, gb, and dx (if dx != NULL) */
static void dense_bwd(const Layer* L, const float* x, const float* y, float* dy, int B, float* gW, float* gb, float* dx) {
  for (int b = 0; b < B; ++b)
    for (int o = 0; o < L->out; ++o) {
      const size_t j = (size_t)b * L->out + o;
      if (L->act == ACT_RELU) dy[j] = y[j] > 0.f ? dy[j] : 0.f;
      else if (L->act == ACT_TANH) { const float t = y[j] * y[j]; dy[j] = dy[j] * (1.f - t); }
    }
  if (gW) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < L->in; ++i) {   /* gW[o, i] = sum over b = 0, 1, ... of dy[b, o] * x[b, i] (o innermost: contiguous, same sums) */
      double acc[2048];
      const int out = L->out;
      for (int o0 = 0; o0 < out; o0 += 2048) {
        const int on = out - o0 < 2048 ? out - o0 : 2048;
        for (int o = 0; o < on; ++o) acc[o] = 0.0;
        for (int b = 0; b < B; ++b) {
          const double xv = (double)x[(size_t)b * L->in + i];
          const float* d = dy + (size_t)b * out + o0;
          for (int o = 0; o < on; ++o) acc[o] += (double)d[o] * xv;
        }
        for (int o = 0; o < on; ++o) gW[(size_t)(o0 + o) + (size_t)out * i] = (float)acc[o];
      }
    }
    for (int o = 0; o < L->out; ++o) {
      double acc = 0.0;
      for (int b = 0; b < B; ++b) acc += (double)dy[(size_t)b * L->out + o];
      gb[o] = (float)acc;
    }
  }
  if (dx) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < L->in; ++i) {
        double acc = 0.0;
        for (int o = 0; o < L->out; ++o) acc += (double)L->W[(size_t)o + (size_t)L->out * i] * (double)dy[(size_t)b * L->out + o];
        dx[(size_t)b * L->in + i] = (float)acc;
      }
  }
}

/* Flux.Optimise.ADAM apply! + update! (x .-= Δ) on one array */
static void adam_array(float* x, const float* g, float* m, float* v, size_t n, double b1, double b2, double eps,
                       double bp1, double bp2, float eta) {
  for (size_t j = 0; j < n; ++j) {
    const float g2 = g[j] * g[j];                                         /* Δ^2 in Float32 */
    m[j] = (float)(b1 * (double)m[j] + (1.0 - b1) * (double)g[j]);
    v[j] = (float)(b2 * (double)v[j] + (1.0 - b2) * (double)g2);
    const double d = (double)m[j] / (1.0 - bp1) / (sqrt((double)v[j] / (1.0 - bp2)) + eps) * (double)eta;
    const float df = (float)d;
    x[j] = x[j] - df;
  }
}
static void adam_net(OracleDdpg* h, int n, float eta) {
  for (int k = 0; k < 3; ++k) {
    Layer* L = &h->net[n].l[k];
    adam_array(L->W, h->gW[n][k], h->mW[n][k].m, h->mW[n][k].v, (size_t)L->in * L->out, h->p.adam_beta1, h->p.adam_beta2,
               h->p.adam_eps, h->beta_pow[n][0], h->beta_pow[n][1], eta);
    adam_array(L->b, h->gb[n][k], h->mb[n][k].m, h->mb[n][k].v, (size_t)L->out, h->p.adam_beta1, h->p.adam_beta2,
               h->p.adam_eps, h->beta_pow[n][0], h->beta_pow[n][1], eta);
  }
  h->beta_pow[n][0] *= h->p.adam_beta1; h->beta_pow[n][1] *= h->p.adam_beta2;
}
/* soft_update! (DDPG.jl:99-103): p_t .= (1f0 - τ) * p_t .+ τ * p_m */
static void soft_update(Net* t, const Net* m, float tau) {
  const float omt = 1.0f - tau;
  for (int k = 0; k < 3; ++k) {
    const size_t nw = (size_t)m->l[k].in * m->l[k].out;
    for (size_t j = 0; j < nw; ++j) { const float a = omt * t->l[k].W[j]; const float b = tau * m->l[k].W[j]; t->l[k].W[j] = a + b; }
    for (int j = 0; j < m->l[k].out; ++j) { const float a = omt * t->l[k].b[j]; const float b = tau * m->l[k].b[j]; t->l[k].b[j] = a + b; }
  }
}

static void net_fwd(const Net* n, const float* x, int B, float* h1, float* h2, float* y) {
  dense_fwd(&n->l[0], x, B, h1); dense_fwd(&n->l[1], h1, B, h2); dense_fwd(&n->l[2], h2, B, y);
}

/* act (DDPG.jl:148-176) for n states with caller noise (or NULL): a = clamp(actor(normalize(s)) + noise, -1, 1) */
void oracle_ddpg_act(OracleDdpg* h, const float* obs /*[S][n]*/, int n, const float* noise /*[A][n] or NULL*/,
                     float* a_out /*[A][n]*/, float* scaled_out /*[A][n] or NULL*/) {
  const int S = h->p.state_size, A = h->p.action_size, l1 = h->p.l1, l2 = h->p.l2;
  float* x = (float*)malloc(sizeof(float) * (size_t)n * S);
  float* h1 = (float*)malloc(sizeof(float) * (size_t)n * l1);
  float* h2 = (float*)malloc(sizeof(float) * (size_t)n * l2);
  float* y = (float*)malloc(sizeof(float) * (size_t)n * A);
  normalize_into(h, obs, n, x, S);
  net_fwd(&h->net[DDPG_NET_ACTOR], x, n, h1, h2, y);
  for (int b = 0; b < n; ++b) {
    float av[2];
    for (int k = 0; k < A; ++k) {
      float v = y[(size_t)b * A + k];
      if (noise) v = v + noise[(size_t)k * n + b];
      v = v > 1.f ? 1.f : (v < -1.f ? -1.f : v);
      a_out[(size_t)k * n + b] = v; av[k] = v;
    }
    if (scaled_out) { float sc[2]; oracle_scale_action(av, h->p.act_lo, h->p.act_hi, sc); for (int k = 0; k < A; ++k) scaled_out[(size_t)k * n + b] = sc[k]; }
  }
  free(x); free(h1); free(h2); free(y);
}

/* sample_noise(ou::OUNoise, rng) (DDPG.jl:49-55) for n instances, each with its own X (ou_x [A][n], advanced in place) and
 * its standard normal draws z [A][n] (Julia: randn(length(ou.X)), Float64).  OUNoise(μ, σ, θ, dt, X) holds Float32 values
 * (input.jl:216-234):  dx = θ .* (μ .- X) .* dt  is Float32;  dx .+= σ .* sqrt(dt) .* randn  evaluates
 * Float64(dx) + Float64(Float32(σ*√dt)) * z and stores Float32;  X .+= dx in Float32;  returns Float32.(X). */
void oracle_ou_noise(float theta, float mu, float sigma, float dt, float* ou_x, const double* z, int n, int A, float* noise_out) {
  volatile float ssd = sigma * sqrtf(dt);
  for (int k = 0; k < A; ++k)
    for (int b = 0; b < n; ++b) {
      const size_t e = (size_t)k * n + b;
      volatile float d0 = mu - ou_x[e];
      volatile float d1 = theta * d0;
      volatile float dx = d1 * dt;
      volatile float dx2 = (float)((double)dx + (double)ssd * z[e]);
      volatile float xn = ou_x[e] + dx2;
      ou_x[e] = xn;
      noise_out[e] = xn;
    }
}

/* replay() body on a given minibatch (DDPG.jl:131-143). Arrays are SoA [k][B]. */
void oracle_ddpg_update_batch(OracleDdpg* h, const float* s, const float* a, const float* r, const float* s2, const float* done) {
  const int S = h->p.state_size, A = h->p.action_size, l1 = h->p.l1, l2 = h->p.l2, B = h->p.batch, C = S + A;
  float* xs = (float*)calloc((size_t)B * C, sizeof(float));   /* vcat(normalize(s), a)  */
  float* xs2 = (float*)calloc((size_t)B * C, sizeof(float));  /* vcat(normalize(s'), a') */
  float* sn = (float*)calloc((size_t)B * S, sizeof(float));
  float* s2n = (float*)calloc((size_t)B * S, sizeof(float));
  float* h1 = (float*)calloc((size_t)B * l1, sizeof(float)); float* h2 = (float*)calloc((size_t)B * l2, sizeof(float));
  float* g1 = (float*)calloc((size_t)B * l1, sizeof(float)); float* g2 = (float*)calloc((size_t)B * l2, sizeof(float));
  float* ah1 = (float*)calloc((size_t)B * l1, sizeof(float)); float* ah2 = (float*)calloc((size_t)B * l2, sizeof(float));
  float* a2 = (float*)calloc((size_t)B * A, sizeof(float)); float* api = (float*)calloc((size_t)B * A, sizeof(float));
  float* q = (float*)calloc((size_t)B, sizeof(float)); float* q2 = (float*)calloc((size_t)B, sizeof(float));
  float* y = (float*)calloc((size_t)B, sizeof(float)); float* dq = (float*)calloc((size_t)B, sizeof(float));
  float* dx = (float*)calloc((size_t)B * C, sizeof(float)); float* da = (float*)calloc((size_t)B * A, sizeof(float));

  normalize_into(h, s, B, sn, S); normalize_into(h, s2, B, s2n, S);
  /* a' = actor_target(normalize(s')); q' = critic_target(vcat(normalize(s'), a'))   :131-132 */
  net_fwd(&h->net[DDPG_NET_ACTOR_TARGET], s2n, B, h1, h2, a2);
  for (int b = 0; b < B; ++b) {
    for (int k = 0; k < S; ++k) { xs2[(size_t)b * C + k] = s2n[(size_t)b * S + k]; xs[(size_t)b * C + k] = sn[(size_t)b * S + k]; }
    for (int k = 0; k < A; ++k) { xs2[(size_t)b * C + S + k] = a2[(size_t)b * A + k]; xs[(size_t)b * C + S + k] = a[(size_t)k * B + b]; }
  }
  net_fwd(&h->net[DDPG_NET_CRITIC_TARGET], xs2, B, h1, h2, q2);
  /* y = r .+ γ .* (1 .- done) .* q'   :133 (Float32) */
  for (int b = 0; b < B; ++b) {
    const float nd = 1.f - (done ? done[b] : 0.f);
    const float t = h->p.gamma * nd;
    const float u = t * q2[b];
    y[b] = r[b] + u;
  }
  /* critic step :137, :114: loss = mean((critic(vcat(s_n, a)) - y)^2) */
  net_fwd(&h->net[DDPG_NET_CRITIC], xs, B, h1, h2, q);
  {
    double acc = 0.0;
    for (int b = 0; b < B; ++b) { const float d = q[b] - y[b]; acc += (double)(d * d); dq[b] = 2.f * d / (float)B; }
    h->loss_crit = (float)(acc / B);
  }
  {
    const Net* n = &h->net[DDPG_NET_CRITIC];
    dense_bwd(&n->l[2], h2, q, dq, B, h->gW[1][2], h->gb[1][2], g2);
    dense_bwd(&n->l[1], h1, h2, g2, B, h->gW[1][1], h->gb[1][1], g1);
    dense_bwd(&n->l[0], xs, h1, g1, B, h->gW[1][0], h->gb[1][0], NULL);
  }
  adam_net(h, DDPG_NET_CRITIC, h->p.lr_critic);
  /* actor step :140, :116-119: loss = -mean(critic(vcat(s_n, actor(s_n)))) with the UPDATED critic */
  net_fwd(&h->net[DDPG_NET_ACTOR], sn, B, ah1, ah2, api);
  for (int b = 0; b < B; ++b) for (int k = 0; k < A; ++k) xs[(size_t)b * C + S + k] = api[(size_t)b * A + k];
  net_fwd(&h->net[DDPG_NET_CRITIC], xs, B, h1, h2, q);
  {
    double acc = 0.0;
    for (int b = 0; b < B; ++b) { acc += (double)q[b]; dq[b] = -1.f / (float)B; }
    h->loss_act = (float)(-acc / B);
  }
  {
    const Net* c = &h->net[DDPG_NET_CRITIC];
    dense_bwd(&c->l[2], h2, q, dq, B, NULL, NULL, g2);
    dense_bwd(&c->l[1], h1, h2, g2, B, NULL, NULL, g1);
    dense_bwd(&c->l[0], xs, h1, g1, B, NULL, NULL, dx);
    for (int b = 0; b < B; ++b) for (int k = 0; k < A; ++k) da[(size_t)b * A + k] = dx[(size_t)b * C + S + k];
    const Net* n = &h->net[DDPG_NET_ACTOR];
    dense_bwd(&n->l[2], ah2, api, da, B, h->gW[0][2], h->gb[0][2], g2);
    dense_bwd(&n->l[1], ah1, ah2, g2, B, h->gW[0][1], h->gb[0][1], g1);
    dense_bwd(&n->l[0], sn, ah1, g1, B, h->gW[0][0], h->gb[0][0], NULL);
  }
  adam_net(h, DDPG_NET_ACTOR, h->p.lr_actor);
  /* :142-143 */
  soft_update(&h->net[DDPG_NET_ACTOR_TARGET], &h->net[DDPG_NET_ACTOR], h->p.tau);
  soft_update(&h->net[DDPG_NET_CRITIC_TARGET], &h->net[DDPG_NET_CRITIC], h->p.tau);

  free(xs); free(xs2); free(sn); free(s2n); free(h1); free(h2); free(g1); free(g2); free(ah1); free(ah2);
  free(a2); free(api); free(q); free(q2); free(y); free(dq); free(dx); free(da);
}

/* getData index draw (memory_plotting_saving.jl:31-33): batch i.i.d. indices in [0, len) from
 * Philox(seed, id = draw number, ctr = update counter) — this repo's spec for StatsBase.sample */
enum { STREAM_SAMPLE = 0x534du };
void oracle_sample_indices(uint64_t seed, uint32_t update, long long len, int batch, int* idx_out) {
  for (int j = 0; j < batch; ++j) {
    uint32_t r[4];
    oracle_philox(seed, (uint64_t)j, update, STREAM_SAMPLE, r);
    long long k = (long long)(oracle_u53(r[0], r[1]) * (double)len);
    if (k >= len) k = len - 1;
    idx_out[j] = (int)k;
  }
}
