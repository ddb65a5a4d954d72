// common.h — internals shared by the translation units of libshems_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/shems_b200.h"

// thread-local last-error string (shems_last_error)
void shems_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      shems_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SHEMS_ERR_CUDA;                                                               \
    }                                                                                      \
  } while (0)

#define REQUIRE(cond, code, ...)  \
  do {                            \
    if (!(cond)) {                \
      shems_set_error(__VA_ARGS__); \
      return (code);              \
    }                             \
  } while (0)

// RAII device guard: every entry point runs on its handle's device and restores the caller's
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define GUARD(dev)                                             \
  DeviceGuard _guard(dev);                                     \
  if (!_guard.ok) {                                            \
    cudaError_t _e = cudaGetLastError();                       \
    shems_set_error("cannot select CUDA device %d: %s", (dev), cudaGetErrorString(_e)); \
    return SHEMS_ERR_CUDA;                                     \
  }

// constants of the env with the derived values every step needs (passed by value to kernels)
struct DevParams {
  float pv_eta, b_eta, smin, smax, loss, evmin, evmax, evR, pw;
  float one_m_l;     // Float32(1 - b.loss)
  float one_m_l_e;   // Float32(Float32(1 - b.loss) - 1f-7)
  float C;           // ev.soc_max - ev.soc_min
  float span;        // b.soc_max - b.soc_min
  float R_f;         // Float32(b.rate_max) (only where the Float64 min is provably equivalent)
  double R, sell, dw, pot;
  double pw_d;               // Float64 penalty weight of the sibling envs (pen_f64 != 0)
  int pen_f64, reward_form;
  int reward_mode;           // 0: w*discomfort^2 in form 0 (shems_LU1, the fast path); 1: anything else (shems_discomfort_term)
  double eta_d, one_m_l_d, C_d, smax95;
  // division by the constants eta / span / C as multiply + 2 FMA (see fdiv_const / ddiv_const in shems_device.cuh)
  float r_eta_f, r_span_f;   // Float32(1/eta), Float32(1/span)
  double r_eta_d, r_C_d;     // 1/Float64(eta), 1/Float64(C)
  int fast_eta_f, fast_span_f;  // host-verified over all 2^23 significands: the 3-op sequence == IEEE division
  int fast_d;                   // eta_d and C_d are Float32-valued: the 3-op sequence is provably correctly rounded
  double span_d, r_span_d;      // Float64(span), 1/Float64(span): the Float32 quotient Soc_b / span through Float64 (fdiv_const_wide)
  int fast_all;                 // fast_d && span Float32-valued && Float32 penalty weight && shems_LU1's reward form: kernels take the
                                // FAST instantiation (no flag tests, no guarded division sequences) — same results
};

// ---- replay ring layout: 22 fields per transition, tiled so that one transition's fields sit at fixed
// 128-byte strides: element (slot, k) lives at ((slot >> 5) * 22 + k) * 32 + (slot & 31).
// A warp writing 32 consecutive slots stores one full 128-byte line per field, with the field offset an
// immediate (no per-field address arithmetic in the rollout kernel).
#define RING_FIELDS 22
#define RING_S 0      // s[0..8]
#define RING_A 9      // a[0..1]
#define RING_R 11     // r
#define RING_S2 12    // s'[0..8]
#define RING_DONE 21  // done
__host__ __device__ __forceinline__ size_t ring_base(long long slot) { return ((size_t)(slot >> 5) * RING_FIELDS) * 32 + (size_t)(slot & 31); }
__host__ __device__ __forceinline__ size_t ring_off(long long slot, int k) { return ring_base(slot) + (size_t)k * 32; }

#define SHEMS_MAX_GROUPS 16

struct ShemsReplay {
  int device;
  cudaStream_t stream;
  int64_t capacity, length, head;  // head = physical slot of the next push
  float* ring;  // tiled SoA, see ring_off(): ceil(cap/32) tiles x 22 fields x 32 slots
  int32_t* idx_scratch;  // device scratch for sampled indices
  int64_t idx_scratch_n;
  float* minmax_scratch; // [18]
  void* group_refs; int group_refs_n;  // device scratch of replay_push_groups (ring / capacity / head of every learner's memory)
};

struct ShemsEnv {
  int device;
  cudaStream_t stream;
  DevParams dp;
  ShemsParams params;
  int32_t nrows, maxsteps;
  int64_t n;
  float4* series;   // [nrows][2] rows: (soc_ev, h_countdown, electkwh, PV_generation | p_buy, hour_cos, hour_sin, season)
  float* obs;       // [9][n]  == env.state of every instance
  int32_t* idx;     // [n] 1-based row (env.idx)
  int32_t* d_maxidx;  // device scalar: max idx after reset
  int32_t max_idx;    // host mirror: the largest row any instance is on, or (while max_pending) an upper bound of it; +1 per step
  int32_t* h_maxidx;  // pinned host word the reset kernel's maximum is copied to (asynchronously)
  cudaEvent_t ev_maxidx; bool max_pending;  // the exact value is fetched only when the upper bound does not settle a bounds check
  int32_t step;       // env.step (instances run in lockstep)
  bool was_reset;
  bool consistent;    // state fields 2..8 equal series row idx (false after shems_set_state)
  int32_t* scratch_i; float* scratch_f; // staging for reset host draws
  // groups of instances with their own constants (several chargers in one handle): group g = instances gstart[g] .. gstart[g+1]-1
  int n_groups;
  DevParams gdp[SHEMS_MAX_GROUPS];
  float4* gseries[SHEMS_MAX_GROUPS];
  long long gstart[SHEMS_MAX_GROUPS + 1];
};

int replay_after_rollout(ShemsReplay* rp, int64_t n_written);
// may every instance advance `need` more rows (next_state! reads row idx+1)?  0 = yes, else the furthest instance's row (env.cu)
int32_t ensure_rows(ShemsEnv* e, int64_t need);
