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
  double eta_d, one_m_l_d, C_d, smax95;
};

struct ShemsReplay {
  int device;
  cudaStream_t stream;
  int64_t capacity, length, head;  // head = physical slot of the next push
  float* s;     // [9][cap]
  float* a;     // [2][cap]
  float* r;     // [cap]
  float* s2;    // [9][cap]
  float* done;  // [cap]
  int32_t* idx_scratch;  // device scratch for sampled indices
  int64_t idx_scratch_n;
  float* minmax_scratch; // [18]
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
  int32_t max_idx;    // host mirror (advances by 1 per step)
  int32_t step;       // env.step (instances run in lockstep)
  bool was_reset;
  bool consistent;    // state fields 2..8 equal series row idx (false after shems_set_state)
  int32_t* scratch_i; float* scratch_f; // staging for reset host draws
};

int replay_after_rollout(ShemsReplay* rp, int64_t n_written);
