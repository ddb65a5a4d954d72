// replay.cu — device-resident replay memory: `memory = CircularBuffer{Any}(MEM_SIZE)` of
// [s, a, r, s', done] (RL-SHEMS/input.jl:140; src/memory_plotting_saving.jl:31-57) as a
// structure-of-arrays ring in HBM with on-device sampling.
//
// Layout: 22 float32 fields per transition (s 9, a 2, r 1, s' 9, done 1 = 88 B) in 32-slot tiles, see
// ring_off() in common.h.  Logical index i (0 = oldest) lives in physical slot (head - length + i) mod cap.
#include <new>
#include <vector>
#include <string.h>

#include "common.h"
#include "philox.cuh"

extern "C" int32_t replay_create(int64_t capacity, int32_t device, ShemsReplay** out) {
  REQUIRE(out && capacity >= 1, SHEMS_ERR_INVALID, "replay_create: capacity=%lld", (long long)capacity);
  REQUIRE(shems_device_count() > 0, SHEMS_ERR_CUDA, "replay_create: no CUDA device (this library has no CPU fallback)");
  GUARD(device);
  ShemsReplay* rp = new (std::nothrow) ShemsReplay();
  REQUIRE(rp, SHEMS_ERR_INVALID, "replay_create: out of host memory");
  memset(rp, 0, sizeof(*rp));
  rp->device = device; rp->capacity = capacity;
  const size_t tiles = ((size_t)capacity + 31) / 32;
  cudaError_t st;
  if ((st = cudaMalloc(&rp->ring, sizeof(float) * RING_FIELDS * 32 * tiles)) != cudaSuccess ||
      (st = cudaMemset(rp->ring, 0, sizeof(float) * RING_FIELDS * 32 * tiles)) != cudaSuccess ||
      (st = cudaMalloc(&rp->minmax_scratch, sizeof(float) * 18)) != cudaSuccess) {
    shems_set_error("replay_create: %s", cudaGetErrorString(st));
    replay_destroy(rp);
    return SHEMS_ERR_CUDA;
  }
  *out = rp;
  return SHEMS_OK;
}
extern "C" int32_t replay_destroy(ShemsReplay* rp) {
  if (!rp) return SHEMS_OK;
  GUARD(rp->device);
  cudaFree(rp->ring); cudaFree(rp->idx_scratch); cudaFree(rp->minmax_scratch); cudaFree(rp->group_refs);
  delete rp;
  return SHEMS_OK;
}
extern "C" int32_t replay_set_stream(ShemsReplay* rp, void* s) {
  REQUIRE(rp, SHEMS_ERR_INVALID, "replay_set_stream: NULL handle");
  rp->stream = (cudaStream_t)s;
  return SHEMS_OK;
}
extern "C" int64_t replay_length(const ShemsReplay* rp) { return rp ? rp->length : 0; }
extern "C" int64_t replay_capacity(const ShemsReplay* rp) { return rp ? rp->capacity : 0; }

int replay_after_rollout(ShemsReplay* rp, int64_t n_written) {
  rp->head = (rp->head + n_written) % rp->capacity;
  rp->length = rp->length + n_written > rp->capacity ? rp->capacity : rp->length + n_written;
  return SHEMS_OK;
}

// remember() for n transitions: slot = (head + i) mod cap.  When n > cap only the last cap survive.
__global__ void __launch_bounds__(256)
replay_push_kernel(float* __restrict__ ring, long long cap, long long head, const float* __restrict__ s, const float* __restrict__ a,
                   const float* __restrict__ r, const float* __restrict__ s2, const float* __restrict__ done, long long n, long long first) {
  const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long slot = head + i;
  slot -= (slot / cap) * cap;
  float* q = ring + ring_base(slot);
#pragma unroll
  for (int k = 0; k < 9; ++k) q[(RING_S + k) * 32] = s[k * n + i];
  q[(RING_A + 0) * 32] = a[i];
  q[(RING_A + 1) * 32] = a[n + i];
  q[RING_R * 32] = r[i];
#pragma unroll
  for (int k = 0; k < 9; ++k) q[(RING_S2 + k) * 32] = s2[k * n + i];
  q[RING_DONE * 32] = done ? done[i] : 0.0f;
}

extern "C" int32_t replay_push(ShemsReplay* rp, const float* s_dev, const float* a_dev, const float* r_dev, const float* s2_dev,
                               const float* done_dev, int64_t n) {
  REQUIRE(rp && s_dev && a_dev && r_dev && s2_dev, SHEMS_ERR_INVALID, "replay_push: NULL argument");
  REQUIRE(n >= 0, SHEMS_ERR_INVALID, "replay_push: n=%lld", (long long)n);
  if (n == 0) return SHEMS_OK;
  GUARD(rp->device);
  // only the newest `cap` of the n transitions can survive; skipping the rest also keeps slots single-writer
  const int64_t first = n > rp->capacity ? n - rp->capacity : 0;
  const int64_t cnt = n - first;
  replay_push_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, rp->stream>>>(rp->ring, rp->capacity, rp->head, s_dev, a_dev, r_dev, s2_dev,
                                                                           done_dev, n, first);
  CUDA_TRY(cudaGetLastError());
  return replay_after_rollout(rp, n);
}

// remember() for a population: arrays are SoA over all N = P*n_per instances, learner l = instances l*n_per .. l*n_per+n_per-1
// pushes its slice into its own ring (blockIdx.y = learner)
struct RingRef { float* ring; long long cap, head; };
__global__ void __launch_bounds__(256)
replay_push_groups_kernel(const RingRef* __restrict__ refs, const float* __restrict__ s, const float* __restrict__ a, const float* __restrict__ r,
                          const float* __restrict__ s2, const float* __restrict__ done, long long n_per, long long N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per) return;
  const RingRef ref = refs[blockIdx.y];
  const long long g = (long long)blockIdx.y * n_per + i;  // global instance
  long long slot = ref.head + i;
  slot -= (slot / ref.cap) * ref.cap;
  float* q = ref.ring + ring_base(slot);
#pragma unroll
  for (int k = 0; k < 9; ++k) q[(RING_S + k) * 32] = s[k * N + g];
  q[(RING_A + 0) * 32] = a[g];
  q[(RING_A + 1) * 32] = a[N + g];
  q[RING_R * 32] = r[g];
#pragma unroll
  for (int k = 0; k < 9; ++k) q[(RING_S2 + k) * 32] = s2[k * N + g];
  q[RING_DONE * 32] = done ? done[g] : 0.0f;
}

extern "C" int32_t replay_push_groups(ShemsReplay* const* rps, int32_t n_groups, const float* s_dev, const float* a_dev, const float* r_dev,
                                      const float* s2_dev, const float* done_dev, int64_t n_per) {
  REQUIRE(rps && s_dev && a_dev && r_dev && s2_dev, SHEMS_ERR_INVALID, "replay_push_groups: NULL argument");
  REQUIRE(n_groups >= 1 && n_per >= 1, SHEMS_ERR_INVALID, "replay_push_groups: n_groups=%d n_per=%lld", n_groups, (long long)n_per);
  ShemsReplay* r0 = rps[0];
  REQUIRE(r0, SHEMS_ERR_INVALID, "replay_push_groups: rps[0] is NULL");
  GUARD(r0->device);
  std::vector<RingRef> refs((size_t)n_groups);
  for (int g = 0; g < n_groups; ++g) {
    ShemsReplay* rp = rps[g];
    REQUIRE(rp && rp->device == r0->device, SHEMS_ERR_INVALID, "replay_push_groups: rps[%d] missing or on another device", g);
    REQUIRE(n_per <= rp->capacity, SHEMS_ERR_INVALID, "replay_push_groups: n_per=%lld exceeds the capacity %lld of rps[%d]", (long long)n_per,
            (long long)rp->capacity, g);
    refs[g].ring = rp->ring; refs[g].cap = rp->capacity; refs[g].head = rp->head;
  }
  if (r0->group_refs_n < n_groups) {
    CUDA_TRY(cudaStreamSynchronize(r0->stream));
    cudaFree(r0->group_refs); r0->group_refs = nullptr; r0->group_refs_n = 0;
    CUDA_TRY(cudaMalloc(&r0->group_refs, sizeof(RingRef) * (size_t)n_groups));
    r0->group_refs_n = n_groups;
  }
  CUDA_TRY(cudaMemcpyAsync(r0->group_refs, refs.data(), sizeof(RingRef) * (size_t)n_groups, cudaMemcpyHostToDevice, r0->stream));
  replay_push_groups_kernel<<<dim3((unsigned)((n_per + 255) / 256), n_groups), 256, 0, r0->stream>>>((const RingRef*)r0->group_refs, s_dev, a_dev,
                                                                                                r_dev, s2_dev, done_dev, n_per,
                                                                                                n_per * n_groups);
  CUDA_TRY(cudaGetLastError());
  for (int g = 0; g < n_groups; ++g) replay_after_rollout(rps[g], n_per);
  return SHEMS_OK;
}

// getData(): gather B sampled transitions into SoA minibatch arrays [k][B].
// idx != NULL: logical indices; else Philox(seed, id = draw j, ctr = update) -> floor(u * len).
__device__ __forceinline__ long long replay_slot(long long logical, long long head, long long len, long long cap) {
  long long slot = head - len + logical;
  if (slot < 0) slot += cap;
  return slot;
}
__global__ void __launch_bounds__(128)
replay_sample_kernel(const float* __restrict__ ring, long long cap, long long head, long long len, const int32_t* __restrict__ idx,
                     unsigned long long seed, unsigned update, int B, float* __restrict__ s, float* __restrict__ a, float* __restrict__ r,
                     float* __restrict__ s2, float* __restrict__ done) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= B) return;
  long long li;
  if (idx) li = idx[j];
  else {
    uint32_t w[4];
    philox4x32_10(seed, (uint64_t)j, update, STREAM_SAMPLE, w);
    li = (long long)(u53(w[0], w[1]) * (double)len);
    if (li >= len) li = len - 1;
  }
  const float* q = ring + ring_base(replay_slot(li, head, len, cap));
#pragma unroll
  for (int k = 0; k < 9; ++k) s[k * B + j] = q[(RING_S + k) * 32];
  a[j] = q[(RING_A + 0) * 32];
  a[B + j] = q[(RING_A + 1) * 32];
  r[j] = q[RING_R * 32];
#pragma unroll
  for (int k = 0; k < 9; ++k) s2[k * B + j] = q[(RING_S2 + k) * 32];
  done[j] = q[RING_DONE * 32];
}

static int ensure_idx_scratch(ShemsReplay* rp, int64_t n) {
  if (rp->idx_scratch_n >= n) return SHEMS_OK;
  cudaFree(rp->idx_scratch);
  rp->idx_scratch = nullptr; rp->idx_scratch_n = 0;
  CUDA_TRY(cudaMalloc(&rp->idx_scratch, sizeof(int32_t) * (size_t)n));
  rp->idx_scratch_n = n;
  return SHEMS_OK;
}

extern "C" int32_t replay_sample(ShemsReplay* rp, int32_t batch, const int32_t* idx_host, uint64_t seed, float* s_dev, float* a_dev,
                                 float* r_dev, float* s2_dev, float* done_dev) {
  REQUIRE(rp && s_dev && a_dev && r_dev && s2_dev && done_dev, SHEMS_ERR_INVALID, "replay_sample: NULL argument");
  REQUIRE(batch >= 1, SHEMS_ERR_INVALID, "replay_sample: batch=%d", batch);
  REQUIRE(rp->length > 0, SHEMS_ERR_STATE, "replay_sample: memory is empty (sample from an empty collection)");
  GUARD(rp->device);
  const int32_t* didx = nullptr;
  if (idx_host) {
    for (int j = 0; j < batch; ++j)
      REQUIRE(idx_host[j] >= 0 && idx_host[j] < rp->length, SHEMS_ERR_INVALID, "replay_sample: idx[%d]=%d outside 0..%lld", j, idx_host[j],
              (long long)rp->length - 1);
    int st = ensure_idx_scratch(rp, batch);
    if (st) return st;
    CUDA_TRY(cudaMemcpyAsync(rp->idx_scratch, idx_host, sizeof(int32_t) * batch, cudaMemcpyHostToDevice, rp->stream));
    didx = rp->idx_scratch;
  }
  replay_sample_kernel<<<(batch + 127) / 128, 128, 0, rp->stream>>>(rp->ring, rp->capacity, rp->head, rp->length, didx, seed, 0u, batch,
                                                                   s_dev, a_dev, r_dev, s2_dev, done_dev);
  CUDA_TRY(cudaGetLastError());
  if (idx_host) CUDA_TRY(cudaStreamSynchronize(rp->stream));  // idx_host was staged asynchronously
  return SHEMS_OK;
}

// min_max_buffer(): min/max of the 9 state fields over a with-replacement sample of n draws.
// Float min/max are exact and order-independent -> atomics on the int-ordered bit pattern.
__device__ __forceinline__ void atomic_minf(float* addr, float v) {
  if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_maxf(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__global__ void __launch_bounds__(256)
replay_minmax_kernel(const float* __restrict__ ring, long long cap, long long head, long long len, const int32_t* __restrict__ idx,
                     unsigned long long seed, long long n, float* __restrict__ out /* [0..8]=min, [9..17]=max */) {
  float mn[9], mx[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { mn[k] = INFINITY; mx[k] = -INFINITY; }
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    long long li;
    if (idx) li = idx[j];
    else {
      uint32_t w[4];
      philox4x32_10(seed, (uint64_t)j, 0u, STREAM_SAMPLE, w);
      li = (long long)(u53(w[0], w[1]) * (double)len);
      if (li >= len) li = len - 1;
    }
    const float* q = ring + ring_base(replay_slot(li, head, len, cap));
#pragma unroll
    for (int k = 0; k < 9; ++k) { const float v = q[(RING_S + k) * 32]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
    if ((threadIdx.x & 31) == 0) { atomic_minf(out + k, mn[k]); atomic_maxf(out + 9 + k, mx[k]); }
  }
}

extern "C" int32_t replay_minmax(ShemsReplay* rp, int64_t n_samples, const int32_t* idx_host, uint64_t seed, float* s_min_host, float* s_max_host) {
  REQUIRE(rp && s_min_host && s_max_host, SHEMS_ERR_INVALID, "replay_minmax: NULL argument");
  REQUIRE(n_samples >= 1, SHEMS_ERR_INVALID, "replay_minmax: n_samples=%lld", (long long)n_samples);
  REQUIRE(rp->length > 0, SHEMS_ERR_STATE, "replay_minmax: memory is empty");
  GUARD(rp->device);
  const int32_t* didx = nullptr;
  if (idx_host) {
    for (int64_t j = 0; j < n_samples; ++j)
      REQUIRE(idx_host[j] >= 0 && idx_host[j] < rp->length, SHEMS_ERR_INVALID, "replay_minmax: idx[%lld]=%d outside 0..%lld", (long long)j,
              idx_host[j], (long long)rp->length - 1);
    int st = ensure_idx_scratch(rp, n_samples);
    if (st) return st;
    CUDA_TRY(cudaMemcpyAsync(rp->idx_scratch, idx_host, sizeof(int32_t) * (size_t)n_samples, cudaMemcpyHostToDevice, rp->stream));
    didx = rp->idx_scratch;
  }
  float init[18];
  for (int k = 0; k < 9; ++k) { init[k] = INFINITY; init[9 + k] = -INFINITY; }
  CUDA_TRY(cudaMemcpyAsync(rp->minmax_scratch, init, sizeof(init), cudaMemcpyHostToDevice, rp->stream));
  long long blocks = (n_samples + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  replay_minmax_kernel<<<(unsigned)blocks, 256, 0, rp->stream>>>(rp->ring, rp->capacity, rp->head, rp->length, didx, seed, n_samples, rp->minmax_scratch);
  CUDA_TRY(cudaGetLastError());
  float res[18];
  CUDA_TRY(cudaMemcpyAsync(res, rp->minmax_scratch, sizeof(res), cudaMemcpyDeviceToHost, rp->stream));
  CUDA_TRY(cudaStreamSynchronize(rp->stream));
  memcpy(s_min_host, res, sizeof(float) * 9);
  memcpy(s_max_host, res + 9, sizeof(float) * 9);
  return SHEMS_OK;
}

extern "C" int32_t replay_get(ShemsReplay* rp, float* s_host, float* a_host, float* r_host, float* s2_host, float* done_host) {
  REQUIRE(rp, SHEMS_ERR_INVALID, "replay_get: NULL handle");
  GUARD(rp->device);
  const int64_t len = rp->length, cap = rp->capacity;
  if (len == 0) return SHEMS_OK;
  CUDA_TRY(cudaStreamSynchronize(rp->stream));
  const size_t tiles = ((size_t)cap + 31) / 32;
  std::vector<float> host(RING_FIELDS * 32 * tiles);
  CUDA_TRY(cudaMemcpy(host.data(), rp->ring, sizeof(float) * host.size(), cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < len; ++i) {  // logical order, oldest first
    int64_t slot = rp->head - len + i;
    if (slot < 0) slot += cap;
    const float* q = host.data() + ring_base(slot);
    for (int k = 0; k < 9; ++k) {
      if (s_host) s_host[(size_t)k * len + i] = q[(RING_S + k) * 32];
      if (s2_host) s2_host[(size_t)k * len + i] = q[(RING_S2 + k) * 32];
    }
    if (a_host) { a_host[i] = q[(RING_A + 0) * 32]; a_host[(size_t)len + i] = q[(RING_A + 1) * 32]; }
    if (r_host) r_host[i] = q[RING_R * 32];
    if (done_host) done_host[i] = q[RING_DONE * 32];
  }
  return SHEMS_OK;
}
