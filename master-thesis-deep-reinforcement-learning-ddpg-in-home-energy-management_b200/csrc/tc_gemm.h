// tc_gemm.h — TF32 tcgen05 GEMM used by the large-batch DDPG update (csrc/tc_gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

enum { TC_EPI_NONE = 0, TC_EPI_BIAS_RELU = 1, TC_EPI_RELU_MASK = 2 };

struct TcOperand {
  const float* ptr;
  long long ld;      // floats between consecutive rows (K-major) or consecutive k (MN-major)
  bool mn_major;     // false: (row, k) at ptr[row*ld + k]; true: (row, k) at ptr[k*ld + row]
};
#define TC_MAX_PROBLEMS 3
struct TcGemmArgs {
  int M, N, K, splits, epi, tma_store, batch, nprob;
  float* D[TC_MAX_PROBLEMS]; long long ldd, split_stride;
  const float* bias[TC_MAX_PROBLEMS]; const float* aux[TC_MAX_PROBLEMS]; long long auxld, bs_bias, bs_aux;
};
// `count` independent products of one shape in a launch (a population of learners): strides in floats between entries
struct TcBatch {
  int count = 1;
  long long sA = 0, sB = 0, sD = 0, sBias = 0, sAux = 0;
};
// D[M×N] (row-major, ldd) = epilogue(sum_k A(m,k) · B(n,k))
int tc_gemm(cudaStream_t st, const TcOperand& A, const TcOperand& B, float* D, long long ldd, int M, int N, int K, int epi,
            const float* bias, const float* aux, long long auxld, int splits, float* workspace, const TcBatch& batch = TcBatch());
// up to TC_MAX_PROBLEMS products of one shape in a single launch (e.g. the layer-2 forward passes of actor_target / critic / actor)
int tc_gemm_multi(cudaStream_t st, int nprob, const TcOperand* A, const TcOperand* B, float* const* D, long long ldd, int M, int N, int K, int epi,
                  const float* const* bias, const float* const* aux, long long auxld, int splits, float* workspace, const TcBatch& batch = TcBatch());
// sets the kernels' shared-memory attribute on the current device; call once per device before capturing launches in a graph
int tc_gemm_prepare();
