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
  int M, N, K, splits, epi, tma_store, batch, nprob, pdl;
  float* D[TC_MAX_PROBLEMS]; long long ldd, split_stride;
  const float* bias[TC_MAX_PROBLEMS]; const float* aux[TC_MAX_PROBLEMS]; long long auxld, bs_bias, bs_aux;
};
// `count` independent products of one shape in a launch (a population of learners): strides in floats between entries
struct TcBatch {
  int count = 1;
  long long sA = 0, sB = 0, sD = 0, sBias = 0, sAux = 0;
  bool pdl = false;   // launch as a programmatic dependent of the predecessor in the stream (the kernel waits before touching its operands)
};
// D[M×N] (row-major, ldd) = epilogue(sum_k A(m,k) · B(n,k))
int tc_gemm(cudaStream_t st, const TcOperand& A, const TcOperand& B, float* D, long long ldd, int M, int N, int K, int epi,
            const float* bias, const float* aux, long long auxld, int splits, float* workspace, const TcBatch& batch = TcBatch());
// up to TC_MAX_PROBLEMS products of one shape in a single launch (e.g. the layer-2 forward passes of actor_target / critic / actor)
int tc_gemm_multi(cudaStream_t st, int nprob, const TcOperand* A, const TcOperand* B, float* const* D, long long ldd, int M, int N, int K, int epi,
                  const float* const* bias, const float* const* aux, long long auxld, int splits, float* workspace, const TcBatch& batch = TcBatch());
// sets the kernels' shared-memory attribute on the current device; call once per device before capturing launches in a graph
int tc_gemm_prepare();

// The whole forward pass of up to TC_MAX_PROBLEMS nets (same widths l1 <= 256, l2 <= 512) over a large minibatch in ONE launch:
// layer 1 (fp32 SIMT) -> layer 2 (tcgen05, TF32) -> output layer (1 or 2 units) in the epilogue; 128 rows per CTA.
enum { TC_OUT_TANH = 0, TC_OUT_ID = 1, TC_OUT_TD = 2 };
struct TcFwdChainArgs {
  int M, L1, L2, nprob, ldx, ldh1, ldh2;
  const float* X[TC_MAX_PROBLEMS]; int K1[TC_MAX_PROBLEMS];        // inputs [M][ldx] (K1 <= 12 columns used)
  const float *W1[TC_MAX_PROBLEMS], *b1[TC_MAX_PROBLEMS];          // Flux layout Wt[in][out]
  const float *W2[TC_MAX_PROBLEMS], *b2[TC_MAX_PROBLEMS];
  const float *W3[TC_MAX_PROBLEMS], *b3[TC_MAX_PROBLEMS]; int J[TC_MAX_PROBLEMS];
  float* H1[TC_MAX_PROBLEMS];                                      // [M][ldh1] layer-1 activations, or NULL (target nets: no backward pass)
  float* H2[TC_MAX_PROBLEMS];                                      // [M][ldh2] layer-2 activations, or NULL
  int out_mode[TC_MAX_PROBLEMS]; float* out[TC_MAX_PROBLEMS]; int ldo[TC_MAX_PROBLEMS];   // out[row*ldo + j]
  const float *td_r, *td_done, *td_q; float* td_dq; float gamma, inv_batch;               // TC_OUT_TD
  int pdl, early_weights;          // programmatic dependent launch; early_weights: the predecessor in the stream did not write the net's weights
  int pop; long long pop_stride;   // a population of learners (grid.y): every pointer above moves by pop_stride floats per learner
};
int tc_fwd_chain(cudaStream_t st, const TcFwdChainArgs& a);
