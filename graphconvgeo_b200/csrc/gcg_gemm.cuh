// Shared between the FFMA (gcg_gemm.cu) and tcgen05 (gcg_gemm_tc.cu) dense kernels.
#pragma once
#include "gcg_common.cuh"

namespace gcg {

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int64_t M, N, K;
  float beta;
  const float* bias; int act;
  const float* mask; int64_t ld_mask; int mask_act;
  int n_tiles_n;
  int split_k; int64_t k_per_split;
  float* part;  // split-K partials [split][M][N]
  int vecA, vecB, vecC;
  // optional pre-split tf32 operands (same leading dimensions as A / B); NULL = split inside the call
  const float *A_hi, *A_lo, *B_hi, *B_lo;
};


__device__ __forceinline__ float gemm_epilogue(const GemmArgs& g, float v, int64_t m, int64_t n) {
  if (g.beta != 0.f) v += g.beta * g.C[m * g.ldc + n];
  if (g.bias) v += __ldg(g.bias + n);
  v = apply_act(v, g.act);
  if (g.mask) v *= act_grad_from_out(g.mask[m * g.ld_mask + n], g.mask_act);
  return v;
}


// tcgen05 path (gcg_gemm_tc.cu).  Returns GCG_ERR_UNSUPPORTED when the operands cannot be
// described by TMA tensor maps (unaligned base / leading dimension), in which case the
// caller falls back to the FFMA tiles.
int gemm_tc_launch(const GemmArgs& g, int transA, int transB, int mode, void* workspace,
                   int64_t workspace_bytes, cudaStream_t st);
int64_t gemm_tc_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K, int mode,
                                int split_k);
int gemm_tc_auto_split(int64_t M, int64_t N, int64_t K);
int tf32_split_launch(const float* x, int64_t n_floats, float* hi, float* lo, cudaStream_t st);

}  // namespace gcg
