// Dense projections of the GCN hot path: C = act(op(A).op(B) + beta*C + bias) [* act'(mask)].
//
// Replaces T.dot (lasagne_layers.py:82) and the two gradient products theano.grad
// derives from it (dW = H^T.dZ, dH = dZ.W^T; mlpconv.py:263).
//
// GCG_GEMM_FMA: fp32 FFMA tiles, 128x128x16 CTA tile, 8x8 per thread, register
// double buffering.  Deterministic split-K (fixed slice order) for the tall-skinny
// weight-gradient products (K = number of graph nodes).
// GCG_GEMM_TF32X3 / TF32: tcgen05 path, see gcg_gemm_tc.cu.
#include <algorithm>

#include "gcg_common.cuh"
#include "gcg_gemm.cuh"

namespace gcg {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

// 4 consecutive elements along the contiguous dimension, zero padded.
__device__ __forceinline__ float4 load4(const float* __restrict__ base, int64_t outer, int64_t inner,
                                        int64_t n_outer, int64_t n_inner, int64_t ld, int vec) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (outer >= n_outer || inner >= n_inner) return r;
  const float* p = base + outer * ld + inner;
  if (vec && inner + 3 < n_inner) return __ldg(reinterpret_cast<const float4*>(p));
  r.x = __ldg(p);
  if (inner + 1 < n_inner) r.y = __ldg(p + 1);
  if (inner + 2 < n_inner) r.z = __ldg(p + 2);
  if (inner + 3 < n_inner) r.w = __ldg(p + 3);
  return r;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 2) gemm_fma_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int mt = blockIdx.x / g.n_tiles_n, nt = blockIdx.x - mt * g.n_tiles_n;
  const int64_t m0 = (int64_t)mt * BM, n0 = (int64_t)nt * BN;
  const int64_t kb = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t ke = min(g.K, kb + g.k_per_split);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!TA) {  // A stored [M][K]: row = m, inner = k
        const int row = (tid >> 2) + i * 64, kq = (tid & 3) * 4;
        ra[i] = load4(g.A, m0 + row, k0 + kq, g.M, ke, g.lda, g.vecA);
      } else {    // A stored [K][M]: row = k, inner = m
        const int k = (tid >> 5) + i * 8, m4 = (tid & 31) * 4;
        ra[i] = load4(g.A, k0 + k, m0 + m4, ke, g.M, g.lda, g.vecA);
      }
      if (!TB) {  // B stored [K][N]
        const int k = (tid >> 5) + i * 8, n4 = (tid & 31) * 4;
        rb[i] = load4(g.B, k0 + k, n0 + n4, ke, g.N, g.ldb, g.vecB);
      } else {    // B stored [N][K]
        const int row = (tid >> 2) + i * 64, kq = (tid & 3) * 4;
        rb[i] = load4(g.B, n0 + row, k0 + kq, g.N, ke, g.ldb, g.vecB);
      }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!TA) {
        const int row = (tid >> 2) + i * 64, kq = (tid & 3) * 4;
        As[buf][kq + 0][row] = ra[i].x; As[buf][kq + 1][row] = ra[i].y;
        As[buf][kq + 2][row] = ra[i].z; As[buf][kq + 3][row] = ra[i].w;
      } else {
        const int k = (tid >> 5) + i * 8, m4 = (tid & 31) * 4;
        *reinterpret_cast<float4*>(&As[buf][k][m4]) = ra[i];
      }
      if (!TB) {
        const int k = (tid >> 5) + i * 8, n4 = (tid & 31) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][k][n4]) = rb[i];
      } else {
        const int row = (tid >> 2) + i * 64, kq = (tid & 3) * 4;
        Bs[buf][kq + 0][row] = rb[i].x; Bs[buf][kq + 1][row] = rb[i].y;
        Bs[buf][kq + 2][row] = rb[i].z; Bs[buf][kq + 3][row] = rb[i].w;
      }
    }
  };

  int buf = 0;
  if (kb < ke) {
    fetch(kb);
    stash(0);
  }
  __syncthreads();
  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
    const bool more = (k0 + BK) < ke;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int64_t n = n0 + jh * 64 + tx * 4;
      if (n >= g.N) continue;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if (g.split_k > 1) {
        float* dst = g.part + ((int64_t)blockIdx.z * g.M + m) * g.N + n;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < g.N) dst[e] = v[e];
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < g.N) v[e] = gemm_epilogue(g, v[e], m, n + e);
        float* dst = g.C + m * g.ldc + n;
        if (g.vecC && n + 3 < g.N) {
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < g.N) dst[e] = v[e];
        }
      }
    }
  }
}

// C = epilogue( sum_z part[z] ) in slice order z = 0..split-1 (deterministic).
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(const GemmArgs g) {
  const int64_t total = g.M * g.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < g.split_k; ++z) s += g.part[(int64_t)z * total + i];
    const int64_t m = i / g.N, n = i - m * g.N;
    g.C[m * g.ldc + n] = gemm_epilogue(g, s, m, n);
  }
}

static int auto_split(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ceil_div(M, BM) * ceil_div(N, BN);
  if (tiles >= 2 * kNumSMs || K < 4096) return 1;
  int64_t s = ceil_div(4 * (int64_t)kNumSMs, tiles);
  s = std::min<int64_t>(s, K / (16 * BK));
  return (int)std::max<int64_t>(1, std::min<int64_t>(s, 256));
}

int launch_splitk_reduce(const GemmArgs& g, cudaStream_t st) {
  const unsigned rg = (unsigned)std::min<int64_t>(ceil_div(g.M * g.N, 256), (int64_t)kNumSMs * 16);
  gemm_splitk_reduce_kernel<<<rg, 256, 0, st>>>(g);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

}  // namespace gcg

using namespace gcg;

extern "C" int64_t gcg_gemm_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K,
                                            int mode, int32_t split_k) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int64_t fma_bytes = 0;
  {
    int s = split_k <= 0 ? auto_split(M, N, K) : split_k;
    if (s > 1) fma_bytes = (int64_t)s * M * N * (int64_t)sizeof(float);
  }
  if (mode == GCG_GEMM_FMA) return fma_bytes;
  return std::max(fma_bytes, gemm_tc_workspace_bytes(transA, transB, M, N, K, mode, split_k));
}

static int gemm_impl(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                     const float* B, int64_t ldb, float* C, int64_t ldc, float beta, const float* bias, int act,
                     const float* mask, int64_t ld_mask, int mask_act, int mode, int32_t split_k, void* workspace,
                     int64_t workspace_bytes, void* stream, const float* A_hi, const float* A_lo,
                     const float* B_hi, const float* B_lo);

extern "C" int gcg_gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K,
                            const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                            int64_t ldc, float beta, const float* bias, int act, const float* mask,
                            int64_t ld_mask, int mask_act, int mode, int32_t split_k,
                            void* workspace, int64_t workspace_bytes, void* stream) {
  GCG_RECORD("gcg_gemm_f32", gcg_gemm_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, beta, bias, act, mask, ld_mask, mask_act, mode, split_k, workspace, workspace_bytes, s__));
  return gemm_impl(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, beta, bias, act, mask, ld_mask, mask_act, mode,
                   split_k, workspace, workspace_bytes, stream, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int gcg_gemm_presplit_f32(int transA, int transB, int64_t M, int64_t N, int64_t K,
                                     const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                                     int64_t ldc, float beta, const float* bias, int act, const float* mask,
                                     int64_t ld_mask, int mask_act, int mode, int32_t split_k,
                                     void* workspace, int64_t workspace_bytes, void* stream,
                                     const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo) {
  GCG_RECORD("gcg_gemm_presplit_f32", gcg_gemm_presplit_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, beta, bias, act, mask, ld_mask, mask_act, mode, split_k, workspace, workspace_bytes, s__, A_hi, A_lo, B_hi, B_lo));
  GCG_CHECK_ARG((A_hi == nullptr) == (A_lo == nullptr) && (B_hi == nullptr) == (B_lo == nullptr),
                "gcg_gemm_presplit_f32: hi and lo go together");
  return gemm_impl(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, beta, bias, act, mask, ld_mask, mask_act, mode,
                   split_k, workspace, workspace_bytes, stream, A_hi, A_lo, B_hi, B_lo);
}

static int gemm_impl(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                     const float* B, int64_t ldb, float* C, int64_t ldc, float beta, const float* bias, int act,
                     const float* mask, int64_t ld_mask, int mask_act, int mode, int32_t split_k, void* workspace,
                     int64_t workspace_bytes, void* stream, const float* A_hi, const float* A_lo,
                     const float* B_hi, const float* B_lo) {
  GCG_CHECK_ARG(A && B && C, "gcg_gemm_f32: NULL operand");
  GCG_CHECK_SHAPE(M >= 0 && N >= 0 && K >= 0, "gcg_gemm_f32: negative dimension");
  GCG_CHECK_SHAPE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N,
                  "gcg_gemm_f32: leading dimension too small (lda=%lld ldb=%lld ldc=%lld)",
                  (long long)lda, (long long)ldb, (long long)ldc);
  GCG_CHECK_ARG(act >= GCG_ACT_IDENTITY && act <= GCG_ACT_SIGMOID, "gcg_gemm_f32: bad act %d", act);
  GCG_CHECK_ARG(!mask || ld_mask >= N, "gcg_gemm_f32: ld_mask too small");
  GCG_CHECK_ARG(mode == GCG_GEMM_FMA || mode == GCG_GEMM_TF32X3 || mode == GCG_GEMM_TF32 || mode == GCG_GEMM_TF32X3_CHAINED,
                "gcg_gemm_f32: unknown mode %d", mode);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M == 0 || N == 0) return GCG_OK;

  GemmArgs g;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.beta = beta; g.bias = bias; g.act = act;
  g.mask = mask; g.ld_mask = ld_mask; g.mask_act = mask_act;
  g.n_tiles_n = (int)ceil_div(N, BN);
  g.vecA = aligned16(A) && lda % 4 == 0;
  g.vecB = aligned16(B) && ldb % 4 == 0;
  g.vecC = aligned16(C) && ldc % 4 == 0;
  g.A_hi = A_hi; g.A_lo = A_lo; g.B_hi = B_hi; g.B_lo = B_lo;
  const int32_t split_req = split_k;
  if (mode != GCG_GEMM_FMA && K > 0) {
    g.split_k = split_req > 0 ? split_req : gemm_tc_auto_split(M, N, K);
    g.k_per_split = 0;
    g.part = nullptr;
    const int rc = gemm_tc_launch(g, transA, transB, mode, workspace, workspace_bytes, st);
    if (rc != GCG_ERR_UNSUPPORTED) return rc;     // ran (or failed hard) on the tensor cores
  }
  if (split_k <= 0) split_k = auto_split(M, N, K);
  if (K == 0) split_k = 1;
  if (split_k > 1 && (!workspace || workspace_bytes < (int64_t)split_k * M * N * (int64_t)sizeof(float)) &&
      mode != GCG_GEMM_FMA)
    split_k = 1;                                  // FFMA fallback of a tensor-core request: no spare workspace
  g.split_k = split_k;
  g.k_per_split = std::max<int64_t>(BK, ceil_div(ceil_div(std::max<int64_t>(K, 1), split_k), BK) * BK);
  g.split_k = (int)std::max<int64_t>(1, ceil_div(std::max<int64_t>(K, 1), g.k_per_split));
  g.part = reinterpret_cast<float*>(workspace);
  if (g.split_k > 1)
    GCG_CHECK_ARG(workspace && workspace_bytes >= (int64_t)g.split_k * M * N * (int64_t)sizeof(float),
                  "gcg_gemm_f32: split-K workspace too small");
  const int64_t tiles = ceil_div(M, BM) * g.n_tiles_n;
  GCG_CHECK_SHAPE(tiles < INT32_MAX, "gcg_gemm_f32: grid too large");
  dim3 grid((unsigned)tiles, 1, (unsigned)g.split_k);
  if (!transA && !transB) gemm_fma_kernel<false, false><<<grid, 256, 0, st>>>(g);
  else if (transA && !transB) gemm_fma_kernel<true, false><<<grid, 256, 0, st>>>(g);
  else if (!transA && transB) gemm_fma_kernel<false, true><<<grid, 256, 0, st>>>(g);
  else gemm_fma_kernel<true, true><<<grid, 256, 0, st>>>(g);
  GCG_LAUNCH_CHECK();
  if (g.split_k > 1) {
    const unsigned rg = (unsigned)std::min<int64_t>(ceil_div(M * N, 256), (int64_t)kNumSMs * 16);
    gemm_splitk_reduce_kernel<<<rg, 256, 0, st>>>(g);
    GCG_LAUNCH_CHECK();
  }
  return GCG_OK;
}
