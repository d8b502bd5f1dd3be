// Epilogue / reduction / head / optimiser kernels of the GCN hot path (sm_100a).
// All are HBM-bound streaming kernels: 128-bit accesses when the layout allows
// it, grids sized in multiples of the SM count, deterministic reductions.
#include <algorithm>

#include <vector>

#include "gcg_common.cuh"

namespace gcg {

template <int VEC> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<1> { using T = float; };

__device__ __forceinline__ float4 ld(const float4* p) { return *p; }
__device__ __forceinline__ float ld(const float* p) { return *p; }

// true when every (ptr, ld) pair allows float4 access over round_up(F,4) columns
static bool vec_ok(int64_t F, std::initializer_list<std::pair<const void*, int64_t>> ops) {
  const int64_t fp = (F + 3) / 4 * 4;
  for (auto& o : ops) {
    if (o.first == nullptr) continue;
    if (!aligned16(o.first) || (o.second % 4) != 0 || o.second < fp) return false;
  }
  return true;
}

static inline unsigned grid_for(int64_t total, int threads) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, threads), (int64_t)kNumSMs * 32));
}

// ------------------------------------------------------------------ colsum
// pass 1: block (cx, ry) sums rows ry, ry+gridDim.y, ... of a 32*VEC-column strip
template <int VEC>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, int64_t ld,
                                                             int64_t n_rows, int64_t F,
                                                             float* __restrict__ part) {
  __shared__ float sm[8][32 * VEC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c0 = ((int64_t)blockIdx.x * 32 + lane) * VEC;
  const int64_t rows_per = ceil_div(n_rows, gridDim.y);
  const int64_t r0 = (int64_t)blockIdx.y * rows_per;
  const int64_t r1 = min(n_rows, r0 + rows_per);
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  if (c0 < F) {
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      if (VEC == 4) {
        const float4 v = *reinterpret_cast<const float4*>(X + r * ld + c0);
        acc[0] += v.x; acc[VEC > 1 ? 1 : 0] += v.y; acc[VEC > 2 ? 2 : 0] += v.z; acc[VEC > 3 ? 3 : 0] += v.w;
      } else {
        acc[0] += X[r * ld + c0];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) sm[warp][lane * VEC + e] = acc[e];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sm[w][lane * VEC + e];
      if (c0 + e < F) part[(int64_t)blockIdx.y * F + c0 + e] = s;
    }
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int n_parts, int64_t F,
                                    float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= F) return;
  float s = 0.f;
  for (int p = 0; p < n_parts; ++p) s += part[(int64_t)p * F + c];
  out[c] = s;
}

static int colsum_parts(int64_t n_rows, int64_t F) {
  const int64_t strips = ceil_div(F, 128);
  int64_t parts = std::max<int64_t>(1, (int64_t)kNumSMs * 8 / std::max<int64_t>(1, strips));
  parts = std::min<int64_t>(parts, ceil_div(n_rows, 64));
  return (int)std::max<int64_t>(1, std::min<int64_t>(parts, 4096));
}

// ------------------------------------------------------------------ act_bwd
template <int VEC>
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dA, int64_t ld_da,
                                                      const float* __restrict__ A, int64_t ld_a,
                                                      float* __restrict__ dP, int64_t ld_dp,
                                                      int64_t n_rows, int64_t W, int act) {
  const int64_t total = n_rows * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / W, c = (i - r * W) * VEC;
    if (VEC == 4) {
      const float4 d = *reinterpret_cast<const float4*>(dA + r * ld_da + c);
      const float4 a = *reinterpret_cast<const float4*>(A + r * ld_a + c);
      float4 o;
      o.x = d.x * act_grad_from_out(a.x, act); o.y = d.y * act_grad_from_out(a.y, act);
      o.z = d.z * act_grad_from_out(a.z, act); o.w = d.w * act_grad_from_out(a.w, act);
      *reinterpret_cast<float4*>(dP + r * ld_dp + c) = o;
    } else {
      dP[r * ld_dp + c] = dA[r * ld_da + c] * act_grad_from_out(A[r * ld_a + c], act);
    }
  }
}

// --------------------------------------------------------------- highway_bwd
__device__ __forceinline__ void highway_bwd_one(float d, float g, float hc, float hin, int act,
                                                float& dp, float& dg, float& dh) {
  dp = g * d * act_grad_from_out(hc, act);
  dg = d * (hc - hin) * g * (1.f - g);
  dh = (1.f - g) * d;
}
template <int VEC>
__global__ void __launch_bounds__(256) highway_bwd_kernel(
    const float* __restrict__ dO, int64_t ld_do, const float* __restrict__ g, int64_t ld_g,
    const float* __restrict__ Hc, int64_t ld_hc, const float* __restrict__ Hin, int64_t ld_hin,
    float* dP, int64_t ld_dp, float* dG, int64_t ld_dg, float* dH, int64_t ld_dh, int64_t n_rows,
    int64_t W, int act) {
  const int64_t total = n_rows * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / W, c = (i - r * W) * VEC;
    if (VEC == 4) {
      const float4 d = *reinterpret_cast<const float4*>(dO + r * ld_do + c);
      const float4 gg = *reinterpret_cast<const float4*>(g + r * ld_g + c);
      const float4 hc = *reinterpret_cast<const float4*>(Hc + r * ld_hc + c);
      const float4 hi = *reinterpret_cast<const float4*>(Hin + r * ld_hin + c);
      float4 p, q, h;
      highway_bwd_one(d.x, gg.x, hc.x, hi.x, act, p.x, q.x, h.x);
      highway_bwd_one(d.y, gg.y, hc.y, hi.y, act, p.y, q.y, h.y);
      highway_bwd_one(d.z, gg.z, hc.z, hi.z, act, p.z, q.z, h.z);
      highway_bwd_one(d.w, gg.w, hc.w, hi.w, act, p.w, q.w, h.w);
      *reinterpret_cast<float4*>(dP + r * ld_dp + c) = p;
      *reinterpret_cast<float4*>(dG + r * ld_dg + c) = q;
      *reinterpret_cast<float4*>(dH + r * ld_dh + c) = h;
    } else {
      float p, q, h;
      highway_bwd_one(dO[r * ld_do + c], g[r * ld_g + c], Hc[r * ld_hc + c], Hin[r * ld_hin + c], act, p, q, h);
      dP[r * ld_dp + c] = p; dG[r * ld_dg + c] = q; dH[r * ld_dh + c] = h;
    }
  }
}

// --------------------------------------------------------------- highway_fwd
// O = g*Hc + (1-g)*Hin with separately rounded operations (same arithmetic as the fused SpMM epilogue)
template <int VEC>
__global__ void __launch_bounds__(256) highway_fwd_kernel(const float* __restrict__ Hc, int64_t ld_hc,
                                                          const float* __restrict__ g, int64_t ld_g,
                                                          const float* __restrict__ Hin, int64_t ld_hin,
                                                          float* O, int64_t ld_o, int64_t n_rows, int64_t W) {
  const int64_t total = n_rows * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / W, c = (i - r * W) * VEC;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float gg = g[r * ld_g + c + e], hc = Hc[r * ld_hc + c + e], hi = Hin[r * ld_hin + c + e];
      O[r * ld_o + c + e] = __fadd_rn(__fmul_rn(gg, hc), __fmul_rn(__fsub_rn(1.f, gg), hi));
    }
  }
}

// ----------------------------------------------------- column-slice pack / unpack
// row layout [n, F] <-> P column slices [P][n][Fp] (zero padded), the two transposes around the
// feature-sliced multi-GPU propagation.
// one warp per row, lanes over the P*Fp/4 float4 of the sliced side: no 64-bit divisions in the loop
__global__ void __launch_bounds__(256) pack_cols_kernel(const float* __restrict__ src, int64_t ld, int64_t n,
                                                        int64_t F, int P, int64_t Fp, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int fp4 = (int)(Fp >> 2), row_f4 = P * fp4;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (int64_t)gridDim.x * 8) {
    const float* srow = src + r * ld;
    for (int j = lane; j < row_f4; j += 32) {
      const int q = j / fp4, c4 = j - q * fp4;
      const int64_t col = (int64_t)q * Fp + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col + 3 < F) v = *reinterpret_cast<const float4*>(srow + col);
      else {
        if (col < F) v.x = srow[col];
        if (col + 1 < F) v.y = srow[col + 1];
        if (col + 2 < F) v.z = srow[col + 2];
      }
      reinterpret_cast<float4*>(dst + ((int64_t)q * n + r) * Fp)[c4] = v;
    }
  }
}
__global__ void __launch_bounds__(256) unpack_cols_kernel(const float* __restrict__ src, int64_t n, int64_t F, int P,
                                                          int64_t Fp, float* __restrict__ dst, int64_t ld) {
  const int lane = threadIdx.x & 31;
  const int fp4 = (int)(Fp >> 2), row_f4 = P * fp4;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (int64_t)gridDim.x * 8) {
    float* drow = dst + r * ld;
    for (int j = lane; j < row_f4; j += 32) {
      const int q = j / fp4, c4 = j - q * fp4;
      const int64_t col = (int64_t)q * Fp + c4 * 4;
      if (col >= F) continue;
      const float4 v = reinterpret_cast<const float4*>(src + ((int64_t)q * n + r) * Fp)[c4];
      if (col + 3 < F) *reinterpret_cast<float4*>(drow + col) = v;
      else {
        drow[col] = v.x;
        if (col + 1 < F) drow[col + 1] = v.y;
        if (col + 2 < F) drow[col + 2] = v.z;
      }
    }
  }
}

// ------------------------------------------------------------- softmax + CE
// one warp per target row; the row (<= a few KB) is re-read from L1.
__global__ void __launch_bounds__(256) softmax_ce_kernel(
    const float* __restrict__ L, int64_t ld_l, const int32_t* __restrict__ y, int64_t n_idx,
    int64_t C, float denom, float* __restrict__ probs, int64_t ld_p, float* __restrict__ G,
    int64_t ld_g, float* __restrict__ ce, float* __restrict__ hit, int64_t* __restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_idx) return;
  const float* row = L + i * ld_l;
  float m = -INFINITY;
  int64_t am = INT64_MAX;
  for (int64_t c = lane; c < C; c += 32) {
    const float v = row[c];
    if (v > m) { m = v; am = c; }  // strict: the first maximum wins, like np.argmax
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int64_t oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  if (am == INT64_MAX) am = 0;  // row of -inf
  float s = 0.f;
  for (int64_t c = lane; c < C; c += 32) s += expf(row[c] - m);
  s = warp_sum(s);
  const int yi = y ? y[i] : -1;
  if (lane == 0) {
    if (pred) pred[i] = am;
    if (hit) hit[i] = (am == (int64_t)yi) ? 1.f : 0.f;
    // a label outside [0, C) has no cross-entropy (the reference would raise an index error): NaN, never stale data
    if (ce) ce[i] = (yi >= 0 && yi < C) ? (m + logf(s)) - row[yi] : __int_as_float(0x7fc00000);
  }
  if (probs || G) {
    for (int64_t c = lane; c < C; c += 32) {
      const float p = __fdiv_rn(expf(row[c] - m), s);
      if (probs) probs[i * ld_p + c] = p;
      if (G) G[i * ld_g + c] = __fdiv_rn(c == yi ? p - 1.f : p, denom);
    }
  }
}

// ------------------------------------------------------------------- sum
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, int64_t n, float scale,
                                                   float* __restrict__ out) {
  __shared__ float sm[1024];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += x[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sm[0] * scale;
}

// ------------------------------------------------------- scatter / gather
template <int VEC>
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ G, int64_t ld_g,
                                                           const int32_t* __restrict__ pos_ptr,
                                                           const int32_t* __restrict__ pos_idx,
                                                           int64_t n_rows, int64_t W,
                                                           float* __restrict__ dP, int64_t ld_dp) {
  using T = typename Vec<VEC>::T;
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int b = __ldg(pos_ptr + r), e = __ldg(pos_ptr + r + 1);
  for (int64_t c = lane; c < W; c += 32) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = b; k < e; ++k) {
      const int64_t p = __ldg(pos_idx + k);
      const T v = ld(reinterpret_cast<const T*>(G + p * ld_g) + c);
      if (VEC == 4) {
        const float4 q = *reinterpret_cast<const float4*>(&v);
        acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
      } else {
        acc[0] += *reinterpret_cast<const float*>(&v);
      }
    }
    if (VEC == 4)
      *(reinterpret_cast<float4*>(dP + r * ld_dp) + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else
      dP[r * ld_dp + c] = acc[0];
  }
}
template <int VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ X, int64_t ld_x,
                                                          const int32_t* __restrict__ idx, int64_t n_idx,
                                                          int64_t W, float* __restrict__ out, int64_t ld_out) {
  using T = typename Vec<VEC>::T;
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_idx) return;
  const int64_t r = __ldg(idx + i);
  for (int64_t c = lane; c < W; c += 32)
    *(reinterpret_cast<T*>(out + i * ld_out) + c) = ld(reinterpret_cast<const T*>(X + r * ld_x) + c);
}

template <int VEC>
__global__ void __launch_bounds__(256) put_rows_kernel(const float* __restrict__ src, int64_t ld_src,
                                                       const int32_t* __restrict__ idx, int64_t n_idx,
                                                       int64_t W, float* __restrict__ dst, int64_t ld_dst) {
  using T = typename Vec<VEC>::T;
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_idx) return;
  const int64_t r = __ldg(idx + i);
  for (int64_t c = lane; c < W; c += 32)
    *(reinterpret_cast<T*>(dst + r * ld_dst) + c) = ld(reinterpret_cast<const T*>(src + i * ld_src) + c);
}

// --------------------------------------------------------------------- Adam
constexpr int kAdamMaxTensors = 32;
constexpr int kAdamChunk = 4096;  // elements per block
struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  int64_t n[kAdamMaxTensors];
  float reg[kAdamMaxTensors];
  int block0[kAdamMaxTensors + 1];  // first block of each tensor
  int n_tensors;
};

// t <- t+1 ; a_t = lr*sqrt(1-b2^t)/(1-b1^t)   (lasagne.updates.adam)
__global__ void adam_tick_kernel(float* t, float lr, float b1, float b2) {
  const float tn = t[0] + 1.f;
  t[0] = tn;
  t[1] = lr * sqrtf(1.f - powf(b2, tn)) / (1.f - powf(b1, tn));
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamTable tb, const float* __restrict__ tstate,
                                                   float b1, float b2, float eps,
                                                   float* __restrict__ reg_part) {
  __shared__ float sm[256];
  int k = 0;
  while (k + 1 < tb.n_tensors && (int)blockIdx.x >= tb.block0[k + 1]) ++k;
  const int64_t off = (int64_t)(blockIdx.x - tb.block0[k]) * kAdamChunk;
  const int64_t n = tb.n[k];
  float* __restrict__ p = tb.p[k];
  const float* __restrict__ g = tb.g[k];
  float* __restrict__ m = tb.m[k];
  float* __restrict__ v = tb.v[k];
  const float creg = tb.reg[k];
  const float a_t = tstate[1];
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  float racc = 0.f;
  const int64_t end = min(n, off + kAdamChunk);
  for (int64_t i = off + threadIdx.x; i < end; i += 256) {
    const float pi = p[i];
    float gi = g[i];
    if (creg != 0.f) {
      const float sg = (pi > 0.f) ? 1.f : ((pi < 0.f) ? -1.f : 0.f);
      gi = gi + 0.5f * creg * (sg + 2.f * pi);
      racc += 0.5f * creg * (fabsf(pi) + pi * pi);
    }
    const float mi = b1 * m[i] + omb1 * gi;
    const float vi = b2 * v[i] + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - a_t * mi / (sqrtf(vi) + eps);
  }
  if (reg_part) {
    sm[threadIdx.x] = racc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) reg_part[blockIdx.x] = sm[0];
  }
}

// sum_k reg[k]*0.5*(|p|_1 + |p|_2^2): block partials (same table as Adam)
__global__ void __launch_bounds__(256) elastic_net_kernel(const AdamTable tb, float* __restrict__ reg_part) {
  __shared__ float sm[256];
  int k = 0;
  while (k + 1 < tb.n_tensors && (int)blockIdx.x >= tb.block0[k + 1]) ++k;
  const int64_t off = (int64_t)(blockIdx.x - tb.block0[k]) * kAdamChunk;
  const int64_t end = min(tb.n[k], off + kAdamChunk);
  const float* __restrict__ p = tb.p[k];
  const float creg = tb.reg[k];
  float racc = 0.f;
  if (creg != 0.f)
    for (int64_t i = off + threadIdx.x; i < end; i += 256) {
      const float pi = p[i];
      racc += 0.5f * creg * (fabsf(pi) + pi * pi);
    }
  sm[threadIdx.x] = racc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) reg_part[blockIdx.x] = sm[0];
}

}  // namespace gcg

using namespace gcg;

// dst[i] = sum over s (ascending) of src[s * slab + i]: the per-document-block partial rows of X^T.dZ1 reduced in
// block order (deterministic); float4 path, slab must be a multiple of 4 floats
__global__ void __launch_bounds__(256) sum_slabs_kernel(const float4* __restrict__ src, int n_slabs, int64_t slab4,
                                                        float4* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slab4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = src[i];
    for (int sgi = 1; sgi < n_slabs; ++sgi) {
      const float4 b = src[(int64_t)sgi * slab4 + i];
      a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
    }
    dst[i] = a;
  }
}

extern "C" int gcg_sum_slabs_f32(const float* src, int32_t n_slabs, int64_t slab_floats, float* dst, void* stream) {
  GCG_RECORD("gcg_sum_slabs_f32", gcg_sum_slabs_f32(src, n_slabs, slab_floats, dst, s__));
  GCG_CHECK_ARG(src && dst && n_slabs > 0 && slab_floats >= 0, "gcg_sum_slabs_f32: bad argument");
  GCG_CHECK_SHAPE(slab_floats % 4 == 0 && aligned16(src) && aligned16(dst), "gcg_sum_slabs_f32: needs 16-byte aligned slabs of 4k floats");
  if (slab_floats == 0) return GCG_OK;
  const int64_t slab4 = slab_floats / 4;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(slab4, 256), (int64_t)kNumSMs * 16);
  sum_slabs_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(src), n_slabs, slab4,
                                                                              reinterpret_cast<float4*>(dst));
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int64_t gcg_colsum_workspace_bytes(int64_t n_rows, int64_t F) {
  if (n_rows <= 0 || F <= 0) return 0;
  return (int64_t)colsum_parts(n_rows, F) * F * (int64_t)sizeof(float);
}

extern "C" int gcg_colsum_f32(const float* X, int64_t ld, int64_t n_rows, int64_t F, float* out,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  GCG_RECORD("gcg_colsum_f32", gcg_colsum_f32(X, ld, n_rows, F, out, workspace, workspace_bytes, s__));
  GCG_CHECK_ARG(X && out, "gcg_colsum_f32: NULL argument");
  GCG_CHECK_SHAPE(F > 0 && n_rows >= 0 && ld >= F, "gcg_colsum_f32: bad shape n=%lld F=%lld ld=%lld",
                  (long long)n_rows, (long long)F, (long long)ld);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_rows == 0) {
    GCG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * F, st));
    return GCG_OK;
  }
  const int parts = colsum_parts(n_rows, F);
  GCG_CHECK_ARG(workspace && workspace_bytes >= (int64_t)parts * F * (int64_t)sizeof(float),
                "gcg_colsum_f32: workspace too small");
  float* part = reinterpret_cast<float*>(workspace);
  if (vec_ok(F, {{X, ld}})) {
    dim3 grid((unsigned)ceil_div(F, 128), (unsigned)parts);
    colsum_partial_kernel<4><<<grid, 256, 0, st>>>(X, ld, n_rows, F, part);
  } else {
    dim3 grid((unsigned)ceil_div(F, 32), (unsigned)parts);
    colsum_partial_kernel<1><<<grid, 256, 0, st>>>(X, ld, n_rows, F, part);
  }
  GCG_LAUNCH_CHECK();
  colsum_final_kernel<<<(unsigned)ceil_div(F, 256), 256, 0, st>>>(part, parts, F, out);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_act_bwd_f32(const float* dA, int64_t ld_da, const float* A, int64_t ld_a,
                               float* dP, int64_t ld_dp, int64_t n_rows, int64_t F, int act,
                               void* stream) {
  GCG_RECORD("gcg_act_bwd_f32", gcg_act_bwd_f32(dA, ld_da, A, ld_a, dP, ld_dp, n_rows, F, act, s__));
  GCG_CHECK_ARG(dA && A && dP, "gcg_act_bwd_f32: NULL argument");
  GCG_CHECK_SHAPE(F > 0 && ld_da >= F && ld_a >= F && ld_dp >= F, "gcg_act_bwd_f32: bad leading dimension");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_rows == 0) return GCG_OK;
  if (vec_ok(F, {{dA, ld_da}, {A, ld_a}, {dP, ld_dp}})) {
    const int64_t W = (F + 3) / 4;
    act_bwd_kernel<4><<<grid_for(n_rows * W, 256), 256, 0, st>>>(dA, ld_da, A, ld_a, dP, ld_dp, n_rows, W, act);
  } else {
    act_bwd_kernel<1><<<grid_for(n_rows * F, 256), 256, 0, st>>>(dA, ld_da, A, ld_a, dP, ld_dp, n_rows, F, act);
  }
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_highway_bwd_f32(const float* dO, int64_t ld_do, const float* g, int64_t ld_g,
                                   const float* Hc, int64_t ld_hc, const float* Hin, int64_t ld_hin,
                                   float* dP, int64_t ld_dp, float* dGpre, int64_t ld_dg,
                                   float* dHin, int64_t ld_dh, int64_t n_rows, int64_t F, int act,
                                   void* stream) {
  GCG_RECORD("gcg_highway_bwd_f32", gcg_highway_bwd_f32(dO, ld_do, g, ld_g, Hc, ld_hc, Hin, ld_hin, dP, ld_dp, dGpre, ld_dg, dHin, ld_dh, n_rows, F, act, s__));
  GCG_CHECK_ARG(dO && g && Hc && Hin && dP && dGpre && dHin, "gcg_highway_bwd_f32: NULL argument");
  GCG_CHECK_SHAPE(F > 0 && ld_do >= F && ld_g >= F && ld_hc >= F && ld_hin >= F && ld_dp >= F &&
                      ld_dg >= F && ld_dh >= F,
                  "gcg_highway_bwd_f32: bad leading dimension");
  GCG_CHECK_ARG(dP != dGpre && dP != dHin && dGpre != dHin, "gcg_highway_bwd_f32: outputs alias each other");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_rows == 0) return GCG_OK;
  if (vec_ok(F, {{dO, ld_do}, {g, ld_g}, {Hc, ld_hc}, {Hin, ld_hin}, {dP, ld_dp}, {dGpre, ld_dg}, {dHin, ld_dh}})) {
    const int64_t W = (F + 3) / 4;
    highway_bwd_kernel<4><<<grid_for(n_rows * W, 256), 256, 0, st>>>(
        dO, ld_do, g, ld_g, Hc, ld_hc, Hin, ld_hin, dP, ld_dp, dGpre, ld_dg, dHin, ld_dh, n_rows, W, act);
  } else {
    highway_bwd_kernel<1><<<grid_for(n_rows * F, 256), 256, 0, st>>>(
        dO, ld_do, g, ld_g, Hc, ld_hc, Hin, ld_hin, dP, ld_dp, dGpre, ld_dg, dHin, ld_dh, n_rows, F, act);
  }
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_softmax_ce_f32(const float* L, int64_t ld_l, const int32_t* y, int64_t n_idx,
                                  int64_t C, float denom, float* probs, int64_t ld_p, float* G,
                                  int64_t ld_g, float* ce, float* hit, int64_t* pred, void* stream) {
  GCG_RECORD("gcg_softmax_ce_f32", gcg_softmax_ce_f32(L, ld_l, y, n_idx, C, denom, probs, ld_p, G, ld_g, ce, hit, pred, s__));
  GCG_CHECK_ARG(L != nullptr, "gcg_softmax_ce_f32: logits NULL");
  GCG_CHECK_SHAPE(C > 0 && ld_l >= C && (!probs || ld_p >= C) && (!G || ld_g >= C),
                  "gcg_softmax_ce_f32: bad shape C=%lld", (long long)C);
  GCG_CHECK_ARG(!(G || ce || hit) || y, "gcg_softmax_ce_f32: labels needed for G / ce / hit");
  GCG_CHECK_ARG(!G || denom != 0.f, "gcg_softmax_ce_f32: denom is 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_idx == 0) return GCG_OK;
  softmax_ce_kernel<<<(unsigned)ceil_div(n_idx, 8), 256, 0, st>>>(L, ld_l, y, n_idx, C, denom, probs, ld_p,
                                                                   G, ld_g, ce, hit, pred);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_sum_f32(const float* x, int64_t n, float scale, float* out, void* stream) {
  GCG_RECORD("gcg_sum_f32", gcg_sum_f32(x, n, scale, out, s__));
  GCG_CHECK_ARG(out && (x || n == 0), "gcg_sum_f32: NULL argument");
  sum_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, scale, out);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_scatter_rows_f32(const float* G, int64_t ld_g, const int32_t* pos_ptr,
                                    const int32_t* pos_idx, int64_t n_rows, int64_t C, float* dP,
                                    int64_t ld_dp, void* stream) {
  GCG_RECORD("gcg_scatter_rows_f32", gcg_scatter_rows_f32(G, ld_g, pos_ptr, pos_idx, n_rows, C, dP, ld_dp, s__));
  GCG_CHECK_ARG(G && pos_ptr && pos_idx && dP, "gcg_scatter_rows_f32: NULL argument");
  GCG_CHECK_SHAPE(C > 0 && ld_g >= C && ld_dp >= C, "gcg_scatter_rows_f32: bad leading dimension");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_rows == 0) return GCG_OK;
  if (vec_ok(C, {{G, ld_g}, {dP, ld_dp}}))
    scatter_rows_kernel<4><<<(unsigned)ceil_div(n_rows, 8), 256, 0, st>>>(G, ld_g, pos_ptr, pos_idx, n_rows, (C + 3) / 4, dP, ld_dp);
  else
    scatter_rows_kernel<1><<<(unsigned)ceil_div(n_rows, 8), 256, 0, st>>>(G, ld_g, pos_ptr, pos_idx, n_rows, C, dP, ld_dp);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_gather_rows_f32(const float* X, int64_t ld_x, const int32_t* idx, int64_t n_idx,
                                   int64_t C, float* out, int64_t ld_out, void* stream) {
  GCG_RECORD("gcg_gather_rows_f32", gcg_gather_rows_f32(X, ld_x, idx, n_idx, C, out, ld_out, s__));
  GCG_CHECK_ARG(X && idx && out, "gcg_gather_rows_f32: NULL argument");
  GCG_CHECK_SHAPE(C > 0 && ld_x >= C && ld_out >= C, "gcg_gather_rows_f32: bad leading dimension");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_idx == 0) return GCG_OK;
  if (vec_ok(C, {{X, ld_x}, {out, ld_out}}))
    gather_rows_kernel<4><<<(unsigned)ceil_div(n_idx, 8), 256, 0, st>>>(X, ld_x, idx, n_idx, (C + 3) / 4, out, ld_out);
  else
    gather_rows_kernel<1><<<(unsigned)ceil_div(n_idx, 8), 256, 0, st>>>(X, ld_x, idx, n_idx, C, out, ld_out);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_put_rows_f32(const float* src, int64_t ld_src, const int32_t* idx, int64_t n_idx,
                                int64_t C, float* dst, int64_t ld_dst, void* stream) {
  GCG_RECORD("gcg_put_rows_f32", gcg_put_rows_f32(src, ld_src, idx, n_idx, C, dst, ld_dst, s__));
  GCG_CHECK_ARG(src && idx && dst, "gcg_put_rows_f32: NULL argument");
  GCG_CHECK_SHAPE(C > 0 && ld_src >= C && ld_dst >= C, "gcg_put_rows_f32: bad leading dimension");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_idx == 0) return GCG_OK;
  if (vec_ok(C, {{src, ld_src}, {dst, ld_dst}}))
    put_rows_kernel<4><<<(unsigned)ceil_div(n_idx, 8), 256, 0, st>>>(src, ld_src, idx, n_idx, (C + 3) / 4, dst, ld_dst);
  else
    put_rows_kernel<1><<<(unsigned)ceil_div(n_idx, 8), 256, 0, st>>>(src, ld_src, idx, n_idx, C, dst, ld_dst);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

static int64_t adam_blocks(int32_t n_tensors, const int64_t* sizes) {
  int64_t b = 0;
  for (int i = 0; i < n_tensors; ++i) b += ceil_div(sizes[i], kAdamChunk);
  return b;
}

extern "C" int64_t gcg_adam_workspace_bytes(int32_t n_tensors, const int64_t* h_sizes) {
  if (n_tensors <= 0 || !h_sizes) return 0;
  return adam_blocks(n_tensors, h_sizes) * (int64_t)sizeof(float);
}

extern "C" int gcg_adam_step_f32(int32_t n_tensors, float* const* h_params, const float* const* h_grads,
                                 float* const* h_m, float* const* h_v, const int64_t* h_sizes,
                                 const float* h_reg, float lr, float beta1, float beta2, float eps,
                                 float* d_t, float* reg_out, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  if (gcg::epoch_recording() && n_tensors > 0 && h_params && h_grads && h_m && h_v && h_sizes && h_reg) {
    // the pointer / size / coefficient tables live on the host: the closure keeps its own copies
    std::vector<float*> cp(h_params, h_params + n_tensors), cm(h_m, h_m + n_tensors), cv(h_v, h_v + n_tensors);
    std::vector<const float*> cg(h_grads, h_grads + n_tensors);
    std::vector<int64_t> cs(h_sizes, h_sizes + n_tensors);
    std::vector<float> cr(h_reg, h_reg + n_tensors);
    gcg::epoch_record("gcg_adam_step_f32", [=](void* s__) -> int {
      return gcg_adam_step_f32(n_tensors, cp.data(), cg.data(), cm.data(), cv.data(), cs.data(), cr.data(), lr, beta1,
                               beta2, eps, d_t, reg_out, workspace, workspace_bytes, s__);
    });
  }
  GCG_CHECK_ARG(n_tensors > 0 && n_tensors <= kAdamMaxTensors, "gcg_adam_step_f32: n_tensors=%d (max %d)",
                n_tensors, kAdamMaxTensors);
  GCG_CHECK_ARG(h_params && h_grads && h_m && h_v && h_sizes && d_t, "gcg_adam_step_f32: NULL argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  AdamTable tb;
  int64_t blocks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    GCG_CHECK_ARG(h_params[i] && h_grads[i] && h_m[i] && h_v[i] && h_sizes[i] >= 0,
                  "gcg_adam_step_f32: tensor %d has a NULL pointer", i);
    tb.p[i] = h_params[i]; tb.g[i] = h_grads[i]; tb.m[i] = h_m[i]; tb.v[i] = h_v[i];
    tb.n[i] = h_sizes[i]; tb.reg[i] = h_reg ? h_reg[i] : 0.f;
    tb.block0[i] = (int)blocks;
    blocks += ceil_div(h_sizes[i], kAdamChunk);
  }
  tb.block0[n_tensors] = (int)blocks;
  tb.n_tensors = n_tensors;
  GCG_CHECK_SHAPE(blocks < INT32_MAX, "gcg_adam_step_f32: too many elements");
  float* part = nullptr;
  if (reg_out) {
    GCG_CHECK_ARG(workspace && workspace_bytes >= blocks * (int64_t)sizeof(float),
                  "gcg_adam_step_f32: workspace too small");
    part = reinterpret_cast<float*>(workspace);
  }
  adam_tick_kernel<<<1, 1, 0, st>>>(d_t, lr, beta1, beta2);
  GCG_LAUNCH_CHECK();
  if (blocks > 0) {
    adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(tb, d_t, beta1, beta2, eps, part);
    GCG_LAUNCH_CHECK();
  }
  if (reg_out) {
    sum_kernel<<<1, 1024, 0, st>>>(part, blocks, 1.f, reg_out);
    GCG_LAUNCH_CHECK();
  }
  return GCG_OK;
}

extern "C" int gcg_elastic_net_f32(int32_t n_tensors, const float* const* h_params, const int64_t* h_sizes,
                                   const float* h_reg, float* out, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
  if (gcg::epoch_recording() && n_tensors > 0 && h_params && h_sizes && h_reg) {
    std::vector<const float*> cp(h_params, h_params + n_tensors);
    std::vector<int64_t> cs(h_sizes, h_sizes + n_tensors);
    std::vector<float> cr(h_reg, h_reg + n_tensors);
    gcg::epoch_record("gcg_elastic_net_f32", [=](void* s__) -> int {
      return gcg_elastic_net_f32(n_tensors, cp.data(), cs.data(), cr.data(), out, workspace, workspace_bytes, s__);
    });
  }
  GCG_CHECK_ARG(n_tensors > 0 && n_tensors <= kAdamMaxTensors, "gcg_elastic_net_f32: n_tensors=%d (max %d)",
                n_tensors, kAdamMaxTensors);
  GCG_CHECK_ARG(h_params && h_sizes && h_reg && out, "gcg_elastic_net_f32: NULL argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  AdamTable tb;
  int64_t blocks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    GCG_CHECK_ARG(h_params[i] && h_sizes[i] >= 0, "gcg_elastic_net_f32: tensor %d is NULL", i);
    tb.p[i] = const_cast<float*>(h_params[i]); tb.g[i] = nullptr; tb.m[i] = nullptr; tb.v[i] = nullptr;
    tb.n[i] = h_sizes[i]; tb.reg[i] = h_reg[i];
    tb.block0[i] = (int)blocks;
    blocks += ceil_div(h_sizes[i], kAdamChunk);
  }
  tb.block0[n_tensors] = (int)blocks;
  tb.n_tensors = n_tensors;
  GCG_CHECK_ARG(blocks == 0 || (workspace && workspace_bytes >= blocks * (int64_t)sizeof(float)),
                "gcg_elastic_net_f32: workspace too small");
  float* part = reinterpret_cast<float*>(workspace);
  if (blocks > 0) {
    elastic_net_kernel<<<(unsigned)blocks, 256, 0, st>>>(tb, part);
    GCG_LAUNCH_CHECK();
  }
  sum_kernel<<<1, 1024, 0, st>>>(part, blocks, 1.f, out);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_highway_fwd_f32(const float* Hc, int64_t ld_hc, const float* g, int64_t ld_g,
                                   const float* Hin, int64_t ld_hin, float* O, int64_t ld_o,
                                   int64_t n_rows, int64_t F, void* stream) {
  GCG_RECORD("gcg_highway_fwd_f32", gcg_highway_fwd_f32(Hc, ld_hc, g, ld_g, Hin, ld_hin, O, ld_o, n_rows, F, s__));
  GCG_CHECK_ARG(Hc && g && Hin && O, "gcg_highway_fwd_f32: NULL argument");
  GCG_CHECK_SHAPE(F > 0 && ld_hc >= F && ld_g >= F && ld_hin >= F && ld_o >= F, "gcg_highway_fwd_f32: bad leading dimension");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_rows == 0) return GCG_OK;
  if (vec_ok(F, {{Hc, ld_hc}, {g, ld_g}, {Hin, ld_hin}, {O, ld_o}}))
    highway_fwd_kernel<4><<<grid_for(n_rows * ((F + 3) / 4), 256), 256, 0, st>>>(Hc, ld_hc, g, ld_g, Hin, ld_hin, O, ld_o, n_rows, (F + 3) / 4);
  else
    highway_fwd_kernel<1><<<grid_for(n_rows * F, 256), 256, 0, st>>>(Hc, ld_hc, g, ld_g, Hin, ld_hin, O, ld_o, n_rows, F);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_pack_cols_f32(const float* src, int64_t ld, int64_t n_rows, int64_t F, int32_t P, int64_t Fp,
                                 float* dst, void* stream) {
  GCG_RECORD("gcg_pack_cols_f32", gcg_pack_cols_f32(src, ld, n_rows, F, P, Fp, dst, s__));
  GCG_CHECK_ARG(src && dst && P > 0, "gcg_pack_cols_f32: bad argument");
  GCG_CHECK_SHAPE(Fp % 4 == 0 && ld % 4 == 0 && ld >= F && (int64_t)P * Fp >= F && aligned16(src) && aligned16(dst),
                  "gcg_pack_cols_f32: needs 16-byte aligned operands, ld %% 4 == 0, Fp %% 4 == 0");
  if (n_rows == 0) return GCG_OK;
  pack_cols_kernel<<<grid_for(n_rows * 32, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, ld, n_rows, F, P, Fp, dst);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_unpack_cols_f32(const float* src, int64_t n_rows, int64_t F, int32_t P, int64_t Fp, float* dst,
                                   int64_t ld, void* stream) {
  GCG_RECORD("gcg_unpack_cols_f32", gcg_unpack_cols_f32(src, n_rows, F, P, Fp, dst, ld, s__));
  GCG_CHECK_ARG(src && dst && P > 0, "gcg_unpack_cols_f32: bad argument");
  GCG_CHECK_SHAPE(Fp % 4 == 0 && ld % 4 == 0 && ld >= F && (int64_t)P * Fp >= F && aligned16(src) && aligned16(dst),
                  "gcg_unpack_cols_f32: needs 16-byte aligned operands, ld %% 4 == 0, Fp %% 4 == 0");
  if (n_rows == 0) return GCG_OK;
  unpack_cols_kernel<<<grid_for(n_rows * 32, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, n_rows, F, P, Fp, dst, ld);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}
