// Input smoothing X_conv = A_hat * X (sparse x sparse) and the device-side CSR helpers of the
// minibatch path (main.py:528-534, tensormain.py:112-118, mlp.py:81-91).
//
// SpGEMM design (B200): the product is latency/HBM bound integer + float64 work, no tensor cores.
// A fixed pool of "workers" (one warp each, 4 per CTA, a multiple of the 148 SMs) pulls output rows
// from an atomic counter.  Every worker owns, in the workspace, a dense accumulator acc[n_cols_b]
// (float64, or float32 for a float32 A) that is all-zero between rows, and a bitmap of the columns
// touched by the current row.  A's entries are visited strictly in stored order and the 32 lanes split
// one row of B, so every output entry receives its additions `sums[k] += a*b` (from 0, multiply and add
// rounded separately) in exactly scipy's csr_matmat order -> bit-identical values after the single
// float32 rounding.  The row is then emitted by one sweep over the non-empty 1024-column groups of the
// bitmap (a second-level summary bitmap names them, so short rows over a wide column space -- the graph
// projection -- do not pay for the range they span), which yields ascending columns (the canonical form
// `X_conv.tocsr().astype('float32')` ends up in, because scipy's astype sorts the indices) and resets
// accumulator and bitmap for the next row.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>

#include "gcg_common.cuh"

namespace gcg {
namespace {

constexpr int kWarpsPerCta = 4;
constexpr int64_t kSpgemmBudget = 12LL << 30;  // bytes of dense accumulators across all workers (HBM is 180 GB)
constexpr int kMaxWarpsPerSM = 32;             // latency-bound kernel: ncu showed 22 % occupancy at 14 warps/SM

struct SpgemmLayout {
  int workers;
  int64_t vpad;              // columns padded to a multiple of 1024 (32 bitmap words = one group)
  int64_t sum_words;         // summary words per worker (one bit per group), a multiple of 32
  int64_t bitmap_off, summary_off, acc_off, total;
};

SpgemmLayout spgemm_layout(int64_t n_cols) {
  SpgemmLayout L;
  L.vpad = (n_cols + 1023) / 1024 * 1024;
  if (L.vpad == 0) L.vpad = 1024;
  L.sum_words = ((L.vpad >> 10) + 1023) / 1024 * 32;
  int64_t w = kSpgemmBudget / (L.vpad * 8 + L.vpad / 8 + L.sum_words * 4);
  w = std::min<int64_t>(w, (int64_t)kNumSMs * kMaxWarpsPerSM);
  w = std::max<int64_t>(w, (int64_t)kNumSMs);
  L.workers = (int)(w / kWarpsPerCta * kWarpsPerCta);
  L.bitmap_off = 256;
  L.summary_off = L.bitmap_off + (int64_t)L.workers * (L.vpad / 8);
  L.acc_off = L.summary_off + (int64_t)L.workers * L.sum_words * 4;
  L.total = L.acc_off + (int64_t)L.workers * L.vpad * 8;
  return L;
}

struct SpgemmArgs {
  const int32_t* a_indptr;
  const int32_t* a_indices;
  const void* a_vals;
  const int32_t* b_indptr;
  const int32_t* b_indices;
  const float* b_vals;
  int32_t* row_nnz;
  const int64_t* c_indptr;
  int32_t* c_indices;
  float* c_vals;
  int32_t* counter;
  uint32_t* bitmap;
  uint32_t* summary;
  double* acc;
  int64_t vpad;
  int64_t sum_words;
  int n_rows;
  int drop_diagonal;   // skip column == row (graph projection: no self edges)
  int pattern_only;    // NUMERIC pass without values (c_vals == NULL)
};

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// NUMERIC = false: count the distinct columns of every output row.
// NUMERIC = true : accumulate in AT (float64 A -> float64 sums, float32 A -> float32 sums, as scipy's
//                  type promotion does) and emit the row with ascending columns.
template <bool NUMERIC, typename AT>
__global__ void __launch_bounds__(kWarpsPerCta * 32) spgemm_rows_kernel(SpgemmArgs a) {
  const int lane = threadIdx.x & 31;
  const int worker = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  uint32_t* bitmap = a.bitmap + (size_t)worker * (a.vpad >> 5);
  uint32_t* summary = a.summary + (size_t)worker * a.sum_words;     // bit g: group g (1024 columns) is non-empty
  AT* acc = reinterpret_cast<AT*>(a.acc + (size_t)worker * a.vpad);
  const AT* a_vals = reinterpret_cast<const AT*>(a.a_vals);

  for (;;) {
    int row = 0;
    if (lane == 0) row = atomicAdd(a.counter, 1);
    row = __shfl_sync(0xffffffffu, row, 0);
    if (row >= a.n_rows) break;
    const int pb = a.a_indptr[row], pe = a.a_indptr[row + 1];
    int kmin = INT32_MAX, kmax = -1;
    for (int p = pb; p < pe; ++p) {
      const int j = a.a_indices[p];
      AT av = AT(0);
      if (NUMERIC && !a.pattern_only) av = a_vals[p];
      const int qb = a.b_indptr[j], qe = a.b_indptr[j + 1];
      for (int q0 = qb; q0 < qe; q0 += 32) {
        const int q = q0 + lane;
        const int k = (q < qe) ? a.b_indices[q] : -1;
        // group marks: lanes whose left neighbour is in the same 1024-column group skip theirs (B's rows are
        // normally sorted, so a chunk issues a handful); both atomics are fire-and-forget reductions
        const int g = k >> 10;
        const int g_left = __shfl_up_sync(0xffffffffu, g, 1);
        if (k >= 0) {
          kmin = min(kmin, k);
          kmax = max(kmax, k);
          atomicOr(&bitmap[k >> 5], 1u << (k & 31));
          if (lane == 0 || g != g_left) atomicOr(&summary[g >> 5], 1u << (g & 31));
          if (NUMERIC && !a.pattern_only) acc[k] = add_rn(acc[k], mul_rn(av, (AT)a.b_vals[q]));   // sums[k] += v * Bx[kk]
        }
        __syncwarp();   // the next chunk may touch the same columns from other lanes
      }
    }
    kmin = warp_min_i(kmin);
    kmax = warp_max_i(kmax);
    __syncwarp();
    // sweep the touched range of the bitmap: ascending columns; resets bitmap (and acc) for the next row
    int cnt = 0;
    int32_t* out_idx = nullptr;
    float* out_val = nullptr;
    if (NUMERIC) {
      const int64_t base = a.c_indptr[row];
      out_idx = a.c_indices + base;
      out_val = a.c_vals + base;
    }
    if (kmax >= 0) {
      // two-level sweep: summary words -> non-empty groups -> the group's 32 bitmap words (one per lane)
      const int s_end = kmax >> 15;
      for (int s0 = kmin >> 15; s0 <= s_end; s0 += 32) {
        const int sw = s0 + lane;
        const uint32_t sbits = (sw <= s_end) ? __ldcg(&summary[sw]) : 0u;   // set by atomics at L2: bypass L1
        if (sbits) summary[sw] = 0u;
        unsigned lanes = __ballot_sync(0xffffffffu, sbits != 0u);
        while (lanes) {
          const int src = __ffs(lanes) - 1;
          lanes &= lanes - 1;
          uint32_t groups = __shfl_sync(0xffffffffu, sbits, src);
          const int gbase = (s0 + src) << 5;
          while (groups) {
            const int g = gbase + __ffs(groups) - 1;
            groups &= groups - 1;
            const int w = (g << 5) + lane;
            uint32_t bits = __ldcg(&bitmap[w]);
            if (bits) bitmap[w] = 0u;
            if (a.drop_diagonal && w == (row >> 5)) bits &= ~(1u << (row & 31));
            const int c = __popc(bits);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += t;
            }
            if (NUMERIC && bits) {
              int pos = cnt + incl - c;
              while (bits) {
                const int bpos = __ffs(bits) - 1;
                bits &= bits - 1;
                const int k = (w << 5) + bpos;
                out_idx[pos] = k;
                if (!a.pattern_only) {
                  out_val[pos] = (float)acc[k];
                  acc[k] = AT(0);
                }
                ++pos;
              }
            }
            cnt += __shfl_sync(0xffffffffu, incl, 31);
          }
        }
      }
    }
    if (!NUMERIC && lane == 0) a.row_nnz[row] = cnt;
    __syncwarp();
  }
}

__global__ void csr_gather_rows_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                       const float* __restrict__ vals, const int32_t* __restrict__ rows,
                                       int n_sel, const int32_t* __restrict__ out_indptr,
                                       int32_t* __restrict__ out_indices, float* __restrict__ out_vals) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_sel; s += warps) {
    const int r = rows[s];
    const int sb = indptr[r], len = indptr[r + 1] - sb;
    const int db = out_indptr[s];
    for (int t = lane; t < len; t += 32) {
      out_indices[db + t] = indices[sb + t];
      out_vals[db + t] = vals[sb + t];
    }
  }
}

// entry e of row r -> row_of[e] = r, ident[e] = e   (offsets relative to indptr[0])
__global__ void csr_expand_rows_kernel(const int32_t* __restrict__ indptr, int n_rows, int32_t* __restrict__ row_of,
                                       int32_t* __restrict__ ident) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int base = indptr[0];
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
    const int b = indptr[r] - base, e = indptr[r + 1] - base;
    for (int t = b + lane; t < e; t += 32) {
      row_of[t] = r;
      ident[t] = t;
    }
  }
}

__global__ void csr_transpose_finish_kernel(const int32_t* __restrict__ sorted_cols, const int32_t* __restrict__ perm,
                                            const int32_t* __restrict__ row_of, const float* __restrict__ vals,
                                            int nnz, int n_cols, int32_t* __restrict__ t_indptr,
                                            int32_t* __restrict__ t_indices, float* __restrict__ t_vals) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= nnz; i += stride) {
    const int c_prev = (i == 0) ? -1 : sorted_cols[i - 1];
    const int c_cur = (i == nnz) ? n_cols : sorted_cols[i];
    for (int v = c_prev + 1; v <= c_cur; ++v) t_indptr[v] = i;
    if (i < nnz) {
      const int e = perm[i];
      t_indices[i] = row_of[e];
      t_vals[i] = vals[e];
    }
  }
}

int bits_for(int64_t n) {
  int b = 1;
  while (b < 31 && (1LL << b) < n) ++b;
  return b;
}

size_t radix_temp_bytes(int64_t nnz, int end_bit) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)nnz, 0, end_bit);
  return temp;
}

int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

}  // namespace
}  // namespace gcg

using namespace gcg;

extern "C" int64_t gcg_spgemm_workspace_bytes(int64_t n_cols_b) {
  if (n_cols_b < 0) return 0;
  return spgemm_layout(n_cols_b).total;
}

static int spgemm_common(bool numeric, int64_t n_rows, int64_t n_cols_b, SpgemmArgs& a, int a_is_f64,
                         void* workspace, int64_t workspace_bytes, void* stream) {
  GCG_CHECK_ARG(n_rows >= 0 && n_cols_b >= 0 && n_rows < INT32_MAX && n_cols_b < INT32_MAX,
                "gcg_spgemm: sizes out of range");
  GCG_CHECK_ARG(a.a_indptr && a.b_indptr, "gcg_spgemm: NULL indptr");
  const SpgemmLayout L = spgemm_layout(n_cols_b);
  GCG_CHECK_ARG(workspace && workspace_bytes >= L.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                "gcg_spgemm: workspace too small or not 256-byte aligned (%lld < %lld)",
                (long long)workspace_bytes, (long long)L.total);
  if (n_rows == 0) return GCG_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  a.counter = reinterpret_cast<int32_t*>(ws);
  a.bitmap = reinterpret_cast<uint32_t*>(ws + L.bitmap_off);
  a.summary = reinterpret_cast<uint32_t*>(ws + L.summary_off);
  a.sum_words = L.sum_words;
  a.acc = reinterpret_cast<double*>(ws + L.acc_off);
  a.vpad = L.vpad;
  a.n_rows = (int)n_rows;
  // row counter, bitmaps and (numeric pass) accumulators start at zero; the kernel leaves them zero
  GCG_CUDA(cudaMemsetAsync(ws, 0, (size_t)((numeric && !a.pattern_only) ? L.total : L.acc_off), st));
  const unsigned grid = (unsigned)(L.workers / kWarpsPerCta);
  if (!numeric)
    spgemm_rows_kernel<false, float><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
  else if (a_is_f64)
    spgemm_rows_kernel<true, double><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
  else
    spgemm_rows_kernel<true, float><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_spgemm_count_csr(int64_t n_rows, int64_t n_cols_b, const int32_t* a_indptr,
                                    const int32_t* a_indices, const int32_t* b_indptr, const int32_t* b_indices,
                                    int drop_diagonal, int32_t* row_nnz, void* workspace, int64_t workspace_bytes,
                                    void* stream) {
  GCG_CHECK_ARG(row_nnz || n_rows == 0, "gcg_spgemm_count_csr: row_nnz is NULL");
  SpgemmArgs a{};
  a.a_indptr = a_indptr; a.a_indices = a_indices;
  a.b_indptr = b_indptr; a.b_indices = b_indices;
  a.row_nnz = row_nnz;
  a.drop_diagonal = drop_diagonal ? 1 : 0;
  return spgemm_common(false, n_rows, n_cols_b, a, 0, workspace, workspace_bytes, stream);
}

extern "C" int gcg_spgemm_fill_pattern_csr(int64_t n_rows, int64_t n_cols_b, const int32_t* a_indptr,
                                           const int32_t* a_indices, const int32_t* b_indptr,
                                           const int32_t* b_indices, int drop_diagonal, const int64_t* c_indptr,
                                           int32_t* c_indices, void* workspace, int64_t workspace_bytes,
                                           void* stream) {
  GCG_CHECK_ARG(c_indptr, "gcg_spgemm_fill_pattern_csr: c_indptr is NULL");
  SpgemmArgs a{};
  a.a_indptr = a_indptr; a.a_indices = a_indices;
  a.b_indptr = b_indptr; a.b_indices = b_indices;
  a.c_indptr = c_indptr; a.c_indices = c_indices;
  a.drop_diagonal = drop_diagonal ? 1 : 0;
  a.pattern_only = 1;
  return spgemm_common(true, n_rows, n_cols_b, a, 0, workspace, workspace_bytes, stream);
}

extern "C" int gcg_spgemm_fill_csr_f32(int64_t n_rows, int64_t n_cols_b, const int32_t* a_indptr,
                                       const int32_t* a_indices, const void* a_vals, int a_is_f64,
                                       const int32_t* b_indptr, const int32_t* b_indices, const float* b_vals,
                                       const int64_t* c_indptr, int32_t* c_indices, float* c_vals,
                                       void* workspace, int64_t workspace_bytes, void* stream) {
  GCG_CHECK_ARG(c_indptr, "gcg_spgemm_fill_csr_f32: c_indptr is NULL");
  SpgemmArgs a{};
  a.a_indptr = a_indptr; a.a_indices = a_indices; a.a_vals = a_vals;
  a.b_indptr = b_indptr; a.b_indices = b_indices; a.b_vals = b_vals;
  a.c_indptr = c_indptr; a.c_indices = c_indices; a.c_vals = c_vals;
  return spgemm_common(true, n_rows, n_cols_b, a, a_is_f64, workspace, workspace_bytes, stream);
}

extern "C" int gcg_csr_gather_rows_device(const int32_t* d_indptr, const int32_t* d_indices, const float* d_vals,
                                          const int32_t* d_rows, int64_t n_sel, const int32_t* d_out_indptr,
                                          int32_t* d_out_indices, float* d_out_vals, void* stream) {
  GCG_CHECK_ARG(n_sel >= 0 && n_sel < INT32_MAX, "gcg_csr_gather_rows_device: bad n_sel");
  if (n_sel == 0) return GCG_OK;
  GCG_CHECK_ARG(d_indptr && d_rows && d_out_indptr, "gcg_csr_gather_rows_device: NULL argument");
  const int64_t blocks = std::min<int64_t>(ceil_div(n_sel, 8), (int64_t)kNumSMs * 8);
  csr_gather_rows_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_indptr, d_indices, d_vals, d_rows, (int)n_sel, d_out_indptr, d_out_indices, d_out_vals);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int64_t gcg_csr_transpose_device_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t nnz) {
  if (n_rows < 0 || n_cols < 0 || nnz < 0 || nnz >= INT32_MAX) return 0;
  const int64_t arr = align256(4 * std::max<int64_t>(nnz, 1));
  return 4 * arr + align256((int64_t)radix_temp_bytes(std::max<int64_t>(nnz, 1), bits_for(n_cols))) + 256;
}

extern "C" int gcg_csr_transpose_device(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* d_indptr,
                                        const int32_t* d_indices, const float* d_vals, int32_t* t_indptr,
                                        int32_t* t_indices, float* t_vals, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
  GCG_CHECK_ARG(n_rows >= 0 && n_cols >= 0 && nnz >= 0 && n_rows < INT32_MAX && n_cols < INT32_MAX - 1 &&
                    nnz < INT32_MAX, "gcg_csr_transpose_device: sizes out of range");
  GCG_CHECK_ARG(d_indptr && t_indptr, "gcg_csr_transpose_device: NULL indptr");
  const int64_t need = gcg_csr_transpose_device_workspace_bytes(n_rows, n_cols, nnz);
  GCG_CHECK_ARG(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                "gcg_csr_transpose_device: workspace too small or not 256-byte aligned (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t arr = align256(4 * std::max<int64_t>(nnz, 1));
  char* ws = reinterpret_cast<char*>(workspace);
  int32_t* row_of = reinterpret_cast<int32_t*>(ws);
  int32_t* ident = reinterpret_cast<int32_t*>(ws + arr);
  int32_t* perm = reinterpret_cast<int32_t*>(ws + 2 * arr);
  int32_t* sorted_cols = reinterpret_cast<int32_t*>(ws + 3 * arr);
  void* temp = ws + 4 * arr;
  const int end_bit = bits_for(n_cols);
  if (nnz > 0) {
    GCG_CHECK_ARG(d_indices && d_vals && t_indices && t_vals, "gcg_csr_transpose_device: NULL CSR array");
    const int64_t blocks = std::min<int64_t>(ceil_div(std::max<int64_t>(n_rows, 1), 8), (int64_t)kNumSMs * 8);
    csr_expand_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_indptr, (int)n_rows, row_of, ident);
    GCG_LAUNCH_CHECK();
    size_t temp_bytes = radix_temp_bytes(nnz, end_bit);
    GCG_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, d_indices, sorted_cols, (const int32_t*)ident, perm,
                                             (int)nnz, 0, end_bit, st));
    count_launch(1);
  }
  const int64_t fb = std::min<int64_t>(ceil_div(nnz + 1, 256), (int64_t)kNumSMs * 8);
  csr_transpose_finish_kernel<<<(unsigned)fb, 256, 0, st>>>(sorted_cols, perm, row_of, d_vals, (int)nnz, (int)n_cols,
                                                            t_indptr, t_indices, t_vals);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}
