// Device-side A_hat construction (SURVEY 8f "next" row 1): the normalisation of
// tensormain.py:170-180,221 for a BINARY adjacency pattern that is already a device CSR.
//   adj.setdiag(1); d = rowsum; A_hat = D^-1/2 adj D^-1/2 in float64; cast to float32.
// Bit-identical to gcg_ahat_build_host / the oracle: degrees are exact integers, 1/sqrt in IEEE
// float64 (__drcp/__dsqrt round-to-nearest), value = float((dinv_i * 1.0) * dinv_j).
#include "gcg_common.cuh"

namespace gcg {

// need[r] = 1 if row r lacks its diagonal entry (columns sorted -> binary search)
__global__ void __launch_bounds__(256) ahat_mark_kernel(int64_t n, const int* __restrict__ indptr,
                                                        const int* __restrict__ indices, int* __restrict__ need) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int lo = indptr[r], hi = indptr[r + 1];
  bool found = false;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int c = indices[mid];
    if (c == (int)r) { found = true; break; }
    if (c < (int)r) lo = mid + 1; else hi = mid;
  }
  need[r] = found ? 0 : 1;
}

// three-pass exclusive scan of `need` added to indptr: out[r] = indptr[r] + sum_{q<r} need[q]
constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(kScanBlock) scan_block_sums_kernel(int64_t n, const int* __restrict__ need,
                                                                     int* __restrict__ block_sums) {
  __shared__ int sm[kScanBlock];
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  sm[threadIdx.x] = i < n ? need[i] : 0;
  __syncthreads();
  for (int o = kScanBlock / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = sm[0];
}
__global__ void scan_block_offsets_kernel(int n_blocks, int* __restrict__ block_sums, int* __restrict__ total) {
  // single thread: n_blocks <= ~2M/1024; cheap and deterministic
  int acc = 0;
  for (int b = 0; b < n_blocks; ++b) { const int v = block_sums[b]; block_sums[b] = acc; acc += v; }
  *total = acc;
}
__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(int64_t n, const int* __restrict__ need,
                                                                const int* __restrict__ block_off,
                                                                const int* __restrict__ indptr,
                                                                int* __restrict__ out_indptr) {
  __shared__ int sm[kScanBlock];
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int v = i < n ? need[i] : 0;
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < kScanBlock; o <<= 1) {      // Hillis-Steele inclusive scan
    const int t = (int)threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
    __syncthreads();
    sm[threadIdx.x] += t;
    __syncthreads();
  }
  const int excl = sm[threadIdx.x] - v + block_off[blockIdx.x];
  if (i < n) out_indptr[i] = indptr[i] + excl;
  if (i == n - 1) out_indptr[n] = indptr[n] + excl + v;
}

// one warp per row: copy the columns inserting the diagonal; dinv[r] = 1/sqrt(degree incl. self loop)
__global__ void __launch_bounds__(256) ahat_fill_kernel(int64_t n, const int* __restrict__ indptr,
                                                        const int* __restrict__ indices,
                                                        const int* __restrict__ out_indptr,
                                                        int* __restrict__ out_indices, double* __restrict__ dinv) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const int b = indptr[r], e = indptr[r + 1], ob = out_indptr[r], oe = out_indptr[r + 1];
  const bool insert = (oe - ob) != (e - b);
  for (int k = b + lane; k < e; k += 32) {
    const int c = indices[k];
    out_indices[ob + (k - b) + ((insert && c > (int)r) ? 1 : 0)] = c;
  }
  if (insert && lane == 0) {
    // position of the diagonal = number of columns < r
    int lo = b, hi = e;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (indices[mid] < (int)r) lo = mid + 1; else hi = mid; }
    out_indices[ob + (lo - b)] = (int)r;
  }
  if (lane == 0) {
    const double d = (double)(oe - ob);            // binary adjacency with unit diagonal: exact row sum
    dinv[r] = d > 0.0 ? __ddiv_rn(1.0, __dsqrt_rn(d)) : 0.0;
  }
}

__global__ void __launch_bounds__(256) ahat_vals_kernel(int64_t n, const int* __restrict__ out_indptr,
                                                        const int* __restrict__ out_indices,
                                                        const double* __restrict__ dinv, float* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const double di = dinv[r];
  for (int k = out_indptr[r] + lane; k < out_indptr[r + 1]; k += 32)
    vals[k] = (float)__dmul_rn(__dmul_rn(di, 1.0), dinv[out_indices[k]]);   // (D*adj)*D, then astype(float32)
}

}  // namespace gcg

using namespace gcg;

extern "C" int64_t gcg_ahat_device_workspace_bytes(int64_t n) {
  const int64_t blocks = ceil_div(n > 0 ? n : 1, kScanBlock);
  return 4 * n + 4 * blocks + 16 + 8 * n + 64;
}

// phase 1: out_indptr (int32[n+1]) and the output nnz (device int32 at d_total) for the pattern with unit diagonal
extern "C" int gcg_ahat_indptr_device(int64_t n, const int32_t* d_indptr, const int32_t* d_indices,
                                      int32_t* out_indptr, int32_t* d_total, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
  GCG_CHECK_ARG(n >= 0 && d_indptr && out_indptr && d_total && workspace, "gcg_ahat_indptr_device: NULL argument");
  GCG_CHECK_ARG(workspace_bytes >= gcg_ahat_device_workspace_bytes(n), "gcg_ahat_indptr_device: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) {
    GCG_CUDA(cudaMemsetAsync(out_indptr, 0, 4, st));
    GCG_CUDA(cudaMemsetAsync(d_total, 0, 4, st));
    return GCG_OK;
  }
  int* need = reinterpret_cast<int*>(workspace);
  int* block_sums = need + n;
  const int blocks = (int)ceil_div(n, kScanBlock);
  ahat_mark_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(n, d_indptr, d_indices, need);
  GCG_LAUNCH_CHECK();
  scan_block_sums_kernel<<<blocks, kScanBlock, 0, st>>>(n, need, block_sums);
  GCG_LAUNCH_CHECK();
  scan_block_offsets_kernel<<<1, 1, 0, st>>>(blocks, block_sums, d_total);
  GCG_LAUNCH_CHECK();
  scan_apply_kernel<<<blocks, kScanBlock, 0, st>>>(n, need, block_sums, d_indptr, out_indptr);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

// phase 2: column indices and float32 values of A_hat (outputs sized from out_indptr[n])
extern "C" int gcg_ahat_fill_device(int64_t n, const int32_t* d_indptr, const int32_t* d_indices,
                                    const int32_t* out_indptr, int32_t* out_indices, float* out_vals,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
  GCG_CHECK_ARG(n >= 0 && d_indptr && out_indptr && workspace, "gcg_ahat_fill_device: NULL argument");
  GCG_CHECK_ARG(workspace_bytes >= gcg_ahat_device_workspace_bytes(n), "gcg_ahat_fill_device: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) return GCG_OK;
  const int64_t blocks = ceil_div(n, kScanBlock);
  uintptr_t p = reinterpret_cast<uintptr_t>(workspace) + 4 * n + 4 * blocks + 16;
  double* dinv = reinterpret_cast<double*>((p + 7) & ~uintptr_t(7));
  ahat_fill_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(n, d_indptr, d_indices, out_indptr, out_indices, dinv);
  GCG_LAUNCH_CHECK();
  ahat_vals_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(n, out_indptr, out_indices, dinv, out_vals);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

// ---------------------------------------------------------------------------------------------
// Label pipeline around the hot path: haversine distances in float64.
//   data.py:416-419   dev/test label = argmin_c haversine(point, median_c)   (brute force, first minimum)
//   tensormain.py:38-54  geo_eval: haversine(true location, median of the predicted region)
namespace gcg {
__device__ __forceinline__ double haversine_km(double lat1, double lon1, double lat2, double lon2) {
  const double k = 0.017453292519943295;          // pi / 180
  lat1 *= k; lon1 *= k; lat2 *= k; lon2 *= k;
  const double s1 = sin((lat2 - lat1) * 0.5), s2 = sin((lon2 - lon1) * 0.5);
  const double d = s1 * s1 + cos(lat1) * cos(lat2) * s2 * s2;
  return 2.0 * 6371.0088 * asin(sqrt(d));         // haversine package: AVG_EARTH_RADIUS_KM
}

__global__ void __launch_bounds__(256) haversine_nearest_kernel(const double* __restrict__ pts, int64_t n,
                                                                const double* __restrict__ med, int32_t n_med,
                                                                int64_t* __restrict__ out, double* __restrict__ out_km) {
  extern __shared__ double sm[];                  // medians staged in shared memory
  for (int i = threadIdx.x; i < 2 * n_med; i += blockDim.x) sm[i] = med[i];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double lat = pts[2 * i], lon = pts[2 * i + 1];
  double best = 1e300;
  int64_t arg = 0;
  for (int c = 0; c < n_med; ++c) {
    const double d = haversine_km(lat, lon, sm[2 * c], sm[2 * c + 1]);
    if (d < best) { best = d; arg = c; }          // strict: first minimum wins
  }
  out[i] = arg;
  if (out_km) out_km[i] = best;
}

__global__ void __launch_bounds__(256) haversine_pairs_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                              int64_t n, double* __restrict__ out_km) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out_km[i] = haversine_km(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}
}  // namespace gcg

extern "C" int gcg_haversine_nearest_f64(const double* d_points, int64_t n, const double* d_medians, int32_t n_medians,
                                         int64_t* d_out_idx, double* d_out_km, void* stream) {
  GCG_CHECK_ARG(d_points && d_medians && d_out_idx && n >= 0 && n_medians > 0, "gcg_haversine_nearest_f64: bad argument");
  GCG_CHECK_SHAPE((int64_t)n_medians * 16 <= 200 * 1024, "gcg_haversine_nearest_f64: too many medians for shared memory");
  if (n == 0) return GCG_OK;
  const int smem = n_medians * 16;
  GCG_CUDA(cudaFuncSetAttribute(gcg::haversine_nearest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  gcg::haversine_nearest_kernel<<<(unsigned)gcg::ceil_div(n, 256), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_points, n, d_medians, n_medians, d_out_idx, d_out_km);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_haversine_pairs_f64(const double* d_a, const double* d_b, int64_t n, double* d_out_km, void* stream) {
  GCG_CHECK_ARG(d_a && d_b && d_out_km && n >= 0, "gcg_haversine_pairs_f64: bad argument");
  if (n == 0) return GCG_OK;
  gcg::haversine_pairs_kernel<<<(unsigned)gcg::ceil_div(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_a, d_b, n, d_out_km);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}
