// Peer-memory (NVLink P2P) transposes for the feature-sliced multi-GPU propagation.
//
// Instead of pack -> NCCL all-to-all -> (SpMM) -> NCCL all-to-all -> unpack, every rank WRITES its
// column slices / row groups straight into the peers' buffers with plain global stores on
// IPC-mapped pointers (ld/st over NVLink through the UVA aperture): one kernel per direction,
// no staging copy and no collective launch.  Visibility is ordered by a cross-rank barrier that the
// caller issues after the kernel (a 1-element NCCL all-reduce in stream order).
#include <algorithm>
#include <cstring>

#include "gcg_common.cuh"

namespace gcg {
constexpr int kMaxPeers = 16;
struct PeerTable { float* p[kMaxPeers]; };
struct RowOff { int64_t off[kMaxPeers + 1]; };

// dst_q[(dst_row0 + i) * Fp + c] = src[i, q*Fp + c]  (zero padded beyond F)
__global__ void __launch_bounds__(256) push_cols_kernel(const float* __restrict__ src, int64_t ld, int64_t n, int64_t F,
                                                        int P, int64_t Fp, PeerTable dst, int64_t dst_row0) {
  const int64_t fp4 = Fp >> 2, total = (int64_t)P * n * fp4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // consecutive threads walk one source row across all slices -> coalesced reads, 16 B stores per peer
    const int64_t r = i / (P * fp4), rem = i - r * P * fp4, q = rem / fp4, c4 = rem - q * fp4;
    const int64_t col = q * Fp + c4 * 4;
    const float* s = src + r * ld + col;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col + 3 < F) v = *reinterpret_cast<const float4*>(s);
    else {
      if (col < F) v.x = s[0];
      if (col + 1 < F) v.y = s[1];
      if (col + 2 < F) v.z = s[2];
    }
    reinterpret_cast<float4*>(dst.p[q] + (dst_row0 + r) * Fp)[c4] = v;
  }
  __threadfence_system();
}

// for every owner q: dst_q[slot*slot_stride + i*Fp + c] = src[(off[q] + i) * Fp + c],  i < off[q+1]-off[q]
__global__ void __launch_bounds__(256) push_rows_kernel(const float* __restrict__ src, RowOff ro, int P, int64_t Fp,
                                                        PeerTable dst, int64_t slot_offset_floats) {
  const int64_t fp4 = Fp >> 2, total = ro.off[P] * fp4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / fp4, c4 = i - row * fp4;
    int q = 0;
    while (q + 1 < P && row >= ro.off[q + 1]) ++q;
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<float4*>(dst.p[q] + slot_offset_floats + (row - ro.off[q]) * Fp)[c4] = v;
  }
  __threadfence_system();
}
}  // namespace gcg

using namespace gcg;

extern "C" int gcg_peer_alloc(int64_t bytes, void** d_ptr, void* handle64) {
  GCG_CHECK_ARG(bytes > 0 && d_ptr && handle64, "gcg_peer_alloc: bad argument");
  GCG_CUDA(cudaMalloc(d_ptr, (size_t)bytes));
  GCG_CUDA(cudaMemset(*d_ptr, 0, (size_t)bytes));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  GCG_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), *d_ptr));
  return GCG_OK;
}
extern "C" int gcg_peer_free(void* d_ptr) {
  if (d_ptr) GCG_CUDA(cudaFree(d_ptr));
  return GCG_OK;
}
extern "C" int gcg_peer_open(const void* handle64, void** d_ptr) {
  GCG_CHECK_ARG(handle64 && d_ptr, "gcg_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  GCG_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GCG_OK;
}
extern "C" int gcg_peer_close(void* d_ptr) {
  if (d_ptr) GCG_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return GCG_OK;
}

extern "C" int gcg_push_cols_f32(const float* src, int64_t ld, int64_t n_rows, int64_t F, int32_t P, int64_t Fp,
                                 void* const* h_peer_dst, int64_t dst_row0, void* stream) {
  GCG_CHECK_ARG(src && h_peer_dst && P > 0 && P <= kMaxPeers, "gcg_push_cols_f32: bad argument");
  GCG_CHECK_SHAPE(Fp % 4 == 0 && ld % 4 == 0 && ld >= F && (int64_t)P * Fp >= F && aligned16(src),
                  "gcg_push_cols_f32: needs 16-byte aligned source, ld %% 4 == 0, Fp %% 4 == 0");
  if (n_rows == 0) return GCG_OK;
  PeerTable t;
  for (int q = 0; q < P; ++q) {
    GCG_CHECK_ARG(h_peer_dst[q] && aligned16(h_peer_dst[q]), "gcg_push_cols_f32: peer pointer %d invalid", q);
    t.p[q] = reinterpret_cast<float*>(h_peer_dst[q]);
  }
  const int64_t total = (int64_t)P * n_rows * (Fp / 4);
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16));
  push_cols_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, ld, n_rows, F, P, Fp, t, dst_row0);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_push_rows_f32(const float* src, const int64_t* h_row_off, int32_t P, int64_t Fp,
                                 void* const* h_peer_dst, int64_t slot_offset_floats, void* stream) {
  GCG_CHECK_ARG(src && h_row_off && h_peer_dst && P > 0 && P <= kMaxPeers, "gcg_push_rows_f32: bad argument");
  GCG_CHECK_SHAPE(Fp % 4 == 0 && aligned16(src) && slot_offset_floats % 4 == 0, "gcg_push_rows_f32: alignment");
  PeerTable t;
  RowOff ro;
  for (int q = 0; q < P; ++q) {
    GCG_CHECK_ARG(h_peer_dst[q] && aligned16(h_peer_dst[q]), "gcg_push_rows_f32: peer pointer %d invalid", q);
    t.p[q] = reinterpret_cast<float*>(h_peer_dst[q]);
  }
  for (int q = 0; q <= P; ++q) ro.off[q] = h_row_off[q];
  const int64_t total = ro.off[P] * (Fp / 4);
  if (total == 0) return GCG_OK;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16));
  push_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, ro, P, Fp, t, slot_offset_floats);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}
