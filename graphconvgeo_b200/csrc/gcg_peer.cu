// Peer-memory (NVLink P2P) transposes for the feature-sliced multi-GPU propagation.
//
// Instead of pack -> NCCL all-to-all -> (SpMM) -> NCCL all-to-all -> unpack, every rank WRITES its
// column slices / row groups straight into the peers' buffers through IPC-mapped pointers over
// NVLink: no staging copy and no collective launch.  Visibility is ordered by a cross-rank barrier that the
// caller issues after the kernel (a 1-element NCCL all-reduce in stream order): the stores of a
// finished kernel are performed at system scope before later work of the stream starts.
#include <algorithm>
#include <cstring>

#include "gcg_common.cuh"

namespace gcg {
constexpr int kMaxPeers = 16;
struct PeerTable { float* p[kMaxPeers]; };
struct RowOff { int64_t off[kMaxPeers + 1]; };

// one warp per source row: the row is read once (coalesced) and its P slices are written to the P
// peers as contiguous Fp*4-byte runs of 128-bit stores; dst_q[(dst_row0 + r) * Fp + c] = src[r, q*Fp + c]
__global__ void __launch_bounds__(256) push_cols_kernel(const float* __restrict__ src, int64_t ld, int64_t n, int64_t F,
                                                        int P, int64_t Fp, PeerTable dst, int64_t dst_row0) {
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
      reinterpret_cast<float4*>(dst.p[q] + (dst_row0 + r) * Fp)[c4] = v;
    }
  }
}

// one warp per source row of the column slice: owner q by comparison with the row offsets,
// dst_q[slot_offset + (row - off[q]) * Fp + c] = src[row * Fp + c]
__global__ void __launch_bounds__(256) push_rows_kernel(const float* __restrict__ src, RowOff ro, int P, int64_t Fp,
                                                        PeerTable dst, int64_t slot_offset_floats) {
  const int lane = threadIdx.x & 31;
  const int fp4 = (int)(Fp >> 2);
  const int64_t n = ro.off[P];
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * 8) {
    int q = 0;
    while (q + 1 < P && row >= ro.off[q + 1]) ++q;
    const float4* s = reinterpret_cast<const float4*>(src + row * Fp);
    float4* d = reinterpret_cast<float4*>(dst.p[q] + slot_offset_floats + (row - ro.off[q]) * Fp);
    for (int j = lane; j < fp4; j += 32) d[j] = s[j];
  }
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

// Measured alternatives (4 x B200, 213 MB per push): this store kernel and strided DMA on the copy
// engines (cudaMemcpy2DAsync per peer) both take ~1.1 ms (~200 GB/s per GPU); the kernel keeps the
// following barrier short, so it is the one used.
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
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_rows, 8), (int64_t)kNumSMs * 16));
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
  if (ro.off[P] == 0) return GCG_OK;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(ro.off[P], 8), (int64_t)kNumSMs * 16));
  push_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, ro, P, Fp, t, slot_offset_floats);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}
