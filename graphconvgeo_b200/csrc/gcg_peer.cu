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
struct PeerFlags { int* p[kMaxPeers]; };

// push_cols: blockIdx.y = peer q.  For one peer the destination rows are CONTIGUOUS (dst rows are Fp floats
// wide and consecutive), so a warp writes 32 consecutive float4 of the peer's buffer per step -- whole 128-byte
// lines over NVLink -- and gathers the matching pieces of the source rows (local, cached).  The first version
// walked a source row and let one warp-wide store straddle two or three peers (304-byte runs): 150-200 GB/s per
// GPU inside the epoch against 540-690 GB/s for contiguous runs in the 2-GPU probe (profiles/r01_p2p_probe_2gpu.json).
//   dst_q[(dst_row0 + r) * Fp + c] = src[r, q*Fp + c]   (zero padded beyond F)
__global__ void __launch_bounds__(256) push_cols_kernel(const float* __restrict__ src, int64_t ld, int64_t n, int64_t F,
                                                        int P, int64_t Fp, PeerTable dst, int64_t dst_row0) {
  const int q = blockIdx.y;
  const int fp4 = (int)(Fp >> 2);
  const int64_t total = n * fp4;                       // float4 elements this peer receives
  float4* __restrict__ d = reinterpret_cast<float4*>(dst.p[q] + dst_row0 * Fp);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = j / fp4;
    const int c4 = (int)(j - r * fp4);
    const int64_t col = (int64_t)q * Fp + c4 * 4;
    const float* srow = src + r * ld;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col + 3 < F) v = *reinterpret_cast<const float4*>(srow + col);
    else {
      if (col < F) v.x = srow[col];
      if (col + 1 < F) v.y = srow[col + 1];
      if (col + 2 < F) v.z = srow[col + 2];
    }
    d[j] = v;
  }
}

// push_rows: blockIdx.y = peer q.  The row group of owner q is one contiguous block of the column slice and
// lands contiguously in q's buffer: a flat float4 copy.
//   dst_q[slot_offset + i] = src[off[q] * Fp + i],  i < (off[q+1] - off[q]) * Fp
__global__ void __launch_bounds__(256) push_rows_kernel(const float* __restrict__ src, RowOff ro, int P, int64_t Fp,
                                                        PeerTable dst, int64_t slot_offset_floats) {
  const int q = blockIdx.y;
  const int64_t total = (ro.off[q + 1] - ro.off[q]) * (Fp >> 2);
  const float4* __restrict__ s = reinterpret_cast<const float4*>(src + ro.off[q] * Fp);
  float4* __restrict__ d = reinterpret_cast<float4*>(dst.p[q] + slot_offset_floats);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x)
    d[j] = s[j];
}

// Cross-GPU barrier in stream order without a collective launch: every rank stores `seq` into slot `rank` of
// every peer's flag array (after a system-scope fence, so the peer stores of the kernels before it on this
// stream are visible first) and then waits until all P slots of its OWN array have reached `seq`.
// One warp; lane q talks to peer q.  A rank that never arrives would hang the others, so the wait gives up
// after ~4 s and raises `*err` (checked by the host at the next synchronisation point).
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerFlags peers, volatile int* own, int P, int rank, int seq,
                                                          int* err) {
  const int q = threadIdx.x;
  __threadfence_system();
  if (q < P) {
    volatile int* f = peers.p[q] + rank;
    *f = seq;
  }
  __threadfence_system();
  if (q < P) {
    const long long t0 = clock64();
    while (own[q] - seq < 0) {
      if (clock64() - t0 > 8000000000LL) { atomicExch(err, 1); break; }
      __nanosleep(200);
    }
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
  const int64_t per_peer = n_rows * (Fp >> 2);
  const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(per_peer, 256 * 4), (int64_t)kNumSMs * 4)), (unsigned)P);
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
  int64_t most = 0;
  for (int q = 0; q < P; ++q) most = std::max(most, (ro.off[q + 1] - ro.off[q]) * (Fp >> 2));
  const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(most, 256 * 4), (int64_t)kNumSMs * 4)), (unsigned)P);
  push_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, ro, P, Fp, t, slot_offset_floats);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

extern "C" int gcg_peer_barrier(void* const* h_peer_flags, void* own_flags, int32_t P, int32_t rank, int32_t seq,
                                void* err_flag, void* stream) {
  GCG_CHECK_ARG(h_peer_flags && own_flags && err_flag && P > 0 && P <= kMaxPeers && rank >= 0 && rank < P,
                "gcg_peer_barrier: bad argument");
  PeerFlags f;
  for (int q = 0; q < P; ++q) {
    GCG_CHECK_ARG(h_peer_flags[q] != nullptr, "gcg_peer_barrier: peer pointer %d is NULL", q);
    f.p[q] = reinterpret_cast<int*>(h_peer_flags[q]);
  }
  peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(f, reinterpret_cast<volatile int*>(own_flags), P,
                                                                           rank, seq, reinterpret_cast<int*>(err_flag));
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}
