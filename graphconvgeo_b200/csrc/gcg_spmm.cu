// CSR x dense SpMM for the GCN propagation hot path (sm_100a).
//
// Replaces S.dot(H, activation) / S.dot(input, W) of lasagne_layers.py:26,65,67,84
// and, through A_hat^T = A_hat and a pre-transposed X, their gradients.
//
// Layout / algorithm
//   * A "group" of G lanes (G in {4,8,16,32}) owns one output row inside one
//     column panel; each lane keeps VPL float4 accumulators, so a panel is
//     4*G*VPL floats wide.  Lanes of a group read G*16 contiguous bytes of the
//     gathered dense row per load -> fully coalesced 128-bit loads.
//   * Column indices / values are fetched G at a time (coalesced, streaming,
//     L2 evict_first) and broadcast with group-wide shuffles.
//   * Products are summed in CSR order with separately rounded mul and add
//     (__fmul_rn / __fadd_rn): bit-identical to scipy's csr_matvecs.
//   * Power-law hubs: rows longer than the plan's threshold T are cut into
//     segments of T non-zeros.  Segment blocks are scheduled FIRST in every
//     panel; each writes a partial row to the workspace and a small finalize
//     kernel reduces the partials in segment order (deterministic, no atomics)
//     and applies the epilogue.
//   * Grid is panel-major: all row blocks of panel 0, then panel 1, ... so the
//     set of gathered bytes live at any time is n_cols*panel_bytes, which the
//     caller sizes to stay resident in the 126 MB L2 (evict_last on gathers).
//   * Epilogue fused: +bias, relu/tanh/sigmoid, highway mix g*Hc + (1-g)*H.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "gcg_spmm.cuh"

// Plans of minibatch matrices are created and destroyed thousands of times per fit; cudaMalloc / cudaFree
// cost ~1 ms each (and much more on a busy driver), so the long-row tables come from a small size-bucketed
// cache of device blocks that are only returned to the driver when the cache is full.
namespace {
struct PendingBlock {
  int device;
  size_t cap;
  void* p;
  cudaEvent_t ev;       // last launch that read the block
};
struct PlanBlockCache {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_blocks;   // (device, capacity) -> block
  std::vector<PendingBlock> pending;                          // released, last reader possibly still running
  size_t cached_bytes = 0;                                    // free + pending
  static constexpr size_t kMaxCached = 256u << 20;
} g_plan_blocks;

// move released blocks whose last reader has finished to the free list (caller holds the lock)
void plan_blocks_collect() {
  auto& pd = g_plan_blocks.pending;
  for (size_t i = 0; i < pd.size();) {
    const cudaError_t q = cudaEventQuery(pd[i].ev);
    if (q == cudaErrorNotReady) { ++i; continue; }
    if (q == cudaSuccess) {
      cudaEventDestroy(pd[i].ev);
      g_plan_blocks.free_blocks.insert({{pd[i].device, pd[i].cap}, pd[i].p});
    } else {
      // the event was recorded inside a stream capture (or the context is failing): its completion cannot be
      // observed, so the block is never handed out again (it stays allocated; a rare path)
      (void)cudaGetLastError();
      g_plan_blocks.cached_bytes -= pd[i].cap;
    }
    pd[i] = pd.back();
    pd.pop_back();
  }
}

size_t plan_block_capacity(size_t bytes) {
  size_t cap = 4096;
  while (cap < bytes) cap <<= 1;
  return cap;
}

cudaError_t plan_block_acquire(int device, size_t cap, void** out) {
  {
    std::lock_guard<std::mutex> lk(g_plan_blocks.mu);
    plan_blocks_collect();
    auto it = g_plan_blocks.free_blocks.find({device, cap});
    if (it != g_plan_blocks.free_blocks.end()) {
      *out = it->second;
      g_plan_blocks.free_blocks.erase(it);
      g_plan_blocks.cached_bytes -= cap;
      return cudaSuccess;
    }
  }
  return cudaMalloc(out, cap);
}

// ``last_use``: event recorded after the last launch that read the block (nullptr: never read on the device).
// No synchronisation here -- this runs from CSRMatrix.__del__, possibly while a CUDA graph is being captured.
void plan_block_release(int device, size_t cap, void* p, cudaEvent_t last_use) {
  std::lock_guard<std::mutex> lk(g_plan_blocks.mu);
  if (!last_use) {
    if (g_plan_blocks.cached_bytes + cap <= PlanBlockCache::kMaxCached) {
      g_plan_blocks.free_blocks.insert({{device, cap}, p});
      g_plan_blocks.cached_bytes += cap;
    } else {
      cudaFree(p);          // never used by a kernel: nothing to wait for
    }
    return;
  }
  // over the cache limit the block still waits in the pending list (a cudaFree would synchronise the device);
  // gcg_plan_create_csr trims the free list back under the limit
  g_plan_blocks.pending.push_back({device, cap, p, last_use});
  g_plan_blocks.cached_bytes += cap;
}

// called from gcg_plan_create_csr (a synchronising call anyway): give memory back when the cache is over its limit
void plan_blocks_trim() {
  std::lock_guard<std::mutex> lk(g_plan_blocks.mu);
  plan_blocks_collect();
  while (g_plan_blocks.cached_bytes > PlanBlockCache::kMaxCached && !g_plan_blocks.free_blocks.empty()) {
    auto it = g_plan_blocks.free_blocks.begin();
    g_plan_blocks.cached_bytes -= it->first.second;
    cudaFree(it->second);
    g_plan_blocks.free_blocks.erase(it);
  }
}
}  // namespace

namespace gcg {

// U = gathered rows in flight per group, MINB = resident CTAs per SM the register budget must allow
// defaults from the B200 sweep (profiles/r01_spmm_notes.md): no spills, most bytes in flight per SM
template <int G, int VPL, int U>
__device__ __forceinline__ void spmm_vec_body(const SpmmArgs& a) {
  constexpr int RPW = 32 / G;
  constexpr int RPB = 8 * RPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / G, s = lane % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  const int bpp = a.seg_blocks + a.row_blocks;
  const int panel = blockIdx.x / bpp;
  const int bip = blockIdx.x - panel * bpp;
  const uint64_t keep = policy_evict_last(), strm = policy_evict_first();

  int beg, end;
  int64_t row;
  bool is_seg;
  if (bip < a.seg_blocks) {
    const int seg = bip * RPB + warp * RPW + grp;
    if (seg >= a.n_seg) return;
    const int r = __ldg(a.seg_row + seg);
    beg = __ldg(a.seg_beg + seg);
    end = min(beg + a.long_thresh, __ldg(a.indptr + r + 1));
    row = seg;
    is_seg = true;
  } else {
    const int r = (bip - a.seg_blocks) * RPB + warp * RPW + grp;
    if (r >= a.n_rows) return;
    beg = __ldg(a.indptr + r);
    end = __ldg(a.indptr + r + 1);
    if (a.only_segments || end - beg > a.long_thresh) return;  // segment blocks + finalize own this row
    // plain accumulation of an empty row is a no-op: no read-modify-write (column-blocked products
    // visit every row once per block and most (row, block) pairs are empty)
    if (beg == end && a.accumulate && !a.bias && a.act == GCG_ACT_IDENTITY && !a.gate) return;
    row = r;
    is_seg = false;
  }

  const int c_base = panel * a.panel_f4 + s;
  bool cv[VPL];
  float4 acc[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    cv[k] = (c_base + k * G) < a.f4_total;
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4* __restrict__ Bp = reinterpret_cast<const float4*>(a.B) + c_base;
  const int64_t ldb4 = a.ldb >> 2;

  for (int base = beg; base < end; base += G) {
    const int kk = base + s;
    int myc = 0;
    float myv = 0.f;
    if (kk < end) {
      myc = ldg_i32_stream(a.indices + kk, strm);
      myv = ldg_f32_stream(a.vals + kk, strm);
    }
    const int cnt = min(G, end - base);
    int j = 0;
    for (; j + U <= cnt; j += U) {
      float4 x[U][VPL];
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int col = __shfl_sync(gmask, myc, j + u, G);
        v[u] = __shfl_sync(gmask, myv, j + u, G);
        const float4* src = Bp + (int64_t)col * ldb4;
#pragma unroll
        for (int k = 0; k < VPL; ++k)
          if (cv[k]) x[u][k] = ldg_f4_keep(src + k * G, keep);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < VPL; ++k)
          if (cv[k]) acc[k] = f4_axpy_exact(acc[k], v[u], x[u][k]);
    }
    for (; j < cnt; ++j) {
      const int col = __shfl_sync(gmask, myc, j, G);
      const float v = __shfl_sync(gmask, myv, j, G);
      const float4* src = Bp + (int64_t)col * ldb4;
      float4 x[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k)
        if (cv[k]) x[k] = ldg_f4_keep(src + k * G, keep);
#pragma unroll
      for (int k = 0; k < VPL; ++k)
        if (cv[k]) acc[k] = f4_axpy_exact(acc[k], v, x[k]);
    }
  }

  if (is_seg) {
    float4* out = reinterpret_cast<float4*>(a.part + row * a.ldc) + c_base;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (cv[k]) out[k * G] = acc[k];  // re-read soon by finalize: default policy
  } else {
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (cv[k]) epilogue_store(a, row, c_base + k * G, acc[k], strm);
  }
}

template <int G, int VPL, int U = ((VPL >= 2) ? 2 : 4), int MINB = ((VPL >= 3) ? 2 : 4)>
__global__ void __launch_bounds__(256, MINB) spmm_vec_kernel(const SpmmArgs a) {
  spmm_vec_body<G, VPL, U>(a);
}
// same body, register budget left to the compiler's own heuristic (no min-blocks hint)
template <int G, int VPL, int U>
__global__ void __launch_bounds__(256) spmm_vec_kernel_nb(const SpmmArgs a) {
  spmm_vec_body<G, VPL, U>(a);
}

// One warp per long row: sum the segment partials in order, apply the epilogue.
__global__ void __launch_bounds__(256) spmm_finalize_kernel(const SpmmArgs a) {
  const int lane = threadIdx.x & 31;
  const int li = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (li >= a.n_long) return;
  const uint64_t strm = policy_evict_first();
  const int row = __ldg(a.long_rows + li);
  const int s0 = __ldg(a.long_segptr + li), s1 = __ldg(a.long_segptr + li + 1);
  for (int c4 = lane; c4 < a.f4_total; c4 += 32) {
    float4 acc = *(reinterpret_cast<const float4*>(a.part + (int64_t)s0 * a.ldc) + c4);
    for (int sgi = s0 + 1; sgi < s1; ++sgi) {
      const float4 p = *(reinterpret_cast<const float4*>(a.part + (int64_t)sgi * a.ldc) + c4);
      acc.x = __fadd_rn(acc.x, p.x); acc.y = __fadd_rn(acc.y, p.y);
      acc.z = __fadd_rn(acc.z, p.z); acc.w = __fadd_rn(acc.w, p.w);
    }
    epilogue_store(a, row, c4, acc, strm);
  }
}

// Scalar fallback: unaligned pointers or leading dimensions that are not
// multiples of 4.  One warp per row, 32*8 columns per pass, no row splitting.
__global__ void __launch_bounds__(256) spmm_scalar_kernel(const SpmmArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= a.n_rows) return;
  const int beg = __ldg(a.indptr + row), end = __ldg(a.indptr + row + 1);
  for (int64_t c0 = 0; c0 < a.F; c0 += 256) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int base = beg; base < end; base += 32) {
      const int kk = base + lane;
      const int myc = kk < end ? __ldg(a.indices + kk) : 0;
      const float myv = kk < end ? __ldg(a.vals + kk) : 0.f;
      const int cnt = min(32, end - base);
      for (int j = 0; j < cnt; ++j) {
        const int col = __shfl_sync(0xffffffffu, myc, j);
        const float v = __shfl_sync(0xffffffffu, myv, j);
        const float* src = a.B + (int64_t)col * a.ldb + c0 + lane;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (c0 + lane + 32 * k < a.F) acc[k] = __fadd_rn(acc[k], __fmul_rn(v, __ldg(src + 32 * k)));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t c = c0 + lane + 32 * k;
      if (c >= a.F) continue;
      float v = acc[k];
      float* cp = out_row_ptr(a, row) + c;
      if (a.accumulate) v = __fadd_rn(*cp, v);
      if (a.bias) v = __fadd_rn(v, __ldg(a.bias + c));
      v = apply_act(v, a.act);
      if (a.gate) {
        if (a.conv_out) a.conv_out[row * a.ld_conv + c] = v;
        v = gate_mix(a.gate[row * a.ld_gate + c], v, a.carry[row * a.ld_carry + c]);
      }
      *cp = v;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Bulk-copy (TMA) staged variant: whole rows per warp, gathered rows land in shared memory.
//
// The register-gather kernel above can only keep U*VPL 16-byte loads per lane in flight (126
// registers -> 16 warps/SM -> ~50 KB in flight per SM, ncu: 16 % warps active, DRAM 43 % busy).
// Here every gathered dense row (F*4 contiguous bytes) is fetched by ONE cp.async.bulk
// (global -> shared, mbarrier complete_tx) issued by lane 0, into a per-warp ring of RING slots:
// 8 warps x 8 slots x 2400 B = 154 KB of gathers in flight per SM with no register cost.  A warp
// owns `rows_per_warp` CONSECUTIVE rows, so its non-zeros are one contiguous CSR slice that is
// streamed without draining the ring at row boundaries; column indices / values are fetched 32 at
// a time (plus a prefetched next chunk) and broadcast by shuffle.  Accumulation order and rounding
// are identical to the register kernel (CSR order, __fmul_rn/__fadd_rn) -> bit-identical results.
// Long rows are skipped here; their segments run in spmm_vec_kernel (only_segments) + finalize.
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sm_u32(dst)), "l"(src), "r"(bytes), "r"(sm_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t"
      "}" ::"r"(sm_u32(bar)), "r"(parity) : "memory");
}

// first non-empty, non-long row at or after r (warp-uniform)
__device__ __forceinline__ bool seek_row(const SpmmArgs& a, int& r, int& k, int& kend, int r1) {
  while (r < r1) {
    const int b = __ldg(a.indptr + r), e = __ldg(a.indptr + r + 1);
    if (e > b && (e - b) <= a.long_thresh) { k = b; kend = e; return true; }
    ++r;
  }
  return false;
}

template <int VPL>
__global__ void __launch_bounds__(256, 1)
spmm_tma_kernel(const SpmmArgs a, const int rows_per_warp, const int ring, const int slot_bytes) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* slots = smem_dyn + (size_t)warp * ring * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dyn + (size_t)8 * ring * slot_bytes) + warp * ring;
  if (lane == 0) {
    for (int s = 0; s < ring; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm_u32(bars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int r0 = (blockIdx.x * 8 + warp) * rows_per_warp;
  const int r1 = min(a.n_rows, r0 + rows_per_warp);
  if (r0 >= a.n_rows) return;
  const uint64_t strm = policy_evict_first();
  const uint32_t row_bytes = (uint32_t)a.f4_total * 16u;
  bool cv[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) cv[j] = (lane + 32 * j) < a.f4_total;

  // ---- producer cursor (column indices) with a prefetched next chunk
  int pr = r0, pk = 0, pend = 0;
  bool pvalid = seek_row(a, pr, pk, pend, r1);
  int pbase = pvalid ? pk : 0;
  auto ld_idx = [&](int base) { const int q = base + lane; return q < a.nnz_total ? ldg_i32_stream(a.indices + q, strm) : 0; };
  auto ld_val = [&](int base) { const int q = base + lane; return q < a.nnz_total ? ldg_f32_stream(a.vals + q, strm) : 0.f; };
  int pidx = ld_idx(pbase), pidx_nx = ld_idx(pbase + 32);
  int issued = 0, consumed = 0;
  auto issue_one = [&]() {
    if (pk >= pbase + 32 || pk < pbase) {
      if (pk >= pbase + 32 && pk < pbase + 64) { pbase += 32; pidx = pidx_nx; }
      else { pbase = pk; pidx = ld_idx(pbase); }
      pidx_nx = ld_idx(pbase + 32);
    }
    const int col = __shfl_sync(0xffffffffu, pidx, pk - pbase);
    const int slot = issued % ring;
    if (lane == 0) bulk_g2s(slots + (size_t)slot * slot_bytes, a.B + (int64_t)col * a.ldb, row_bytes, bars + slot);
    ++issued;
    if (++pk == pend) { ++pr; pvalid = seek_row(a, pr, pk, pend, r1); }
  };
  while (pvalid && issued < ring) issue_one();

  // ---- consumer
  int cbase = -64;
  float cval = 0.f, cval_nx = 0.f;
  for (int r = r0; r < r1; ++r) {
    const int b = __ldg(a.indptr + r), e = __ldg(a.indptr + r + 1);
    if (e - b > a.long_thresh) continue;                 // segments + finalize own this row
    float4 acc[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = b; k < e; ++k) {
      if (k >= cbase + 32 || k < cbase) {
        if (k >= cbase + 32 && k < cbase + 64) { cbase += 32; cval = cval_nx; }
        else { cbase = k; cval = ld_val(cbase); }
        cval_nx = ld_val(cbase + 32);
      }
      const float v = __shfl_sync(0xffffffffu, cval, k - cbase);
      const int slot = consumed % ring;
      bar_wait(bars + slot, (uint32_t)((consumed / ring) & 1));
      const float4* src = reinterpret_cast<const float4*>(slots + (size_t)slot * slot_bytes) + lane;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (cv[j]) acc[j] = f4_axpy_exact(acc[j], v, src[32 * j]);
      ++consumed;
      __syncwarp();                                      // every lane is done with the slot
      if (pvalid) issue_one();                           // refill it (issued % ring == this slot)
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j)
      if (cv[j]) epilogue_store(a, r, lane + 32 * j, acc[j], strm);
  }
}

template <int VPL>
static cudaError_t launch_tma(const SpmmArgs& a, int rows_per_warp, int ring, int slot_bytes, cudaStream_t st) {
  const int smem_bytes = 8 * ring * slot_bytes + 8 * ring * 8;
  cudaError_t e = cudaFuncSetAttribute(spmm_tma_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)ceil_div(a.n_rows, 8 * rows_per_warp);
  spmm_tma_kernel<VPL><<<grid, 256, smem_bytes, st>>>(a, rows_per_warp, ring, slot_bytes);
  return cudaGetLastError();
}

cudaError_t spmm_finalize_launch(const SpmmArgs& a, int n_long, cudaStream_t st) {
  spmm_finalize_kernel<<<(unsigned)ceil_div(n_long, 8), 256, 0, st>>>(a);
  return cudaGetLastError();
}

static int g_tune_u = 0, g_tune_minb = 0;   // experiment knobs (gcg_spmm_set_tuning); 0 = defaults
static int g_tune_g = 0, g_tune_vpl = 0;    // experiment knob (gcg_spmm_set_group): lanes per row x float4 per lane

template <int G, int VPL>
static cudaError_t launch_vec(const SpmmArgs& a, cudaStream_t st) {
  const int64_t grid = (int64_t)a.n_panels * (a.seg_blocks + a.row_blocks);
  if (G == 32 && g_tune_u > 0) {
#define GCG_TUNE(UU, MB) if (g_tune_u == UU && g_tune_minb == MB) { spmm_vec_kernel<32, VPL, UU, MB><<<(unsigned)grid, 256, 0, st>>>(a); return cudaGetLastError(); }
    GCG_TUNE(2, 2) GCG_TUNE(2, 3) GCG_TUNE(3, 2) GCG_TUNE(4, 2) GCG_TUNE(2, 4)
#undef GCG_TUNE
    if (g_tune_minb == 0 && g_tune_u == 2) { spmm_vec_kernel_nb<32, VPL, 2><<<(unsigned)grid, 256, 0, st>>>(a); return cudaGetLastError(); }
    if (g_tune_minb == 0 && g_tune_u == 4) { spmm_vec_kernel_nb<32, VPL, 4><<<(unsigned)grid, 256, 0, st>>>(a); return cudaGetLastError(); }
  }
  // A/B on B200 (profiles/r01_spmm_notes.md): whole 600-wide rows run 14 % faster when the register
  // budget is left to the compiler (126 regs) than under a (256, 2) bound; F = 256 is fastest at 64 regs
  if (G == 32 && VPL >= 3) spmm_vec_kernel_nb<32, VPL, 2><<<(unsigned)grid, 256, 0, st>>>(a);
  else spmm_vec_kernel<G, VPL><<<(unsigned)grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gcg

using namespace gcg;

extern "C" void gcg_spmm_set_tuning(int u, int minb) { gcg::g_tune_u = u; gcg::g_tune_minb = minb; }
extern "C" void gcg_spmm_set_group(int lanes, int vpl) { gcg::g_tune_g = lanes; gcg::g_tune_vpl = vpl; }

extern "C" int gcg_plan_create_csr(int64_t n_rows, int64_t n_cols, int64_t nnz,
                                   const int32_t* d_indptr, const int32_t* d_indices,
                                   const float* d_vals, const int32_t* h_indptr,
                                   int32_t long_row_threshold, gcg_plan** out) {
  GCG_CHECK_ARG(out != nullptr, "gcg_plan_create_csr: out is NULL");
  GCG_CHECK_ARG(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "gcg_plan_create_csr: negative size");
  GCG_CHECK_ARG(n_rows < INT32_MAX && n_cols < INT32_MAX && nnz < INT32_MAX,
                "gcg_plan_create_csr: int32 CSR indices overflow");
  GCG_CHECK_ARG(d_indptr && (nnz == 0 || (d_indices && d_vals)), "gcg_plan_create_csr: NULL CSR array");
  if (long_row_threshold <= 0) long_row_threshold = 256;
  std::vector<int32_t> hp;
  if (!h_indptr) {
    hp.resize(n_rows + 1);
    GCG_CUDA(cudaMemcpy(hp.data(), d_indptr, sizeof(int32_t) * (n_rows + 1), cudaMemcpyDeviceToHost));
    h_indptr = hp.data();
  }
  // a plan may describe a contiguous row slice of a larger CSR (absolute offsets)
  GCG_CHECK_SHAPE(h_indptr[0] >= 0 && h_indptr[n_rows] <= nnz,
                  "gcg_plan_create_csr: indptr[0]=%d indptr[n]=%d but nnz=%lld", h_indptr[0],
                  h_indptr[n_rows], (long long)nnz);
  std::vector<int32_t> seg_row, seg_beg, long_rows, long_segptr;
  int64_t max_deg = 0;
  const int32_t T = long_row_threshold;
  for (int64_t r = 0; r < n_rows; ++r) {
    const int32_t b = h_indptr[r], e = h_indptr[r + 1];
    GCG_CHECK_SHAPE(e >= b, "gcg_plan_create_csr: indptr not monotone at row %lld", (long long)r);
    const int32_t deg = e - b;
    max_deg = std::max<int64_t>(max_deg, deg);
    if (deg > T) {
      long_rows.push_back((int32_t)r);
      long_segptr.push_back((int32_t)seg_row.size());
      for (int32_t s = b; s < e; s += T) {
        seg_row.push_back((int32_t)r);
        seg_beg.push_back(s);
      }
    }
  }
  long_segptr.push_back((int32_t)seg_row.size());
  gcg_plan* p = new gcg_plan();
  p->n_rows = n_rows; p->n_cols = n_cols; p->nnz = nnz;
  p->indptr = d_indptr; p->indices = d_indices; p->vals = d_vals;
  p->long_thresh = T;
  p->n_long = (int64_t)long_rows.size();
  p->n_seg = (int64_t)seg_row.size();
  p->max_deg = max_deg;
  p->d_seg_row = p->d_seg_beg = p->d_long_rows = p->d_long_segptr = nullptr;
  p->d_block = nullptr;
  p->block_cap = 0;
  p->device = 0;
  p->stream = nullptr;
  p->last_use = nullptr;
  plan_blocks_trim();
  p->h_indptr.assign(h_indptr, h_indptr + n_rows + 1);
  if (p->n_long > 0) {
    // one block, one upload: [seg_row | seg_beg | long_rows | long_segptr]
    std::vector<int32_t> packed;
    packed.reserve(2 * seg_row.size() + 2 * long_rows.size() + 1);
    packed.insert(packed.end(), seg_row.begin(), seg_row.end());
    packed.insert(packed.end(), seg_beg.begin(), seg_beg.end());
    packed.insert(packed.end(), long_rows.begin(), long_rows.end());
    packed.insert(packed.end(), long_segptr.begin(), long_segptr.end());
    const size_t bytes = packed.size() * sizeof(int32_t);
    cudaError_t e = cudaGetDevice(&p->device);
    p->block_cap = plan_block_capacity(bytes);
    if (e == cudaSuccess) e = plan_block_acquire(p->device, p->block_cap, &p->d_block);
    // synchronous copy: also orders the upload after any kernel of a previous owner of the block
    if (e == cudaSuccess) e = cudaMemcpy(p->d_block, packed.data(), bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("gcg_plan_create_csr: %s", cudaGetErrorString(e));
      gcg_plan_destroy(p);
      return GCG_ERR_CUDA;
    }
    int32_t* base = reinterpret_cast<int32_t*>(p->d_block);
    p->d_seg_row = base;
    p->d_seg_beg = base + seg_row.size();
    p->d_long_rows = base + 2 * seg_row.size();
    p->d_long_segptr = base + 2 * seg_row.size() + long_rows.size();
  }
  *out = p;
  return GCG_OK;
}

extern "C" int gcg_plan_destroy(gcg_plan* p) {
  if (!p) return GCG_OK;
  if (p->stream) { gcg::stream_state_destroy(p->stream); p->stream = nullptr; }
  if (p->d_block) {
    // kernels still reading the tables must finish before the block can serve another plan: the block waits in the
    // pool's pending list until the event recorded after the plan's last launch has completed
    plan_block_release(p->device, p->block_cap, p->d_block, p->last_use);
  } else if (p->last_use) {
    cudaEventDestroy(p->last_use);
  }
  delete p;
  return GCG_OK;
}

extern "C" int64_t gcg_plan_workspace_bytes(const gcg_plan* p, int64_t ldc) {
  if (!p || ldc <= 0) return 0;
  return p->n_seg * ldc * (int64_t)sizeof(float);
}

extern "C" int gcg_plan_info(const gcg_plan* p, int64_t* info) {
  GCG_CHECK_ARG(p && info, "gcg_plan_info: NULL argument");
  info[0] = p->n_rows; info[1] = p->n_cols; info[2] = p->nnz; info[3] = p->n_long;
  info[4] = p->n_seg; info[5] = p->max_deg; info[6] = p->long_thresh; info[7] = 0;
  return GCG_OK;
}

struct OwnerRoute {
  int n_owner;
  void* const* base;
  const int64_t* off;
};

static int spmm_run_impl(const gcg_plan* p, const float* B, int64_t ldb, int64_t F,
                         float* C, int64_t ldc, const float* bias, int act,
                         int accumulate, const float* gate, int64_t ld_gate,
                         const float* carry, int64_t ld_carry, float* conv_out,
                         int64_t ld_conv, int32_t panel_cols, void* workspace,
                         int64_t workspace_bytes, void* stream, const OwnerRoute* route);

static int spmm_run(const gcg_plan* p, const float* B, int64_t ldb, int64_t F,
                    float* C, int64_t ldc, const float* bias, int act,
                    int accumulate, const float* gate, int64_t ld_gate,
                    const float* carry, int64_t ld_carry, float* conv_out,
                    int64_t ld_conv, int32_t panel_cols, void* workspace,
                    int64_t workspace_bytes, void* stream, const OwnerRoute* route) {
  const int rc = spmm_run_impl(p, B, ldb, F, C, ldc, bias, act, accumulate, gate, ld_gate, carry, ld_carry, conv_out,
                               ld_conv, panel_cols, workspace, workspace_bytes, stream, route);
  if (rc == GCG_OK && p && p->d_block) {
    // fence for the pooled long-row tables (see plan_block_release)
    if (!p->last_use && cudaEventCreateWithFlags(&p->last_use, cudaEventDisableTiming) != cudaSuccess) {
      p->last_use = nullptr;
      GCG_CUDA(cudaGetLastError());
    }
    if (p->last_use) GCG_CUDA(cudaEventRecord(p->last_use, reinterpret_cast<cudaStream_t>(stream)));
  }
  return rc;
}

extern "C" int gcg_spmm_csr_f32(const gcg_plan* p, const float* B, int64_t ldb, int64_t F,
                                float* C, int64_t ldc, const float* bias, int act,
                                int accumulate, const float* gate, int64_t ld_gate,
                                const float* carry, int64_t ld_carry, float* conv_out,
                                int64_t ld_conv, int32_t panel_cols, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  GCG_RECORD("gcg_spmm_csr_f32", gcg_spmm_csr_f32(p, B, ldb, F, C, ldc, bias, act, accumulate, gate, ld_gate, carry, ld_carry, conv_out, ld_conv, panel_cols, workspace, workspace_bytes, s__));
  return spmm_run(p, B, ldb, F, C, ldc, bias, act, accumulate, gate, ld_gate, carry, ld_carry, conv_out, ld_conv,
                  panel_cols, workspace, workspace_bytes, stream, nullptr);
}

extern "C" int gcg_spmm_csr_routed_f32(const gcg_plan* p, const float* B, int64_t ldb, int64_t F, int32_t n_owner,
                                       void* const* h_owner_base, const int64_t* h_owner_row_off, int64_t ldc,
                                       const float* bias, int act, int32_t panel_cols, void* workspace,
                                       int64_t workspace_bytes, void* stream) {
  GCG_CHECK_ARG(p != nullptr, "gcg_spmm_csr_routed_f32: plan is NULL");
  GCG_CHECK_ARG(n_owner > 0 && n_owner <= 16 && h_owner_base && h_owner_row_off, "gcg_spmm_csr_routed_f32: bad owner table");
  GCG_CHECK_SHAPE(h_owner_row_off[0] == 0 && h_owner_row_off[n_owner] == p->n_rows,
                  "gcg_spmm_csr_routed_f32: owner offsets must cover rows [0, %lld)", (long long)p->n_rows);
  for (int q = 0; q < n_owner; ++q) {
    GCG_CHECK_ARG(h_owner_base[q] != nullptr && aligned16(h_owner_base[q]), "gcg_spmm_csr_routed_f32: owner %d buffer invalid", q);
    GCG_CHECK_SHAPE(h_owner_row_off[q + 1] >= h_owner_row_off[q], "gcg_spmm_csr_routed_f32: offsets not monotone");
  }
  OwnerRoute r{n_owner, h_owner_base, h_owner_row_off};
  // C is only used for argument validation below: any owner buffer will do
  return spmm_run(p, B, ldb, F, reinterpret_cast<float*>(h_owner_base[0]), ldc, bias, act, 0, nullptr, 0, nullptr, 0, nullptr, 0,
                  panel_cols, workspace, workspace_bytes, stream, &r);
}

static int spmm_run_impl(const gcg_plan* p, const float* B, int64_t ldb, int64_t F,
                         float* C, int64_t ldc, const float* bias, int act,
                         int accumulate, const float* gate, int64_t ld_gate,
                         const float* carry, int64_t ld_carry, float* conv_out,
                         int64_t ld_conv, int32_t panel_cols, void* workspace,
                         int64_t workspace_bytes, void* stream, const OwnerRoute* route) {
  GCG_CHECK_ARG(p != nullptr, "gcg_spmm_csr_f32: plan is NULL");
  GCG_CHECK_ARG(B && C, "gcg_spmm_csr_f32: NULL dense operand");
  GCG_CHECK_ARG(B != C, "gcg_spmm_csr_f32: B and C must not alias");
  GCG_CHECK_SHAPE(F > 0 && ldb >= F && ldc >= F, "gcg_spmm_csr_f32: F=%lld ldb=%lld ldc=%lld",
                  (long long)F, (long long)ldb, (long long)ldc);
  GCG_CHECK_ARG(act >= GCG_ACT_IDENTITY && act <= GCG_ACT_SIGMOID, "gcg_spmm_csr_f32: bad act %d", act);
  GCG_CHECK_ARG((gate == nullptr) == (carry == nullptr), "gcg_spmm_csr_f32: gate and carry go together");
  if (gate) {
    GCG_CHECK_SHAPE(ld_gate >= F && ld_carry >= F && (!conv_out || ld_conv >= F),
                    "gcg_spmm_csr_f32: gate/carry/conv leading dimensions smaller than F");
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p->n_rows == 0) return GCG_OK;

  SpmmArgs a;
  a.indptr = p->indptr; a.indices = p->indices; a.vals = p->vals;
  a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc; a.F = F;
  a.n_rows = (int)p->n_rows; a.long_thresh = p->long_thresh;
  a.seg_row = p->d_seg_row; a.seg_beg = p->d_seg_beg; a.n_seg = (int)p->n_seg;
  a.long_rows = p->d_long_rows; a.long_segptr = p->d_long_segptr; a.n_long = (int)p->n_long;
  a.part = reinterpret_cast<float*>(workspace);
  a.bias = bias; a.act = act; a.accumulate = accumulate;
  a.gate = gate; a.ld_gate = ld_gate; a.carry = carry; a.ld_carry = ld_carry;
  a.conv_out = conv_out; a.ld_conv = ld_conv;
  a.f4_total = (int)((F + 3) / 4);
  a.n_owner = 0;
  if (route) {
    a.n_owner = route->n_owner;
    for (int q = 0; q < route->n_owner; ++q) {
      a.owner_base[q] = reinterpret_cast<float*>(route->base[q]);
      a.owner_off[q] = (int)route->off[q];
    }
    a.owner_off[route->n_owner] = (int)route->off[route->n_owner];
  }

  const int64_t f_pad = 4 * (int64_t)a.f4_total;
  bool vec = aligned16(B) && aligned16(C) && (ldb % 4 == 0) && (ldc % 4 == 0) && ldb >= f_pad && ldc >= f_pad;
  if (gate)
    vec = vec && aligned16(gate) && aligned16(carry) && (ld_gate % 4 == 0) && (ld_carry % 4 == 0) &&
          ld_gate >= f_pad && ld_carry >= f_pad &&
          (!conv_out || (aligned16(conv_out) && ld_conv % 4 == 0 && ld_conv >= f_pad));
  if (!vec) {
    a.n_seg = 0; a.n_long = 0; a.long_thresh = INT32_MAX;
    a.panel_f4 = a.f4_total; a.n_panels = 1; a.seg_blocks = 0; a.row_blocks = 0;
    spmm_scalar_kernel<<<(unsigned)ceil_div(p->n_rows, 8), 256, 0, st>>>(a);
    GCG_LAUNCH_CHECK();
    return GCG_OK;
  }
  if (p->n_seg > 0) {
    GCG_CHECK_ARG(workspace != nullptr && workspace_bytes >= gcg_plan_workspace_bytes(p, ldc),
                  "gcg_spmm_csr_f32: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                  (long long)gcg_plan_workspace_bytes(p, ldc));
    GCG_CHECK_ARG(aligned16(workspace), "gcg_spmm_csr_f32: workspace must be 16-byte aligned");
  }

  a.nnz_total = (int)p->nnz;
  a.only_segments = 0;
  if (panel_cols == -2) {
    // nnz-balanced streaming variant (gcg_spmm_stream.cu); shapes it does not cover fall through
    const int rc = spmm_stream_launch(p, a, st);
    if (rc < 0) return rc;
    if (rc == 1) return GCG_OK;
    panel_cols = 0;
  }
  if (panel_cols < 0 && a.f4_total <= 256) {
    // bulk-copy staged variant: rows here, long-row segments in the register kernel
    const int slot_bytes = a.f4_total * 16;
    int ring = std::min(8, (200 * 1024 / 8) / slot_bytes);
    if (ring >= 2) {
      const int vpl = (int)ceil_div(a.f4_total, 32);
      const int rows_per_warp = 16;
      cudaError_t e2;
      if (p->n_seg > 0) {
        SpmmArgs sa = a;
        sa.only_segments = 1;
        sa.panel_f4 = 32 * std::min(5, vpl);
        sa.n_panels = (int)ceil_div(a.f4_total, sa.panel_f4);
        sa.seg_blocks = (int)ceil_div(p->n_seg, 8);
        sa.row_blocks = 0;
        switch (std::min(5, vpl)) {
          case 1: e2 = launch_vec<32, 1>(sa, st); break;
          case 2: e2 = launch_vec<32, 2>(sa, st); break;
          case 3: e2 = launch_vec<32, 3>(sa, st); break;
          case 4: e2 = launch_vec<32, 4>(sa, st); break;
          default: e2 = launch_vec<32, 5>(sa, st); break;
        }
        if (e2 != cudaSuccess) { set_error("gcg_spmm_csr_f32: segment launch failed: %s", cudaGetErrorString(e2)); return GCG_ERR_CUDA; }
        count_launch();
      }
      switch (vpl) {
        case 1: e2 = launch_tma<1>(a, rows_per_warp, ring, slot_bytes, st); break;
        case 2: e2 = launch_tma<2>(a, rows_per_warp, ring, slot_bytes, st); break;
        case 3: e2 = launch_tma<3>(a, rows_per_warp, ring, slot_bytes, st); break;
        case 4: e2 = launch_tma<4>(a, rows_per_warp, ring, slot_bytes, st); break;
        case 5: e2 = launch_tma<5>(a, rows_per_warp, ring, slot_bytes, st); break;
        case 6: e2 = launch_tma<6>(a, rows_per_warp, ring, slot_bytes, st); break;
        case 7: e2 = launch_tma<7>(a, rows_per_warp, ring, slot_bytes, st); break;
        default: e2 = launch_tma<8>(a, rows_per_warp, ring, slot_bytes, st); break;
      }
      if (e2 != cudaSuccess) { set_error("gcg_spmm_csr_f32: bulk-copy launch failed: %s", cudaGetErrorString(e2)); return GCG_ERR_CUDA; }
      count_launch();
      if (p->n_long > 0) {
        spmm_finalize_kernel<<<(unsigned)ceil_div(p->n_long, 8), 256, 0, st>>>(a);
        GCG_LAUNCH_CHECK();
      }
      return GCG_OK;
    }
  }
  if (panel_cols < 0) panel_cols = 0;
  // choose (G, VPL): panel width in float4 = G*VPL
  int want_f4 = a.f4_total;
  if (panel_cols > 0) want_f4 = (int)std::min<int64_t>(a.f4_total, std::max<int64_t>(1, (panel_cols + 3) / 4));
  // (G lanes per row, VPL float4 per lane): the narrowest power-of-two group that covers the panel
  int G, VPL;
  if (want_f4 <= 4) { G = 4; VPL = 1; }
  else if (want_f4 <= 8) { G = 8; VPL = 1; }
  else if (want_f4 <= 16) { G = 16; VPL = 1; }
  else { G = 32; VPL = (int)std::min<int64_t>(5, ceil_div(want_f4, 32)); }   // (8,3)/(16,2) groups measured slower than (32,1) at f4 = 19
  // narrow operands (the F/P column slices of the multi-GPU propagation: 19 float4 at P = 8): a 32-lane group
  // leaves 13 lanes idle; 4 lanes x 5 float4 (8 rows per warp) use all of them
  static const bool narrow_groups = getenv("GCG_SPMM_NARROW_GROUPS") ? atoi(getenv("GCG_SPMM_NARROW_GROUPS")) != 0 : false;
  if (narrow_groups && panel_cols == 0 && want_f4 > 16 && want_f4 <= 20) { G = 4; VPL = 5; }
  if (panel_cols == 0 && g_tune_g > 0 && g_tune_g * g_tune_vpl >= want_f4) { G = g_tune_g; VPL = g_tune_vpl; }
  a.panel_f4 = G * VPL;
  a.n_panels = (int)ceil_div(a.f4_total, a.panel_f4);
  const int rpb = 8 * (32 / G);
  a.seg_blocks = (int)ceil_div(p->n_seg, rpb);
  a.row_blocks = (int)ceil_div(p->n_rows, rpb);
  GCG_CHECK_SHAPE((int64_t)a.n_panels * (a.seg_blocks + a.row_blocks) < INT32_MAX, "gcg_spmm_csr_f32: grid too large");

  cudaError_t e;
  switch (G * 8 + VPL) {
    case 4 * 8 + 1: e = launch_vec<4, 1>(a, st); break;
    case 8 * 8 + 1: e = launch_vec<8, 1>(a, st); break;
    case 16 * 8 + 1: e = launch_vec<16, 1>(a, st); break;
    case 4 * 8 + 5: e = launch_vec<4, 5>(a, st); break;
    case 4 * 8 + 3: e = launch_vec<4, 3>(a, st); break;
    case 8 * 8 + 3: e = launch_vec<8, 3>(a, st); break;
    case 8 * 8 + 2: e = launch_vec<8, 2>(a, st); break;
    case 16 * 8 + 2: e = launch_vec<16, 2>(a, st); break;
    case 32 * 8 + 1: e = launch_vec<32, 1>(a, st); break;
    case 32 * 8 + 2: e = launch_vec<32, 2>(a, st); break;
    case 32 * 8 + 3: e = launch_vec<32, 3>(a, st); break;
    case 32 * 8 + 4: e = launch_vec<32, 4>(a, st); break;
    default: e = launch_vec<32, 5>(a, st); break;
  }
  if (e != cudaSuccess) {
    set_error("gcg_spmm_csr_f32: launch failed: %s", cudaGetErrorString(e));
    return GCG_ERR_CUDA;
  }
  count_launch();
  if (p->n_long > 0) {
    spmm_finalize_kernel<<<(unsigned)ceil_div(p->n_long, 8), 256, 0, st>>>(a);
    GCG_LAUNCH_CHECK();
  }
  return GCG_OK;
}
