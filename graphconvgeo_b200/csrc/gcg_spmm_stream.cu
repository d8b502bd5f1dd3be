// Streaming CSR x dense SpMM (sm_100a): nnz-balanced spans, deep gather pipelines.
//
// Same contract as the register-gather kernel of gcg_spmm.cu (S.dot of lasagne_layers.py:26,65,67,84:
// products summed in CSR order with separately rounded multiply and add -> bit-identical to scipy's
// csr_matvecs), different execution shape:
//
//   * VIRTUAL ROWS.  The CSR rows, with every long row (> long_thresh non-zeros) cut into long_thresh-sized
//     segments, form one ascending list of non-zero ranges vptr[0..n_v].  vdst[v] >= 0 names the output row,
//     vdst[v] < 0 is ~segment: the partial goes to the workspace and spmm_finalize_kernel reduces the
//     segments of a row in order (deterministic, no atomics) -- exactly the scheme of the other variants.
//   * SPANS.  A span = a run of consecutive virtual rows holding about `span_nnz` non-zeros, restricted to a
//     panel of float4 columns.  One warp executes one span: its non-zeros are ONE contiguous slice of the CSR
//     arrays, streamed without draining the pipeline at row boundaries.  Every warp of a CTA carries the same
//     amount of gather work, so no warp slot idles behind a hub row (the register-gather kernel's static 8
//     rows per CTA lost 40 % of its warp time that way on the power-law mention graph).
//   * SCHEDULE.  Spans are emitted block by block (gcg_plan_set_schedule): a block is a range of rows with a
//     panel count.  Panel-major inside a block keeps `distinct columns of the block x panel bytes` resident
//     in L2 while the block's rows gather from it (2-D tiling for communities larger than L2); interleaved
//     order puts the panels of the same rows into the same CTA (thin per-warp state, whole-row DRAM locality).
//   * TRANSPORT.  Gathered rows travel either through a per-warp shared-memory ring filled by cp.async
//     (LDGSTS, 16 B per lane, no register cost: D-1 rows in flight per warp, ~150-200 KB per SM), or through
//     a D-deep rotating register buffer (LDG.128).  Both are software pipelines over the span's non-zero
//     stream: iteration t issues the gather of non-zero t and consumes non-zero t-(D-1).
#include <algorithm>
#include <limits.h>

#include "gcg_spmm.cuh"

namespace gcg {

struct StreamSpan { int32_t v_beg, v_end, f4_beg, f4_cnt; };

struct StreamSchedule {
  int f4_total = 0, vplmax = 0, n_spans = 0;
  StreamSpan* d_spans = nullptr;
};

struct StreamState {
  std::mutex mu;
  int64_t n_v = 0;
  int32_t* d_vptr = nullptr;            // [n_v + 1]
  int32_t* d_vdst = nullptr;            // [n_v]
  int32_t* d_cursor = nullptr;          // next span to hand out (persistent grid with dynamic span distribution)
  std::vector<int32_t> h_vptr;          // host copies
  std::vector<int32_t> h_row_v;         // [n_rows + 1] first virtual row of every row
  std::vector<int32_t> blk_rows;        // [n_blocks + 1] row ranges of the schedule (empty = one block)
  std::vector<int32_t> blk_panels;      // [n_blocks] >0: panel-major inside the block, <0: interleaved, |x| panels
  std::map<int64_t, StreamSchedule> scheds;   // key: f4_total * 2^20 + span_nnz
  int near_window = 0;                  // rows; 0 = every gather is kept (gcg_plan_set_near_window)
};

static int g_stream_variant = 0;     // 0 = auto
static int g_stream_span_nnz = 0;    // 0 = default
// Non-zeros per span (= per warp) and how spans reach the warps.  All 148 x 16 warps advance together, so the rows
// whose neighbourhoods compete for L2 at any moment are the span length times 2,368: SHORT spans keep that window
// narrow.  An LRU model of the benched graph (scripts/l2_model.py, profiles/r02_l2_model.json) reproduces the
// measured traffic at 256 non-zeros per span (47.7 GB modelled, 49.6 GB measured) and at 384 (51.9 GB; 8.30 vs
// 7.65 ms measured = the same ratio) and predicts 42 / 38.7 GB at 128 / 64.  Measured (profiles/r02_spmm_spans.md):
//   * one CTA per 16 spans (round-2 default until the last day): shorter spans LOSE (8.39 ms at 384, 9.42 at 64) --
//     200 KB of shared memory means one resident CTA, so the SM drains and refills at every CTA boundary;
//   * persistent grid, static stride over the spans: 9.6 ms -- the warps drift apart and the window widens;
//   * persistent grid, spans handed out IN ORDER by an atomic cursor, next span's descriptor fetched while the
//     current one drains: 8.30 ms at 128 against 8.85 ms for the old default on the same box, X^T.dZ1's block-major
//     matrix 18.8 -> 15.2 ms, epoch 209.2 -> 203.3 ms.  Below 128 the per-span start-up wins again, and at
//     ~8.4 TB/s of gathered bytes the kernel is now near the L2 -> SM rate rather than the DRAM rate.
constexpr int kDefaultSpanNnz = 128;
constexpr int kStreamPersistent = 2;   // 0: one CTA per 16 spans; 1: persistent, static stride; 2: persistent, cursor
static int g_stream_near = -1;       // -1 = plan's own choice, 0 = every gather evict_last, > 0 = window in rows

void stream_state_destroy(StreamState* s) {
  if (!s) return;
  if (s->d_vptr) cudaFree(s->d_vptr);
  if (s->d_vdst) cudaFree(s->d_vdst);
  if (s->d_cursor) cudaFree(s->d_cursor);
  for (auto& kv : s->scheds)
    if (kv.second.d_spans) cudaFree(kv.second.d_spans);
  delete s;
}

// ------------------------------------------------------------------------------------------ device side
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// Epilogue of one float4 of a finished row, LEAN flavour: identity / +bias / relu, no gate, no accumulation --
// every SpMM of an epoch except the gated layer's.  Code size matters here: the first version inlined the
// general epilogue (tanhf / expf / gate mix) VPLMAX times per flush site and unrolled the loop D times; the
// loop body outgrew the instruction cache and ncu showed the ring kernels stalled on instruction fetch
// (stall_no_instruction 4-7 warps per issue) instead of on memory.  The kernel is therefore specialised on the
// epilogue class at compile time (LEAN) and the ring loop body exists exactly once.
__device__ __forceinline__ void epilogue_store_lean(const SpmmArgs& a, int64_t row, int c4, float4 v, uint64_t strm) {
  if (a.bias) {
    const int64_t c = 4 * (int64_t)c4;
    v.x = __fadd_rn(v.x, __ldg(a.bias + c));
    v.y = __fadd_rn(v.y, (c + 1 < a.F) ? __ldg(a.bias + c + 1) : 0.f);
    v.z = __fadd_rn(v.z, (c + 2 < a.F) ? __ldg(a.bias + c + 2) : 0.f);
    v.w = __fadd_rn(v.w, (c + 3 < a.F) ? __ldg(a.bias + c + 3) : 0.f);
  }
  if (a.act == GCG_ACT_RELU) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  stg_f4_stream(reinterpret_cast<float4*>(out_row_ptr(a, row) + 4 * (int64_t)c4), v, strm);
}

// C += v, nothing else: the column-blocked products (X^T.dZ1 one document block at a time) accumulate into a
// compact buffer that is re-read soon -> default cache policy, no streaming hint
__device__ __forceinline__ void epilogue_accumulate_only(const SpmmArgs& a, int64_t row, int c4, float4 v) {
  float4* cp = reinterpret_cast<float4*>(out_row_ptr(a, row) + 4 * (int64_t)c4);
  const float4 o = *cp;
  v.x = __fadd_rn(o.x, v.x); v.y = __fadd_rn(o.y, v.y);
  v.z = __fadd_rn(o.z, v.z); v.w = __fadd_rn(o.w, v.w);
  *cp = v;
}

// +bias, relu / identity, then the highway mix g*Hc + (1-g)*H (Hc optionally stored): the gated layer's epilogue
// without the transcendental activations of the general one (a fifth of its code)
__device__ __forceinline__ void epilogue_store_gate(const SpmmArgs& a, int64_t row, int c4, float4 v, uint64_t strm) {
  const int64_t c = 4 * (int64_t)c4;
  if (a.bias) {
    v.x = __fadd_rn(v.x, __ldg(a.bias + c));
    v.y = __fadd_rn(v.y, (c + 1 < a.F) ? __ldg(a.bias + c + 1) : 0.f);
    v.z = __fadd_rn(v.z, (c + 2 < a.F) ? __ldg(a.bias + c + 2) : 0.f);
    v.w = __fadd_rn(v.w, (c + 3 < a.F) ? __ldg(a.bias + c + 3) : 0.f);
  }
  if (a.act == GCG_ACT_RELU) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  if (a.conv_out) stg_f4_stream(reinterpret_cast<float4*>(a.conv_out + row * a.ld_conv + c), v, strm);
  const float4 g = ldg_f4_stream(reinterpret_cast<const float4*>(a.gate + row * a.ld_gate + c), strm);
  const float4 h = ldg_f4_stream(reinterpret_cast<const float4*>(a.carry + row * a.ld_carry + c), strm);
  v.x = gate_mix(g.x, v.x, h.x); v.y = gate_mix(g.y, v.y, h.y);
  v.z = gate_mix(g.z, v.z, h.z); v.w = gate_mix(g.w, v.w, h.w);
  stg_f4_stream(reinterpret_cast<float4*>(a.C + row * a.ldc + c), v, strm);
}

// One warp = one span.  VPLMAX float4 per lane and row, D = pipeline depth (ring slots / register rows),
// SMEM = transport.  Iteration t issues the gather of non-zero t and consumes non-zero t - (D - 1); the loop
// runs D - 1 iterations past the span's last non-zero so that the last rows drain through the same (single)
// row-flush site.
template <int VPLMAX, int D, bool SMEM, int LEAN>
__device__ __forceinline__ void spmm_stream_body(const SpmmArgs& a, const StreamSpan* __restrict__ spans, int n_spans,
                                                 const int* __restrict__ vptr, const int* __restrict__ vdst, int n_v,
                                                 int warps_per_cta, int near_window, int* __restrict__ cursor) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t keep = policy_evict_last(), strm = policy_evict_first();
  const unsigned full = 0xffffffffu;
  // Grid-stride over the spans: with a PERSISTENT grid (one CTA per SM, spmm_stream_launch) warp w runs spans
  // w, w + G, w + 2G, ... (G = warps of the grid), so the spans in flight are still G consecutive ones -- the L2
  // window of the schedule -- but a warp that finishes a span starts its next one while the other 15 warps of the
  // SM keep their gathers in flight.  With one CTA per 16 spans instead, the whole SM drains and refills at every
  // CTA boundary (200 KB of shared memory: one resident CTA), which is what made short spans slow.
  // `cursor` != nullptr: spans are handed out in order by an atomic counter instead (dynamic: a warp that finishes
  // takes the NEXT span of the matrix, so the spans in flight stay consecutive however unevenly the warps advance;
  // with the static stride the warps drift apart and the window widens -- measured 9.6 vs 8.4 ms).
  const int span_stride = (int)gridDim.x * warps_per_cta;
  // the next span of this warp: its index (cursor or stride), descriptor and non-zero range.  Fetched while the
  // CURRENT span drains (its last D - 1 gathers are in flight anyway), so that a new span starts with one level of
  // dependent loads (its first index / value / row-boundary chunks) instead of four.
  int si = blockIdx.x * warps_per_cta + warp;
  int4 spv_n = make_int4(0, 0, 0, 0);
  int k0_n = 0, k1_n = 0;
  auto fetch_next = [&](bool first) {
    if (cursor) {
      int nx = 0;
      if (lane == 0) nx = atomicAdd(cursor, 1);
      si = __shfl_sync(full, nx, 0);
    } else if (!first) {
      si += span_stride;
    }
    if (si < n_spans) {
      spv_n = __ldg(reinterpret_cast<const int4*>(spans) + si);
      k0_n = __ldg(vptr + spv_n.x);
      k1_n = __ldg(vptr + spv_n.y);
    }
  };
  fetch_next(true);
  while (si < n_spans) {
  const int4 spv = spv_n;
  const int v_beg = spv.x, v_end = spv.y, f4_beg = spv.z, f4_cnt = spv.w;
  bool fetched = false;

  bool cv[VPLMAX];
#pragma unroll
  for (int j = 0; j < VPLMAX; ++j) cv[j] = (lane + 32 * j) < f4_cnt;
  const float4* __restrict__ Bp = reinterpret_cast<const float4*>(a.B) + f4_beg + lane;
  const int64_t ldb4 = a.ldb >> 2;
  const int k0 = k0_n, k1 = k1_n;
  const int nnz_total = a.nnz_total;

  // chunk caches: 32 consecutive column indices (producer side), values (consumer side) and virtual-row
  // boundaries / destinations, each with its successor prefetched
  auto ld_idx = [&](int base) { const int q = base + lane; return q < nnz_total ? ldg_i32_stream(a.indices + q, strm) : 0; };
  auto ld_val = [&](int base) { const int q = base + lane; return q < nnz_total ? ldg_f32_stream(a.vals + q, strm) : 0.f; };
  auto ld_vend = [&](int base) { const int q = base + lane + 1; return q <= n_v ? __ldg(vptr + q) : INT_MAX; };
  auto ld_vdst = [&](int base) { const int q = base + lane; return q < n_v ? __ldg(vdst + q) : 0; };
  int pb = k0 & ~31;
  int pidx = ld_idx(pb), pidx_nx = ld_idx(pb + 32);
  int cb = pb;
  float cval = ld_val(cb), cval_nx = ld_val(cb + 32);
  int vb = v_beg;
  int vend_c = ld_vend(vb), vend_nx = ld_vend(vb + 32);
  int vdst_c = ld_vdst(vb), vdst_nx = ld_vdst(vb + 32);
  int v = v_beg;
  int vstart = k0;
  int vend = __shfl_sync(full, vend_c, 0);
  // L2 policy per gathered row: columns within `near_window` rows of the span (the community the node order keeps
  // together) are the reusable set -> evict_last; far columns (the random long-range mentions, never re-read
  // before eviction) stream with evict_first so that they do not push the reusable set out of L2.
  const int r0 = __shfl_sync(full, vdst_c, 0);
  const bool classify = near_window > 0 && r0 >= 0;

  float4 acc[VPLMAX];
#pragma unroll
  for (int j = 0; j < VPLMAX; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto flush_row = [&]() {      // virtual row v is complete: store it, move on to v + 1
    const int dst = __shfl_sync(full, vdst_c, v - vb);
    if (dst >= 0) {
      // plain accumulation of an empty row is a no-op (column-blocked products visit mostly empty rows)
      const bool skip = (vstart == vend) && a.accumulate && !a.bias && a.act == GCG_ACT_IDENTITY && !a.gate;
      if (!skip) {
#pragma unroll
        for (int j = 0; j < VPLMAX; ++j)
          if (cv[j]) {
            if (LEAN == 1) epilogue_store_lean(a, dst, f4_beg + lane + 32 * j, acc[j], strm);
            else if (LEAN == 2) epilogue_accumulate_only(a, dst, f4_beg + lane + 32 * j, acc[j]);
            else if (LEAN == 3) epilogue_store_gate(a, dst, f4_beg + lane + 32 * j, acc[j], strm);
            else epilogue_store(a, dst, f4_beg + lane + 32 * j, acc[j], strm);
          }
      }
    } else {
      float4* out = reinterpret_cast<float4*>(a.part + (int64_t)(~dst) * a.ldc) + f4_beg + lane;
#pragma unroll
      for (int j = 0; j < VPLMAX; ++j)
        if (cv[j]) out[32 * j] = acc[j];             // re-read soon by finalize: default policy
    }
#pragma unroll
    for (int j = 0; j < VPLMAX; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    ++v;
    if (v - vb == 32) {
      vb += 32;
      vend_c = vend_nx; vdst_c = vdst_nx;
      vend_nx = ld_vend(vb + 32); vdst_nx = ld_vdst(vb + 32);
    }
    vstart = vend;
    vend = __shfl_sync(full, vend_c, v - vb);
  };

  if (SMEM) {
    // ---- shared-memory ring: slot index is a run-time counter, the loop body exists once
    constexpr uint32_t SLOT = VPLMAX * 512;
    const uint32_t ring = smem_u32(smem_dyn) + (uint32_t)warp * (D * SLOT) + lane * 16;
    uint32_t ps = 0, cs = (D > 1) ? SLOT : 0;          // producer slot offset; consumer runs one slot ahead of it
    for (int t = k0; t < k1 + D; ++t) {
      if (t == k1) { fetch_next(false); fetched = true; }     // first drain iteration: see fetch_next
      if (t < k1) {
        if (t >= pb + 32) { pb += 32; pidx = pidx_nx; pidx_nx = ld_idx(pb + 32); }
        const int col = __shfl_sync(full, pidx, t - pb);
        const float4* src = Bp + (int64_t)col * ldb4;
        const uint64_t pol = (classify && abs(col - r0) > near_window) ? strm : keep;
#pragma unroll
        for (int j = 0; j < VPLMAX; ++j)
          if (cv[j]) cp_async16(ring + ps + j * 512, src + 32 * j, pol);
      }
      cp_async_commit();                                // one group per iteration (possibly empty)
      const int c = t - (D - 1);
      if (c >= k0) {
        while (v < v_end && c >= vend) flush_row();     // row boundaries, empty rows, and the tail at c == k1
        if (c < k1) {
          if (c >= cb + 32) { cb += 32; cval = cval_nx; cval_nx = ld_val(cb + 32); }
          const float val = __shfl_sync(full, cval, c - cb);
          cp_async_wait<D - 1>();
          float4 xv[VPLMAX];
#pragma unroll
          for (int j = 0; j < VPLMAX; ++j)
            if (cv[j]) xv[j] = lds_f4(ring + cs + j * 512);
#pragma unroll
          for (int j = 0; j < VPLMAX; ++j)
            if (cv[j]) acc[j] = f4_axpy_exact(acc[j], val, xv[j]);
        }
      }
      ps = (ps + SLOT == D * SLOT) ? 0u : ps + SLOT;
      cs = (cs + SLOT == D * SLOT) ? 0u : cs + SLOT;
    }
  } else {
    // ---- register pipeline: D rows rotate through statically indexed registers (unrolled by D)
    float4 x[D][VPLMAX];
    for (int base = k0; base < k1 + D; base += D) {
#pragma unroll
      for (int u = 0; u < D; ++u) {
        const int t = base + u;
        if (t < k1) {
          if (t >= pb + 32) { pb += 32; pidx = pidx_nx; pidx_nx = ld_idx(pb + 32); }
          const int col = __shfl_sync(full, pidx, t - pb);
          const float4* src = Bp + (int64_t)col * ldb4;
          const uint64_t pol = (classify && abs(col - r0) > near_window) ? strm : keep;
#pragma unroll
          for (int j = 0; j < VPLMAX; ++j)
            if (cv[j]) x[u][j] = ldg_f4_keep(src + 32 * j, pol);
        }
        const int c = t - (D - 1);
        if (c >= k0) {
          while (v < v_end && c >= vend) flush_row();
          if (c < k1) {
            if (c >= cb + 32) { cb += 32; cval = cval_nx; cval_nx = ld_val(cb + 32); }
            const float val = __shfl_sync(full, cval, c - cb);
            const int cu = (u + 1 == D) ? 0 : u + 1;    // compile-time after unrolling
#pragma unroll
            for (int j = 0; j < VPLMAX; ++j)
              if (cv[j]) acc[j] = f4_axpy_exact(acc[j], val, x[cu][j]);
          }
        }
      }
    }
  }
  while (v < v_end) flush_row();                         // spans made of empty rows only
  if (!fetched) fetch_next(false);                       // register-pipeline variants
  }   // next span of this warp
}

template <int VPLMAX, int D, int WARPS, int MINB, bool SMEM, int LEAN>
__global__ void __launch_bounds__(WARPS * 32, MINB)
spmm_stream_kernel(const SpmmArgs a, const StreamSpan* __restrict__ spans, int n_spans, const int* __restrict__ vptr,
                   const int* __restrict__ vdst, int n_v, int near_window, int* __restrict__ cursor) {
  spmm_stream_body<VPLMAX, D, SMEM, LEAN>(a, spans, n_spans, vptr, vdst, n_v, WARPS, near_window, cursor);
}

// ------------------------------------------------------------------------------------------ host side
template <int VPLMAX, int D, int WARPS, int MINB, bool SMEM, int LEAN>
static cudaError_t launch_stream_epi(const SpmmArgs& a, const StreamSchedule& s, const StreamState& ss, cudaStream_t st) {
  auto kern = spmm_stream_kernel<VPLMAX, D, WARPS, MINB, SMEM, LEAN>;
  const int smem_bytes = SMEM ? WARPS * D * VPLMAX * 512 : 0;
  if (SMEM) {
    static bool attr_set = false;       // per instantiation
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
  }
  unsigned grid = (unsigned)ceil_div(s.n_spans, WARPS);
  // 0: one CTA per 16 spans; 1: persistent grid, static stride; 2: persistent grid, spans handed out by an atomic cursor
  static const int persistent = getenv("GCG_STREAM_PERSISTENT") ? atoi(getenv("GCG_STREAM_PERSISTENT")) : kStreamPersistent;
  int* cursor = nullptr;
  if (persistent && grid > (unsigned)(kNumSMs * MINB)) {
    grid = (unsigned)(kNumSMs * MINB);
    if (persistent == 2 && ss.d_cursor) {
      cursor = ss.d_cursor;
      cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(int32_t), st);
      if (e != cudaSuccess) return e;
    }
  }
  const int near = g_stream_near >= 0 ? g_stream_near : ss.near_window;
  kern<<<grid, WARPS * 32, smem_bytes, st>>>(a, s.d_spans, s.n_spans, ss.d_vptr, ss.d_vdst, (int)ss.n_v, near, cursor);
  return cudaGetLastError();
}

template <int VPLMAX, int D, int WARPS, int MINB, bool SMEM>
static cudaError_t launch_stream(const SpmmArgs& a, const StreamSchedule& s, const StreamState& ss, cudaStream_t st) {
  const bool lean = a.act <= GCG_ACT_RELU && !a.gate && !a.accumulate;
  const bool acc_only = a.accumulate && !a.bias && a.act == GCG_ACT_IDENTITY && !a.gate && a.n_owner == 0;
  if (lean) return launch_stream_epi<VPLMAX, D, WARPS, MINB, SMEM, 1>(a, s, ss, st);
  if (acc_only) return launch_stream_epi<VPLMAX, D, WARPS, MINB, SMEM, 2>(a, s, ss, st);
  const bool gate_only = a.gate && a.act <= GCG_ACT_RELU && !a.accumulate && a.n_owner == 0;
  if (gate_only) return launch_stream_epi<VPLMAX, D, WARPS, MINB, SMEM, 3>(a, s, ss, st);
  return launch_stream_epi<VPLMAX, D, WARPS, MINB, SMEM, 0>(a, s, ss, st);
}

// (VPLMAX, variant) -> kernel.  Variant 0 picks the measured default of the width class.
//   variants 1..4: shared-memory ring;   5..8: register pipeline
static cudaError_t dispatch_stream(int vplmax, int variant, const SpmmArgs& a, const StreamSchedule& s,
                                   const StreamState& ss, cudaStream_t st, bool* found) {
  *found = true;
#define GCG_ST(V, VAR, D, W, MB, SM) \
  if (vplmax == V && variant == VAR) return launch_stream<V, D, W, MB, SM>(a, s, ss, st);
  // ---- 1 float4 per lane (<= 128 floats per panel)
  GCG_ST(1, 1, 16, 16, 1, true)  GCG_ST(1, 2, 12, 32, 1, true)  GCG_ST(1, 3, 8, 16, 2, true)
  GCG_ST(1, 5, 8, 8, 3, false)   GCG_ST(1, 6, 4, 8, 4, false)
  // ---- 2 float4 per lane (<= 256 floats)
  GCG_ST(2, 1, 12, 16, 1, true)  GCG_ST(2, 2, 8, 24, 1, true)   GCG_ST(2, 3, 6, 16, 2, true)
  GCG_ST(2, 5, 4, 8, 3, false)   GCG_ST(2, 6, 2, 8, 4, false)
  // ---- 3 float4 per lane (<= 384 floats)
  GCG_ST(3, 1, 8, 16, 1, true)   GCG_ST(3, 2, 5, 24, 1, true)
  GCG_ST(3, 5, 3, 8, 3, false)   GCG_ST(3, 6, 2, 8, 3, false)
  // ---- 4 float4 per lane (<= 512 floats)
  GCG_ST(4, 1, 6, 16, 1, true)   GCG_ST(4, 2, 4, 24, 1, true)
  GCG_ST(4, 5, 3, 8, 2, false)   GCG_ST(4, 6, 2, 8, 2, false)
  // ---- 5 float4 per lane (<= 640 floats: hidden 600)
  GCG_ST(5, 1, 5, 16, 1, true)   GCG_ST(5, 2, 4, 16, 1, true)   GCG_ST(5, 3, 4, 20, 1, true)   GCG_ST(5, 4, 3, 24, 1, true)
  GCG_ST(5, 5, 3, 8, 2, false)   GCG_ST(5, 7, 2, 8, 2, false)   GCG_ST(5, 8, 2, 32, 1, true)   GCG_ST(5, 9, 3, 28, 1, true)
  // ---- 8 float4 per lane (<= 1024 floats: 1024 regions)
  GCG_ST(8, 1, 3, 16, 1, true)   GCG_ST(8, 2, 2, 24, 1, true)
  GCG_ST(8, 5, 2, 8, 2, false)
#undef GCG_ST
  *found = false;
  return cudaSuccess;
}

static int default_variant(int vplmax) { return 1; }

static int build_vrows(const gcg_plan* p, StreamState* s) {
  const int64_t n = p->n_rows;
  const int32_t T = p->long_thresh;
  const std::vector<int32_t>& ip = p->h_indptr;
  std::vector<int32_t> vdst;
  s->h_vptr.clear();
  s->h_row_v.assign(n + 1, 0);
  s->h_vptr.reserve(n + p->n_seg + 1);
  vdst.reserve(n + p->n_seg);
  int32_t seg = 0;
  for (int64_t r = 0; r < n; ++r) {
    s->h_row_v[r] = (int32_t)vdst.size();
    const int32_t b = ip[r], e = ip[r + 1];
    if (e - b > T) {
      for (int32_t k = b; k < e; k += T) {
        s->h_vptr.push_back(k);
        vdst.push_back(~seg);
        ++seg;
      }
    } else {
      s->h_vptr.push_back(b);
      vdst.push_back((int32_t)r);
    }
  }
  s->h_row_v[n] = (int32_t)vdst.size();
  s->h_vptr.push_back(ip[n]);
  s->n_v = (int64_t)vdst.size();
  if (seg != p->n_seg) { set_error("stream plan: segment count mismatch (%d vs %lld)", seg, (long long)p->n_seg); return GCG_ERR_SHAPE; }
  GCG_CUDA(cudaMalloc(&s->d_vptr, sizeof(int32_t) * (s->n_v + 1)));
  GCG_CUDA(cudaMalloc(&s->d_vdst, sizeof(int32_t) * std::max<int64_t>(1, s->n_v)));
  GCG_CUDA(cudaMalloc(&s->d_cursor, sizeof(int32_t)));
  GCG_CUDA(cudaMemcpy(s->d_vptr, s->h_vptr.data(), sizeof(int32_t) * (s->n_v + 1), cudaMemcpyHostToDevice));
  if (s->n_v > 0) GCG_CUDA(cudaMemcpy(s->d_vdst, vdst.data(), sizeof(int32_t) * s->n_v, cudaMemcpyHostToDevice));
  return GCG_OK;
}

static int build_schedule(const gcg_plan* p, StreamState* s, int f4_total, int span_nnz, StreamSchedule* out) {
  std::vector<StreamSpan> spans;
  const int64_t n = p->n_rows;
  std::vector<int32_t> one_rows = {0, (int32_t)n};
  std::vector<int32_t> one_panels = {1};
  const std::vector<int32_t>& brows = s->blk_rows.empty() ? one_rows : s->blk_rows;
  const std::vector<int32_t>& bpan = s->blk_rows.empty() ? one_panels : s->blk_panels;
  int vplmax = 1;
  const size_t nb = bpan.size();
  for (size_t b = 0; b < nb; ++b) {
    const int32_t r0 = brows[b], r1 = brows[b + 1];
    if (r1 <= r0) continue;
    const bool interleave = bpan[b] < 0;
    int np = std::max(1, std::abs(bpan[b]));
    // panel width: a multiple of 8 float4 (128 B) unless one panel covers the row
    int pw = (int)ceil_div(f4_total, np);
    if (np > 1) pw = (int)ceil_div(pw, 8) * 8;
    pw = std::min(pw, 256);                            // VPLMAX <= 8
    np = (int)ceil_div(f4_total, pw);
    const int vpl = (int)ceil_div(std::min(pw, f4_total), 32);
    vplmax = std::max(vplmax, vpl);
    const int32_t v0 = s->h_row_v[r0], v1 = s->h_row_v[r1];
    // spans of the block for one panel (the same cut for every panel)
    std::vector<std::pair<int32_t, int32_t>> cuts;
    const int64_t target = std::max<int64_t>(32, (int64_t)span_nnz * 5 / std::max(1, std::min(vpl, 5)));
    int32_t vb = v0;
    while (vb < v1) {
      int32_t ve = vb + 1;
      const int64_t kb = s->h_vptr[vb];
      while (ve < v1 && (int64_t)s->h_vptr[ve + 1] - kb <= target && ve - vb < 4096) ++ve;
      cuts.push_back({vb, ve});
      vb = ve;
    }
    if (interleave) {
      for (auto& c : cuts)
        for (int q = 0; q < np; ++q)
          spans.push_back({c.first, c.second, q * pw, std::min(pw, f4_total - q * pw)});
    } else {
      for (int q = 0; q < np; ++q)
        for (auto& c : cuts)
          spans.push_back({c.first, c.second, q * pw, std::min(pw, f4_total - q * pw)});
    }
  }
  out->f4_total = f4_total;
  out->vplmax = vplmax == 6 || vplmax == 7 ? 8 : vplmax;
  out->n_spans = (int)spans.size();
  out->d_spans = nullptr;
  if (!spans.empty()) {
    GCG_CUDA(cudaMalloc(&out->d_spans, sizeof(StreamSpan) * spans.size()));
    GCG_CUDA(cudaMemcpy(out->d_spans, spans.data(), sizeof(StreamSpan) * spans.size(), cudaMemcpyHostToDevice));
  }
  return GCG_OK;
}

int spmm_stream_launch(const gcg_plan* p, SpmmArgs& a, cudaStream_t st) {
  if (a.f4_total > 256) return 0;
  gcg_plan* mp = const_cast<gcg_plan*>(p);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  if (!mp->stream) {
    if (cap != cudaStreamCaptureStatusNone) {
      set_error("gcg_spmm_csr_f32: the streaming variant must run once outside CUDA-graph capture (it builds its schedule on first use)");
      return GCG_ERR_UNSUPPORTED;
    }
    mp->stream = new StreamState();
    const int rc = build_vrows(p, mp->stream);
    if (rc != GCG_OK) return rc;
  }
  StreamState* s = mp->stream;
  // default span: 384 non-zeros, shorter for small matrices so that the grid still covers the 148 SMs twice over
  int span_nnz = g_stream_span_nnz;
  if (span_nnz <= 0) {
    static const int env_span = getenv("GCG_STREAM_SPAN") ? atoi(getenv("GCG_STREAM_SPAN")) : 0;
    const int64_t nz = (int64_t)p->h_indptr[p->n_rows] - p->h_indptr[0];
    span_nnz = (int)std::min<int64_t>(env_span > 0 ? env_span : kDefaultSpanNnz,
                                      std::max<int64_t>(32, nz / (kNumSMs * 16 * 2)));
  }
  StreamSchedule sched;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    const int64_t key = ((int64_t)a.f4_total << 20) + span_nnz;
    auto it = s->scheds.find(key);
    if (it == s->scheds.end()) {
      if (cap != cudaStreamCaptureStatusNone) {
        set_error("gcg_spmm_csr_f32: streaming schedule for F4=%d missing during CUDA-graph capture", a.f4_total);
        return GCG_ERR_UNSUPPORTED;
      }
      StreamSchedule ns;
      const int rc = build_schedule(p, s, a.f4_total, span_nnz, &ns);
      if (rc != GCG_OK) return rc;
      it = s->scheds.emplace(key, ns).first;
    }
    sched = it->second;
  }
  if (sched.n_spans == 0) return 1;
  int variant = g_stream_variant > 0 ? g_stream_variant : default_variant(sched.vplmax);
  bool found = false;
  cudaError_t e = dispatch_stream(sched.vplmax, variant, a, sched, *s, st, &found);
  if (!found) {
    variant = default_variant(sched.vplmax);
    e = dispatch_stream(sched.vplmax, variant, a, sched, *s, st, &found);
    if (!found) return 0;
  }
  if (e != cudaSuccess) { set_error("gcg_spmm_csr_f32: streaming launch failed: %s", cudaGetErrorString(e)); return GCG_ERR_CUDA; }
  count_launch();
  if (p->n_long > 0) {
    e = spmm_finalize_launch(a, (int)p->n_long, st);
    if (e != cudaSuccess) { set_error("gcg_spmm_csr_f32: finalize launch failed: %s", cudaGetErrorString(e)); return GCG_ERR_CUDA; }
    count_launch();
  }
  return 1;
}

}  // namespace gcg

using namespace gcg;

extern "C" void gcg_spmm_stream_tuning(int variant, int span_nnz, int near_window) {
  gcg::g_stream_variant = variant;
  gcg::g_stream_span_nnz = span_nnz;
  gcg::g_stream_near = near_window;
}

extern "C" int gcg_plan_set_near_window(gcg_plan* p, int32_t near_window_rows) {
  GCG_CHECK_ARG(p != nullptr, "gcg_plan_set_near_window: plan is NULL");
  GCG_CHECK_ARG(near_window_rows >= 0, "gcg_plan_set_near_window: negative window");
  if (!p->stream) {
    p->stream = new StreamState();
    const int rc = build_vrows(p, p->stream);
    if (rc != GCG_OK) return rc;
  }
  p->stream->near_window = near_window_rows;
  return GCG_OK;
}

extern "C" int gcg_plan_set_schedule(gcg_plan* p, int64_t n_blocks, const int32_t* h_block_rows,
                                     const int32_t* h_block_panels) {
  GCG_CHECK_ARG(p != nullptr, "gcg_plan_set_schedule: plan is NULL");
  GCG_CHECK_ARG(n_blocks >= 0 && (n_blocks == 0 || (h_block_rows && h_block_panels)), "gcg_plan_set_schedule: NULL arrays");
  if (n_blocks > 0) {
    GCG_CHECK_SHAPE(h_block_rows[0] == 0 && h_block_rows[n_blocks] == p->n_rows,
                    "gcg_plan_set_schedule: blocks must cover rows [0, %lld)", (long long)p->n_rows);
    for (int64_t b = 0; b < n_blocks; ++b) {
      GCG_CHECK_SHAPE(h_block_rows[b + 1] >= h_block_rows[b], "gcg_plan_set_schedule: block %lld is reversed", (long long)b);
      GCG_CHECK_ARG(h_block_panels[b] != 0 && std::abs(h_block_panels[b]) <= 64, "gcg_plan_set_schedule: panel count %d", h_block_panels[b]);
    }
  }
  if (!p->stream) {
    p->stream = new StreamState();
    const int rc = build_vrows(p, p->stream);
    if (rc != GCG_OK) return rc;
  }
  StreamState* s = p->stream;
  std::lock_guard<std::mutex> lk(s->mu);
  // schedules built for the previous blocks are stale; kernels that may still read them must drain first
  if (!s->scheds.empty()) {
    GCG_CUDA(cudaDeviceSynchronize());
    for (auto& kv : s->scheds)
      if (kv.second.d_spans) cudaFree(kv.second.d_spans);
    s->scheds.clear();
  }
  s->blk_rows.assign(h_block_rows, h_block_rows + (n_blocks > 0 ? n_blocks + 1 : 0));
  s->blk_panels.assign(h_block_panels, h_block_panels + n_blocks);
  return GCG_OK;
}
