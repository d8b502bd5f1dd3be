// Shared pieces of the SpMM kernels (gcg_spmm.cu: register-gather / bulk-copy variants,
// gcg_spmm_stream.cu: nnz-balanced streaming variants).
#pragma once
#include <map>
#include <mutex>
#include <vector>

#include "gcg_common.cuh"

namespace gcg { struct StreamState; }

struct gcg_plan {
  int64_t n_rows, n_cols, nnz;
  const int32_t* indptr;
  const int32_t* indices;
  const float* vals;
  int32_t long_thresh;
  int64_t n_long, n_seg, max_deg;
  int32_t* d_seg_row;      // [n_seg]
  int32_t* d_seg_beg;      // [n_seg]
  int32_t* d_long_rows;    // [n_long]
  int32_t* d_long_segptr;  // [n_long+1]
  void* d_block;           // one pooled allocation holding the four arrays above
  size_t block_cap;
  int device;
  std::vector<int32_t> h_indptr;   // host copy of indptr: the streaming variants build their span schedule
                                   // from it on first use
  gcg::StreamState* stream;        // owned; gcg_spmm_stream.cu
  mutable cudaEvent_t last_use;    // recorded after every launch that reads d_block; the pooled block is handed to
                                   // another plan only once this event has completed (no sync in gcg_plan_destroy)
};

namespace gcg {

struct SpmmArgs {
  const int* indptr;
  const int* indices;
  const float* vals;
  const float* B;
  int64_t ldb;
  float* C;
  int64_t ldc;
  int64_t F;
  int n_rows;
  int long_thresh;
  const int* seg_row;
  const int* seg_beg;
  int n_seg;
  const int* long_rows;
  const int* long_segptr;
  int n_long;
  float* part;
  const float* bias;
  int act;
  int accumulate;
  const float* gate;
  int64_t ld_gate;
  const float* carry;
  int64_t ld_carry;
  float* conv_out;
  int64_t ld_conv;
  int f4_total;
  int panel_f4;
  int n_panels;
  int seg_blocks;
  int row_blocks;
  int nnz_total;
  // > 0: output row r belongs to owner q with owner_off[q] <= r < owner_off[q+1] and is written to
  // owner_base[q] + (r - owner_off[q]) * ldc -- peer memory of the owning GPU (gcg_spmm_csr_routed_f32)
  int n_owner;
  int owner_off[17];
  float* owner_base[16];
  int only_segments;   // vec kernel launched for the long-row segments only (rows go to the bulk-copy kernel)
};

__device__ __forceinline__ float4 f4_axpy_exact(float4 acc, float v, float4 x) {
  acc.x = __fadd_rn(acc.x, __fmul_rn(v, x.x));
  acc.y = __fadd_rn(acc.y, __fmul_rn(v, x.y));
  acc.z = __fadd_rn(acc.z, __fmul_rn(v, x.z));
  acc.w = __fadd_rn(acc.w, __fmul_rn(v, x.w));
  return acc;
}

__device__ __forceinline__ float gate_mix(float g, float hc, float h) {
  // g*Hc + (1-g)*H with separately rounded operations (matches the oracle)
  return __fadd_rn(__fmul_rn(g, hc), __fmul_rn(__fsub_rn(1.f, g), h));
}

__device__ __forceinline__ float* out_row_ptr(const SpmmArgs& a, int64_t row) {
  if (a.n_owner > 0) {
    int q = 0;
    while (q + 1 < a.n_owner && row >= a.owner_off[q + 1]) ++q;
    return a.owner_base[q] + (row - a.owner_off[q]) * a.ldc;
  }
  return a.C + row * a.ldc;
}

// bias + act + optional highway mix for one float4 of output row `row`
// at float4 column `c4`; writes C (and conv_out).
__device__ __forceinline__ void epilogue_store(const SpmmArgs& a, int64_t row, int c4, float4 v,
                                               uint64_t strm) {
  float* cp = out_row_ptr(a, row) + 4 * (int64_t)c4;
  if (a.accumulate) {
    const float4 o = *reinterpret_cast<const float4*>(cp);
    v.x = __fadd_rn(o.x, v.x); v.y = __fadd_rn(o.y, v.y);
    v.z = __fadd_rn(o.z, v.z); v.w = __fadd_rn(o.w, v.w);
  }
  if (a.bias) {
    const int64_t c = 4 * (int64_t)c4;
    v.x = __fadd_rn(v.x, __ldg(a.bias + c));
    v.y = __fadd_rn(v.y, (c + 1 < a.F) ? __ldg(a.bias + c + 1) : 0.f);
    v.z = __fadd_rn(v.z, (c + 2 < a.F) ? __ldg(a.bias + c + 2) : 0.f);
    v.w = __fadd_rn(v.w, (c + 3 < a.F) ? __ldg(a.bias + c + 3) : 0.f);
  }
  if (a.act != GCG_ACT_IDENTITY) {
    v.x = apply_act(v.x, a.act); v.y = apply_act(v.y, a.act);
    v.z = apply_act(v.z, a.act); v.w = apply_act(v.w, a.act);
  }
  if (a.gate) {
    if (a.conv_out)
      stg_f4_stream(reinterpret_cast<float4*>(a.conv_out + row * a.ld_conv + 4 * (int64_t)c4), v, strm);
    const float4 g = ldg_f4_stream(reinterpret_cast<const float4*>(a.gate + row * a.ld_gate + 4 * (int64_t)c4), strm);
    const float4 h = ldg_f4_stream(reinterpret_cast<const float4*>(a.carry + row * a.ld_carry + 4 * (int64_t)c4), strm);
    v.x = gate_mix(g.x, v.x, h.x); v.y = gate_mix(g.y, v.y, h.y);
    v.z = gate_mix(g.z, v.z, h.z); v.w = gate_mix(g.w, v.w, h.w);
  }
  stg_f4_stream(reinterpret_cast<float4*>(cp), v, strm);
}



// gcg_spmm_stream.cu
void stream_state_destroy(StreamState* s);
// launches the streaming SpMM for `a` (vector path already validated); returns 0 when this variant cannot run
// the shape (caller falls back to the register-gather kernel), 1 on success, <0 = GCG error code
int spmm_stream_launch(const gcg_plan* p, SpmmArgs& a, cudaStream_t st);
// one warp per long row: sums the segment partials in order and applies the epilogue (gcg_spmm.cu)
cudaError_t spmm_finalize_launch(const SpmmArgs& a, int n_long, cudaStream_t st);

}  // namespace gcg
