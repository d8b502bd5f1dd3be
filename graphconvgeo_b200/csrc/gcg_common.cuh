// Shared helpers for libgcg.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <functional>

#include "gcg.h"

namespace gcg {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define GCG_CHECK_ARG(cond, ...)             \
  do {                                       \
    if (!(cond)) {                           \
      gcg::set_error(__VA_ARGS__);           \
      return GCG_ERR_BAD_ARG;                \
    }                                        \
  } while (0)

#define GCG_CHECK_SHAPE(cond, ...)           \
  do {                                       \
    if (!(cond)) {                           \
      gcg::set_error(__VA_ARGS__);           \
      return GCG_ERR_SHAPE;                  \
    }                                        \
  } while (0)

#define GCG_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      gcg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,               \
                     cudaGetErrorString(e__));                                  \
      return GCG_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

#define GCG_LAUNCH_CHECK()                                                      \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      gcg::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__,           \
                     cudaGetErrorString(e__));                                  \
      return GCG_ERR_CUDA;                                                      \
    }                                                                           \
    gcg::count_launch();                                                        \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- epoch programs (gcg_epoch_*, gcg_core.cu) --------------------------------
// While a thread records, every stream-taking compute entry point appends a closure of itself (its arguments by
// value, host-side tables deep-copied, the stream left open) to the program, and still runs.  gcg_epoch_run()
// calls the closures in order on the stream it is given: the host-side C++ epoch of SURVEY section 8 row a13.
bool epoch_recording();
void epoch_record(const char* name, std::function<int(void*)> call);
#define GCG_RECORD(NAME, CALL)                                                              \
  do {                                                                                      \
    if (gcg::epoch_recording()) gcg::epoch_record(NAME, [=](void* s__) -> int { return CALL; }); \
  } while (0)

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- device helpers ---------------------------------------------------------
#ifdef __CUDACC__

// L2 eviction policies: the gathered dense operand is the only data with reuse
// (evict_last); CSR arrays and the output stream through once (evict_first).
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ float4 ldg_f4_keep(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ldg_f1_keep(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int ldg_i32_stream(const int* p, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;"
               : "=r"(v)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ldg_f32_stream(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;"
               : "=f"(v)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float4 ldg_f4_stream(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_f4_stream(float4* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
               :
               : "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void stg_f1_stream(float* p, float v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;"
               :
               : "l"(p), "f"(v), "l"(pol)
               : "memory");
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case GCG_ACT_RELU: return fmaxf(x, 0.f);
    case GCG_ACT_TANH: return tanhf(x);
    case GCG_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
// act'(p) written in terms of the activation OUTPUT a = act(p)
__device__ __forceinline__ float act_grad_from_out(float a, int act) {
  switch (act) {
    case GCG_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case GCG_ACT_TANH: return 1.f - a * a;
    case GCG_ACT_SIGMOID: return a * (1.f - a);
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__

}  // namespace gcg
