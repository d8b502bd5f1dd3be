// Multi-GPU entry points of the C ABI: one process per GPU, A_hat row-partitioned.
//
//   gcg_comm_init                      NCCL communicator for this rank (+ a communication stream and two events)
//   gcg_spmm_rowpart_allgather_f32     one propagation A_hat[rows of this rank, :] . Z of north_star's design: the local
//                                      slab of Z is all-gathered IN PLACE over NVLink on the communication stream while
//                                      the calling stream already runs the SpMM over the DIAGONAL column block (which
//                                      reads the local slab only); the off-diagonal block follows with accumulate + the
//                                      fused epilogue once the gather has landed
//   gcg_allreduce_grads_f32            sum of the partial parameter gradients over ranks (one NCCL group)
//   gcg_comm_wait                      fence: the calling stream waits for the collectives issued so far
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a PyTorch process this is the copy torch has already
// loaded), so libgcg.so itself has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "gcg_common.cuh"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

bool nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.ok) return true;
  if (!g_nccl.handle) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
      gcg::set_error("gcg_comm: cannot load libnccl.so.2 (%s)", dlerror());
      return false;
    }
  }
  bool all = true;
#define GCG_NCCL_SYM(field, sym)                                                    \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, sym)); \
  all = all && (g_nccl.field != nullptr);
  GCG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  GCG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  GCG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  GCG_NCCL_SYM(AllGather, "ncclAllGather")
  GCG_NCCL_SYM(AllReduce, "ncclAllReduce")
  GCG_NCCL_SYM(GroupStart, "ncclGroupStart")
  GCG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  GCG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef GCG_NCCL_SYM
  if (!all) {
    gcg::set_error("gcg_comm: libnccl.so.2 lacks a required symbol");
    return false;
  }
  g_nccl.ok = true;
  return true;
}

#define GCG_NCCL(call)                                                                             \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) {                                                                      \
      gcg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));     \
      return GCG_ERR_NCCL;                                                                         \
    }                                                                                              \
  } while (0)

}  // namespace

struct gcg_comm {
  ncclComm_t comm = nullptr;
  int world = 0, rank = 0, device = 0;
  cudaStream_t cs = nullptr;     // communication stream (collectives overlap the caller's stream)
  cudaEvent_t ready = nullptr;   // caller's stream -> communication stream: operands are written
  cudaEvent_t done = nullptr;    // communication stream -> caller's stream: the collective has landed
  bool pending = false;          // `done` has been recorded and not yet waited for
};

extern "C" int gcg_comm_unique_id(void* id128) {
  GCG_CHECK_ARG(id128 != nullptr, "gcg_comm_unique_id: id128 is NULL");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  if (!nccl_load()) return GCG_ERR_NCCL;
  GCG_NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
  return GCG_OK;
}

extern "C" int gcg_comm_init(const void* id128, int32_t world, int32_t rank, gcg_comm** out) {
  GCG_CHECK_ARG(out != nullptr && id128 != nullptr, "gcg_comm_init: NULL argument");
  GCG_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "gcg_comm_init: rank %d of %d", rank, world);
  if (!nccl_load()) return GCG_ERR_NCCL;
  gcg_comm* c = new gcg_comm();
  c->world = world;
  c->rank = rank;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->cs, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    gcg::set_error("gcg_comm_init: %s", cudaGetErrorString(e));
    delete c;
    return GCG_ERR_CUDA;
  }
  const ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);   // collective over all ranks
  if (r != ncclSuccess) {
    gcg::set_error("gcg_comm_init: ncclCommInitRank -> %s", g_nccl.GetErrorString(r));
    cudaEventDestroy(c->ready);
    cudaEventDestroy(c->done);
    cudaStreamDestroy(c->cs);
    delete c;
    return GCG_ERR_NCCL;
  }
  *out = c;
  return GCG_OK;
}

extern "C" int gcg_comm_destroy(gcg_comm* c) {
  if (!c) return GCG_OK;
  if (c->cs) cudaStreamSynchronize(c->cs);
  if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
  if (c->ready) cudaEventDestroy(c->ready);
  if (c->done) cudaEventDestroy(c->done);
  if (c->cs) cudaStreamDestroy(c->cs);
  delete c;
  return GCG_OK;
}

extern "C" int gcg_comm_info(const gcg_comm* c, int32_t* world, int32_t* rank) {
  GCG_CHECK_ARG(c && world && rank, "gcg_comm_info: NULL argument");
  *world = c->world;
  *rank = c->rank;
  return GCG_OK;
}

extern "C" int gcg_comm_wait(gcg_comm* c, void* stream) {
  GCG_CHECK_ARG(c != nullptr, "gcg_comm_wait: comm is NULL");
  if (c->pending) {
    GCG_CUDA(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(stream), c->done, 0));
    c->pending = false;
  }
  return GCG_OK;
}

extern "C" int gcg_allgather_rows_f32(gcg_comm* c, float* full, int64_t floats_per_rank, int32_t wait, void* stream) {
  GCG_CHECK_ARG(c && full, "gcg_allgather_rows_f32: NULL argument");
  GCG_CHECK_SHAPE(floats_per_rank >= 0, "gcg_allgather_rows_f32: negative size");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (floats_per_rank == 0) return GCG_OK;
  GCG_CUDA(cudaEventRecord(c->ready, st));
  GCG_CUDA(cudaStreamWaitEvent(c->cs, c->ready, 0));
  GCG_NCCL(g_nccl.AllGather(full + (int64_t)c->rank * floats_per_rank, full, (size_t)floats_per_rank, ncclFloat, c->comm,
                            c->cs));
  GCG_CUDA(cudaEventRecord(c->done, c->cs));
  c->pending = true;
  if (wait) return gcg_comm_wait(c, stream);
  return GCG_OK;
}

extern "C" int gcg_allreduce_grads_f32(gcg_comm* c, int32_t n_tensors, float* const* h_bufs, const int64_t* h_sizes,
                                       int32_t wait, void* stream) {
  GCG_CHECK_ARG(c != nullptr, "gcg_allreduce_grads_f32: comm is NULL");
  GCG_CHECK_ARG(n_tensors >= 0 && (n_tensors == 0 || (h_bufs && h_sizes)), "gcg_allreduce_grads_f32: bad tensor list");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < n_tensors; ++i)
    GCG_CHECK_ARG(h_sizes[i] >= 0 && (h_sizes[i] == 0 || h_bufs[i]), "gcg_allreduce_grads_f32: tensor %d invalid", i);
  if (n_tensors == 0) return GCG_OK;
  GCG_CUDA(cudaEventRecord(c->ready, st));
  GCG_CUDA(cudaStreamWaitEvent(c->cs, c->ready, 0));
  GCG_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n_tensors; ++i) {
    if (h_sizes[i] == 0) continue;
    const ncclResult_t r = g_nccl.AllReduce(h_bufs[i], h_bufs[i], (size_t)h_sizes[i], ncclFloat, ncclSum, c->comm, c->cs);
    if (r != ncclSuccess) {
      g_nccl.GroupEnd();
      gcg::set_error("gcg_allreduce_grads_f32: ncclAllReduce(tensor %d) -> %s", i, g_nccl.GetErrorString(r));
      return GCG_ERR_NCCL;
    }
  }
  GCG_NCCL(g_nccl.GroupEnd());
  GCG_CUDA(cudaEventRecord(c->done, c->cs));
  c->pending = true;
  if (wait) return gcg_comm_wait(c, stream);
  return GCG_OK;
}

extern "C" int gcg_spmm_rowpart_allgather_f32(gcg_comm* c, const gcg_plan* diag, const gcg_plan* off, float* Z_full,
                                              int64_t ld, int64_t F, int64_t n_loc, float* C, int64_t ldc,
                                              const float* bias, int act, const float* gate, int64_t ld_gate,
                                              const float* carry, int64_t ld_carry, float* conv_out, int64_t ld_conv,
                                              int32_t panel_cols, void* workspace, int64_t workspace_bytes,
                                              void* stream) {
  GCG_CHECK_ARG(c && diag && off && Z_full && C, "gcg_spmm_rowpart_allgather_f32: NULL argument");
  GCG_CHECK_SHAPE(n_loc > 0 && F > 0 && ld >= F, "gcg_spmm_rowpart_allgather_f32: n_loc=%lld F=%lld ld=%lld",
                  (long long)n_loc, (long long)F, (long long)ld);
  int64_t di[8], oi[8];
  int rc = gcg_plan_info(diag, di);
  if (rc != GCG_OK) return rc;
  rc = gcg_plan_info(off, oi);
  if (rc != GCG_OK) return rc;
  GCG_CHECK_SHAPE(di[0] == oi[0] && di[1] == oi[1] && di[1] == n_loc * c->world,
                  "gcg_spmm_rowpart_allgather_f32: blocks are [%lld x %lld] / [%lld x %lld], operand has %lld rows",
                  (long long)di[0], (long long)di[1], (long long)oi[0], (long long)oi[1], (long long)(n_loc * c->world));
  // (1) the gather starts as soon as the local slab is complete on the caller's stream
  rc = gcg_allgather_rows_f32(c, Z_full, n_loc * ld, 0, stream);
  if (rc != GCG_OK) return rc;
  // (2) diagonal block: reads rows [rank*n_loc, (rank+1)*n_loc) of Z_full only -- overlaps the gather
  if (di[0] > 0) {
    rc = gcg_spmm_csr_f32(diag, Z_full, ld, F, C, ldc, nullptr, GCG_ACT_IDENTITY, 0, nullptr, 0, nullptr, 0, nullptr, 0,
                          panel_cols, workspace, workspace_bytes, stream);
    if (rc != GCG_OK) return rc;
  }
  // (3) the rest of the columns once the other ranks' slabs have landed, with the layer's epilogue
  rc = gcg_comm_wait(c, stream);
  if (rc != GCG_OK) return rc;
  if (di[0] == 0) return GCG_OK;
  return gcg_spmm_csr_f32(off, Z_full, ld, F, C, ldc, bias, act, 1, gate, ld_gate, carry, ld_carry, conv_out, ld_conv,
                          panel_cols, workspace, workspace_bytes, stream);
}
