/*
 * gcg.h -- C ABI of libgcg.so: the B200 (sm_100a) implementation of the
 * graphconvgeo GCN propagation hot path.
 *
 * The reference (afcarl/graphconvgeo) has no FFI: its boundary for this path is
 * the Lasagne Layer protocol (lasagne_layers.py:20-89) whose bodies call
 * Theano's S.dot / T.dot.  Each entry point below replaces one of those calls
 * (or the Theano-generated ops around it); the reference call site is cited
 * per function.  graphconvgeo_b200/lasagne_layers.py re-creates the Layer
 * classes on top of these functions via ctypes.
 *
 * Conventions
 *   - every function returns 0 (GCG_OK) or a negative gcg_status; the message
 *     is available from gcg_last_error() (thread-local).
 *   - all matrices are float32 row-major with an explicit leading dimension
 *     (in elements); CSR = (indptr int32[n_rows+1], indices int32[nnz],
 *     vals float32[nnz]) with column indices sorted within a row.
 *   - pointers named d_* / unprefixed data pointers are DEVICE pointers owned
 *     by the caller (e.g. torch tensors); h_* are host pointers.  The library
 *     never frees or retains caller memory beyond the call, except a
 *     gcg_plan, which keeps the CSR device pointers it was created with (the
 *     caller keeps them alive until gcg_plan_destroy).
 *   - every device function is asynchronous on `stream` (a cudaStream_t passed
 *     as void*), performs no allocation and no synchronisation, and is
 *     therefore CUDA-graph capturable.  *_host functions run on the CPU.
 */
#ifndef GCG_H_
#define GCG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  GCG_OK = 0,
  GCG_ERR_BAD_ARG = -1,
  GCG_ERR_SHAPE = -2,
  GCG_ERR_CUDA = -3,
  GCG_ERR_NCCL = -4,
  GCG_ERR_NOMEM = -5,
  GCG_ERR_UNSUPPORTED = -6
} gcg_status;

/* non-linearities of the hot path (mlpconv.py:186-193; lang2loc.py:285-287) */
typedef enum {
  GCG_ACT_IDENTITY = 0,
  GCG_ACT_RELU = 1,
  GCG_ACT_TANH = 2,
  GCG_ACT_SIGMOID = 3
} gcg_act;

/* dense contraction engines (gcg_gemm_f32 `mode`) */
typedef enum {
  GCG_GEMM_FMA = 0,     /* fp32 FFMA tiles                                   */
  GCG_GEMM_TF32X3 = 1,  /* tcgen05 TF32 tensor cores, 3-term split (~fp32)   */
  GCG_GEMM_TF32 = 2,    /* tcgen05 TF32 single pass (10-bit mantissa inputs) */
  GCG_GEMM_TF32X3_CHAINED = 3 /* TF32X3 with accumulation chains of 4 K blocks summed in fp32 round-to-nearest:
                           the tensor core's round-toward-zero accumulate then costs 16 instead of 75 truncations
                           at K = 600 -- for outputs that cancel (the logits); a few % slower */
} gcg_gemm_mode;

typedef struct gcg_plan gcg_plan; /* opaque: one per sparse matrix */

int gcg_version(void);
const char* gcg_last_error(void);
/* number of kernels this library has launched in this process (bench.py's
 * gpu_launches evidence); gcg_launch_count_reset() zeroes it. */
int64_t gcg_launch_count(void);
void gcg_launch_count_reset(void);

/* ------------------------------------------------------------------ plans */

/* Analyse a device CSR matrix once: finds the "long" rows (degree >
 * long_row_threshold) of the power-law graph and cuts them into fixed-size
 * segments so that no warp ever owns more than long_row_threshold non-zeros
 * (deterministic two-pass reduction, no atomics).  h_indptr may be NULL, in
 * which case indptr is copied back from the device (synchronous).
 * Replaces: nothing in the reference (scipy needs no plan); it is the
 * B200-side preparation for S.dot, lasagne_layers.py:26,65,67,84. */
int gcg_plan_create_csr(int64_t n_rows, int64_t n_cols, int64_t nnz,
                        const int32_t* d_indptr, const int32_t* d_indices,
                        const float* d_vals, const int32_t* h_indptr,
                        int32_t long_row_threshold, gcg_plan** out);
int gcg_plan_destroy(gcg_plan* plan);
/* bytes of scratch gcg_spmm_csr_f32 needs for an output of leading dim ldc */
int64_t gcg_plan_workspace_bytes(const gcg_plan* plan, int64_t ldc);
/* info[0..7] = n_rows, n_cols, nnz, n_long_rows, n_segments, max_degree,
 *              long_row_threshold, reserved */
int gcg_plan_info(const gcg_plan* plan, int64_t* info);

/* ------------------------------------------------------------------- SpMM */

/* C = epilogue( A . B )            A: the plan's CSR [n_rows, n_cols]
 *   P    = A.B (+ C if accumulate)                    S.dot, lasagne_layers.py:26,65,67,84
 *   Hc   = act(P + bias)                              :27-29, :69-71, :86-89
 *   C    = gate ? gate*Hc + (1-gate)*carry : Hc       highway gate (north_star; not in reference)
 *   conv_out (optional) also receives Hc when gating (needed by the backward).
 * The same call serves the backward A^T.dP (A_hat is symmetric, so the forward
 * plan is reused) and X^T.dZ (a plan of the transposed X).
 * Per output element the products are summed in CSR order with separately
 * rounded multiply and add, i.e. bit-identical to scipy's csr_matvecs for rows
 * that are not split; split (long) rows differ only by summation order.
 * B: [n_cols, F] ld ldb;  C: [n_rows, F] ld ldc.  The float4 path needs
 * 16-byte aligned B/C and ldb, ldc multiples of 4 with ld >= round_up(F, 4)
 * (pad columns of C are then overwritten with unspecified values); anything
 * else takes the scalar path.
 * panel_cols: -2 = nnz-balanced streaming variant (csrc/gcg_spmm_stream.cu): every warp executes one "span" of
 * consecutive rows holding an equal share of the non-zeros, gathered rows travel through a deep per-warp
 * pipeline (cp.async shared-memory ring or rotating registers), spans follow the plan's row-block x column-panel
 * schedule (gcg_plan_set_schedule).  Builds its tables on first use -> call once outside graph capture.
 * -1 = whole rows per warp with the gathered rows staged in shared memory by
 * cp.async.bulk (TMA) copies into a per-warp ring (deepest memory-level parallelism; F <= 1024);
 * 0 = whole rows per warp gathered straight into registers; >0 = process the feature dimension in
 * column panels of that many floats, panel-major over the grid, so that one
 * panel of B (n_cols*panel_cols*4 bytes) stays L2-resident while it is gathered.
 * workspace: gcg_plan_workspace_bytes(plan, ldc) bytes (may be NULL if 0). */
int gcg_spmm_csr_f32(const gcg_plan* plan, const float* B, int64_t ldb, int64_t F,
                     float* C, int64_t ldc, const float* bias, int act,
                     int accumulate, const float* gate, int64_t ld_gate,
                     const float* carry, int64_t ld_carry, float* conv_out,
                     int64_t ld_conv, int32_t panel_cols, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* Row-block x column-panel schedule of the streaming variant (panel_cols = -2).  Rows [block_rows[b],
 * block_rows[b+1]) form block b, processed in |block_panels[b]| column panels: > 0 panel-major inside the block
 * (all rows of panel 0, then panel 1, ...: the block's gathered columns x panel bytes stay L2-resident -- 2-D
 * tiling for communities larger than L2), < 0 interleaved (the panels of the same rows run side by side in one
 * CTA).  n_blocks = 0 restores the default (one block, whole rows).  Host arrays; the plan copies them.
 * Results are bit-identical for every schedule (each output element still sums its row in CSR order).
 * Replaces: nothing in the reference; B200-side preparation for S.dot, lasagne_layers.py:67,84. */
int gcg_plan_set_schedule(gcg_plan* plan, int64_t n_blocks, const int32_t* h_block_rows,
                          const int32_t* h_block_panels);
/* L2 residency hint of the streaming variant for a square matrix whose node order keeps communities together:
 * gathered rows whose column lies within near_window_rows of the gathering row are loaded with the L2 evict_last
 * policy (the reusable set), all others with evict_first (long-range edges are never re-read before eviction and
 * would only push the reusable set out).  0 = every gather evict_last (default). */
int gcg_plan_set_near_window(gcg_plan* plan, int32_t near_window_rows);
/* experiment knobs of the streaming variant: kernel variant (0 = default of the width class; 1-4 shared-memory
 * ring depths, 5-8 register pipelines), non-zeros per span (0 = 384), near window (-1 = the plan's) */
void gcg_spmm_stream_tuning(int variant, int span_nnz, int near_window);

/* ------------------------------------------------------------------- GEMM */

/* C = act( op(A).op(B) + beta*C + bias ) [ * act'(mask) ]
 *   op(A): [M,K] (transA=0: stored [M,K] ld lda; transA=1: stored [K,M])
 *   op(B): [K,N] (transB=0: stored [K,N] ld ldb; transB=1: stored [N,K])
 * Replaces T.dot (lasagne_layers.py:82) and its two gradient products
 * (dW = H^T.dZ, dH = dZ.W^T) that theano.grad derives (mlpconv.py:263).
 * mask (optional, [M,N] ld ld_mask): multiplies the result by mask_act'(mask)
 * expressed from the activation OUTPUT (relu: mask>0; tanh: 1-mask^2) -- fuses
 * dP = dA * act'(.) into the producer of dA.
 * split_k > 1 cuts K into that many slices reduced in a fixed order
 * (deterministic); workspace must then hold split_k*M*N floats. 0 = auto. */
int gcg_gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K,
                 const float* A, int64_t lda, const float* B, int64_t ldb,
                 float* C, int64_t ldc, float beta, const float* bias, int act,
                 const float* mask, int64_t ld_mask, int mask_act, int mode,
                 int32_t split_k, void* workspace, int64_t workspace_bytes,
                 void* stream);
int64_t gcg_gemm_workspace_bytes(int transA, int transB, int64_t M, int64_t N,
                                 int64_t K, int mode, int32_t split_k);
/* tf32 hi/lo split of a matrix (the pre-pass of GCG_GEMM_TF32X3): hi = tf32_rn(x), lo = tf32_rn(x - hi),
 * over n_rows*ld floats (hi / lo have x's leading dimension).  gcg_gemm_presplit_f32 is gcg_gemm_f32 with
 * operands split beforehand (NULL pair = split inside the call), so that an activation read by several
 * GEMMs of one training step (H.W, H.W_g, H^T.dZ ...) is split once.  Ignored by the FFMA engine. */
int gcg_tf32_split_f32(const float* x, int64_t ld, int64_t n_rows, float* hi, float* lo, void* stream);
int gcg_gemm_presplit_f32(int transA, int transB, int64_t M, int64_t N, int64_t K,
                          const float* A, int64_t lda, const float* B, int64_t ldb,
                          float* C, int64_t ldc, float beta, const float* bias, int act,
                          const float* mask, int64_t ld_mask, int mask_act, int mode,
                          int32_t split_k, void* workspace, int64_t workspace_bytes, void* stream,
                          const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo);
/* 1 when the tcgen05 engines (GCG_GEMM_TF32X3 / GCG_GEMM_TF32) can run: sm_100 device and
 * a driver that exports cuTensorMapEncodeTiled.  Requests that cannot be described by TMA
 * tensor maps (unaligned base or leading dimension) fall back to the FFMA tiles. */
int gcg_gemm_tc_available(void);

/* ------------------------------------------------- epilogues / reductions */

/* out[j] = sum_i X[i,j]   (db = colsum(dP); bias grads of lasagne_layers.py:27-28,69-70,86-87)
 * deterministic two-pass; workspace >= gcg_colsum_workspace_bytes(n_rows, F). */
int gcg_colsum_f32(const float* X, int64_t ld, int64_t n_rows, int64_t F, float* out,
                   void* workspace, int64_t workspace_bytes, void* stream);
int64_t gcg_colsum_workspace_bytes(int64_t n_rows, int64_t F);

/* dP = dA * act'(A) (from the activation output A), optionally in place (dP == dA).
 * Backward of the non-linearity at lasagne_layers.py:29,71,89. */
int gcg_act_bwd_f32(const float* dA, int64_t ld_da, const float* A, int64_t ld_a,
                    float* dP, int64_t ld_dp, int64_t n_rows, int64_t F, int act,
                    void* stream);

/* Stand-alone highway mix O = g*Hc + (1-g)*Hin (O may alias Hc).  Normally fused into the SpMM
 * epilogue; used where the mix cannot be fused (feature-partitioned multi-GPU propagation). */
int gcg_highway_fwd_f32(const float* Hc, int64_t ld_hc, const float* g, int64_t ld_g,
                        const float* Hin, int64_t ld_hin, float* O, int64_t ld_o, int64_t n_rows,
                        int64_t F, void* stream);

/* Backward of the highway mix  O = g*Hc + (1-g)*Hin,  Hc = act(P):
 *   dP    = g*dO * act'(Hc)          dGpre = dO*(Hc-Hin) * g*(1-g)
 *   dHin  = (1-g)*dO                 (all [n_rows,F]; any output may alias dO)
 * north_star formula; not in the reference. */
int gcg_highway_bwd_f32(const float* dO, int64_t ld_do, const float* g, int64_t ld_g,
                        const float* Hc, int64_t ld_hc, const float* Hin, int64_t ld_hin,
                        float* dP, int64_t ld_dp, float* dGpre, int64_t ld_dg,
                        float* dHin, int64_t ld_dh, int64_t n_rows, int64_t F, int act,
                        void* stream);

/* Output head on the gathered logits L [n_idx, C] (mlpconv.py:216,223,227-233,252):
 *   probs = softmax_rows(L) (max-subtracted);  ce[i] = -log probs[i, y[i]];
 *   pred[i] = argmax (first maximum, like np.argmax);  hit[i] = (pred[i]==y[i]);
 *   G = (probs - onehot(y)) / denom   (gradient of mean CE w.r.t. L; denom = n_idx)
 * probs / G / ce / hit / pred may each be NULL.  y may be NULL when only
 * probs / pred are wanted (predict / predict_proba, mlpconv.py:320-336). */
int gcg_softmax_ce_f32(const float* L, int64_t ld_l, const int32_t* y, int64_t n_idx,
                       int64_t C, float denom, float* probs, int64_t ld_p, float* G,
                       int64_t ld_g, float* ce, float* hit, int64_t* pred, void* stream);

/* out[0] = scale * sum(x[0..n))  -- deterministic (fixed tree), single block. */
int gcg_sum_f32(const float* x, int64_t n, float scale, float* out, void* stream);
/* dst[i] = sum_s src[s*slab_floats + i], s ascending (deterministic): reduces the per-document-block partial rows
 * of the blocked X^T.dZ1 product (Dot.grad of lasagne_layers.py:26,65). */
int gcg_sum_slabs_f32(const float* src, int32_t n_slabs, int64_t slab_floats, float* dst, void* stream);

/* dP[r,:] = sum over the positions p in pos_idx[pos_ptr[r] .. pos_ptr[r+1]) of G[p,:]
 * (zero where a node has no target position).  Deterministic scatter-ADD that is
 * the gradient of the row gather activation[target_indices,:] (lasagne_layers.py:88);
 * duplicates in target_indices accumulate (tensormain.py:226 samples with replacement). */
int gcg_scatter_rows_f32(const float* G, int64_t ld_g, const int32_t* pos_ptr,
                         const int32_t* pos_idx, int64_t n_rows, int64_t C, float* dP,
                         int64_t ld_dp, void* stream);
/* out[i,:] = X[idx[i],:]  (lasagne_layers.py:88) */
int gcg_gather_rows_f32(const float* X, int64_t ld_x, const int32_t* idx, int64_t n_idx,
                        int64_t C, float* out, int64_t ld_out, void* stream);
/* dst[idx[i],:] = src[i,:] for DISTINCT idx (the inverse placement of gcg_gather_rows_f32): the rows of
 * dW1 = X^T.dZ1 (Dot.grad of lasagne_layers.py:26,65) that the dense head block computed, put back in place. */
int gcg_put_rows_f32(const float* src, int64_t ld_src, const int32_t* idx, int64_t n_idx,
                     int64_t C, float* dst, int64_t ld_dst, void* stream);

/* Transposes around the feature-sliced multi-GPU propagation (graphconvgeo_b200/dist.py):
 * pack:   src [n_rows, F] (ld)  ->  dst [P][n_rows][Fp], slice q = columns [q*Fp, (q+1)*Fp), zero padded;
 * unpack: the inverse (pad columns dropped).  Fp and ld multiples of 4, 16-byte aligned operands. */
int gcg_pack_cols_f32(const float* src, int64_t ld, int64_t n_rows, int64_t F, int32_t P, int64_t Fp,
                      float* dst, void* stream);
int gcg_unpack_cols_f32(const float* src, int64_t n_rows, int64_t F, int32_t P, int64_t Fp, float* dst,
                        int64_t ld, void* stream);

/* Peer-memory (NVLink P2P) variant of the two transposes: each rank WRITES its slices straight into the
 * peers' buffers (IPC-mapped device pointers, 128-bit stores over NVLink, one kernel per direction),
 * no staging copy and no collective; the caller orders visibility with a cross-rank barrier.
 *  gcg_peer_alloc/open/close/free: cudaMalloc + 64-byte cudaIpcMemHandle exchange helpers.
 *  push_cols: dst_q[(dst_row0 + i), :Fp] = src[i, q*Fp:(q+1)*Fp]   for every peer q (zero padded)
 *  push_rows: dst_q[slot_offset + i*Fp ...] = src[row_off[q] + i, :Fp], i < row_off[q+1]-row_off[q]
 * h_peer_dst: host array of P device pointers (peer q's buffer as mapped in THIS process). */
int gcg_peer_alloc(int64_t bytes, void** d_ptr, void* handle64);
int gcg_peer_free(void* d_ptr);
int gcg_peer_open(const void* handle64, void** d_ptr);
int gcg_peer_close(void* d_ptr);
/* Cross-GPU barrier in stream order (no collective launch): stores `seq` into slot `rank` of every peer's int32
 * flag array (system-scope fence first: peer stores of earlier kernels on this stream become visible before it),
 * then waits until all P slots of the own array reached `seq`.  seq must grow by one per call on every rank.
 * A peer that never arrives sets *err_flag (device int32) after ~4 s instead of hanging the GPU. */
int gcg_peer_barrier(void* const* h_peer_flags, void* own_flags, int32_t P, int32_t rank, int32_t seq,
                     void* err_flag, void* stream);
/* gcg_spmm_csr_f32 whose output rows are ROUTED to their owners: row r with owner_row_off[q] <= r <
 * owner_row_off[q+1] is written to owner_base[q] + (r - owner_row_off[q]) * ldc.  With peer-mapped owner buffers
 * the second transpose of the feature-sliced propagation (push_rows) rides on the SpMM's own epilogue stores:
 * NVLink traffic overlaps the gather instead of following it.  Same summation order as gcg_spmm_csr_f32
 * (S.dot, lasagne_layers.py:67,84); bias + activation fused, no gate / accumulate. */
int gcg_spmm_csr_routed_f32(const gcg_plan* plan, const float* B, int64_t ldb, int64_t F, int32_t n_owner,
                            void* const* h_owner_base, const int64_t* h_owner_row_off, int64_t ldc,
                            const float* bias, int act, int32_t panel_cols, void* workspace,
                            int64_t workspace_bytes, void* stream);
int gcg_push_cols_f32(const float* src, int64_t ld, int64_t n_rows, int64_t F, int32_t P, int64_t Fp,
                      void* const* h_peer_dst, int64_t dst_row0, void* stream);
int gcg_push_rows_f32(const float* src, const int64_t* h_row_off, int32_t P, int64_t Fp,
                      void* const* h_peer_dst, int64_t slot_offset_floats, void* stream);

/* SpMM tuning knob for experiments (scripts/spmm_sweep.py): u = gathered rows in flight per lane group,
 * minb = __launch_bounds__ min blocks (0 = compiler's choice); (0, 0) restores the defaults. */
void gcg_spmm_set_tuning(int u, int minb);
/* experiment knob of the register-gather kernel: lanes per output row x float4 per lane (0, 0 = automatic) */
void gcg_spmm_set_group(int lanes, int vpl);

/* ---------------------------------------------------------------- optimiser */

/* One fused multi-tensor step of lasagne.updates.adam (mlpconv.py:263) with the
 * elastic-net sub-gradient of mlpconv.py:235-244 folded in:
 *   g' = g + reg_coef[k]*0.5*(sign(p) + 2p);  m,v update;  p -= a_t*m/(sqrt(v)+eps)
 * where a_t = lr*sqrt(1-b2^t)/(1-b1^t) is computed on the DEVICE from the step
 * state d_t (float32[2]: d_t[0] = step counter t, incremented by this call;
 * d_t[1] = a_t, written by this call) so the step is CUDA-graph replayable.
 * reg_out (optional, float[1]): receives sum_k reg_coef[k]*0.5*(|p|_1 + |p|_2^2) of
 * the PRE-update parameters (the penalty term of the loss f_train returns).
 * h_params/h_grads/h_m/h_v: host arrays of n_tensors device pointers; h_sizes: element
 * counts (tensors must be contiguous); h_reg: per-tensor coefficient (0 for biases).
 * workspace >= gcg_adam_workspace_bytes(n_tensors, h_sizes). */
int gcg_adam_step_f32(int32_t n_tensors, float* const* h_params, const float* const* h_grads,
                      float* const* h_m, float* const* h_v, const int64_t* h_sizes,
                      const float* h_reg, float lr, float beta1, float beta2, float eps,
                      float* d_t, float* reg_out, void* workspace, int64_t workspace_bytes,
                      void* stream);
int64_t gcg_adam_workspace_bytes(int32_t n_tensors, const int64_t* h_sizes);

/* out[0] = sum_k h_reg[k]*0.5*(|p_k|_1 + |p_k|_2^2): the elastic-net penalty of
 * mlpconv.py:235-245 alone (eval_loss of f_val adds it without an update).
 * workspace >= gcg_adam_workspace_bytes(n_tensors, h_sizes). */
int gcg_elastic_net_f32(int32_t n_tensors, const float* const* h_params, const int64_t* h_sizes,
                        const float* h_reg, float* out, void* workspace, int64_t workspace_bytes,
                        void* stream);

/* -------------------------------------------------- multi-GPU (NCCL over NVLink) */

/* One process per GPU, A_hat row-partitioned (the reference is single-process).  NCCL is bound
 * at run time (dlopen of libnccl.so.2: inside a PyTorch process the copy torch already loaded), so single-GPU users
 * of this library never touch it.  The unique id is made on rank 0 (gcg_comm_unique_id, 128 bytes) and handed to
 * the other ranks by the launcher's own means (the Python shim broadcasts it with torch.distributed). */
typedef struct gcg_comm gcg_comm;
int gcg_comm_unique_id(void* id128);
/* collective over all ranks; binds the communicator to the CURRENT device and creates its communication stream */
int gcg_comm_init(const void* id128, int32_t world, int32_t rank, gcg_comm** out);
int gcg_comm_destroy(gcg_comm* comm);
int gcg_comm_info(const gcg_comm* comm, int32_t* world, int32_t* rank);
/* `stream` waits for every collective issued so far on the communicator (no host synchronisation) */
int gcg_comm_wait(gcg_comm* comm, void* stream);
/* in-place all-gather of `full` = [world][floats_per_rank]: slab `rank` is this rank's contribution; starts when the
 * work already enqueued on `stream` is done; wait != 0: `stream` also waits for the result (else gcg_comm_wait) */
int gcg_allgather_rows_f32(gcg_comm* comm, float* full, int64_t floats_per_rank, int32_t wait, void* stream);
/* sum over ranks, in place, of n_tensors gradient buffers in one NCCL group (theano.grad of a loss that is a mean
 * over ALL targets, mlpconv.py:230,263: each rank holds the partial sum over its rows); wait as above */
int gcg_allreduce_grads_f32(gcg_comm* comm, int32_t n_tensors, float* const* h_bufs, const int64_t* h_sizes,
                            int32_t wait, void* stream);
/* One propagation S.dot(H, Z) (lasagne_layers.py:67,84) for the row block of this rank -- north_star's design:
 * Z_full = [world*n_loc, ld] holds this rank's slab at rows [rank*n_loc, (rank+1)*n_loc); it is all-gathered in place
 * on the communication stream WHILE `stream` runs C = diag . Z_full (`diag`: the columns of the local slab), then
 * C = epilogue(C + off . Z_full) (`off`: all other columns) once the gather has landed.  Both plans are
 * [n_rows_local x world*n_loc]; epilogue arguments as in gcg_spmm_csr_f32. */
int gcg_spmm_rowpart_allgather_f32(gcg_comm* comm, const gcg_plan* diag, const gcg_plan* off, float* Z_full,
                                   int64_t ld, int64_t F, int64_t n_loc, float* C, int64_t ldc,
                                   const float* bias, int act, const float* gate, int64_t ld_gate,
                                   const float* carry, int64_t ld_carry, float* conv_out, int64_t ld_conv,
                                   int32_t panel_cols, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ epoch */

/* The epoch as one native object.  The reference's epoch is ONE call of a compiled
 * Theano function: f_train = theano.function([X_sym, y_sym, train_indices_sym], [loss, acc], updates=...)
 * (mlpconv.py:265), called once per epoch (mlpconv.py:295).  A gcg_epoch is the counterpart of that compiled
 * object: the ordered list of libgcg calls of one f_train (forward a1-a4/a9, head a10, backward a5-a8, elastic
 * net + Adam a11-a12), each with its arguments, recorded ONCE from the layer code between
 * gcg_epoch_record_begin / gcg_epoch_record_end on the calling thread (the calls still execute while being
 * recorded; host-side tables such as Adam's pointer lists are copied), and replayed by gcg_epoch_run() on the
 * stream it is given: one C call per epoch, every launch enqueued from C++, CUDA-graph capturable.
 * The caller keeps every buffer the recorded calls point to alive and unchanged in place (the same rule a CUDA
 * graph has).  Recorded: every stream-taking compute entry point of this header except the peer-memory ones
 * (gcg_push_*, gcg_peer_barrier, gcg_spmm_csr_routed_f32 -- their sequence numbers change per call). */
typedef struct gcg_epoch gcg_epoch;
int gcg_epoch_create(gcg_epoch** out);
int gcg_epoch_destroy(gcg_epoch* epoch);
int gcg_epoch_record_begin(gcg_epoch* epoch);
int gcg_epoch_record_end(gcg_epoch* epoch);
/* number of recorded calls, and the entry-point name of call i ("" out of range) */
int64_t gcg_epoch_size(const gcg_epoch* epoch);
const char* gcg_epoch_call_name(const gcg_epoch* epoch, int64_t i);
/* enqueue the recorded calls, in order, on `stream`; stops at the first failing call and returns its status */
int gcg_epoch_run(const gcg_epoch* epoch, void* stream);

/* --------------------------------------------------------------- host side */

/* k-d tree region labels, bit-exact restatement of kdtree.py:84-118,126-147
 * (float64 compares, np.median split, left-before-right leaf numbering). */
int gcg_kdtree_fit_host(const double* h_points, int64_t n, int32_t dims, int64_t bucket_size,
                        int64_t* h_labels, int64_t* n_leaves);

/* A_hat = D^-1/2 (A, unit diagonal) D^-1/2 in float64, cast to float32 --
 * tensormain.py:170-180,221.  Input: CSR pattern of the adjacency (sorted
 * columns; weights NULL = binary).  Call gcg_ahat_nnz_host first to size the
 * outputs (rows lacking a diagonal entry gain one). */
int64_t gcg_ahat_nnz_host(int64_t n, const int32_t* h_indptr, const int32_t* h_indices);
int gcg_ahat_build_host(int64_t n, const int32_t* h_indptr, const int32_t* h_indices,
                        const double* h_weights, int32_t* out_indptr, int32_t* out_indices,
                        float* out_vals);

/* Device version of the two calls above for a BINARY adjacency pattern that is already a device CSR
 * (sorted columns): phase 1 writes out_indptr[n+1] and the output nnz (*d_total, device int32);
 * phase 2 writes the column indices (unit diagonal inserted) and the float32 values.  Bit-identical
 * to gcg_ahat_build_host.  workspace >= gcg_ahat_device_workspace_bytes(n), the same buffer for both. */
int64_t gcg_ahat_device_workspace_bytes(int64_t n);
int gcg_ahat_indptr_device(int64_t n, const int32_t* d_indptr, const int32_t* d_indices,
                           int32_t* out_indptr, int32_t* d_total, void* workspace,
                           int64_t workspace_bytes, void* stream);
int gcg_ahat_fill_device(int64_t n, const int32_t* d_indptr, const int32_t* d_indices,
                         const int32_t* out_indptr, int32_t* out_indices, float* out_vals,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* Label pipeline around the hot path, float64 haversine (R = 6371.0088 km):
 *  nearest: out_idx[i] = argmin_c haversine(points[i], medians[c]) (first minimum) -- the brute-force
 *           1-NN that assigns dev/test users to regions, data.py:416-419; out_km optional.
 *  pairs:   out_km[i] = haversine(a[i], b[i]) -- the per-user error of geo_eval, tensormain.py:38-54.
 * points / medians / a / b are [n, 2] (lat, lon) float64 device arrays. */
int gcg_haversine_nearest_f64(const double* d_points, int64_t n, const double* d_medians, int32_t n_medians,
                              int64_t* d_out_idx, double* d_out_km, void* stream);
int gcg_haversine_pairs_f64(const double* d_a, const double* d_b, int64_t n, double* d_out_km, void* stream);

/* CSR transpose (stable: rows ascending inside each output row), used once per
 * fit to build X^T for dW1 = X^T.dZ1 (Dot.grad of lasagne_layers.py:26,65). */
int gcg_csr_transpose_host(int64_t n_rows, int64_t n_cols, const int32_t* h_indptr,
                           const int32_t* h_indices, const float* h_vals, int32_t* t_indptr,
                           int32_t* t_indices, float* t_vals);

/* out CSR = rows idx[0..n_idx) of the input CSR (duplicates allowed).  Used to turn
 * activation[target_indices,:] of a propagated matrix into a propagation with
 * A_hat[target_indices,:] so that only the wanted rows are ever computed.
 * Two-call protocol: out_indices == NULL returns the nnz needed. */
int64_t gcg_csr_gather_rows_host(int64_t n_rows, const int32_t* h_indptr,
                                 const int32_t* h_indices, const float* h_vals,
                                 const int32_t* idx, int64_t n_idx, int32_t* out_indptr,
                                 int32_t* out_indices, float* out_vals);

/* Column-blocked split of selected rows (document-blocked X^T.dZ, see graphconvgeo_b200/sparse.py
 * BlockedRows): block b keeps columns [b*block_cols, (b+1)*block_cols).  out_indptr is
 * [n_blocks][n_sel+1] with offsets relative to the block's slice of out_indices / out_vals, whose
 * boundaries are out_block_off[n_blocks+1].  Call with out_indices == NULL to size the outputs. */
int gcg_csr_split_colblocks_host(const int32_t* h_indptr, const int32_t* h_indices, const float* h_vals,
                                 const int32_t* row_sel, int64_t n_sel, int64_t block_cols, int64_t n_blocks,
                                 int32_t* out_indptr, int64_t* out_block_off, int32_t* out_indices,
                                 float* out_vals);

/* Node reordering for gather locality: out = P A Q^T where row i of out is row
 * order[i] of A and column j of A becomes col_map[j] (col_map == NULL keeps the
 * columns); columns are re-sorted inside every row.  The GCN is permutation
 * equivariant, so running it on (P A_hat P^T, P X) and mapping target_indices
 * through P changes no result of lasagne_layers.py:60-89, only which dense rows
 * sit next to each other in HBM / L2.  Output arrays have the input's nnz. */
int gcg_csr_permute_host(int64_t n_rows, const int32_t* h_indptr, const int32_t* h_indices,
                         const float* h_vals, const int32_t* order, const int32_t* col_map,
                         int32_t* out_indptr, int32_t* out_indices, float* out_vals);

/* ------------------------------------------- input smoothing + minibatches */
/* One-shot smoothing X_conv = A_hat * X (sparse x sparse) -- main.py:528-530,
 * tensormain.py:112-114: `X_conv = H * X; X_conv = X_conv.tocsr().astype('float32')`.
 * There H is FLOAT64 and X float32, so scipy's csr_matmat accumulates every output entry in
 * float64, in the order the rows j of A's row i are visited (CSR order); the result is rounded to
 * float32 once and -- because scipy's astype() canonicalises -- ends up with ascending columns in every
 * row.  The two kernels reproduce exactly that: one warp per output row, A's entries strictly in stored
 * order, the 32 lanes over the entries of B's row j, a per-warp dense accumulator plus a bitmap of touched
 * columns (workspace), then one sweep of the bitmap emits the row in ascending column order.  Values are
 * bit-identical to scipy.  Differences: entries whose sum is exactly 0.0 are kept (scipy's csr_matmat
 * drops them); B must not hold duplicate column entries inside a row.
 *
 * Protocol: (1) gcg_spgemm_count_csr (drop_diagonal = 0) -> row_nnz[n_rows]; (2) caller builds c_indptr
 * (int64[n_rows+1], exclusive scan) and allocates c_indices / c_vals; (3) gcg_spgemm_fill_csr_f32.
 * a_vals is float64 when a_is_f64 != 0 (float64 sums, the reference's case), else float32 (float32
 * sums, scipy's promotion rule for float32 * float32).
 * Workspace: gcg_spgemm_workspace_bytes(n_cols_b); zero-initialised by the calls themselves. */
int64_t gcg_spgemm_workspace_bytes(int64_t n_cols_b);
int gcg_spgemm_count_csr(int64_t n_rows, int64_t n_cols_b, const int32_t* a_indptr, const int32_t* a_indices,
                         const int32_t* b_indptr, const int32_t* b_indices, int drop_diagonal, int32_t* row_nnz,
                         void* workspace, int64_t workspace_bytes, void* stream);
int gcg_spgemm_fill_csr_f32(int64_t n_rows, int64_t n_cols_b, const int32_t* a_indptr, const int32_t* a_indices,
                            const void* a_vals, int a_is_f64, const int32_t* b_indptr, const int32_t* b_indices,
                            const float* b_vals, const int64_t* c_indptr, int32_t* c_indices, float* c_vals,
                            void* workspace, int64_t workspace_bytes, void* stream);

/* Pattern-only product (no values): the sorted column pattern of A * B, optionally without the diagonal.
 * With A = R^T, B = R (R[m, t] = 1 when target user t is a neighbour of node m in the @-mention graph,
 * self loops included) this is the user-user graph of efficient_collaboration_weighted_projected_graph2
 * (data.py:226-250): every node's target neighbours become a clique.  Count with gcg_spgemm_count_csr
 * (same drop_diagonal), then fill. */
int gcg_spgemm_fill_pattern_csr(int64_t n_rows, int64_t n_cols_b, const int32_t* a_indptr, const int32_t* a_indices,
                                const int32_t* b_indptr, const int32_t* b_indices, int drop_diagonal,
                                const int64_t* c_indptr, int32_t* c_indices, void* workspace,
                                int64_t workspace_bytes, void* stream);

/* Device-side minibatch slicing, `inputs[excerpt]` of iterate_minibatches (mlp.py:81-91):
 * out CSR = rows d_rows[0..n_sel) of the input CSR, entry order inside a row preserved.
 * out_indptr (int32[n_sel+1], the exclusive scan of the selected rows' lengths) is supplied by the
 * caller; indptr / out_indptr may hold absolute offsets into indices / out_indices. */
int gcg_csr_gather_rows_device(const int32_t* d_indptr, const int32_t* d_indices, const float* d_vals,
                               const int32_t* d_rows, int64_t n_sel, const int32_t* d_out_indptr,
                               int32_t* d_out_indices, float* d_out_vals, void* stream);

/* Device-side stable CSR transpose (rows ascending inside each output row) of a [n_rows, n_cols]
 * matrix with nnz entries starting at d_indptr[0]; gives X_batch^T for dW = X_batch^T.dP
 * (Dot.grad of lasagne_layers.py:26) without a host round trip.  Uses a radix sort (CUB) on the
 * column index.  Workspace: gcg_csr_transpose_device_workspace_bytes(n_rows, n_cols, nnz). */
int64_t gcg_csr_transpose_device_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t nnz);
int gcg_csr_transpose_device(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* d_indptr,
                             const int32_t* d_indices, const float* d_vals, int32_t* t_indptr,
                             int32_t* t_indices, float* t_vals, void* workspace, int64_t workspace_bytes,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCG_H_ */
