// tcgen05 (5th-gen tensor core) path for the dense projections of the GCN hot path.
//
//   C = act(op(A).op(B) + beta*C + bias) [* act'(mask)],  fp32 in / fp32 out,
//   TF32 tensor-core products with fp32 accumulation in TMEM.
//
// GCG_GEMM_TF32X3 (default engine for contraction-bound shapes): every operand x is split as
//   x = hi + lo,  hi = tf32(x) (low 13 mantissa bits cleared),  lo = x - hi  (exact in fp32),
// and three MMAs are accumulated per K step:  hi.hi + lo.hi + hi.lo  (lo.lo ~ 2^-22 is dropped).
// That restores fp32-level accuracy (north_star asks 1e-4 relative; plain TF32 gives ~1e-3) at a
// third of the TF32 rate -- still several times the FFMA tiles.  GCG_GEMM_TF32 issues hi.hi only.
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of the A/B (hi, lo) tiles into a
//               SWIZZLE_128B shared-memory ring, completion on mbarriers (expect_tx)
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::tf32
//               (M=128, N=128, K=8) from shared-memory descriptors; tcgen05.commit frees the
//               ring slot / publishes the accumulator
//   warp 2      TMEM allocator (2 accumulator stages x 128 columns)
//   warps 4-7   epilogue: tcgen05.ld 32x32b.x32 -> registers -> bias/act/mask -> global,
//               overlapped with the next tile's main loop through the second TMEM stage.
//               The accumulator holds the TRANSPOSED tile (operands swapped at issue) so that a
//               warp's lanes are consecutive columns of C: all epilogue traffic is coalesced.
// Both operand majors are handled by descriptors (K-major and MN-major SWIZZLE_128B canonical
// layouts), so NN / TN / NT / TT need no transposes.  Split-K (weight gradients: K = #nodes)
// writes per-slice partials that are reduced in fixed order (deterministic).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "gcg_gemm.cuh"

namespace gcg {

constexpr int TM = 128, TN = 128, TK = 32;       // CTA tile; TK*4 B = the 128 B swizzle span
constexpr int UK = 8;                            // K of one tcgen05.mma.kind::tf32
constexpr int TILE_BYTES = TM * TK * 4;          // 16 KB per operand tile
constexpr int ACC_STAGES = 2;
// TMEM: per accumulator stage TWO D^T tiles (TN lanes x TM columns each): the hi.hi products and, for 3xTF32,
// the two cross terms in a tile of their own.  The tensor core adds into D with round-toward-zero, one
// truncation per MMA; with all three terms in one accumulator a K = 600 chain made 225 truncating adds
// (measured error 2.5e-5 of the output magnitude, over the 1e-4 relative bound on cancelling logits).  The
// cross terms are 2^-11 smaller: kept apart they cost the big accumulator nothing and its chain is 3x shorter.
// The epilogue adds the two tiles in fp32 round-to-nearest.
constexpr int TMEM_COLS = ACC_STAGES * 2 * TM;   // 512 = all of TMEM
constexpr int TC_THREADS = 256;

int launch_splitk_reduce(const GemmArgs& g, cudaStream_t st);   // gcg_gemm.cu

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// multicast flavours for 2-CTA clusters: the TMA write lands at the same shared-memory offset (and signals the
// mbarrier at the same offset) in every CTA of `mask`; the commit arrives on the barrier of every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// Shared-memory matrix descriptor (SM100 UMMA), SWIZZLE_128B canonical layouts:
//   K-major : rows of 128 B (32 tf32 along K), 8-row groups SBO = 1024 B apart; LBO unused (1)
//   MN-major: K rows of 128 B (32 tf32 along M/N); 32-wide MN blocks LBO apart, 8-K groups SBO = 1024 B apart
// MN-major 32-bit (tf32) operands only exist as SWIZZLE_128B_BASE32B (layout type 1, Swizzle<2,5,2>):
// 4-K-row groups of 128 B rows, SBO = 512 B between groups; TMA writes it with SWIZZLE_128B_ATOM_32B.
template <bool MN>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)(((MN ? 4096u : 16u) >> 4) & 0x3fff) << 16;    // LBO
  d |= (uint64_t)(((MN ? 512u : 1024u) >> 4) & 0x3fff) << 32;   // SBO
  d |= (uint64_t)1 << 46;                                       // descriptor version (Blackwell)
  d |= (uint64_t)(MN ? 1 : 2) << 61;                            // SWIZZLE_128B_BASE32B : SWIZZLE_128B
  return d;
}

// compact (non-inlined) activation for the general epilogue path: keeps the unrolled body small
__device__ __noinline__ float epi_act(float x, int act) {
  switch (act) {
    case GCG_ACT_RELU: return fmaxf(x, 0.f);
    case GCG_ACT_TANH: return tanhf(x);
    case GCG_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}

struct TcArgs {
  GemmArgs g;
  int m_tiles, n_tiles, splits, num_kb_total;   // K blocks over the whole K
  int kb_per_split;
  int x3;                                       // 1: hi.hi + lo.hi + hi.lo ; 0: hi.hi
  int chunk_kb;                                 // 3xTF32: K blocks per accumulation chain (0 = one chain per tile)
};

template <bool A_MN, bool B_MN, int CS, bool CHAINED>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tiles_per_stage = a.x3 ? 4 : 2;
  const int stages = a.x3 ? 3 : 6;
  const uint32_t stage_bytes = tiles_per_stage * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* full = bars;                 // [stages]
  uint64_t* empty = bars + 8;            // [stages]
  uint64_t* acc_full = bars + 16;        // [ACC_STAGES]
  uint64_t* acc_empty = bars + 18;       // [ACC_STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Clusters of CS = 2 CTAs work on two M tiles of the SAME N tile and K range: each CTA loads its own A tile and
  // HALF of the B (weight) tile, multicast into both CTAs' rings.  The kernel is bound by L2 -> SM operand traffic
  // (64 KB per K block per CTA for 768 MMA cycles; measured 150-180 TFLOP/s effective = the L2 cap, not the tensor
  // pipe): sharing B cuts it to 48 KB.  CS = 1 (plain launch) degenerates to the single-CTA schedule.
  constexpr int cs = CS;                      // compile-time: the single-CTA kernel carries none of the cluster code
  const int crank = (CS > 1) ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
  const int m_groups = (a.m_tiles + cs - 1) / cs;
  const int total_tiles = m_groups * a.n_tiles * a.splits;          // tile GROUPS (one tile per CTA of the cluster)
  const int tile0 = blockIdx.x / cs, tile_step = gridDim.x / cs;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
    if (a.x3) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, cs); }   // slot free = every CTA's MMAs retired
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (cs > 1) cluster_sync_all();            // the peer's barriers are initialised before anything signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = tile0; t < total_tiles; t += tile_step) {
        const int nt = t % a.n_tiles, mt = ((t / a.n_tiles) % m_groups) * cs + crank, z = t / (a.n_tiles * m_groups);
        const int m0 = mt * TM, n0 = nt * TN;     // mt may be one past the last M tile (odd count): TMA zero-fills
        const int kb0 = z * a.kb_per_split, kb1 = min(a.num_kb_total, kb0 + a.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* sA = smem + stage * stage_bytes;
          uint8_t* sB = sA + (a.x3 ? 2 : 1) * TILE_BYTES;
          mbar_expect_tx(full + stage, stage_bytes);
          const int k0 = kb * TK;
          if (!A_MN) {
            tma_load_2d(&tmAh, full + stage, sA, k0, m0);                       // box {32 K, 128 M}
            if (a.x3) tma_load_2d(&tmAl, full + stage, sA + TILE_BYTES, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                        // box {32 M, 32 K} x4
              tma_load_2d(&tmAh, full + stage, sA + j * 4096, m0 + 32 * j, k0);
              if (a.x3) tma_load_2d(&tmAl, full + stage, sA + TILE_BYTES + j * 4096, m0 + 32 * j, k0);
            }
          }
          if (cs == 1) {
            if (!B_MN) {
              tma_load_2d(&tmBh, full + stage, sB, k0, n0);                       // box {32 K, 128 N}
              if (a.x3) tma_load_2d(&tmBl, full + stage, sB + TILE_BYTES, k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                tma_load_2d(&tmBh, full + stage, sB + j * 4096, n0 + 32 * j, k0);
                if (a.x3) tma_load_2d(&tmBl, full + stage, sB + TILE_BYTES + j * 4096, n0 + 32 * j, k0);
              }
            }
          } else {
            // this CTA's half of the B tile (N rows [64*crank, 64*crank + 64)), delivered to both CTAs
            if (!B_MN) {
              const int hoff = crank * (TILE_BYTES / 2);                        // box {32 K, 64 N}: 8 KB
              tma_load_2d_mc(&tmBh, full + stage, sB + hoff, k0, n0 + 64 * crank, cmask);
              if (a.x3) tma_load_2d_mc(&tmBl, full + stage, sB + TILE_BYTES + hoff, k0, n0 + 64 * crank, cmask);
            } else {
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const int j = 2 * crank + jj;
                tma_load_2d_mc(&tmBh, full + stage, sB + j * 4096, n0 + 32 * j, k0, cmask);
                if (a.x3) tma_load_2d_mc(&tmBl, full + stage, sB + TILE_BYTES + j * 4096, n0 + 32 * j, k0, cmask);
              }
            }
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    if (lane == 0) {
      // instruction descriptor: D=F32 (c_format 1), A/B = TF32 (2), majors, N>>3, M>>4
      // Operands are SWAPPED: the UMMA "A" (M side, TMEM lanes) is our B tile (n), the UMMA "B" (N side,
      // TMEM columns) is our A tile (m), i.e. the accumulator holds C^T -- see the epilogue.
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 15) |
                             ((A_MN ? 1u : 0u) << 16) | ((uint32_t)(TM >> 3) << 17) | ((uint32_t)(TN >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = tile0; t < total_tiles; t += tile_step) {
        const int z = t / (a.n_tiles * m_groups);
        const int kb0 = z * a.kb_per_split, kb1 = min(a.num_kb_total, kb0 + a.kb_per_split);
        const int chunk = (CHAINED && a.chunk_kb > 0) ? a.chunk_kb : (kb1 - kb0);
        mbar_wait(acc_empty + acc, acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t d_tmem = tmem_base + acc * 2 * TM;
        uint32_t d_small = d_tmem + TM;                     // cross terms lo.hi + hi.lo (3xTF32 only)
        int chain0 = kb0;                                   // first K block of the current accumulation chain
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kb - chain0 == chunk) {
            // chain complete: hand this accumulator stage to the epilogue (which adds it to its running sums in
            // fp32 round-to-nearest) and continue in the other stage -- the tensor core's round-toward-zero
            // accumulate errs with the length of a chain, so chains are kept short
            umma_commit(acc_full + acc);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            mbar_wait(acc_empty + acc, acc_phase ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            d_tmem = tmem_base + acc * 2 * TM;
            d_small = d_tmem + TM;
            chain0 = kb;
          }
          mbar_wait(full + stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sA = smem_u32(smem + stage * stage_bytes);
          const uint32_t sB = sA + (a.x3 ? 2 : 1) * TILE_BYTES;
#pragma unroll
          for (int k = 0; k < TK / UK; ++k) {
            // K-major: +32 B per K step inside the 128 B swizzle row; MN-major: +1024 B (8 K rows)
            const uint32_t aoff = A_MN ? k * 1024 : k * 32;
            const uint32_t boff = B_MN ? k * 1024 : k * 32;
            const uint64_t ah = make_desc<A_MN>(sA + aoff);
            const uint64_t bh = make_desc<B_MN>(sB + boff);
            const uint32_t first = (kb > chain0 || k > 0) ? 1u : 0u;
            if (a.x3) {
              const uint64_t al = make_desc<A_MN>(sA + TILE_BYTES + aoff);
              const uint64_t bl = make_desc<B_MN>(sB + TILE_BYTES + boff);
              umma_tf32(d_small, bh, al, idesc, first);
              umma_tf32(d_small, bl, ah, idesc, 1u);
              umma_tf32(d_tmem, bh, ah, idesc, first);
            } else {
              umma_tf32(d_tmem, bh, ah, idesc, first);
            }
          }
          if (cs == 1) umma_commit(empty + stage);       // frees the ring slot when the MMAs retire
          else umma_commit_mc(empty + stage, cmask);     // ... in both CTAs: the peer's multicast writes into it too
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full + acc);                     // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ========================================================================= epilogue
    // The MMA computes the TRANSPOSED tile D[n][m] (operands swapped, see the issuer), so a TMEM
    // lane -- hence a thread -- is a column n of C and its registers run along m: for a fixed m
    // the 32 lanes of a warp touch 32 consecutive floats of one row of C -> every global access
    // of the epilogue (C, beta*C, bias, mask, split-K partials) is a single 128 B wavefront.
    const GemmArgs& g = a.g;
    const int ew = warp & 3;                             // TMEM lane quarter owned by this warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = tile0; t < total_tiles; t += tile_step) {
      const int nt = t % a.n_tiles, mt = ((t / a.n_tiles) % m_groups) * cs + crank, z = t / (a.n_tiles * m_groups);
      const int64_t n = (int64_t)nt * TN + ew * 32 + lane;
      const int64_t m0 = (int64_t)mt * TM;
      const bool n_ok = n < g.N;
      const float bias_n = (g.bias && n_ok && a.splits == 1) ? __ldg(g.bias + n) : 0.f;
      const bool relu = g.act == GCG_ACT_RELU;
      const bool simple = g.beta == 0.f && g.mask == nullptr && (g.act == GCG_ACT_IDENTITY || relu);
      const bool gate_like = g.beta == 0.f && g.mask == nullptr && g.act == GCG_ACT_SIGMOID;
      const bool accum_like = g.act == GCG_ACT_IDENTITY && g.bias == nullptr &&
                              (g.mask == nullptr || g.mask_act == GCG_ACT_RELU || g.mask_act == GCG_ACT_TANH);
      const int kb0e = z * a.kb_per_split, kb1e = min(a.num_kb_total, kb0e + a.kb_per_split);
      const int n_chains = (CHAINED && a.chunk_kb > 0) ? (kb1e - kb0e + a.chunk_kb - 1) / a.chunk_kb : 1;
      mbar_wait(acc_full + acc, acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (CHAINED && n_chains > 1) {
        // running sums of the chains in registers (fp32 round-to-nearest); the total goes back into the LAST
        // chain's TMEM stage so that the (code-size critical) store phase below stays as it is
        float sum[TM / 32][32];
#pragma unroll
        for (int cg = 0; cg < TM / 32; ++cg)
#pragma unroll
          for (int e = 0; e < 32; ++e) sum[cg][e] = 0.f;
        for (int ci = 0; ci < n_chains; ++ci) {
          if (ci > 0) {
            mbar_wait(acc_full + acc, acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          const uint32_t tb = tmem_base + acc * 2 * TM + ((uint32_t)(ew * 32) << 16);
#pragma unroll
          for (int cg = 0; cg < TM / 32; ++cg) {
            uint32_t v[32], w[32];
            tmem_ld32(tb + cg * 32, v);
            tmem_ld32(tb + TM + cg * 32, w);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 32; ++e)
              sum[cg][e] = __fadd_rn(sum[cg][e], __fadd_rn(__uint_as_float(v[e]), __uint_as_float(w[e])));
          }
          if (ci + 1 < n_chains) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + acc);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
          } else {
#pragma unroll
            for (int cg = 0; cg < TM / 32; ++cg) {
              uint32_t o[32], zr[32];
#pragma unroll
              for (int e = 0; e < 32; ++e) { o[e] = __float_as_uint(sum[cg][e]); zr[e] = 0u; }
              tmem_st32(tb + cg * 32, o);
              tmem_st32(tb + TM + cg * 32, zr);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
        }
      }
#pragma unroll 1
      for (int c = 0; c < TM; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + acc * 2 * TM + c + ((uint32_t)(ew * 32) << 16), v);
        if (a.x3) {
          uint32_t w[32];
          tmem_ld32(tmem_base + acc * 2 * TM + TM + c + ((uint32_t)(ew * 32) << 16), w);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__fadd_rn(__uint_as_float(v[e]), __uint_as_float(w[e])));
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        // The bodies below are kept tiny on purpose: the 32-way unroll is needed to keep v[] in
        // registers, and a fat per-element body (bounds + beta + act switch + mask) made the kernel
        // 85 KB of SASS whose instruction-cache misses cost 27 us per tile (ncu: no_instruction stalls).
        const int64_t mrow = m0 + c;
        const int rows = (int)min((int64_t)32, g.M - mrow);
        if (!n_ok || rows <= 0) continue;
        if (a.splits > 1) {
          float* dst = g.part + ((int64_t)z * g.M + mrow) * g.N + n;
          if (rows == 32) {
#pragma unroll
            for (int e = 0; e < 32; ++e) { *dst = __uint_as_float(v[e]); dst += g.N; }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) { if (e < rows) *dst = __uint_as_float(v[e]); dst += g.N; }
          }
        } else if (simple) {
          float* dst = g.C + mrow * g.ldc + n;
          if (rows == 32) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float o = __uint_as_float(v[e]) + bias_n;
              *dst = relu ? fmaxf(o, 0.f) : o;
              dst += g.ldc;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float o = __uint_as_float(v[e]) + bias_n;
              if (e < rows) *dst = relu ? fmaxf(o, 0.f) : o;
              dst += g.ldc;
            }
          }
        } else if (gate_like) {
          // act(x + bias) with act = sigmoid (the highway gate), beta = 0, no mask
          float* dst = g.C + mrow * g.ldc + n;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float o = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(v[e]) + bias_n)));
            if (e < rows) *dst = o;
            dst += g.ldc;
          }
        } else if (accum_like) {
          // (x + beta*C) [* act'(mask)], identity activation: the backward dH products
          float* dst = g.C + mrow * g.ldc + n;
          const float* msk = g.mask ? g.mask + mrow * g.ld_mask + n : dst;
          const int64_t mstep = g.mask ? g.ld_mask : g.ldc;
          const bool has_mask = g.mask != nullptr, mrelu = g.mask_act == GCG_ACT_RELU;
          const bool has_beta = g.beta != 0.f;
          // two phases per half chunk: all loads first (16 independent 128 B wavefronts in flight),
          // then the math and the stores -- a load/store-interleaved loop serialises on possible aliasing
#pragma unroll
          for (int h = 0; h < 32; h += 16) {
            float cv[16], mk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const bool ok = (h + e) < rows;
              cv[e] = (has_beta && ok) ? __ldcg(dst + (int64_t)e * g.ldc) : 0.f;
              mk[e] = (has_mask && ok) ? __ldcg(msk + (int64_t)e * mstep) : 1.f;
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float o = fmaf(g.beta, cv[e], __uint_as_float(v[h + e]));
              if (has_mask) o *= mrelu ? (mk[e] > 0.f ? 1.f : 0.f) : (1.f - mk[e] * mk[e]);
              if ((h + e) < rows) dst[(int64_t)e * g.ldc] = o;
            }
            dst += 16 * g.ldc;
            msk += 16 * mstep;
          }
        } else {
          float* dst = g.C + mrow * g.ldc + n;
          const float* msk = g.mask ? g.mask + mrow * g.ld_mask + n : nullptr;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (e < rows) {
              float o = __uint_as_float(v[e]);
              if (g.beta != 0.f) o = fmaf(g.beta, *dst, o);
              o = epi_act(o + bias_n, g.act);
              if (msk) o *= act_grad_from_out(*msk, g.mask_act);
              *dst = o;
            }
            dst += g.ldc;
            if (msk) msk += g.ld_mask;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (cs > 1) cluster_sync_all();            // no CTA leaves while its peer may still multicast into it
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// hi = tf32 round-to-nearest of x, lo = tf32_rn(x - hi)  (x - hi is exact in fp32, |lo| <= 2^-11 |x|).
// Round-to-nearest keeps the dropped lo.lo term and the representation error of lo UNBIASED; a
// truncating split makes every product err towards zero and the error grows like sum|a.b|
// (measured: 3e-5 at K=600) instead of sqrt(K).
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                         float* __restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 h, l;
    h.x = tf32_rn(v.x); l.x = tf32_rn(v.x - h.x);
    h.y = tf32_rn(v.y); l.y = tf32_rn(v.y - h.y);
    h.z = tf32_rn(v.z); l.z = tf32_rn(v.z - h.z);
    h.w = tf32_rn(v.w); l.w = tf32_rn(v.w - h.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
}

int tf32_split_launch(const float* x, int64_t n_floats, float* hi, float* lo, cudaStream_t st) {
  const int64_t n4 = n_floats / 4;
  if (n4 == 0) return GCG_OK;
  split_tf32_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n4, 256), (int64_t)kNumSMs * 16), 256, 0, st>>>(x, hi, lo, n4);
  GCG_LAUNCH_CHECK();
  return GCG_OK;
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

// matrix stored [rows][cols] (cols contiguous, leading dim ld); box = {32 cols, box_rows}
static bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                     bool mn_major) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}


// The tensor core accumulates in fp32 with round-toward-zero, so the error of one accumulation
// chain grows linearly with its length (measured ~2.5e-5 relative at K=600, 2e-4 at K=66k).
// Chains are therefore capped at kMaxChainKB K-blocks (K = 8192): longer contractions (the weight
// gradients, K = number of nodes) are cut into more split-K slices whose partials are summed in
// fp32 round-to-nearest by the reduce kernel.  Within the cap the split is chosen for load balance.
constexpr int kMaxChainKB = 256;
int gemm_tc_auto_split(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ceil_div(M, TM) * ceil_div(N, TN);
  const int64_t num_kb = ceil_div(K, TK);
  const int s_min = (int)ceil_div(num_kb, kMaxChainKB);
  if (s_min <= 1 && (tiles >= kNumSMs || num_kb < 128)) return 1;
  int best = s_min;
  double best_cost = (double)ceil_div(tiles * s_min, kNumSMs) / s_min;
  for (int s = s_min + 1; s <= std::max(64, 2 * s_min) && num_kb / s >= 32; ++s) {
    const double cost = (double)ceil_div(tiles * s, kNumSMs) / s;
    if (cost < best_cost * 0.97) { best_cost = cost; best = s; }
  }
  return std::max(1, best);
}

static int64_t round16(int64_t b) { return (b + 15) / 16 * 16; }

int64_t gemm_tc_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K, int mode, int split_k) {
  // conservative: operands are assumed densely packed up to a 4-float padded leading dimension;
  // callers with larger leading dimensions get the FFMA fallback if the workspace is short
  if (split_k <= 0) split_k = gemm_tc_auto_split(M, N, K);
  int64_t b = 0;
  if (mode == GCG_GEMM_TF32X3 || mode == GCG_GEMM_TF32X3_CHAINED) {
    const int64_t a_el = (transA ? K : M) * (((transA ? M : K) + 3) / 4 * 4);
    const int64_t b_el = (transB ? N : K) * (((transB ? K : N) + 3) / 4 * 4);
    b += 2 * round16(a_el * 4) + 2 * round16(b_el * 4);     // hi and lo copies of both operands
  }
  if (split_k > 1) b += round16((int64_t)split_k * M * N * 4);
  return b + 64;
}

int gemm_tc_launch(const GemmArgs& g0, int transA, int transB, int mode, void* workspace, int64_t workspace_bytes,
                   cudaStream_t st) {
  if (!encode_fn()) return GCG_ERR_UNSUPPORTED;
  GemmArgs g = g0;
  if (!(g.vecA && g.vecB) || g.K < 1) return GCG_ERR_UNSUPPORTED;
  const int x3 = (mode == GCG_GEMM_TF32X3 || mode == GCG_GEMM_TF32X3_CHAINED);
  const int64_t a_rows = transA ? g.K : g.M, a_cols = transA ? g.M : g.K;
  const int64_t b_rows = transB ? g.N : g.K, b_cols = transB ? g.K : g.N;
  // carve the workspace
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int64_t off = 0;
  auto carve = [&](int64_t bytes) -> float* {
    float* p = reinterpret_cast<float*>(ws + off);
    off += round16(bytes);
    return p;
  };
  const float *Ah = g.A, *Bh = g.B;
  float *Al = nullptr, *Bl = nullptr, *Ahx = nullptr, *Bhx = nullptr;
  const bool split_a = x3 && !(g.A_hi && g.A_lo), split_b = x3 && !(g.B_hi && g.B_lo);
  if (split_a) {
    Al = carve(a_rows * g.lda * 4);
    Ahx = carve(a_rows * g.lda * 4);       // the round-to-nearest split needs its own hi copy
  }
  if (split_b) {
    Bl = carve(b_rows * g.ldb * 4);
    Bhx = carve(b_rows * g.ldb * 4);
  }
  int split = g.split_k;
  const int num_kb = (int)ceil_div(g.K, TK);
  split = std::max(1, std::min(split, num_kb));
  int kbps = (int)ceil_div(num_kb, split);
  split = (int)ceil_div(num_kb, kbps);
  if (split > 1) g.part = carve((int64_t)split * g.M * g.N * 4);
  g.split_k = split;
  if (off > workspace_bytes || (off > 0 && (!workspace || !aligned16(workspace)))) return GCG_ERR_UNSUPPORTED;

  if (x3) {
    if (split_a) {
      const int rc = tf32_split_launch(g.A, a_rows * g.lda, Ahx, Al, st);
      if (rc != GCG_OK) return rc;
      Ah = Ahx;
    } else {
      Ah = g.A_hi;
      Al = const_cast<float*>(g.A_lo);
    }
    if (split_b) {
      const int rc = tf32_split_launch(g.B, b_rows * g.ldb, Bhx, Bl, st);
      if (rc != GCG_OK) return rc;
      Bh = Bhx;
    } else {
      Bh = g.B_hi;
      Bl = const_cast<float*>(g.B_lo);
    }
  } else {
    // single-pass TF32: the tensor core TRUNCATES fp32 inputs to 10 mantissa bits (a bias of ~ -3.5e-4 per
    // operand); a caller that holds round-to-nearest hi copies (gcg_gemm_presplit_f32) gets unbiased inputs
    if (g.A_hi) Ah = g.A_hi;
    if (g.B_hi) Bh = g.B_hi;
  }
  // 2-CTA clusters share the B tile by multicast (see the kernel): worth it when there are enough M tiles to pair up
  // MEASURED (B200, profiles/r02_gemm_notes.md): the clustered schedule is bit-identical but SLOWER (3xTF32 140 vs
  // 148 TFLOP/s at 450k x 600 x 600, plain TF32 279 vs 343): the two CTAs advance in lockstep through a 3-stage ring
  // and stall on each other's slot releases, which costs more than the halved B traffic saves.  Default: off.
  const int cluster_env = getenv("GCG_GEMM_CLUSTER") ? atoi(getenv("GCG_GEMM_CLUSTER")) : 1;          // read per call:
  const int cluster_min = getenv("GCG_GEMM_CLUSTER_MIN_TILES") ? atoi(getenv("GCG_GEMM_CLUSTER_MIN_TILES")) : 64;   // tests flip them
  const int64_t m_tiles64 = ceil_div(g.M, TM);
  const int cs = (cluster_env >= 2 && m_tiles64 >= cluster_min) ? 2 : 1;
  CUtensorMap mAh, mAl, mBh, mBl;
  // MN-major operands load {32 MN, 32 K} boxes; a clustered K-major B is loaded in two {32 K, 64 N} halves
  const int a_box = transA ? 32 : TM, b_box = transB ? (cs == 2 ? TN / 2 : TN) : 32;
  const bool a_mn = transA != 0, b_mn = transB == 0;
  bool ok = make_map(&mAh, Ah, a_rows, a_cols, g.lda, a_box, a_mn) && make_map(&mBh, Bh, b_rows, b_cols, g.ldb, b_box, b_mn);
  if (x3) ok = ok && make_map(&mAl, Al, a_rows, a_cols, g.lda, a_box, a_mn) && make_map(&mBl, Bl, b_rows, b_cols, g.ldb, b_box, b_mn);
  else { mAl = mAh; mBl = mBh; }
  if (!ok) {
    set_error("gemm_tc_launch: cuTensorMapEncodeTiled failed");
    return GCG_ERR_CUDA;
  }
  TcArgs ta;
  ta.g = g;
  ta.m_tiles = (int)ceil_div(g.M, TM);
  ta.n_tiles = (int)ceil_div(g.N, TN);
  ta.splits = split;
  ta.num_kb_total = num_kb;
  ta.kb_per_split = kbps;
  ta.x3 = x3;
  // 3xTF32: accumulation chains of at most 4 K blocks (K = 128): 16 truncating adds into the hi.hi accumulator
  // instead of 75 at K = 600 (measured at Twitter-World: logits error 2.1x -> 1.0x -> below the 1e-4 bound)
  // measured in the Twitter-World epoch: chains of 4 K blocks everywhere cost +15 ms of 55 ms of GEMMs (the epilogue
  // warps drain TMEM once per chain), so only callers that ask for it (GCG_GEMM_TF32X3_CHAINED) get short chains
  const int chunk_env = getenv("GCG_GEMM_CHUNK_KB") ? atoi(getenv("GCG_GEMM_CHUNK_KB")) : 0;
  const int chunk = (mode == GCG_GEMM_TF32X3_CHAINED) ? 4 : chunk_env;
  ta.chunk_kb = (x3 && chunk > 0 && kbps > chunk) ? chunk : 0;
  const int64_t total = ceil_div(ta.m_tiles, cs) * ta.n_tiles * split;        // tile groups (one tile per CTA of a cluster)
  if (total * cs >= INT32_MAX) return GCG_ERR_UNSUPPORTED;
  const int smem_bytes = (x3 ? 3 * 4 : 6 * 2) * TILE_BYTES + 1024 + 256;
  const unsigned grid = (unsigned)(std::min<int64_t>(total, kNumSMs / cs) * cs);
  // A is MN-major when it is stored [K][M] (transA); B is MN-major when stored [K][N] (!transB)
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    if (cs == 1) {
      kern<<<grid, TC_THREADS, smem_bytes, st>>>(mAh, mAl, mBh, mBl, ta);
      return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, mAh, mAl, mBh, mBl, ta);
  };
  cudaError_t e;
  const bool ch = ta.chunk_kb > 0;
#define GCG_TC(AM, BM) \
  (cs == 1 ? (ch ? launch(gemm_tc_kernel<AM, BM, 1, true>) : launch(gemm_tc_kernel<AM, BM, 1, false>)) \
           : (ch ? launch(gemm_tc_kernel<AM, BM, 2, true>) : launch(gemm_tc_kernel<AM, BM, 2, false>)))
  if (!transA && !transB) e = GCG_TC(false, true);
  else if (transA && !transB) e = GCG_TC(true, true);
  else if (!transA && transB) e = GCG_TC(false, false);
  else e = GCG_TC(true, false);
#undef GCG_TC
  if (e != cudaSuccess) {
    set_error("gemm_tc_launch: %s", cudaGetErrorString(e));
    return GCG_ERR_CUDA;
  }
  count_launch();
  if (split > 1) return launch_splitk_reduce(g, st);
  return GCG_OK;
}

}  // namespace gcg

extern "C" int gcg_tf32_split_f32(const float* x, int64_t ld, int64_t n_rows, float* hi, float* lo, void* stream) {
  GCG_RECORD("gcg_tf32_split_f32", gcg_tf32_split_f32(x, ld, n_rows, hi, lo, s__));
  GCG_CHECK_ARG(x && hi && lo && n_rows >= 0, "gcg_tf32_split_f32: NULL argument");
  GCG_CHECK_SHAPE(ld % 4 == 0 && gcg::aligned16(x) && gcg::aligned16(hi) && gcg::aligned16(lo),
                  "gcg_tf32_split_f32: needs 16-byte aligned operands and ld %% 4 == 0");
  return gcg::tf32_split_launch(x, n_rows * ld, hi, lo, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gcg_gemm_tc_available(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return (major == 10 && gcg::encode_fn() != nullptr) ? 1 : 0;
}
