// 2-CTA variant of gemm.cuh's kernel: a thread-block cluster of two CTAs (one TPC) computes a 256 x BN output tile
// with tcgen05.mma.cta_group::2 (UMMA M = 256).  Each CTA stages its own 128 A rows and HALF of the B tile, so the
// operand bytes that cross L2 -> shared memory per FLOP drop from (1/128 + 1/BN) to (1/256 + 1/BN) ... the 1-CTA
// kernel pulls ~0.7 GB through L2 per ViT GEMM and tops out near 1000 TFLOP/s (profiles/prof_gemm_r01.md).
//
// Protocol (leader = cluster rank 0 issues all MMAs):
//   * both CTAs' producers TMA-load into their own shared memory with .cta_group::2, completing bytes on the LEADER's
//     full barrier (address with the peer bit cleared); the leader arms expect_tx for both CTAs' bytes, the peer adds
//     a plain remote arrive (full barrier count 2);
//   * tcgen05.commit.cta_group::2 ... multicast::cluster signals `empty` / `tmem full` in both CTAs;
//   * each CTA's 8 epilogue warps drain their own 128 TMEM lanes; all 16 warps arrive on the leader's `tmem empty`.
// Normal (non-transposed) epilogues only; the epilogue body is gemm.cuh's epi_store_normal().
#pragma once
#include "gemm.cuh"

namespace gitb200 {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into this CTA's smem, transaction bytes counted on the leader CTA's barrier at the same offset
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// arrive (count 1, no tx) on the leader CTA's barrier at the same offset as `bar`
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {  // arrives on `bar`'s offset in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN>
struct Gemm2Cfg {
  static constexpr int BM = 128;                 // rows per CTA (256 per pair)
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;   // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 128 + EPI_WARPS * 32;
  static constexpr int STAGING_BYTES = EPI_WARPS * 32 * 128;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_LIMIT = 227 * 1024;
  static constexpr int STAGES_RAW = (SMEM_LIMIT - 1024 - BAR_BYTES - STAGING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + BAR_BYTES;
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "BN");
  static_assert(B_BYTES % 1024 == 0, "B half tile must keep 1024B alignment for SWIZZLE_128B");
  static_assert((2 * STAGES + 4) * 8 + 8 <= BAR_BYTES, "barrier area");
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg<BN>::THREADS, 1)
gemm2_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmParams p) {
  using C = Gemm2Cfg<BN>;
  static_assert((EPI & (EPI_TRANSPOSED | EPI_PARTIAL)) == 0, "2-CTA kernel: normal epilogues only");
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // keep the shared-memory provenance of the pointer (plain pointer arithmetic) so staging accesses compile to LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  uint8_t* sStage = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sStage + C::STAGING_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 2);    // leader: arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty[s], 1);   // multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 2 * C::EPI_WARPS);   // both CTAs' epilogue warps (meaningful in the leader)
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // barriers of both CTAs are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (p.M + 2 * C::BM - 1) / (2 * C::BM);   // 256-row tiles per pair
  const int n_tiles = (p.N + BN - 1) / BN;
  const int kb_total = (p.K + C::BK - 1) / C::BK;
  const int kb_per = (kb_total + p.k_splits - 1) / p.k_splits;
  const int mn_tiles = m_tiles * n_tiles;
  const int num_tiles = mn_tiles * p.k_splits;

  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int split = tile / mn_tiles;
        const int rem = tile - split * mn_tiles;
        const int m_blk = rem / n_tiles;
        const int n_blk = rem - m_blk * n_tiles;
        const int kb0 = split * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C::STAGE_BYTES);   // both CTAs' bytes land on this barrier
          tma_load_2d_2cta(sA + stage * C::A_BYTES, &tmA, &full[stage], kb * C::BK, m_blk * 2 * C::BM + static_cast<int>(rank) * C::BM);
          tma_load_2d_2cta(sB + stage * C::B_BYTES, &tmB, &full[stage], kb * C::BK, n_blk * BN + static_cast<int>(rank) * (BN / 2));
          if (!leader) mbar_arrive_leader(&full[stage]);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ UMMA issuer (leader CTA) ----------------
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * C::BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int accum = 0;
      uint32_t accum_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int split = tile / mn_tiles;
        const int kb0 = split * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        mbar_wait(&tempty[accum], accum_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + accum * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * C::A_BYTES);
          const uint32_t b_base = smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < C::BK / 16; ++k) {
            umma_bf16_2cta(d_tmem, umma_desc_sw128(a_base + k * 32), umma_desc_sw128(b_base + k * 32), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2cta(&empty[stage]);  // frees the smem slot in both CTAs once these MMAs have read it
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta(&tfull[accum]);  // accumulators complete in both CTAs -> epilogues
        accum ^= 1;
        if (accum == 0) accum_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ----------------------------------
    // 8 warps: warp w reads TMEM lanes 32*(w%4).. (hardware restriction) and the 32-column chunks
    // c == (w-4)/4 (mod 2).  Every 32x32 fp32 chunk is transposed through a per-warp swizzled staging buffer so
    // that each lane ends up with 4 consecutive output elements of one row: coalesced 128-bit accesses.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint8_t* stg = sStage + (warp - 4) * (32 * 128);
    const int rsub = lane >> 3;
    int accum = 0;
    uint32_t accum_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int rem = tile % mn_tiles;
      const int m_blk = rem / n_tiles;
      const int n_blk = rem - m_blk * n_tiles;
      const int row0 = m_blk * 2 * C::BM + static_cast<int>(rank) * C::BM + q * 32;  // first tile row of this warp
      long long ooff[8];
      const uint32_t okmask = epi_row_offsets(p, row0, rsub, true, ooff);
      mbar_wait(&tfull[accum], accum_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int n0 = n_blk * BN + c * 32;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + accum * BN + c * 32, r);
        tmem_ld_wait();
        epi_store_normal<EPI>(p, r, stg, lane, n0, row0, ooff, okmask);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[accum]);   // leader's barrier (local for rank 0)
      accum ^= 1;
      if (accum == 0) accum_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's shared memory / TMEM stay valid until the leader's last MMA has retired
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace gitb200
