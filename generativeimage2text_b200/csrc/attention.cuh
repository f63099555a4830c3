// Multi-head attention kernels (head dim 64).
//
// flash_attn_kernel : non-causal softmax(q k^T / 8) v over S keys for every (batch, head); used by the
//   ViT blocks (reference layers/CLIP/model.py:189-197 -> nn.MultiheadAttention -> SDPA, no mask) and by
//   the one-off image-row pass of the decoder (image rows attend image rows only, reference
//   layers/decoder.py:119-120; layers/bert/modeling_bert.py:41-47,138-152).  Online-softmax over 64-key
//   chunks, cp.async double buffering, warp-level bf16 tensor-core MMAs with fp32 accumulation.
//
// decode_attn_kernel : one new text row per sequence against [image K/V || text K/V] (the KV-cached form of
//   reference layers/decoder.py:121-123 + modeling_bert.py:124-152).  Pure HBM streaming: every K/V row is
//   read once with 128-bit loads; image K/V are shared by the beams of an image; the new token's K/V are
//   appended to the text cache by the same kernel.
#pragma once
#include "ptx.cuh"
#include "rowops.cuh"

namespace gitb200 {

struct AttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* out;
  int B, S, H;
  long long q_rs, kv_rs, q_bs, kv_bs, o_rs, o_bs;  // row / batch strides in elements
  float scale_log2;                                 // (1/sqrt(64)) * log2(e)
};

__device__ __forceinline__ void attn_load_tile(uint32_t smem_base, const __nv_bfloat16* g, long long row_stride,
                                               int row0, int nrows, int S, int tid, int nthreads) {
  for (int idx = tid; idx < nrows * 8; idx += nthreads) {
    const int r = idx >> 3;
    const int c = idx & 7;
    const bool valid = (row0 + r) < S;
    const int gr = valid ? (row0 + r) : (S - 1);
    const __nv_bfloat16* src = g + static_cast<long long>(gr) * row_stride + c * 8;
    cp_async_16(smem_base + r * 128 + ((c ^ (r & 7)) << 4), src, valid);
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) flash_attn_kernel(const AttnParams p) {
  constexpr int QROWS = NW * 16;
  constexpr int KC = 64;
  __shared__ __align__(128) uint8_t sQ[QROWS * 128];
  __shared__ __align__(128) uint8_t sK[2][KC * 128];
  __shared__ __align__(128) uint8_t sV[2][KC * 128];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q_row0 = blockIdx.x * QROWS;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const __nv_bfloat16* qg = p.q + b * p.q_bs + h * 64;
  const __nv_bfloat16* kg = p.k + b * p.kv_bs + h * 64;
  const __nv_bfloat16* vg = p.v + b * p.kv_bs + h * 64;
  const int nchunks = (p.S + KC - 1) / KC;

  attn_load_tile(smem_u32(sQ), qg, p.q_rs, q_row0, QROWS, p.S, tid, NW * 32);
  attn_load_tile(smem_u32(sK[0]), kg, p.kv_rs, 0, KC, p.S, tid, NW * 32);
  attn_load_tile(smem_u32(sV[0]), vg, p.kv_rs, 0, KC, p.S, tid, NW * 32);
  cp_async_commit();

  uint32_t qa[4][4];
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) {
      attn_load_tile(smem_u32(sK[buf ^ 1]), kg, p.kv_rs, (c + 1) * KC, KC, p.S, tid, NW * 32);
      attn_load_tile(smem_u32(sV[buf ^ 1]), vg, p.kv_rs, (c + 1) * KC, KC, p.S, tid, NW * 32);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (c == 0) {
      const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int chunk = 2 * kk + (lane >> 4);
        ldmatrix_x4(qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], smem_u32(sQ) + row * 128 + ((chunk ^ (row & 7)) << 4));
      }
    }
    // ---- S = Q K^T for this chunk -------------------------------------------------------------
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      const int krow = 8 * j + (lane & 7);
#pragma unroll
      for (int kk2 = 0; kk2 < 2; ++kk2) {
        const int chunk = 4 * kk2 + (lane >> 3);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, smem_u32(sK[buf]) + krow * 128 + ((chunk ^ (krow & 7)) << 4));
        mma_bf16_16816(s[j], qa[2 * kk2], b0, b1);
        mma_bf16_16816(s[j], qa[2 * kk2 + 1], b2, b3);
      }
    }
    // ---- mask the tail, online softmax ---------------------------------------------------------
    const int key0 = c * KC + 2 * (lane & 3);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = key0 + 8 * j;
      if (key >= p.S) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (key + 1 >= p.S) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
    float corr[2], m_new[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      m_new[i] = fmaxf(m_run[i], mx[i]);
      corr[i] = exp2f((m_run[i] - m_new[i]) * p.scale_log2);
      m_run[i] = m_new[i];
      l_run[i] *= corr[i];
    }
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f((s[j][0] - m_new[0]) * p.scale_log2);
      const float p1 = exp2f((s[j][1] - m_new[0]) * p.scale_log2);
      const float p2 = exp2f((s[j][2] - m_new[1]) * p.scale_log2);
      const float p3 = exp2f((s[j][3] - m_new[1]) * p.scale_log2);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
      o[j][0] *= corr[0];
      o[j][1] *= corr[0];
      o[j][2] *= corr[1];
      o[j][3] *= corr[1];
    }
    // ---- O += P V -------------------------------------------------------------------------------
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int vrow = 16 * kk + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int chunk = 2 * jj + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(b0, b1, b2, b3, smem_u32(sV[buf]) + vrow * 128 + ((chunk ^ (vrow & 7)) << 4));
        mma_bf16_16816(o[2 * jj], pa[kk], b0, b1);
        mma_bf16_16816(o[2 * jj + 1], pa[kk], b2, b3);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 1);
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 2);
  }
  const float inv0 = 1.0f / l_run[0];
  const float inv1 = 1.0f / l_run[1];
  const int r0 = q_row0 + warp * 16 + (lane >> 2);
  const int r1 = r0 + 8;
  __nv_bfloat16* og = p.out + b * p.o_bs + h * 64 + 2 * (lane & 3);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (r0 < p.S)
      *reinterpret_cast<uint32_t*>(og + static_cast<long long>(r0) * p.o_rs + 8 * j) = pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
    if (r1 < p.S)
      *reinterpret_cast<uint32_t*>(og + static_cast<long long>(r1) * p.o_rs + 8 * j) = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
  }
}

// ------------------------------------------------------------------------------------------------
// flash_attn_tc_kernel: the same non-causal attention on the 5th-generation tensor cores, for sequences that fit the
// tensor memory in one piece (S <= 512 keys: the ViT blocks and the image-row prefill of the image models).
//   One CTA per (batch, head).  K and V of the head are fetched ONCE by TMA (128B-swizzled boxes) and stay in shared
//   memory; per 128-row query tile:  TMA Q -> tcgen05.mma S = Q K^T (fp32, TMEM columns [0, Spad)) -> the four softmax
//   warps read their TMEM lane (= query row, so row max / row sum need no shuffles), write P = exp2(...) as bf16 into
//   shared memory in the K-major swizzled operand layout -> tcgen05.mma O = P V with V as an MN-major B operand (V stays
//   [key][dim] as TMA delivered it), O overwriting TMEM columns [0, 64) -> the softmax warps scale by 1 / row sum and
//   store bf16 rows.
//   Warp 0: TMA + MMA issue (one lane); warps 1-8: softmax / epilogue, two per TMEM lane quadrant (they split the key
//   columns).  Q and K share shared memory with P (K is re-fetched per query tile -- an L2 hit), which keeps a CTA at
//   ~92 KB for S = 197 so that TWO CTAs per SM overlap each other's load / MMA / softmax latencies.
// ------------------------------------------------------------------------------------------------
struct AttnTcParams {
  __nv_bfloat16* out;
  int B, S, H;
  int spad;                 // keys padded to a multiple of 16 (MMA N / K granularity)
  int kv_box_rows, kv_boxes;   // K / V arrive as kv_boxes TMA boxes of kv_box_rows rows (kv_box_rows * kv_boxes >= spad)
  long long q_rows_per_batch, kv_rows_per_batch;   // row coordinate of batch b in the tensor maps = b * rows_per_batch
  int q_col0, k_col0, v_col0;                      // column of head 0 inside the tensor maps (heads are 64 columns apart)
  long long o_rs, o_bs;
  float scale_log2;
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

constexpr int kAttnTcThreads = 288;      // warp 0: TMA + MMA issue; warps 1-8: softmax / epilogue, two per TMEM lane quadrant
__host__ __device__ inline size_t attn_tc_kv_bytes(int kv_box_rows, int kv_boxes) {
  return (static_cast<size_t>(kv_box_rows) * kv_boxes * 128 + 1023) / 1024 * 1024;
}
// shared memory: V | [ Q tile | K ] -- the P blocks (128 rows x 64 keys each) overlay Q and K, which are dead once S sits in TMEM
__host__ __device__ inline size_t attn_tc_p_bytes(int spad, int kv_box_rows, int kv_boxes) {
  const size_t p = static_cast<size_t>((spad + 63) / 64) * 16384;
  const size_t qk = 16384 + attn_tc_kv_bytes(kv_box_rows, kv_boxes);
  return p > qk ? p : qk;
}
__host__ __device__ inline size_t attn_tc_smem_bytes(int spad, int kv_box_rows, int kv_boxes) {
  return 1024 + attn_tc_kv_bytes(kv_box_rows, kv_boxes) + attn_tc_p_bytes(spad, kv_box_rows, kv_boxes) + 64 + 2048;
}

// one MUFU.EX2 (exp2f() adds a denormal-range fix-up the softmax does not need: those terms vanish against a row sum >= 1)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// bounded wait: a protocol bug must show up as wrong numbers in a test, never as a hung device
__device__ __forceinline__ void mbar_wait_lim(uint64_t* bar, uint32_t parity) {
  for (unsigned int i = 0; i < (1u << 22); ++i)
    if (mbar_try_wait(bar, parity)) return;
}

__global__ void __launch_bounds__(kAttnTcThreads, 1)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t attn_tc_raw[];
  uint8_t* smem = attn_tc_raw + ((1024u - (smem_u32(attn_tc_raw) & 1023u)) & 1023u);
  const size_t kv_bytes = attn_tc_kv_bytes(p.kv_box_rows, p.kv_boxes);
  uint8_t* sV = smem;
  uint8_t* sP = smem + kv_bytes;                     // P blocks; block 0 doubles as the Q tile ...
  uint8_t* sK = sP + 16384;                          // ... and K (re-fetched per query tile, an L2 hit) sits in blocks 1..
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + attn_tc_p_bytes(p.spad, p.kv_box_rows, p.kv_boxes));
  uint64_t* bar_v = bars;        // V landed (once)
  uint64_t* bar_q = bars + 1;    // Q tile + K landed
  uint64_t* bar_s = bars + 2;    // S complete in TMEM
  uint64_t* bar_p = bars + 3;    // P written (8 warps)
  uint64_t* bar_o = bars + 4;    // O complete in TMEM
  uint64_t* bar_free = bars + 5; // the softmax warps are done with O (8 warps): TMEM and the Q / K / P region may be reused
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* row_sums = reinterpret_cast<float*>(bars + 8);     // [2 halves][128 rows]
  float* row_max = row_sums + 256;                          // [2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int n_tiles = (p.S + 127) / 128;
  const uint32_t tmem_cols = (p.spad <= 256) ? 256u : 512u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_v, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 8);
    mbar_init(bar_o, 1);
    mbar_init(bar_free, 8);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t kbytes = static_cast<uint32_t>(p.kv_boxes * p.kv_box_rows * 128);
      mbar_arrive_expect_tx(bar_v, kbytes);
      for (int i = 0; i < p.kv_boxes; ++i)
        tma_load_2d(sV + static_cast<size_t>(i) * p.kv_box_rows * 128, &tmV, bar_v, p.v_col0 + h * 64,
                    static_cast<int>(b * p.kv_rows_per_batch) + i * p.kv_box_rows);
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);     // B operand (V) is MN-major
      for (int tile = 0; tile < n_tiles; ++tile) {
        const uint32_t ph = tile & 1;
        if (tile > 0) { mbar_wait_lim(bar_free, ph ^ 1); tc_fence_after(); }
        mbar_arrive_expect_tx(bar_q, 128 * 128 + kbytes);
        tma_load_2d(sP, &tmQ, bar_q, p.q_col0 + h * 64, static_cast<int>(b * p.q_rows_per_batch) + tile * 128);
        for (int i = 0; i < p.kv_boxes; ++i)
          tma_load_2d(sK + static_cast<size_t>(i) * p.kv_box_rows * 128, &tmK, bar_q, p.k_col0 + h * 64,
                      static_cast<int>(b * p.kv_rows_per_batch) + i * p.kv_box_rows);
        mbar_wait_lim(bar_q, ph);
        tc_fence_after();
        // S = Q K^T: N in pieces of <= 256 key columns, K = 64 head dims = 4 MMAs each
        for (int n0 = 0; n0 < p.spad; n0 += 256) {
          const int nn = min(256, p.spad - n0);
          const uint32_t idesc = umma_idesc_bf16(128, nn);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + n0, umma_desc_sw128(smem_u32(sP) + k * 32), umma_desc_sw128(smem_u32(sK) + n0 * 128 + k * 32), idesc,
                      k > 0 ? 1u : 0u);
        }
        umma_commit(bar_s);
        // O = P V once the softmax warps have written P
        if (tile == 0) mbar_wait_lim(bar_v, 0);
        mbar_wait_lim(bar_p, ph);
        tc_fence_after();
        const int nk16 = p.spad / 16;
        for (int kb = 0; kb < nk16; ++kb)
          umma_bf16(tmem_base, umma_desc_sw128(smem_u32(sP) + (kb >> 2) * 16384 + (kb & 3) * 32),
                    umma_desc_sw128(smem_u32(sV) + kb * 2048), idesc_pv, kb > 0 ? 1u : 0u);
        umma_commit(bar_o);
      }
    }
  } else {
    // ---- softmax / epilogue: thread = query row = TMEM lane; the two warps of a quadrant split the key columns ----
    const int quad = warp & 3;
    const int half = (warp - 1) >> 2;
    const int row = quad * 32 + lane;                 // row of the tile
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int n16 = p.spad / 16;
    for (int tile = 0; tile < n_tiles; ++tile) {
      const uint32_t ph = tile & 1;
      mbar_wait_lim(bar_s, ph);
      tc_fence_after();
      // pass 1: the row maximum.  Each warp of a quadrant scans its own 16-column chunks (c = half, half + 2, ...), two
      // TMEM loads in flight; the two partial maxima meet through shared memory.  Chunks below n_full hold no padding
      // column, so the bulk of the row runs without per-column predicates.
      const int n_full = p.S >> 4;
      auto chunk_max = [&](const uint32_t (&r)[16], int c, float mx_in) -> float {
        float m0, m1;
        if (c < n_full) {
          m0 = fmaxf(fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3])));
          m1 = fmaxf(fmaxf(__uint_as_float(r[4]), __uint_as_float(r[5])), fmaxf(__uint_as_float(r[6]), __uint_as_float(r[7])));
          m0 = fmaxf(m0, fmaxf(fmaxf(__uint_as_float(r[8]), __uint_as_float(r[9])), fmaxf(__uint_as_float(r[10]), __uint_as_float(r[11]))));
          m1 = fmaxf(m1, fmaxf(fmaxf(__uint_as_float(r[12]), __uint_as_float(r[13])), fmaxf(__uint_as_float(r[14]), __uint_as_float(r[15]))));
        } else {
          m0 = m1 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c * 16 + j < p.S) m0 = fmaxf(m0, __uint_as_float(r[j]));
        }
        return fmaxf(mx_in, fmaxf(m0, m1));
      };
      float mx = -INFINITY;
      {
        int c = half;
        for (; c + 2 < n16; c += 4) {
          uint32_t ra[16], rb[16];
          tmem_ld_32x32b_x16(t_lane + c * 16, ra);
          tmem_ld_32x32b_x16(t_lane + (c + 2) * 16, rb);
          tmem_ld_wait();
          mx = chunk_max(ra, c, mx);
          mx = chunk_max(rb, c + 2, mx);
        }
        for (; c < n16; c += 2) {
          uint32_t ra[16];
          tmem_ld_32x32b_x16(t_lane + c * 16, ra);
          tmem_ld_wait();
          mx = chunk_max(ra, c, mx);
        }
      }
      row_max[half * 128 + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      mx = fmaxf(mx, row_max[(half ^ 1) * 128 + row]);
      // pass 2: this warp's 16-column chunks -> P (bf16, K-major swizzled operand layout), partial row sum
      float sum = 0.f;
      const float mb = mx * p.scale_log2;
      auto chunk_p = [&](const uint32_t (&r)[16], int c) {
        uint32_t pk[8];
        float s0 = 0.f, s1 = 0.f;
        if (c < n_full) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), p.scale_log2, -mb));
            const float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2, -mb));
            s0 += p0;
            s1 += p1;
            pk[j] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float p0 = (c * 16 + 2 * j < p.S) ? ex2_approx(fmaf(__uint_as_float(r[2 * j]), p.scale_log2, -mb)) : 0.f;
            const float p1 = (c * 16 + 2 * j + 1 < p.S) ? ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2, -mb)) : 0.f;
            s0 += p0;
            s1 += p1;
            pk[j] = pack_bf16(p0, p1);
          }
        }
        sum += s0 + s1;
        uint8_t* blk = sP + (c >> 2) * 16384 + row * 128;
        const int ch = 2 * (c & 3);
        *reinterpret_cast<uint4*>(blk + (((ch) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(blk + (((ch + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      };
      {
        int c = half;
        for (; c + 2 < n16; c += 4) {
          uint32_t ra[16], rb[16];
          tmem_ld_32x32b_x16(t_lane + c * 16, ra);
          tmem_ld_32x32b_x16(t_lane + (c + 2) * 16, rb);
          tmem_ld_wait();
          chunk_p(ra, c);
          chunk_p(rb, c + 2);
        }
        for (; c < n16; c += 2) {
          uint32_t ra[16];
          tmem_ld_32x32b_x16(t_lane + c * 16, ra);
          tmem_ld_wait();
          chunk_p(ra, c);
        }
      }
      row_sums[half * 128 + row] = sum;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");   // the partner warp's partial row sum
      const float inv = 1.0f / (row_sums[row] + row_sums[128 + row]);
      // ---- O / row sum -> bf16 row (each warp: two of the four 16-column chunks) ----
      mbar_wait_lim(bar_o, ph);
      tc_fence_after();
      const int qrow = tile * 128 + row;
      __nv_bfloat16* orow = p.out + b * p.o_bs + static_cast<long long>(qrow) * p.o_rs + h * 64;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * cc + half;
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_lane + c * 16, r);
        tmem_ld_wait();
        if (qrow < p.S) {
          uint4 v0, v1;
          v0.x = pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
          v0.y = pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
          v0.z = pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
          v0.w = pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
          v1.x = pack_bf16(__uint_as_float(r[8]) * inv, __uint_as_float(r[9]) * inv);
          v1.y = pack_bf16(__uint_as_float(r[10]) * inv, __uint_as_float(r[11]) * inv);
          v1.z = pack_bf16(__uint_as_float(r[12]) * inv, __uint_as_float(r[13]) * inv);
          v1.w = pack_bf16(__uint_as_float(r[14]) * inv, __uint_as_float(r[15]) * inv);
          reinterpret_cast<uint4*>(orow + c * 16)[0] = v0;
          reinterpret_cast<uint4*>(orow + c * 16)[1] = v1;
        }
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");   // row_sums are rewritten by the next tile
      if (lane == 0) mbar_arrive(bar_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// flash_attn_tc_long_kernel: the same attention on tcgen05 for sequences that do NOT fit the tensor memory in one piece
// (S > 512: the 6-frame video prefill with 1182 keys, 30 x 40 VQA grids with 1201).
//   One CTA per (head, batch, 128-row query tile), two CTAs per SM.  The keys are walked in blocks of 128, twice:
//     pass 0:  S_blk = Q K_blk^T (tcgen05.mma into TMEM columns [0, 128))  ->  the softmax warps fold the block into the row maximum;
//     pass 1:  S_blk again  ->  P_blk = exp2((S_blk - max) * scale) as bf16 in the K-major operand layout  ->
//              O += P_blk V_blk (TMEM columns [128, 192), accumulated by the tensor core across the blocks).
//   Recomputing S costs a second pass of QK^T MMAs -- the tensor pipe idles below 15 % in these kernels, the softmax warps'
//   instruction issue is the limit -- and buys a softmax without running-maximum corrections of O (which would be a TMEM
//   load / scale / store round trip per block).  K blocks are double buffered (the next block's TMA overlaps the current
//   block's MMA and softmax), V and P single buffered; 192 of 256 allocated TMEM columns, 99 KB of shared memory.
//   Warp 0: TMA + MMA issue (one lane); warps 1-8: softmax / epilogue, two per TMEM lane quadrant, splitting the 16-column
//   chunks of a block like flash_attn_tc_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnLongBlk = 128;        // keys per block
__host__ __device__ inline size_t attn_tc_long_smem_bytes() {
  return 1024 + 16384 /*Q*/ + 2 * 16384 /*K*/ + 16384 /*V*/ + 32768 /*P*/ + 128 /*barriers*/ + 2048 /*row max / row sum*/;
}

__global__ void __launch_bounds__(kAttnTcThreads, 2)
flash_attn_tc_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t attn_tcl_raw[];
  uint8_t* smem = attn_tcl_raw + ((1024u - (smem_u32(attn_tcl_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;                  // two buffers of 16 KB
  uint8_t* sV = smem + 3 * 16384;
  uint8_t* sP = smem + 4 * 16384;              // two sub-blocks [128 rows x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * 16384);
  uint64_t* bar_q = bars;          // Q tile landed (once)
  uint64_t* bar_k = bars + 1;      // [2] K block landed
  uint64_t* bar_v = bars + 3;      // V block landed
  uint64_t* bar_s = bars + 4;      // S block complete in TMEM
  uint64_t* bar_d = bars + 5;      // the 8 softmax warps are done with the S block (pass 1: and have written P)
  uint64_t* bar_o = bars + 6;      // P V of the block complete: P and V shared memory may be rewritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* row_sums = reinterpret_cast<float*>(bars + 16);    // [2 halves][128 rows]
  float* row_max = row_sums + 256;                          // [2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y, tile = blockIdx.z;
  const int nblk = (p.S + kAttnLongBlk - 1) / kAttnLongBlk;
  const int n_it = 2 * nblk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1);
    mbar_init(&bar_k[0], 1);
    mbar_init(&bar_k[1], 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_d, 8);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      const int kv_row0 = static_cast<int>(b * p.kv_rows_per_batch);
      const uint32_t idesc_s = umma_idesc_bf16(128, kAttnLongBlk);
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);     // B operand (V) is MN-major
      mbar_arrive_expect_tx(bar_q, 16384);
      tma_load_2d(sQ, &tmQ, bar_q, p.q_col0 + h * 64, static_cast<int>(b * p.q_rows_per_batch) + tile * 128);
      mbar_arrive_expect_tx(&bar_k[0], 16384);
      tma_load_2d(sK, &tmK, &bar_k[0], p.k_col0 + h * 64, kv_row0);
      mbar_wait_lim(bar_q, 0);
      for (int it = 0; it < n_it; ++it) {
        const int pass = it >= nblk ? 1 : 0;
        const int blk = it - pass * nblk;
        const int buf = it & 1;
        // the next iteration's K block (its buffer was last read by the S MMA of iteration it - 1, whose commit we waited for)
        if (it + 1 < n_it) {
          const int nb = (it + 1) - ((it + 1) >= nblk ? nblk : 0);
          mbar_arrive_expect_tx(&bar_k[buf ^ 1], 16384);
          tma_load_2d(sK + (buf ^ 1) * 16384, &tmK, &bar_k[buf ^ 1], p.k_col0 + h * 64, kv_row0 + nb * kAttnLongBlk);
        }
        if (pass == 1) {
          if (blk > 0) mbar_wait_lim(bar_o, (blk - 1) & 1);       // the previous P V has finished reading V (and P)
          mbar_arrive_expect_tx(bar_v, 16384);
          tma_load_2d(sV, &tmV, bar_v, p.v_col0 + h * 64, kv_row0 + blk * kAttnLongBlk);
        }
        mbar_wait_lim(&bar_k[buf], (it >> 1) & 1);
        if (it > 0) mbar_wait_lim(bar_d, (it - 1) & 1);           // the softmax warps have read the previous S block
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, umma_desc_sw128(smem_u32(sQ) + k * 32), umma_desc_sw128(smem_u32(sK) + buf * 16384 + k * 32), idesc_s,
                    k > 0 ? 1u : 0u);
        umma_commit(bar_s);
        if (pass == 1) {
          mbar_wait_lim(bar_v, blk & 1);
          mbar_wait_lim(bar_d, it & 1);                           // P of this block written
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 8; ++kb)
            umma_bf16(tmem_o, umma_desc_sw128(smem_u32(sP) + (kb >> 2) * 16384 + (kb & 3) * 32),
                      umma_desc_sw128(smem_u32(sV) + kb * 2048), idesc_pv, (blk > 0 || kb > 0) ? 1u : 0u);
          umma_commit(bar_o);
        } else {
          mbar_wait_lim(bar_s, it & 1);                           // (so that the K buffer is free for the prefetch above)
        }
      }
    }
  } else {
    // ---- softmax / epilogue: thread = query row = TMEM lane; the two warps of a quadrant split the 16-column chunks ----
    const int quad = warp & 3;
    const int half = (warp - 1) >> 2;
    const int row = quad * 32 + lane;                 // row of the tile
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    // ---- pass 0: the row maximum over all key blocks ----
    float mx = -INFINITY;
    for (int blk = 0; blk < nblk; ++blk) {
      mbar_wait_lim(bar_s, blk & 1);
      tc_fence_after();
      const int valid = min(kAttnLongBlk, p.S - blk * kAttnLongBlk);      // keys of this block that exist
      const int n_full = valid >> 4;
      for (int c = half; c < 8; c += 4) {
        uint32_t ra[16], rb[16];
        tmem_ld_32x32b_x16(t_lane + c * 16, ra);
        tmem_ld_32x32b_x16(t_lane + (c + 2) * 16, rb);
        tmem_ld_wait();
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint32_t (&r)[16] = w ? rb : ra;
          const int cc = c + 2 * w;
          if (cc < n_full) {
            float m0 = fmaxf(fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3])));
            float m1 = fmaxf(fmaxf(__uint_as_float(r[4]), __uint_as_float(r[5])), fmaxf(__uint_as_float(r[6]), __uint_as_float(r[7])));
            m0 = fmaxf(m0, fmaxf(fmaxf(__uint_as_float(r[8]), __uint_as_float(r[9])), fmaxf(__uint_as_float(r[10]), __uint_as_float(r[11]))));
            m1 = fmaxf(m1, fmaxf(fmaxf(__uint_as_float(r[12]), __uint_as_float(r[13])), fmaxf(__uint_as_float(r[14]), __uint_as_float(r[15]))));
            mx = fmaxf(mx, fmaxf(m0, m1));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (cc * 16 + j < valid) mx = fmaxf(mx, __uint_as_float(r[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_d);
    }
    row_max[half * 128 + row] = mx;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
    mx = fmaxf(mx, row_max[(half ^ 1) * 128 + row]);
    // ---- pass 1: P blocks and the partial row sum ----
    float sum = 0.f;
    const float mb = mx * p.scale_log2;
    for (int blk = 0; blk < nblk; ++blk) {
      const int it = nblk + blk;
      mbar_wait_lim(bar_s, it & 1);
      if (blk > 0) mbar_wait_lim(bar_o, (blk - 1) & 1);           // the previous P V has finished reading P
      tc_fence_after();
      const int valid = min(kAttnLongBlk, p.S - blk * kAttnLongBlk);
      const int n_full = valid >> 4;
      for (int c = half; c < 8; c += 4) {
        uint32_t ra[16], rb[16];
        tmem_ld_32x32b_x16(t_lane + c * 16, ra);
        tmem_ld_32x32b_x16(t_lane + (c + 2) * 16, rb);
        tmem_ld_wait();
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint32_t (&r)[16] = w ? rb : ra;
          const int cc = c + 2 * w;
          uint32_t pk[8];
          float s0 = 0.f, s1 = 0.f;
          if (cc < n_full) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), p.scale_log2, -mb));
              const float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2, -mb));
              s0 += p0;
              s1 += p1;
              pk[j] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float p0 = (cc * 16 + 2 * j < valid) ? ex2_approx(fmaf(__uint_as_float(r[2 * j]), p.scale_log2, -mb)) : 0.f;
              const float p1 = (cc * 16 + 2 * j + 1 < valid) ? ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), p.scale_log2, -mb)) : 0.f;
              s0 += p0;
              s1 += p1;
              pk[j] = pack_bf16(p0, p1);
            }
          }
          sum += s0 + s1;
          uint8_t* pb = sP + (cc >> 2) * 16384 + row * 128;
          const int ch = 2 * (cc & 3);
          *reinterpret_cast<uint4*>(pb + (((ch) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(pb + (((ch + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_d);
    }
    row_sums[half * 128 + row] = sum;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
    const float inv = 1.0f / (row_sums[row] + row_sums[128 + row]);
    // ---- O / row sum -> bf16 row (each warp: two of the four 16-column chunks) ----
    mbar_wait_lim(bar_o, (nblk - 1) & 1);
    tc_fence_after();
    const int qrow = tile * 128 + row;
    __nv_bfloat16* orow = p.out + b * p.o_bs + static_cast<long long>(qrow) * p.o_rs + h * 64;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = 2 * cc + half;
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_lane + 128 + c * 16, r);
      tmem_ld_wait();
      if (qrow < p.S) {
        uint4 v0, v1;
        v0.x = pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
        v0.y = pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
        v0.z = pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
        v0.w = pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
        v1.x = pack_bf16(__uint_as_float(r[8]) * inv, __uint_as_float(r[9]) * inv);
        v1.y = pack_bf16(__uint_as_float(r[10]) * inv, __uint_as_float(r[11]) * inv);
        v1.z = pack_bf16(__uint_as_float(r[12]) * inv, __uint_as_float(r[13]) * inv);
        v1.w = pack_bf16(__uint_as_float(r[14]) * inv, __uint_as_float(r[15]) * inv);
        reinterpret_cast<uint4*>(orow + c * 16)[0] = v0;
        reinterpret_cast<uint4*>(orow + c * 16)[1] = v1;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------
struct DecAttnParams {
  const float* qkv;                // [n_partials][R, 3*D] fp32 (q | k | v): the QKV GEMM's split-K partial sums, added
  int n_partials;                  //   here in split order (1..4 buffers, partial_stride elements apart)
  long long partial_stride;
  const float* bqkv;               // [3*D] bias, added here
  const __nv_bfloat16* img_k;      // [B, M, D]
  const __nv_bfloat16* img_v;
  __nv_bfloat16* txt_k;            // [R, T_alloc, D]
  __nv_bfloat16* txt_v;
  const int* src_row;              // [R, T_alloc] physical row holding text position j of logical row r (null = r)
  __nv_bfloat16* ctx;              // [R, D]
  int B, M, T_alloc, D;
  const StepState* state;          // text position = state->pos (or pos_fixed when null)
  int pos_fixed;
  int chunk_rows;                  // image keys staged per TMA round (<= 512), box_rows * n_boxes
  int box_rows;                    // rows per TMA box (<= 256)
  ChainSync chain;
};

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// One q / k / v element of this step: the split-K partial sums in split order (fixed order: bit-reproducible).
__device__ __forceinline__ float ld_partials(const float* ptr, int n, long long stride) {
  const float a = __ldcg(ptr);
  const float b = n > 1 ? __ldcg(ptr + stride) : 0.f;
  const float c = n > 2 ? __ldcg(ptr + 2 * stride) : 0.f;
  const float d = n > 3 ? __ldcg(ptr + 3 * stride) : 0.f;
  return ((a + b) + c) + d;
}

// Online-softmax state of one 8-lane key group for one query: running max m, running sum l, 8 output dims.
__device__ __forceinline__ void dec_attn_update(float (&sc)[4], const uint4 (&w)[4], float& m, float& l, float (&acc)[8]) {
  float m_new = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])), m);
  if (m_new == -INFINITY) return;  // nothing valid seen yet
  const float scale = __expf(m - m_new);
  float p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __expf(sc[i] - m_new);
  l = l * scale + (p[0] + p[1]) + (p[2] + p[3]);
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] *= scale;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float f[8];
    bf16x8_to_f32(w[i], f);
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = fmaf(p[i], f[d], acc[d]);
  }
  m = m_new;
}

// Persistent, double-buffered streaming kernel: CTA c handles the (image, head) items c, c+G, c+2G, ...  Each item's
// image K/V slice (M rows of 128 B, constant during decoding) is fetched by TMA into one of two shared-memory
// buffers; the first two fetches are issued BEFORE the dependency wait, so under programmatic dependent launch most
// of this kernel's HBM stream overlaps the QKV GEMM that precedes it, and later fetches overlap the arithmetic of
// the previous item.  128 threads = 16 key groups x 8 lanes (8 head dims each, 128-bit accesses); scores are folded
// into per-group online-softmax states merged at the end (flash-decoding style), single pass over K and V.
//
// kPipe (NQ == 1 only): software pipelining of the two global-memory round trips an item used to expose -- the step's own
// q/k/v of item k+1 are requested at the top of item k, and the text K/V rows of an item are requested before its
// image-key loop (shared memory) and consumed after it.  At 256 rows a CTA walks ~10 items, so the exposed latencies
// (not the HBM stream) bounded the kernel: 58 us for 158 MB.
template <int NQ, bool kPipe = false>
__global__ void __launch_bounds__(128)
decode_attn_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const DecAttnParams p) {
  constexpr bool kPipeOn = kPipe && NQ == 1;
  extern __shared__ uint8_t attn_dyn[];
  uint8_t* sbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(attn_dyn) + 127) & ~uintptr_t(127));
  const size_t kv_bytes = static_cast<size_t>(p.chunk_rows) * 128;  // one of K / V of one unit
  __shared__ uint64_t bars[2];
  __shared__ float q_s[NQ][64];
  __shared__ float red_m[4][NQ];
  __shared__ float red_l[4][NQ];
  __shared__ float red_acc[4][NQ][64];

  const int tid = threadIdx.x;
  const int D = p.D;
  const int H = D / 64;
  const int n_items = p.B * H;
  const int G = gridDim.x;
  const int n_chunks = (p.M + p.chunk_rows - 1) / p.chunk_rows;
  const int n_my = (n_items - static_cast<int>(blockIdx.x) + G - 1) / G;
  const int n_units = n_my * n_chunks;

  griddep_launch_early();
  tl_mark(100003);
  if (tid == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue_unit = [&](int u) {  // one thread
    const int item = blockIdx.x + (u / n_chunks) * G;
    const int c = u - (u / n_chunks) * n_chunks;
    const int b = item / H, h = item - b * H;
    uint8_t* sK = sbase + static_cast<size_t>(u & 1) * 2 * kv_bytes;
    uint8_t* sV = sK + kv_bytes;
    const int rows_c = min(p.chunk_rows, p.M - c * p.chunk_rows);
    const int nb = (rows_c + p.box_rows - 1) / p.box_rows;
    mbar_arrive_expect_tx(&bars[u & 1], static_cast<uint32_t>(2 * nb * p.box_rows * 128));
    for (int i = 0; i < nb; ++i) {
      const int grow = b * p.M + c * p.chunk_rows + i * p.box_rows;
      const int gcol = h * 64;
      tma_load_2d(sK + static_cast<size_t>(i) * p.box_rows * 128, &tmK, &bars[u & 1], gcol, grow);
      tma_load_2d(sV + static_cast<size_t>(i) * p.box_rows * 128, &tmV, &bars[u & 1], gcol, grow);
    }
  };
  if (tid == 0) {
    issue_unit(0);
    if (n_units > 1) issue_unit(1);
  }
  bool finished;
  if (p.chain.counters != nullptr) {
    // `finished` only changes between steps (full dependency): in a finished step no kernel waits or signals
    finished = (p.state != nullptr && p.state->finished);
    if (!finished) chain_wait(p.chain);
  } else {
    griddep_wait();
    finished = (p.state != nullptr && p.state->finished);
  }
  if (finished) {  // never leave with a bulk copy in flight into this CTA's shared memory
    mbar_wait(&bars[0], 0);
    if (n_units > 1) mbar_wait(&bars[1], 0);
    return;
  }
  tl_mark(3);
  const int pos = (p.state != nullptr) ? p.state->pos : p.pos_fixed;
  const int n_txt = pos + 1;
  const int grp = tid >> 3;
  const int gl = tid & 7;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  // This step's q/k/v of every item of this CTA are requested together (one L2 round trip instead of one per item):
  // threads 0-63 hold (q, k) of dim tid, threads 64-127 hold v of dim tid-64.
  constexpr int kPre = (NQ == 1) ? 4 : 1;
  float pre_a[kPre], pre_b[kPre];
#pragma unroll
  for (int k = 0; k < kPre; ++k) {
    pre_a[k] = 0.f;
    pre_b[k] = 0.f;
    if (NQ == 1 && k < n_my) {
      const int item = blockIdx.x + k * G;
      const int b = item / H, h = item - b * H;
      const float* row = p.qkv + static_cast<long long>(b) * 3 * D + h * 64;
      if (tid < 64) {
        pre_a[k] = ld_partials(row + tid, p.n_partials, p.partial_stride);
        pre_b[k] = ld_partials(row + D + tid, p.n_partials, p.partial_stride);
      } else {
        pre_a[k] = ld_partials(row + 2 * D + tid - 64, p.n_partials, p.partial_stride);
      }
    }
  }

  // kPipe: (q, k | v) of the NEXT item, requested one item ahead
  float nxt_a = 0.f, nxt_b = 0.f;
  auto request_qkv = [&](int k, float& a, float& b2) {
    if (k < n_my) {
      const int item = blockIdx.x + k * G;
      const int b = item / H, h = item - b * H;
      const float* row = p.qkv + static_cast<long long>(b) * 3 * D + h * 64;
      if (tid < 64) {
        a = ld_partials(row + tid, p.n_partials, p.partial_stride);
        b2 = ld_partials(row + D + tid, p.n_partials, p.partial_stride);
      } else {
        a = ld_partials(row + 2 * D + tid - 64, p.n_partials, p.partial_stride);
      }
    }
  };
  if (kPipeOn) request_qkv(kPre, nxt_a, nxt_b);   // items 0..kPre-1 were requested above

  for (int k = 0; k < n_my; ++k) {
    const int item = blockIdx.x + k * G;
    const int b = item / H, h = item - b * H;
    float cur_a = 0.f, cur_b = 0.f;
    if (kPipeOn && k >= kPre) {
      cur_a = nxt_a;
      cur_b = nxt_b;
      request_qkv(k + 1, nxt_a, nxt_b);
    }
    // ---- q (scaled by 1/8 in fp32 like the reference scales Q), append this step's K/V (bf16) ----
    for (int qi = 0; qi < NQ; ++qi) {
      const int r = b * NQ + qi;
      const float* row = p.qkv + static_cast<long long>(r) * 3 * D + h * 64;
      const float* bias = p.bqkv + h * 64;
      const bool have = (NQ == 1) && (k < kPre);
      if (tid < 64) {
        const float qv = have ? pre_a[k < kPre ? k : 0] : (kPipeOn ? cur_a : ld_partials(row + tid, p.n_partials, p.partial_stride));
        const float kv = have ? pre_b[k < kPre ? k : 0] : (kPipeOn ? cur_b : ld_partials(row + D + tid, p.n_partials, p.partial_stride));
        q_s[qi][tid] = (qv + bias[tid]) * 0.125f;
        p.txt_k[(static_cast<long long>(r) * p.T_alloc + pos) * D + h * 64 + tid] = __float2bfloat16_rn(kv + bias[D + tid]);
      } else {
        const int d = tid - 64;
        const float vv = have ? pre_a[k < kPre ? k : 0] : (kPipeOn ? cur_a : ld_partials(row + 2 * D + d, p.n_partials, p.partial_stride));
        p.txt_v[(static_cast<long long>(r) * p.T_alloc + pos) * D + h * 64 + d] = __float2bfloat16_rn(vv + bias[2 * D + d]);
      }
    }
    __syncthreads();

    float qreg[NQ][8];
    float m_run[NQ], l_run[NQ], acc[NQ][8];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
      m_run[qi] = -INFINITY;
      l_run[qi] = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        qreg[qi][d] = q_s[qi][gl * 8 + d];
        acc[qi][d] = 0.f;
      }
    }
    // ---- text keys (global loads): each beam row has its own history (through the src_row indirection) ----
    auto text_chunk_load = [&](int r, int base, uint4 (&u)[4], uint4 (&w)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = base + grp + 16 * i;
        if (j < n_txt) {
          const int pr = (p.src_row != nullptr) ? p.src_row[r * p.T_alloc + j] : r;
          const long long off = (static_cast<long long>(pr) * p.T_alloc + j) * D + h * 64 + gl * 8;
          u[i] = *reinterpret_cast<const uint4*>(p.txt_k + off);
          w[i] = *reinterpret_cast<const uint4*>(p.txt_v + off);
        } else {
          u[i] = make_uint4(0, 0, 0, 0);
          w[i] = make_uint4(0, 0, 0, 0);
        }
      }
    };
    auto text_chunk_use = [&](int qi, int base, const uint4 (&u)[4], const uint4 (&w)[4]) {
      float sc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float kf[8];
        bf16x8_to_f32(u[i], kf);
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) a = fmaf(qreg[qi][d], kf[d], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        sc[i] = (base + grp + 16 * i < n_txt) ? a : -INFINITY;
      }
      dec_attn_update(sc, w, m_run[qi], l_run[qi], acc[qi]);
    };
    uint4 tu[4], tw[4];                       // kPipe: text positions 0..63 of this item, in flight across the image loop
    if (kPipeOn) {
      text_chunk_load(b, 0, tu, tw);
    } else {
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
        const int r = b * NQ + qi;
        for (int base = 0; base < n_txt; base += 64) {
          uint4 u[4], w[4];
          text_chunk_load(r, base, u, w);
          text_chunk_use(qi, base, u, w);
        }
      }
    }
    // ---- image keys from shared memory: shared by the NQ beams of this image ----
    for (int c = 0; c < n_chunks; ++c) {
      const int u_idx = k * n_chunks + c;
      const uint8_t* sK = sbase + static_cast<size_t>(u_idx & 1) * 2 * kv_bytes;
      const uint8_t* sV = sK + kv_bytes;
      mbar_wait(&bars[u_idx & 1], static_cast<uint32_t>((u_idx >> 1) & 1));
      const int rows_c = min(p.chunk_rows, p.M - c * p.chunk_rows);
      for (int base = 0; base < rows_c; base += 64) {  // uniform trip count: the shuffles below stay converged
        uint4 u[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int s = base + grp + 16 * i;
          const bool ok = s < rows_c;
          u[i] = ok ? *reinterpret_cast<const uint4*>(sK + static_cast<size_t>(s) * 128 + gl * 16) : make_uint4(0, 0, 0, 0);
          w[i] = ok ? *reinterpret_cast<const uint4*>(sV + static_cast<size_t>(s) * 128 + gl * 16) : make_uint4(0, 0, 0, 0);
        }
        float kf[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) bf16x8_to_f32(u[i], kf[i]);
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          float sc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < 8; ++d) a = fmaf(qreg[qi][d], kf[i][d], a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            sc[i] = (base + grp + 16 * i < rows_c) ? a : -INFINITY;
          }
          dec_attn_update(sc, w, m_run[qi], l_run[qi], acc[qi]);
        }
      }
      __syncthreads();  // everyone is done with this buffer: refill it with the unit after next
      if (tid == 0 && u_idx + 2 < n_units) issue_unit(u_idx + 2);
    }
    if (kPipeOn) {
      text_chunk_use(0, 0, tu, tw);
      for (int base = 64; base < n_txt; base += 64) {   // captions longer than 64 tokens: the rest the plain way
        uint4 u[4], w[4];
        text_chunk_load(b, base, u, w);
        text_chunk_use(0, base, u, w);
      }
    }
    // ---- merge the 16 group states: 4 groups of a warp by shuffles, the 4 warps through smem ----
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        const float m_o = __shfl_xor_sync(0xffffffffu, m_run[qi], o);
        const float l_o = __shfl_xor_sync(0xffffffffu, l_run[qi], o);
        const float m_n = fmaxf(m_run[qi], m_o);
        const float sa = (m_run[qi] == -INFINITY) ? 0.f : __expf(m_run[qi] - m_n);
        const float sb = (m_o == -INFINITY) ? 0.f : __expf(m_o - m_n);
        l_run[qi] = l_run[qi] * sa + l_o * sb;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          const float a_o = __shfl_xor_sync(0xffffffffu, acc[qi][d], o);
          acc[qi][d] = acc[qi][d] * sa + a_o * sb;
        }
        m_run[qi] = m_n;
      }
      if (lane < 8) {
        if (lane == 0) { red_m[warp][qi] = m_run[qi]; red_l[warp][qi] = l_run[qi]; }
#pragma unroll
        for (int d = 0; d < 8; ++d) red_acc[warp][qi][lane * 8 + d] = acc[qi][d];
      }
    }
    __syncthreads();
    for (int i = tid; i < NQ * 64; i += 128) {
      const int qi = i >> 6;
      const int d = i & 63;
      const float mm = fmaxf(fmaxf(red_m[0][qi], red_m[1][qi]), fmaxf(red_m[2][qi], red_m[3][qi]));
      float l = 0.f, a = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float sw = (red_m[w][qi] == -INFINITY) ? 0.f : __expf(red_m[w][qi] - mm);
        l += red_l[w][qi] * sw;
        a += red_acc[w][qi][d] * sw;
      }
      p.ctx[static_cast<long long>(b * NQ + qi) * D + h * 64 + d] = __float2bfloat16_rn(a / l);
    }
    __syncthreads();  // q_s / red_* are reused by the next item
  }
  tl_mark(200003);
  chain_signal(p.chain);
}

// ------------------------------------------------------------------------------------------------
// fp32-grade parity mode (engine option "parity"): attention in plain fp32 -- q, k, v and both K/V caches stay fp32 and the
// context rows leave in the GEMMs' split operand format [hi | lo | hi].  One warp per (sequence, head, query row); scores
// live in shared memory.  Not a fast path: it exists so that the whole engine can be compared with the fp32 reference at
// the 1e-3 logit tolerance of BASELINE.json's north star.
// ------------------------------------------------------------------------------------------------
struct AttnF32Params {
  const float* q;
  const float* k;
  const float* v;
  __nv_bfloat16* out;            // split3 rows: [S, 3 * d_model] per batch element
  int B, S, H, d_model;
  long long q_rs, kv_rs, q_bs, kv_bs, o_bs;   // row / batch strides in elements (output row stride = 3 * d_model)
};

__device__ __forceinline__ void store_split3_pair(__nv_bfloat16* row, int d_model, int col, float a, float b) {
  uint32_t hi, lo;
  pack_split2(a, b, hi, lo);
  *reinterpret_cast<uint32_t*>(row + col) = hi;
  *reinterpret_cast<uint32_t*>(row + d_model + col) = lo;
  *reinterpret_cast<uint32_t*>(row + 2 * d_model + col) = hi;
}

__global__ void __launch_bounds__(128) attn_f32_kernel(const AttnF32Params p) {
  extern __shared__ float attn_f32_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q_s = attn_f32_smem + warp * (64 + p.S);
  float* sc = q_s + 64;
  const long long item = static_cast<long long>(blockIdx.x) * 4 + warp;
  const long long total = static_cast<long long>(p.B) * p.H * p.S;
  if (item >= total) return;
  const int row = static_cast<int>(item % p.S);
  const int h = static_cast<int>((item / p.S) % p.H);
  const int b = static_cast<int>(item / (static_cast<long long>(p.S) * p.H));
  const float* qg = p.q + b * p.q_bs + static_cast<long long>(row) * p.q_rs + h * 64;
  const float* kg = p.k + b * p.kv_bs + h * 64;
  const float* vg = p.v + b * p.kv_bs + h * 64;
  q_s[lane] = qg[lane] * 0.125f;             // Q / sqrt(64) before the product (reference layers/bert/modeling_bert.py:42-43)
  q_s[lane + 32] = qg[lane + 32] * 0.125f;
  __syncwarp();
  float mx = -INFINITY;
  for (int j0 = 0; j0 < p.S; j0 += 32) {
    const int j = j0 + lane;
    if (j < p.S) {
      const float4* kr = reinterpret_cast<const float4*>(kg + static_cast<long long>(j) * p.kv_rs);
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 kk = kr[d];
        a = fmaf(q_s[4 * d], kk.x, a); a = fmaf(q_s[4 * d + 1], kk.y, a);
        a = fmaf(q_s[4 * d + 2], kk.z, a); a = fmaf(q_s[4 * d + 3], kk.w, a);
      }
      sc[j] = a;
      mx = fmaxf(mx, a);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < p.S; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float a0 = 0.f, a1 = 0.f;
  for (int j = 0; j < p.S; ++j) {
    const float2 vv = *reinterpret_cast<const float2*>(vg + static_cast<long long>(j) * p.kv_rs + 2 * lane);
    a0 = fmaf(sc[j], vv.x, a0);
    a1 = fmaf(sc[j], vv.y, a1);
  }
  __nv_bfloat16* orow = p.out + b * p.o_bs + static_cast<long long>(row) * 3 * p.d_model;
  store_split3_pair(orow, p.d_model, h * 64 + 2 * lane, a0 / sum, a1 / sum);
}

struct DecAttnF32Params {
  const float* qkv;               // split-K partial sums of this step's q | k | v, as in DecAttnParams
  int n_partials;
  long long partial_stride;
  const float* bqkv;
  const float* img_k;             // fp32 [B, M, D]
  const float* img_v;
  float* txt_k;                   // fp32 [R, T_alloc, D]
  float* txt_v;
  const int* src_row;
  __nv_bfloat16* ctx;             // split3 rows [R, 3 * D]
  int R, beam, M, T_alloc, D;
  const StepState* state;
  int pos_fixed;
  ChainSync chain;
};

// one warp per (sequence r, head h); 4 warps per CTA
__global__ void __launch_bounds__(128) decode_attn_f32_kernel(const DecAttnF32Params p) {
  extern __shared__ float attn_f32_smem[];
  griddep_launch_early();
  bool finished;
  if (p.chain.counters != nullptr) {
    finished = (p.state != nullptr && p.state->finished);
    if (!finished) chain_wait(p.chain);
  } else {
    griddep_wait();
    finished = (p.state != nullptr && p.state->finished);
  }
  if (finished) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = p.D, H = D / 64;
  const int pos = (p.state != nullptr) ? p.state->pos : p.pos_fixed;
  const int n_keys = p.M + pos + 1;
  float* q_s = attn_f32_smem + warp * (192 + p.M + p.T_alloc);
  float* k_s = q_s + 64;
  float* v_s = q_s + 128;
  float* sc = q_s + 192;
  const int item = blockIdx.x * 4 + warp;
  if (item < p.R * H) {
    const int r = item / H, h = item - r * H;
    const int b = r / p.beam;
    const float* row = p.qkv + static_cast<long long>(r) * 3 * D + h * 64;
    const float* bias = p.bqkv + h * 64;
    for (int d = lane; d < 64; d += 32) {
      q_s[d] = (ld_partials(row + d, p.n_partials, p.partial_stride) + bias[d]) * 0.125f;
      const float kn = ld_partials(row + D + d, p.n_partials, p.partial_stride) + bias[D + d];
      const float vn = ld_partials(row + 2 * D + d, p.n_partials, p.partial_stride) + bias[2 * D + d];
      k_s[d] = kn;
      v_s[d] = vn;
      p.txt_k[(static_cast<long long>(r) * p.T_alloc + pos) * D + h * 64 + d] = kn;
      p.txt_v[(static_cast<long long>(r) * p.T_alloc + pos) * D + h * 64 + d] = vn;
    }
    __syncwarp();
    auto key_ptr = [&](int j, bool want_v) -> const float* {   // row of key j (j != newest)
      if (j < p.M) return (want_v ? p.img_v : p.img_k) + (static_cast<long long>(b) * p.M + j) * D + h * 64;
      const int t = j - p.M;
      const int pr = (p.src_row != nullptr) ? p.src_row[r * p.T_alloc + t] : r;
      return (want_v ? p.txt_v : p.txt_k) + (static_cast<long long>(pr) * p.T_alloc + t) * D + h * 64;
    };
    float mx = -INFINITY;
    for (int j0 = 0; j0 < n_keys; j0 += 32) {
      const int j = j0 + lane;
      if (j < n_keys) {
        float a = 0.f;
        if (j == n_keys - 1) {
#pragma unroll
          for (int d = 0; d < 64; ++d) a = fmaf(q_s[d], k_s[d], a);
        } else {
          const float4* kr = reinterpret_cast<const float4*>(key_ptr(j, false));
#pragma unroll
          for (int d = 0; d < 16; ++d) {
            const float4 kk = kr[d];
            a = fmaf(q_s[4 * d], kk.x, a); a = fmaf(q_s[4 * d + 1], kk.y, a);
            a = fmaf(q_s[4 * d + 2], kk.z, a); a = fmaf(q_s[4 * d + 3], kk.w, a);
          }
        }
        sc[j] = a;
        mx = fmaxf(mx, a);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n_keys; j += 32) {
      const float e = expf(sc[j] - mx);
      sc[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j + 1 < n_keys; ++j) {
      const float2 vv = *reinterpret_cast<const float2*>(key_ptr(j, true) + 2 * lane);
      a0 = fmaf(sc[j], vv.x, a0);
      a1 = fmaf(sc[j], vv.y, a1);
    }
    a0 = fmaf(sc[n_keys - 1], v_s[2 * lane], a0);
    a1 = fmaf(sc[n_keys - 1], v_s[2 * lane + 1], a1);
    store_split3_pair(p.ctx + static_cast<long long>(r) * 3 * D, D, h * 64 + 2 * lane, a0 / sum, a1 / sum);
  }
  chain_signal(p.chain);
}

}  // namespace gitb200
