// One persistent kernel per greedy decode step (<= 64 sequences): the 6 decoder layers, the tied LM head with the
// arg-max / log-sum-exp folded into it, the reference's greedy bookkeeping and the next step's token embedding
// (reference layers/decoder.py:313-417, 65-78; layers/bert/modeling_bert.py:92-334) -- replacing the 45-launch chain of
// gitb200.cu::step_layers for that case.
//
// Why: at <= 64 rows every kernel of the chain is latency bound (TMA -> tcgen05 -> TMEM -> epilogue -> flag, ~7 us per
// hop, ~400 us per step against an HBM floor of ~50 us).  Here the HBM stream is decoupled from the dependency chain:
//   * 148 CTAs (one per SM, cooperative launch), 8 compute warps + 1 producer warp each;
//   * every byte the step reads from HBM -- the weight tiles a CTA owns in each GEMM phase, the image K/V of its attention
//     items and the text K/V so far -- flows through a 12 x 16 KB shared-memory ring that the producer warp fills in
//     program order with TMA (bulk copies of pre-packed weight tiles, 64-key swizzled K | V box pairs), running as far
//     ahead of the compute warps as the ring allows, across phase boundaries (only the text K/V of the CURRENT layer waits
//     for that layer's QKV phase);
//   * phases (QKV | attention | out-proj | LN | fc1 | fc2 | LN per layer, then LM head | selection + embedding) are
//     separated by a grid barrier (one release-add + acquire-spin on a global counter, 1.7 us on B200 -- the cheapest of
//     the variants tools/barrier_bench.cu times); what crosses a barrier is only the <= 64-row activations, read straight
//     from L2 into mma.sync fragments;
//   * weights are stationary per CTA: a GEMM phase gives CTA c a few 8-feature tiles over a 768-long reduction; the 8 warps
//     split a tile 4 row tiles x 2 K halves.  fc2 (K = 3072) is dealt as 4 k slices x 96 feature tiles over 128 CTAs, its
//     four partial sums meet in slice order in the LayerNorm phase: no atomics anywhere, the step is bit-reproducible;
//   * attention: the 64-key chunks of a CTA's (sequence, head) items are dealt round-robin to its 8 warps, online-softmax
//     states in shared memory, fixed-order merge.
// The skinny GEMMs and the 1-row attention are HBM-bound byte work (arithmetic intensity ~rows FLOP/B): they use the
// warp-level mma.sync path fed from shared memory (measured limit here: ~0.77 us per 64 x 8 x 768 tile per SM, which is
// what the 26-tile LM-head slice of a CTA costs); tcgen05 / TMEM stay with the compute-bound encoder and prefill GEMMs:
// with the weights as the M operand a tcgen05 tile needs >= 64 weight rows per CTA (3/4 of the SMs, and of the HBM stream,
// would idle in the layer GEMMs); with the activations as the M operand every CTA would have to stage the 64 x 768
// activations (96 KB) in shared memory in every phase, which the 192 KB ring leaves no room for.
#pragma once
#include "ptx.cuh"
#include "rowops.cuh"

namespace gitb200 {

constexpr int kMegaComputeWarps = 8;
constexpr int kMegaThreads = (kMegaComputeWarps + 1) * 32;
constexpr int kMegaSlots = 12;
constexpr int kMegaSlotBytes = 16384;
constexpr int kMegaTileBytes = 12288;      // 8 output features x 768 k x bf16, in mma-fragment order (pack_tiles_kernel)
constexpr int kMegaKvRows = 64;            // keys per attention chunk: one ring slot = K box (8 KB) | V box (8 KB)
constexpr int kMegaMaxRows = 64;
constexpr int kMegaD = 768, kMegaF = 3072, kMegaH = 12;
constexpr int kMegaAttItems = 6;           // (sequence, head) attention items of the busiest CTA: ceil(64 * 12 / 148)
constexpr int kMegaAttState = 72;          // floats per (item, warp) softmax state: 64 output dims, running max, 4 lane sums
constexpr unsigned int kMegaSpinLimit = 1u << 18;   // bounded waits: a protocol bug must end in an error code, not a hung device

struct MegaLayer {
  const uint8_t* wqkv;     // [288] tiles: features 8t .. 8t+7 of the fused q | k | v projection
  const uint8_t* wo;       // [96]
  const uint8_t* w1;       // [384]
  const uint8_t* w2;       // [96][4]: feature tile x 768-wide k slice (CTA c < 128: slice c & 3 of tiles 3 (c >> 2) .. + 2)
  const float* bqkv; const float* bo; const float* b1; const float* b2;
  const float* lnag; const float* lnab; const float* lnog; const float* lnob;
  __nv_bfloat16* txt_k;    // [R, T_alloc, 768]
  __nv_bfloat16* txt_v;
};

struct MegaParams {
  MegaLayer layer[6];
  const uint8_t* lm;       // [ceil(V / 8)] tiles of the tied word-embedding matrix
  const float* lm_bias;
  const float* words;      // fp32 [V, 768] (embedding gather)
  const float* positions;  // fp32 [max_pos, 768]
  const float* lnemb_g; const float* lnemb_b;
  int R, M, T_alloc, V, n_layers;
  // activations (global, L2 resident)
  float* x;                // [R, 768] residual stream (post-LayerNorm)
  float* y;                // [R, 768] pre-LayerNorm sum (after the attention block)
  float* ypart;            // [4][R, 768] fc2 partial sums of the four 768-wide k slices (summed in slice order by the LN)
  __nv_bfloat16* hb;       // [R, 768] bf16 copy of x (GEMM operand)
  __nv_bfloat16* qb;       // [R, 768] q (+bias) / 8
  __nv_bfloat16* ctx;      // [R, 768]
  __nv_bfloat16* ub;       // [R, 3072]
  float* part_max; float* part_sum; int* part_arg;   // [R, gridDim.x] LM-head partials
  // search state (the same objects greedy_select_kernel works on)
  StepState* state;
  long long* tokens_out; int max_steps; float* logprob_sum; long long* next_token; const long long* forced;
  float* step_logits; int eos;
  const long long* row_prefix; int row_prefix_stride; const int* row_prefix_lens;   // per-row prefixes (see SelectParams)
  unsigned int* barrier;   // [2] grid-barrier counters, used alternately by successive steps
  int* error;              // set non-zero when a bounded spin gave up (the host reports it)
};

// ---- packing: [N, K] row-major bf16 -> tiles of 8 features x 768 k in fragment order ------------------------------------
// Tile layout: 48 k-steps x 32 lanes x 8 bytes.  Lane (g = lane / 4, t = lane % 4) of k-step s holds the four k values
// k0 + 64 * (s / 4) + 16 t + 4 (s % 4) + {0, 1, 2, 3} of feature 8 tile + g: the B fragment (b0 = first pair, b1 = second
// pair) of an m16n8k16 MMA whose k index has been permuted so that the matching A fragment is 16 CONTIGUOUS bf16 per
// thread and group of four k-steps -- one 256-bit load from the row-major activation matrix, and the four lanes of a row
// together fetch one whole 128-byte line (with 128-bit loads every line was touched by two instructions, and the L1
// wavefronts of the A loads, not L2 bandwidth, set the pace of the GEMM phases: profiles/decode_mega_timeline_r02_call18.txt).
__global__ void __launch_bounds__(256) pack_tiles_kernel(const __nv_bfloat16* __restrict__ W, long long ldw, int n_feat, int k0,
                                                         uint8_t* __restrict__ dst, long long n_tiles, int tile_stride_tiles,
                                                         int tile_offset) {
  const long long total = n_tiles * 48 * 32;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int lane = static_cast<int>(i & 31);
    const int s = static_cast<int>((i >> 5) % 48);
    const long long tile = i / (48 * 32);
    const int g = lane >> 2, t = lane & 3;
    const long long f = tile * 8 + g;
    const int k = k0 + 64 * (s >> 2) + 16 * t + 4 * (s & 3);
    uint2 v = make_uint2(0u, 0u);
    if (f < n_feat) v = *reinterpret_cast<const uint2*>(W + f * ldw + k);
    *reinterpret_cast<uint2*>(dst + (tile * tile_stride_tiles + tile_offset) * kMegaTileBytes + (s * 32 + lane) * 8) = v;
  }
}

// ---- small device helpers ------------------------------------------------------------------------------------------------
// Fine-grained marks of the debug timeline build (tools/mega_timeline.py): thread 0 of CTA 0 stores (%clock64, id) pairs
// into a slice of the timeline buffer reserved once per launch -- no atomics or %globaltimer reads inside the phases.
#ifdef GITB200_TIMELINE
struct MegaTl {
  unsigned long long* buf;
  unsigned int n;
  __device__ __forceinline__ void begin() {
    buf = nullptr; n = 0;
    if (g_tl_buf != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      const unsigned int i = atomicAdd(&g_tl_count, 160u);
      if (i + 160u <= kTimelineMax) {
        buf = g_tl_buf + 2 * i;
        for (int k = 0; k < 160; ++k) { buf[2 * k] = 0ull; buf[2 * k + 1] = 0ull; }
        mark(700000); buf[2 * n] = globaltimer_ns(); buf[2 * n + 1] = 700001ull; ++n;
      }
    }
  }
  __device__ __forceinline__ void mark(int kid) {
    if (buf != nullptr && n < 160u) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
      buf[2 * n] = t; buf[2 * n + 1] = static_cast<unsigned long long>(kid); ++n;
    }
  }
  __device__ __forceinline__ void end() {
    if (buf != nullptr) { mark(700002); if (n < 160u) { buf[2 * n] = globaltimer_ns(); buf[2 * n + 1] = 700003ull; ++n; } }
  }
};
#define MEGA_TL(kid) tlf.mark(kid)
#else
struct MegaTl {
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ void mark(int) {}
  __device__ __forceinline__ void end() {}
};
#define MEGA_TL(kid)
#endif

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// false: gave up (or another wait already had: once the error word is set every further wait returns at once, so a broken
// launch drains in microseconds instead of timing out chunk by chunk)
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, const int* error) {
  for (unsigned int i = 0; i < kMegaSpinLimit; ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((i & 1023u) == 1023u && *reinterpret_cast<const volatile int*>(error) != 0) return false;
  }
  return false;
}

struct MegaRing {
  uint8_t* base;
  uint64_t* full;
  uint64_t* empty;
  uint32_t idx;      // chunks consumed so far (identical in every compute warp)
  int* error;
  volatile uint32_t* issued;   // chunks the producer has armed so far (shared memory)
  __device__ __forceinline__ const uint8_t* acquire() {
    const uint32_t slot = idx % kMegaSlots;
    if (!mbar_wait_bounded(&full[slot], (idx / kMegaSlots) & 1, error)) *error = 2;
    return base + slot * kMegaSlotBytes;
  }
  // chunk idx + n (n < kMegaSlots) without consuming it: the GEMM phases take all tiles of a batch first, so that the MMAs
  // of one tile overlap the shared-memory reads, the K-half exchange and the epilogue of its neighbours
  __device__ __forceinline__ const uint8_t* acquire_ahead(uint32_t n) {
    const uint32_t i = idx + n, slot = i % kMegaSlots;
    if (!mbar_wait_bounded(&full[slot], (i / kMegaSlots) & 1, error)) *error = 2;
    return base + slot * kMegaSlotBytes;
  }
  __device__ __forceinline__ void release() {   // every compute warp, once per chunk
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[idx % kMegaSlots]);
    ++idx;
  }
  // a chunk that ONE warp consumes alone (attention items): that warp waits for it by index and frees it for all eight
  // (A parity wait alone would be ambiguous here: the warps of a CTA walk different items, so a warp may ask for a chunk
  //  whose slot is still two uses back.  It first waits until the producer has ARMED chunk i -- from then until this very
  //  warp releases it, the slot's barrier can only be in chunk i's phase -- and only then for the phase to complete.)
  __device__ __forceinline__ uint8_t* acquire_at(uint32_t i) {
    const uint32_t slot = i % kMegaSlots;
    unsigned int spins = 0;
    while (*issued <= i) {
      if (++spins > (kMegaSpinLimit << 4) || ((spins & 1023u) == 1023u && *reinterpret_cast<const volatile int*>(error) != 0)) { *error = 5; break; }
    }
    if (!mbar_wait_bounded(&full[slot], (i / kMegaSlots) & 1, error)) *error = 2;
    return base + slot * kMegaSlotBytes;
  }
  __device__ __forceinline__ void release_at(uint32_t i) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&empty[i % kMegaSlots])), "r"(kMegaComputeWarps) : "memory");
  }
};

// Grid barrier between phases (compute warps only: 256 threads; the producer warp never waits for a phase).
__device__ __forceinline__ void mega_grid_sync(unsigned int* counter, unsigned int& epoch, int* error, MegaTl& tlf, int tl_id = 0) {
  if (tl_id) tlf.mark(tl_id + 5);                                   // this warp's phase work issued
  named_bar_sync(1, kMegaComputeWarps * 32);
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    if (tl_id) tlf.mark(tl_id + 6);                                 // all compute warps of the CTA are here
    else tl_mark_one(500000 + static_cast<int>(epoch / gridDim.x)); // this CTA arrived at barrier #n
    // release-RMW at gpu scope: cumulative over the CTA's writes that the bar.sync above ordered before this thread, so
    // no separate fence -- __threadfence() is a sequentially-consistent fence (MEMBAR.SC.GPU + L1 invalidate) that cost
    // 0.5-2.5 us per barrier here on top of the release's own MEMBAR.ALL.GPU
    if (tl_id) tlf.mark(tl_id + 7);
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    if (tl_id) tlf.mark(tl_id + 8);
    unsigned int spins = 0;
    while (ld_acquire_gpu(counter) < epoch) {
      if (++spins > kMegaSpinLimit || ((spins & 255u) == 255u && *reinterpret_cast<const volatile int*>(error) != 0)) {
        if (*reinterpret_cast<const volatile int*>(error) == 0) *error = 1;
        break;
      }
    }
    if (tl_id) tlf.mark(tl_id + 9);
    else tl_mark_one(600000 + static_cast<int>(epoch / gridDim.x)); // barrier #n released
  }
  named_bar_sync(1, kMegaComputeWarps * 32);
  if (tl_id) tlf.mark(tl_id + 10);
}

// A operand of one GEMM phase: this warp's 16 rows x 384 k of a row-major bf16 activation matrix, straight from L2.
struct MegaAFrag {
  uint32_t lo[6][8];   // row g     : 16 contiguous k per entry (four k-steps)
  uint32_t hi[6][8];   // row g + 8
};
__device__ __forceinline__ void ldcg_256(uint32_t (&r)[8], const void* p) {
  asm volatile("ld.global.cg.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void mega_load_a(MegaAFrag& a, const __nv_bfloat16* A, long long lda, int rows, int mt, int kh, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = mt * 16 + g, r1 = r0 + 8;
  const __nv_bfloat16* p0 = A + static_cast<long long>(r0) * lda + kh * 384 + 16 * t;
  const __nv_bfloat16* p1 = A + static_cast<long long>(r1) * lda + kh * 384 + 16 * t;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    if (r0 < rows) ldcg_256(a.lo[j], p0 + 64 * j);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) a.lo[j][e] = 0u;
    }
    if (r1 < rows) ldcg_256(a.hi[j], p1 + 64 * j);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) a.hi[j][e] = 0u;
    }
  }
}
// c += A(16 x 384 of this warp) * tile(8 features, this warp's k half), NT tiles at once.  Per tile two independent
// accumulators (even / odd k-steps, added in a fixed order: bit-reproducible) -- a single one would serialise 24 dependent
// MMAs; across the NT tiles the 2 NT chains of a warp keep the tensor pipe busy over the MMA latency, and the NT tiles
// share one K-half exchange.  (9 warps cap the kernel at 168 registers per thread: NT <= 3.)
template <int NT>
__device__ __forceinline__ void mega_mma_tiles(float (&c)[NT][4], const MegaAFrag& a, const uint8_t* const (&tile)[NT], int kh, int lane) {
  const uint2* bp[NT];
  float c1[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    bp[n] = reinterpret_cast<const uint2*>(tile[n]) + kh * 24 * 32 + lane;
    c1[n][0] = c1[n][1] = c1[n][2] = c1[n][3] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      const uint32_t a0[4] = {a.lo[j][2 * i], a.hi[j][2 * i], a.lo[j][2 * i + 1], a.hi[j][2 * i + 1]};
      const uint32_t a1[4] = {a.lo[j][2 * i + 2], a.hi[j][2 * i + 2], a.lo[j][2 * i + 3], a.hi[j][2 * i + 3]};
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint2 b0 = bp[n][(4 * j + i) * 32];
        const uint2 b1 = bp[n][(4 * j + i + 1) * 32];
        mma_bf16_16816(c[n], a0, b0.x, b0.y);
        mma_bf16_16816(c1[n], a1, b1.x, b1.y);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) c[n][e] += c1[n][e];
}
// Sum of the two K halves of NT tiles behind one named barrier: the kh = 1 warp parks its accumulators in shared memory,
// its kh = 0 partner adds them.  Returns true in the warp that owns the result.  `red` = [2 buffers][3 tiles][4 row tiles][32 lanes] float4 (12 KB).
template <int NT>
__device__ __forceinline__ bool mega_combine_n(float (&c)[NT][4], float4* red, int buf, int mt, int kh, int lane) {
  float4* slot = red + (buf * 12 + mt) * 32 + lane;
  if (kh == 1) {
#pragma unroll
    for (int n = 0; n < NT; ++n) slot[n * 128] = make_float4(c[n][0], c[n][1], c[n][2], c[n][3]);
  }
  named_bar_sync(2 + mt, 64);
  if (kh == 0) {
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float4 o = slot[n * 128];
      c[n][0] += o.x; c[n][1] += o.y; c[n][2] += o.z; c[n][3] += o.w;
    }
  }
  return kh == 0;
}
// Two-way variant for the LM head: both warps of a pair end up with the sum (kh0 + kh1, the same operand order in both),
// so that each can run the statistics of ONE of the two rows a thread owns.  `red2` = [2 buffers][2 tiles][8 warps][32 lanes] float4.
template <int NT>
__device__ __forceinline__ void mega_combine_both_n(float (&c)[NT][4], float4* red2, int buf, int warp, int mt, int kh, int lane) {
  float4* mine = red2 + (buf * 16 + warp) * 32 + lane;
  const float4* other = red2 + (buf * 16 + (warp ^ 4)) * 32 + lane;
#pragma unroll
  for (int n = 0; n < NT; ++n) mine[n * 256] = make_float4(c[n][0], c[n][1], c[n][2], c[n][3]);
  named_bar_sync(2 + mt, 64);
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const float4 o = other[n * 256];
    if (kh == 0) { c[n][0] += o.x; c[n][1] += o.y; c[n][2] += o.z; c[n][3] += o.w; }
    else { c[n][0] = o.x + c[n][0]; c[n][1] = o.y + c[n][1]; c[n][2] = o.z + c[n][2]; c[n][3] = o.w + c[n][3]; }
  }
}

__device__ __forceinline__ float2 ldcg_f2(const float* p) { return __ldcg(reinterpret_cast<const float2*>(p)); }

// ---- the kernel ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMegaThreads, 1)
decode_mega_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmTXT, const MegaParams p) {
  extern __shared__ __align__(1024) uint8_t mega_smem_raw[];
  uint8_t* smem = mega_smem_raw + ((1024u - (smem_u32(mega_smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;                                                  // 12 x 16 KB
  // 16 KB used by two phases that never overlap inside a CTA: attention -- per compute warp a 16-row x 128 B q tile
  // (swizzled); GEMM phases -- the K-half exchange buffers of mega_combine_n / mega_combine_both_n
  uint8_t* q_s = smem + kMegaSlots * kMegaSlotBytes;
  float4* redv = reinterpret_cast<float4*>(q_s);
  // attention: softmax states [item][warp][kMegaAttState floats], then the staged q rows of the CTA's items; the LM head keeps
  // this CTA's bias slice here
  float* att_part = reinterpret_cast<float*>(q_s + kMegaComputeWarps * 2048);
  uint4* q_stage = reinterpret_cast<uint4*>(att_part + kMegaAttItems * kMegaComputeWarps * kMegaAttState);     // [items][8] x 16 B
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(q_stage) + kMegaAttItems * 128);
  uint64_t* empty = full + kMegaSlots;
  volatile uint32_t* issued = reinterpret_cast<volatile uint32_t*>(empty + kMegaSlots);

  StepState* st = p.state;
  if (st->finished) return;                     // stable: only the previous launch's selection phase writes it
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int R = p.R, M = p.M;
  const int pos = st->pos, step = st->step, cur_len = st->cur_len;
  const int n_kv = (M + kMegaKvRows - 1) / kMegaKvRows;
  const int n_txt = pos / 64 + 1;               // 64-position boxes of the text K/V cache holding positions 0 .. pos
  const int n_items = R * kMegaH;
  const int my_cta_rev = G - 1 - cta;           // attention items are dealt from the last CTA down (those own fewer weights)
  const int n_my_items = (n_items > my_cta_rev) ? (n_items - my_cta_rev + G - 1) / G : 0;
  const int lm_tiles_total = (p.V + 7) / 8;
  const int lm_per = (lm_tiles_total + G - 1) / G;
  const int lm_t0 = cta * lm_per;
  const int lm_n = max(0, min(lm_per, lm_tiles_total - lm_t0));

  if (tid == 0) {
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmTXT);
    for (int s = 0; s < kMegaSlots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kMegaComputeWarps);
    }
    mbar_fence_init();
    *issued = 0;
    if (cta == 0) p.barrier[(step + 1) & 1] = 0;   // the counter the NEXT step will use
  }
  __syncthreads();

  if (warp == kMegaComputeWarps) {
    // ============================== producer: the step's HBM stream, in program order ==============================
    if (lane == 0) {
      uint32_t i = 0;
      auto slot_ready = [&]() -> uint8_t* {
        const uint32_t slot = i % kMegaSlots;
        if (i >= kMegaSlots && !mbar_wait_bounded(&empty[slot], ((i / kMegaSlots) - 1) & 1, p.error)) *p.error = 3;
        return ring + slot * kMegaSlotBytes;
      };
      auto publish = [&]() {          // the chunk's barrier is armed: consumers may now wait for its phase
        ++i;
        __threadfence_block();
        *issued = i;
      };
      auto tile = [&](const uint8_t* src) {
        uint8_t* dst = slot_ready();
        mbar_arrive_expect_tx(&full[i % kMegaSlots], kMegaTileBytes);
        bulk_load_1d(dst, src, kMegaTileBytes, &full[i % kMegaSlots]);
        publish();
      };
      // one attention chunk: 64 keys of one (sequence, head): K box at the slot's start, V box 8 KB in
      auto kv_pair = [&](const CUtensorMap* tm, int row_k, int row_v, int col) {
        uint8_t* dst = slot_ready();
        mbar_arrive_expect_tx(&full[i % kMegaSlots], kMegaSlotBytes);
        tma_load_2d(dst, tm, &full[i % kMegaSlots], col, row_k);
        tma_load_2d(dst + 8192, tm, &full[i % kMegaSlots], col, row_v);
        publish();
      };
      for (int l = 0; l < p.n_layers; ++l) {
        const MegaLayer& L = p.layer[l];
        if (cta < 144) for (int j = 0; j < 2; ++j) tile(L.wqkv + static_cast<size_t>(cta * 2 + j) * kMegaTileBytes);
        // attention chunks ("units") of this CTA's items, round-major: unit u = r * n_my_items + k is keys 64r .. 64r + 63 of
        // item k (image keys first, then the text rounds); the consumer deals the units to its 8 warps round-robin
        for (int r = 0; r < n_kv + n_txt; ++r) {
          if (r == n_kv) {
            // the text K/V of position `pos` exist once every CTA has passed barrier 7l + 1 (after this layer's QKV phase)
            const unsigned int target = static_cast<unsigned int>(G) * (7u * l + 1u);
            unsigned int spins = 0;
            while (ld_acquire_gpu(p.barrier + (step & 1)) < target) {
              if (++spins > kMegaSpinLimit || ((spins & 255u) == 255u && *reinterpret_cast<const volatile int*>(p.error) != 0)) {
                if (*reinterpret_cast<const volatile int*>(p.error) == 0) *p.error = 4;
                break;
              }
            }
            asm volatile("fence.proxy.async;" ::: "memory");   // other SMs' generic-proxy stores -> this thread's TMA reads
          }
          for (int kk = 0; kk < n_my_items; ++kk) {
            const int item = my_cta_rev + kk * G;
            const int b = item / kMegaH, h = item - b * kMegaH;
            if (r < n_kv)
              kv_pair(&tmKV, (l * 2 + 0) * R * M + b * M + r * kMegaKvRows, (l * 2 + 1) * R * M + b * M + r * kMegaKvRows, h * 64);
            else
              kv_pair(&tmTXT, ((l * 2 + 0) * R + b) * p.T_alloc + (r - n_kv) * 64, ((l * 2 + 1) * R + b) * p.T_alloc + (r - n_kv) * 64, h * 64);
          }
        }
        if (cta < 96) tile(L.wo + static_cast<size_t>(cta) * kMegaTileBytes);
        if (cta < 128) for (int j = 0; j < 3; ++j) tile(L.w1 + static_cast<size_t>(cta * 3 + j) * kMegaTileBytes);
        if (cta < 128) for (int j = 0; j < 3; ++j) tile(L.w2 + static_cast<size_t>(((cta >> 2) * 3 + j) * 4 + (cta & 3)) * kMegaTileBytes);
      }
      for (int j = 0; j < lm_n; ++j) tile(p.lm + static_cast<size_t>(lm_t0 + j) * kMegaTileBytes);
    }
    return;
  }

  // ================================== compute warps ==================================
  MegaRing rg{ring, full, empty, 0u, p.error, issued};
  unsigned int* bar = p.barrier + (step & 1);
  unsigned int epoch = 0;
  MegaTl tlf;
  tlf.begin();
#ifdef GITB200_TIMELINE
#define MEGA_TL_ID(ph) ((l == 2) ? 710000 + (ph) * 100 : 0)
#define MEGA_TL_DEP(a) if (((a).lo[5][7] ^ (a).hi[5][7] ^ (a).lo[0][0]) == 0x9E3779B9u) *p.error = 99;
#else
#define MEGA_TL_ID(ph) 0
#define MEGA_TL_DEP(a)
#endif
  const int mt = warp & 3, kh = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int r0 = mt * 16 + g, r1 = r0 + 8;
  int red_buf = 0;

  for (int l = 0; l < p.n_layers; ++l) {
    const MegaLayer& L = p.layer[l];
    // ------------------------------------------------ P1: q | k | v ------------------------------------------------
    if (cta < 144) {
      MegaAFrag a;
      if (MEGA_TL_ID(1)) tlf.mark(MEGA_TL_ID(1) + 0);
      mega_load_a(a, p.hb, kMegaD, R, mt, kh, lane);
      if (MEGA_TL_ID(1)) { tlf.mark(MEGA_TL_ID(1) + 1); MEGA_TL_DEP(a) tlf.mark(MEGA_TL_ID(1) + 2); }
      float2 bias_j[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) bias_j[j] = __ldg(reinterpret_cast<const float2*>(L.bqkv + (cta * 2 + j) * 8 + 2 * t));
      float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      const uint8_t* tb[2] = {rg.acquire_ahead(0), rg.acquire_ahead(1)};
      if (MEGA_TL_ID(1)) tlf.mark(MEGA_TL_ID(1) + 3);
      mega_mma_tiles<2>(c, a, tb, kh, lane);
      rg.release();
      rg.release();
      const bool own1 = mega_combine_n<2>(c, redv, red_buf, mt, kh, lane);
      if (MEGA_TL_ID(1)) tlf.mark(MEGA_TL_ID(1) + 4);
      if (own1) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int f = (cta * 2 + j) * 8 + 2 * t;
          const float2 bias = bias_j[j];
          const int seg = f / kMegaD, fo = f - seg * kMegaD;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int r = hh ? r1 : r0;
            if (r >= R) continue;
            const float v0 = c[j][2 * hh] + bias.x, v1 = c[j][2 * hh + 1] + bias.y;
            if (seg == 0) {
              *reinterpret_cast<uint32_t*>(p.qb + static_cast<long long>(r) * kMegaD + fo) = pack_bf16(v0 * 0.125f, v1 * 0.125f);
            } else {
              __nv_bfloat16* dst = (seg == 1 ? L.txt_k : L.txt_v) + (static_cast<long long>(r) * p.T_alloc + pos) * kMegaD + fo;
              *reinterpret_cast<uint32_t*>(dst) = pack_bf16(v0, v1);
            }
          }
        }
      }
      red_buf ^= 1;
    }
    mega_grid_sync(bar, epoch, p.error, tlf, MEGA_TL_ID(1));
    // ------------------------------------------------ P2: attention ------------------------------------------------
    // The CTA's 5-6 (sequence, head) items x (image + text) 64-key chunks form U units; warp w takes units w, w + 8, ...
    // (whatever item they belong to), so the tensor pipes of the four SM sub-partitions carry the same load -- one warp
    // per item left two of them with twice the MMAs of the others, and mma.sync issue rate is what bounds this phase.
    // The online-softmax state of (item, warp) lives in shared memory between the units of a warp; S = q K^T and
    // O += P V run on mma.sync with a 16-row q tile whose row 0 is the query (rows 1..15 zero).  At the end warp k merges
    // the 8 states of item k in warp order (bit-reproducible).  Two block-level barriers per phase, none per unit.
    {
      const int rounds = n_kv + n_txt;                      // ring chunks per item (64 keys each)
      const uint32_t att_base = rg.idx;
      const int U = n_my_items * rounds;
      uint8_t* qw = q_s + warp * 2048;
      constexpr float kLog2e = 1.44269504088896340736f;
      // phase start: this warp's q tile zeroed (row 0 is rewritten per unit), its softmax states reset, the q rows staged
#pragma unroll
      for (int jq = 0; jq < 4; ++jq) reinterpret_cast<uint4*>(qw)[lane + 32 * jq] = make_uint4(0, 0, 0, 0);
      if (lane < 4) {
        for (int k = 0; k < n_my_items; ++k) {
          float* sl = att_part + (k * kMegaComputeWarps + warp) * kMegaAttState;
#pragma unroll
          for (int j = 0; j < 8; ++j) { sl[8 * j + 2 * lane] = 0.f; sl[8 * j + 2 * lane + 1] = 0.f; }
          sl[65 + lane] = 0.f;
          if (lane == 0) sl[64] = -INFINITY;
        }
      }
      if (warp < n_my_items && lane < 8) {
        const int item = my_cta_rev + warp * G;
        const int b = item / kMegaH, h = item - b * kMegaH;
        q_stage[warp * 8 + lane] = __ldcg(reinterpret_cast<const uint4*>(p.qb + static_cast<long long>(b) * kMegaD + h * 64) + lane);
      }
      named_bar_sync(1, kMegaComputeWarps * 32);
      for (int u = warp; u < U; u += kMegaComputeWarps) {
        const int r = u / n_my_items, k = u - r * n_my_items;
        // q tile (already scaled by 1/8), 128B-swizzled like the K/V boxes: row 0 <- the item's q (chunk c of row 0 sits at c << 4)
        if (lane < 8) *reinterpret_cast<uint4*>(qw + (lane << 4)) = q_stage[k * 8 + lane];
        __syncwarp();
        uint32_t qa[4][4];
        {
          const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int chunk = 2 * kk + (lane >> 4);
            ldmatrix_x4(qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], smem_u32(qw) + row * 128 + ((chunk ^ (row & 7)) << 4));
          }
        }
        float* sl = att_part + (k * kMegaComputeWarps + warp) * kMegaAttState;
        float o[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;       // row g of the q tile (only g == 0 is a real row)
        if (g == 0) {
          m_run = sl[64];
          l_run = sl[65 + t];
#pragma unroll
          for (int j = 0; j < 8; ++j) { o[j][0] = sl[8 * j + 2 * t]; o[j][1] = sl[8 * j + 2 * t + 1]; }
        }
        // 64 keys: rows [row0, row0 + 64) of a K box at sK and of a V box at sV; key index = key0 + row, valid below key_end
        auto keys64 = [&](uint32_t sK, uint32_t sV, int row0, int key0, int key_end) {
          float sc[8][4];
#pragma unroll
          for (int jn = 0; jn < 8; ++jn) {
            sc[jn][0] = sc[jn][1] = sc[jn][2] = sc[jn][3] = 0.f;
            const int krow = row0 + 8 * jn + (lane & 7);
#pragma unroll
            for (int kk2 = 0; kk2 < 2; ++kk2) {
              const int chunk = 4 * kk2 + (lane >> 3);
              uint32_t b0, b1, b2, b3;
              ldmatrix_x4(b0, b1, b2, b3, sK + krow * 128 + ((chunk ^ (krow & 7)) << 4));
              mma_bf16_16816(sc[jn], qa[2 * kk2], b0, b1);
              mma_bf16_16816(sc[jn], qa[2 * kk2 + 1], b2, b3);
            }
          }
          float mx = -INFINITY;
#pragma unroll
          for (int jn = 0; jn < 8; ++jn) {
            const int key = key0 + 8 * jn + 2 * t;
            if (key >= key_end) sc[jn][0] = -INFINITY;
            if (key + 1 >= key_end) sc[jn][1] = -INFINITY;
            mx = fmaxf(mx, fmaxf(sc[jn][0], sc[jn][1]));
          }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          const float m_new = fmaxf(m_run, mx);
          const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
          const float corr = ex2_approx((m_run - m_use) * kLog2e);
          m_run = m_new;
          l_run *= corr;
#pragma unroll
          for (int j = 0; j < 8; ++j) { o[j][0] *= corr; o[j][1] *= corr; }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {                  // 16 keys per k-step = score tiles 2kk, 2kk + 1
            const float p0 = ex2_approx((sc[2 * kk][0] - m_use) * kLog2e), p1 = ex2_approx((sc[2 * kk][1] - m_use) * kLog2e);
            const float p2 = ex2_approx((sc[2 * kk + 1][0] - m_use) * kLog2e), p3 = ex2_approx((sc[2 * kk + 1][1] - m_use) * kLog2e);
            l_run += (p0 + p1) + (p2 + p3);
            const uint32_t pa[4] = {pack_bf16(p0, p1), 0u, pack_bf16(p2, p3), 0u};   // rows 8..15 of the q tile do not exist
            const int vrow = row0 + 16 * kk + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int chunk = 2 * jj + (lane >> 4);
              uint32_t b0, b1, b2, b3;
              ldmatrix_x4_trans(b0, b1, b2, b3, sV + vrow * 128 + ((chunk ^ (vrow & 7)) << 4));
              mma_bf16_16816(o[2 * jj], pa, b0, b1);
              mma_bf16_16816(o[2 * jj + 1], pa, b2, b3);
            }
          }
        };
        const uint32_t ci = att_base + static_cast<uint32_t>(u);
        const uint32_t sK = smem_u32(rg.acquire_at(ci));
        if (r < n_kv) keys64(sK, sK + 8192, 0, r * kMegaKvRows, M);               // image keys 64r ..
        else keys64(sK, sK + 8192, 0, (r - n_kv) * 64, pos + 1);                 // text positions 64(r - n_kv) ..
        rg.release_at(ci);
        if (g == 0) {
          if (t == 0) sl[64] = m_run;
          sl[65 + t] = l_run;
#pragma unroll
          for (int j = 0; j < 8; ++j) { sl[8 * j + 2 * t] = o[j][0]; sl[8 * j + 2 * t + 1] = o[j][1]; }
        }
        __syncwarp();                                       // the q tile's row 0 and the state are rewritten by the next unit
      }
      named_bar_sync(1, kMegaComputeWarps * 32);
      if (warp < n_my_items) {
        // merge the 8 softmax states of item `warp` (fixed order: bit-reproducible); lane -> output dims 2 lane, 2 lane + 1
        const int item = my_cta_rev + warp * G;
        const int b = item / kMegaH, h = item - b * kMegaH;
        const float* p0 = att_part + (warp * kMegaComputeWarps) * kMegaAttState;
        float mm = -INFINITY;
        for (int i = 0; i < kMegaComputeWarps; ++i) mm = fmaxf(mm, p0[i * kMegaAttState + 64]);
        float lsum = 0.f, a0 = 0.f, a1 = 0.f;
        for (int i = 0; i < kMegaComputeWarps; ++i) {
          const float* pi = p0 + i * kMegaAttState;
          const float mi = pi[64];
          const float wgt = (mi == -INFINITY) ? 0.f : exp2f((mi - mm) * kLog2e);
          lsum += ((pi[65] + pi[66]) + (pi[67] + pi[68])) * wgt;
          a0 += pi[2 * lane] * wgt;
          a1 += pi[2 * lane + 1] * wgt;
        }
        *reinterpret_cast<uint32_t*>(p.ctx + static_cast<long long>(b) * kMegaD + h * 64 + 2 * lane) = pack_bf16(a0 / lsum, a1 / lsum);
      }
      rg.idx = att_base + static_cast<uint32_t>(U);
    }
    mega_grid_sync(bar, epoch, p.error, tlf, MEGA_TL_ID(2));
    // ------------------------------------------------ P3: attention output projection (+bias +residual) ------------------
    if (cta < 96) {
      MegaAFrag a;
      mega_load_a(a, p.ctx, kMegaD, R, mt, kh, lane);
      const int f = cta * 8 + 2 * t;
      const float2 bias = __ldg(reinterpret_cast<const float2*>(L.bo + f));
      float2 xr[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};      // residual: requested before the MMA needs the tile
      if (kh == 0 && r0 < R) xr[0] = ldcg_f2(p.x + static_cast<long long>(r0) * kMegaD + f);
      if (kh == 0 && r1 < R) xr[1] = ldcg_f2(p.x + static_cast<long long>(r1) * kMegaD + f);
      float c[1][4] = {{0.f, 0.f, 0.f, 0.f}};
      const uint8_t* tb[1] = {rg.acquire_ahead(0)};
      mega_mma_tiles<1>(c, a, tb, kh, lane);
      rg.release();
      if (mega_combine_n<1>(c, redv, red_buf, mt, kh, lane)) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int r = hh ? r1 : r0;
          if (r >= R) continue;
          *reinterpret_cast<float2*>(p.y + static_cast<long long>(r) * kMegaD + f) =
              make_float2(xr[hh].x + (c[0][2 * hh] + bias.x), xr[hh].y + (c[0][2 * hh + 1] + bias.y));
        }
      }
      red_buf ^= 1;
    }
    mega_grid_sync(bar, epoch, p.error, tlf, MEGA_TL_ID(3));
    // ------------------------------------------------ P4 / P7: LayerNorm(y) -> x, hb (warp 0 of CTA r: row r) ----------------
    // from_parts: the input row is x + ((p0 + p1) + p2) + p3 + bias (fc2's four k-slice partials, fixed order)
    auto layer_norm_rows = [&](const float* gamma, const float* beta, bool from_parts, const float* bias) {
      const int row = cta;           // one row per CTA (R <= 64 of them): 8 rows per CTA made 8 SMs pull all the rows through their L2 ports
      if (row < R && warp == 0) {
        float4 v[6];
        if (!from_parts) {
          const float4* yp = reinterpret_cast<const float4*>(p.y + static_cast<long long>(row) * kMegaD);
#pragma unroll
          for (int i = 0; i < 6; ++i) v[i] = __ldcg(yp + i * 32 + lane);
        } else {
          const long long ps = static_cast<long long>(R) * kMegaD / 4;     // float4s per partial buffer
          const float4* pp = reinterpret_cast<const float4*>(p.ypart + static_cast<long long>(row) * kMegaD);
          const float4* xp = reinterpret_cast<const float4*>(p.x + static_cast<long long>(row) * kMegaD);
          float4 q0[6], q1[6], q2[6], q3[6], xx[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            q0[i] = __ldcg(pp + i * 32 + lane); q1[i] = __ldcg(pp + ps + i * 32 + lane);
            q2[i] = __ldcg(pp + 2 * ps + i * 32 + lane); q3[i] = __ldcg(pp + 3 * ps + i * 32 + lane);
            xx[i] = __ldcg(xp + i * 32 + lane);
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + i * 32 + lane);
            v[i].x = xx[i].x + ((((q0[i].x + q1[i].x) + q2[i].x) + q3[i].x) + bb.x);
            v[i].y = xx[i].y + ((((q0[i].y + q1[i].y) + q2[i].y) + q3[i].y) + bb.y);
            v[i].z = xx[i].z + ((((q0[i].z + q1[i].z) + q2[i].z) + q3[i].z) + bb.z);
            v[i].w = xx[i].w + ((((q0[i].w + q1[i].w) + q2[i].w) + q3[i].w) + bb.w);
          }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float mean = warp_sum(s) * (1.0f / kMegaD);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float a = v[i].x - mean, b2 = v[i].y - mean, c2 = v[i].z - mean, d2 = v[i].w - mean;
          ss += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
        }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / kMegaD) + 1e-12f);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
          const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
          float4 ov;
          ov.x = (v[i].x - mean) * rstd * gm.x + bt.x;
          ov.y = (v[i].y - mean) * rstd * gm.y + bt.y;
          ov.z = (v[i].z - mean) * rstd * gm.z + bt.z;
          ov.w = (v[i].w - mean) * rstd * gm.w + bt.w;
          reinterpret_cast<float4*>(p.x + static_cast<long long>(row) * kMegaD)[i * 32 + lane] = ov;
          uint2 pk;
          pk.x = pack_bf16(ov.x, ov.y);
          pk.y = pack_bf16(ov.z, ov.w);
          reinterpret_cast<uint2*>(p.hb + static_cast<long long>(row) * kMegaD)[i * 32 + lane] = pk;
        }
      }
    };
    layer_norm_rows(L.lnag, L.lnab, false, nullptr);
    mega_grid_sync(bar, epoch, p.error, tlf, MEGA_TL_ID(4));
    // ------------------------------------------------ P5: fc1 + erf-GELU ------------------------------------------------
    if (cta < 128) {
      MegaAFrag a;
      if (MEGA_TL_ID(5)) tlf.mark(MEGA_TL_ID(5) + 0);
      mega_load_a(a, p.hb, kMegaD, R, mt, kh, lane);
      if (MEGA_TL_ID(5)) { tlf.mark(MEGA_TL_ID(5) + 1); MEGA_TL_DEP(a) tlf.mark(MEGA_TL_ID(5) + 2); }
      float2 bias_j[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) bias_j[j] = __ldg(reinterpret_cast<const float2*>(L.b1 + (cta * 3 + j) * 8 + 2 * t));
      float c[3][4];
#pragma unroll
      for (int j = 0; j < 3; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
      const uint8_t* tb[3] = {rg.acquire_ahead(0), rg.acquire_ahead(1), rg.acquire_ahead(2)};
      if (MEGA_TL_ID(5)) tlf.mark(MEGA_TL_ID(5) + 3);
      mega_mma_tiles<3>(c, a, tb, kh, lane);
      rg.release();
      rg.release();
      rg.release();
      const bool own5 = mega_combine_n<3>(c, redv, red_buf, mt, kh, lane);
      if (MEGA_TL_ID(5)) tlf.mark(MEGA_TL_ID(5) + 4);
      if (own5) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int f = (cta * 3 + j) * 8 + 2 * t;
          const float2 bias = bias_j[j];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int r = hh ? r1 : r0;
            if (r >= R) continue;
            *reinterpret_cast<uint32_t*>(p.ub + static_cast<long long>(r) * kMegaF + f) =
                pack_bf16(apply_act(c[j][2 * hh] + bias.x, ACT_GELU_ERF), apply_act(c[j][2 * hh + 1] + bias.y, ACT_GELU_ERF));
          }
        }
      }
      red_buf ^= 1;
    }
    mega_grid_sync(bar, epoch, p.error, tlf, MEGA_TL_ID(5));
    // ------------------------------------------------ P6: fc2, split over CTAs: 32 groups of 24 features x 4 k slices ------
    // (each CTA reads ONE 768-wide slice of the activations; the four partial sums of a feature meet, in slice order, in
    //  the LayerNorm phase below -- bit-reproducible, no atomics)
    if (cta < 128) {
      const int ks = cta & 3, fg = cta >> 2;
      MegaAFrag a;
      mega_load_a(a, p.ub + ks * kMegaD, kMegaF, R, mt, kh, lane);
      float* yp = p.ypart + static_cast<long long>(ks) * R * kMegaD;
      float c[3][4];
#pragma unroll
      for (int j = 0; j < 3; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
      const uint8_t* tb[3] = {rg.acquire_ahead(0), rg.acquire_ahead(1), rg.acquire_ahead(2)};
      mega_mma_tiles<3>(c, a, tb, kh, lane);
      rg.release();
      rg.release();
      rg.release();
      if (mega_combine_n<3>(c, redv, red_buf, mt, kh, lane)) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int f = (fg * 3 + j) * 8 + 2 * t;
          if (r0 < R) *reinterpret_cast<float2*>(yp + static_cast<long long>(r0) * kMegaD + f) = make_float2(c[j][0], c[j][1]);
          if (r1 < R) *reinterpret_cast<float2*>(yp + static_cast<long long>(r1) * kMegaD + f) = make_float2(c[j][2], c[j][3]);
        }
      }
      red_buf ^= 1;
    }
    mega_grid_sync(bar, epoch, p.error, tlf, MEGA_TL_ID(6));
    layer_norm_rows(L.lnog, L.lnob, true, L.b2);
    mega_grid_sync(bar, epoch, p.error, tlf, MEGA_TL_ID(7));
  }

  // ------------------------------------------------ LM head with the greedy statistics folded in ---------------------------
  // After the two K halves of a tile are summed, the kh = 0 warp of a pair keeps the statistics of row g, the kh = 1 warp
  // those of row g + 8: for its row and its 2 columns per tile a thread maintains the running (max, arg max, sum exp) over
  // this CTA's features -- the logits themselves never leave the SM (unless the parity hook asks for them).
  {
    const int my_row = kh ? r1 : r0;
    long long last = -1;
    bool first = (step == 0);                       // no no-repeat mask at a row's first real decision
    if (my_row < R) {
      last = p.next_token[my_row];
      if (p.row_prefix != nullptr) first = cur_len <= p.row_prefix_lens[my_row];
    }
    float smax = -INFINITY, ssum = 0.f;
    int sarg = 0x7fffffff;
    float* bias_s = att_part;                       // this CTA's slice of the output bias (<= 8 * lm_per floats)
    for (int i = tid; i < lm_n * 8; i += kMegaComputeWarps * 32) {
      const int col = lm_t0 * 8 + i;
      bias_s[i] = (col < p.V) ? __ldg(p.lm_bias + col) : 0.f;
    }
    named_bar_sync(1, kMegaComputeWarps * 32);
    if (lm_n > 0) {
      MegaAFrag a;
      mega_load_a(a, p.hb, kMegaD, R, mt, kh, lane);
      // statistics of one tile's two columns of this thread's row (columns arrive in increasing order)
      auto lm_stats = [&](const float (&cc)[4], int j) {
        if (my_row >= R) return;
        const int f = (lm_t0 + j) * 8 + 2 * t;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = f + e;
          if (col >= p.V) continue;
          float v = (kh ? cc[2 + e] : cc[e]) + bias_s[j * 8 + 2 * t + e];
          if (p.step_logits != nullptr) p.step_logits[(static_cast<long long>(step) * R + my_row) * p.V + col] = v;
          if (!first && col == static_cast<int>(last)) v = -10000.0f;        // no-repeat (reference :330)
          if (v > smax) {            // the lowest index wins exact ties
            ssum = ssum * __expf(smax - v) + 1.0f;
            smax = v;
            sarg = col;
          } else {
            ssum += __expf(v - smax);
          }
        }
      };
      int j = 0;
      MEGA_TL(720000);                                // LM head: A loads issued
      for (; j + 2 <= lm_n; j += 2) {                 // two tiles in flight per warp (see mega_mma_tiles)
        float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const uint8_t* tb[2] = {rg.acquire_ahead(0), rg.acquire_ahead(1)};
        MEGA_TL(720100 + j);                          // the pair's tiles landed
        mega_mma_tiles<2>(c, a, tb, kh, lane);
        rg.release();
        rg.release();
        mega_combine_both_n<2>(c, redv, red_buf, warp, mt, kh, lane);
        MEGA_TL(720200 + j);                          // MMAs + exchange done
        red_buf ^= 1;
        lm_stats(c[0], j);
        lm_stats(c[1], j + 1);
        MEGA_TL(720300 + j);                          // statistics done
      }
      if (j < lm_n) {
        float c[1][4] = {{0.f, 0.f, 0.f, 0.f}};
        const uint8_t* tb[1] = {rg.acquire_ahead(0)};
        mega_mma_tiles<1>(c, a, tb, kh, lane);
        rg.release();
        mega_combine_both_n<1>(c, redv, red_buf, warp, mt, kh, lane);
        red_buf ^= 1;
        lm_stats(c[0], j);
      }
    }
    // combine the 4 lanes of a quad (they hold the same row, interleaved column pairs)
#pragma unroll
    for (int o2 = 1; o2 <= 2; o2 <<= 1) {
      const float m_o = __shfl_xor_sync(0xffffffffu, smax, o2);
      const float s_o = __shfl_xor_sync(0xffffffffu, ssum, o2);
      const int a_o = __shfl_xor_sync(0xffffffffu, sarg, o2);
      const float mn = fmaxf(smax, m_o);
      const float sa = (smax == -INFINITY) ? 0.f : __expf(smax - mn);
      const float sb = (m_o == -INFINITY) ? 0.f : __expf(m_o - mn);
      ssum = ssum * sa + s_o * sb;
      if (m_o > smax || (m_o == smax && a_o < sarg)) sarg = a_o;
      smax = mn;
    }
    if (t == 0 && my_row < R) {
      p.part_max[static_cast<long long>(my_row) * G + cta] = smax;
      p.part_sum[static_cast<long long>(my_row) * G + cta] = ssum;
      p.part_arg[static_cast<long long>(my_row) * G + cta] = sarg;
    }
  }
  mega_grid_sync(bar, epoch, p.error, tlf, 0);

  tlf.end();
  // ------------------------------------------------ selection (greedy bookkeeping) + next token's embedding ----------------
  if (cta < R && warp == 0) {
    const int row = cta;
    const int own_prefix = (p.row_prefix != nullptr) ? p.row_prefix_lens[row] : 0;
    const bool in_prefix = (p.row_prefix != nullptr) && cur_len < own_prefix;
    const bool first = (p.row_prefix != nullptr) ? (cur_len == own_prefix) : (step == 0);
    float gm = -INFINITY, gs = 0.f;
    int ga = 0x7fffffff;
    for (int k = lane; k < G; k += 32) {           // increasing CTA order = increasing column order
      const float pm = __ldcg(p.part_max + static_cast<long long>(row) * G + k);
      const float ps = __ldcg(p.part_sum + static_cast<long long>(row) * G + k);
      const int pa2 = __ldcg(p.part_arg + static_cast<long long>(row) * G + k);
      const float mn = fmaxf(gm, pm);
      const float sa = (gm == -INFINITY) ? 0.f : __expf(gm - mn);
      const float sb = (pm == -INFINITY) ? 0.f : __expf(pm - mn);
      gs = gs * sa + ps * sb;
      if (pm > gm || (pm == gm && pa2 < ga)) ga = pa2;
      gm = mn;
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
      const float m_o = __shfl_xor_sync(0xffffffffu, gm, o2);
      const float s_o = __shfl_xor_sync(0xffffffffu, gs, o2);
      const int a_o = __shfl_xor_sync(0xffffffffu, ga, o2);
      const float mn = fmaxf(gm, m_o);
      const float sa = (gm == -INFINITY) ? 0.f : __expf(gm - mn);
      const float sb = (m_o == -INFINITY) ? 0.f : __expf(m_o - mn);
      gs = gs * sa + s_o * sb;
      if (m_o > gm || (m_o == gm && a_o < ga)) ga = a_o;
      gm = mn;
    }
    const long long last = p.next_token[row];
    const bool row_done = (!first) && (!in_prefix) && (last == p.eos);
    long long tok = row_done ? p.eos : ga;                         // EOS forcing: one-hot distribution, log-prob 0
    float lp = row_done ? 0.f : -logf(gs);
    if (in_prefix) {                                               // still feeding this row's prefix
      tok = p.row_prefix[static_cast<long long>(row) * p.row_prefix_stride + cur_len];
      lp = 0.f;
    }
    long long nxt = tok;
    if (p.forced != nullptr) nxt = p.forced[static_cast<long long>(row) * p.max_steps + cur_len];
    if (lane == 0) {
      p.tokens_out[static_cast<long long>(row) * p.max_steps + cur_len] = tok;
      p.logprob_sum[row] += lp;
      p.next_token[row] = nxt;
    }
    // embedding of the token the next step feeds: e = LN(words[nxt] + positions[pos + 1], eps 1e-8)
    {
      long long tk = nxt < 0 ? 0 : (nxt >= p.V ? p.V - 1 : nxt);
      const float4* wp = reinterpret_cast<const float4*>(p.words + tk * kMegaD);
      const float4* pp = reinterpret_cast<const float4*>(p.positions + static_cast<long long>(pos + 1) * kMegaD);
      float4 v[6];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float4 a = __ldg(wp + i * 32 + lane);
        const float4 b2 = __ldg(pp + i * 32 + lane);
        v[i] = make_float4(a.x + b2.x, a.y + b2.y, a.z + b2.z, a.w + b2.w);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
      const float mean = warp_sum(s) * (1.0f / kMegaD);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float a = v[i].x - mean, b2 = v[i].y - mean, c2 = v[i].z - mean, d2 = v[i].w - mean;
        ss += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
      }
      const float rstd = rsqrtf(warp_sum(ss) * (1.0f / kMegaD) + 1e-8f);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float4 gmm = __ldg(reinterpret_cast<const float4*>(p.lnemb_g) + i * 32 + lane);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(p.lnemb_b) + i * 32 + lane);
        float4 ov = make_float4((v[i].x - mean) * rstd * gmm.x + bt.x, (v[i].y - mean) * rstd * gmm.y + bt.y,
                                (v[i].z - mean) * rstd * gmm.z + bt.z, (v[i].w - mean) * rstd * gmm.w + bt.w);
        reinterpret_cast<float4*>(p.x + static_cast<long long>(row) * kMegaD)[i * 32 + lane] = ov;
        uint2 pk;
        pk.x = pack_bf16(ov.x, ov.y);
        pk.y = pack_bf16(ov.z, ov.w);
        reinterpret_cast<uint2*>(p.hb + static_cast<long long>(row) * kMegaD)[i * 32 + lane] = pk;
      }
    }
    if (lane == 0) {
      // loop state: the row that draws the last ticket advances it (same protocol as greedy_select_kernel)
      if (nxt != p.eos) atomicAdd(&st->not_eos, 1);
      __threadfence();
      const unsigned int tr = atomicAdd(&st->ticket, 1u);
      if (tr == static_cast<unsigned int>(R) - 1) {
        __threadfence();
        const int not_eos = atomicAdd(&st->not_eos, 0);
        st->ticket = 0;
        st->not_eos = 0;
        st->cur_len = cur_len + 1;
        st->final_len = cur_len + 1;
        st->pos = pos + 1;
        st->step = step + 1;
        if (not_eos == 0) {
          st->finished = 1;
          if (step == 0 && p.row_prefix == nullptr) st->empty_caption = 1;
        }
        if (cur_len + 1 >= p.max_steps) st->finished = 1;
        __threadfence();
      }
    }
  }
}

constexpr size_t kMegaSmemBytes = 1024 + kMegaSlots * kMegaSlotBytes + kMegaComputeWarps * 2048 +
                                  kMegaAttItems * kMegaComputeWarps * kMegaAttState * 4 + kMegaAttItems * 128 + 2 * kMegaSlots * 8 + 64 + 64;

}  // namespace gitb200
