// Latency-lean GEMM for the decode step: out[r][f] (+)= sum_k X[r][k] * W[f][k] with R <= 64 activation rows.
//
// At 64 rows these GEMMs are pure latency chains (weights 1-5 MB, a few hundred KFLOP per CTA): what matters is the
// number of dependent hops between "predecessor finished" and "my result is visible".  The tcgen05 swap-AB path
// (gemm.cuh, still used for the 30522-wide LM head) needs TMA -> mbarrier -> UMMA -> commit -> tcgen05.ld -> smem
// transpose -> store, ~6 us in situ; this kernel needs cp.async -> ldmatrix/mma.sync -> one smem reduction -> store,
// and prefetches its whole weight slice before the dependency wait.
//
// CTA = 64 rows x 32 features x KS k-depth (grid = N/32 x K/KS), 8 warps.  Warp w owns k-slice w of the CTA's depth
// for ALL 64x32 outputs (each operand byte is read from shared memory exactly once per CTA), partial sums of the 8
// warps are reduced through shared memory, then bias / activation / store (fp32, bf16 or red.v4 accumulate).
#pragma once
#include "ptx.cuh"

namespace gitb200 {

struct SkinnyParams {
  const __nv_bfloat16* X;   // [R, K] activations
  long long ldx;
  const __nv_bfloat16* W;   // [N, K] weights (never written during decoding)
  long long ldw;
  int R, N, K, KS;          // KS: k-depth per CTA, multiple of 128
  const float* bias;        // [N] or null (added by the k-split 0 CTA only)
  int act;
  int mode;                 // 0: store fp32, 1: store bf16, 2: red.global.add.v4.f32 (split-K)
  void* out;
  long long ldo;
  const int* skip;
  int pdl;
  ChainSync chain;
};

constexpr int kSkinnyFT = 32;    // features per CTA
constexpr int kSkinnyRows = 64;

__host__ __device__ constexpr size_t skinny_smem_bytes(int KS) {
  const size_t operands = static_cast<size_t>(kSkinnyFT + kSkinnyRows) * KS * 2;
  const size_t reduce = static_cast<size_t>(8) * kSkinnyRows * 36 * 4;
  return (operands > reduce ? operands : reduce) + 128;
}

// element (row, k) of a [rows][KS] bf16 tile stored as KS/64 column blocks of [rows][128 B], 16 B units XOR-swizzled
__device__ __forceinline__ uint32_t skinny_off(int rows, int row, int kchunk16) {  // kchunk16: index of the 16 B unit
  return static_cast<uint32_t>(((kchunk16 >> 3) * rows + row) * 128 + (((kchunk16 & 7) ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(256, 1) skinny_mma_kernel(const SkinnyParams p) {
  extern __shared__ __align__(128) uint8_t sk_smem[];
  if (p.pdl) griddep_launch_early();
  if (p.pdl) tl_mark(100000 + 5000 + static_cast<int>(gridDim.x * gridDim.y));
  if ((!p.pdl || p.chain.counters != nullptr) && p.skip != nullptr && *p.skip != 0) return;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int f_base = blockIdx.x * kSkinnyFT;
  const int k_base = blockIdx.y * p.KS;
  const int KS = p.KS;
  const int units = KS >> 3;  // 16-byte units per row
  uint8_t* sW = sk_smem;
  uint8_t* sX = sk_smem + static_cast<size_t>(kSkinnyFT) * KS * 2;
  const uint32_t sW_u = smem_u32(sW), sX_u = smem_u32(sX);

  // ---- weights first (constant data): whole [32 x KS] slice in flight before the dependency wait ----
  for (int id = tid; id < kSkinnyFT * units; id += 256) {
    const int row = id / units, c = id - row * units;
    const int f = f_base + row;
    const bool ok = f < p.N && (k_base + c * 8) < p.K;
    const __nv_bfloat16* src = p.W + static_cast<long long>(ok ? f : 0) * p.ldw + (ok ? k_base + c * 8 : 0);
    cp_async_16(sW_u + skinny_off(kSkinnyFT, row, c), src, ok);
  }
  cp_async_commit();
  if (p.chain.counters != nullptr) chain_wait(p.chain);
  else if (p.pdl) griddep_wait();
  if (p.pdl) griddep_launch_late();
  if (p.pdl) tl_mark(5000 + static_cast<int>(gridDim.x * gridDim.y));
  // ---- activations (written by the predecessor: cp.async.cg reads through L2) ----
  for (int id = tid; id < kSkinnyRows * units; id += 256) {
    const int row = id / units, c = id - row * units;
    const bool ok = row < p.R && (k_base + c * 8) < p.K;
    const __nv_bfloat16* src = p.X + static_cast<long long>(ok ? row : 0) * p.ldx + (ok ? k_base + c * 8 : 0);
    cp_async_16(sX_u + skinny_off(kSkinnyRows, row, c), src, ok);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // ---- warp w: k-slice [w*KS/8, (w+1)*KS/8) for all 64 x 32 outputs ----
  float acc[4][4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
  const int steps = KS >> 7;            // 16-wide k-steps per warp
  const int u0 = warp * (KS >> 6);      // first 16 B unit of this warp's slice
  for (int ks = 0; ks < steps; ++ks) {
    const int u = u0 + ks * 2;
    uint32_t a[4][4], b[4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int row = m * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldmatrix_x4(a[m][0], a[m][1], a[m][2], a[m][3], sX_u + skinny_off(kSkinnyRows, row, u + (lane >> 4)));
    }
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      const int row = np * 16 + ((lane >> 4) & 1) * 8 + (lane & 7);
      ldmatrix_x4(b[2 * np][0], b[2 * np][1], b[2 * np + 1][0], b[2 * np + 1][1],
                  sW_u + skinny_off(kSkinnyFT, row, u + ((lane >> 3) & 1)));
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n) mma_bf16_16816(acc[m][n], a[m], b[n][0], b[n][1]);
  }
  __syncthreads();  // operands no longer needed: the reduction buffer overlays them
  float* red = reinterpret_cast<float*>(sk_smem);  // [8 warps][64 rows][36]
  {
    const int g = lane >> 2, t = lane & 3;
    float* base = red + static_cast<size_t>(warp) * kSkinnyRows * 36;
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        *reinterpret_cast<float2*>(base + (m * 16 + g) * 36 + n * 8 + 2 * t) = make_float2(acc[m][n][0], acc[m][n][1]);
        *reinterpret_cast<float2*>(base + (m * 16 + g + 8) * 36 + n * 8 + 2 * t) = make_float2(acc[m][n][2], acc[m][n][3]);
      }
  }
  __syncthreads();
  {
    const int row = tid >> 2;           // 0..63
    const int c0 = (tid & 3) * 8;       // 8 consecutive features
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float4 x0 = *reinterpret_cast<const float4*>(red + (static_cast<size_t>(w) * kSkinnyRows + row) * 36 + c0);
      const float4 x1 = *reinterpret_cast<const float4*>(red + (static_cast<size_t>(w) * kSkinnyRows + row) * 36 + c0 + 4);
      v[0] += x0.x; v[1] += x0.y; v[2] += x0.z; v[3] += x0.w;
      v[4] += x1.x; v[5] += x1.y; v[6] += x1.z; v[7] += x1.w;
    }
    const int f0 = f_base + c0;
    if (row < p.R && f0 < p.N) {        // N is a multiple of 8 for every decode GEMM (checked on the host)
      if (p.bias != nullptr && blockIdx.y == 0) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + f0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + f0 + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (p.act != ACT_NONE) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = apply_act(v[e], p.act);
      }
      const long long off = static_cast<long long>(row) * p.ldo + f0;
      if (p.mode == 1) {
        uint4 pk;
        pk.x = pack_bf16(v[0], v[1]); pk.y = pack_bf16(v[2], v[3]);
        pk.z = pack_bf16(v[4], v[5]); pk.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = pk;
      } else if (p.mode == 0) {
        float* dst = reinterpret_cast<float*>(p.out) + off;
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        float* dst = reinterpret_cast<float*>(p.out) + off;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
      }
    }
  }
  if (p.pdl) tl_mark(200000 + 5000 + static_cast<int>(gridDim.x * gridDim.y));
  chain_signal(p.chain);
}

}  // namespace gitb200
