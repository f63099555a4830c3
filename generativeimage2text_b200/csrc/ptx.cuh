// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA / TMEM),
// legacy mma.sync + ldmatrix (used by the attention kernels), and small math helpers.
// Bit layouts of the UMMA shared-memory and instruction descriptors follow the PTX ISA tables
// (cross-checked against cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS headers).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gitb200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Optional in-situ timeline (debug builds only: nvcc -DGITB200_TIMELINE, see build.py / tools/step_timeline2.py): block 0 /
// thread 0 of every decode-step kernel appends (%globaltimer, kernel id) at its start, right after its dependency wait
// (i.e. when its predecessor has fully completed) and at its end.  The production library contains none of this.
#ifdef GITB200_TIMELINE
__device__ unsigned long long* g_tl_buf = nullptr;
__device__ unsigned int g_tl_count = 0;
constexpr unsigned int kTimelineMax = 8192;
__device__ __forceinline__ void tl_mark_one(int kid) {   // one designated thread of block 0
  if (g_tl_buf != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    const unsigned int i = atomicAdd(&g_tl_count, 1u);
    if (i < kTimelineMax) {
      g_tl_buf[2 * i] = globaltimer_ns();
      g_tl_buf[2 * i + 1] = static_cast<unsigned long long>(kid);
    }
  }
}
__device__ __forceinline__ void tl_mark(int kid) {
  if (threadIdx.x == 0) tl_mark_one(kid);
}
#else
__device__ __forceinline__ void tl_mark(int) {}
__device__ __forceinline__ void tl_mark_one(int) {}
#endif

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor is still running; griddep_wait() blocks until the predecessor grid has completed
// and its writes are visible, griddep_launch() lets the successor's prologue begin. Both are no-ops otherwise.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// The decode-chain kernels release their PDL successor at kernel entry: its prologue and weight prefetch overlap as much
// as possible (a late release, once the kernel's own dependency is satisfied, was measured in round 1: no gain).
__device__ __forceinline__ void griddep_launch_early() { griddep_launch(); }

// Flag-based ordering of the decode-step kernel chain.  Kernel k of a step (launched with the PDL attribute, so
// it may become resident while kernel k-1 still runs, but WITHOUT griddepcontrol.wait) spins until every CTA of
// kernel k-1 has published its completion; kernel boundaries then cost one L2 round trip instead of a full grid
// drain + memory flush (~4 us measured).  Safe against deadlock because a PDL successor is only scheduled once
// every CTA of its predecessor has started.  Data written under this scheme must be read with L1-bypassing
// loads (__ldcg / TMA).  Counters are re-zeroed by the last kernel of the step.
struct ChainSync {
  unsigned int* counters;   // [64] (null: disabled)
  int idx;                  // position of this kernel in the step's chain
  unsigned int pred_ctas;   // CTAs of kernel idx-1
};
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// all threads of the CTA
__device__ __forceinline__ void chain_wait(const ChainSync& c) {
  if (c.counters != nullptr && c.idx > 0) {
    if (threadIdx.x == 0) {
      while (ld_acquire_gpu(c.counters + c.idx - 1) < c.pred_ctas) {
      }
      asm volatile("fence.proxy.async;" ::: "memory");  // thread 0 is also the TMA issuer of every kernel here
    }
    __syncthreads();
  }
}
// thread 0, after a __syncthreads() that follows the CTA's last global write
__device__ __forceinline__ void chain_signal_thread0(const ChainSync& c) {
  // release-RMW: orders this thread's (and, through the preceding bar.sync, the CTA's) prior writes before the count
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c.counters + c.idx), "r"(1u) : "memory");
}
__device__ __forceinline__ void chain_signal(const ChainSync& c) {
  if (c.counters != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) chain_signal_thread0(c);
  }
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------------------------------------
// TMA: 2-D tiled bulk tensor load, completion counted on an mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c_inner,
                                            int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner),
      "r"(c_outer)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, UMMA issue, commit, TMEM loads
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// Arrive on an mbarrier once all tcgen05.mma previously issued by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, single CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 bytes with the
// 128-byte swizzle (what TMA's CU_TENSOR_MAP_SWIZZLE_128B writes): 8-row core groups are 1024 B
// apart (SBO), LBO is unused for swizzled K-major layouts (set to 1 like CUTLASS), version = 1.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);     // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                        // [16,30) leading byte offset >> 4
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                // [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                        // [46,48) descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                        // [61,64) SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, dense, M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (lane i of the warp reads
// TMEM lane (warp%4)*32+i).  Caller must tmem_ld_wait() before using r[].
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// legacy warp-level tensor-core path (attention core: 4 % of the encoder FLOPs)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                                                  uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------------
// math / packing
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

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

enum Act : int { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_GELU_ERF = 2, ACT_QUICKGELU_EXACT = 3 };

// fp32 -> (hi, lo) bf16 pair with hi + lo == x to ~2^-17 relative: the operand format of the engine's fp32-grade parity
// mode, where every GEMM runs as sum_k (a_hi w_hi + a_lo w_hi + a_hi w_lo) through the same tcgen05 kernel by laying the
// three products side by side along K ("[hi | lo | hi]" activations against "[hi | hi | lo]" weights).
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ void pack_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat16 ah, al, bh, bl;
  split_bf16(a, ah, al);
  split_bf16(b, bh, bl);
  hi = static_cast<uint32_t>(__bfloat16_as_ushort(ah)) | (static_cast<uint32_t>(__bfloat16_as_ushort(bh)) << 16);
  lo = static_cast<uint32_t>(__bfloat16_as_ushort(al)) | (static_cast<uint32_t>(__bfloat16_as_ushort(bl)) << 16);
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_QUICKGELU) {
    // reference layers/CLIP/model.py:171-173: x * sigmoid(1.702 x) == 0.5 x (1 + tanh(0.851 x)).
    // One MUFU op per element (tanh.approx, rel. error 2^-11 -- below the bf16 rounding of the stored result):
    // the ex2 + rcp form made the c_fc epilogue MUFU-bound (16 ops/clk/SM).
    const float hx = 0.5f * x;
    return fmaf(hx, tanh_approx(0.851f * x), hx);
  } else if (act == ACT_QUICKGELU_EXACT) {   // parity mode: the sigmoid itself (expf, full-precision division)
    return x / (1.0f + expf(-1.702f * x));
  } else if (act == ACT_GELU_ERF) {
    // reference layers/bert/activations.py:16-23: x * 0.5 * (1 + erf(x / sqrt(2))).
    // erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32-exact for our purposes) with one ex2 and
    // one rcp instead of erff()'s ~40-instruction polynomial: the decode-step fc1 epilogue was bound by it.
    const float z = x * 0.70710678118654752f;
    const float az = fabsf(z);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    const float erf_abs = 1.0f - poly * __expf(-az * az);
    return x * 0.5f * (1.0f + copysignf(erf_abs, z));
  }
  return x;
}

}  // namespace gitb200
